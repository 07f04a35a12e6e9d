"""NDArray / IndexedSlices — mirrors python/hetu/ndarray.py (every tensor is float32, ids
included: ndarray.py:401-418).  CPU arrays live in pinned host memory so copies are async."""
import ctypes

import numpy as np

from ._base import _LIB, check_call, c_array


class DLContext(ctypes.Structure):
    _fields_ = [("device_id", ctypes.c_int), ("device_type", ctypes.c_int)]
    _NAMES = {1: "cpu", 2: "gpu"}

    def __init__(self, device_id=0, device_type=1):
        super().__init__()
        self.device_id = device_id
        self.device_type = device_type

    def __repr__(self):
        return "%s(%d)" % (self._NAMES.get(self.device_type, "?"), self.device_id)

    def __hash__(self):
        return hash((self.device_type, self.device_id))

    def __eq__(self, other):
        return isinstance(other, DLContext) and hash(self) == hash(other)

    def __ne__(self, other):
        return not self == other


class DLArray(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("ctx", DLContext), ("ndim", ctypes.c_int),
                ("shape", ctypes.POINTER(ctypes.c_int64)),
                ("stride", ctypes.POINTER(ctypes.c_int64))]


DLArrayHandle = ctypes.POINTER(DLArray)


def cpu(dev_id=0):
    return DLContext(dev_id, 1)


def gpu(dev_id=0):
    return DLContext(dev_id, 2)


def is_gpu_ctx(ctx):
    return bool(ctx) and ctx.device_type == 2


def shape_to_stride(shape):
    stride, run = [], 1
    for dim in reversed(tuple(shape)):
        stride.append(run)
        run *= dim
    return tuple(reversed(stride))


def _view_of_numpy(arr):
    """A DLArray struct aliasing a C-contiguous float32 numpy array (no copy)."""
    assert arr.flags["C_CONTIGUOUS"] and arr.dtype == np.float32
    view = DLArray()
    keep = (c_array(ctypes.c_int64, arr.shape), c_array(ctypes.c_int64, shape_to_stride(arr.shape)))
    view.data = arr.ctypes.data_as(ctypes.c_void_p)
    view.shape, view.stride = keep
    view.ndim = arr.ndim
    view.ctx = cpu(0)
    return view, keep


class NDArray(object):
    """Buffer object over a DLArray handle; no arithmetic."""
    __slots__ = ["handle", "no_free"]

    def __init__(self, handle):
        self.handle = handle
        self.no_free = False

    def __del__(self):
        if not self.no_free and self.handle and _LIB is not None:
            _LIB.DLArrayFree(self.handle)

    @property
    def shape(self):
        c = self.handle.contents
        return tuple(c.shape[i] for i in range(c.ndim))

    @property
    def stride(self):
        c = self.handle.contents
        return tuple(c.stride[i] for i in range(c.ndim))

    @property
    def ctx(self):
        return self.handle.contents.ctx

    @property
    def data_ptr(self):
        return self.handle.contents.data

    def __setitem__(self, in_slice, value):
        if not isinstance(in_slice, slice) or in_slice.start is not None or in_slice.stop is not None:
            raise ValueError("Array only support set from numpy array")
        if isinstance(value, NDArray):
            if value.handle is not self.handle:
                value.copyto(self)
        elif isinstance(value, (np.ndarray, np.generic)):
            self._sync_copyfrom(value)
        else:
            raise TypeError("type %s not supported" % str(type(value)))

    def _sync_copyfrom(self, source, data_type=np.float32):
        source = np.ascontiguousarray(source, dtype=np.float32)
        if source.shape != self.shape:
            raise ValueError("array shape do not match the shape of NDArray")
        view, keep = _view_of_numpy(source)
        check_call(_LIB.DLArrayCopyFromTo(ctypes.byref(view), self.handle, None))
        del keep

    def _async_copyfrom(self, source_array, stream_handle, event_handle=None):
        check_call(_LIB.DLArrayCopyFromTo(source_array.handle, self.handle, stream_handle.handle))
        if event_handle is not None:
            check_call(_LIB.DLEventRecord(stream_handle.handle, event_handle.handle))

    def async_h2d(self, source_array, stream_handle, event_handle=None):
        if isinstance(source_array, tuple):
            source_array = source_array[0]
        if isinstance(source_array, np.ndarray):
            source_array = array(source_array, cpu(0))
        assert is_gpu_ctx(self.ctx) and not is_gpu_ctx(source_array.ctx) and stream_handle
        self._async_copyfrom(source_array, stream_handle, event_handle)

    def async_d2h(self, source_array, stream_handle, event_handle=None):
        assert not is_gpu_ctx(self.ctx) and is_gpu_ctx(source_array.ctx) and stream_handle
        self._async_copyfrom(source_array, stream_handle, event_handle)

    def asnumpy(self):
        out = np.empty(self.shape, dtype=np.float32)
        view, keep = _view_of_numpy(out)
        check_call(_LIB.DLArrayCopyFromTo(self.handle, ctypes.byref(view), None))
        del keep
        return out

    def host_view(self):
        """numpy view (no copy) of a HOST NDArray's pinned memory — herald_b200 extension, used to
        read a result where it landed instead of copying it again."""
        assert not is_gpu_ctx(self.ctx)
        n = int(np.prod(self.shape))
        buf = (ctypes.c_float * n).from_address(self.data_ptr)
        return np.frombuffer(buf, dtype=np.float32).reshape(self.shape)

    def copyto(self, target):
        if isinstance(target, DLContext):
            target = empty(self.shape, target)
        if not isinstance(target, NDArray):
            raise ValueError("Unsupported target type %s" % str(type(target)))
        check_call(_LIB.DLArrayCopyFromTo(self.handle, target.handle, None))
        return target

    def reshape(self, shape, target):
        """Alias this buffer under another shape (target does not own the memory)."""
        c = self.handle.contents
        arr = DLArray()
        arr.data, arr.ctx, arr.ndim = c.data, c.ctx, len(shape)
        arr.shape = c_array(ctypes.c_int64, tuple(shape))
        arr.stride = c_array(ctypes.c_int64, shape_to_stride(shape))
        target.handle = ctypes.pointer(arr)
        target.no_free = True


def empty(shape, ctx=cpu(0)):
    shape = tuple(int(s) for s in shape)
    cshape = c_array(ctypes.c_int64, shape)
    cstride = c_array(ctypes.c_int64, shape_to_stride(shape))
    handle = DLArrayHandle()
    check_call(_LIB.DLArrayAlloc(cshape, cstride, ctypes.c_int64(len(shape)), ctx,
                                 ctypes.byref(handle)))
    return NDArray(handle)


def array(arr, ctx, data_type=np.float32):
    if not isinstance(arr, np.ndarray):
        arr = np.array(arr, dtype=data_type)
    out = empty(arr.shape, ctx)
    out._sync_copyfrom(arr)
    return out


class IndexedSlices(object):
    """Sparse gradient container: python/hetu/ndarray.py:503-611.

    ``deduplicate`` is where the reference copies the ids to the host, runs ``np.unique`` and
    copies unique ids + inverse back (ndarray.py:532-554).  Here the same ascending-unique
    contract is met on the device (HBUniqueIndexedSlices: radix sort + scan); only the unique
    COUNT crosses to the host because the output NDArrays are exactly ``[U]`` / ``[U, D]``.
    """
    __slots__ = ["indices", "values", "dense_shape", "deduplicated", "lazy", "to_dense_flag",
                 "push_indices"]

    def __init__(self, indices=None, values=None, dense_shape=None, push_indices=None):
        self.indices = indices
        self.push_indices = push_indices
        self.values = values
        self.dense_shape = dense_shape
        self.deduplicated = False
        self.lazy = False
        self.to_dense_flag = False

    def get_dense_shape(self):
        assert self.dense_shape is not None
        return self.dense_shape

    def get_sparse_shape(self):
        assert isinstance(self.values, NDArray)
        return self.values.shape

    def update(self, indices, values, dense_shape, push_indices=None):
        self.indices = indices
        self.push_indices = push_indices
        self.values = values
        if self.dense_shape is not None:
            assert tuple(self.dense_shape) == tuple(dense_shape)
        else:
            self.dense_shape = dense_shape

    @staticmethod
    def _unique_on_device(ids, stream):
        """-> (unique NDArray[U], inverse NDArray[n], U)."""
        n = int(np.prod(ids.shape)) if ids.shape else 1
        ctx = ids.ctx
        uniq_full = empty((n,), ctx)
        inverse = empty((n,), ctx)
        count = empty((2,), ctx)  # 8 bytes: one int64 written by the kernel
        sh = stream.handle if stream else None
        check_call(_LIB.HBUniqueIndexedSlices(ids.handle, uniq_full.handle, inverse.handle,
                                              ctypes.c_void_p(count.data_ptr), sh))
        if stream:
            stream.sync()
        num_unique = int(count.asnumpy().view(np.int64)[0])
        uniq = empty((num_unique,), ctx)
        # device-to-device copy of the first U entries
        head = NDArray(None)
        uniq_full.reshape((n,), head)
        head.handle.contents.shape[0] = num_unique
        check_call(_LIB.DLArrayCopyFromTo(head.handle, uniq.handle, sh))
        if stream:
            stream.sync()
        return uniq, inverse, num_unique

    def deduplicate(self, stream):
        assert is_gpu_ctx(self.indices.ctx)
        uniq, inverse, num_unique = self._unique_on_device(self.indices, stream)
        self.indices = uniq
        if self.push_indices is not None:
            self.push_indices, _, _ = self._unique_on_device(self.push_indices, stream)
        new_values = empty((num_unique, self.values.shape[-1]), ctx=self.values.ctx)
        sh = stream.handle if stream else None
        check_call(_LIB.DLGpuArraySet(new_values.handle, ctypes.c_float(0), sh))
        check_call(_LIB.DeduplicateIndexedSlices(self.values.handle, inverse.handle,
                                                 new_values.handle, sh))
        self.values = new_values
        self.deduplicated = True

    def free_deduplicate(self):
        if self.deduplicated:
            self.indices = None
            self.push_indices = None
            self.values = None
            self.deduplicated = False

    def to_dense(self, stream):
        assert is_gpu_ctx(self.indices.ctx)
        dense_shape = tuple(self.get_dense_shape())
        new_values = empty(dense_shape, ctx=self.values.ctx)
        sh = stream.handle if stream else None
        check_call(_LIB.DLGpuArraySet(new_values.handle, ctypes.c_float(0), sh))
        check_call(_LIB.IndexedSlices2Dense(self.values.handle, self.indices.handle,
                                            new_values.handle, sh))
        all_rows = array(np.arange(dense_shape[0], dtype=np.float32), ctx=self.indices.ctx)
        self.free_deduplicate()
        self.values = new_values
        self.indices = all_rows
        self.to_dense_flag = True

    def free_dense(self):
        if self.to_dense_flag:
            self.indices = None
            self.push_indices = None
            self.values = None
            self.to_dense_flag = False
