"""Streams and events — mirrors python/hetu/stream.py."""
import ctypes

from ._base import _LIB, check_call
from . import ndarray


class DLStream(ctypes.Structure):
    _fields_ = [("device_id", ctypes.c_int), ("handle", ctypes.c_void_p)]


DLStreamHandle = ctypes.POINTER(DLStream)


class Stream(object):
    __slots__ = ["handle"]

    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        if self.handle and _LIB is not None:
            _LIB.DLStreamDestroy(self.handle)

    def sync(self):
        check_call(_LIB.DLStreamSync(self.handle))


def create_stream_handle(ctx):
    assert ndarray.is_gpu_ctx(ctx)
    handle = DLStreamHandle()
    check_call(_LIB.DLStreamCreate(ctypes.c_size_t(ctx.device_id), ctypes.byref(handle)))
    return Stream(handle)


class DLEvent(ctypes.Structure):
    _fields_ = [("device_id", ctypes.c_int), ("handle", ctypes.c_void_p)]


DLEventHandle = ctypes.POINTER(DLEvent)


class Event(object):
    __slots__ = ["handle"]

    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        if self.handle and _LIB is not None:
            _LIB.DLEventDestroy(self.handle)

    def sync(self):
        check_call(_LIB.DLEventSync(self.handle))

    def record(self, stream_handle):
        check_call(_LIB.DLEventRecord(stream_handle.handle, self.handle))

    def time_since(self, event):
        out = ctypes.c_float()
        check_call(_LIB.DLEventElapsedTime(event.handle, self.handle, ctypes.byref(out)))
        return out.value


def create_event_handle(ctx):
    assert ndarray.is_gpu_ctx(ctx)
    handle = DLEventHandle()
    check_call(_LIB.DLEventCreate(ctypes.c_size_t(ctx.device_id), ctypes.byref(handle)))
    return Event(handle)


class CSEvent(object):
    """Event of a cache-backed embedding parameter: a list of pending cache calls
    (python/hetu/stream.py:86-105)."""
    __slots__ = ["tss"]

    def __init__(self, comm=None, nid=None):
        self.tss = []

    def update_ts(self, ts):
        self.tss.append(ts)

    def update(self):
        pass

    def sync(self):
        for ts in self.tss:
            ts.wait()
        self.tss = []
