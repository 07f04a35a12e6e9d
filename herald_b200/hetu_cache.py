"""`hetu_cache` — the reference's pybind module surface (src/hetu_cache/src/python_api.cc:10-79)
on top of the hb_cache_* C-ABI.  Import it as ``from herald_b200 import hetu_cache`` where the
reference does ``import hetu_cache`` (python/hetu/cstable.py:23-24).

The cache, its index and the rows live in HBM.  Every batch method accepts host memory (the
reference's contract: numpy arrays or raw host pointers) or device pointers, returns immediately
with a ``_waittype`` and completes on ``.wait()``; callers keep keys/dest/grads alive until
then, exactly as the reference requires (cstable.py:38-45).
"""
import ctypes

import numpy as np

from ._base import _LIB, check_call

_sz = ctypes.c_size_t
_vp = ctypes.c_void_p

KEYS_U64, KEYS_F32 = 0, 1
_POLICY = {"lru": 0, "lfu": 1, "lfuopt": 2}


class hb_perf(ctypes.Structure):
    _fields_ = [("num_all", ctypes.c_int64), ("num_unique", ctypes.c_int64),
                ("num_miss", ctypes.c_int64), ("num_evict", ctypes.c_int64),
                ("num_transfered", ctypes.c_int64), ("is_full", ctypes.c_int64),
                ("size", ctypes.c_int64), ("error", ctypes.c_int64),
                ("time_ms", ctypes.c_float), ("sort_ms", ctypes.c_float),
                ("lookup_ms", ctypes.c_float), ("transfer_ms", ctypes.c_float),
                ("copy_ms", ctypes.c_float), ("insert_ms", ctypes.c_float),
                ("kernel_ms", ctypes.c_float), ("num_remote", ctypes.c_int64)]


def _proto():
    L = _LIB
    L.hb_cache_create.argtypes = [ctypes.c_int, _sz, _sz, _sz, ctypes.c_int, ctypes.POINTER(_vp)]
    L.hb_cache_destroy.argtypes = [_vp]
    L.hb_cache_set_bounds.argtypes = [_vp, ctypes.c_int64, ctypes.c_int64]
    L.hb_cache_get_bounds.argtypes = [_vp, ctypes.POINTER(ctypes.c_int64),
                                      ctypes.POINTER(ctypes.c_int64)]
    L.hb_cache_set_bypass.argtypes = [_vp, ctypes.c_int]
    L.hb_cache_set_grad_scale.argtypes = [_vp, ctypes.c_float]
    L.hb_cache_get_grad_scale.argtypes = [_vp, ctypes.POINTER(ctypes.c_float)]
    L.hb_cache_after_stream.argtypes = [_vp, _vp]
    L.hb_cache_set_reduce_mode.argtypes = [_vp, ctypes.c_int]
    L.hb_cache_get_reduce_mode.argtypes = [_vp, ctypes.POINTER(ctypes.c_int)]
    L.hb_cache_set_perf.argtypes = [_vp, ctypes.c_int]
    L.hb_cache_set_perf_sampling.argtypes = [_vp, ctypes.c_uint]
    L.hb_cache_reserve.argtypes = [_vp, _sz]
    L.hb_cache_stream.argtypes = [_vp, ctypes.POINTER(_vp)]
    L.hb_cache_flush.argtypes = [_vp]
    L.hb_cache_lookup.argtypes = [_vp, _vp, ctypes.c_int, _sz, _vp]
    L.hb_cache_update.argtypes = [_vp, _vp, ctypes.c_int, _sz, _vp]
    L.hb_cache_update_with_push_keys.argtypes = [_vp, _vp, ctypes.c_int, _sz, _vp, ctypes.c_int,
                                                 _sz, _vp]
    L.hb_cache_push_pull.argtypes = [_vp, _vp, ctypes.c_int, _sz, _vp, _vp, ctypes.c_int, _sz, _vp]
    L.hb_cache_wait.argtypes = [_vp, ctypes.POINTER(hb_perf)]
    L.hb_cache_last_call.argtypes = [_vp, ctypes.POINTER(ctypes.c_uint64)]
    L.hb_cache_wait_call.argtypes = [_vp, ctypes.c_uint64, ctypes.POINTER(hb_perf)]
    L.hb_cache_perf_range.argtypes = [_vp, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(hb_perf),
                                      ctypes.POINTER(ctypes.c_int)]
    L.hb_cache_perf_history.argtypes = [_vp, ctypes.POINTER(hb_perf), ctypes.POINTER(ctypes.c_int),
                                        ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    L.hb_cache_score.argtypes = [_vp, _vp, ctypes.c_int, _sz, _sz, _vp, _sz, ctypes.c_int, _vp]
    L.hb_cache_probe.argtypes = [_vp, _vp, ctypes.c_int, _sz, _vp]
    L.hb_cache_size.argtypes = [_vp, ctypes.POINTER(_sz)]
    L.hb_cache_count.argtypes = [_vp, ctypes.c_uint64, ctypes.POINTER(ctypes.c_int)]
    L.hb_cache_keys.argtypes = [_vp, _vp, _sz, ctypes.POINTER(_sz)]
    L.hb_cache_peek.argtypes = [_vp, ctypes.c_uint64, ctypes.POINTER(ctypes.c_int),
                                ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
                                _vp, _vp]
    L.hb_cache_touch.argtypes = [_vp, ctypes.c_uint64, ctypes.POINTER(ctypes.c_int),
                                 ctypes.POINTER(ctypes.c_int64), _vp]
    L.hb_cache_insert.argtypes = [_vp, ctypes.c_uint64, ctypes.c_int64, _vp]


_proto()


def _check_c_contiguous(arr, what):
    # binding.h:51-57 raises std::runtime_error("Array not continuous in C: ...")
    if not arr.flags["C_CONTIGUOUS"]:
        raise RuntimeError("Array not continuous in C: " + what)


class _waittype(object):
    """Handle of ONE enqueued cache call (python_api.cc:16-19): ``wait()`` blocks until that call
    (for a lookup into host memory: including the download of its rows) has finished and records
    the perf entries up to it.  Calls enqueued after it keep running, so a host caller can overlap
    update(t+1)'s upload with lookup(t+1)'s download."""

    def __init__(self, cache, keepalive, seq):
        self._cache = cache
        self._keep = keepalive
        self._seq = seq

    def wait(self):
        self._cache._drain(self._seq)
        self._keep = None


class Embedding(object):
    """One cache line as seen from Python (python_api.cc:21-30)."""

    def __init__(self, key, version, data, grad=None, updates=0):
        data = np.ascontiguousarray(data, dtype=np.float32)
        assert data.ndim == 1
        self.key = int(key)
        self.version = int(version)
        self.data = data
        self.grad = np.zeros_like(data) if grad is None else grad
        self.updates = int(updates)

    def mean(self):
        return float(np.sum(self.data.astype(np.float64)) / self.data.size)

    def var(self):
        d = self.data.astype(np.float64)
        return float(np.sum((d - d.mean()) ** 2) / d.size)

    def __repr__(self):
        return "<hetu.Embedding : key:%d, len:%d, version:%d, mean:%g, var:%g>" % (
            self.key, self.data.size, self.version, self.mean(), self.var())


class CacheBase(object):
    _policy = None

    def __init__(self, limit, length, width, node_id):
        self._h = _vp()
        self._limit, self._length, self._width, self._node_id = (int(limit), int(length),
                                                                 int(width), int(node_id))
        check_call(_LIB.hb_cache_create(_POLICY[self._policy], self._limit, self._length,
                                        self._width, self._node_id, ctypes.byref(self._h)))
        self._perf = []
        self._perf_enabled = False
        self._recorded = 0  # device calls [0, _recorded) have had their perf entry collected

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _LIB is not None:
            _LIB.hb_cache_destroy(h)
            self._h = None

    # ---- properties (python_api.cc:32-41) ------------------------------------------------
    @property
    def limit(self):
        return self._limit

    @property
    def width(self):
        return self._width

    @property
    def perf(self):
        self._drain()
        return self._perf

    @property
    def perf_enabled(self):
        return self._perf_enabled

    def set_perf_sampling(self, every):
        """Phase timings on every `every`-th update/lookup pair only (herald_b200 extension)."""
        check_call(_LIB.hb_cache_set_perf_sampling(self._h, int(every)))

    @perf_enabled.setter
    def perf_enabled(self, value):
        self._drain()
        self._perf_enabled = bool(value)
        check_call(_LIB.hb_cache_set_perf(self._h, int(self._perf_enabled)))

    def _bounds(self):
        a, b = ctypes.c_int64(), ctypes.c_int64()
        check_call(_LIB.hb_cache_get_bounds(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    @property
    def pull_bound(self):
        return self._bounds()[0]

    @pull_bound.setter
    def pull_bound(self, v):
        check_call(_LIB.hb_cache_set_bounds(self._h, int(v), self._bounds()[1]))

    @property
    def push_bound(self):
        return self._bounds()[1]

    @push_bound.setter
    def push_bound(self, v):
        check_call(_LIB.hb_cache_set_bounds(self._h, self._bounds()[0], int(v)))

    @property
    def grad_scale(self):
        """Factor every gradient value is multiplied by inside the update kernel (herald_b200
        extension: the -lr fold the reference does on the host, ParameterServerCommunicate.py:58-59)."""
        v = ctypes.c_float()
        check_call(_LIB.hb_cache_get_grad_scale(self._h, ctypes.byref(v)))
        return v.value

    @grad_scale.setter
    def grad_scale(self, v):
        check_call(_LIB.hb_cache_set_grad_scale(self._h, float(v)))

    @property
    def reduce_mode(self):
        """"exact": every row's occurrences are added in order, bit-identical to the reference;
        "split": rows with more than 1024 occurrences in a call use a fixed two-level order
        (deterministic, within 1e-5 relative of the reference) — herald_b200 extension."""
        m = ctypes.c_int()
        check_call(_LIB.hb_cache_get_reduce_mode(self._h, ctypes.byref(m)))
        return "split" if m.value else "exact"

    @reduce_mode.setter
    def reduce_mode(self, mode):
        check_call(_LIB.hb_cache_set_reduce_mode(self._h, {"exact": 0, "split": 1}[mode]))

    def after(self, stream_handle):
        """Order the cache's work behind a caller's DLStream (device-pointer callers whose keys /
        gradients are still being produced on that stream)."""
        if stream_handle is None:
            return
        raw = ctypes.cast(stream_handle.handle.contents.handle, ctypes.POINTER(_vp)).contents
        check_call(_LIB.hb_cache_after_stream(self._h, raw))

    def bypass(self):
        check_call(_LIB.hb_cache_set_bypass(self._h, 1))

    def undo_bypass(self):
        check_call(_LIB.hb_cache_set_bypass(self._h, 0))

    def reserve(self, max_keys):
        check_call(_LIB.hb_cache_reserve(self._h, int(max_keys)))

    def flush(self):
        """Push every dirty line to its owner whatever the bound (pending victims and resident
        lines) — not in the reference, whose ParamSave loses the updates still held in worker
        caches.  Call before SaveParam.  Synchronous."""
        self._drain()
        first = self._last_call() + 1
        check_call(_LIB.hb_cache_flush(self._h))
        self._drain(skip_from=first)   # the flush runs one (empty) update call on the device

    @property
    def stream(self):
        """cudaStream_t (as int) that orders this cache's work — for device-pointer callers."""
        s = _vp()
        check_call(_LIB.hb_cache_stream(self._h, ctypes.byref(s)))
        return s.value

    # ---- completion / perf -------------------------------------------------------------
    def _last_call(self):
        """Sequence number of the device call enqueued last, -1 if there is none."""
        seq = ctypes.c_uint64()
        if _LIB.hb_cache_last_call(self._h, ctypes.byref(seq)) != 0:
            return -1
        return int(seq.value)

    def _drain(self, upto=None, skip_from=None):
        """Wait for device call `upto` (default: everything enqueued) and collect the perf entries
        of the calls finished by then (cache.cc:89-106,179-196).  Calls from `skip_from` on are
        internal (flush) and leave no entry."""
        if upto is None:
            check_call(_LIB.hb_cache_wait(self._h, None))
            upto = self._last_call()
        else:
            check_call(_LIB.hb_cache_wait_call(self._h, int(upto), None))
        pending = upto + 1 - self._recorded
        if pending <= 0:
            return
        if self._perf_enabled:
            first = self._recorded
            if pending > 1024:          # hb_cache::kRing records are kept
                first, pending = upto + 1 - 1024, 1024
            buf = (hb_perf * pending)()
            kinds = (ctypes.c_int * pending)()
            check_call(_LIB.hb_cache_perf_range(self._h, first, pending, buf, kinds))
            for k in range(pending):
                p = buf[k]
                if kinds[k] == 2:   # push_pull records nothing in the reference
                    continue
                if skip_from is not None and first + k >= skip_from:
                    continue
                d = {"type": "Pull" if kinds[k] == 0 else "Push", "is_full": bool(p.is_full),
                     "num_all": int(p.num_all), "num_unique": int(p.num_unique),
                     "num_miss": int(p.num_miss), "num_transfered": int(p.num_transfered),
                     "time": float(p.time_ms), "sort_time": float(p.sort_ms),
                     "lookup_time": float(p.lookup_ms), "transfer_time": float(p.transfer_ms),
                     "copy_time": float(p.copy_ms), "num_remote": int(p.num_remote)}
                if kinds[k] == 0:
                    d["prepare_time"] = 0.0
                    d["insert_time"] = float(p.insert_ms)
                else:
                    d["num_evict"] = int(p.num_evict)
                    d["cleanup_time"] = 0.0
                    d["kernel_time"] = float(p.kernel_ms)   # accumulate kernel alone (extension)
                self._perf.append(d)
        self._recorded = upto + 1

    def _issue(self, *keepalive):
        seq = self._last_call()
        # the library keeps the records of the last 1024 calls: with perf on, collect the old ones
        # before they are overwritten (they belong to calls that finished long ago)
        if self._perf_enabled and seq - self._recorded > 768:
            self._drain(seq - 256)
        return _waittype(self, keepalive, seq)

    # ---- numpy entry points (uint64 keys) ------------------------------------------------
    def embedding_lookup(self, keys, dest):
        _check_c_contiguous(keys, "keys")
        _check_c_contiguous(dest, "dest")
        assert keys.dtype == np.uint64 and dest.dtype == np.float32
        assert dest.size == keys.size * self._width
        check_call(_LIB.hb_cache_lookup(self._h, keys.ctypes.data, KEYS_U64, keys.size,
                                        dest.ctypes.data))
        return self._issue(keys, dest)

    def embedding_update(self, keys, grads):
        _check_c_contiguous(keys, "keys")
        _check_c_contiguous(grads, "grads")
        assert keys.dtype == np.uint64 and grads.dtype == np.float32
        assert grads.size == keys.size * self._width
        check_call(_LIB.hb_cache_update(self._h, keys.ctypes.data, KEYS_U64, keys.size,
                                        grads.ctypes.data))
        return self._issue(keys, grads)

    def embedding_update_with_push_keys(self, keys, push_keys, grads):
        for a, w in ((keys, "keys"), (push_keys, "push_keys"), (grads, "grads")):
            _check_c_contiguous(a, w)
        assert keys.dtype == np.uint64 and push_keys.dtype == np.uint64
        assert grads.dtype == np.float32 and grads.size == keys.size * self._width
        check_call(_LIB.hb_cache_update_with_push_keys(
            self._h, keys.ctypes.data, KEYS_U64, keys.size, push_keys.ctypes.data, KEYS_U64,
            push_keys.size, grads.ctypes.data))
        return self._issue(keys, push_keys, grads)

    # ---- raw-pointer entry points (float32-carried ids; host OR device pointers) ----------
    def embedding_lookup_raw(self, keys_ptr, dest_ptr, num_keys):
        check_call(_LIB.hb_cache_lookup(self._h, keys_ptr, KEYS_F32, int(num_keys), dest_ptr))
        return self._issue()

    def embedding_update_raw(self, keys_ptr, grads_ptr, num_keys):
        check_call(_LIB.hb_cache_update(self._h, keys_ptr, KEYS_F32, int(num_keys), grads_ptr))
        return self._issue()

    def embedding_push_pull_raw(self, pullkeys_ptr, dest_ptr, num_pull_keys, pushkeys_ptr,
                                grads_ptr, num_push_keys):
        check_call(_LIB.hb_cache_push_pull(self._h, pullkeys_ptr, KEYS_F32, int(num_pull_keys),
                                           dest_ptr, pushkeys_ptr, KEYS_F32, int(num_push_keys),
                                           grads_ptr))
        return self._issue()

    def embedding_update_with_push_keys_np_raw(self, keys_ptr, push_keys, grads_ptr, num_keys):
        _check_c_contiguous(push_keys, "push_keys")
        assert push_keys.dtype == np.uint64
        check_call(_LIB.hb_cache_update_with_push_keys(
            self._h, keys_ptr, KEYS_F32, int(num_keys), push_keys.ctypes.data, KEYS_U64,
            push_keys.size, grads_ptr))
        return self._issue(push_keys)

    def embedding_update_with_push_keys_raw(self, keys_ptr, push_keys_ptr, grads_ptr, num_keys,
                                            num_push_keys):
        check_call(_LIB.hb_cache_update_with_push_keys(
            self._h, keys_ptr, KEYS_F32, int(num_keys), push_keys_ptr, KEYS_F32,
            int(num_push_keys), grads_ptr))
        return self._issue()

    # ---- Laia / Herald scoring against the real index (herald_b200 extension) --
    def score_samples(self, sample_ids, table_order=None, top_k=None, fresh=False):
        """uint32[num_samples]: per sample row of `sample_ids` ([S, T], uint64 or float32-carried),
        how many of its first `top_k` tables (in `table_order`) hold an id resident in this cache."""
        ids = np.ascontiguousarray(sample_ids)
        assert ids.ndim == 2 and ids.dtype in (np.uint64, np.float32)
        S, T = ids.shape
        order = None if table_order is None else np.ascontiguousarray(table_order, np.uint32)
        k = T if top_k is None else int(top_k)
        if order is not None:
            k = min(k, order.size)
        out = np.zeros(S, np.uint32)
        check_call(_LIB.hb_cache_score(self._h, ids.ctypes.data, KEYS_U64 if ids.dtype == np.uint64 else KEYS_F32,
                                       S, T, None if order is None else order.ctypes.data, k, int(fresh),
                                       out.ctypes.data))
        return out

    def resident(self, keys):
        """bool[n]: which of `keys` have a line in this cache (no policy touch)."""
        keys = np.ascontiguousarray(keys)
        assert keys.dtype in (np.uint64, np.float32)
        out = np.zeros(keys.size, np.uint8)
        check_call(_LIB.hb_cache_probe(self._h, keys.ctypes.data, KEYS_U64 if keys.dtype == np.uint64 else KEYS_F32,
                                       keys.size, out.ctypes.data))
        return out.astype(bool).reshape(keys.shape)

    # ---- debug surface (python_api.cc:56-60) -----------------------------------------------
    def size(self):
        self._drain()
        n = _sz()
        check_call(_LIB.hb_cache_size(self._h, ctypes.byref(n)))
        return n.value

    def count(self, k):
        self._drain()
        out = ctypes.c_int()
        check_call(_LIB.hb_cache_count(self._h, int(k), ctypes.byref(out)))
        return out.value

    def keys(self):
        self._drain()
        cap = self.size()
        buf = np.empty(max(cap, 1), np.uint64)
        n = _sz()
        check_call(_LIB.hb_cache_keys(self._h, buf.ctypes.data, buf.size, ctypes.byref(n)))
        return buf[:n.value].copy()

    def lookup(self, k):
        """Policy lookup of one key (touches replacement state, like the reference)."""
        self._drain()
        first = self._last_call() + 1
        found, ver = ctypes.c_int(), ctypes.c_int64()
        data = np.zeros(self._width, np.float32)
        check_call(_LIB.hb_cache_touch(self._h, int(k), ctypes.byref(found), ctypes.byref(ver),
                                       data.ctypes.data))
        self._drain(skip_from=first)   # single-key calls leave no perf entry (as in the reference)
        return Embedding(k, ver.value, data) if found.value else None

    def peek(self, k):
        """Read one line WITHOUT touching replacement state (not in the reference)."""
        self._drain()
        found, ver, upd = ctypes.c_int(), ctypes.c_int64(), ctypes.c_int64()
        data = np.zeros(self._width, np.float32)
        grad = np.zeros(self._width, np.float32)
        check_call(_LIB.hb_cache_peek(self._h, int(k), ctypes.byref(found), ctypes.byref(ver),
                                      ctypes.byref(upd), data.ctypes.data, grad.ctypes.data))
        return Embedding(k, ver.value, data, grad, upd.value) if found.value else None

    def insert(self, e):
        self._drain()
        assert e.data.size == self._width
        first = self._last_call() + 1
        check_call(_LIB.hb_cache_insert(self._h, e.key, e.version, e.data.ctypes.data))
        self._drain(skip_from=first)

    def __repr__(self):
        pull, push = self._bounds()
        return "<Cache : %d/%d , id:%d , width:%d , bound:%d %d>" % (
            self.size(), self._limit, self._node_id, self._width, pull, push)


class LRUCache(CacheBase):
    _policy = "lru"


class LFUCache(CacheBase):
    _policy = "lfu"


class LFUOptCache(CacheBase):
    _policy = "lfuopt"


def debug():
    from ._base import version
    print("herald_b200 hetu_cache: %s" % version())
