"""TEST INFRASTRUCTURE: ctypes front-end of oracle/hetu_port.cc (the CPU port).

Mirrors the reference's pybind surface (src/hetu_cache/src/python_api.cc:32-76)
closely enough that tests can drive port, reference and CUDA product with the
same call sequence.  Synchronous: every call completes before returning.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libhetu_port.so")


def build():
    """Compile the port (g++, no CUDA, no reference sources needed)."""
    subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)


def _load():
    if not os.path.exists(_LIBPATH):
        build()
    lib = ctypes.CDLL(_LIBPATH)
    vp, sz, i64, u64p, f32p, i64p = (ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int64,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p)
    lib.hp_server_create.restype = vp
    lib.hp_server_create.argtypes = [sz, sz]
    lib.hp_server_destroy.argtypes = [vp]
    lib.hp_server_load.argtypes = [vp, f32p]
    lib.hp_server_read.argtypes = [vp, f32p, i64p]
    lib.hp_cache_create.restype = vp
    lib.hp_cache_create.argtypes = [vp, ctypes.c_int, sz, sz]
    lib.hp_cache_destroy.argtypes = [vp]
    lib.hp_cache_set_bounds.argtypes = [vp, i64, i64]
    lib.hp_cache_set_bypass.argtypes = [vp, ctypes.c_int]
    lib.hp_cache_lookup.argtypes = [vp, u64p, sz, f32p, i64p]
    lib.hp_cache_update.argtypes = [vp, u64p, sz, f32p, i64p]
    lib.hp_cache_update_push_keys.argtypes = [vp, u64p, sz, u64p, sz, f32p, i64p]
    lib.hp_cache_push_pull.argtypes = [vp, u64p, sz, f32p, u64p, sz, f32p]
    lib.hp_cache_size.restype = sz
    lib.hp_cache_size.argtypes = [vp]
    lib.hp_cache_num_evicted_pending.restype = sz
    lib.hp_cache_num_evicted_pending.argtypes = [vp]
    lib.hp_cache_keys.argtypes = [vp, u64p]
    lib.hp_cache_line.restype = ctypes.c_int
    lib.hp_cache_line.argtypes = [vp, ctypes.c_uint64, i64p, i64p, f32p, f32p]
    lib.hp_cache_touch.restype = ctypes.c_int
    lib.hp_cache_touch.argtypes = [vp, ctypes.c_uint64]
    lib.hp_cache_insert_line.argtypes = [vp, ctypes.c_uint64, i64, f32p]
    lib.hp_unique.argtypes = [u64p, sz, u64p, u64p, ctypes.c_void_p]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


POLICY = {"lru": 0, "lfu": 1, "lfuopt": 2}
_PERF_PULL = ("num_all", "num_unique", "num_miss", "num_evict", "num_transfered", "is_full")


def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def unique(keys):
    """Sorted unique + inverse (src/hetu_cache/include/unqiue_tools.h:27-48)."""
    keys = _u64(keys).reshape(-1)
    uniq = np.empty(keys.size, np.uint64)
    inv = np.empty(keys.size, np.uint64)
    n = ctypes.c_size_t(0)
    lib().hp_unique(keys.ctypes.data, keys.size, uniq.ctypes.data, inv.ctypes.data,
                    ctypes.byref(n))
    return uniq[:n.value].copy(), inv.astype(np.int64)


class Server:
    """Owner-side table: rows + versions (ps-lite CacheTable)."""

    def __init__(self, length, width, rows=None):
        self.length, self.width = int(length), int(width)
        self.h = lib().hp_server_create(self.length, self.width)
        if rows is not None:
            self.load(rows)

    def load(self, rows):
        rows = _f32(rows)
        assert rows.shape == (self.length, self.width)
        lib().hp_server_load(self.h, rows.ctypes.data)

    def rows(self):
        out = np.empty((self.length, self.width), np.float32)
        lib().hp_server_read(self.h, out.ctypes.data, None)
        return out

    def versions(self):
        out = np.empty(self.length, np.int64)
        lib().hp_server_read(self.h, None, out.ctypes.data)
        return out

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.hp_server_destroy(self.h)
            self.h = None


class Cache:
    """Worker cache against one Server; policy in {'lru','lfu','lfuopt'}."""

    def __init__(self, server, policy, limit, bound=None):
        self.server = server
        self.width = server.width
        self.limit = int(limit)
        self.h = lib().hp_cache_create(server.h, POLICY[policy.lower()], self.limit, self.width)
        self.perf = []
        self.pull_bound = self.push_bound = 5
        if bound is not None:
            self.set_bounds(bound, bound)

    def set_bounds(self, pull_bound, push_bound):
        self.pull_bound, self.push_bound = int(pull_bound), int(push_bound)
        lib().hp_cache_set_bounds(self.h, self.pull_bound, self.push_bound)

    def bypass(self, on=True):
        lib().hp_cache_set_bypass(self.h, int(on))

    def _perf(self, kind, raw):
        d = dict(zip(_PERF_PULL, (int(x) for x in raw)))
        d["type"] = kind
        d["is_full"] = bool(d["is_full"])
        if kind == "Pull":
            d.pop("num_evict")
        self.perf.append(d)
        return d

    def embedding_lookup(self, keys, dest=None):
        keys = _u64(keys).reshape(-1)
        if dest is None:
            dest = np.empty((keys.size, self.width), np.float32)
        assert dest.dtype == np.float32 and dest.flags.c_contiguous
        raw = np.zeros(6, np.int64)
        lib().hp_cache_lookup(self.h, keys.ctypes.data, keys.size, dest.ctypes.data,
                              raw.ctypes.data)
        self._perf("Pull", raw)
        return dest

    def embedding_update(self, keys, grads, push_keys=None):
        keys = _u64(keys).reshape(-1)
        grads = _f32(grads).reshape(keys.size, self.width)
        raw = np.zeros(6, np.int64)
        if push_keys is None:
            lib().hp_cache_update(self.h, keys.ctypes.data, keys.size, grads.ctypes.data,
                                  raw.ctypes.data)
        else:
            pk = _u64(push_keys).reshape(-1)
            lib().hp_cache_update_push_keys(self.h, keys.ctypes.data, keys.size,
                                            pk.ctypes.data, pk.size, grads.ctypes.data,
                                            raw.ctypes.data)
        return self._perf("Push", raw)

    def embedding_push_pull(self, pull_keys, push_keys, grads, dest=None):
        pull_keys = _u64(pull_keys).reshape(-1)
        push_keys = _u64(push_keys).reshape(-1)
        grads = _f32(grads).reshape(push_keys.size, self.width)
        if dest is None:
            dest = np.empty((pull_keys.size, self.width), np.float32)
        lib().hp_cache_push_pull(self.h, pull_keys.ctypes.data, pull_keys.size,
                                 dest.ctypes.data, push_keys.ctypes.data, push_keys.size,
                                 grads.ctypes.data)
        return dest

    def size(self):
        return int(lib().hp_cache_size(self.h))

    def num_evicted_pending(self):
        return int(lib().hp_cache_num_evicted_pending(self.h))

    def keys(self):
        out = np.empty(self.size(), np.uint64)
        lib().hp_cache_keys(self.h, out.ctypes.data)
        return out

    def line(self, key):
        """-> None | dict(version, updates, data, grad)."""
        ver, upd = ctypes.c_int64(0), ctypes.c_int64(0)
        data = np.zeros(self.width, np.float32)
        grad = np.zeros(self.width, np.float32)
        ok = lib().hp_cache_line(self.h, int(key), ctypes.byref(ver), ctypes.byref(upd),
                                 data.ctypes.data, grad.ctypes.data)
        if not ok:
            return None
        return dict(version=ver.value, updates=upd.value, data=data, grad=grad)

    def touch(self, key):
        return bool(lib().hp_cache_touch(self.h, int(key)))

    def insert(self, key, version, data):
        data = _f32(data)
        lib().hp_cache_insert_line(self.h, int(key), int(version), data.ctypes.data)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.hp_cache_destroy(self.h)
            self.h = None
