"""TEST INFRASTRUCTURE: loader for the compiled reference oracle (oracle/_ref).

``oracle/_ref/hetu_cache*.so`` is the reference's own worker cache and server
handlers (built from /root/reference by oracle/Makefile, sources never copied)
with the ZMQ/RDMA transport replaced by oracle/ref_shim/transport_shim.cc.  It
is git-ignored but travels to the GPU box.  Drive it synchronously
(``.wait()`` after every call): the reference moves ``evict_`` without a lock
(src/hetu_cache/src/cache.cc:144).
"""
import ctypes
import glob
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")

_mod = None
_lib = None


def available():
    return bool(glob.glob(os.path.join(_REF, "hetu_cache*.so")))


def module():
    """The reference ``hetu_cache`` pybind module."""
    global _mod, _lib
    if _mod is None:
        if not available():
            raise RuntimeError("oracle/_ref is not built: run `make -C oracle ref` "
                               "(needs /root/reference)")
        sys.path.insert(0, _REF)
        try:
            import hetu_cache  # noqa: the reference module name
        finally:
            sys.path.remove(_REF)
        _mod = hetu_cache
        _lib = ctypes.CDLL(hetu_cache.__file__)
        sz = ctypes.c_size_t
        _lib.oracle_set_servers.argtypes = [ctypes.c_int]
        _lib.oracle_init_table.argtypes = [ctypes.c_int, sz, sz, ctypes.c_int, ctypes.c_double,
                                           ctypes.c_double, ctypes.c_ulonglong]
        _lib.oracle_load_rows.argtypes = [ctypes.c_int, sz, sz, ctypes.c_void_p]
        _lib.oracle_read_rows.argtypes = [ctypes.c_int, ctypes.c_void_p]
        _lib.oracle_read_versions.argtypes = [ctypes.c_int, ctypes.c_void_p]
        _lib.oracle_clear_table.argtypes = [ctypes.c_int]
        _lib.oracle_read_rows_at.argtypes = [ctypes.c_int, ctypes.c_void_p, sz, ctypes.c_void_p,
                                             ctypes.c_void_p]
        _lib.oracle_add_rows_at.argtypes = [ctypes.c_int, ctypes.c_void_p, sz, ctypes.c_void_p]
    return _mod


_next_id = [1000]


class Server:
    """A kCacheTable on the in-process reference server(s)."""

    def __init__(self, length, width, rows=None, nserver=1, init=None):
        module()
        self.length, self.width = int(length), int(width)
        self.id = _next_id[0]
        _next_id[0] += 1
        _lib.oracle_set_servers(nserver)
        if rows is not None:
            rows = np.ascontiguousarray(rows, np.float32)
            assert rows.shape == (self.length, self.width)
            _lib.oracle_load_rows(self.id, self.length, self.width, rows.ctypes.data)
        else:
            kind, a, b, seed = init or (0, 0.0, 0.0, 0)
            _lib.oracle_init_table(self.id, self.length, self.width, kind, a, b, seed)

    def rows(self):
        out = np.empty((self.length, self.width), np.float32)
        _lib.oracle_read_rows(self.id, out.ctypes.data)
        return out

    def versions(self):
        out = np.empty(self.length, np.int64)
        _lib.oracle_read_versions(self.id, out.ctypes.data)
        return out

    # ---- selected rows only: tables too large to copy whole (bench-scale parity) ------------
    def rows_at(self, keys):
        """(rows, versions) of ascending `keys` as a never-synced worker would be sent them."""
        keys = np.ascontiguousarray(keys, np.uint64).reshape(-1)
        assert keys.size < 2 or bool(np.all(keys[1:] > keys[:-1])), "ascending unique keys"
        rows = np.empty((keys.size, self.width), np.float32)
        ver = np.empty(keys.size, np.int64)
        if keys.size:
            _lib.oracle_read_rows_at(self.id, keys.ctypes.data, keys.size, rows.ctypes.data,
                                     ver.ctypes.data)
        return rows, ver

    def load_rows_at(self, keys, rows):
        """rows[keys] += given rows (SparsePush): on a zero-initialised table this loads them."""
        keys = np.ascontiguousarray(keys, np.uint64).reshape(-1)
        assert keys.size < 2 or bool(np.all(keys[1:] > keys[:-1])), "ascending unique keys"
        rows = np.ascontiguousarray(rows, np.float32)
        assert rows.shape == (keys.size, self.width)
        if keys.size:
            _lib.oracle_add_rows_at(self.id, keys.ctypes.data, keys.size, rows.ctypes.data)

    def close(self):
        if self.id is not None:
            _lib.oracle_clear_table(self.id)
            self.id = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Cache:
    """Reference cache object with the same convenience surface as oracle.port.Cache."""

    def __init__(self, server, policy, limit, bound=None):
        m = module()
        cls = {"lru": m.LRUCache, "lfu": m.LFUCache, "lfuopt": m.LFUOptCache}[policy.lower()]
        self.server = server
        self.width = server.width
        self.limit = int(limit)
        self.c = cls(self.limit, server.length, server.width, server.id)
        self.c.perf_enabled = True
        if bound is not None:
            self.set_bounds(bound, bound)

    def set_bounds(self, pull_bound, push_bound):
        self.c.pull_bound = int(pull_bound)
        self.c.push_bound = int(push_bound)

    def bypass(self, on=True):
        self.c.bypass() if on else self.c.undo_bypass()

    @property
    def perf(self):
        return self.c.perf

    def embedding_lookup(self, keys, dest=None):
        keys = np.ascontiguousarray(keys, np.uint64).reshape(-1)
        if dest is None:
            dest = np.empty((keys.size, self.width), np.float32)
        self.c.embedding_lookup(keys, dest).wait()
        return dest

    def embedding_update(self, keys, grads, push_keys=None):
        keys = np.ascontiguousarray(keys, np.uint64).reshape(-1)
        grads = np.ascontiguousarray(grads, np.float32).reshape(keys.size, self.width)
        if push_keys is None:
            self.c.embedding_update(keys, grads).wait()
        else:
            pk = np.ascontiguousarray(push_keys, np.uint64).reshape(-1)
            self.c.embedding_update_with_push_keys(keys, pk, grads).wait()
        return self.c.perf[-1]

    def embedding_push_pull(self, pull_keys, push_keys, grads, dest=None):
        # the reference only exposes the float32-keyed raw entry point for this call
        pk = np.ascontiguousarray(pull_keys, np.float32).reshape(-1)
        sk = np.ascontiguousarray(push_keys, np.float32).reshape(-1)
        grads = np.ascontiguousarray(grads, np.float32).reshape(sk.size, self.width)
        if dest is None:
            dest = np.empty((pk.size, self.width), np.float32)
        self.c.embedding_push_pull_raw(pk.ctypes.data, dest.ctypes.data, pk.size,
                                       sk.ctypes.data, grads.ctypes.data, sk.size).wait()
        return dest

    def size(self):
        return self.c.size()

    def keys(self):
        return np.asarray(self.c.keys(), np.uint64)

    def line(self, key):
        # NOTE: `lookup` is a policy touch in the reference (python_api.cc:57); use
        # `count` first and only call this at the end of a test sequence.
        if not self.c.count(int(key)):
            return None
        e = self.c.lookup(int(key))
        return dict(version=e.version, data=np.array(e.data, np.float32))
