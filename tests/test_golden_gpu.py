"""CUDA path replayed against the committed golden vectors (outputs of the reference itself,
tests/golden/make_golden.py) — needs no oracle at run time."""
import numpy as np
import pytest

from common import assert_bits_equal
import golden_util

pytestmark = pytest.mark.gpu


class GpuAdapter(object):
    _next = [0]

    def __init__(self, policy, limit, bound, rows0):
        import herald_b200 as hb
        from herald_b200 import hetu_cache, ps
        self.hb = hb
        GpuAdapter._next[0] += 1
        self.node_id = 8000 + GpuAdapter._next[0]
        V, D = rows0.shape
        self.D = D
        self.comm = hb.get_worker_communicate()
        self.table = self.comm.InitTensor(self.node_id, ps.kCacheTable, V, D, ps.Constant, 0.0)
        self.table.load_rows(rows0)
        cls = {"lru": hetu_cache.LRUCache, "lfu": hetu_cache.LFUCache,
               "lfuopt": hetu_cache.LFUOptCache}[policy]
        self.c = cls(limit, V, D, self.node_id)
        self.c.pull_bound = self.c.push_bound = bound
        self.c.perf_enabled = True

    def lookup(self, keys):
        keys = np.ascontiguousarray(keys, np.uint64)
        dest = np.zeros((keys.size, self.D), np.float32)
        self.c.embedding_lookup(keys, dest).wait()
        return dest

    def update(self, keys, grads, push_keys):
        keys = np.ascontiguousarray(keys, np.uint64)
        grads = np.ascontiguousarray(grads, np.float32)
        if push_keys is None:
            self.c.embedding_update(keys, grads).wait()
        else:
            self.c.embedding_update_with_push_keys(keys, np.ascontiguousarray(push_keys, np.uint64),
                                                   grads).wait()

    def push_pull(self, k1, k2, g):
        hb = self.hb
        pk, sk = hb.array(k1.astype(np.float32), hb.cpu(0)), hb.array(k2.astype(np.float32), hb.cpu(0))
        gr = hb.array(np.ascontiguousarray(g, np.float32), hb.cpu(0))
        dest = hb.empty((len(k1), self.D), hb.cpu(0))
        self.c.embedding_push_pull_raw(pk.data_ptr, dest.data_ptr, len(k1), sk.data_ptr, gr.data_ptr,
                                       len(k2)).wait()
        return dest.asnumpy()

    def last_perf(self):
        return self.c.perf[-1]

    def keys(self):
        return self.c.keys()

    def rows(self):
        return self.table.read_rows()

    def versions(self):
        return self.table.read_versions()

    def line(self, k):
        e = self.c.peek(k)
        return e.version, e.data

    def __del__(self):
        self.c = None
        self.comm.ClearTensor(self.node_id)


@pytest.mark.parametrize("path", golden_util.fixtures(), ids=lambda p: p.split("/")[-1][:-4])
def test_cuda_matches_reference_golden(path):
    golden_util.replay(path, GpuAdapter, assert_bits_equal)
