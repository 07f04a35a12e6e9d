"""Shared helpers for the parity tests: drive the CUDA cache and an oracle with one sequence."""
import numpy as np

PULL_KEYS = ("num_all", "num_unique", "num_miss", "num_transfered", "is_full")
PUSH_KEYS = ("num_all", "num_unique", "num_miss", "num_evict", "num_transfered", "is_full")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bits_equal(a, b, what=""):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not np.array_equal(bits(a), bits(b)):
        bad = np.argwhere(bits(a) != bits(b))
        raise AssertionError("%s: %d of %d elements differ, first at %s: %r vs %r" % (
            what, len(bad), a.size, bad[0], a[tuple(bad[0])], b[tuple(bad[0])]))


def zipf_keys(rng, n, vocab, a=1.05):
    return ((rng.zipf(a, n) - 1) % vocab).astype(np.uint64)


def perf_subset(d, keys):
    return {k: (bool(d[k]) if k == "is_full" else int(d[k])) for k in keys}


class GpuHarness(object):
    """A table + cache on the GPU next to the same pair in an oracle implementation."""
    _next_id = [0]

    def __init__(self, oracle, policy, limit, bound, rows, push_bound=None):
        import herald_b200 as hb
        from herald_b200 import hetu_cache, ps
        self.hb = hb
        V, D = rows.shape
        self.V, self.D = V, D
        GpuHarness._next_id[0] += 1
        self.node_id = 5000 + GpuHarness._next_id[0]
        self.comm = hb.get_worker_communicate()
        self.table = self.comm.InitTensor(self.node_id, ps.kCacheTable, V, D, ps.Constant, 0.0)
        self.table.load_rows(rows)
        cls = {"lru": hetu_cache.LRUCache, "lfu": hetu_cache.LFUCache,
               "lfuopt": hetu_cache.LFUOptCache}[policy]
        self.gc = cls(limit, V, D, self.node_id)
        self.gc.perf_enabled = True
        pb = bound if push_bound is None else push_bound
        self.gc.pull_bound, self.gc.push_bound = bound, pb
        self.osrv = oracle.Server(V, D, rows)
        self.oc = oracle.Cache(self.osrv, policy, limit)
        self.oc.set_bounds(bound, pb)

    def close(self):
        self.gc = None
        self.comm.ClearTensor(self.node_id)

    # --- one call on both sides, compared ---
    def lookup(self, keys, what=""):
        keys = np.ascontiguousarray(keys, np.uint64)
        dest = np.zeros((keys.size, self.D), np.float32)
        self.gc.embedding_lookup(keys, dest).wait()
        exp = self.oc.embedding_lookup(keys)
        assert_bits_equal(dest, exp, "gathered rows " + what)
        g, o = self.gc.perf[-1], self.oc.perf[-1]
        assert perf_subset(g, PULL_KEYS) == perf_subset(o, PULL_KEYS), (what, g, dict(o))
        return dest

    def update(self, keys, grads, push_keys=None, what=""):
        keys = np.ascontiguousarray(keys, np.uint64)
        grads = np.ascontiguousarray(grads, np.float32)
        if push_keys is None:
            self.gc.embedding_update(keys, grads).wait()
        else:
            push_keys = np.ascontiguousarray(push_keys, np.uint64)
            self.gc.embedding_update_with_push_keys(keys, push_keys, grads).wait()
        self.oc.embedding_update(keys, grads, push_keys)
        g, o = self.gc.perf[-1], self.oc.perf[-1]
        assert perf_subset(g, PUSH_KEYS) == perf_subset(o, PUSH_KEYS), (what, g, dict(o))

    def push_pull(self, pull_keys, push_keys, grads, what=""):
        hb = self.hb
        pk = hb.array(np.asarray(pull_keys, np.float32), hb.cpu(0))
        sk = hb.array(np.asarray(push_keys, np.float32), hb.cpu(0))
        gr = hb.array(np.asarray(grads, np.float32), hb.cpu(0))
        dest = hb.empty((len(pull_keys), self.D), hb.cpu(0))
        self.gc.embedding_push_pull_raw(pk.data_ptr, dest.data_ptr, len(pull_keys), sk.data_ptr,
                                        gr.data_ptr, len(push_keys)).wait()
        exp = self.oc.embedding_push_pull(pull_keys, push_keys, grads)
        assert_bits_equal(dest.asnumpy(), exp, "push_pull rows " + what)

    def check_state(self, what=""):
        assert np.array_equal(self.gc.keys(), self.oc.keys()), what + " resident key sets differ"
        assert_bits_equal(self.table.read_rows(), self.osrv.rows(), "owner rows " + what)
        assert np.array_equal(self.table.read_versions(), self.osrv.versions()), \
            what + " owner versions differ"

    def check_lines(self, what="", max_lines=200):
        """Per-line data/version (uses the oracle's touching lookup: call last)."""
        keys = self.gc.keys()
        if len(keys) > max_lines:
            keys = keys[:: max(1, len(keys) // max_lines)]
        for k in keys:
            g = self.gc.peek(int(k))
            o = self.oc.line(int(k))
            assert g is not None and o is not None, (what, int(k))
            assert g.version == o["version"], (what, int(k), g.version, o["version"])
            assert_bits_equal(g.data, o["data"], "%s line %d data" % (what, int(k)))
