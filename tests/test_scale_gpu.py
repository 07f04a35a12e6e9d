"""GPU parity AT THE BENCHMARKED SCALE (VERDICT r1 "parity gaps"): the same checks as
tests/test_cache_gpu.py — gathered rows, per-call counters, owner rows + versions, resident key
set, all bit-exact against the compiled reference (oracle/_ref) — but at BASELINE C2's shape
(B = 8192 x 26 fields = 212 992 float32-carried ids per call, V = 33 762 577, LRU limit 10 %,
D = 128), at D = 512 / bound 10 / LFU with calls larger than the sort's direct-sum limit, and the
sort's chained look-back path on its own (n = 300 000 and 1 700 000).

The table is too large to mirror whole on the host: the GPU table is initialised on the device,
the rows the run will touch are read back (hb_table_read_rows_at) and loaded into a
zero-initialised oracle server (oracle.ref.Server.load_rows_at) before the first call.
"""
import ctypes

import numpy as np
import pytest

from common import PULL_KEYS, PUSH_KEYS, assert_bits_equal, perf_subset

pytestmark = pytest.mark.gpu

FIELDS = 26


def _mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable"):
                    return int(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


@pytest.fixture(scope="module")
def ref_oracle():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref (the compiled reference) is not built")
    return ref


class ScaleRun(object):
    """One table + cache on the GPU next to a sparse mirror on the reference oracle."""
    _next = [0]

    def __init__(self, ref, V, D, policy, limit, bound, touched):
        import herald_b200 as hb
        from herald_b200 import ps
        from herald_b200.cstable import CacheSparseTable
        self.hb, self.V, self.D = hb, V, D
        ScaleRun._next[0] += 1
        self.node_id = 8100 + ScaleRun._next[0]
        self.comm = hb.get_worker_communicate()
        self.table = self.comm.InitTensor(self.node_id, ps.kCacheTable, V, D, ps.Normal, 0.0, 0.01, 77)
        self.cst = CacheSparseTable(limit, V, D, self.node_id, policy, bound)
        self.cst.perf_enabled(True)
        self.touched = np.unique(np.asarray(touched, np.uint64))
        rows0, ver0 = self.table.read_rows_at(self.touched)
        assert np.all(ver0 == 0) and np.any(rows0 != 0)
        self.osrv = ref.Server(V, D)                      # zero rows, zero versions
        self.osrv.load_rows_at(self.touched, rows0)
        del rows0
        self.oc = ref.Cache(self.osrv, policy, limit, bound)
        self.host = hb.cpu(0)

    def close(self):
        self.cst = None
        self.oc = None
        self.osrv.close()
        self.comm.ClearTensor(self.node_id)

    def _nd(self, a):
        return self.hb.array(np.ascontiguousarray(a, np.float32), self.host)

    def lookup(self, ids_f32, what):
        """ids as Hetu carries them: float32 (cache.cc:49-58 casts them to uint64)."""
        ids_f32 = np.ascontiguousarray(ids_f32, np.float32).reshape(-1)
        dest = self.hb.empty((ids_f32.size, self.D), self.host)
        self.cst.embedding_lookup(self._nd(ids_f32), dest, sync=True)
        exp = self.oc.embedding_lookup(ids_f32.astype(np.uint64))
        assert_bits_equal(dest.host_view(), exp, "gathered rows " + what)
        g, o = self.cst.perf[-1], self.oc.perf[-1]
        assert perf_subset(g, PULL_KEYS) == perf_subset(o, PULL_KEYS), (what, g, dict(o))

    def update(self, ids_f32, grads, what, push_keys=None):
        ids_f32 = np.ascontiguousarray(ids_f32, np.float32).reshape(-1)
        if push_keys is None:
            self.cst.embedding_update(self._nd(ids_f32), self._nd(grads), sync=True)
        else:
            self.cst.embedding_update_with_push_keys(self._nd(ids_f32), push_keys, self._nd(grads),
                                                     sync=True)
        self.oc.embedding_update(ids_f32.astype(np.uint64), grads, push_keys)
        g, o = self.cst.perf[-1], self.oc.perf[-1]
        assert perf_subset(g, PUSH_KEYS) == perf_subset(o, PUSH_KEYS), (what, g, dict(o))

    def check_state(self, what):
        rows, ver = self.table.read_rows_at(self.touched)
        orows, over = self.osrv.rows_at(self.touched)
        assert_bits_equal(rows, orows, "owner rows " + what)
        assert np.array_equal(ver, over), what + ": owner versions differ"
        assert np.array_equal(self.cst.keys(), self.oc.keys()), what + ": resident key sets differ"


def test_c2_shape_lru(ref_oracle):
    """BASELINE configs[1]: B = 8192, 26 fields, D = 128, V = 33 762 577, LRU ratio 0.1, bound 0,
    Zipf(1.05) ids carried as float32 (ids > 2^24 round), cache pre-filled with the hottest ids —
    the exact workload bench.py times — for 5 update + lookup steps."""
    import bench
    V, D, B = bench.VOCAB, 128, 8192
    if _mem_available_gb() < V * D * 4 / 1e9 * 1.5 + 12:
        V = 4_000_000                                  # host RAM binds: same shape, smaller table
    limit = bench.cache_limit(V, 0.1)
    steps = 5
    ids = [bench.make_ids(s, B, V).reshape(-1) for s in range(steps + 1)]
    fill = [bench.hottest_ids(lo, min(lo + (1 << 20), limit), V, np.float32)
            for lo in range(0, limit, 1 << 20)]
    touched = np.concatenate([a.astype(np.uint64) for a in ids + fill])
    run = ScaleRun(ref_oracle, V, D, "lru", limit, 0, touched)
    try:
        for k, f in enumerate(fill):
            run.lookup(f, "fill %d" % k)
        grads = (np.random.default_rng(7).normal(0, 1e-3, (B * FIELDS, D)) * 1e-2).astype(np.float32)
        run.lookup(ids[0], "first lookup")
        for s in range(steps):
            run.update(ids[s], grads, "update %d" % s)
            run.lookup(ids[s + 1], "lookup %d" % (s + 1))
        run.check_state("after %d steps" % steps)
        assert len(run.cst.perf) and run.cst.perf[-1]["num_unique"] > 100_000
    finally:
        run.close()


@pytest.mark.parametrize("policy", ["lfu", "lru"])
def test_d512_bound10_large_calls(ref_oracle, policy):
    """BASELINE configs[3] flavour: D = 512, bound 10 (lines carry pending gradients, pushes and
    pulls are rare), calls of 301 600 keys (above the 229 376 of the sort's direct-sum path, so
    the cache's sort takes the chained look-back), every other update driven by a plan."""
    V, D, B = 1_000_003, 512, 11_600
    limit, bound, steps = 120_000, 10, 4
    rng = np.random.default_rng(11)
    ids = [((rng.zipf(1.05, B * FIELDS) - 1) * 7919 % V).astype(np.float32) for _ in range(steps + 1)]
    touched = np.concatenate([a.astype(np.uint64) for a in ids])
    run = ScaleRun(ref_oracle, V, D, policy, limit, bound, touched)
    try:
        grads = rng.normal(0, 1e-4, (B * FIELDS, D)).astype(np.float32)
        run.lookup(ids[0], "first lookup")
        for s in range(steps):
            plan = None
            if s % 2 == 1:
                u = np.unique(ids[s].astype(np.uint64))
                plan = u[rng.random(u.size) < 0.4]
            run.update(ids[s], grads, "update %d" % s, push_keys=plan)
            run.lookup(ids[s + 1], "lookup %d" % (s + 1))
        run.check_state("after %d steps" % steps)
    finally:
        run.close()


@pytest.mark.parametrize("n,vocab", [(300_000, 33_762_577), (1_700_000, 100_000_000)])
def test_unique_inverse_lookback_path(n, vocab):
    """Sorts of more than 56 tiles (n > 229 376) take the chained decoupled look-back
    (hb_sort.cuh kSortDirectTiles); C5's largest batch is 64k x 26 = 1.7 M ids of a 1e8-row table."""
    import herald_b200 as hb
    from herald_b200._base import _LIB, check_call
    from oracle import ops_port
    rng = np.random.default_rng(n)
    ids = ((rng.zipf(1.05, n) - 1) * 6180339 % vocab).astype(np.float32)
    d_ids = hb.array(ids, hb.gpu(0))
    uq, iv, cnt = hb.empty((n,), hb.gpu(0)), hb.empty((n,), hb.gpu(0)), hb.empty((2,), hb.gpu(0))
    for _ in range(2):                                  # twice: the epoch-tagged status words are reused
        check_call(_LIB.HBUniqueIndexedSlices(d_ids.handle, uq.handle, iv.handle,
                                              ctypes.c_void_p(cnt.data_ptr), None))
        U = int(cnt.asnumpy().view(np.int64)[0])
        uniq, inv = ops_port.unique_inverse(ids)
        assert U == len(uniq)
        assert np.array_equal(uq.asnumpy()[:U], uniq)
        assert np.array_equal(iv.asnumpy().astype(np.int64), inv)
