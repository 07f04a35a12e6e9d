"""GPU parity: herald_b200.hetu_cache (CUDA) against the oracle, same seeded call sequences.

Bit-exact: gathered rows, resident key sets, perf counters (num_unique / num_miss /
num_transfered / num_evict / is_full), owner rows and versions, cached line data and versions.
"""
import numpy as np
import pytest

from common import GpuHarness, assert_bits_equal, zipf_keys

pytestmark = pytest.mark.gpu


def _rows(rng, V, D):
    return rng.normal(0, 0.01, (V, D)).astype(np.float32)


def _run_sequence(oracle, policy, limit, bound, V=300, D=8, steps=60, seed=0, max_n=80,
                  push_keys=False, push_pull=False, a=1.3):
    rng = np.random.default_rng(seed)
    h = GpuHarness(oracle, policy, limit, bound, _rows(rng, V, D))
    try:
        for t in range(steps):
            tag = "%s limit=%d bound=%d step=%d" % (policy, limit, bound, t)
            n = int(rng.integers(1, max_n))
            keys = zipf_keys(rng, n, V, a)
            h.lookup(keys, tag)
            ukeys = keys if rng.random() < 0.7 else zipf_keys(rng, int(rng.integers(1, max_n)), V, a)
            grads = rng.normal(0, 1e-3, (len(ukeys), D)).astype(np.float32)
            pk = None
            if push_keys:
                pk = np.unique(rng.choice(ukeys, size=max(1, len(ukeys) // 3)))
            h.update(ukeys, grads, pk, tag)
            if push_pull and t % 3 == 0:
                k1 = zipf_keys(rng, n, V, a)
                k2 = zipf_keys(rng, int(rng.integers(1, max_n)), V, a)
                h.push_pull(k1, k2, rng.normal(0, 1e-3, (len(k2), D)).astype(np.float32), tag)
            if t % 10 == 9:
                h.check_state(tag)
        h.check_state("final")
        h.check_lines("final")
    finally:
        h.close()


@pytest.mark.parametrize("policy", ["lru", "lfu", "lfuopt"])
@pytest.mark.parametrize("limit", [5, 30, 150, 400])
@pytest.mark.parametrize("bound", [0, 2, 10])
def test_lookup_update_sequence(oracle_impl, policy, limit, bound):
    _run_sequence(oracle_impl, policy, limit, bound, seed=limit * 31 + bound)


@pytest.mark.parametrize("policy", ["lru", "lfu", "lfuopt"])
@pytest.mark.parametrize("limit", [5, 30, 150])
@pytest.mark.parametrize("bound", [0, 10])
def test_update_with_push_keys(oracle_impl, policy, limit, bound):
    # Laia/Herald plan path: cache.cc:248-334 (version quirk s7 of SURVEY §9 included)
    _run_sequence(oracle_impl, policy, limit, bound, seed=7 + limit, push_keys=True)


@pytest.mark.parametrize("policy", ["lru", "lfu", "lfuopt"])
@pytest.mark.parametrize("limit", [5, 30, 150])
@pytest.mark.parametrize("bound", [0, 2])
def test_push_pull(oracle_impl, policy, limit, bound):
    # ASP prefetch path: cache.cc:356-422
    _run_sequence(oracle_impl, policy, limit, bound, seed=13 + limit, push_pull=True)


@pytest.mark.parametrize("policy", ["lru", "lfu", "lfuopt"])
def test_wide_rows_and_larger_batches(oracle_impl, policy):
    # D = 128 (the north-star width): one float4 per lane; batches of a few thousand keys
    _run_sequence(oracle_impl, policy, limit=2000, bound=0, V=20000, D=128, steps=12, seed=3,
                  max_n=6000, a=1.05)


@pytest.mark.parametrize("policy,bound,D", [("lru", 0, 8), ("lfu", 2, 128), ("lru", 10, 40),
                                            ("lfuopt", 0, 6)])
@pytest.mark.parametrize("mode", ["plain", "plan", "pushpull"])
def test_hot_segment_path(oracle_impl, policy, bound, D, mode):
    """Ids occurring more than 2 times go through the column-split hot path (production
    threshold: 64); results stay bit-identical to the serial reference."""
    from herald_b200._base import set_hot_threshold
    set_hot_threshold(2)
    try:
        _run_sequence(oracle_impl, policy, limit=60, bound=bound, V=300, D=D, steps=25, seed=21,
                      max_n=600, push_keys=mode == "plan", push_pull=mode == "pushpull", a=1.2)
    finally:
        set_hot_threshold(64)


@pytest.mark.parametrize("D,bound", [(128, 0), (40, 0), (8, 2), (512, 10), (6, 0), (20, 0)])
def test_very_hot_id(oracle_impl, D, bound):
    """One id repeated thousands of times (> kVeryHot: the 16-column chunks of the hot phase, with
    partial last chunks at D = 40 / 20 / 8 and the 4-byte copy path at D = 6) among cold ids."""
    rng = np.random.default_rng(17)
    V = 400
    h = GpuHarness(oracle_impl, "lru", 100, bound, _rows(rng, V, D))
    try:
        for t in range(3):
            keys = zipf_keys(rng, 6000, V, 1.3)
            keys[rng.random(6000) < 0.5] = 7      # ~3000 occurrences of id 7
            keys[rng.random(6000) < 0.1] = 11     # ~600 of id 11
            h.lookup(keys)
            h.update(keys, rng.normal(0, 1e-3, (6000, D)).astype(np.float32))
        h.check_state()
        h.check_lines()
    finally:
        h.close()


def test_width_not_multiple_of_four(oracle_impl):
    _run_sequence(oracle_impl, "lru", limit=40, bound=1, V=200, D=6, steps=30, seed=5)


def test_d512_bound10(oracle_impl):
    # config C4 shape: emb 512, bound 10
    _run_sequence(oracle_impl, "lru", limit=500, bound=10, V=5000, D=512, steps=10, seed=11,
                  max_n=2000, a=1.05)


def test_semantics_s2_lru_eviction_order(oracle_impl):
    # limit 4, cache {10,20,30,40}, touch {10,20}, miss {50,60} => {30,40} evicted
    rng = np.random.default_rng(0)
    h = GpuHarness(oracle_impl, "lru", 4, 0, _rows(rng, 100, 8))
    try:
        h.lookup(np.array([10, 20, 30, 40], np.uint64))
        h.lookup(np.array([10, 20], np.uint64))
        h.lookup(np.array([50, 60], np.uint64))
        assert list(h.gc.keys()) == [10, 20, 50, 60]
        h.check_state()
    finally:
        h.close()


def test_semantics_s3_dirty_eviction_flushed_by_next_update(oracle_impl):
    rng = np.random.default_rng(1)
    rows = _rows(rng, 100, 8)
    h = GpuHarness(oracle_impl, "lru", 4, 5, rows)
    try:
        h.lookup(np.array([1, 2, 3, 4], np.uint64))
        h.update(np.array([1, 2], np.uint64), np.full((2, 8), 0.5, np.float32))
        h.lookup(np.array([5, 6, 7, 8], np.uint64))  # evicts 1,2 (dirty) and 3,4 (clean)
        assert_bits_equal(h.table.read_rows()[1], rows[1], "row 1 untouched before the flush")
        h.lookup(np.array([11], np.uint64))
        h.update(np.array([11], np.uint64), np.zeros((1, 8), np.float32))
        p = h.gc.perf[-1]
        assert p["num_evict"] == 2 and p["num_transfered"] == 2
        assert_bits_equal(h.table.read_rows()[1], rows[1] + np.float32(0.5), "row 1 after flush")
        h.check_state()
    finally:
        h.close()


def test_semantics_s4_limit_smaller_than_batch(oracle_impl):
    rng = np.random.default_rng(2)
    h = GpuHarness(oracle_impl, "lru", 2, 0, _rows(rng, 50, 8))
    try:
        keys = np.array([4, 9, 14, 19], np.uint64)
        h.lookup(keys)
        h.update(keys, np.ones((4, 8), np.float32))
        p = h.gc.perf[-1]
        assert p["num_miss"] == 2 and p["num_transfered"] == 4
        h.check_state()
    finally:
        h.close()


def test_semantics_s5_bound_two_caches_one_table(oracle_impl):
    """Two workers (A, B) against one owner table, bound 2: A pushes on its 3rd update and B
    re-pulls exactly then."""
    import herald_b200 as hb
    from herald_b200 import hetu_cache, ps
    rng = np.random.default_rng(3)
    V, D = 40, 8
    rows = _rows(rng, V, D)
    comm = hb.get_worker_communicate()
    table = comm.InitTensor(7777, ps.kCacheTable, V, D, ps.Constant, 0.0)
    table.load_rows(rows)
    osrv = oracle_impl.Server(V, D, rows)
    try:
        ga, gb = hetu_cache.LRUCache(10, V, D, 7777), hetu_cache.LRUCache(10, V, D, 7777)
        oa, ob = oracle_impl.Cache(osrv, "lru", 10, 2), oracle_impl.Cache(osrv, "lru", 10, 2)
        for c in (ga, gb):
            c.pull_bound = c.push_bound = 2
            c.perf_enabled = True
        k = np.array([7], np.uint64)
        g = np.full((1, D), 0.25, np.float32)
        for step in range(7):
            da, db = np.zeros((1, D), np.float32), np.zeros((1, D), np.float32)
            ga.embedding_lookup(k, da).wait()
            gb.embedding_lookup(k, db).wait()
            assert_bits_equal(da, oa.embedding_lookup(k), "A step %d" % step)
            assert_bits_equal(db, ob.embedding_lookup(k), "B step %d" % step)
            assert ga.perf[-1]["num_transfered"] == oa.perf[-1]["num_transfered"]
            assert gb.perf[-1]["num_transfered"] == ob.perf[-1]["num_transfered"]
            ga.embedding_update(k, g).wait()
            oa.embedding_update(k, g)
            assert ga.perf[-1]["num_transfered"] == oa.perf[-1]["num_transfered"]
        assert np.array_equal(table.read_versions(), osrv.versions())
        assert_bits_equal(table.read_rows(), osrv.rows(), "owner rows")
        del ga, gb
    finally:
        comm.ClearTensor(7777)


@pytest.mark.parametrize("policy", ["lru", "lfu", "lfuopt"])
def test_bypass(oracle_impl, policy):
    rng = np.random.default_rng(4)
    h = GpuHarness(oracle_impl, policy, 20, 0, _rows(rng, 100, 8))
    try:
        h.gc.bypass()
        h.oc.bypass(True)
        for t in range(5):
            keys = zipf_keys(rng, 30, 100, 1.3)
            h.lookup(keys, "bypass")
            h.update(keys, rng.normal(0, 1e-3, (30, 8)).astype(np.float32), None, "bypass")
        assert h.gc.size() == 0
        h.gc.undo_bypass()
        h.oc.bypass(False)
        keys = zipf_keys(rng, 30, 100, 1.3)
        h.lookup(keys)
        h.check_state()
    finally:
        h.close()


def test_empty_and_single_key_batches(oracle_impl):
    rng = np.random.default_rng(5)
    h = GpuHarness(oracle_impl, "lru", 8, 0, _rows(rng, 50, 8))
    try:
        h.lookup(np.array([3], np.uint64))
        h.update(np.array([3], np.uint64), np.ones((1, 8), np.float32))
        h.lookup(np.array([3, 3, 3, 3], np.uint64))
        h.update(np.array([3, 3, 3, 3], np.uint64), np.ones((4, 8), np.float32))
        empty_keys = np.zeros(0, np.uint64)
        dest = np.zeros((0, 8), np.float32)
        h.gc.embedding_lookup(empty_keys, dest).wait()
        h.gc.embedding_update(empty_keys, dest).wait()
        h.lookup(np.array([3, 4], np.uint64))
        h.check_state()
    finally:
        h.close()


def test_float32_raw_keys_round_like_the_reference(oracle_impl):
    """Raw (NDArray) entry points carry ids as float32: key = (uint64)(float)id, so ids above
    2^24 collapse onto the nearest representable float."""
    import herald_b200 as hb
    rng = np.random.default_rng(6)
    V, D = (1 << 24) + 64, 4
    rows = np.zeros((V, D), np.float32)
    rows[:, 0] = np.arange(V, dtype=np.float64)  # exact up to 2^24, rounded above
    h = GpuHarness(oracle_impl, "lru", 16, 0, rows)
    try:
        ids = np.array([(1 << 24) + 1, (1 << 24) + 3, 5, (1 << 24) + 2], np.int64)
        f32 = ids.astype(np.float32)               # what Hetu's dataloader produces
        keys = hb.array(f32, hb.cpu(0))
        dest = hb.empty((4, D), hb.cpu(0))
        h.gc.embedding_lookup_raw(keys.data_ptr, dest.data_ptr, 4).wait()
        exp = h.oc.embedding_lookup(f32.astype(np.uint64))
        assert_bits_equal(dest.asnumpy(), exp, "rows for float32-rounded keys")
        assert h.gc.perf[-1]["num_unique"] == h.oc.perf[-1]["num_unique"]
    finally:
        h.close()


def test_device_pointer_entry_points(oracle_impl):
    """Keys, dest and grads resident in HBM (the B200-native call path: no PCIe in the step)."""
    import herald_b200 as hb
    rng = np.random.default_rng(8)
    V, D = 500, 128
    h = GpuHarness(oracle_impl, "lru", 100, 0, _rows(rng, V, D))
    try:
        for t in range(6):
            ids = zipf_keys(rng, 400, V, 1.2)
            keys_d = hb.array(ids.astype(np.float32), hb.gpu(0))
            dest_d = hb.empty((400, D), hb.gpu(0))
            h.gc.embedding_lookup_raw(keys_d.data_ptr, dest_d.data_ptr, 400).wait()
            assert_bits_equal(dest_d.asnumpy(), h.oc.embedding_lookup(ids), "device dest")
            g = rng.normal(0, 1e-3, (400, D)).astype(np.float32)
            grads_d = hb.array(g, hb.gpu(0))
            h.gc.embedding_update_raw(keys_d.data_ptr, grads_d.data_ptr, 400).wait()
            h.oc.embedding_update(ids, g)
        h.check_state()
    finally:
        h.close()


def test_single_key_debug_surface(oracle_impl):
    from herald_b200 import hetu_cache
    rng = np.random.default_rng(9)
    h = GpuHarness(oracle_impl, "lru", 3, 0, _rows(rng, 50, 8))
    try:
        h.lookup(np.array([1, 2, 3], np.uint64))
        assert h.gc.count(2) == 1 and h.gc.count(9) == 0
        e = h.gc.lookup(1)                     # touches: 1 becomes most recent
        assert e is not None and e.key == 1
        assert h.gc.lookup(40) is None
        h.oc.touch(1) if hasattr(h.oc, "touch") else h.oc.c.lookup(1)
        h.lookup(np.array([4], np.uint64))     # evicts 2 (least recent), not 1
        assert list(h.gc.keys()) == [1, 3, 4]
        h.gc.insert(hetu_cache.Embedding(9, 0, np.arange(8, dtype=np.float32)))
        assert h.gc.count(9) == 1 and h.gc.size() == 3
        assert "Cache : 3/3" in repr(h.gc)
    finally:
        h.close()


# ---- checkpoint: flush of dirty lines + shard save / load -------------------------
@pytest.mark.parametrize("policy", ["lru", "lfu"])
def test_flush_then_save_load_roundtrip(tmp_path, policy):
    """hb_cache_flush pushes what a bound of 10 keeps in the cache; the shard file has the
    reference's layout (PSAgent.h:447-476: "<dir>/<node>_<part>.dat", raw row-major float32)."""
    from oracle import port
    rng = np.random.default_rng(5)
    V, D, limit = 200, 8, 40
    h = GpuHarness(port, policy, limit, 10, _rows(rng, V, D))
    try:
        for t in range(25):
            keys = zipf_keys(rng, int(rng.integers(1, 60)), V, 1.3)
            h.lookup(keys, "t%d" % t)
            h.update(keys, rng.normal(0, 1e-3, (len(keys), D)).astype(np.float32), None, "t%d" % t)
        # expected owner state: the oracle flushes evict_ on any update call (cache.cc:142-166) ...
        h.oc.embedding_update(np.zeros(0, np.uint64), np.zeros((0, D), np.float32))
        rows, vers = h.osrv.rows().copy(), h.osrv.versions().copy()
        dirty = 0
        for k in h.oc.keys():  # ... and a push of a resident line is row += grad; ver += updates
            ln = h.oc.line(int(k))
            if ln["updates"]:
                rows[int(k)] = rows[int(k)] + ln["grad"]
                vers[int(k)] += ln["updates"]
                dirty += 1
        assert dirty > 0, "sequence left nothing dirty: the test would be vacuous"
        h.gc.flush()
        assert_bits_equal(h.table.read_rows(), rows, "owner rows after flush")
        assert np.array_equal(h.table.read_versions(), vers)
        for k in h.gc.keys():
            assert h.gc.peek(int(k)).updates == 0
        # second flush is a no-op
        h.gc.flush()
        assert_bits_equal(h.table.read_rows(), rows, "owner rows after 2nd flush")
        # save, clobber, load
        h.comm.SaveParam(h.node_id, str(tmp_path))
        raw = np.fromfile(str(tmp_path / ("%d_0.dat" % h.node_id)), np.float32).reshape(V, D)
        assert_bits_equal(raw, rows, "file contents")
        h.table.load_rows(np.zeros((V, D), np.float32))
        h.comm.LoadParam(h.node_id, str(tmp_path))
        assert_bits_equal(h.table.read_rows(), rows, "owner rows after load")
        assert np.array_equal(h.table.read_versions(), vers)
        # a file written by the reference (rows only, no .ver) loads too and keeps the versions
        (tmp_path / ("%d_0.ver" % h.node_id)).unlink()
        (rows * 2).astype(np.float32).tofile(str(tmp_path / ("%d_0.dat" % h.node_id)))
        h.comm.LoadParam(h.node_id, str(tmp_path))
        assert_bits_equal(h.table.read_rows(), rows * 2, "rows from a reference-format file")
        assert np.array_equal(h.table.read_versions(), vers)
    finally:
        h.close()


@pytest.mark.parametrize("slots", [16, 256, 4096])
@pytest.mark.parametrize("limit,bound,mode", [(5, 0, "plain"), (30, 2, "plain"), (150, 0, "plan"),
                                               (30, 0, "pushpull"), (400, 10, "plain")])
def test_lru_stamp_log_fallback(oracle_impl, monkeypatch, slots, limit, bound, mode):
    """LRU victims come from the stamp-log walk while [floor, now) fits the log and from the
    histogram select otherwise (the device decides per call).  A tiny log makes calls alternate
    between the two paths; hit/miss/evict sets and rows must stay those of the reference."""
    monkeypatch.setenv("HERALD_STAMP_LOG_SLOTS", str(slots))
    _run_sequence(oracle_impl, "lru", limit, bound, seed=7 * slots + limit, steps=80,
                  push_keys=mode == "plan", push_pull=mode == "pushpull")


def test_perf_sampling(oracle_impl):
    """hb_cache_set_perf_sampling: phase times only on every n-th pair of calls, counters always."""
    rng = np.random.default_rng(5)
    h = GpuHarness(oracle_impl, "lru", 50, 0, _rows(rng, 300, 8))
    try:
        h.gc.set_perf_sampling(4)
        for t in range(16):
            keys = zipf_keys(rng, 40, 300, 1.3)
            h.lookup(keys, "step %d" % t)
            h.update(keys, rng.normal(0, 1e-3, (len(keys), 8)).astype(np.float32), None, "step %d" % t)
        perf = list(h.gc.perf)[-32:]
        timed = [p for p in perf if p["lookup_time"] > 0]
        assert 0 < len(timed) < len(perf)
        assert all(p["num_all"] == 40 for p in perf)             # counters on every call
        assert all(p["time"] > 0 for p in timed)                 # durations on the sampled calls
    finally:
        h.close()


@pytest.mark.parametrize("policy,bound", [("lru", 10), ("lfu", 2)])
def test_laia_plan_as_push_keys(oracle_impl, policy, bound):
    """run_laia's loop for one worker of three: the planner (herald_b200.laia, host C++) chooses
    the samples and the push plan, the GPU cache executes lookup + update_with_push_keys; rows,
    versions and counters against the oracle driven with the same calls."""
    from herald_b200.laia import LaiaScheduler
    rng = np.random.default_rng(17)
    W, mini, T, nb, V, D, cap = 3, 24, 6, 6, 300, 8, 40
    emb = ((rng.zipf(1.15, (W * mini * nb, T)) - 1) % V).astype(np.uint64)
    s = LaiaScheduler()
    s.start(emb, emb.shape[0], T, 1, mini, nb, W, 0, cap, 2)
    h = GpuHarness(oracle_impl, policy, cap, bound, _rows(rng, V, D))
    try:
        assert s.step()
        dist = s.dist_of(0)
        for b in range(nb):
            keys = emb[dist.astype(np.int64)].reshape(-1)
            h.lookup(keys, "batch %d" % b)
            assert s.step()                                           # plan for batch b + 1
            grads = rng.normal(0, 1e-3, (keys.size, D)).astype(np.float32)
            h.update(keys, grads, s.plan_of(0), "batch %d" % b)
            dist = s.dist_of(0)
        h.check_state("final")
        h.check_lines("final")
    finally:
        h.close()


@pytest.mark.parametrize("n,V,D", [(50, 300, 8), (6000, 20000, 128), (300000, 400000, 8)])
def test_update_of_a_different_batch_of_the_same_size(oracle_impl, n, V, D):
    """The update's batch has the size of the batch its workspace holds but other ids: the
    device-side comparison (same_keys) fails and the sort kernels of the side stream really run,
    over the workspace the previous lookup filled.  Alternated with updates of the looked-up batch
    itself (the sort kernels return at once)."""
    rng = np.random.default_rng(n)
    h = GpuHarness(oracle_impl, "lru", max(5, V // 10), 0, _rows(rng, V, D))
    try:
        for t in range(6):
            keys = zipf_keys(rng, n, V, 1.1)
            h.lookup(keys, "step %d" % t)
            ukeys = keys if t % 2 else zipf_keys(rng, n, V, 1.1)
            h.update(ukeys, rng.normal(0, 1e-3, (n, D)).astype(np.float32), None, "step %d" % t)
        h.check_state("final")
        h.check_lines("final")
    finally:
        h.close()
