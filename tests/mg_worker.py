"""One rank of the multi-GPU parity run (launched by tests/test_multi_gpu.py under torchrun).

Every rank drives its own cache (worker role) over the row-sharded table (owner role) through
the reference-facing API, with per-rank seeded batches, in Hetu's BSP order: update(batch t) on
every rank, then lookup(batch t+1).  Every rank also replays the WHOLE group on the oracle —
one reference server, one reference cache per rank, calls issued in rank order — and compares
bit-for-bit: its gathered rows and perf counters each step, and at the end its table shard
(rows and versions) and resident key set.  The owner applies pushes in source-rank order, which
is what makes the multi-worker result reproducible (the reference's own order is arrival order).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    policy = os.environ.get("MG_POLICY", "lru")
    bound = int(os.environ.get("MG_BOUND", "0"))
    V, D = int(os.environ.get("MG_V", "1003")), int(os.environ.get("MG_D", "32"))
    limit = int(os.environ.get("MG_LIMIT", "120"))
    steps, max_n = int(os.environ.get("MG_STEPS", "25")), int(os.environ.get("MG_N", "300"))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    import herald_b200 as hb
    from herald_b200 import ps, hetu_cache, partition
    from common import assert_bits_equal, perf_subset, PULL_KEYS, PUSH_KEYS, zipf_keys
    from oracle import ref, port
    oracle = ref if ref.available() else port

    def exchange(b):
        obj = [b]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]

    ps.group_init(rank, world, local, exchange)
    comm = hb.worker_init(local)
    rows = np.random.default_rng(99).normal(0, 0.01, (V, D)).astype(np.float32)
    table = comm.InitTensor(7001, ps.kCacheTable, V, D, ps.Constant, 0.0)
    table.load_rows(rows)                       # clipped to this rank's shard
    cls = {"lru": hetu_cache.LRUCache, "lfu": hetu_cache.LFUCache, "lfuopt": hetu_cache.LFUOptCache}[policy]
    gc = cls(limit, V, D, 7001)
    gc.perf_enabled = True
    gc.pull_bound, gc.push_bound = bound, bound
    comm.BarrierWorker()
    if os.environ.get("MG_TABLES", "1") == "2":
        two_tables(rank, world, comm, oracle, gc, table, rows, policy, bound, V, D, limit, steps, max_n)
        del gc
        comm.ClearTensor(7001)
        ps.group_finalize()
        dist.destroy_process_group()
        print("mg_worker rank %d/%d ok (%s bound %d, two tables)" % (rank, world, policy, bound), flush=True)
        return

    osrv = oracle.Server(V, D, rows)
    ocs = [oracle.Cache(osrv, policy, limit) for _ in range(world)]
    for oc in ocs:
        oc.set_bounds(bound, bound)

    def batch(w, t):
        rng = np.random.default_rng(1000 * t + w)
        n = int(rng.integers(1, max_n))
        keys = zipf_keys(rng, n, V, 1.2)
        grads = rng.normal(0, 1e-3, (n, D)).astype(np.float32)
        return keys, grads

    def lookup_all(t):
        for w in range(world):
            keys, _ = batch(w, t)
            exp = ocs[w].embedding_lookup(keys)
            if w == rank:
                dest = np.zeros((keys.size, D), np.float32)
                gc.embedding_lookup(keys, dest).wait()
                assert_bits_equal(dest, exp, "rank %d step %d gathered rows" % (rank, t))
                g, o = gc.perf[-1], ocs[w].perf[-1]
                assert perf_subset(g, PULL_KEYS) == perf_subset(o, PULL_KEYS), (rank, t, g, dict(o))

    lookup_all(0)
    for t in range(steps):
        for w in range(world):                  # pushes reach the owner in rank order
            keys, grads = batch(w, t)
            ocs[w].embedding_update(keys, grads)
            if w == rank:
                gc.embedding_update(keys, grads).wait()
                g, o = gc.perf[-1], ocs[w].perf[-1]
                assert perf_subset(g, PUSH_KEYS) == perf_subset(o, PUSH_KEYS), (rank, t, g, dict(o))
        lookup_all(t + 1)
    comm.BarrierWorker()
    begin, n = partition.shard_range(rank, world, V)
    assert (begin, n) == table.shard()[:2]
    assert_bits_equal(table.read_rows(), osrv.rows()[begin:begin + n], "rank %d owner rows" % rank)
    assert np.array_equal(table.read_versions(), osrv.versions()[begin:begin + n]), "owner versions"
    assert np.array_equal(gc.keys(), ocs[rank].keys()), "resident key set"
    comm.BarrierWorker()
    del gc
    comm.ClearTensor(7001)
    ps.group_finalize()
    dist.destroy_process_group()
    print("mg_worker rank %d/%d ok (%s bound %d)" % (rank, world, policy, bound), flush=True)


def two_tables(rank, world, comm, oracle, gc_a, table_a, rows_a, policy, bound, V, D, limit, steps, max_n):
    """Two embedding tables, each with its own cache, driven INTERLEAVED and asynchronously (both
    updates enqueued before either is waited for): every cache has its own exchange flags and
    epochs, so the two exchanges may overlap in any order on the devices (with one
    process-global barrier epoch this corrupted a mailbox or timed out)."""
    import herald_b200 as hb
    from herald_b200 import ps, hetu_cache
    from common import assert_bits_equal, perf_subset, PULL_KEYS, PUSH_KEYS, zipf_keys
    Vb, Db, limit_b = 777, 16, 90
    rows_b = np.random.default_rng(5).normal(0, 0.01, (Vb, Db)).astype(np.float32)
    table_b = comm.InitTensor(7002, ps.kCacheTable, Vb, Db, ps.Constant, 0.0)
    table_b.load_rows(rows_b)
    cls = {"lru": hetu_cache.LRUCache, "lfu": hetu_cache.LFUCache, "lfuopt": hetu_cache.LFUOptCache}[policy]
    gc_b = cls(limit_b, Vb, Db, 7002)
    gc_b.perf_enabled = True
    gc_b.pull_bound, gc_b.push_bound = bound, bound
    comm.BarrierWorker()
    tabs = [dict(gc=gc_a, table=table_a, V=V, D=D, limit=limit, rows=rows_a, seed=0),
            dict(gc=gc_b, table=table_b, V=Vb, D=Db, limit=limit_b, rows=rows_b, seed=500000)]
    for tb in tabs:
        tb["osrv"] = oracle.Server(tb["V"], tb["D"], tb["rows"])
        tb["ocs"] = [oracle.Cache(tb["osrv"], policy, tb["limit"]) for _ in range(world)]
        for oc in tb["ocs"]:
            oc.set_bounds(bound, bound)

    def batch(tb, w, t):
        rng = np.random.default_rng(tb["seed"] + 1000 * t + w)
        n = int(rng.integers(1, max_n))
        return zipf_keys(rng, n, tb["V"], 1.2), rng.normal(0, 1e-3, (n, tb["D"])).astype(np.float32)

    def lookups(t):
        mine = []
        for tb in tabs:                                        # enqueue both, then wait
            keys, _ = batch(tb, rank, t)
            dest = np.zeros((keys.size, tb["D"]), np.float32)
            mine.append((tb["gc"].embedding_lookup(keys, dest), dest))
        for (w8, dest), tb in zip(mine, tabs):
            w8.wait()
            for w in range(world):
                keys, _ = batch(tb, w, t)
                exp = tb["ocs"][w].embedding_lookup(keys)
                if w == rank:
                    assert_bits_equal(dest, exp, "rank %d step %d rows" % (rank, t))
                    g, o = tb["gc"].perf[-1], tb["ocs"][w].perf[-1]
                    assert perf_subset(g, PULL_KEYS) == perf_subset(o, PULL_KEYS), (rank, t, g, dict(o))

    lookups(0)
    for t in range(steps):
        waits = []
        for tb in tabs:
            keys, grads = batch(tb, rank, t)
            waits.append(tb["gc"].embedding_update(keys, grads))
        for w8, tb in zip(waits, tabs):
            w8.wait()
            for w in range(world):
                keys, grads = batch(tb, w, t)
                tb["ocs"][w].embedding_update(keys, grads)
                if w == rank:
                    g, o = tb["gc"].perf[-1], tb["ocs"][w].perf[-1]
                    assert perf_subset(g, PUSH_KEYS) == perf_subset(o, PUSH_KEYS), (rank, t, g, dict(o))
        comm.BarrierWorker()
        lookups(t + 1)
    comm.BarrierWorker()
    from herald_b200 import partition
    for tb in tabs:
        begin, n = partition.shard_range(rank, world, tb["V"])
        assert_bits_equal(tb["table"].read_rows(), tb["osrv"].rows()[begin:begin + n], "owner rows")
        assert np.array_equal(tb["table"].read_versions(), tb["osrv"].versions()[begin:begin + n])
        assert np.array_equal(tb["gc"].keys(), tb["ocs"][rank].keys())
    comm.BarrierWorker()
    tabs[1]["gc"] = None
    del gc_b
    comm.ClearTensor(7002)


if __name__ == "__main__":
    main()
