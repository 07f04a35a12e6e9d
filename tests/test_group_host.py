"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo groups (no GPU, no compute
calls into the CUDA library).

  * the row-range partition (herald_b200.partition) against ps-lite's AveragePartitioner rule;
  * the BSP exchange order the GPU path implements — every rank's update applied at the owner in
    rank order, then the lookups — replayed on the oracle by both ranks independently: the two
    replays must agree bit-for-bit (the schedule is a function of the inputs only), and each rank's
    shard of the result is what tests/mg_worker.py compares the GPUs against.
"""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from herald_b200 import partition  # noqa: E402  (pure numpy: importable without a GPU)


def test_partition_matches_average_partitioner():
    for length, world in [(10, 3), (33762577, 8), (7, 8), (100000000, 8), (1003, 2), (5, 1)]:
        begins = [partition.shard_range(r, world, length) for r in range(world)]
        # contiguous cover, sizes differ by at most one, larger shards first (partitioner.h:46-57)
        assert begins[0][0] == 0 and sum(n for _, n in begins) == length
        for r in range(1, world):
            assert begins[r][0] == begins[r - 1][0] + begins[r - 1][1]
            assert begins[r - 1][1] - begins[r][1] in (0, 1)
        keys = np.unique(np.random.default_rng(length % 97).integers(0, length, 2000).astype(np.uint64))
        owner, local = partition.owner_of(keys, world, length)
        for k, o, l in zip(keys[:200], owner[:200], local[:200]):
            b, n = begins[o]
            assert b <= k < b + n and l == k - b
        lo = partition.split_sorted(keys, world, length)
        assert lo[0] == 0 and lo[-1] == len(keys)
        for o in range(world):
            assert np.all(owner[lo[o]:lo[o + 1]] == o)


def _replay(world, policy, bound, V, D, limit, steps):
    from oracle import port
    rows = np.random.default_rng(5).normal(0, 0.01, (V, D)).astype(np.float32)
    srv = port.Server(V, D, rows)
    caches = [port.Cache(srv, policy, limit, bound) for _ in range(world)]
    digest = []

    def batch(w, t):
        rng = np.random.default_rng(1000 * t + w)
        n = int(rng.integers(1, 200))
        return (((rng.zipf(1.2, n) - 1) % V).astype(np.uint64),
                rng.normal(0, 1e-3, (n, D)).astype(np.float32))

    for w in range(world):
        digest.append(caches[w].embedding_lookup(batch(w, 0)[0]).tobytes())
    for t in range(steps):
        for w in range(world):
            k, g = batch(w, t)
            caches[w].embedding_update(k, g)
        for w in range(world):
            digest.append(caches[w].embedding_lookup(batch(w, t + 1)[0]).tobytes())
    return srv.rows(), srv.versions(), digest


def _rank_main(rank, world, port_no, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        V, D = 503, 8
        rows, vers, digest = _replay(world, "lru", 2, V, D, 60, 12)
        begin, n = partition.shard_range(rank, world, V)
        mine = (rows[begin:begin + n].tobytes(), vers[begin:begin + n].tobytes(),
                hashlib.sha256(b"".join(digest)).hexdigest())
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        # both ranks computed the same schedule: the lookups agree ...
        assert len({g[2] for g in gathered}) == 1
        # ... and the shards tile the table rank 0 computed
        if rank == 0:
            assert b"".join(g[0] for g in gathered) == rows.tobytes()
            assert b"".join(g[1] for g in gathered) == vers.tobytes()
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover - surfaced by the parent
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_bsp_replay_is_deterministic_across_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port_no = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_rank_main, args=(r, world, port_no, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def _laia_rank_main(rank, world, port_no, out):
    """Every rank runs its own planner (as every Hetu worker does) and pops its own part; together
    the parts must tile every global batch, and a rank's plan must be what its peers computed for
    it (the planner is replicated, not distributed: laia_scheduler.cc:138-139)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from herald_b200.laia import LaiaScheduler
        rng = np.random.default_rng(21)
        mini, T, nb = 6, 5, 4
        emb = ((rng.zipf(1.2, (world * mini * nb, T)) - 1) % 70 + 1).astype(np.uint64)
        s = LaiaScheduler()
        s.start(emb, emb.shape[0], T, 1, mini, nb, world, rank, 25, 2)
        b = 0
        while s.step():
            mine = (s.plan_of(rank).tolist(), s.dist_of(rank).tolist(),
                    [s.plan_of(w).tolist() for w in range(world)])
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
            start = (b * world * mini) % emb.shape[0]
            covered = sorted(p for g in gathered for p in g[1])
            assert covered == list(range(start, start + world * mini)), (b, covered)
            for w in range(world):
                assert gathered[w][0] == mine[2][w], (b, w)           # peers agree on w's plan
            b += 1
        assert b == nb + 1
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover - surfaced by the parent
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_laia_planner_parts_tile_the_batch_across_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port_no = 29800 + os.getpid() % 90
    procs = [ctx.Process(target=_laia_rank_main, args=(r, world, port_no, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
