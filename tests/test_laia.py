"""Laia / Herald embedding scheduler: the C++ planner in
libherald_b200.so (csrc/hb_laia.cu, through herald_b200.laia) against
  * the golden vectors produced by the reference's own planner (python/hetu/laia/laia.pyx,
    tests/golden/make_golden_laia.py),
  * the pure-Python restatement oracle/laia_port.py (itself checked against the same vectors),
  * the reference planner run live when oracle/_ref/laia*.so is present.
Integer work: every plan and distribution must be identical.  Host code: no GPU needed."""
import os

import numpy as np
import pytest

from herald_b200.laia import LaiaScheduler, MiniLRUCache
from oracle import laia_port, laia_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "laia_cases.npz")


def golden_cases():
    z = np.load(GOLDEN)
    for c in range(int(z["ncases"])):
        W, mini, T, nb, cap, ep = (int(x) for x in z["c%d_params" % c])
        emb, plan, off, dist = z["c%d_emb" % c], z["c%d_plan" % c], z["c%d_plan_off" % c], z["c%d_dist" % c]
        batches = []
        for b in range(dist.shape[0]):
            plans = [plan[off[b * W + w]:off[b * W + w + 1]].tolist() for w in range(W)]
            batches.append((plans, dist[b].tolist()))
        yield c, dict(W=W, mini=mini, T=T, nb=nb, cap=cap, ep=ep, emb=emb), batches


CASES = list(golden_cases())


def run_ours(p, threads=3):
    """Every rank runs its own scheduler and reports its own part -> per batch (plans, dist)."""
    S = p["emb"].shape[0]
    scheds = []
    for rank in range(p["W"]):
        s = LaiaScheduler()
        s.start(p["emb"], S, p["T"], p["ep"], p["mini"], p["nb"], p["W"], rank, p["cap"], threads)
        scheds.append(s)
    out = []
    while True:
        more = [s.step() for s in scheds]
        assert all(m == more[0] for m in more)
        if not more[0]:
            break
        # (not through pop(): a plan that is exactly [0] reads as the terminator on the reference's
        # wire, laia_dataloader.py:137-139)
        out.append(([s.plan_of(s.rank).tolist() for s in scheds], [s.dist_of(s.rank).tolist() for s in scheds]))
    return out, scheds


@pytest.mark.parametrize("case", [c[0] for c in CASES])
def test_port_matches_reference_golden(case):
    _, p, expected = CASES[case]
    planner = laia_port.LaiaPlanner(p["emb"], p["mini"], p["W"], p["cap"], p["ep"], p["nb"])
    for b, (plans, dist) in enumerate(expected):
        got = planner.next_all()
        assert got is not None, "port stopped early at batch %d" % b
        assert got[0] == plans, "communication plan, batch %d" % b
        assert got[1] == dist, "sample distribution, batch %d" % b
    assert planner.next_all() is None


@pytest.mark.parametrize("case", [c[0] for c in CASES])
def test_planner_matches_reference_golden(case):
    _, p, expected = CASES[case]
    got, scheds = run_ours(p)
    assert len(got) == len(expected)
    for b, ((plans, dist), (eplans, edist)) in enumerate(zip(got, expected)):
        assert plans == eplans, "communication plan, batch %d" % b
        assert dist == edist, "sample distribution, batch %d" % b
    # every rank plans the whole group: any rank can report any worker's part
    assert scheds[0].plan_of(p["W"] - 1).tolist() == expected[-1][0][p["W"] - 1]


@pytest.mark.parametrize("seed", range(8))
def test_planner_matches_port_random(seed):
    rng = np.random.default_rng(seed)
    W, mini, T = int(rng.integers(1, 7)), int(rng.integers(1, 10)), int(rng.integers(1, 8))
    nb, cap, ep = int(rng.integers(1, 5)), int(rng.integers(0, 50)), int(rng.integers(1, 4))
    S = W * mini * nb + int(rng.integers(0, 3)) * W * mini        # more samples than one epoch uses
    emb = ((rng.zipf(1.3, (S, T)) - 1) % 90).astype(np.uint64)
    p = dict(W=W, mini=mini, T=T, nb=nb, cap=cap, ep=ep, emb=emb)
    got, scheds = run_ours(p, threads=int(rng.integers(1, 5)))
    planner = laia_port.LaiaPlanner(emb, mini, W, cap, ep, nb)
    for b, (plans, dist) in enumerate(got):
        eplans, edist = planner.next_all()
        assert plans == eplans and dist == edist, "batch %d" % b
    assert planner.next_all() is None
    for w in range(W):                                            # simulated caches agree at the end
        assert scheds[0].snapshot_keys(w).tolist() == planner.snaps[w].get_keys()


@pytest.mark.skipif(not laia_ref.available(), reason="oracle/_ref/laia*.so not built")
@pytest.mark.parametrize("seed", range(4))
def test_planner_matches_live_reference(seed):
    rng = np.random.default_rng(50 + seed)
    W, mini, T = int(rng.integers(1, 6)), int(rng.integers(2, 12)), int(rng.integers(1, 6))
    while W * mini < 8:
        mini += 1
    nb, cap, ep = int(rng.integers(1, 5)), int(rng.integers(1, 60)), int(rng.integers(1, 3))
    emb = ((rng.zipf(1.3, (W * mini * nb, T)) - 1) % 80).astype(np.int32)
    expected = laia_ref.run(emb, ep, mini, nb, W, cap)
    got, _ = run_ours(dict(W=W, mini=mini, T=T, nb=nb, cap=cap, ep=ep, emb=emb))
    assert [(pl, d) for pl, d in got] == [(pl, d) for pl, d in expected]


def test_wire_format():
    """laia_dataloader.py:122-143: plan then indices per batch, [0] at the end, length() >= 2
    while batches remain."""
    emb = np.arange(48, dtype=np.uint64).reshape(16, 3) % 7
    s = LaiaScheduler()
    s.start(emb, 16, 3, 1, 4, 1, 2, 1, 5)
    assert s.length() >= 2
    msgs = []
    while True:
        m = s.pop()
        assert isinstance(m, list)
        msgs.append(m)
        if m == [0]:
            break
    assert len(msgs) == 2 * 2 + 1                                 # batch_num + 1 batches in the last epoch
    assert all(len(msgs[i]) == 4 for i in (1, 3))
    assert s.length() == 0
    with pytest.raises(RuntimeError):
        s.pop()
    with pytest.raises(RuntimeError):
        LaiaScheduler().start(np.zeros(4, np.uint64), 4, 1, 1, 1, 1, 1, 0, 1)


def test_mini_lru_return_codes():
    """laia/include/mini_lru_cache.h:69-105: -1 hit, -2 stale hit, 0 miss, 1 miss + valid victim."""
    for impl in (MiniLRUCache, laia_port.MiniLRUCache):
        c = impl(2)
        assert c.get(5) == 0 and c.get(7) == 0                    # misses, room left
        assert c.get(5) == -1                                     # hit
        c.outdate(7)
        assert not c.check(7) and c.check(5)
        assert c.get(7) == -2                                     # stale hit: valid again, most recent
        assert c.check(7)
        assert c.get(9) == 1                                      # evicts 5 (valid)
        assert list(c.get_keys()) == [7, 9]
        c.outdate(7)
        assert c.get(11) == 0                                     # evicts 7 (outdated)
        assert list(c.get_keys()) == [9, 11]
        c.evict(9)
        assert list(c.get_keys()) == [11] and not c.check(9)
        z = impl(0)                                               # capacity 0: nothing stays
        assert z.get(3) == 1 and not z.check(3)


def test_mini_lru_random_against_port():
    rng = np.random.default_rng(3)
    a, b = MiniLRUCache(17), laia_port.MiniLRUCache(17)
    for _ in range(5000):
        k, op = int(rng.integers(0, 60)), int(rng.integers(0, 10))
        if op < 6:
            assert a.get(k) == b.get(k)
        elif op < 8:
            a.outdate(k), b.outdate(k)
        elif op < 9:
            a.evict(k), b.evict(k)
        else:
            assert a.check(k) == b.check(k)
    assert list(a.get_keys()) == b.get_keys()


def test_laia_front_end_pairs_indices_with_the_next_plan():
    """laia_dataloader.py:108-114, 150-169: batch b = (sample indices of b, plan computed for b + 1);
    five batches planned ahead; stepping every dataset forward refills the oldest slot."""
    from herald_b200.laia import LAIAScheduler

    class Config(object):
        nrank, rank, local_rank, cache_limit = 2, 1, 1, 30

    rng = np.random.default_rng(11)
    data = ((rng.zipf(1.3, (160, 4)) - 1) % 50 + 1).astype(np.float32)  # ids >= 1: a plan never reads as [0]
    front = LAIAScheduler(data, batch_size=8)
    front.start(Config, dataset_num=1, epoch_num=1)
    assert front.batch_num == 10 and front.queue_size == 5
    planner = laia_port.LaiaPlanner(data.astype(np.int64), 8, 2, 30, 1, 10)
    seq = []
    while True:
        r = planner.next_all()
        if r is None:
            break
        seq.append((r[0][1], r[1][1]))                                # rank 1: (plan, dist)
    for b in range(8):
        assert front.get_input_index(b % front.batch_num) == seq[b][1], b
        assert front.get_comm_plan(b % front.batch_num) == seq[b + 1][0], b
        front.step_forward(0)


def test_plans_drive_the_cache_update_of_every_worker():
    """run_laia's loop on the CPU oracle (one server, one cache per worker): worker w looks up the
    samples the planner gave it, trains, and updates with the plan of the NEXT batch as `push_keys`
    (laia_dataloader.py:108-114 -> cstable.py:64-80 -> cache.cc:248-334).  Checks the contract
    between planner and cache: a plan is ascending unique keys the cache accepts, and the lines it
    names are exactly the ones pushed (plus dataless lines, which are pushed regardless)."""
    from oracle import port
    rng = np.random.default_rng(4)
    W, mini, T, nb, V, D, cap = 3, 16, 6, 5, 400, 8, 60
    emb = ((rng.zipf(1.15, (W * mini * nb, T)) - 1) % V).astype(np.uint64)
    scheds = []
    for w in range(W):
        s = LaiaScheduler()
        s.start(emb, emb.shape[0], T, 1, mini, nb, W, w, cap, 2)
        scheds.append(s)
    srv = port.Server(V, D, rng.normal(0, 0.01, (V, D)).astype(np.float32))
    caches = [port.Cache(srv, "lru", cap, 10) for _ in range(W)]
    assert all(s.step() for s in scheds)
    dist = [s.dist_of(w) for w, s in enumerate(scheds)]
    for b in range(nb):
        keys = [emb[dist[w].astype(np.int64)].reshape(-1) for w in range(W)]
        for w in range(W):
            caches[w].embedding_lookup(keys[w])
        more = [s.step() for s in scheds]                             # plans for batch b + 1
        assert all(more)
        for w in range(W):
            plan = scheds[w].plan_of(w)
            assert np.all(np.diff(plan.astype(np.int64)) > 0)         # ascending, unique
            grads = rng.normal(0, 1e-3, (keys[w].size, D)).astype(np.float32)
            perf = caches[w].embedding_update(keys[w], grads, plan)
            touched = np.unique(keys[w])
            expect = np.intersect1d(touched, plan).size
            # pushed = planned lines that this batch touched (+ dataless lines of update misses)
            assert perf["num_transfered"] - perf["num_evict"] >= expect - perf["num_miss"]
            assert perf["num_transfered"] - perf["num_evict"] <= expect + perf["num_miss"]
        dist = [s.dist_of(w) for w, s in enumerate(scheds)]


@pytest.mark.parametrize("W,offset,n_keys,mini", [(3, 1 << 33, 5000, 900), (65, 0, 90, 40),
                                                   (66, 1 << 40, 3000, 40)])
def test_planner_paths_without_bitmap_or_holder_bits(W, offset, n_keys, mini):
    """Ids above 2^31 take the radix sort instead of the id bitmap, more than 64 workers probe the
    snapshots again instead of reading holder bits: same plans as the port."""
    rng = np.random.default_rng(W)
    T, nb, cap = 5, 3, 2500 if mini > 100 else 700        # the 900-sample case builds plans of > 2048 entries
    emb = (((rng.zipf(1.2, (W * mini * nb, T)) - 1) % n_keys) + offset).astype(np.uint64)
    p = dict(W=W, mini=mini, T=T, nb=nb, cap=cap, ep=1, emb=emb)
    got, scheds = run_ours(p, threads=4)
    planner = laia_port.LaiaPlanner(emb, mini, W, cap, 1, nb)
    for b, (plans, dist) in enumerate(got):
        eplans, edist = planner.next_all()
        assert plans == eplans and dist == edist, "batch %d" % b
    assert planner.next_all() is None
    assert scheds[0].snapshot_keys(W - 1).tolist() == planner.snaps[W - 1].get_keys()
