"""TopkScheduler — the reference's production planner (laia/src/topk_scheduler.cc:259-502, the one
run_laia.py --local-shared starts) — in libherald_b200.so (csrc/hb_laia.cu, herald_b200.laia)
against
  * golden vectors produced by the reference's own TopkScheduler (tests/golden/
    make_golden_laia_topk.py; oracle/_ref/laia_cache*.so = the reference sources unmodified),
  * the reference run live on random cases when that oracle is built,
and the simple planner against the reference's C++ LaiaScheduler (the Cython laia.pyx is the
oracle of tests/test_laia.py).  Integer work: plans and distributions must be identical.  The
shared-memory rings of the local-shared mode are exercised across processes.  Host code, no GPU."""
import multiprocessing as mp
import os

import numpy as np
import pytest

from herald_b200.laia import LaiaScheduler, TopkScheduler, LAIAScheduler, _ShmRing
from oracle import laia_cpp_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "laia_topk_cases.npz")


def golden_cases():
    z = np.load(GOLDEN)
    for c in range(int(z["ncases"])):
        W, mini, T, nb, cap, ep, th, topk = (int(x) for x in z["c%d_params" % c])
        emb, plan, off, dist = z["c%d_emb" % c], z["c%d_plan" % c], z["c%d_plan_off" % c], z["c%d_dist" % c]
        batches = []
        for b in range(dist.shape[0]):
            plans = [plan[off[b * W + w]:off[b * W + w + 1]].tolist() for w in range(W)]
            batches.append((plans, dist[b].tolist()))
        yield c, dict(W=W, mini=mini, T=T, nb=nb, cap=cap, ep=ep, th=th, topk=topk,
                      ds=str(z["c%d_dataset" % c]), emb=emb), batches


CASES = list(golden_cases())


def run_ours(p):
    S = p["emb"].shape[0]
    scheds = []
    for rank in range(p["W"]):
        s = TopkScheduler()
        s.start(p["emb"], S, p["T"], p["ep"], p["mini"], p["nb"], p["W"], rank, p["cap"], p["th"], p["ds"],
                p["topk"])
        scheds.append(s)
    out = []
    while True:
        more = [s.step() for s in scheds]
        assert all(m == more[0] for m in more)
        if not more[0]:
            break
        out.append(([s.plan_of(s.rank).tolist() for s in scheds], [s.dist_of(s.rank).tolist() for s in scheds]))
    return out


@pytest.mark.parametrize("case", [c[0] for c in CASES])
def test_topk_matches_reference_golden(case):
    _, p, expected = CASES[case]
    got = run_ours(p)
    assert len(got) == len(expected)
    for b, ((plans, dist), (eplans, edist)) in enumerate(zip(got, expected)):
        assert dist == edist, "sample distribution, batch %d" % b
        assert plans == eplans, "communication plan, batch %d" % b


def test_golden_holds_the_erase_while_iterating_result():
    """The reference erases from the plan set inside the loop over it (topk_scheduler.cc:478-482): with
    empty snapshots every other id of the first batch survives (less what the stale tail cells
    re-examine) — about half, not none."""
    _, p, expected = CASES[1]
    first_plans = expected[0][0]
    uniq = [len(np.unique(p["emb"][d].reshape(-1))) for d in expected[0][1]]
    for plan, u in zip(first_plans, uniq):
        assert 0.4 * u < len(plan) < 0.6 * u, (len(plan), u)


needs_ref = pytest.mark.skipif(not laia_cpp_ref.available(), reason="oracle/_ref/laia_cache*.so not built")


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_topk_matches_live_reference_random(seed):
    rng = np.random.default_rng(1000 + seed)
    W = int(rng.integers(1, 7))
    th = int(rng.choice([1, 2, 4]))
    mini = th * int(rng.integers(2, 9))     # (a one-sample distribution [0] would read as the wire's end)
    nb, ep = int(rng.integers(2, 6)), int(rng.integers(1, 3))
    ds, T = [("criteo", 26), ("avazu", 18), ("criteosearch", 17)][seed % 3]
    topk = int(rng.integers(0, T + 1))
    cap = int(rng.integers(5, 400))
    vocab = int(rng.integers(20, 3000))
    emb = ((rng.zipf(1.2, (W * mini * nb, T)) - 1) % vocab + 1).astype(np.int64)
    p = dict(W=W, mini=mini, T=T, nb=nb, cap=cap, ep=ep, th=th, topk=topk, ds=ds, emb=emb)
    expected = laia_cpp_ref.run_topk(emb, ep, mini, nb, W, cap, th, ds, topk)
    got = run_ours(p)
    assert len(got) == len(expected)
    for b, ((plans, dist), (eplans, edist)) in enumerate(zip(got, expected)):
        assert dist == edist, (seed, b)
        assert plans == eplans, (seed, b)


@needs_ref
@pytest.mark.parametrize("seed", range(4))
def test_simple_planner_matches_reference_cpp_laia_scheduler(seed):
    """herald_b200's LaiaScheduler against laia/src/laia_scheduler.cc itself (tests/test_laia.py pins
    it to the older Cython laia.pyx)."""
    rng = np.random.default_rng(2000 + seed)
    W, mini, nb, T = int(rng.integers(1, 6)), int(rng.integers(2, 20)), int(rng.integers(2, 5)), 26
    cap, vocab = int(rng.integers(10, 500)), int(rng.integers(50, 4000))
    emb = ((rng.zipf(1.15, (W * mini * nb, T)) - 1) % vocab + 1).astype(np.int64)
    expected = laia_cpp_ref.run_laia(emb, 1, mini, nb, W, cap)
    scheds = []
    for rank in range(W):
        s = LaiaScheduler()
        s.start(emb, emb.shape[0], T, 1, mini, nb, W, rank, cap, 3)
        scheds.append(s)
    for b, (eplans, edist) in enumerate(expected):
        assert all(s.step() for s in scheds)
        assert [s.dist_of(s.rank).tolist() for s in scheds] == edist, (seed, b)
        assert [s.plan_of(s.rank).tolist() for s in scheds] == eplans, (seed, b)
    assert not any(s.step() for s in scheds)


def test_unsupported_dataset_and_slot_split_are_rejected():
    emb = np.ones((64, 26), np.int64)
    with pytest.raises(Exception, match="dataset not supported"):
        TopkScheduler().start(emb, 64, 26, 1, 4, 4, 2, 0, 10, 2, "imagenet", 5)
    with pytest.raises(Exception, match="multiple of num_threads"):
        TopkScheduler().start(emb, 64, 26, 1, 5, 4, 2, 0, 10, 2, "criteo", 5)


def _ring_reader(name, n_msgs, q):
    r = _ShmRing(name, False)
    q.put([r.recv() for _ in range(n_msgs)])
    r.close()


def test_shm_ring_between_processes():
    name = "hb_test_ring_%d" % os.getpid()
    ring = _ShmRing(name, True, 4096)            # 512 words: the writer has to wait for the reader
    msgs = [list(range(i, i + 1 + (37 * i) % 300)) for i in range(40)] + [[0]]
    q = mp.get_context("spawn").Queue()
    proc = mp.get_context("spawn").Process(target=_ring_reader, args=(name, len(msgs), q), daemon=True)
    proc.start()
    for m in msgs:
        ring.send(m)
    got = q.get(timeout=60)
    proc.join(timeout=30)
    ring.close()
    assert got == msgs


class _Cfg(object):
    def __init__(self, rank, nrank, local_rank, limit, threads, local_size):
        self.rank, self.nrank, self.local_rank, self.cache_limit = rank, nrank, local_rank, limit
        self.laia_threads, self.local_size = threads, local_size


def _local_worker(local_rank, sparse, batch, limit, q):
    sched = LAIAScheduler(sparse, batch, dataset="criteo", local_shared=True)
    sched.start(_Cfg(0 if local_rank == 0 else local_rank, 2, local_rank, limit, 2, 2), dataset_num=1,
                epoch_num=1)
    out = []
    for b in range(3):
        out.append((list(sched.get_input_index(b)), list(sched.get_comm_plan(b))))
        sched.step_forward(0)
    q.put((local_rank, out))


def test_local_shared_front_end_two_processes():
    """LAIAScheduler(local_shared=True): local rank 0 plans for both local workers and feeds worker 1
    through its ring (laia_dataloader.py:72-96).  Each worker must see what the standalone
    TopkScheduler of its rank computes: indices of batch b with the plan of batch b + 1."""
    rng = np.random.default_rng(5)
    T, nrank, batch = 26, 2, 4
    S = nrank * batch * 12
    sparse = ((rng.zipf(1.2, (S, T)) - 1) % 200 + 1).astype(np.float32)
    limit = 60
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_local_worker, args=(lr, sparse, batch, limit, q), daemon=True) for lr in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    emb = sparse.astype(np.intc).astype(np.int64)
    samples_num = S // nrank
    bsz = min(batch, samples_num // 5)
    nb = samples_num // bsz
    for rank in range(2):
        s = TopkScheduler()
        s.start(emb, S, T, 1, bsz, nb, nrank, rank, limit, 2, "criteo", 20)
        seq = []
        while s.step():
            seq.append((s.plan_of(rank).tolist(), s.dist_of(rank).tolist()))
        for b in range(3):
            idx, plan = results[rank][b]
            assert idx == seq[b][1], (rank, b)
            assert plan == seq[b + 1][0], (rank, b)


def test_assign_by_scores_respects_slots_and_prefers_the_best_worker():
    from herald_b200.laia import assign_by_scores
    rng = np.random.default_rng(9)
    W, mini, parts = 4, 8, 2
    scores = rng.integers(0, 6, (W, W * mini))
    dist = assign_by_scores(scores, mini, parts)
    assert sorted(dist.reshape(-1).tolist()) == list(range(W * mini))       # every sample exactly once
    per, room = W * mini // parts, mini // parts
    for t in range(parts):                                                   # slot runs hold the thread's samples
        got = dist[:, t * room:(t + 1) * room].reshape(-1)
        assert set(got.tolist()) == set(range(t * per, (t + 1) * per))
    first = int(dist[int(np.argmax(scores[:, 0])), 0])
    assert first == 0                                                        # sample 0: best worker, first slot
    assert np.array_equal(dist, assign_by_scores(scores, mini, parts))
