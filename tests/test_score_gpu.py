"""Scoring and residency probes against the real cache index (hb_cache_score / hb_cache_probe,
SURVEY 8 f-1) vs the oracle cache's resident key set; then one GpuScoredPlanner batch."""
import numpy as np
import pytest

from common import GpuHarness, zipf_keys

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("policy,bound", [("lru", 0), ("lfu", 3)])
def test_scores_and_probes_match_the_oracle_cache(oracle_impl, policy, bound):
    from herald_b200.laia import TOPK_TABLE_ORDER
    rng = np.random.default_rng(3)
    V, D, T = 3000, 16, 26
    h = GpuHarness(oracle_impl, policy, 400, bound, rng.normal(0, 0.01, (V, D)).astype(np.float32))
    try:
        for t in range(6):
            keys = zipf_keys(rng, 900, V, 1.1)
            h.lookup(keys)
            h.update(keys, rng.normal(0, 1e-3, (len(keys), D)).astype(np.float32))
        samples = ((rng.zipf(1.1, (257, T)) - 1) % V).astype(np.uint64)
        resident = np.isin(samples, h.oc.keys())
        order = np.asarray(TOPK_TABLE_ORDER["criteo"], np.uint32)
        for top_k in (26, 20, 1):
            got = h.gc.score_samples(samples, order, top_k)
            assert np.array_equal(got, resident[:, order[:top_k]].sum(1).astype(np.uint32)), top_k
        assert np.array_equal(h.gc.score_samples(samples.astype(np.float32), None, 7),
                              resident[:, :7].sum(1).astype(np.uint32))            # float32-carried ids
        assert np.array_equal(h.gc.resident(samples), resident)
        # fresh scoring never counts more than plain residency, and with nothing stale it counts the same
        fresh = h.gc.score_samples(samples, order, 26, fresh=True)
        assert np.all(fresh <= resident.sum(1))
        h.check_state("scoring is read-only")
    finally:
        h.close()


def test_gpu_scored_planner_single_worker(oracle_impl):
    from herald_b200.laia import GpuScoredPlanner
    rng = np.random.default_rng(4)
    V, D, T, mini = 2000, 8, 26, 16
    h = GpuHarness(oracle_impl, "lru", 300, 0, rng.normal(0, 0.01, (V, D)).astype(np.float32))
    try:
        embs = ((rng.zipf(1.15, (mini * 6, T)) - 1) % V).astype(np.uint64)
        h.lookup(embs[:mini].reshape(-1))
        planner = GpuScoredPlanner(h.gc, embs, mini, 1, 0)
        for b in range(3):
            pos, plan = planner.plan_batch(b)
            assert sorted(pos.tolist()) == list(range(b * mini, (b + 1) * mini))
            keys = np.unique(embs[pos].reshape(-1))
            exp = keys[np.isin(keys, h.oc.keys())]
            assert np.array_equal(plan, exp)
            h.lookup(embs[pos].reshape(-1))
    finally:
        h.close()
