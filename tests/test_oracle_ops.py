"""Pins oracle/ops_port.py (numpy restatements of the reference CUDA kernels) against the numpy
references the REFERENCE's own tests use — no GPU needed."""
import numpy as np

from oracle import ops_port


def test_adamw_sparse_against_reference_test_formula():
    # reference tests/test_optimizer.py:117-197 (test_adamw_sparse): dict-based dedup + formula :175-181
    rng = np.random.default_rng(0)
    V, D, n = 500, 400, 100
    param = rng.uniform(-10, 10, size=(V, D)).astype(np.float32)
    idx = rng.integers(0, V, n)
    g = rng.uniform(-10, 10, size=(n, D)).astype(np.float32)
    m = rng.uniform(0, 10, size=(V, D)).astype(np.float32)
    v = rng.uniform(0, 10, size=(V, D)).astype(np.float32)
    lr, b1, b2, eps, wd = 1e-2, 0.9, 0.99, 1e-7, 0.1
    b1t, b2t = b1 ** 10, b2 ** 10
    # their reference: accumulate duplicates in a dict, then per unique row
    acc = {}
    for i, k in enumerate(idx):
        acc[k] = acc.get(k, np.zeros(D, np.float32)) + g[i]
    ep, em, ev = param.copy(), m.copy(), v.copy()
    for k, gk in acc.items():
        em[k] = b1 * em[k] + (1 - b1) * gk
        ev[k] = b2 * ev[k] + (1 - b2) * gk * gk
        mc, vc = em[k] / (1 - b1t), ev[k] / (1 - b2t)
        ep[k] = ep[k] - lr * (mc / (np.sqrt(vc) + eps) + wd * ep[k])
    uniq, inv = ops_port.unique_inverse(idx.astype(np.float32))
    cg = ops_port.deduplicate(g, inv, len(uniq))
    p2, m2, v2 = ops_port.adamw_sparse_update(param, uniq, cg, m, v, lr, b1, b2, b1t, b2t, eps, wd)
    np.testing.assert_allclose(p2, ep, atol=1e-5)
    np.testing.assert_allclose(m2, em, atol=1e-5)
    np.testing.assert_allclose(v2, ev, atol=1e-5)


def test_lamb_sparse_against_reference_test_formula():
    # reference tests/test_optimizer.py:200-298 (test_lamb_sparse): dict-based dedup, Adam
    # direction :262-268, norms over the indexed rows only :275-276, step :281, atol 1e-5
    rng = np.random.default_rng(3)
    V, D, n = 500, 400, 100
    param = rng.uniform(-10, 10, size=(V, D)).astype(np.float32)
    idx = rng.integers(0, V, n)
    g = rng.uniform(-10, 10, size=(n, D)).astype(np.float32)
    m = rng.uniform(-10, 10, size=(V, D)).astype(np.float32)
    v = rng.uniform(0, 10, size=(V, D)).astype(np.float32)
    lr, b1, b2, eps, wd = 1e-2, 0.9, 0.99, 1e-7, 0.1
    b1t, b2t = b1 ** 10, b2 ** 10
    acc = {}
    for i, k in enumerate(idx):
        acc[k] = acc.get(k, np.zeros(D, np.float32)) + g[i]
    ep, em, ev = param.copy(), m.copy(), v.copy()
    ups = []
    for k, gk in acc.items():
        em[k] = b1 * em[k] + (1 - b1) * gk
        ev[k] = b2 * ev[k] + (1 - b2) * gk * gk
        ups.append((em[k] / (1 - b1t)) / (np.sqrt(ev[k] / (1 - b2t)) + eps))
    ups = np.array(ups)
    norm_p = np.sqrt(np.sum(np.power(np.array([ep[k] for k in acc]), 2)))
    norm_u = np.sqrt(np.sum(np.power(ups, 2)))
    for k, u in zip(acc, ups):
        ep[k] = ep[k] - lr * norm_p / norm_u * (u + wd * ep[k])
    uniq, inv = ops_port.unique_inverse(idx.astype(np.float32))
    cg = ops_port.deduplicate(g, inv, len(uniq))
    p2, m2, v2 = ops_port.lamb_sparse_update(param, uniq, cg, m, v, lr, b1, b2, b1t, b2t, eps, wd)
    np.testing.assert_allclose(p2, ep, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(m2, em, atol=1e-5)
    np.testing.assert_allclose(v2, ev, atol=1e-5)


def test_toy_embedding_sgd_dup_ids():
    # reference tests/test_embedding_op.py:25-89: 5x5 table, ids [[0,1],[0,1]]; duplicate-id SGD
    rng = np.random.default_rng(1)
    table = rng.normal(size=(5, 5)).astype(np.float32)
    ids = np.array([[0, 1], [0, 1]], np.float32)
    cur, lr = table.copy(), 0.1
    for _ in range(100):
        looked = ops_port.embedding_lookup(cur, ids)
        assert looked.shape == (2, 2, 5)
        cur = ops_port.sgd_sparse_update(cur, ids, looked * np.float32(0.5), lr)
    expect = table.astype(np.float64)
    expect[:2] *= (1 - 2 * lr * 0.5) ** 100       # each of rows 0,1 is hit twice per step
    np.testing.assert_allclose(cur, expect, rtol=1e-5)
    np.testing.assert_array_equal(cur[2:], table[2:])


def test_additive_push_semantics():
    # reference tests/pstests/test_apis.py:105-156: sparse push adds, duplicates accumulate
    base = np.zeros((10, 3), np.float32)
    ids = np.array([1, 1, 4], np.float32)
    vals = np.ones((3, 3), np.float32)
    out = ops_port.indexedslices_oneside_add(ids, vals, base)
    assert out[1, 0] == 2 and out[4, 0] == 1 and out.sum() == 9


def test_dedup_is_segment_sum_in_order():
    rng = np.random.default_rng(2)
    ids = rng.integers(0, 7, 50).astype(np.float32)
    vals = rng.normal(size=(50, 4)).astype(np.float32)
    uniq, inv = ops_port.unique_inverse(ids)
    assert np.array_equal(uniq, np.unique(ids))
    out = ops_port.deduplicate(vals, inv, len(uniq))
    for u, k in enumerate(uniq):
        acc = np.zeros(4, np.float32)
        for row in vals[ids == k]:
            acc = acc + row
        assert np.array_equal(out[u], acc)
