"""GPU parity of the op-level entry points (reference names, float32 ids) against
oracle/ops_port.py.  Integer/index work and pure row moves are bit-exact; the order-fixed adds
(SGD with duplicate ids, dedup, one-side add) are bit-exact too; Adam-family updates compare at
rtol 1e-5 (north_star tolerance for updated rows; FMA contraction differs between nvcc and numpy).
"""
import numpy as np
import pytest

from common import assert_bits_equal

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # BASELINE.json north_star: "updated embedding rows within 1e-5 relative in fp32"


@pytest.fixture(scope="module")
def hb():
    import herald_b200
    return herald_b200


def _dev(hb, a):
    return hb.array(np.ascontiguousarray(a, np.float32), hb.gpu(0))


@pytest.mark.parametrize("V,D,shape", [(50, 8, (7,)), (1000, 128, (64, 26)), (333, 6, (5, 3)),
                                       (2000, 512, (300,))])
def test_embedding_lookup(hb, V, D, shape):
    from herald_b200 import gpu_links
    from oracle import ops_port
    rng = np.random.default_rng(0)
    table = rng.normal(size=(V, D)).astype(np.float32)
    ids = rng.integers(0, V, size=shape).astype(np.float32)
    out = hb.empty(shape + (D,), hb.gpu(0))
    gpu_links.embedding_lookup(_dev(hb, table), _dev(hb, ids), out)
    assert_bits_equal(out.asnumpy(), ops_port.embedding_lookup(table, ids), "gather")


def test_embedding_lookup_with_stream(hb):
    from herald_b200 import gpu_links, stream
    from oracle import ops_port
    rng = np.random.default_rng(1)
    table = rng.normal(size=(100, 128)).astype(np.float32)
    ids = rng.integers(0, 100, size=(40,)).astype(np.float32)
    st = stream.create_stream_handle(hb.gpu(0))
    out = hb.empty((40, 128), hb.gpu(0))
    gpu_links.embedding_lookup(_dev(hb, table), _dev(hb, ids), out, st)
    st.sync()
    assert_bits_equal(out.asnumpy(), ops_port.embedding_lookup(table, ids), "gather on a stream")


@pytest.mark.parametrize("n,V", [(1, 10), (100, 20), (5000, 300), (20000, 1 << 20)])
def test_indexedslices_deduplicate_matches_np_unique(hb, n, V):
    from oracle import ops_port
    rng = np.random.default_rng(2)
    D = 16
    ids = ((rng.zipf(1.2, n) - 1) % V).astype(np.float32)
    vals = rng.normal(size=(n, D)).astype(np.float32)
    sl = hb.IndexedSlices(indices=_dev(hb, ids), values=_dev(hb, vals), dense_shape=(V, D))
    sl.deduplicate(None)
    uniq, inv = ops_port.unique_inverse(ids)
    assert np.array_equal(sl.indices.asnumpy(), uniq)           # ascending unique ids, bit-exact
    assert_bits_equal(sl.values.asnumpy(), ops_port.deduplicate(vals, inv, len(uniq)), "dedup")


def test_unique_inverse_bit_exact(hb):
    import ctypes
    from herald_b200._base import _LIB, check_call
    from oracle import ops_port
    rng = np.random.default_rng(3)
    for n in (1, 31, 32, 33, 1023, 1024, 1025, 2049, 6656, 50000):
        ids = ((rng.zipf(1.05, n) - 1) % 33762577).astype(np.float32)
        d_ids = _dev(hb, ids)
        uq, iv, cnt = hb.empty((n,), hb.gpu(0)), hb.empty((n,), hb.gpu(0)), hb.empty((2,), hb.gpu(0))
        check_call(_LIB.HBUniqueIndexedSlices(d_ids.handle, uq.handle, iv.handle,
                                              ctypes.c_void_p(cnt.data_ptr), None))
        U = int(cnt.asnumpy().view(np.int64)[0])
        uniq, inv = ops_port.unique_inverse(ids)
        assert U == len(uniq), n
        assert np.array_equal(uq.asnumpy()[:U], uniq), n
        assert np.array_equal(iv.asnumpy().astype(np.int64), inv), n


@pytest.mark.parametrize("D", [8, 128])
@pytest.mark.parametrize("hot_threshold", [64, 3])
def test_sgd_sparse_update_duplicate_ids(hb, D, hot_threshold):
    from herald_b200 import gpu_links
    from herald_b200._base import set_hot_threshold
    from oracle import ops_port
    set_hot_threshold(hot_threshold)
    rng = np.random.default_rng(4)
    V, n = 60, 500
    param = rng.normal(size=(V, D)).astype(np.float32)
    ids = ((rng.zipf(1.3, n) - 1) % V).astype(np.float32)
    g = rng.normal(size=(n, D)).astype(np.float32)
    p = _dev(hb, param)
    gpu_links.sgd_update(p, hb.IndexedSlices(_dev(hb, ids), _dev(hb, g), (V, D)), 0.01)
    set_hot_threshold(64)
    assert_bits_equal(p.asnumpy(), ops_port.sgd_sparse_update(param, ids, g, 0.01), "sgd sparse")


def test_reference_toy_case_dup_ids(hb):
    """tests/test_embedding_op.py:25-89 of the reference: 5x5 table, ids [[0,1],[0,1]], repeated
    SGD steps on duplicate ids (there vs TensorFlow at rtol 1e-5; here vs the order-fixed port,
    bit-exact, and vs a float64 closed form at rtol 1e-5)."""
    from herald_b200 import gpu_links
    from oracle import ops_port
    rng = np.random.default_rng(5)
    table = rng.normal(size=(5, 5)).astype(np.float32)
    ids = np.array([[0, 1], [0, 1]], np.float32)
    p = _dev(hb, table)
    ref = table.copy()
    lr = 0.1
    for it in range(200):
        out = hb.empty((2, 2, 5), hb.gpu(0))
        gpu_links.embedding_lookup(p, _dev(hb, ids), out)
        g = out.asnumpy() * np.float32(0.5)           # d(0.25*sum(x^2))/dx
        gpu_links.sgd_update(p, hb.IndexedSlices(_dev(hb, ids), _dev(hb, g), (5, 5)), lr)
        ref_g = ops_port.embedding_lookup(ref, ids) * np.float32(0.5)
        ref = ops_port.sgd_sparse_update(ref, ids, ref_g, lr)
    assert_bits_equal(p.asnumpy(), ref, "toy SGD")
    closed = table.astype(np.float64)
    closed[:2] *= (1 - 2 * lr * 0.5) ** 200
    np.testing.assert_allclose(p.asnumpy(), closed, rtol=RTOL, atol=1e-30)


def test_adam_sparse_after_dedup(hb):
    from herald_b200 import gpu_links
    from oracle import ops_port
    rng = np.random.default_rng(6)
    V, D, n = 500, 400, 100                          # tests/test_optimizer.py:117-197 shapes
    param = rng.normal(size=(V, D)).astype(np.float32)
    m = np.zeros((V, D), np.float32)
    v = np.zeros((V, D), np.float32)
    pd, md, vd = _dev(hb, param), _dev(hb, m), _dev(hb, v)
    b1t = b2t = 1.0
    for it in range(3):
        ids = rng.integers(0, V, n).astype(np.float32)
        g = rng.normal(size=(n, D)).astype(np.float32)
        b1t, b2t = b1t * 0.9, b2t * 0.999           # optimizer.py:384-385: before the update
        gpu_links.adam_update(pd, hb.IndexedSlices(_dev(hb, ids), _dev(hb, g), (V, D)), md, vd,
                              0.01, 0.9, 0.999, b1t, b2t, 1e-7)
        uniq, inv = ops_port.unique_inverse(ids)
        cg = ops_port.deduplicate(g, inv, len(uniq))
        param, m, v = ops_port.adam_sparse_update(param, uniq, cg, m, v, 0.01, 0.9, 0.999, b1t,
                                                  b2t, 1e-7)
    np.testing.assert_allclose(pd.asnumpy(), param, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(md.asnumpy(), m, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(vd.asnumpy(), v, rtol=RTOL, atol=1e-9)


def test_adam_fused_equals_dedup_then_adam(hb):
    from herald_b200 import gpu_links
    rng = np.random.default_rng(7)
    V, D, n = 300, 128, 2000
    param = rng.normal(size=(V, D)).astype(np.float32)
    ids = ((rng.zipf(1.2, n) - 1) % V).astype(np.float32)
    g = rng.normal(size=(n, D)).astype(np.float32)
    pa, ma, va = _dev(hb, param), _dev(hb, np.zeros((V, D))), _dev(hb, np.zeros((V, D)))
    pb, mb, vb = _dev(hb, param), _dev(hb, np.zeros((V, D))), _dev(hb, np.zeros((V, D)))
    gpu_links.adam_update(pa, hb.IndexedSlices(_dev(hb, ids), _dev(hb, g), (V, D)), ma, va,
                          0.01, 0.9, 0.999, 0.9, 0.999, 1e-7)
    gpu_links.adam_update_fused(pb, hb.IndexedSlices(_dev(hb, ids), _dev(hb, g), (V, D)), mb, vb,
                                0.01, 0.9, 0.999, 0.9, 0.999, 1e-7)
    assert_bits_equal(pa.asnumpy(), pb.asnumpy(), "fused adam param")
    assert_bits_equal(ma.asnumpy(), mb.asnumpy(), "fused adam m")
    assert_bits_equal(va.asnumpy(), vb.asnumpy(), "fused adam v")


def test_adamw_sparse_reference_formula(hb):
    """tests/test_optimizer.py:117-197 (test_adamw_sparse): duplicate ids, dict-style dedup,
    formula at :175-181, atol 1e-5."""
    from herald_b200 import gpu_links
    from oracle import ops_port
    rng = np.random.default_rng(8)
    V, D, n = 500, 400, 100
    param = rng.uniform(-10, 10, size=(V, D)).astype(np.float32)
    ids = rng.integers(0, V, n).astype(np.float32)
    g = rng.uniform(-10, 10, size=(n, D)).astype(np.float32)
    m = rng.uniform(0, 10, size=(V, D)).astype(np.float32)
    v = rng.uniform(0, 10, size=(V, D)).astype(np.float32)
    lr, b1, b2, b1t, b2t, eps, wd = 1e-2, 0.9, 0.99, 0.9 ** 10, 0.99 ** 10, 1e-7, 0.1
    pd, md, vd = _dev(hb, param), _dev(hb, m), _dev(hb, v)
    gpu_links.adamw_update(pd, hb.IndexedSlices(_dev(hb, ids), _dev(hb, g), (V, D)), md, vd, lr,
                           b1, b2, b1t, b2t, eps, wd)
    uniq, inv = ops_port.unique_inverse(ids)
    cg = ops_port.deduplicate(g, inv, len(uniq))
    p2, m2, v2 = ops_port.adamw_sparse_update(param, uniq, cg, m, v, lr, b1, b2, b1t, b2t, eps, wd)
    np.testing.assert_allclose(pd.asnumpy(), p2, rtol=RTOL, atol=1e-5)
    np.testing.assert_allclose(md.asnumpy(), m2, rtol=RTOL, atol=1e-5)
    np.testing.assert_allclose(vd.asnumpy(), v2, rtol=RTOL, atol=1e-5)


def test_lamb_sparse_reference_formula(hb):
    """tests/test_optimizer.py:200-298 (test_lamb_sparse): duplicate ids, norms over the indexed
    rows only, atol 1e-5."""
    from herald_b200 import gpu_links
    from oracle import ops_port
    rng = np.random.default_rng(18)
    for (V, D, n) in ((500, 400, 100), (64, 128, 40), (37, 5, 9)):
        param = rng.uniform(-10, 10, size=(V, D)).astype(np.float32)
        ids = rng.integers(0, V, n).astype(np.float32)
        g = rng.uniform(-10, 10, size=(n, D)).astype(np.float32)
        m = rng.uniform(-10, 10, size=(V, D)).astype(np.float32)
        v = rng.uniform(0, 10, size=(V, D)).astype(np.float32)
        lr, b1, b2, b1t, b2t, eps, wd = 1e-2, 0.9, 0.99, 0.9 ** 10, 0.99 ** 10, 1e-7, 0.1
        pd, md, vd = _dev(hb, param), _dev(hb, m), _dev(hb, v)
        gpu_links.lamb_update(pd, hb.IndexedSlices(_dev(hb, ids), _dev(hb, g), (V, D)), md, vd, lr,
                              b1, b2, b1t, b2t, eps, wd)
        uniq, inv = ops_port.unique_inverse(ids)
        cg = ops_port.deduplicate(g, inv, len(uniq))
        p2, m2, v2 = ops_port.lamb_sparse_update(param, uniq, cg, m, v, lr, b1, b2, b1t, b2t, eps,
                                                 wd)
        np.testing.assert_allclose(pd.asnumpy(), p2, rtol=RTOL, atol=1e-5)
        np.testing.assert_allclose(md.asnumpy(), m2, rtol=RTOL, atol=1e-5)
        np.testing.assert_allclose(vd.asnumpy(), v2, rtol=RTOL, atol=1e-5)
    # rows that are not listed stay untouched
    untouched = np.setdiff1d(np.arange(V), ids.astype(np.int64))
    assert np.array_equal(pd.asnumpy()[untouched], param[untouched])


def test_adagrad_and_momentum_sparse(hb):
    from herald_b200 import gpu_links
    from oracle import ops_port
    rng = np.random.default_rng(9)
    V, D, n = 80, 32, 300
    param = rng.normal(size=(V, D)).astype(np.float32)
    ids = ((rng.zipf(1.3, n) - 1) % V).astype(np.float32)
    g = rng.normal(size=(n, D)).astype(np.float32)
    acc = np.zeros((V, D), np.float32)
    pd, ad = _dev(hb, param), _dev(hb, acc)
    gpu_links.adagrad_update(pd, hb.IndexedSlices(_dev(hb, ids), _dev(hb, g), (V, D)), ad, 0.05, 1e-7)
    uniq, inv = ops_port.unique_inverse(ids)
    cg = ops_port.deduplicate(g, inv, len(uniq))
    p2, a2 = ops_port.adagrad_sparse_update(param, uniq, cg, acc, 0.05, 1e-7)
    np.testing.assert_allclose(pd.asnumpy(), p2, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(ad.asnumpy(), a2, rtol=RTOL, atol=1e-9)
    for nesterov in (False, True):
        vel = rng.normal(size=(V, D)).astype(np.float32)
        pd, vd = _dev(hb, param), _dev(hb, vel)
        gpu_links.momentum_update(pd, hb.IndexedSlices(_dev(hb, ids), _dev(hb, g), (V, D)), vd,
                                  0.05, 0.9, nesterov)
        p2, v2 = ops_port.momentum_sparse_update(param, ids, g, vel, 0.05, 0.9, nesterov)
        np.testing.assert_allclose(pd.asnumpy(), p2, rtol=RTOL, atol=1e-7)
        np.testing.assert_allclose(vd.asnumpy(), v2, rtol=RTOL, atol=1e-7)


def test_oneside_add_lookup_gradient_to_dense_l2(hb):
    import ctypes
    from herald_b200 import gpu_links
    from herald_b200._base import _LIB, check_call
    from oracle import ops_port
    rng = np.random.default_rng(10)
    V, D, n = 90, 128, 700
    ids = ((rng.zipf(1.3, n) - 1) % V).astype(np.float32)
    vals = rng.normal(size=(n, D)).astype(np.float32)
    base = rng.normal(size=(V, D)).astype(np.float32)
    out = _dev(hb, base)
    gpu_links.indexedslice_oneside_add(hb.IndexedSlices(_dev(hb, ids), _dev(hb, vals), (V, D)), out)
    assert_bits_equal(out.asnumpy(), ops_port.indexedslices_oneside_add(ids, vals, base), "oneside add")
    gin = hb.empty((V, D), hb.gpu(0))
    gpu_links.embedding_lookup_gradient(_dev(hb, vals), _dev(hb, ids), gin)
    assert_bits_equal(gin.asnumpy(), ops_port.embedding_lookup_gradient(vals, ids, (V, D)), "lookup grad")
    uids = rng.permutation(V)[:40].astype(np.float32)
    uvals = rng.normal(size=(40, D)).astype(np.float32)
    dense = hb.empty((V, D), hb.gpu(0))
    gpu_links.array_set(dense, 0.0)
    d_uvals, d_uids, d_base = _dev(hb, uvals), _dev(hb, uids), _dev(hb, base)   # keep alive
    check_call(_LIB.IndexedSlices2Dense(d_uvals.handle, d_uids.handle, dense.handle, None))
    assert_bits_equal(dense.asnumpy(), ops_port.indexedslices_to_dense(uvals, uids, (V, D)), "to dense")
    gv = _dev(hb, uvals)
    check_call(_LIB.AddL2RegularizationSparse(d_base.handle, d_uids.handle, gv.handle,
                                              ctypes.c_float(0.01), None))
    np.testing.assert_allclose(gv.asnumpy(), ops_port.add_l2_regularization_sparse(base, uids, uvals, 0.01),
                               rtol=RTOL, atol=1e-7)


def test_array_set_and_copies(hb):
    from herald_b200 import gpu_links
    a = hb.empty((1000, 7), hb.gpu(0))
    gpu_links.array_set(a, 3.5)
    assert np.all(a.asnumpy() == np.float32(3.5))
    gpu_links.array_set(a, 0.0)
    assert np.all(a.asnumpy() == 0)
    h = hb.array(np.arange(12, dtype=np.float32).reshape(3, 4), hb.cpu(0))
    d = h.copyto(hb.gpu(0))
    assert np.array_equal(d.asnumpy(), h.asnumpy())


def test_operator_interface(hb):
    """embedding_lookup_op / gradient op keep the reference's operator contract."""
    from herald_b200.gpu_ops.EmbeddingLookUp import Placeholder
    rng = np.random.default_rng(11)
    V, D = 64, 16
    table = rng.normal(size=(V, D)).astype(np.float32)
    emb, idx = Placeholder(ctx=hb.gpu(0)), Placeholder(ctx=hb.gpu(0))
    op = hb.embedding_lookup_op(emb, idx, ctx=hb.gpu(0))
    assert emb.is_embed
    assert op.infer_shape([(V, D), (4, 26)]) == (4, 26, D)
    ids = rng.integers(0, V, (4, 26)).astype(np.float32)
    out = hb.empty((4, 26, D), hb.gpu(0))
    op.compute([_dev(hb, table), _dev(hb, ids)], out)
    assert np.array_equal(out.asnumpy(), table[ids.astype(np.int64)])
    gnode, _ = op.gradient(None)
    op.infer_shape([(V, D), (4, 26)])            # propagates embed_shape to the gradient node
    assert gnode.infer_shape([None, None]) == (V, D)
    sl = hb.IndexedSlices()
    gnode.compute([out, _dev(hb, ids)], sl)
    assert sl.dense_shape == (V, D) and sl.push_indices is None
