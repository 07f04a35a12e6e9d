"""ParameterServerCommunicateOp (python/hetu/gpu_ops/ParameterServerCommunicate.py:12-185, cache
branch) in its three modes — bsp-prefetch, asp-prefetch, no-prefetch — on GPU-resident and on
pinned-host gradients, against the oracle driven the way the reference op drives it: the gradient
multiplied by -lr ON THE HOST (numpy float32, :58-59), then embedding_update /
embedding_push_pull / embedding_lookup.  Here the factor is folded into the update kernel; rows,
counters and owner state must come out bit-identical."""
import numpy as np
import pytest

from common import PULL_KEYS, PUSH_KEYS, assert_bits_equal, perf_subset

pytestmark = pytest.mark.gpu

V, D, B, F = 4000, 128, 48, 26
LR = 0.013


class _Param(object):
    def __init__(self, pid, shape):
        self.id, self.shape, self.event, self.is_embed, self.cache = pid, shape, None, True, None


class _Loader(object):
    """Stands in for the dataloader node: get_next_arr hands out the NEXT batch's ids."""

    def __init__(self, batches):
        self.batches, self.pos = batches, 0

    def get_cur_shape(self, name):
        return self.batches[0].shape

    def get_next_arr(self, name):
        return self.batches[self.pos + 1]

    def step(self):
        self.pos += 1


class _Node(object):
    def __init__(self, loader):
        self.inputs = [None, loader]
        self.ctx = None


class _Config(object):
    def __init__(self, comm, policy, limit, bound, bsp, prefetch, ctx):
        self.ps_comm, self.cstable_policy, self.cache_limit, self.cache_bound = comm, policy, limit, bound
        self.bsp, self.prefetch, self.train_name, self.ps_map = bsp, prefetch, "train", {}
        self.cache_perf_enable = True
        self.embedding_ctx = ctx


@pytest.mark.parametrize("mode", ["bsp_prefetch", "asp_prefetch", "no_prefetch"])
@pytest.mark.parametrize("where", ["gpu", "host"])
@pytest.mark.parametrize("bound", [0, 3])
def test_three_modes_match_host_scaled_reference(oracle_impl, mode, where, bound):
    import herald_b200 as hb
    from herald_b200 import ps, stream
    from herald_b200.gpu_ops.ParameterServerCommunicate import parameterServerCommunicate_op
    rng = np.random.default_rng(sum(map(ord, mode + where)) + bound)
    ctx = hb.gpu(0) if where == "gpu" else hb.cpu(0)
    node_id = 8800 + {"bsp_prefetch": 0, "asp_prefetch": 1, "no_prefetch": 2}[mode] * 4 + \
        (2 if where == "gpu" else 0) + (1 if bound else 0)
    rows = rng.normal(0, 0.01, (V, D)).astype(np.float32)
    comm = hb.get_worker_communicate()
    table = comm.InitTensor(node_id, ps.kCacheTable, V, D, ps.Constant, 0.0)
    table.load_rows(rows)
    steps = 6
    ids = [((rng.zipf(1.1, (B, F)) - 1) % V).astype(np.float32) for _ in range(steps + 2)]
    loader = _Loader([hb.array(a, ctx) for a in ids])
    param = _Param(node_id, (V, D))
    op = parameterServerCommunicate_op(_Node(loader), param, ("sgd", (LR,)))
    prefetch = mode != "no_prefetch"
    cfg = _Config(comm, "LRU", 300, bound, 0 if mode == "bsp_prefetch" else -1, prefetch, ctx)
    op.forward_hook(cfg)                                     # prefetch: pulls batch 1 (pos 0 -> next)
    assert op.compute.__name__ == "_compute_" + mode

    osrv = oracle_impl.Server(V, D, rows)
    oc = oracle_impl.Cache(osrv, "lru", 300, bound)
    u64 = lambda a: a.reshape(-1).astype(np.uint64)
    if prefetch:
        exp = oc.embedding_lookup(u64(ids[1]))
        param.event.sync()
        assert_bits_equal(op.sparse_pull_val.asnumpy().reshape(-1, D), exp, "initial prefetch")
    st = stream.create_stream_handle(hb.gpu(0)) if where == "gpu" else None
    for t in range(steps):
        # the gradient of the batch that was looked up last (prefetch) / of batch t (no prefetch)
        cur = ids[loader.pos + 1] if prefetch else ids[t]
        raw = rng.normal(0, 1e-2, (B, F, D)).astype(np.float32)
        grad = hb.IndexedSlices(indices=hb.array(cur, ctx), values=hb.array(raw, ctx), dense_shape=(V, D))
        loader.step()
        op.compute([grad], None, st)
        param.event.sync()
        # reference: values * learning_rate on the host in float32, then the cache call(s)
        scaled = (raw * np.float32(-LR)).astype(np.float32).reshape(-1, D)
        if mode == "asp_prefetch":
            exp = oc.embedding_push_pull(u64(ids[loader.pos + 1]), u64(cur), scaled)
            assert_bits_equal(op.sparse_pull_val.asnumpy().reshape(-1, D), exp, "push_pull rows %d" % t)
        else:
            oc.embedding_update(u64(cur), scaled)
            g, o = op.cache.perf[-2 if prefetch else -1], oc.perf[-1]
            assert perf_subset(g, PUSH_KEYS) == perf_subset(o, PUSH_KEYS), (t, g, dict(o))
            if prefetch:
                exp = oc.embedding_lookup(u64(ids[loader.pos + 1]))
                assert_bits_equal(op.sparse_pull_val.asnumpy().reshape(-1, D), exp, "prefetched rows %d" % t)
                g, o = op.cache.perf[-1], oc.perf[-1]
                assert perf_subset(g, PULL_KEYS) == perf_subset(o, PULL_KEYS), (t, g, dict(o))
        assert_bits_equal(grad.values.asnumpy(), raw, "the op must not rewrite the gradient")
    assert_bits_equal(table.read_rows(), osrv.rows(), "owner rows")
    assert np.array_equal(table.read_versions(), osrv.versions())
    assert np.array_equal(op.cache.keys(), oc.keys())
    op.cache = None
    param.cache = None
    del op
    comm.ClearTensor(node_id)
