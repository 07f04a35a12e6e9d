"""EmbeddingLookUp / EmbeddingLookUp_Gradient operators — python/hetu/gpu_ops/EmbeddingLookUp.py.

Hetu's graph executor is out of scope; these classes keep the operator contract
(`compute(input_vals, output_val, stream_handle)`, `gradient`, `infer_shape`, the hook that picks
the compute path) so an executor — or a test — drives them exactly like the reference ops.
Differences: the table and the cache are in HBM, so the op stays on the GPU in Hybrid mode
(the reference forces it to cpu(0): EmbeddingLookUp.py:77-86).
"""
from .. import ndarray
from ..gpu_links import embedding_lookup


class Node(object):
    """Minimal stand-in for hetu.gpu_ops.Node.Op: inputs, ctx, event."""

    def __init__(self, inputs, ctx=None):
        self.inputs = list(inputs)
        self.ctx = ctx
        self.raw_ctx = ctx
        self.event = None


class Placeholder(Node):
    """An embedding parameter / index feed: holds a value and, for parameters, the cache."""

    def __init__(self, value=None, ctx=None, is_embed=False, node_id=0):
        super().__init__([], ctx)
        self.value = value
        self.is_embed = is_embed
        self.id = node_id
        self.cache = None


class EmbeddingLookUp(Node):
    def __init__(self, embedding, index, enable_push_index=False, ctx=None):
        super().__init__([embedding, index], ctx)
        self.enable_push_index = enable_push_index
        embedding.is_embed = True
        self.compute = self._compute_gpu

    def _compute_gpu(self, input_vals, output_val, stream_handle=None):
        embedding_lookup(input_vals[0], input_vals[1], output_val, stream_handle)

    def _compute_sparsepull_from_cache(self, input_vals, output_val, stream_handle=None):
        # EmbeddingLookUp.py:36-41
        self.event.sync()
        if self.bsp == 0:
            self.comm.BarrierWorker()
        ts = self.inputs[0].cache.embedding_lookup(input_vals[1], output_val)
        self.event.update_ts(ts)

    def gradient(self, output_grad):
        self.grad_node = embedding_lookup_gradient_op(
            output_grad, self.inputs[1], None, self.enable_push_index, ctx=self.raw_ctx)
        return [self.grad_node, None]

    def infer_shape(self, input_shapes):
        assert len(input_shapes) == 2
        if hasattr(self, "grad_node"):
            self.grad_node.embed_shape = input_shapes[0]
        return tuple(list(input_shapes[1]) + [input_shapes[0][1]])

    def forward_hook(self, config):
        """Pick the compute path from the executor config (EmbeddingLookUp.py:56-75)."""
        if getattr(config, "cstable_policy", None):
            self.event = self.inputs[0].event
            if not getattr(config, "prefetch", True):
                self.bsp = config.bsp
                self.comm = config.ps_comm
                self.compute = self._compute_sparsepull_from_cache
        else:
            self.compute = self._compute_gpu


class EmbeddingLookUp_Gradient(Node):
    def __init__(self, vectors, index, embed_shape, enable_push_index, ctx=None):
        super().__init__([vectors, index], ctx)
        self.embed_shape = embed_shape
        self.enable_push_index = enable_push_index

    def compute(self, input_vals, output_val, stream_handle=None):
        # no arithmetic: wraps (values, indices[, push_indices]) — EmbeddingLookUp.py:95-113
        assert self.embed_shape
        idx = input_vals[1]
        if self.enable_push_index:
            if not isinstance(idx, tuple):
                raise TypeError
            output_val.update(values=input_vals[0], indices=idx[0], push_indices=idx[1],
                              dense_shape=self.embed_shape)
        else:
            if isinstance(idx, tuple):
                idx = idx[0]
            output_val.update(values=input_vals[0], indices=idx, push_indices=None,
                              dense_shape=self.embed_shape)

    def gradient(self, output_grad):
        raise NotImplementedError

    def infer_shape(self, input_shapes):
        assert self.embed_shape
        return self.embed_shape


def embedding_lookup_op(embedding, index, enable_push_index=False, ctx=None):
    """Make a new EmbeddingLookUp node (EmbeddingLookUp.py:128-143)."""
    return EmbeddingLookUp(embedding, index, enable_push_index, ctx=ctx)


def embedding_lookup_gradient_op(vectors, index, embed_shape, enable_push_index=False, ctx=None):
    return EmbeddingLookUp_Gradient(vectors, index, embed_shape, enable_push_index, ctx=ctx)
