"""ParameterServerCommunicateOp for cache-backed embedding parameters —
python/hetu/gpu_ops/ParameterServerCommunicate.py:12-185 restricted to `use_cache_table`.

Keeps the three execution modes of the reference (table in SURVEY §9):
  bsp == 0, prefetch      _compute_bsp_prefetch : update -> wait -> barrier -> lookup(next batch)
  bsp != 0, prefetch      _compute_asp_prefetch : one fused push_pull
  no prefetch             _compute_no_prefetch  : update only (lookup happens in the forward op)
The reference multiplies the sparse gradient by -lr on the host before every push (:24, :58-59).
Here the factor is handed to the cache (`grad_scale`) and applied inside the accumulate kernel —
one exact fp32 product per value, then the exact add: bit-identical to scaling first — so the
gradient is never copied or rewritten, wherever it lives (GPU or pinned host).  `input_val.values`
therefore keeps the RAW gradient.  When the executor passes its stream, the cache's streams are
ordered behind it (the gradient may still be in flight there).
"""
from .. import ndarray
from ..cstable import CacheSparseTable
from ..stream import CSEvent


class ParameterServerCommunicateOp(object):
    def __init__(self, nodeA, parameter, optimizer):
        self.inputs = [nodeA]
        self.parameter = parameter
        self.optimizer = optimizer
        self.learning_rate = -optimizer[1][0]   # only SGD folded into the gradient (:20-24)
        self.cache = None

    def _mult_lr_sparse(self, input_val, stream_handle):
        # folded into the update kernel; ordered behind the stream that produces the gradient
        self.cache.cache.grad_scale = self.learning_rate
        self.cache.cache.after(stream_handle)

    def _push_cache(self, input_val, stream_handle):
        if input_val.push_indices is None:
            return self.cache.embedding_update(input_val.indices, input_val.values)
        return self.cache.embedding_update_with_push_keys(input_val.indices, input_val.push_indices,
                                                          input_val.values)

    def _pull_cache(self):
        return self.cache.embedding_lookup(self.dl_node.get_next_arr(self.dl_name),
                                           self.sparse_pull_val)

    def _push_pull_cache(self, input_val, stream_handle):
        return self.cache.embedding_push_pull(
            pullkeys=self.dl_node.get_next_arr(self.dl_name), dest=self.sparse_pull_val,
            pushkeys=input_val.indices, grads=input_val.values)

    def _compute_bsp_prefetch(self, input_vals, output_val, stream_handle=None):
        self._mult_lr_sparse(input_vals[0], stream_handle)
        self._push_cache(input_vals[0], stream_handle).wait()
        self.comm.BarrierWorker()
        self.parameter.event.update_ts(self._pull_cache())

    def _compute_asp_prefetch(self, input_vals, output_val, stream_handle=None):
        self._mult_lr_sparse(input_vals[0], stream_handle)
        self.parameter.event.update_ts(self._push_pull_cache(input_vals[0], stream_handle))

    def _compute_no_prefetch(self, input_vals, output_val, stream_handle=None):
        self._mult_lr_sparse(input_vals[0], stream_handle)
        self.parameter.event.update_ts(self._push_cache(input_vals[0], stream_handle))

    def forward_hook(self, config):
        """Create the cache from the executor's knobs (:144-185): cstable_policy, cache_limit,
        cache_bound, bsp, prefetch, cache_perf_enable — the --cache/--cache-limit-ratio/--bound
        flags of run_hetu.py / run_laia.py."""
        self.comm = config.ps_comm
        node_shape = self.parameter.shape
        assert len(node_shape) == 2
        if self.parameter.event is None:
            self.parameter.event = CSEvent()
        if config.bsp == 0 and config.prefetch:
            self.compute = self._compute_bsp_prefetch
        elif config.prefetch:
            self.compute = self._compute_asp_prefetch
        else:
            self.compute = self._compute_no_prefetch
        self.cache = CacheSparseTable(config.cache_limit, node_shape[0], node_shape[1],
                                      self.parameter.id, config.cstable_policy, config.cache_bound)
        self.cache.perf_enabled(getattr(config, "cache_perf_enable", False))
        self.parameter.cache = self.cache
        if config.prefetch:
            self.dl_name = config.train_name
            self.dl_node = self.inputs[0].inputs[1]
            local_shape = list(self.dl_node.get_cur_shape(self.dl_name)) + [node_shape[-1]]
            ctx = getattr(config, "embedding_ctx", ndarray.cpu(0))
            self.sparse_pull_val = ndarray.empty(tuple(local_shape), ctx=ctx)
            self.parameter.event.update_ts(self._pull_cache())
            config.ps_map[self.parameter] = self.sparse_pull_val


def parameterServerCommunicate_op(node, parameter, optimizer):
    return ParameterServerCommunicateOp(node, parameter, optimizer)
