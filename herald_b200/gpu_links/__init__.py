"""Link layer — mirrors python/hetu/gpu_links/{EmbeddingLookUpLink,OptimizerLink,IndexedSliceLink}.py.

Same function names, argument order and dedup decisions as the reference (AdaGrad/Adam/AdamW
deduplicate first, SGD/Momentum accept duplicate ids: OptimizerLink.py:23-116)."""
import ctypes

from .._base import _LIB, check_call
from .. import ndarray as _nd


def _s(stream):
    return stream.handle if stream else None


def embedding_lookup(in_mat, ids, out_mat, stream=None):
    assert isinstance(in_mat, _nd.NDArray) and isinstance(ids, _nd.NDArray)
    assert isinstance(out_mat, _nd.NDArray)
    check_call(_LIB.DLGpuEmbeddingLookUp(in_mat.handle, ids.handle, out_mat.handle, _s(stream)))


def embedding_lookup_gradient(grad_out, ids, grad_in, stream=None):
    assert isinstance(grad_out, _nd.NDArray) and isinstance(ids, _nd.NDArray)
    assert isinstance(grad_in, _nd.NDArray)
    check_call(_LIB.DLGpuEmbeddingLookUp_Gradient(grad_out.handle, ids.handle, grad_in.handle,
                                                  _s(stream)))


def indexedslice_oneside_add(indslice, output, stream=None):
    assert isinstance(indslice.indices, _nd.NDArray) and isinstance(indslice.values, _nd.NDArray)
    assert isinstance(output, _nd.NDArray)
    check_call(_LIB.IndexedSlicesOneSideAdd(indslice.indices.handle, indslice.values.handle,
                                            output.handle, _s(stream)))


def array_set(arr, value, stream=None):
    check_call(_LIB.DLGpuArraySet(arr.handle, ctypes.c_float(value), _s(stream)))


def add_l2_regularization(param, grad, l2reg, stream=None):
    assert isinstance(param, _nd.NDArray)
    if not isinstance(grad, _nd.IndexedSlices):
        raise NotImplementedError("dense l2 regularisation is outside the embedding hot path")
    grad.deduplicate(stream)
    grad.to_dense(stream)
    check_call(_LIB.AddL2RegularizationSparse(param.handle, grad.indices.handle, grad.values.handle,
                                              ctypes.c_float(l2reg), _s(stream)))


def _sparse(grad):
    if not isinstance(grad, _nd.IndexedSlices):
        raise NotImplementedError("dense optimizer updates are outside the embedding hot path")
    assert isinstance(grad.indices, _nd.NDArray) and isinstance(grad.values, _nd.NDArray)
    return grad


def sgd_update(param, grad, lr, stream=None):
    assert isinstance(param, _nd.NDArray)
    g = _sparse(grad)
    check_call(_LIB.SGDOptimizerSparseUpdate(param.handle, g.indices.handle, g.values.handle,
                                             ctypes.c_float(lr), _s(stream)))
    g.free_dense()


def momentum_update(param, grad, velocity, lr, momentum, nesterov, stream=None):
    g = _sparse(grad)
    check_call(_LIB.MomentumOptimizerSparseUpdate(
        param.handle, g.indices.handle, g.values.handle, velocity.handle, ctypes.c_float(lr),
        ctypes.c_float(momentum), ctypes.c_bool(nesterov), _s(stream)))
    g.free_dense()


def adagrad_update(param, grad, accumulation, lr, eps, stream=None):
    if not isinstance(grad, _nd.IndexedSlices):
        raise NotImplementedError("dense optimizer updates are outside the embedding hot path")
    grad.deduplicate(stream)
    g = _sparse(grad)
    check_call(_LIB.AdaGradOptimizerSparseUpdate(
        param.handle, g.indices.handle, g.values.handle, accumulation.handle, ctypes.c_float(lr),
        ctypes.c_float(eps), _s(stream)))
    g.free_deduplicate()
    g.free_dense()


def adam_update(param, grad, expavg, expavgsq, lr, beta1, beta2, beta1t, beta2t, eps, stream=None):
    if not isinstance(grad, _nd.IndexedSlices):
        raise NotImplementedError("dense optimizer updates are outside the embedding hot path")
    grad.deduplicate(stream)
    g = _sparse(grad)
    check_call(_LIB.AdamOptimizerSparseUpdate(
        param.handle, g.indices.handle, g.values.handle, expavg.handle, expavgsq.handle,
        ctypes.c_float(lr), ctypes.c_float(beta1), ctypes.c_float(beta2), ctypes.c_float(beta1t),
        ctypes.c_float(beta2t), ctypes.c_float(eps), _s(stream)))
    g.free_deduplicate()
    g.free_dense()


def adam_update_fused(param, grad, expavg, expavgsq, lr, beta1, beta2, beta1t, beta2t, eps,
                      stream=None):
    """Same result as adam_update without materialising unique ids / compressed grads:
    sort + deterministic segment reduce + Adam in one pass (no host round trip)."""
    g = _sparse(grad)
    check_call(_LIB.HBAdamSparseUpdateFused(
        param.handle, g.indices.handle, g.values.handle, expavg.handle, expavgsq.handle,
        ctypes.c_float(lr), ctypes.c_float(beta1), ctypes.c_float(beta2), ctypes.c_float(beta1t),
        ctypes.c_float(beta2t), ctypes.c_float(eps), _s(stream)))


def adamw_update(param, grad, expavg, expavgsq, lr, beta1, beta2, beta1t, beta2t, eps,
                 weight_decay, stream=None):
    if not isinstance(grad, _nd.IndexedSlices):
        raise NotImplementedError("dense optimizer updates are outside the embedding hot path")
    grad.deduplicate(stream)
    g = _sparse(grad)
    check_call(_LIB.AdamWOptimizerSparseUpdate(
        param.handle, g.indices.handle, g.values.handle, expavg.handle, expavgsq.handle,
        ctypes.c_float(lr), ctypes.c_float(beta1), ctypes.c_float(beta2), ctypes.c_float(beta1t),
        ctypes.c_float(beta2t), ctypes.c_float(eps), ctypes.c_float(weight_decay), _s(stream)))
    g.free_deduplicate()


def lamb_update(param, grad, expavg, expavgsq, lr, beta1, beta2, beta1t, beta2t, eps,
                weight_decay, stream=None):
    """python/hetu/gpu_links/OptimizerLink.py:102-116 (sparse branch)."""
    if not isinstance(grad, _nd.IndexedSlices):
        raise NotImplementedError("dense optimizer updates are outside the embedding hot path")
    grad.deduplicate(stream)
    g = _sparse(grad)
    check_call(_LIB.LambOptimizerSparseUpdate(
        param.handle, g.indices.handle, g.values.handle, expavg.handle, expavgsq.handle,
        ctypes.c_float(lr), ctypes.c_float(beta1), ctypes.c_float(beta2), ctypes.c_float(beta1t),
        ctypes.c_float(beta2t), ctypes.c_float(eps), ctypes.c_float(weight_decay), _s(stream)))
    g.free_deduplicate()
    g.free_dense()
