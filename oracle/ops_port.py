"""TEST INFRASTRUCTURE: numpy restatements of the reference's op-level kernels.

These are the non-cache ("table resident on the GPU") variants of the same math
(SURVEY.md §3.3).  ids arrive as float32, exactly like the reference's NDArray
path, and are truncated with ``int(ids[i])`` as the kernels do.

Where the reference CUDA kernel uses ``atomicAdd`` (order undefined) the port
fixes the order to ascending occurrence index, which is what the reference's
own serial CPU variants do (src/dnnl_ops/Optimizers.cpp:51-74,
python/hetu/ndarray.py:556-577) and what the B200 kernels implement.

Parity status: formulas pinned by tests/test_oracle_ops.py against the inline
numpy references of the reference's own tests (tests/test_optimizer.py:117-197,
tests/test_embedding_op.py:25-89 semantics).
"""
import numpy as np

f32 = np.float32


def _ids(ids):
    # `int id = ids[index];`  (src/ops/EmbeddingLookup.cu:9)
    return np.asarray(ids, np.float32).reshape(-1).astype(np.int64)


def embedding_lookup(table, ids):
    """src/ops/EmbeddingLookup.cu:3-14 — out[i,:] = table[int(ids[i]),:]."""
    table = np.asarray(table, np.float32)
    ids = np.asarray(ids, np.float32)
    return table[_ids(ids)].reshape(*ids.shape, table.shape[1])


def unique_inverse(ids):
    """python/hetu/ndarray.py:532-536 — np.unique(return_inverse=True) on float32 ids."""
    ids = np.asarray(ids, np.float32).reshape(-1)
    uniq, inverse = np.unique(ids, return_inverse=True)
    return uniq.astype(np.float32), inverse.astype(np.int64)


def deduplicate(values, inverse, num_unique):
    """src/ops/OptimizersSparse.cu:282-295 — compressed[inverse[n],:] += values[n,:]
    (order fixed to ascending n, as cpu_deduplicate python/hetu/ndarray.py:573-574)."""
    values = np.asarray(values, np.float32)
    width = values.shape[-1]
    flat = values.reshape(-1, width)
    out = np.zeros((num_unique, width), np.float32)
    for n, u in enumerate(np.asarray(inverse).reshape(-1).astype(np.int64)):
        out[u] = out[u] + flat[n]
    return out


def sgd_sparse_update(param, ids, grads, lr):
    """src/ops/OptimizersSparse.cu:53-65 — param[id,:] += -lr*g (per occurrence, in order)."""
    param = np.array(param, np.float32)
    width = param.shape[1]
    flat = np.asarray(grads, np.float32).reshape(-1, width)
    neg_lr = f32(-f32(lr))
    for n, i in enumerate(_ids(ids)):
        param[i] = param[i] + neg_lr * flat[n]
    return param


def adam_sparse_update(param, ids, grads, m, v, lr, beta1, beta2, beta1t, beta2t, eps):
    """src/ops/OptimizersSparse.cu:391-416 — ids must already be unique."""
    param, m, v = (np.array(x, np.float32) for x in (param, m, v))
    width = param.shape[1]
    flat = np.asarray(grads, np.float32).reshape(-1, width)
    lr, b1, b2, b1t, b2t, eps = (f32(x) for x in (lr, beta1, beta2, beta1t, beta2t, eps))
    one = f32(1)
    for n, i in enumerate(_ids(ids)):
        g = flat[n]
        cm = b1 * m[i] + (one - b1) * g
        cv = b2 * v[i] + (one - b2) * g * g
        m[i], v[i] = cm, cv
        cm = cm / (one - b1t)
        cv = cv / (one - b2t)
        param[i] = param[i] - lr * cm / (np.sqrt(cv) + eps)
    return param, m, v


def adamw_sparse_update(param, ids, grads, m, v, lr, beta1, beta2, beta1t, beta2t, eps,
                        weight_decay):
    """src/ops/OptimizersSparse.cu:457-483."""
    param, m, v = (np.array(x, np.float32) for x in (param, m, v))
    width = param.shape[1]
    flat = np.asarray(grads, np.float32).reshape(-1, width)
    lr, b1, b2, b1t, b2t, eps, wd = (f32(x) for x in (lr, beta1, beta2, beta1t, beta2t, eps,
                                                      weight_decay))
    one = f32(1)
    for n, i in enumerate(_ids(ids)):
        g = flat[n]
        cm = b1 * m[i] + (one - b1) * g
        cv = b2 * v[i] + (one - b2) * g * g
        m[i], v[i] = cm, cv
        cm = cm / (one - b1t)
        cv = cv / (one - b2t)
        update = cm / (np.sqrt(cv) + eps)
        param[i] = param[i] - lr * (update + wd * param[i])
    return param, m, v


def lamb_sparse_update(param, ids, grads, m, v, lr, beta1, beta2, beta1t, beta2t, eps,
                       weight_decay):
    """src/ops/OptimizersSparse.cu:538-721 — ids must already be unique.  The two norms are taken
    over the listed rows only (get_indexed_params :521-536, cuDNN NORM2 :667-699); they are
    accumulated in float64 here (cuDNN's reduction order is not specified)."""
    param, m, v = (np.array(x, np.float32) for x in (param, m, v))
    width = param.shape[1]
    flat = np.asarray(grads, np.float32).reshape(-1, width)
    lr, b1, b2, b1t, b2t, eps, wd = (f32(x) for x in (lr, beta1, beta2, beta1t, beta2t, eps,
                                                      weight_decay))
    one = f32(1)
    rows = _ids(ids)
    updates = np.zeros_like(flat)
    for n, i in enumerate(rows):
        g = flat[n]
        cm = b1 * m[i] + (one - b1) * g
        cv = b2 * v[i] + (one - b2) * g * g
        m[i], v[i] = cm, cv
        cm = cm / (one - b1t)
        cv = cv / (one - b2t)
        updates[n] = cm / (np.sqrt(cv) + eps)
    norm_p = f32(np.sqrt(np.sum(param[rows].astype(np.float64) ** 2)))
    norm_u = f32(np.sqrt(np.sum(updates.astype(np.float64) ** 2)))
    for n, i in enumerate(rows):
        param[i] = param[i] - lr * (norm_p / norm_u) * (updates[n] + wd * param[i])
    return param, m, v


def adagrad_sparse_update(param, ids, grads, acc, lr, eps):
    """src/ops/OptimizersSparse.cu:331-350 — ids must already be unique."""
    param, acc = np.array(param, np.float32), np.array(acc, np.float32)
    width = param.shape[1]
    flat = np.asarray(grads, np.float32).reshape(-1, width)
    lr, eps = f32(lr), f32(eps)
    for n, i in enumerate(_ids(ids)):
        g = flat[n]
        ca = acc[i] + g * g
        acc[i] = ca
        param[i] = param[i] - lr * g / (np.sqrt(ca) + eps)
    return param, acc


def momentum_sparse_update(param, ids, grads, velocity, lr, momentum, nesterov):
    """src/ops/OptimizersSparse.cu:101-155 — scatter phase on touched rows, then a dense
    sweep over the WHOLE table (second phase)."""
    param, vel = np.array(param, np.float32), np.array(velocity, np.float32)
    width = param.shape[1]
    flat = np.asarray(grads, np.float32).reshape(-1, width)
    neg_lr, mom = f32(-f32(lr)), f32(momentum)
    if nesterov:
        for n, i in enumerate(_ids(ids)):
            t = neg_lr * flat[n]
            vel[i] = vel[i] + t
            param[i] = param[i] + t
        tv = mom * vel
        vel = tv
        param = param + tv
    else:
        for n, i in enumerate(_ids(ids)):
            vel[i] = vel[i] + neg_lr * flat[n]
        param = param + vel
        vel = mom * vel
    return param.astype(np.float32), vel.astype(np.float32)


def add_l2_regularization_sparse(param, ids, grads, l2reg):
    """src/ops/OptimizersSparse.cu:3-17 — grad[n,:] += l2reg * param[id[n],:]."""
    param = np.asarray(param, np.float32)
    width = param.shape[1]
    flat = np.array(grads, np.float32).reshape(-1, width)
    out = flat + f32(l2reg) * param[_ids(ids)]
    return out.reshape(np.asarray(grads).shape).astype(np.float32)


def indexedslices_to_dense(values, ids, dense_shape):
    """src/ops/OptimizersSparse.cu:233-246 — scatter (last writer wins; ids unique in use)."""
    width = dense_shape[-1]
    out = np.zeros(dense_shape, np.float32)
    flat = np.asarray(values, np.float32).reshape(-1, width)
    for n, i in enumerate(_ids(ids)):
        out[i] = flat[n]
    return out


def indexedslices_oneside_add(ids, values, output):
    """src/ops/IndexedSlices.cu:3-15 — output[id[n],:] += values[n,:] (order fixed to n)."""
    out = np.array(output, np.float32)
    width = out.shape[-1]
    flat = np.asarray(values, np.float32).reshape(-1, width)
    for n, i in enumerate(_ids(ids)):
        out[i] = out[i] + flat[n]
    return out


def embedding_lookup_gradient(grad_out, ids, table_shape):
    """src/ops/EmbeddingLookup.cu:54-73 — dense table gradient: zero, then scatter-add."""
    return indexedslices_oneside_add(ids, grad_out, np.zeros(table_shape, np.float32))


# ---- the two-level reduction order of herald_b200's opt-in "split" mode ---------------------------
# (csrc/hb_rows.cuh kSplitTiles / split_run_length; not a reference algorithm: the reference adds in
# occurrence order, src/hetu_cache/include/embedding.h:78-91.  Restated here so that the GPU's
# re-associated result can be checked BIT FOR BIT against a fixed, documented order.)
SPLIT_TILES = 8
VERY_HOT = 1024


def split_run_length(cnt):
    return (((cnt + SPLIT_TILES - 1) // SPLIT_TILES) + 127) & ~127


def accumulate_in_order(row, grads, scale=1.0):
    """((row + s*g0) + s*g1) + ... in float32 — the reference's order."""
    acc = np.array(row, np.float32)
    s = np.float32(scale)
    for g in np.asarray(grads, np.float32):
        acc = acc + g * s
    return acc


def accumulate_two_level(row, grads, scale=1.0):
    """Rows with more than VERY_HOT occurrences: SPLIT_TILES runs of split_run_length(cnt)
    consecutive occurrences, each summed in order from 0, the run sums added to the row in run order."""
    grads = np.asarray(grads, np.float32)
    cnt = len(grads)
    if cnt <= VERY_HOT:
        return accumulate_in_order(row, grads, scale)
    L = split_run_length(cnt)
    s = np.float32(scale)
    acc = np.array(row, np.float32)
    for r in range(SPLIT_TILES):
        if r * L >= cnt:
            break
        part = np.zeros_like(acc)
        for g in grads[r * L:(r + 1) * L]:
            part = part + g * s
        acc = acc + part
    return acc
