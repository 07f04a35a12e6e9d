// Op-level entry points with the reference's names and signatures (the symbols Hetu's
// gpu_links bind from libc_runtime_api.so).  Table resident in HBM, ids carried as float32.
//
// Where the reference kernels are one-thread-per-element with atomicAdd
// (src/ops/OptimizersSparse.cu:53-65, :282-295; src/ops/IndexedSlices.cu:3-15;
// src/ops/EmbeddingLookup.cu:61-73), these sort the ids on the device and let one warp own one
// destination row, adding its occurrences in ascending index: deterministic, no atomics on the
// value path, every row access a coalesced 128-bit-per-lane transaction.
#include <map>
#include <mutex>

#include "hb_segment.cuh"

namespace hb {
namespace {

std::mutex g_ws_mtx;
std::map<cudaStream_t, KeyWorkspace> g_ws;

KeyWorkspace &workspace(cudaStream_t st, size_t n) {
    std::lock_guard<std::mutex> lock(g_ws_mtx);
    KeyWorkspace &ws = g_ws[st];
    ws.reserve(n);
    return ws;
}

inline bool vec4_ok(size_t D, std::initializer_list<const void *> ptrs) {
    if (D % 4)
        return false;
    for (const void *p : ptrs)
        if (reinterpret_cast<uintptr_t>(p) % 16)
            return false;
    return true;
}

// sort float ids and build segments; returns the workspace holding uniq/seg_start/perm
struct Segments {
    KeyWorkspace *ws;
    SortedKeys sk;
};
Segments build_segments(const float *ids, size_t n, u64 key_space, cudaStream_t st) {
    KeyWorkspace &ws = workspace(st, n);
    ws.reset_scans(st);
    Segments s;
    s.ws = &ws;
    s.sk = radix_sort_keys(ws, ids, HB_KEYS_F32, n, bits_for(key_space), st);
    unique_from_sorted(ws, s.sk, n, st);
    return s;
}

// ---- functors ---------------------------------------------------------------------------
struct IndexFromF32 {
    const float *ids;
    __device__ long long operator()(size_t n) const {
        return (long long)(int)ids[n]; // `int id = ids[index];`
    }
};

// dst[uniq[u], :] (+)= sum of occurrences, starting from the row's current contents
template <int VEC>
struct AddIntoRows {
    using V = RowVec<VEC>;
    const u64 *uniq;
    float *dst;
    float *dst2; // optional second destination (nesterov: velocity and param)
    size_t D;
    float scale; // g is multiplied by scale first (exact product, then exact add); 1 = none
    bool scaled;
    using Ctx = size_t; // destination row
    __device__ void kernel_begin() const {}
    __device__ void kernel_end() const {}
    __device__ bool open(size_t u, u32, Ctx &row) const {
        row = (size_t)uniq[u];
        return true;
    }
    __device__ typename V::T load(const Ctx &row, size_t c) const {
        return V::ld(dst + row * D + c * VEC);
    }
    __device__ typename V::T step(const typename V::T &acc, const typename V::T &g) const {
        if (!scaled)
            return V::add(acc, g);
        const float s = scale;
        return V::map2(acc, g, [s](float a, float b) { return __fadd_rn(a, __fmul_rn(s, b)); });
    }
    __device__ void store(const Ctx &row, size_t c, const typename V::T &acc) const {
        V::st(dst + row * D + c * VEC, acc);
    }
};

// nesterov first phase touches two rows with the same increments
template <int VEC>
struct AddIntoTwoRows {
    using V = RowVec<VEC>;
    struct Acc {
        typename V::T a, b;
    };
    const u64 *uniq;
    float *dst, *dst2;
    size_t D;
    float scale;
    using Ctx = size_t;
    __device__ void kernel_begin() const {}
    __device__ void kernel_end() const {}
    __device__ bool open(size_t u, u32, Ctx &row) const {
        row = (size_t)uniq[u];
        return true;
    }
    __device__ Acc load(const Ctx &row, size_t c) const {
        Acc r;
        r.a = V::ld(dst + row * D + c * VEC);
        r.b = V::ld(dst2 + row * D + c * VEC);
        return r;
    }
    __device__ Acc step(const Acc &acc, const typename V::T &g) const {
        const float s = scale;
        Acc r;
        r.a = V::map2(acc.a, g, [s](float a, float b) { return __fadd_rn(a, __fmul_rn(s, b)); });
        r.b = V::map2(acc.b, g, [s](float a, float b) { return __fadd_rn(a, __fmul_rn(s, b)); });
        return r;
    }
    __device__ void store(const Ctx &row, size_t c, const Acc &acc) const {
        V::st(dst + row * D + c * VEC, acc.a);
        V::st(dst2 + row * D + c * VEC, acc.b);
    }
};

struct AdamScalars {
    float lr, beta1, beta2, beta1t, beta2t, eps, weight_decay;
    bool decoupled; // AdamW
};

__device__ __forceinline__ void adam_elem(float &p, float &m, float &v, float g,
                                          const AdamScalars &s) {
    // same expression shapes as src/ops/OptimizersSparse.cu:408-415 / :474-482
    float cur_m = s.beta1 * m + (1 - s.beta1) * g;
    float cur_v = s.beta2 * v + (1 - s.beta2) * g * g;
    m = cur_m;
    v = cur_v;
    cur_m /= (1 - s.beta1t);
    cur_v /= (1 - s.beta2t);
    if (s.decoupled) {
        float update = cur_m / (sqrtf(cur_v) + s.eps);
        p -= s.lr * (update + s.weight_decay * p);
    } else {
        p -= s.lr * cur_m / (sqrtf(cur_v) + s.eps);
    }
}

// Adam / AdamW on rows named by float ids (unique), gradient row n
template <int VEC>
struct AdamRows {
    const float *ids;
    const float *grads;
    float *param, *m, *v;
    size_t D;
    AdamScalars s;
    __device__ bool begin(size_t) const {
        return true;
    }
    __device__ void end(size_t) const {}
    __device__ void apply(size_t n, size_t c) const {
        const size_t row = (size_t)(int)ids[n];
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            size_t o = row * D + c * VEC + k;
            float pp = param[o], mm = m[o], vv = v[o];
            adam_elem(pp, mm, vv, grads[n * D + c * VEC + k], s);
            param[o] = pp;
            m[o] = mm;
            v[o] = vv;
        }
    }
};

// fused: segment-sum then Adam on the unique row
template <int VEC>
struct AdamSegments {
    using V = RowVec<VEC>;
    const u64 *uniq;
    float *param, *m, *v;
    size_t D;
    AdamScalars s;
    using Ctx = size_t;
    __device__ void kernel_begin() const {}
    __device__ void kernel_end() const {}
    __device__ bool open(size_t u, u32, Ctx &row) const {
        row = (size_t)uniq[u];
        return true;
    }
    __device__ typename V::T load(const Ctx &, size_t) const {
        return V::zero();
    }
    __device__ typename V::T step(const typename V::T &acc, const typename V::T &g) const {
        return V::add(acc, g);
    }
    __device__ void store(const Ctx &row, size_t c, const typename V::T &acc) const {
        const float *g = reinterpret_cast<const float *>(&acc);
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            size_t o = row * D + c * VEC + k;
            float pp = param[o], mm = m[o], vv = v[o];
            adam_elem(pp, mm, vv, g[k], s);
            param[o] = pp;
            m[o] = mm;
            v[o] = vv;
        }
    }
};

template <int VEC>
struct AdaGradRows {
    const float *ids;
    const float *grads;
    float *param, *acc;
    size_t D;
    float lr, eps;
    __device__ bool begin(size_t) const {
        return true;
    }
    __device__ void end(size_t) const {}
    __device__ void apply(size_t n, size_t c) const {
        const size_t row = (size_t)(int)ids[n];
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            size_t o = row * D + c * VEC + k;
            float g = grads[n * D + c * VEC + k];
            float cur_acc = acc[o] + g * g; // src/ops/OptimizersSparse.cu:347-349
            acc[o] = cur_acc;
            param[o] -= lr * g / (sqrtf(cur_acc) + eps);
        }
    }
};

template <int VEC>
struct L2Rows {
    const float *ids;
    const float *param;
    float *grads;
    size_t D;
    float l2reg;
    __device__ bool begin(size_t) const {
        return true;
    }
    __device__ void end(size_t) const {}
    __device__ void apply(size_t n, size_t c) const {
        const size_t row = (size_t)(int)ids[n];
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            size_t o = n * D + c * VEC + k;
            grads[o] = grads[o] + l2reg * param[row * D + c * VEC + k];
        }
    }
};

template <int VEC>
struct ScatterRows {
    using V = RowVec<VEC>;
    const float *ids;
    const float *vals;
    float *dst;
    size_t D;
    __device__ bool begin(size_t) const {
        return true;
    }
    __device__ void end(size_t) const {}
    __device__ void apply(size_t n, size_t c) const {
        const size_t row = (size_t)(int)ids[n];
        V::st(dst + row * D + c * VEC, V::ld(vals + n * D + c * VEC));
    }
};


// ---- Lamb (src/ops/OptimizersSparse.cu:538-721) ------------------------------------------
// The reference gathers the indexed rows, takes their L2 norm with cuDNN, writes the Adam
// direction, takes its norm, then applies the trust ratio.  Here: one warp per listed row
// computes the direction, updates m / v and leaves the row's two sums of squares (fixed lane
// order); one block folds the per-row sums in a fixed order (run-to-run identical, unlike a
// cuDNN tree whose shape depends on the library version); one warp per row applies the step.
__global__ void __launch_bounds__(kRowBlock)
    lamb_direction_kernel(const float *__restrict__ ids, const float *__restrict__ grads,
                          const float *__restrict__ param, float *__restrict__ m,
                          float *__restrict__ v, float *__restrict__ update,
                          float *__restrict__ row_sq, size_t n, size_t D, AdamScalars s) {
    pdl_enter();
    const unsigned lane = lane_id();
    const size_t warp_global = (size_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * kRowWarps;
    for (size_t r = warp_global; r < n; r += nwarps) {
        const size_t row = (size_t)(int)ids[r];
        float sp = 0.f, su = 0.f;
        for (size_t c = lane; c < D; c += 32) {
            const size_t o = row * D + c;
            const float g = grads[r * D + c], p = param[o];
            float cur_m = s.beta1 * m[o] + (1 - s.beta1) * g; // :566-573
            float cur_v = s.beta2 * v[o] + (1 - s.beta2) * g * g;
            m[o] = cur_m;
            v[o] = cur_v;
            cur_m /= (1 - s.beta1t);
            cur_v /= (1 - s.beta2t);
            const float u = cur_m / (sqrtf(cur_v) + s.eps);
            update[r * D + c] = u;
            sp += p * p;
            su += u * u;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            sp += __shfl_xor_sync(FULL, sp, d);
            su += __shfl_xor_sync(FULL, su, d);
        }
        if (lane == 0) {
            row_sq[2 * r] = sp;
            row_sq[2 * r + 1] = su;
        }
    }
}

// norms[0] = ||param[ids]||_2, norms[1] = ||update||_2 (CUDNN_REDUCE_TENSOR_NORM2, :667-699)
__global__ void __launch_bounds__(1024) lamb_norms_kernel(const float *__restrict__ row_sq, size_t n,
                                                          float *norms) {
    pdl_enter();
    __shared__ double sh[2][1024];
    double sp = 0.0, su = 0.0;
    for (size_t r = threadIdx.x; r < n; r += 1024) {
        sp += row_sq[2 * r];
        su += row_sq[2 * r + 1];
    }
    sh[0][threadIdx.x] = sp;
    sh[1][threadIdx.x] = su;
    __syncthreads();
    for (int d = 512; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + d];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + d];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        norms[0] = (float)sqrt(sh[0][0]);
        norms[1] = (float)sqrt(sh[1][0]);
    }
}

__global__ void __launch_bounds__(kRowBlock)
    lamb_step_kernel(const float *__restrict__ ids, const float *__restrict__ update,
                     float *__restrict__ param, const float *__restrict__ norms, size_t n, size_t D,
                     float lr, float weight_decay) {
    pdl_enter();
    const unsigned lane = lane_id();
    const size_t warp_global = (size_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * kRowWarps;
    const float n0 = norms[0], n1 = norms[1];
    for (size_t r = warp_global; r < n; r += nwarps) {
        const size_t row = (size_t)(int)ids[r];
        for (size_t c = lane; c < D; c += 32) {
            const size_t o = row * D + c;
            const float p = param[o]; // :592
            param[o] = p - lr * (n0 / n1) * (update[r * D + c] + weight_decay * p);
        }
    }
}

// per-stream float scratch for ops that need an intermediate of the gradient's size
std::map<cudaStream_t, std::pair<float *, size_t>> g_scratch;
float *float_scratch(cudaStream_t st, size_t count) {
    std::lock_guard<std::mutex> lock(g_ws_mtx);
    auto &e = g_scratch[st];
    if (count > e.second) {
        HB_CUDA(cudaStreamSynchronize(st));
        if (e.first)
            cudaFree(e.first);
        e.first = nullptr;
        size_t cap = std::max<size_t>(count, 1 << 16);
        HB_CUDA(cudaMalloc((void **)&e.first, cap * sizeof(float)));
        e.second = cap;
    }
    return e.first;
}

__global__ void momentum_second_phase(float *param, float *veloc, float momentum, bool nesterov,
                                      size_t size) {
    pdl_enter();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < size; i += stride) {
        if (nesterov) { // src/ops/OptimizersSparse.cu:121-131
            float t = momentum * veloc[i];
            veloc[i] = t;
            param[i] = param[i] + t;
        } else { // :147-155
            param[i] = param[i] + veloc[i];
            veloc[i] = momentum * veloc[i];
        }
    }
}

__global__ void emit_unique_f32(const u64 *uniq, const u32 *inverse, const u32 *num_unique,
                                size_t n, float *unique_out, float *inverse_out, i64 *count_out) {
    pdl_enter();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 U = *num_unique;
    if (i < U)
        unique_out[i] = (float)uniq[i];
    if (i < n)
        inverse_out[i] = (float)inverse[i];
    if (i == 0)
        *count_out = (i64)U;
}

template <class F1, class F4>
void launch_segments(bool v4, const Segments &sg, const float *vals, size_t D, size_t n,
                     cudaStream_t st, F1 f1, F4 f4) {
    run_segment_reduce(*sg.ws, sg.sk.perm, vals, D, n, v4, default_hot_threshold(), st, f1, f4);
}

template <class F1, class F4>
void launch_rows(bool v4, size_t n, size_t D, cudaStream_t st, F1 f1, F4 f4) {
    if (n == 0)
        return;
    int grid = row_grid(n);
    if (v4)
        HB_LAUNCH((foreach_row_kernel<4, F4>), grid, kRowBlock, 0, st, n, nullptr, D, f4);
    else
        HB_LAUNCH((foreach_row_kernel<1, F1>), grid, kRowBlock, 0, st, n, nullptr, D, f1);
    HB_LAUNCHED();
}

void scatter_add(float *dst, size_t dst_rows, const float *ids, const float *vals, size_t n,
                 size_t D, float scale, bool scaled, cudaStream_t st) {
    if (n == 0)
        return;
    Segments sg = build_segments(ids, n, dst_rows, st);
    AddIntoRows<1> f1{sg.ws->uniq, dst, nullptr, D, scale, scaled};
    AddIntoRows<4> f4{sg.ws->uniq, dst, nullptr, D, scale, scaled};
    launch_segments(vec4_ok(D, {dst, vals}), sg, vals, D, n, st, f1, f4);
}

} // namespace
} // namespace hb

using namespace hb;

extern "C" {

int DLGpuEmbeddingLookUp(const DLArrayHandle input, const DLArrayHandle ids, DLArrayHandle output,
                         DLStreamHandle stream_handle) {
    HB_API_BEGIN();
    HB_CHECK(input->ndim == 2, "embedding table must be 2-D");
    size_t n = numel(ids), D = (size_t)input->shape[1];
    HB_CHECK(numel(output) == n * D, "output shape must be ids.shape + [width]");
    if (n) {
        cudaStream_t st = stream_of(stream_handle);
        const float *src = (const float *)input->data;
        float *dst = (float *)output->data;
        IndexFromF32 idx{(const float *)ids->data};
        int grid = row_grid((n + 3) / 4);
        if (vec4_ok(D, {src, dst}))
            HB_LAUNCH((gather_rows_kernel<4, 4, IndexFromF32>), grid, kRowBlock, 0, st, src, dst, n, D, idx);
        else
            HB_LAUNCH((gather_rows_kernel<1, 4, IndexFromF32>), grid, kRowBlock, 0, st, src, dst, n, D, idx);
        HB_LAUNCHED();
    }
    HB_API_END();
}

int DLGpuEmbeddingLookUp_Gradient(const DLArrayHandle output_grad, const DLArrayHandle ids,
                                  DLArrayHandle input_grad, DLStreamHandle stream_handle) {
    HB_API_BEGIN();
    HB_CHECK(input_grad->ndim == 2, "table gradient must be 2-D");
    cudaStream_t st = stream_of(stream_handle);
    size_t n = numel(ids), D = (size_t)input_grad->shape[1];
    HB_CHECK(numel(output_grad) == n * D, "output_grad shape must be ids.shape + [width]");
    HB_CUDA(cudaMemsetAsync(input_grad->data, 0, numel(input_grad) * sizeof(float), st));
    scatter_add((float *)input_grad->data, (size_t)input_grad->shape[0], (const float *)ids->data,
                (const float *)output_grad->data, n, D, 1.f, false, st);
    HB_API_END();
}

int DeduplicateIndexedSlices(const DLArrayHandle origin, const DLArrayHandle inverse,
                             DLArrayHandle compressed, DLStreamHandle stream_handle) {
    HB_API_BEGIN();
    size_t D = (size_t)compressed->shape[compressed->ndim - 1];
    size_t n = numel(inverse);
    HB_CHECK(numel(origin) == n * D, "origin shape must be inverse.shape + [width]");
    size_t rows = numel(compressed) / (D ? D : 1);
    scatter_add((float *)compressed->data, rows, (const float *)inverse->data,
                (const float *)origin->data, n, D, 1.f, false, stream_of(stream_handle));
    HB_API_END();
}

int IndexedSlices2Dense(const DLArrayHandle values, const DLArrayHandle indices,
                        DLArrayHandle new_values, DLStreamHandle stream_handle) {
    HB_API_BEGIN();
    size_t D = (size_t)new_values->shape[new_values->ndim - 1];
    size_t n = numel(indices);
    HB_CHECK(numel(values) == n * D, "values shape must be indices.shape + [width]");
    const float *ids = (const float *)indices->data, *vals = (const float *)values->data;
    float *dst = (float *)new_values->data;
    ScatterRows<1> f1{ids, vals, dst, D};
    ScatterRows<4> f4{ids, vals, dst, D};
    launch_rows(vec4_ok(D, {vals, dst}), n, D, stream_of(stream_handle), f1, f4);
    HB_API_END();
}

int IndexedSlicesOneSideAdd(const DLArrayHandle indices, const DLArrayHandle values,
                            DLArrayHandle output, DLStreamHandle stream_handle) {
    HB_API_BEGIN();
    size_t D = (size_t)output->shape[output->ndim - 1];
    size_t n = numel(indices);
    HB_CHECK(numel(values) == n * D, "values shape must be indices.shape + [width]");
    scatter_add((float *)output->data, numel(output) / (D ? D : 1), (const float *)indices->data,
                (const float *)values->data, n, D, 1.f, false, stream_of(stream_handle));
    HB_API_END();
}

int AddL2RegularizationSparse(const DLArrayHandle param, const DLArrayHandle grad_indices,
                              DLArrayHandle grad_values, float l2reg,
                              DLStreamHandle stream_handle) {
    HB_API_BEGIN();
    size_t D = (size_t)param->shape[1], n = numel(grad_indices);
    HB_CHECK(numel(grad_values) == n * D, "grad_values shape must be indices.shape + [width]");
    const float *ids = (const float *)grad_indices->data, *p = (const float *)param->data;
    float *g = (float *)grad_values->data;
    L2Rows<1> f1{ids, p, g, D, l2reg};
    L2Rows<4> f4{ids, p, g, D, l2reg};
    launch_rows(D % 4 == 0, n, D, stream_of(stream_handle), f1, f4);
    HB_API_END();
}

int SGDOptimizerSparseUpdate(DLArrayHandle param, const DLArrayHandle grad_indices,
                             const DLArrayHandle grad_values, float lr,
                             DLStreamHandle stream_handle) {
    HB_API_BEGIN();
    size_t D = (size_t)param->shape[1], n = numel(grad_indices);
    HB_CHECK(numel(grad_values) == n * D, "grad_values shape must be indices.shape + [width]");
    scatter_add((float *)param->data, (size_t)param->shape[0], (const float *)grad_indices->data,
                (const float *)grad_values->data, n, D, -lr, true, stream_of(stream_handle));
    HB_API_END();
}

int MomentumOptimizerSparseUpdate(DLArrayHandle param, const DLArrayHandle grad_indices,
                                  const DLArrayHandle grad_values, DLArrayHandle velocity,
                                  float lr, float momentum, bool nesterov,
                                  DLStreamHandle stream_handle) {
    HB_API_BEGIN();
    cudaStream_t st = stream_of(stream_handle);
    size_t D = (size_t)param->shape[1], n = numel(grad_indices);
    HB_CHECK(numel(grad_values) == n * D, "grad_values shape must be indices.shape + [width]");
    float *p = (float *)param->data, *vel = (float *)velocity->data;
    const float *vals = (const float *)grad_values->data;
    if (n) {
        if (nesterov) {
            Segments sg = build_segments((const float *)grad_indices->data, n,
                                         (size_t)param->shape[0], st);
            AddIntoTwoRows<1> f1{sg.ws->uniq, vel, p, D, -lr};
            AddIntoTwoRows<4> f4{sg.ws->uniq, vel, p, D, -lr};
            launch_segments(vec4_ok(D, {p, vel, vals}), sg, vals, D, n, st, f1, f4);
        } else {
            scatter_add(vel, (size_t)param->shape[0], (const float *)grad_indices->data, vals, n, D,
                        -lr, true, st);
        }
    }
    size_t total = numel(param);
    int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16);
    HB_LAUNCH(momentum_second_phase, std::max(blocks, 1), 256, 0, st, p, vel, momentum, nesterov, total);
    HB_LAUNCHED();
    HB_API_END();
}

int AdaGradOptimizerSparseUpdate(DLArrayHandle param, const DLArrayHandle grad_indices,
                                 const DLArrayHandle grad_values, DLArrayHandle acc, float lr,
                                 float eps, DLStreamHandle stream_handle) {
    HB_API_BEGIN();
    size_t D = (size_t)param->shape[1], n = numel(grad_indices);
    HB_CHECK(numel(grad_values) == n * D, "grad_values shape must be indices.shape + [width]");
    const float *ids = (const float *)grad_indices->data, *g = (const float *)grad_values->data;
    AdaGradRows<1> f1{ids, g, (float *)param->data, (float *)acc->data, D, lr, eps};
    AdaGradRows<4> f4{ids, g, (float *)param->data, (float *)acc->data, D, lr, eps};
    launch_rows(D % 4 == 0, n, D, stream_of(stream_handle), f1, f4);
    HB_API_END();
}

static int adam_rows(DLArrayHandle param, const DLArrayHandle grad_indices,
                     const DLArrayHandle grad_values, DLArrayHandle expavg, DLArrayHandle expavgsq,
                     const AdamScalars &s, DLStreamHandle stream_handle) {
    HB_API_BEGIN();
    size_t D = (size_t)param->shape[1], n = numel(grad_indices);
    HB_CHECK(numel(grad_values) == n * D, "grad_values shape must be indices.shape + [width]");
    const float *ids = (const float *)grad_indices->data, *g = (const float *)grad_values->data;
    AdamRows<1> f1{ids, g, (float *)param->data, (float *)expavg->data, (float *)expavgsq->data, D, s};
    AdamRows<4> f4{ids, g, (float *)param->data, (float *)expavg->data, (float *)expavgsq->data, D, s};
    launch_rows(D % 4 == 0, n, D, stream_of(stream_handle), f1, f4);
    HB_API_END();
}

int AdamOptimizerSparseUpdate(DLArrayHandle param, const DLArrayHandle grad_indices,
                              const DLArrayHandle grad_values, DLArrayHandle expavg,
                              DLArrayHandle expavgsq, float lr, float beta1, float beta2,
                              float beta1t, float beta2t, float eps,
                              DLStreamHandle stream_handle) {
    AdamScalars s{lr, beta1, beta2, beta1t, beta2t, eps, 0.f, false};
    return adam_rows(param, grad_indices, grad_values, expavg, expavgsq, s, stream_handle);
}

int AdamWOptimizerSparseUpdate(DLArrayHandle param, const DLArrayHandle grad_indices,
                               const DLArrayHandle grad_values, DLArrayHandle expavg,
                               DLArrayHandle expavgsq, float lr, float beta1, float beta2,
                               float beta1t, float beta2t, float eps, float weight_decay,
                               DLStreamHandle stream_handle) {
    AdamScalars s{lr, beta1, beta2, beta1t, beta2t, eps, weight_decay, true};
    return adam_rows(param, grad_indices, grad_values, expavg, expavgsq, s, stream_handle);
}

int LambOptimizerSparseUpdate(DLArrayHandle param, const DLArrayHandle grad_indices,
                              const DLArrayHandle grad_values, DLArrayHandle expavg,
                              DLArrayHandle expavgsq, float lr, float beta1, float beta2,
                              float beta1t, float beta2t, float eps, float weight_decay,
                              DLStreamHandle stream_handle) {
    HB_API_BEGIN();
    cudaStream_t st = stream_of(stream_handle);
    size_t D = (size_t)param->shape[1], n = numel(grad_indices);
    HB_CHECK(numel(grad_values) == n * D, "grad_values shape must be indices.shape + [width]");
    if (n) {
        AdamScalars s{lr, beta1, beta2, beta1t, beta2t, eps, weight_decay, true};
        const float *ids = (const float *)grad_indices->data;
        float *p = (float *)param->data;
        // scratch: update [n, D], per-row sums [n, 2], norms [2]
        float *scratch = float_scratch(st, n * D + 2 * n + 2);
        float *update = scratch, *row_sq = scratch + n * D, *norms = row_sq + 2 * n;
        int grid = row_grid(n);
        HB_LAUNCH(lamb_direction_kernel, grid, kRowBlock, 0, st, ids, (const float *)grad_values->data, p,
                                                          (float *)expavg->data,
                                                          (float *)expavgsq->data, update, row_sq, n,
                                                          D, s);
        HB_LAUNCHED();
        HB_LAUNCH(lamb_norms_kernel, 1, 1024, 0, st, row_sq, n, norms);
        HB_LAUNCHED();
        HB_LAUNCH(lamb_step_kernel, grid, kRowBlock, 0, st, ids, update, p, norms, n, D, lr, weight_decay);
        HB_LAUNCHED();
    }
    HB_API_END();
}

int HBUniqueIndexedSlices(const DLArrayHandle ids, DLArrayHandle unique_ids, DLArrayHandle inverse,
                          int64_t *num_unique_dev, DLStreamHandle stream_handle) {
    HB_API_BEGIN();
    cudaStream_t st = stream_of(stream_handle);
    size_t n = numel(ids);
    HB_CHECK(numel(unique_ids) >= n && numel(inverse) >= n, "outputs must hold n elements");
    // float32 carries integers exactly only below 2^24, but any float id is < 2^32 in practice
    Segments sg = build_segments((const float *)ids->data, n, 1ull << 32, st);
    int blocks = std::max(1, ceil_div(n, 256));
    HB_LAUNCH(emit_unique_f32, blocks, 256, 0, st, sg.ws->uniq, sg.ws->inverse, sg.ws->num_unique, n,
                                            (float *)unique_ids->data, (float *)inverse->data,
                                            (i64 *)num_unique_dev);
    HB_LAUNCHED();
    HB_API_END();
}

int HBAdamSparseUpdateFused(DLArrayHandle param, const DLArrayHandle grad_indices,
                            const DLArrayHandle grad_values, DLArrayHandle expavg,
                            DLArrayHandle expavgsq, float lr, float beta1, float beta2,
                            float beta1t, float beta2t, float eps, DLStreamHandle stream_handle) {
    HB_API_BEGIN();
    cudaStream_t st = stream_of(stream_handle);
    size_t D = (size_t)param->shape[1], n = numel(grad_indices);
    HB_CHECK(numel(grad_values) == n * D, "grad_values shape must be indices.shape + [width]");
    if (n) {
        AdamScalars s{lr, beta1, beta2, beta1t, beta2t, eps, 0.f, false};
        Segments sg = build_segments((const float *)grad_indices->data, n, (size_t)param->shape[0], st);
        const float *vals = (const float *)grad_values->data;
        AdamSegments<1> f1{sg.ws->uniq, (float *)param->data, (float *)expavg->data,
                           (float *)expavgsq->data, D, s};
        AdamSegments<4> f4{sg.ws->uniq, (float *)param->data, (float *)expavg->data,
                           (float *)expavgsq->data, D, s};
        launch_segments(vec4_ok(D, {vals}), sg, vals, D, n, st, f1, f4);
    }
    HB_API_END();
}

} // extern "C"
