// Row-granular device templates shared by the op-level entry points and the cache:
// one warp owns one embedding row at a time, lanes own 128-bit column chunks, so every
// global access of a row is a fully coalesced 16 B x 32 = 512 B (D = 128) transaction.
#pragma once

#include <algorithm>

#include "hb_common.cuh"

namespace hb {

// Vector type per lane: float4 when the row width is a multiple of 4, float otherwise.
template <int VEC>
struct RowVec;
template <>
struct RowVec<4> {
    using T = float4;
    static __device__ __forceinline__ T zero() {
        return make_float4(0.f, 0.f, 0.f, 0.f);
    }
    static __device__ __forceinline__ T add(const T &a, const T &b) {
        return add4(a, b);
    }
    static __device__ __forceinline__ T ld_nc(const float *p) {
        return ld_stream(reinterpret_cast<const float4 *>(p));
    }
    static __device__ __forceinline__ void st_cs(float *p, const T &v) {
        st_stream(reinterpret_cast<float4 *>(p), v);
    }
    static __device__ __forceinline__ T ld(const float *p) {
        return *reinterpret_cast<const float4 *>(p);
    }
    static __device__ __forceinline__ void st(float *p, const T &v) {
        *reinterpret_cast<float4 *>(p) = v;
    }
    template <class F>
    static __device__ __forceinline__ T map2(const T &a, const T &b, F f) {
        return make_float4(f(a.x, b.x), f(a.y, b.y), f(a.z, b.z), f(a.w, b.w));
    }
};
template <>
struct RowVec<1> {
    using T = float;
    static __device__ __forceinline__ T zero() {
        return 0.f;
    }
    static __device__ __forceinline__ T add(const T &a, const T &b) {
        return __fadd_rn(a, b);
    }
    static __device__ __forceinline__ T ld_nc(const float *p) {
        return __ldg(p);
    }
    static __device__ __forceinline__ void st_cs(float *p, const T &v) {
        *p = v;
    }
    static __device__ __forceinline__ T ld(const float *p) {
        return *p;
    }
    static __device__ __forceinline__ void st(float *p, const T &v) {
        *p = v;
    }
    template <class F>
    static __device__ __forceinline__ T map2(const T &a, const T &b, F f) {
        return f(a, b);
    }
};

constexpr int kRowBlock = 256; // 8 warps per CTA
constexpr int kRowWarps = kRowBlock / 32;

inline int row_grid(size_t rows) {
    size_t blocks = (rows + kRowWarps - 1) / kRowWarps;
    size_t cap = (size_t)sm_count() * 8; // 8 resident CTAs of 256 threads per SM = 64 warps
    return (int)std::max<size_t>(1, std::min(blocks, cap));
}

// ---- gather: dst[n,:] = src[index(n),:] ------------------------------------------------
// Index is a functor n -> source row (or < 0 to write zeros).  ROWS rows are in flight per warp
// per iteration: all index loads first, then all row loads, then all stores.
template <int VEC, int ROWS, class Index>
__global__ void __launch_bounds__(kRowBlock)
    gather_rows_kernel(const float *__restrict__ src, float *__restrict__ dst, size_t n, size_t D,
                       Index index) {
    using V = RowVec<VEC>;
    const unsigned lane = lane_id();
    const size_t warp_global = (size_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * kRowWarps;
    const size_t nvec = D / VEC;
    for (size_t row0 = warp_global * ROWS; row0 < n; row0 += nwarps * ROWS) {
        long long srow[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; r++)
            srow[r] = (row0 + r < n) ? index(row0 + r) : -1;
        for (size_t c = lane; c < nvec; c += 32) {
            typename V::T v[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; r++)
                v[r] = srow[r] >= 0 ? V::ld(src + (size_t)srow[r] * D + c * VEC) : V::zero();
#pragma unroll
            for (int r = 0; r < ROWS; r++)
                if (row0 + r < n)
                    V::st_cs(dst + (row0 + r) * D + c * VEC, v[r]);
        }
    }
}

// ---- segment reduce by unique key, applied to a destination row -----------------------
// For unique u (one warp):  F::Ctx ctx; if (!f.begin(u, cnt, ctx)) skip;
//   per 128-bit column chunk c:  acc = f.load(ctx, c);
//                                for each occurrence p in ascending original index:
//                                    acc = f.step(acc, vals[perm[p], c]);
//                                f.store(ctx, c, acc);
//   f.end(ctx)   (all lanes; per-row scalars are written by lane 0 inside)
// The adds happen in occurrence order, so the result is deterministic and equal to a serial
// CPU loop over the batch.
template <int VEC, class F>
__global__ void __launch_bounds__(kRowBlock)
    segment_rows_kernel(const u32 *__restrict__ seg_start, const u32 *__restrict__ perm,
                        const u32 *__restrict__ num_unique, const float *__restrict__ vals,
                        size_t D, u32 hot_threshold, F f) {
    using V = RowVec<VEC>;
    const unsigned lane = lane_id();
    const size_t warp_global = (size_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * kRowWarps;
    const size_t nvec = D / VEC;
    const u32 U = *num_unique;
    f.kernel_begin();
    for (size_t u = warp_global; u < U; u += nwarps) {
        const u32 s0 = seg_start[u], s1 = seg_start[u + 1];
        if (s1 - s0 > hot_threshold)
            continue; // long segments belong to segment_hot_kernel
        typename F::Ctx ctx;
        if (!f.begin(u, s1 - s0, ctx))
            continue;
        for (size_t c = lane; c < nvec; c += 32) {
            auto acc = f.load(ctx, c);
            u32 p = s0;
            for (; p + 4 <= s1; p += 4) {
                typename V::T g[4];
#pragma unroll
                for (int k = 0; k < 4; k++)
                    g[k] = V::ld_nc(vals + (size_t)perm[p + k] * D + c * VEC);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    acc = f.step(acc, g[k]);
            }
            for (; p < s1; p++)
                acc = f.step(acc, V::ld_nc(vals + (size_t)perm[p] * D + c * VEC));
            f.store(ctx, c, acc);
        }
        f.end(ctx);
    }
    f.kernel_end();
}

// ---- long segments (hot ids) -------------------------------------------------------------
// A Zipf batch has ids that occur thousands of times; one warp walking such a segment would
// be the critical path of the whole step.  The ADD ORDER is fixed by the parity contract
// (ascending occurrence), so a segment cannot be split by rows — it is split by COLUMNS: a work
// item is (hot row, 32-float column chunk); a CTA streams the chunk of every occurrence through
// shared memory with all 8 warps (256 rows = 32 KB in flight) while warp 0 adds them in order,
// one column per lane.  Per-row scalars are applied afterwards by segment_hot_finish_kernel,
// once every chunk of the row has read them.
constexpr int kHotTileRows = 256;
constexpr u32 kVeryHot = 1024; // rows above this go first (longest-processing-time-first)

struct HotLists {
    u32 *very_hot; // [cap] unique indices with count > kVeryHot
    u32 *hot;      // [cap] unique indices with hot_threshold < count <= kVeryHot
    u32 *ctrl;     // [0] = #very_hot, [1] = #hot, [2] = work ticket (zeroed with the scan arena)
};

__device__ __forceinline__ u32 rows_warp_append(u32 *counter, bool pred) {
    unsigned m = __ballot_sync(FULL, pred);
    if (!m)
        return 0;
    int leader = __ffs(m) - 1;
    u32 base = 0;
    if ((int)lane_id() == leader)
        base = atomicAdd(counter, (u32)__popc(m));
    base = __shfl_sync(FULL, base, leader);
    return base + __popc(m & lanemask_lt());
}

static __global__ void __launch_bounds__(256)
    build_hot_lists_kernel(const u32 *__restrict__ seg_start, const u32 *__restrict__ num_unique,
                           u32 hot_threshold, HotLists hl) {
    const u32 U = *num_unique;
    const u32 stride = gridDim.x * blockDim.x;
    const u32 rounds = (U + stride - 1) / stride;
    for (u32 it = 0; it < rounds; it++) {
        const u32 u = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        u32 cnt = 0;
        if (u < U)
            cnt = seg_start[u + 1] - seg_start[u];
        const bool a = cnt > kVeryHot && cnt > hot_threshold;
        const bool b = !a && cnt > hot_threshold;
        u32 pa = rows_warp_append(&hl.ctrl[0], a);
        if (a)
            hl.very_hot[pa] = u;
        u32 pb = rows_warp_append(&hl.ctrl[1], b);
        if (b)
            hl.hot[pb] = u;
    }
}

template <class F> // F is the VEC = 1 instantiation: load/step/store address single columns
__global__ void __launch_bounds__(kRowBlock)
    segment_hot_kernel(const u32 *__restrict__ seg_start, const u32 *__restrict__ perm,
                       const float *__restrict__ vals, size_t D, HotLists hl, F f) {
    __shared__ float tile[kHotTileRows][32];
    __shared__ u32 s_item;
    constexpr int ROWS_PER_WARP = kHotTileRows / kRowWarps; // 32
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const u32 Q = (u32)((D + 31) / 32);
    const u32 nA = hl.ctrl[0], nB = hl.ctrl[1];
    const u32 total = (nA + nB) * Q;
    f.kernel_begin();
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0)
            s_item = atomicAdd(&hl.ctrl[2], 1u);
        __syncthreads();
        const u32 t = s_item;
        if (t >= total)
            break;
        const u32 h = t / Q, q = t % Q;
        const u32 u = h < nA ? hl.very_hot[h] : hl.hot[h - nA];
        const u32 s0 = seg_start[u], s1 = seg_start[u + 1];
        typename F::Ctx ctx;
        if (!f.begin(u, s1 - s0, ctx))
            continue;
        const size_t col = (size_t)q * 32 + lane;
        const bool active = col < D;
        decltype(f.load(ctx, 0)) acc;
        if (warp == 0 && active)
            acc = f.load(ctx, col);
        const u32 ntiles = (s1 - s0 + kHotTileRows - 1) / kHotTileRows;
        float reg[ROWS_PER_WARP];
        auto issue = [&](u32 k) {
            const u32 base = s0 + k * kHotTileRows + warp * ROWS_PER_WARP;
#pragma unroll
            for (int j = 0; j < ROWS_PER_WARP; j++) {
                const u32 p = base + j;
                reg[j] = (p < s1 && active) ? __ldg(vals + (size_t)perm[p] * D + col) : 0.f;
            }
        };
        issue(0);
        for (u32 k = 0; k < ntiles; k++) {
#pragma unroll
            for (int j = 0; j < ROWS_PER_WARP; j++)
                tile[warp * ROWS_PER_WARP + j][lane] = reg[j];
            __syncthreads();
            if (k + 1 < ntiles)
                issue(k + 1); // next tile's loads fly while warp 0 adds this one
            if (warp == 0 && active) {
                const u32 rows = min((u32)kHotTileRows, s1 - s0 - k * kHotTileRows);
                u32 r = 0;
                for (; r + 8 <= rows; r += 8) {
                    float g[8];
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        g[j] = tile[r + j][lane];
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        acc = f.step(acc, g[j]);
                }
                for (; r < rows; r++)
                    acc = f.step(acc, tile[r][lane]);
            }
            __syncthreads();
        }
        if (warp == 0 && active)
            f.store(ctx, col, acc);
    }
    f.kernel_end();
}

// per-row scalars of the hot rows, after every chunk has been processed
template <class F>
__global__ void __launch_bounds__(kRowBlock)
    segment_hot_finish_kernel(const u32 *__restrict__ seg_start, HotLists hl, F f) {
    const u32 nA = hl.ctrl[0], nB = hl.ctrl[1];
    const size_t warp_global = (size_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * kRowWarps;
    f.kernel_begin();
    for (size_t h = warp_global; h < nA + nB; h += nwarps) {
        const u32 u = h < nA ? hl.very_hot[h] : hl.hot[h - nA];
        typename F::Ctx ctx;
        if (f.begin(u, seg_start[u + 1] - seg_start[u], ctx))
            f.end(ctx);
    }
    f.kernel_end();
}

// ---- one warp per listed row: f.begin(r); f.apply(r, c) per column chunk; f.end(r) -------------
template <int VEC, class F>
__global__ void __launch_bounds__(kRowBlock)
    foreach_row_kernel(size_t nrows, const u32 *__restrict__ nrows_dev, size_t D, F f) {
    const unsigned lane = lane_id();
    const size_t warp_global = (size_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * kRowWarps;
    const size_t nvec = D / VEC;
    const size_t R = nrows_dev ? (size_t)*nrows_dev : nrows;
    for (size_t r = warp_global; r < R; r += nwarps) {
        if (!f.begin(r))
            continue;
        for (size_t c = lane; c < nvec; c += 32)
            f.apply(r, c);
        f.end(r);
    }
}

} // namespace hb
