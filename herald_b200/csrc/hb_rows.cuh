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
                        size_t D, F f) {
    using V = RowVec<VEC>;
    const unsigned lane = lane_id();
    const size_t warp_global = (size_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * kRowWarps;
    const size_t nvec = D / VEC;
    const u32 U = *num_unique;
    f.kernel_begin();
    for (size_t u = warp_global; u < U; u += nwarps) {
        const u32 s0 = seg_start[u], s1 = seg_start[u + 1];
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
