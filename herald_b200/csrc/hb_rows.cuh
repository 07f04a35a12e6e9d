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
    static __device__ __forceinline__ T ld_rmw(const float *p) { // row read once, then overwritten
        return ld_row(reinterpret_cast<const float4 *>(p));
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
    static __device__ __forceinline__ T ld_rmw(const float *p) {
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
// Index is a functor n -> source row (or < 0 to write zeros).  A warp takes 32 consecutive
// destination rows: their source rows are resolved lane-parallel (one coalesced index load
// instead of 32 broadcast ones), then ROWS rows are in flight per warp per iteration: all row
// loads first, then all stores.
template <int VEC, int ROWS, class Index>
__global__ void __launch_bounds__(kRowBlock)
    gather_rows_kernel(const float *__restrict__ src, float *__restrict__ dst, size_t n, size_t D,
                       Index index) {
    pdl_enter();
    using V = RowVec<VEC>;
    const unsigned lane = lane_id();
    const size_t warp_global = (size_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * kRowWarps;
    const size_t nvec = D / VEC;
    for (size_t base = warp_global * 32; base < n; base += nwarps * 32) {
        const long long mine = base + lane < n ? index(base + lane) : -1;
        const int rows_here = (int)min((size_t)32, n - base);
#pragma unroll 1
        for (int r0 = 0; r0 < rows_here; r0 += ROWS) {
            long long srow[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; r++)
                srow[r] = __shfl_sync(FULL, mine, (r0 + r) & 31);
            for (size_t c = lane; c < nvec; c += 32) {
                typename V::T v[ROWS];
#pragma unroll
                for (int r = 0; r < ROWS; r++)
                    v[r] = (r0 + r < rows_here && srow[r] >= 0)
                               ? V::ld(src + (size_t)srow[r] * D + c * VEC)
                               : V::zero();
#pragma unroll
                for (int r = 0; r < ROWS; r++)
                    if (r0 + r < rows_here)
                        V::st_cs(dst + (base + r0 + r) * D + c * VEC, v[r]);
            }
        }
    }
}

// ---- segment reduce by unique key, applied to a destination row -----------------------
// For unique u:  F::Ctx ctx; if (!f.begin(u, cnt, ctx)) skip;
//   per column chunk c:  acc = f.load(ctx, c);
//                        for each occurrence p in ascending original index:
//                            acc = f.step(acc, vals[perm[p], c]);
//                        f.store(ctx, c, acc);
//   f.end(ctx)   (ONE thread per row: writes the per-row scalars from values begin() read)
// begin() and end() contain no warp-collective operation: the cold phase calls them with a
// different row in every lane.
// The adds happen in occurrence order, so the result is deterministic and bit-identical to a
// serial CPU loop over the batch (Line::accumulate order, src/hetu_cache/include/embedding.h:78-91).
//
// One persistent kernel does the whole reduce:
//   * hot phase — a Zipf batch has ids that occur thousands of times and the add ORDER is fixed
//     by the parity contract, so such a segment cannot be split by rows; it is split by COLUMNS.
//     A work item is (hot row, 32-float column chunk) — 16-float chunks for the very hot rows.
//     The CTA streams the chunk of every occurrence through a cp.async ring in shared memory
//     (hot_stages() deep; all 8 warps issue, 128 occurrences x 128 B per stage; the occurrence
//     indices travel through a second ring, copied ahead in the same groups) while warp 0 adds
//     them in order, one column per lane.  Items are taken from a ticket, longest rows first;
//     the CTA that finishes the last chunk of a row applies the row's scalars (f.end).
//   * cold phase — every warp takes tickets of ticket_rows() (16) uniques, strided over the key
//     range, and walks them ROWS at a time: the metadata of the ticket is loaded lane-parallel
//     (F::peek + the segment bounds first, then F::begin), then the row / gradient / owner-row
//     loads of ROWS segments are in flight together (128-bit per lane) before the first add.
constexpr int kHotTileRows = 128; // occurrences per pipeline stage
// 3 x 128 x 128 B = 48 KB (+ 6 KB of indices) of dynamic shared memory per CTA, x 2 CTAs/SM.  Measured
// on B200 (profiles/r01_segtrace_*.json): the per-occurrence rate of a hot chain does not depend on
// the ring depth (3 .. 10 stages give 6.8 ns), but every 16 KB stage is taken from the SM's L1, and the
// cold phase needs L1 lines for its 12 x 512 B loads in flight per warp: 6 stages -> 111 us, 3 -> 89 us.
constexpr int kHotStagesDefault = 3;
constexpr int kHotStagesMax = 12;
constexpr int kHotStageBytes = kHotTileRows * 32 * 4;
// dynamic shared memory of segment_reduce_kernel: the data ring + the index ring (2 x the ring
// depth of the 16-column variant = 4 x stages tiles of kHotTileRows u32)
constexpr size_t hot_smem_bytes(int stages) {
    return (size_t)stages * kHotStageBytes + (size_t)4 * stages * kHotTileRows * 4;
}
constexpr u32 kVeryHot = 1024; // rows above this go first (longest-processing-time-first)

// ring depth of the hot phase ($HERALD_HOT_STAGES, 3 .. 12): the bytes a CTA keeps in flight are
// (stages - 1) x 16 KB, which is what hides HBM latency under the dependent add chain
int hot_stages();
// uniques per cold-phase ticket ($HERALD_TICKET_ROWS, 4 .. 32)
u32 ticket_rows();

// optional per-CTA timeline of the last segment_reduce launch (diagnostics, HBSegTraceEnable):
// [0] = grid, [1] = hot items; then 4 words per CTA {start, hot phase end, end, items taken};
// then 3 words per hot item {start, end, occurrences} for the first kTraceItems items
constexpr u32 kTraceCtas = 2048, kTraceItems = 1024;
constexpr size_t kTraceWords = 2 + 4 * (size_t)kTraceCtas + 3 * (size_t)kTraceItems;
u64 *seg_trace_buffer(); // null unless enabled

__device__ __forceinline__ u64 global_timer_ns() {
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// broadcast a trivially copyable struct from lane `src` (32 bits at a time)
template <class T>
__device__ __forceinline__ T shfl_pod(const T &v, int src) {
    static_assert(sizeof(T) % 4 == 0, "pad the context to a multiple of 4 bytes");
    union U {
        T t;
        u32 w[sizeof(T) / 4];
        __device__ U() {}
    } in, out;
    in.t = v;
#pragma unroll
    for (size_t i = 0; i < sizeof(T) / 4; i++)
        out.w[i] = __shfl_sync(FULL, in.w[i], src);
    return out.t;
}

// optional two-step row opening: F::peek(u) (loads that depend on nothing but u) + F::begin(pre, u,
// cnt, ctx); functors without peek keep the one-step begin(u, cnt, ctx)
template <class F>
__device__ __forceinline__ auto seg_peek(const F &f, size_t u, int) -> decltype(f.peek(u)) {
    return f.peek(u);
}
template <class F>
__device__ __forceinline__ int seg_peek(const F &, size_t, long) {
    return 0;
}
template <class F, class P>
__device__ __forceinline__ auto seg_begin(const F &f, const P &pre, size_t u, u32 cnt,
                                          typename F::Ctx &x, int) -> decltype(f.begin(pre, u, cnt, x)) {
    return f.begin(pre, u, cnt, x);
}
template <class F, class P>
__device__ __forceinline__ bool seg_begin(const F &f, const P &, size_t u, u32 cnt, typename F::Ctx &x,
                                          long) {
    return f.begin(u, cnt, x);
}

struct HotLists {
    u32 *very_hot; // [cap] unique indices with count > kVeryHot
    u32 *hot;      // [cap] unique indices with hot_threshold < count <= kVeryHot
    u32 *done_a;   // [cap] finished chunks per very-hot row (zeroed by build_hot_lists_kernel)
    u32 *done_b;   // [cap] same for the hot list
    u32 *ctrl;     // [0] = #very_hot, [1] = #hot, [2] = hot ticket, [3] = cold ticket (zeroed with
                   // the scan arena)
    u64 *trace;    // diagnostics timeline or null
    int stages;    // ring depth of the hot phase
    u32 ticket_rows; // uniques per cold ticket (<= 32; $HERALD_TICKET_ROWS)
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
    pdl_enter();
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
        if (a) {
            hl.very_hot[pa] = u;
            hl.done_a[pa] = 0;
        }
        u32 pb = rows_warp_append(&hl.ctrl[1], b);
        if (b) {
            hl.hot[pb] = u;
            hl.done_b[pb] = 0;
        }
    }
}

__device__ __forceinline__ void cp_async_f32(float *smem_dst, const float *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_16(float *smem_dst, const float *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// One hot work item: columns [q*W, q*W + W) of one hot row, all its occurrences in order.
// The CTA streams the W-float chunk of every occurrence through an S-deep cp.async ring
// ([S][kHotTileRows][W] floats); warp 0 adds them, one column per lane (lanes >= W idle).
// WIDE (rows 16 B aligned, D % 4 == 0): a thread copies 16 B, W/4 threads cover one occurrence
// (LDGSTS costs ~8 cycles per warp instruction whatever its width, so 16 B copies are what keeps
// the ring ahead of the adder); otherwise 4 B per thread and W == 32, one occurrence per warp
// instruction.
template <int W, bool WIDE, class F1>
__device__ __forceinline__ void hot_chunk(const F1 &f1, const typename F1::Ctx &ctx, float *s_ring,
                                          u32 *s_perm, u32 S, const u32 *__restrict__ perm,
                                          const float *__restrict__ vals, size_t D, u32 s0, u32 s1,
                                          u32 q) {
    static_assert(WIDE || W == 32, "the 4-byte copy path moves one 32-column chunk per warp");
    constexpr int RPW = kHotTileRows / kRowWarps;              // rows of a stage per warp (4 B path)
    constexpr int TPR = W / 4;                                 // threads per occurrence (16 B path)
    constexpr int RPI = kRowBlock / TPR;                       // occurrences per CTA-wide copy round
    constexpr int CPT = WIDE ? kHotTileRows / RPI : RPW;       // copies per thread per stage
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const u32 cnt = s1 - s0;
    const size_t col = (size_t)q * W + lane;
    const bool active = lane < (unsigned)W && col < D;
    decltype(f1.load(ctx, 0)) acc;
    if (warp == 0 && active)
        acc = f1.load(ctx, col);
    const u32 ntiles = (cnt + kHotTileRows - 1) / kHotTileRows;
    const u32 my_row = WIDE ? (threadIdx.x / TPR) : warp * RPW;       // first row it copies
    const u32 my_off = WIDE ? (threadIdx.x % TPR) * 4 : lane;         // float offset in the chunk
    const bool cp_active = (size_t)q * W + my_off < D;
    const float *my_src = vals + (size_t)q * W + my_off;
    // The occurrence indices (perm) travel through their own shared-memory ring of 2S tiles,
    // copied S - 1 tiles ahead of the data copies that read them and committed in the same
    // cp.async groups: a data copy never waits for an index load from global memory (with the
    // indices prefetched one tile ahead in registers, every tile paid one L2/HBM latency:
    // 10 ns per occurrence instead of the ~2.5 ns of the dependent FADD chain).
    const u32 PS = 2 * S;
    const u32 *seg_perm = perm + s0;
    auto copy_perm = [&](u32 tile, u32 slot) {
        const u32 i = tile * kHotTileRows + threadIdx.x;
        if (threadIdx.x < kHotTileRows && i < cnt)
            cp_async_f32(reinterpret_cast<float *>(s_perm + slot * kHotTileRows + threadIdx.x),
                         reinterpret_cast<const float *>(seg_perm + i));
    };
    u32 issue_slot = 0; // (next tile to issue) % S, kept without a division
    u32 pslot_r = 0;    // (next tile to issue) % 2S
    u32 next_tile = 0;
    auto issue = [&]() {
        float *stage = s_ring + (size_t)issue_slot * kHotTileRows * W;
        const u32 *pt = s_perm + pslot_r * kHotTileRows;
        issue_slot = issue_slot + 1 == S ? 0 : issue_slot + 1;
        pslot_r = pslot_r + 1 == PS ? 0 : pslot_r + 1;
        const u32 base = next_tile * kHotTileRows;
        next_tile++;
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            const u32 row = my_row + (WIDE ? RPI * j : j);
            if (base + row < cnt && cp_active) {
                const u32 pi = pt[row];
                if (WIDE)
                    cp_async_16(stage + row * W + my_off, my_src + (size_t)pi * D);
                else
                    cp_async_f32(stage + row * W + my_off, my_src + (size_t)pi * D);
            }
        }
    };
    // prologue: indices of tiles 0 .. 2S-3, then the data of tiles 0 .. S-2
    for (u32 t = 0; t + 2 < PS && t < ntiles; t++)
        copy_perm(t, t);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
#pragma unroll 1
    for (u32 k = 0; k < S - 1; k++) {
        issue();
        cp_async_commit();
    }
    u32 pslot_w = PS - 2; // slot of index tile k + 2S - 2
    u32 read_slot = 0;
    for (u32 k = 0; k < ntiles; k++) {
        // groups are committed one per tile, in order: at most S - 2 newer than tile k may still
        // be pending (wait_group takes an immediate, so the depth is switched; waiting for more
        // than necessary on a deeper ring is still correct)
        if (S >= 24)
            cp_async_wait<22>();
        else if (S >= 20)
            cp_async_wait<18>();
        else if (S >= 16)
            cp_async_wait<14>();
        else if (S >= 12)
            cp_async_wait<10>();
        else if (S >= 10)
            cp_async_wait<8>();
        else if (S >= 8)
            cp_async_wait<6>();
        else if (S >= 6)
            cp_async_wait<4>();
        else if (S == 5)
            cp_async_wait<3>();
        else if (S == 4)
            cp_async_wait<2>();
        else
            cp_async_wait<1>();
        __syncthreads(); // ... everyone's has; and stage k-1 has been consumed
        issue();         // tile k + S - 1 into the slot tile k - 1 occupied
        copy_perm(k + PS - 2, pslot_w); // its slot held tile k - 2, last read S + 1 iterations ago
        pslot_w = pslot_w + 1 == PS ? 0 : pslot_w + 1;
        cp_async_commit();
        if (warp == 0 && active) {
            const float *stage = s_ring + (size_t)read_slot * kHotTileRows * W + lane;
            const u32 rows = min((u32)kHotTileRows, cnt - k * kHotTileRows);
            if (rows == kHotTileRows) {
                // full tile: completely unrolled, so the shared-memory loads run ahead of the
                // dependent adds as far as the register file allows
                float g[kHotTileRows];
#pragma unroll
                for (int j = 0; j < kHotTileRows; j++)
                    g[j] = stage[j * W];
#pragma unroll
                for (int j = 0; j < kHotTileRows; j++)
                    acc = f1.step(acc, g[j]);
            } else {
                u32 r = 0;
                for (; r + 16 <= rows; r += 16) {
                    float g[16];
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        g[j] = stage[(r + j) * W];
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        acc = f1.step(acc, g[j]);
                }
                for (; r < rows; r++)
                    acc = f1.step(acc, stage[r * W]);
            }
        }
        read_slot = read_slot + 1 == S ? 0 : read_slot + 1;
    }
    cp_async_wait<0>();
    if (warp == 0 && active)
        f1.store(ctx, col, acc);
}

// FV: the functor instantiated for the cold path's vector width VEC; F1: the same functor for
// VEC = 1 (the hot path addresses single columns).
template <int VEC, int ROWS, class FV, class F1>
__global__ void __launch_bounds__(kRowBlock, 2)
    segment_reduce_kernel(const u32 *__restrict__ seg_start, const u32 *__restrict__ perm,
                          const u32 *__restrict__ num_unique, const float *__restrict__ vals,
                          size_t D, u32 hot_threshold, HotLists hl, FV fv, F1 f1) {
    pdl_enter();
    extern __shared__ __align__(16) float s_ring[]; // [stages][kHotTileRows][32], then the index ring
    __shared__ u32 s_item;
    const u32 S = (u32)hl.stages;
    u32 *const s_perm = reinterpret_cast<u32 *>(s_ring + (size_t)S * kHotTileRows * 32); // [4S][kHotTileRows]
    u64 *const trace = hl.trace;
    u32 items_taken = 0;
    if (trace && threadIdx.x == 0 && blockIdx.x < kTraceCtas) {
        trace[2 + 4 * blockIdx.x] = global_timer_ns();
        if (blockIdx.x == 0)
            trace[0] = gridDim.x;
    }
    using V = RowVec<VEC>;
    constexpr bool WIDE = VEC == 4; // rows are 16 B aligned and D % 4 == 0
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const u32 U = *num_unique;
    fv.kernel_begin();

    // Every CTA runs the hot items first, then the cold tickets.  (Measured alternative: letting one
    // CTA of every SM start with the cold tickets makes the chains 2.5x slower — a hot chain's copies
    // queue behind the cold neighbour's loads for the whole run instead of the second half — and
    // the kernel went from 89 us to 200 us.)
    // ------------------------------- hot phase -------------------------------------------
    if (hot_threshold != 0xffffffffu) {
        // very hot rows are cut into 16-column chunks when the rows allow 16 B copies: half the
        // bytes per occurrence in the ring = twice the occurrences in flight on the longest chains
        const u32 QB = (u32)((D + 31) / 32);
        const u32 QA = WIDE ? (u32)((D + 15) / 16) : QB;
        const u32 nA = hl.ctrl[0], nB = hl.ctrl[1];
        const u32 itemsA = nA * QA;
        const u32 total = itemsA + nB * QB;
        while (true) {
            __syncthreads(); // the ring and s_item of the previous item are no longer in use
            if (threadIdx.x == 0)
                s_item = atomicAdd(&hl.ctrl[2], 1u);
            __syncthreads();
            const u32 t = s_item;
            if (t >= total)
                break;
            items_taken++;
            if (trace && threadIdx.x == 0 && t < kTraceItems)
                trace[2 + 4 * (size_t)kTraceCtas + 3 * t] = global_timer_ns();
            const bool very = t < itemsA;
            const u32 h = very ? t / QA : nA + (t - itemsA) / QB;
            const u32 q = very ? t % QA : (t - itemsA) % QB;
            const u32 u = very ? hl.very_hot[h] : hl.hot[h - nA];
            u32 *done = very ? &hl.done_a[h] : &hl.done_b[h - nA];
            const u32 nchunks = very ? QA : QB;
            const u32 s0 = seg_start[u], s1 = seg_start[u + 1];
            const u32 cnt = s1 - s0;
            typename F1::Ctx ctx;
            const bool ok = f1.begin(u, cnt, ctx);
            if (ok) {
                if constexpr (WIDE) {
                    if (very)
                        hot_chunk<16, true>(f1, ctx, s_ring, s_perm, S * 2, perm, vals, D, s0, s1, q);
                    else
                        hot_chunk<32, true>(f1, ctx, s_ring, s_perm, S, perm, vals, D, s0, s1, q);
                } else {
                    hot_chunk<32, false>(f1, ctx, s_ring, s_perm, S, perm, vals, D, s0, s1, q);
                }
            }
            // the CTA that completes the row's last chunk applies the per-row scalars
            if (warp == 0) {
                u32 prev = 0;
                __syncwarp();
                if (lane == 0) {
                    __threadfence();
                    prev = atomicAdd(done, 1u);
                }
                prev = __shfl_sync(FULL, prev, 0);
                if (prev == nchunks - 1 && ok && lane == 0) {
                    __threadfence();
                    f1.end(ctx);
                }
            }
            if (trace && threadIdx.x == 0 && t < kTraceItems) {
                trace[2 + 4 * (size_t)kTraceCtas + 3 * t + 1] = global_timer_ns();
                trace[2 + 4 * (size_t)kTraceCtas + 3 * t + 2] = cnt;
            }
        }
        if (trace && threadIdx.x == 0 && blockIdx.x == 0)
            trace[1] = total;
    }
    if (trace && threadIdx.x == 0 && blockIdx.x < kTraceCtas) {
        trace[2 + 4 * blockIdx.x + 1] = global_timer_ns();
        trace[2 + 4 * blockIdx.x + 3] = items_taken;
    }

    // ------------------------------- cold phase ------------------------------------------
    // Ticket t covers the uniques t, t + T, t + 2T, ... (T tickets): ids that are hot tend to be
    // neighbours in key order, a strided ticket spreads their longer segments over many warps.
    const size_t nvec = D / VEC;
    const u32 TK = hl.ticket_rows;
    const u32 T = (U + TK - 1) / TK;
    while (true) {
        u32 t = 0;
        if (lane == 0)
            t = atomicAdd(&hl.ctrl[3], 1u);
        t = __shfl_sync(FULL, t, 0);
        if (t >= T)
            break;
        // lane-parallel metadata of the 32 uniques of this ticket
        const u32 my_u = t + lane * T;
        const bool mine = lane < TK && my_u < U;
        const auto my_pre = seg_peek(fv, mine ? my_u : 0u, 0);
        const u32 my_s0 = mine ? seg_start[my_u] : 0;
        const u32 my_cnt = mine ? seg_start[my_u + 1] - my_s0 : 0;
        const bool my_cold = my_cnt > 0 && my_cnt <= hot_threshold;
        const u32 my_p0 = my_cold ? perm[my_s0] : 0;
        const int nrows = __popc(__ballot_sync(FULL, mine)); // valid lanes are 0 .. nrows-1
        // every lane opens its own row: the dependent scalar loads of 32 rows overlap
        typename FV::Ctx my_ctx;
        const bool my_ok = my_cold && seg_begin(fv, my_pre, my_u, my_cnt, my_ctx, 0);
#pragma unroll 1
        for (int g0 = 0; g0 < nrows; g0 += ROWS) {
            typename FV::Ctx ctx[ROWS];
            u32 s0[ROWS], cnt[ROWS], p0[ROWS];
            bool ok[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
                const int from = (g0 + r) & 31;
                s0[r] = __shfl_sync(FULL, my_s0, from);
                cnt[r] = __shfl_sync(FULL, my_cnt, from);
                p0[r] = __shfl_sync(FULL, my_p0, from);
                ok[r] = g0 + r < nrows && __shfl_sync(FULL, my_ok, from);
                ctx[r] = shfl_pod(my_ctx, from);
            }
            // warp-uniform column loop (lanes beyond the row width idle): the shuffles below
            // need every lane
            for (size_t c0 = 0; c0 < nvec; c0 += 32) {
                const size_t c = c0 + lane;
                const bool cv = c < nvec;
                decltype(fv.load(ctx[0], 0)) acc[ROWS];
                typename V::T g[ROWS];
#pragma unroll
                for (int r = 0; r < ROWS; r++)
                    if (ok[r] && cv) {
                        acc[r] = fv.load(ctx[r], c);
                        g[r] = V::ld_nc(vals + (size_t)p0[r] * D + c * VEC);
                    }
#pragma unroll
                for (int r = 0; r < ROWS; r++)
                    if (ok[r] && cv)
                        acc[r] = fv.step(acc[r], g[r]);
                // further occurrences (few rows have any): 32 at a time, their source rows
                // fetched lane-parallel, four gradient rows in flight
#pragma unroll
                for (int r = 0; r < ROWS; r++) {
                    if (!ok[r] || cnt[r] < 2)
                        continue;
                    for (u32 b = 1; b < cnt[r]; b += 32) {
                        const u32 nb = min(32u, cnt[r] - b);
                        const u32 pl = lane < nb ? perm[s0[r] + b + lane] : 0;
                        for (u32 j = 0; j < nb; j += 4) {
                            typename V::T gg[4];
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                const u32 pi = __shfl_sync(FULL, pl, (j + k) & 31);
                                if (j + k < nb && cv)
                                    gg[k] = V::ld_nc(vals + (size_t)pi * D + c * VEC);
                            }
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                if (j + k < nb && cv)
                                    acc[r] = fv.step(acc[r], gg[k]);
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < ROWS; r++)
                    if (ok[r] && cv)
                        fv.store(ctx[r], c, acc[r]);
            }
        }
        if (my_ok)
            fv.end(my_ctx);
    }
    if (trace && blockIdx.x < kTraceCtas) {
        __syncthreads();
        if (threadIdx.x == 0)
            trace[2 + 4 * blockIdx.x + 2] = global_timer_ns();
    }
    fv.kernel_end();
}

// ---- one warp per listed row: f.begin(r); f.apply(r, c) per column chunk; f.end(r) -------------
template <int VEC, class F>
__global__ void __launch_bounds__(kRowBlock)
    foreach_row_kernel(size_t nrows, const u32 *__restrict__ nrows_dev, size_t D, F f) {
    pdl_enter();
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
