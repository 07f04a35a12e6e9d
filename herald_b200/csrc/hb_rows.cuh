// Row-granular device templates shared by the op-level entry points and the cache:
// one warp owns one embedding row at a time, lanes own 128-bit column chunks, so every
// global access of a row is a fully coalesced 16 B x 32 = 512 B (D = 128) transaction.
#pragma once

#include <algorithm>
#include <type_traits>

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
    static __device__ __forceinline__ T mul(const T &a, float s) { // exact fp32 products, never fused
        return make_float4(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s), __fmul_rn(a.w, s));
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
    static __device__ __forceinline__ void st_keep(float *p, const T &v) { // row the next kernels read again
        hb::st_keep(reinterpret_cast<float4 *>(p), v);
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
    static __device__ __forceinline__ T mul(const T &a, float s) {
        return __fmul_rn(a, s);
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
    static __device__ __forceinline__ void st_keep(float *p, const T &v) {
        *p = v;
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
        if (VEC == 4 && mine >= 0) // all 32 source rows of the group start moving towards L2 now
            prefetch_l2(src + (size_t)mine * D, (unsigned)(D * sizeof(float)));
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

__device__ __forceinline__ unsigned smem_u32(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}

// ---- gather through the bulk-copy engine (1-D TMA) --------------------------------------------
// dst[n,:] = src[index(n),:] with no register staging: a lane issues ONE bulk load of its source
// row into the warp's shared-memory stage (cp.async.bulk global -> shared, completion on an
// mbarrier) and, once the round's bytes have landed, ONE bulk store of the stage row to its
// destination (cp.async.bulk shared -> global, bulk-group completion).  Two stages per warp: the
// loads of round k + 1 are in flight while the stores of round k drain.  Rows must be multiples of
// 16 bytes at 16-byte aligned addresses; a negative index writes zeros with plain stores.
template <class Index>
__global__ void __launch_bounds__(kRowBlock)
    gather_bulk_kernel(const float *__restrict__ src, float *__restrict__ dst, size_t n, size_t D, Index index,
                       int rows_per_round) {
    pdl_enter();
    extern __shared__ __align__(128) float s_gstage[]; // [kRowWarps][2][R][D], then kRowWarps x 2 mbarriers
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const int R = rows_per_round;
    const unsigned row_bytes = (unsigned)(D * sizeof(float));
    float *stage0 = s_gstage + (size_t)warp * 2 * R * D;
    u64 *bars = reinterpret_cast<u64 *>(s_gstage + (size_t)kRowWarps * 2 * R * D) + warp * 2;
    if (lane < 2)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + lane)) : "memory");
    __syncthreads();
    unsigned phase[2] = {0u, 0u};
    const size_t warp_global = (size_t)blockIdx.x * kRowWarps + warp;
    const size_t nwarps = (size_t)gridDim.x * kRowWarps;
    int st = 0; // stage of the next round
    for (size_t base = warp_global * 32; base < n; base += nwarps * 32) {
        const long long mine = base + lane < n ? index(base + lane) : -1;
        const int rows_here = (int)min((size_t)32, n - base);
        for (int r0 = 0; r0 < rows_here; r0 += R) {
            const int nr = min(R, rows_here - r0);
            float *stage = stage0 + (size_t)st * R * D;
            const unsigned bar_a = smem_u32(bars + st);
            // the stage's previous stores (two rounds ago) must have finished READING it
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            const bool in_round = (int)lane >= r0 && (int)lane < r0 + nr;
            const unsigned valid = __ballot_sync(FULL, in_round && mine >= 0);
            if (lane == 0 && valid)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a),
                             "r"(row_bytes * (unsigned)__popc(valid))
                             : "memory");
            __syncwarp();
            if (in_round && mine >= 0)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(stage + (size_t)(lane - r0) * D)),
                             "l"(src + (size_t)mine * D), "r"(row_bytes), "r"(bar_a)
                             : "memory");
            if (valid) {
                unsigned done = 0;
                while (!done)
                    asm volatile("{\n\t.reg .pred p;\n\t"
                                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                                 "selp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done)
                                 : "r"(bar_a), "r"(phase[st])
                                 : "memory");
                phase[st] ^= 1u;
            }
            if (in_round) {
                float *drow = dst + (base + lane) * D;
                if (mine >= 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(drow),
                                 "r"(smem_u32(stage + (size_t)(lane - r0) * D)), "r"(row_bytes)
                                 : "memory");
                } else {
                    for (size_t k = 0; k < D; k += 4)
                        *reinterpret_cast<float4 *>(drow + k) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            st ^= 1;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // the stores must be complete before the grid ends
}

// ---- segment reduce by unique key, applied to a destination row -----------------------
// Functor contract (instantiated for the cold path's vector width VEC and, as F1, for VEC = 1):
//   bool open(u, cnt, Ctx &)   ONCE per unique key, by one thread of the plan kernel: reads (and
//                              may write) the row's scalars and leaves in Ctx (<= 16 B, trivially
//                              copyable) everything the data phase needs
//   per column chunk c:  acc = load(ctx, c);
//                        for each occurrence p in ascending original index:
//                            acc = step(acc, vals[perm[p], c]);
//                        store(ctx, c, acc);
//   kernel_begin() / kernel_end()   block-level hooks of the plan kernel (counters)
// The adds happen in occurrence order, so the result is deterministic and bit-identical to a
// serial CPU loop over the batch (Line::accumulate order, src/hetu_cache/include/embedding.h:78-91).
//
// Two kernels:
//   seg_plan_kernel  one thread per unique: open() + one 32-byte work item {first sorted position,
//       occurrences, first occurrence's source row, class, Ctx}, written in TICKET order (ticket t
//       holds the uniques t, t + T, t + 2T, ...: ids that are hot tend to be neighbours in key
//       order, a strided ticket spreads them), and the work lists of the long segments.  The data
//       kernel reads its metadata with ONE coalesced load per ticket instead of three dependent
//       levels of scattered ones.
//   segment_reduce_kernel  persistent, warp-specialised:
//     * hot group (warps 0 .. kHotWarps-1 of every CTA) — a Zipf batch has ids that occur thousands
//       of times and the add ORDER is fixed by the parity contract, so such a segment cannot be split
//       by rows; it is split by COLUMNS.  A work item is (hot row, 32-float column chunk) — 16-float
//       chunks for the very hot rows.  The producer warps stream the chunk of every occurrence into
//       a shared-memory ring with cp.async (the occurrence indices prefetched in registers one
//       turn ahead) and signal each stage's `full` mbarrier when their copies have landed
//       (cp.async.mbarrier.arrive); warp 0 only adds: it waits on `full`, runs the dependent FADD
//       chain out of shared memory (loads one batch ahead of the adds, the next stage's barrier
//       tested three batches ahead) and releases the stage through its `empty` mbarrier.  No
//       CTA-wide barrier anywhere in the chain.
//     * every other warp, and the hot group once the hot items are gone: first the MEDIUM rows
//       (more than `med` occurrences: one warp per row, eight gradient rows in flight), then the
//       cold tickets, kColdRows rows at a time with the row / gradient / owner-row loads of all of
//       them in flight together (128-bit per lane); the next ticket's number and work items are
//       fetched while the current one is processed.
constexpr int kHotWarps = 3;                   // 1 adder + 2 producers
constexpr int kHotProducers = kHotWarps - 1;
constexpr int kHotStageFloats = 2048;          // 8 KB per ring stage: 128 occurrences x 16 columns
constexpr int kHotStageBytes = kHotStageFloats * 4; // or 64 occurrences x 32 columns
constexpr int kHotStagesDefault = 6;
constexpr int kHotStagesMax = 12;
// dynamic shared memory of segment_reduce_kernel: the ring, then full[S] / empty[S] mbarriers
constexpr size_t hot_smem_bytes(int stages) {
    return (size_t)stages * kHotStageBytes + (size_t)2 * stages * 8 + 16;
}
constexpr u32 kVeryHot = 1024; // rows above this go first (longest-processing-time-first)
// Two-level reduction of the very hot rows (opt-in, HotLists::split).  The add ORDER of a row is
// part of the bit-parity contract, so a row with 11 000 occurrences is one 11 000-long dependent
// FADD chain — at C2 that single chain, not HBM, bounds the whole kernel (70 of 78 us).  With the
// split on, a row with more than kVeryHot occurrences is cut into kSplitTiles runs of
// L = round_up(ceil(cnt / kSplitTiles), 128) consecutive occurrences; every run is summed in
// occurrence order from 0 (independent chains, any CTA), then the run sums are added to the row in
// run order.  The result depends on cnt only — deterministic, identical on every machine — but it is
// not the serial order: updated rows then agree with the reference within fp32 re-association
// (1e-5 relative, the north star's tolerance) instead of bit for bit.  Every other row keeps the
// exact order.
constexpr u32 kSplitTiles = 8;
__host__ __device__ inline u32 split_run_length(u32 cnt) {
    return (((cnt + kSplitTiles - 1) / kSplitTiles) + 127u) & ~127u;
}
constexpr u32 kMediumDefault = 4;

// ring depth of the hot phase ($HERALD_HOT_STAGES, 2 .. 12): the bytes the producers keep in
// flight are what hides HBM latency under the dependent add chain
int hot_stages();
// uniques per cold-phase ticket ($HERALD_TICKET_ROWS, 4 .. 32)
u32 ticket_rows();
// segments longer than this (and not hot) are taken one warp per row ($HERALD_MED_THRESHOLD)
u32 medium_threshold();

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

// classes of a work item
enum : u32 { SEG_SKIP = 0, SEG_COLD = 1, SEG_LISTED = 2 };

// One unique key's work item, 32 bytes (two 128-bit loads per lane).
struct __align__(16) SegItemHead {
    u32 s0;  // first sorted position
    u32 cnt; // occurrences
    u32 p0;  // perm[s0]: source row of the first occurrence
    u32 cls;
};
template <class Ctx>
struct __align__(16) SegItem {
    SegItemHead h;
    Ctx ctx;
    static_assert(sizeof(Ctx) <= 16, "segment contexts travel in 16 bytes");
};
constexpr size_t kSegItemBytes = 32;

struct HotLists {
    u32 *very_hot; // [cap] item indices of the rows with count > kVeryHot
    u32 *hot;      // [cap] ... hot_threshold < count <= kVeryHot
    u32 *medium;   // [cap] ... med_threshold < count <= hot_threshold
    u32 *ctrl;     // [0] = #very_hot, [1] = #hot, [2] = hot ticket, [3] = cold ticket, [4] = #medium,
                   // [5] = medium ticket (zeroed with the scan arena)
    void *items;   // [cap + 32] work items in ticket order
    u64 *trace;    // diagnostics timeline or null
    int stages;    // ring depth of the hot phase
    u32 ticket_rows; // uniques per cold ticket (<= 32)
    u32 med_threshold;
    u32 mode;        // diagnostics ($HERALD_SEG_MODE): bit 0 = cold work waits for every hot group
    u32 split;       // two-level reduction of the very hot rows (functors that opt in)
    float *partials; // [very hot row][kSplitTiles][D] run sums
    u32 *split_done; // [very hot row] runs x column chunks finished (zero on entry)
};

// A functor opts in to the split with `static constexpr bool kSplit = true` and two members:
// pre(g) = what one gradient value contributes (its scaled value), step_pre(acc, p) = step() for an
// already-scaled contribution.
template <class F, class = void>
struct can_split : std::false_type {};
template <class F>
struct can_split<F, std::enable_if_t<F::kSplit>> : std::true_type {};

// run sum of one tile: ((0 + pre(g0)) + pre(g1)) + ... into partials
template <class F1>
struct RunSum {
    const F1 &f;
    float *dst; // the run's row of partials
    struct Ctx {};
    __device__ float load(const Ctx &, size_t) const {
        return 0.f;
    }
    __device__ float step(float acc, float g) const {
        return __fadd_rn(acc, f.pre(g));
    }
    __device__ void store(const Ctx &, size_t col, float acc) const {
        dst[col] = acc;
    }
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

// The work item of unique `u` at position `i` of the ticket order (a padding item when u >= U or
// i >= total).  Warp-collective: every lane of a full warp calls it exactly once per round (the list
// appends use ballots).  Shared by seg_plan_kernel and by kernels that plan while they still hold the
// unique in registers (the cache's resolve, hb_cache.cu).
template <class F>
__device__ __forceinline__ void seg_plan_item(u32 i, u32 u, u32 U, u32 total, const u32 *__restrict__ seg_start,
                                              const u32 *__restrict__ perm, u32 hot_threshold, const HotLists &hl,
                                              const F &f) {
    using Item = SegItem<typename F::Ctx>;
    static_assert(sizeof(Item) == kSegItemBytes, "work items are 32 bytes");
    Item *items = reinterpret_cast<Item *>(hl.items);
    const u32 med = min(hl.med_threshold, hot_threshold);
    Item item;
    item.h.s0 = item.h.cnt = item.h.p0 = 0;
    item.h.cls = SEG_SKIP;
    item.ctx = typename F::Ctx();
    if (i < total && u < U) {
        item.h.s0 = seg_start[u];
        item.h.cnt = seg_start[u + 1] - item.h.s0;
        if (item.h.cnt > 0 && f.open((size_t)u, item.h.cnt, item.ctx)) {
            item.h.p0 = perm[item.h.s0];
            item.h.cls = item.h.cnt > med ? SEG_LISTED : SEG_COLD;
        }
    }
    if (i < total)
        items[i] = item;
    const bool listed = item.h.cls == SEG_LISTED;
    const bool a = listed && item.h.cnt > hot_threshold && item.h.cnt > kVeryHot;
    const bool b = listed && !a && item.h.cnt > hot_threshold;
    const bool m = listed && !a && !b;
    const u32 pa = rows_warp_append(&hl.ctrl[0], a);
    if (a)
        hl.very_hot[pa] = i;
    const u32 pb = rows_warp_append(&hl.ctrl[1], b);
    if (b)
        hl.hot[pb] = i;
    const u32 pm = rows_warp_append(&hl.ctrl[4], m);
    if (m)
        hl.medium[pm] = i;
}

template <class F>
__global__ void __launch_bounds__(256)
    seg_plan_kernel(const u32 *__restrict__ seg_start, const u32 *__restrict__ perm,
                    const u32 *__restrict__ num_unique, u32 hot_threshold, HotLists hl, F f) {
    pdl_enter();
    const u32 U = *num_unique;
    const u32 TK = hl.ticket_rows;
    const u32 T = (U + TK - 1) / TK;
    const u32 total = T * TK;
    f.kernel_begin();
    const u32 stride = gridDim.x * blockDim.x;
    const u32 rounds = (total + stride - 1) / stride;
    for (u32 it = 0; it < rounds; it++) { // block-uniform trip count: the appends are warp-collective
        const u32 i = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const u32 t = i / TK, l = i - t * TK;
        seg_plan_item(i, t + l * T, U, total, seg_start, perm, hot_threshold, hl, f);
    }
    f.kernel_end();
}

__device__ __forceinline__ void cp_async_f32(float *smem_dst, const float *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_16(float *smem_dst, const float *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

// ---- mbarrier (shared::cta) -------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(u64 *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64 *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the executing thread's arrival happens when all its earlier cp.async copies have landed; the
// barrier's expected count already includes it (.noinc)
__device__ __forceinline__ void mbar_arrive_on_copies(u64 *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(u64 *bar, unsigned parity) { // non-blocking
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(u64 *bar, unsigned parity) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "HB_WAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra HB_DONE_%=;\n\t"
                 "bra HB_WAIT_%=;\n\t"
                 "HB_DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
// named barrier of the hot group (barrier 0 is __syncthreads)
__device__ __forceinline__ void hot_group_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(kHotWarps * 32) : "memory");
}

// ring stage / phase parity of the CTA's kt-th hot tile
struct HotRing {
    float *ring;
    u64 *full, *empty;
    u32 S;
    u32 stage, phase; // of the next tile
    __device__ __forceinline__ void advance() {
        if (++stage == S) {
            stage = 0;
            phase ^= 1u;
        }
    }
    __device__ __forceinline__ void advance(u32 n) { // n < 2 * S steps at a time are enough here
        for (u32 i = 0; i < n; i++)
            advance();
    }
};

// Producer warp `p` of kHotProducers: tiles p, p + P, ... of one hot item (columns [q*W, q*W + W)
// of every occurrence).  WIDE (rows 16 B aligned, D % 4 == 0): a thread copies 16 B, W/4 threads
// cover one occurrence (LDGSTS costs ~8 cycles per warp instruction whatever its width);
// otherwise 4 B per thread and W == 32, one occurrence per warp instruction.
template <int W, bool WIDE>
__device__ __forceinline__ void hot_produce(HotRing hr, u32 p, const u32 *__restrict__ seg_perm,
                                            const float *__restrict__ vals, size_t D, u32 cnt, u32 q) {
    static_assert(WIDE || W == 32, "the 4-byte copy path moves one 32-column chunk per warp");
    constexpr int TILE = kHotStageFloats / W;  // occurrences per stage
    constexpr int TPR = WIDE ? W / 4 : 32;     // threads per occurrence
    constexpr int OPI = 32 / TPR;              // occurrences per warp instruction
    constexpr int NI = TILE / OPI;             // copy instructions per tile
    constexpr int NR = TILE / 32;              // index registers per lane and tile
    const unsigned lane = lane_id();
    const u32 sub = lane / TPR;
    const u32 off = WIDE ? (lane % TPR) * 4 : lane; // float offset inside the chunk
    const bool col_ok = (size_t)q * W + off < D;
    const float *src = vals + (size_t)q * W + off;
    const u32 ntiles = (cnt + TILE - 1) / TILE;
    hr.advance(p);
    // The occurrence indices of this warp's next tile are loaded one turn ahead into a second
    // register set; the two sets swap roles by unrolling the turn loop twice (a register move
    // from the set being loaded would wait for the load: one memory latency per tile).
    auto load_idx = [&](u32 (&idx)[NR], u32 k) {
#pragma unroll
        for (int j = 0; j < NR; j++) {
            const u32 o = k * TILE + j * 32 + lane;
            idx[j] = (k < ntiles && o < cnt) ? seg_perm[o] : 0;
        }
    };
    auto fill = [&](const u32 (&idx)[NR], u32 k) {
        mbar_wait(&hr.empty[hr.stage], hr.phase ^ 1u);
        float *stage = hr.ring + (size_t)hr.stage * kHotStageFloats;
        const u32 base = k * TILE;
#pragma unroll
        for (int i = 0; i < NI; i++) {
            const u32 o = i * OPI + sub;
            const u32 pi = __shfl_sync(FULL, idx[(i * OPI) / 32], ((i * OPI) % 32) + sub);
            if (base + o < cnt && col_ok) {
                if (WIDE)
                    cp_async_16(stage + o * W + off, src + (size_t)pi * D);
                else
                    cp_async_f32(stage + o * W + off, src + (size_t)pi * D);
            }
        }
        mbar_arrive_on_copies(&hr.full[hr.stage]);
        hr.advance(kHotProducers);
    };
    u32 ia[NR], ib[NR];
    u32 k = p;
    load_idx(ia, k);
    while (k < ntiles) {
        load_idx(ib, k + kHotProducers);
        fill(ia, k);
        k += kHotProducers;
        if (k >= ntiles)
            break;
        load_idx(ia, k + kHotProducers);
        fill(ib, k);
        k += kHotProducers;
    }
}

// The adder warp: one column per lane, every occurrence in order.  Lanes >= W mirror lanes < W
// (same shared-memory words, results dropped): the chain runs without a divergent region.
// A full stage is consumed by fully unrolled code with a rolling window: value j + kAhead is
// loaded right before value j is added, so the loads stay kAhead occurrences in front of the
// dependent FADD chain and no register is copied.
template <int W, class F1>
__device__ __forceinline__ void hot_add(HotRing hr, const F1 &f1, const typename F1::Ctx &ctx, size_t D,
                                        u32 cnt, u32 q, u64 &waited) {
    constexpr int TILE = kHotStageFloats / W;
    constexpr int kAhead = 16;
    static_assert(TILE >= 4 * kAhead, "the next stage is tested three windows before the end");
    const unsigned lane = lane_id();
    const unsigned rl = lane & (W - 1);
    const size_t col = (size_t)q * W + lane;
    const bool active = lane < (unsigned)W && col < D;
    const size_t lcol = (size_t)q * W + rl;
    decltype(f1.load(ctx, 0)) acc = f1.load(ctx, lcol < D ? lcol : (size_t)q * W);
    const u32 ntiles = (cnt + TILE - 1) / TILE;
    {
        const long long c0 = clock64();
        mbar_wait(&hr.full[hr.stage], hr.phase);
        waited += (u64)(clock64() - c0);
    }
    for (u32 k = 0; k < ntiles; k++) {
        const float *st = hr.ring + (size_t)hr.stage * kHotStageFloats + rl;
        u64 *const my_empty = &hr.empty[hr.stage];
        const u32 rows = min((u32)TILE, cnt - k * TILE);
        hr.advance();
        const bool more = k + 1 < ntiles;
        bool next_ready = false;
        if (rows == TILE) {
            float v[TILE];
#pragma unroll
            for (int j = 0; j < kAhead; j++)
                v[j] = st[j * W];
#pragma unroll
            for (int j = 0; j < TILE; j++) {
                if (j + kAhead < TILE)
                    v[j + kAhead] = st[(j + kAhead) * W];
                if (j == TILE - 3 * kAhead && more) // its result is back before the last add
                    next_ready = mbar_test(&hr.full[hr.stage], hr.phase);
                acc = f1.step(acc, v[j]);
            }
        } else {
            u32 r = 0;
            for (; r + kAhead <= rows; r += kAhead) {
                float g[kAhead];
#pragma unroll
                for (int j = 0; j < kAhead; j++)
                    g[j] = st[(r + j) * W];
#pragma unroll
                for (int j = 0; j < kAhead; j++)
                    acc = f1.step(acc, g[j]);
            }
            for (; r < rows; r++)
                acc = f1.step(acc, st[r * W]);
        }
        // every value of the stage has been consumed (the adds above depend on the loads)
        __syncwarp();
        if (lane == 0)
            mbar_arrive(my_empty);
        if (more && !next_ready) {
            const long long c0 = clock64();
            mbar_wait(&hr.full[hr.stage], hr.phase);
            waited += (u64)(clock64() - c0);
        }
    }
    if (active)
        f1.store(ctx, col, acc);
}

// FV: the functor instantiated for the cold path's vector width VEC; F1: the same functor for
// VEC = 1 (the hot path addresses single columns).
template <int VEC, int ROWS, class FV, class F1>
__global__ void __launch_bounds__(kRowBlock, 2)
    segment_reduce_kernel(const u32 *__restrict__ perm, const u32 *__restrict__ num_unique,
                          const float *__restrict__ vals, size_t D, u32 hot_threshold, HotLists hl,
                          FV fv, F1 f1) {
    pdl_enter();
    extern __shared__ __align__(16) float s_ring[]; // [stages][kHotStageFloats], then the mbarriers
    __shared__ u32 s_item;
    using Item = SegItem<typename FV::Ctx>;
    static_assert(sizeof(typename FV::Ctx) == sizeof(typename F1::Ctx), "one context type");
    const Item *__restrict__ items = reinterpret_cast<const Item *>(hl.items);
    const u32 S = (u32)hl.stages;
    u64 *const s_full = reinterpret_cast<u64 *>(s_ring + (size_t)S * kHotStageFloats);
    u64 *const s_empty = s_full + S;
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
    const bool hot_on = hot_threshold != 0xffffffffu;

    // ------------------------------- hot group --------------------------------------------
    if (hot_on) {
        if (threadIdx.x == 0) {
            for (u32 i = 0; i < S; i++) {
                mbar_init(&s_full[i], 32); // the 32 lanes of the producer warp that fills the stage
                mbar_init(&s_empty[i], 1); // the adder's lane 0
            }
        }
        __syncthreads();
    }
    if (hot_on && warp < (unsigned)kHotWarps) {
        // very hot rows are cut into 16-column chunks when the rows allow 16 B copies: half the
        // bytes per occurrence in the ring = twice the occurrences in flight on the longest chains
        const u32 QB = (u32)((D + 31) / 32);
        const u32 QA = WIDE ? (u32)((D + 15) / 16) : QB;
        const u32 nA = hl.ctrl[0], nB = hl.ctrl[1];
        constexpr bool kCanSplit = can_split<F1>::value;
        const bool split = kCanSplit && hl.split != 0;
        // item space.  exact: [very hot rows x QA chunks][hot rows x QB chunks]
        //              split: [very hot rows x kSplitTiles runs x QA chunks][hot rows x QB][very hot rows x QB combines]
        const u32 itemsA = split ? nA * kSplitTiles * QA : nA * QA;
        const u32 itemsB = nB * QB;
        const u32 itemsC = split ? nA * QB : 0u;
        const u32 total = itemsA + itemsB + itemsC;
        HotRing hr{s_ring, s_full, s_empty, S, 0u, 0u};
        while (true) {
            hot_group_sync(); // s_item of the previous item has been read by everyone
            if (threadIdx.x == 0)
                s_item = atomicAdd(&hl.ctrl[2], 1u);
            hot_group_sync();
            const u32 t = s_item;
            if (t >= total)
                break;
            items_taken++;
            if (trace && threadIdx.x == 0 && t < kTraceItems)
                trace[2 + 4 * (size_t)kTraceCtas + 3 * t] = global_timer_ns();
            const bool very = t < itemsA;
            const bool combine = t >= itemsA + itemsB;
            u32 h, q, run = 0;
            if (very && split) { // runs outer, chunks inner: neighbouring tickets share gradient rows
                h = t / (kSplitTiles * QA);
                const u32 r = t - h * (kSplitTiles * QA);
                run = r / QA;
                q = r - run * QA;
            } else if (very) {
                h = t / QA;
                q = t % QA;
            } else if (combine) {
                h = (t - itemsA - itemsB) / QB;
                q = (t - itemsA - itemsB) % QB;
            } else {
                h = (t - itemsA) / QB;
                q = (t - itemsA) % QB;
            }
            const Item it = items[(very || combine) ? hl.very_hot[h] : hl.hot[h]];
            u32 cnt = it.h.cnt;
            u32 first = 0; // first occurrence of the chain this item runs
            if (very && split) {
                const u32 L = split_run_length(cnt);
                first = min(run * L, cnt);
                cnt = min(L, cnt - first);
            }
            const u32 W = (very && WIDE) ? 16u : 32u;
            u32 ntiles = (cnt + kHotStageFloats / W - 1) / (kHotStageFloats / W);
            u64 waited = 0;
            if (combine) {
                ntiles = 0; // nothing travels through the ring
                if constexpr (kCanSplit) {
                    if (warp == 0) {
                        // every run of this row has been summed (the run items precede the combines in
                        // ticket order and every earlier ticket is held by a resident group)
                        const u32 need = kSplitTiles * QA;
                        while (*reinterpret_cast<volatile u32 *>(&hl.split_done[h]) < need)
                            __nanosleep(100);
                        __threadfence();
                        typename F1::Ctx ctx;
                        memcpy(&ctx, &it.ctx, sizeof(ctx));
                        const size_t col = (size_t)q * 32 + lane;
                        if (col < D) {
                            auto acc = f1.load(ctx, col);
                            const u32 L = split_run_length(it.h.cnt);
#pragma unroll
                            for (u32 r = 0; r < kSplitTiles; r++)
                                if (r * L < it.h.cnt) // (an empty run contributes nothing, not even + 0)
                                    acc = f1.step_pre(acc, __ldcg(hl.partials + ((size_t)h * kSplitTiles + r) * D + col));
                            f1.store(ctx, col, acc);
                        }
                    }
                }
            } else if (very && split) {
                if constexpr (kCanSplit) {
                    if (cnt) {
                        constexpr int WS = WIDE ? 16 : 32;
                        if (warp == 0) {
                            const RunSum<F1> rs{f1, hl.partials + ((size_t)h * kSplitTiles + run) * D};
                            typename RunSum<F1>::Ctx rc;
                            hot_add<WS>(hr, rs, rc, D, cnt, q, waited);
                        } else {
                            hot_produce<WS, WIDE>(hr, warp - 1, perm + it.h.s0 + first, vals, D, cnt, q);
                        }
                    }
                    if (warp == 0) {
                        __syncwarp();
                        __threadfence();
                        if (lane == 0)
                            atomicAdd(&hl.split_done[h], 1u);
                    }
                }
            } else if (warp == 0) {
                typename F1::Ctx ctx;
                memcpy(&ctx, &it.ctx, sizeof(ctx));
                if (very && WIDE)
                    hot_add<16>(hr, f1, ctx, D, cnt, q, waited);
                else
                    hot_add<32>(hr, f1, ctx, D, cnt, q, waited);
            } else {
                const u32 *sp = perm + it.h.s0;
                if constexpr (WIDE) {
                    if (very)
                        hot_produce<16, true>(hr, warp - 1, sp, vals, D, cnt, q);
                    else
                        hot_produce<32, true>(hr, warp - 1, sp, vals, D, cnt, q);
                } else {
                    hot_produce<32, false>(hr, warp - 1, sp, vals, D, cnt, q);
                }
            }
            // every warp of the group moves its view of the ring past this item's tiles
            for (u32 i = 0; i < ntiles % (2 * S); i++)
                hr.advance();
            if (trace && threadIdx.x == 0 && t < kTraceItems) {
                trace[2 + 4 * (size_t)kTraceCtas + 3 * t + 1] = global_timer_ns();
                trace[2 + 4 * (size_t)kTraceCtas + 3 * t + 2] = (u64)cnt | ((waited >> 4) << 32);
            }
        }
        if ((hl.mode & 1u) && threadIdx.x == 0) {
            __threadfence();
            atomicAdd(&hl.ctrl[6], 1u);
        }
        if (trace && threadIdx.x == 0 && blockIdx.x == 0)
            trace[1] = total;
        if (trace && threadIdx.x == 0 && blockIdx.x < kTraceCtas) {
            trace[2 + 4 * blockIdx.x + 1] = global_timer_ns();
            trace[2 + 4 * blockIdx.x + 3] = items_taken;
        }
    }

    if (hot_on && (hl.mode & 1u)) { // diagnostics: the hot chains run on an otherwise idle chip
        while (*reinterpret_cast<volatile u32 *>(&hl.ctrl[6]) < gridDim.x)
            __nanosleep((hl.mode & 2u) ? 20000 : 200);
    }
    const size_t nvec = D / VEC;
    // ------------------------------- medium rows ------------------------------------------
    // one warp per row, eight gradient rows in flight; taken before the cold tickets so that the
    // longest serial pieces of the kernel start first
    {
        const u32 nM = hl.ctrl[4];
        while (nM) {
            u32 t = 0;
            if (lane == 0)
                t = atomicAdd(&hl.ctrl[5], 1u);
            t = __shfl_sync(FULL, t, 0);
            if (t >= nM)
                break;
            const Item it = items[hl.medium[t]];
            const u32 s0 = it.h.s0, cnt = it.h.cnt;
            for (size_t c0 = 0; c0 < nvec; c0 += 32) {
                const size_t c = c0 + lane;
                const bool cv = c < nvec;
                decltype(fv.load(it.ctx, 0)) acc;
                if (cv)
                    acc = fv.load(it.ctx, c);
                for (u32 b = 0; b < cnt; b += 32) {
                    const u32 nb = min(32u, cnt - b);
                    const u32 pl = lane < nb ? perm[s0 + b + lane] : 0;
                    for (u32 j = 0; j < nb; j += 8) {
                        typename V::T gg[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const u32 pi = __shfl_sync(FULL, pl, (j + k) & 31);
                            if (j + k < nb && cv)
                                gg[k] = V::ld_nc(vals + (size_t)pi * D + c * VEC);
                        }
#pragma unroll
                        for (int k = 0; k < 8; k++)
                            if (j + k < nb && cv)
                                acc = fv.step(acc, gg[k]);
                    }
                }
                if (cv)
                    fv.store(it.ctx, c, acc);
            }
        }
    }

    // ------------------------------- cold tickets ------------------------------------------
    // Ticket t = work items [t*TK, t*TK + TK), one per lane.  The ticket after the next is drawn
    // and the next ticket's items are loaded while the current ticket is processed.
    {
        const u32 TK = hl.ticket_rows;
        const u32 T = (U + TK - 1) / TK;
        constexpr int XPR = 32 / ROWS; // further occurrences per row whose indices one load covers
        auto draw = [&]() {
            u32 t = 0;
            if (lane == 0)
                t = atomicAdd(&hl.ctrl[3], 1u);
            return t; // lane 0's value counts; broadcast where it is used
        };
        auto fetch = [&](u32 t) {
            Item it;
            it.h.s0 = it.h.cnt = it.h.p0 = 0;
            it.h.cls = SEG_SKIP;
            if (t < T && lane < TK)
                it = items[(size_t)t * TK + lane];
            return it;
        };
        u32 t_cur = __shfl_sync(FULL, draw(), 0);
        Item it_cur = fetch(t_cur);
        u32 t_next_raw = draw();
        while (t_cur < T) {
            const u32 t_next = __shfl_sync(FULL, t_next_raw, 0);
            const Item it_next = fetch(t_next);
            t_next_raw = draw();
            const Item my = it_cur;
            const bool my_cold = my.h.cls == SEG_COLD;
            const unsigned cold_mask = __ballot_sync(FULL, my_cold);
#pragma unroll 1
            for (u32 g0 = 0; g0 < TK; g0 += ROWS) {
                if (((cold_mask >> g0) & ((1u << ROWS) - 1u)) == 0)
                    continue;
                typename FV::Ctx ctx[ROWS];
                u32 cnt[ROWS], p0[ROWS];
                bool ok[ROWS];
                u32 max_cnt = 0;
#pragma unroll
                for (int r = 0; r < ROWS; r++) {
                    const int from = (g0 + r) & 31;
                    ok[r] = (cold_mask >> from) & 1u;
                    const u32 rc = __shfl_sync(FULL, my.h.cnt, from);
                    cnt[r] = ok[r] ? rc : 0;
                    p0[r] = __shfl_sync(FULL, my.h.p0, from);
                    ctx[r] = shfl_pod(my.ctx, from);
                    max_cnt = max(max_cnt, cnt[r]);
                }
                // source rows of the further occurrences (few rows have any): lane r*XPR + j holds
                // the one of occurrence 1 + j of row r; issued together with the row loads
                const int xr = (int)(lane / XPR), xj = (int)(lane % XPR);
                const int xfrom = (g0 + xr) & 31;
                const u32 xs0 = __shfl_sync(FULL, my.h.s0, xfrom);
                const u32 xc = __shfl_sync(FULL, my.h.cnt, xfrom);
                const u32 xcnt = ((cold_mask >> xfrom) & 1u) ? xc : 0;
                u32 xi = 0;
                if (max_cnt > 1 && 1u + xj < xcnt)
                    xi = perm[xs0 + 1 + xj];
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
                    // further occurrences: one of every row per round, all rows' loads in flight
                    for (u32 jb = 1; jb < max_cnt; jb += XPR) {
                        u32 xcur = xi;
                        if (jb > 1) // (only with a medium threshold above XPR)
                            xcur = jb + xj < xcnt ? perm[xs0 + jb + xj] : 0;
                        const u32 jend = min(max_cnt, jb + (u32)XPR);
                        for (u32 j = jb; j < jend; j++) {
#pragma unroll
                            for (int r = 0; r < ROWS; r++) {
                                const u32 pi = __shfl_sync(FULL, xcur, r * XPR + (int)(j - jb));
                                if (j < cnt[r] && cv)
                                    g[r] = V::ld_nc(vals + (size_t)pi * D + c * VEC);
                            }
#pragma unroll
                            for (int r = 0; r < ROWS; r++)
                                if (j < cnt[r] && cv)
                                    acc[r] = fv.step(acc[r], g[r]);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < ROWS; r++)
                        if (ok[r] && cv)
                            fv.store(ctx[r], c, acc[r]);
                }
            }
            t_cur = t_next;
            it_cur = it_next;
        }
    }
    if (trace && blockIdx.x < kTraceCtas) {
        __syncthreads();
        if (threadIdx.x == 0) {
            trace[2 + 4 * blockIdx.x + 2] = global_timer_ns();
            if (!hot_on)
                trace[2 + 4 * blockIdx.x + 1] = trace[2 + 4 * blockIdx.x];
        }
    }
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
