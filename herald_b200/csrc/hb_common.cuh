// Shared host/device helpers for libherald_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <stdexcept>
#include <string>
#include <utility>

#include "../../include/herald_b200.h"

namespace hb {

using u8 = uint8_t;
using u32 = uint32_t;
using i32 = int32_t;
using u64 = unsigned long long; // matches CUDA atomics' 64-bit type
using i64 = long long;

// ---- error plumbing --------------------------------------------------------
void set_last_error(const std::string &msg);
extern std::atomic<uint64_t> g_launches;

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define HB_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            throw hb::Error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " + \
                            __FILE__ + ":" + std::to_string(__LINE__));                    \
    } while (0)

#define HB_CHECK(cond, msg)                                                                \
    do {                                                                                   \
        if (!(cond))                                                                       \
            throw hb::Error(std::string("check failed: ") + #cond + ": " + (msg));         \
    } while (0)

#define HB_API_BEGIN() try {
#define HB_API_END()                                                                       \
    }                                                                                      \
    catch (const std::exception &_ex) {                                                    \
        hb::set_last_error(_ex.what());                                                    \
        return -1;                                                                         \
    }                                                                                      \
    return 0;

// Count + check a kernel launch.
#define HB_LAUNCHED()                                                                      \
    do {                                                                                   \
        hb::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
        HB_CUDA(cudaGetLastError());                                                       \
    } while (0)

inline cudaStream_t stream_of(DLStreamHandle h) {
    return h ? *(cudaStream_t *)h->handle : (cudaStream_t)0;
}

inline size_t numel(const DLArray *a) {
    size_t n = 1;
    for (int i = 0; i < a->ndim; i++)
        n *= (size_t)a->shape[i];
    return n;
}

inline int ceil_div(size_t a, size_t b) {
    return (int)((a + b - 1) / b);
}

int sm_count();
u32 default_hot_threshold();
bool pdl_enabled(); // $HERALD_PDL != 0 (default on)

#ifdef __CUDACC__
// ---- launches: programmatic dependent launch (PDL) ----------------------------------------
// A step is 30+ short kernels on one stream; with plain stream order every boundary costs a full
// launch latency.  Every kernel of the library is launched with programmatic stream
// serialization and starts with pdl_enter(): griddepcontrol.wait (block until the previous kernel
// has COMPLETED and its writes are visible), then griddepcontrol.launch_dependents (the next
// kernel's CTAs may become resident now and park at their own wait).  No kernel touches memory
// before its wait, so ordering is exactly stream order; only the launch latency is overlapped,
// one kernel deep.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <class... P, class... A>
inline void launch_kernel(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                          A &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    HB_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(std::forward<A>(args))...));
}
#define HB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    hb::launch_kernel(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, ##__VA_ARGS__)

// ---- device helpers --------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned lane_id() {
    return threadIdx.x & 31;
}
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// 128-bit streaming load/store (read-once / write-once data: bypass L1 allocation).
__device__ __forceinline__ float4 ld_stream(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float4 *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
// coherent 128-bit load that does not allocate in L1 (rows that are read, modified and written once:
// an L1 line would only displace the per-row metadata other warps are about to read)
__device__ __forceinline__ float4 ld_row(const float4 *p) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
// Store of a row the next kernels read again — the cache rows the sync has just pulled are read by
// the gather right behind it and rewritten by the update one call later — with an evict-last L2
// policy (createpolicy + st.L2::cache_hint): 43 MB per lookup at the headline config, well inside the
// 126 MB L2, stay resident although ten times that streams through in between.  Measured per step:
// 0.272 -> 0.265 ms (segment_reduce 85 -> 80 us, gather 58 -> 56.5 us).  Tagging more was measured
// and is worse: evict-last LOADS in the gather slow the gather by 3 us, evict-last stores of the
// update's rows and evict-first hints on the one-shot streams (gradients, gathered output, owner
// rows) change nothing or cost 2 us.
__device__ __forceinline__ void st_keep(float4 *p, const float4 &v) {
    u64 policy;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void st_row(float4 *p, const float4 &v) {
    *p = v;
}
// Ask the memory system to bring a row (`bytes` at `p`) into L2: one prefetch.global.L2 per 128-byte
// line, no register or shared-memory destination.  The row movers know their next rows a few hundred
// nanoseconds before they read them (the next ticket's work items, the 32 source rows of a gather
// group), one row per LANE: prefetched, the demand loads pay L2 latency instead of HBM latency, so
// the same loads in flight per warp sustain more bandwidth.  (cp.async.bulk.prefetch.L2 would take
// the whole row in one instruction but issues from the uniform datapath — one address per WARP: the
// compiler serialises 32 different lane addresses, measured slower here.)
__device__ __forceinline__ void prefetch_l2(const void *p, unsigned bytes) {
    const char *q = static_cast<const char *>(p);
    for (unsigned o = 0; o < bytes; o += 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q + o) : "memory");
}
// one whole row with a single instruction, for a WARP-UNIFORM address (sm_90+ bulk prefetch)
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// exact (never contracted) fp32 add: keeps ((a+g1)+g2)... bit-identical to the CPU path
__device__ __forceinline__ float4 add4(const float4 &a, const float4 &b) {
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z),
                       __fadd_rn(a.w, b.w));
}

// (uint64)(float) as the reference's raw entry points cast ids (cache.cc:53-55);
// negative / NaN ids are undefined behaviour there, here they saturate to 0.
__device__ __forceinline__ u64 key_from_f32(float f) {
    return f > 0.f ? (u64)f : 0ull;
}

// ---- single-pass (decoupled look-back) exclusive scan of one u32 per thread --
// status word per tile: (flag << 32) | value, flag 0 = not ready, 1 = tile aggregate,
// 2 = inclusive prefix.  Tile ids come from an atomic ticket so that a tile never waits on
// a tile that has not started.  The caller zeroes `status[0..ntiles)` and `*ticket`.
struct ScanState {
    u64 *status;
    u32 *ticket;
};

// Block-wide: every thread passes its value; .excl is the exclusive prefix over the whole grid
// in ticket (tile) order, .tile_prefix the sum of all earlier tiles, .tile_total this tile's
// own sum.  In the last tile tile_prefix + tile_total is the grand total.  BLOCK = blockDim.x,
// a multiple of 32, <= 1024.  Safe to call more than once per kernel.
struct ScanResult {
    u32 excl, tile_prefix, tile_total;
};

template <int BLOCK>
__device__ __forceinline__ ScanResult grid_exclusive_scan(ScanState st, u32 value, u32 tile) {
    __shared__ u32 s_warp[BLOCK / 32];
    __shared__ u32 s_prefix, s_total;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    // inclusive scan inside the warp
    u32 incl = value;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(FULL, incl, d);
        if (lane >= (unsigned)d)
            incl += t;
    }
    if (lane == 31)
        s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u32 w = lane < BLOCK / 32 ? s_warp[lane] : 0;
        u32 wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 t = __shfl_up_sync(FULL, wi, d);
            if (lane >= (unsigned)d)
                wi += t;
        }
        if (lane < BLOCK / 32)
            s_warp[lane] = wi - w; // exclusive prefix of each warp
        u32 block_sum = __shfl_sync(FULL, wi, 31);
        // publish the aggregate, then look back
        u32 prefix = 0;
        if (tile == 0) {
            if (lane == 0)
                atomicExch(&st.status[0], (2ull << 32) | block_sum);
        } else {
            if (lane == 0)
                atomicExch(&st.status[tile], (1ull << 32) | block_sum);
            // kGroups x 32 predecessors are loaded before the first is examined: the tiles of a
            // small problem start together, so nobody finds a finished prefix early and a
            // round trip per 32 predecessors would be paid serially
            constexpr int kGroups = 8;
            int look = (int)tile - 1;
            bool found = false;
            while (!found) {
                u64 s[kGroups];
#pragma unroll
                for (int g = 0; g < kGroups; g++) {
                    const int idx = look - g * 32 - (int)lane;
                    s[g] = idx >= 0 ? *((volatile u64 *)&st.status[idx]) : (2ull << 32);
                }
#pragma unroll
                for (int g = 0; g < kGroups; g++) {
                    if (found)
                        continue;
                    const int idx = look - g * 32 - (int)lane;
                    if (idx >= 0)
                        while ((s[g] >> 32) == 0)
                            s[g] = *((volatile u64 *)&st.status[idx]);
                    unsigned is_prefix = __ballot_sync(FULL, (s[g] >> 32) == 2);
                    u32 v = (u32)s[g];
                    if (is_prefix) {
                        int first = __ffs(is_prefix) - 1; // nearest tile holding a full prefix
                        if ((int)lane > first)
                            v = 0;
                    }
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1)
                        v += __shfl_xor_sync(FULL, v, d);
                    prefix += v;
                    found = is_prefix != 0;
                }
                look -= kGroups * 32;
            }
            if (lane == 0)
                atomicExch(&st.status[tile], (2ull << 32) | (prefix + block_sum));
        }
        if (lane == 0) {
            s_prefix = prefix;
            s_total = block_sum;
        }
    }
    __syncthreads();
    ScanResult r;
    r.tile_prefix = s_prefix;
    r.tile_total = s_total;
    r.excl = s_prefix + s_warp[warp] + (incl - value);
    __syncthreads();
    return r;
}

__device__ __forceinline__ u32 take_ticket(u32 *ticket) {
    __shared__ u32 s_tile;
    if (threadIdx.x == 0)
        s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    return s_tile;
}
#endif // __CUDACC__

} // namespace hb
