// Worker-side embedding cache and owner-side table shard, both resident in HBM.
//
// What each call restates (reference file:line):
//   hb_cache_lookup                  CacheBase::_embeddingLookup            src/hetu_cache/src/cache.cc:60-107
//   hb_cache_update                  CacheBase::_embeddingUpdate            cache.cc:132-196
//   hb_cache_update_with_push_keys   CacheBase::_embeddingUpdateWithPushKeys cache.cc:248-334
//   hb_cache_push_pull               CacheBase::_embeddingPushPull          cache.cc:356-422
//   sync / push between cache and owner   hetu_client.cc:6-55 + ps-lite/src/PSFhandle_embedding.cc:5-79
//   replacement policies             lru_cache.cc, lfu_cache.cc, lfuopt_cache.cc
//
// The reference walks the sorted unique keys of a batch serially through a hash map and linked
// lists.  Here the same decisions are made in parallel:
//   * every policy touch / insert gets a stamp from one monotone clock, in the order the serial
//     walk would perform it (sorted rank inside a batch), so "position in the list" becomes a
//     64-bit priority word per slot: (use count << 52) | stamp;
//   * the victims of a batch of inserts are the k smallest priorities of the policy's victim
//     class (LRU: all lines; LFU: lines used once; LFUOpt: lines never re-used), found with a
//     12-bit-per-level radix select over the slot array, plus closed-form handling of the
//     corner cases in which freshly inserted lines evict each other;
//   * gradients are accumulated per unique row in occurrence order by one warp (bit-identical
//     to Line::accumulate, embedding.h:78-91), fused with the push to the owner row.
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>

#include "hb_cache.cuh"
#include "hb_segment.cuh"

namespace hb {
namespace {

// =====================================================================================
// small device helpers
// =====================================================================================
__device__ __forceinline__ u32 hash_key(u64 k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (u32)k;
}

// returns slot or -1
__device__ __forceinline__ i32 ht_find(const HtEntry *ht, u32 mask, u64 key) {
    u32 h = hash_key(key) & mask;
    for (u32 probes = 0; probes <= mask; probes++) {
        const ulonglong2 e = *reinterpret_cast<const ulonglong2 *>(&ht[h]);
        if (e.x == key)
            return (i32)(u32)e.y;
        if (e.x == HT_EMPTY)
            return -1;
        h = (h + 1) & mask;
    }
    return -1;
}

// key must be absent.  Returns false when the table has no free entry.
// `fresh` counts the EMPTY entries taken (the caller adds the warp's sum to regs->ht_occupied with
// one atomic: add_occupied()).
__device__ __forceinline__ bool ht_insert(HtEntry *ht, u32 mask, u64 key, u32 slot, u32 &fresh) {
    u32 h = hash_key(key) & mask;
    for (u32 probes = 0; probes <= mask;) {
        u64 cur = *reinterpret_cast<volatile u64 *>(&ht[h].key);
        if (cur == HT_EMPTY || cur == HT_TOMB) {
            u64 old = atomicCAS(&ht[h].key, cur, key);
            if (old == cur) {
                ht[h].slot = slot;
                if (cur == HT_EMPTY)
                    fresh++;
                return true;
            }
            continue; // lost the race for this entry: look at it again
        }
        h = (h + 1) & mask;
        probes++;
    }
    return false;
}

__device__ __forceinline__ void ht_erase(HtEntry *ht, u32 mask, u64 key) {
    u32 h = hash_key(key) & mask;
    for (u32 probes = 0; probes <= mask; probes++) {
        u64 cur = ht[h].key;
        if (cur == key) {
            ht[h].key = HT_TOMB;
            return;
        }
        if (cur == HT_EMPTY)
            return;
        h = (h + 1) & mask;
    }
}

// End of a kernel (all lanes converged): one atomic per warp for the index entries it took.
__device__ __forceinline__ void add_occupied(u32 *occupied, u32 fresh) {
    __syncwarp();
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        fresh += __shfl_xor_sync(0xffffffffu, fresh, d);
    if ((threadIdx.x & 31) == 0 && fresh)
        atomicAdd(occupied, fresh);
}

// One atomic per warp: every lane of the warp must call this (pred may differ per lane).
__device__ __forceinline__ u32 warp_append(u32 *counter, bool pred) {
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

__device__ __forceinline__ u32 *block_counter() {
    __shared__ u32 s_counter;
    return &s_counter;
}
__device__ __forceinline__ u32 *block_counter2() {
    __shared__ u32 s_counter2;
    return &s_counter2;
}

// Split barriers over peer-mapped flags (no launch of their own): a kernel that is about to read
// what peers wrote spins, one lane per peer, until every flag has reached `epoch`; the flags are
// raised by the kernel that finished the writes.  Bounded: a peer that never shows up fails the
// call (E_BARRIER) instead of hanging the GPU.
__device__ __forceinline__ void wait_flags(const CacheView &c, const u64 *flags, u64 epoch) {
    if (threadIdx.x < (unsigned)c.pv.world && epoch) {
        const u64 t0 = global_timer_ns();
        u64 seen = 0;
        unsigned spins = 0;
        while (true) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flags + threadIdx.x) : "memory");
            if (seen >= epoch)
                break;
            if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > c.pv.timeout_ns) {
                atomicMax(&c.regs->xerror, (u32)E_BARRIER); // (reported by the next epilogue: op_end_body)
                break;
            }
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void raise_flag(u64 *flag, u64 epoch) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
}

__device__ __forceinline__ int class_use_of(int policy) {
    return policy == HB_POLICY_LFU ? 1 : 0;
}

__device__ __forceinline__ u64 make_prio(u32 use, u64 stamp) {
    return ((u64)min(use, kUseSat) << kStampBits) | (stamp & kStampMask);
}

// =====================================================================================
// call prologue / epilogue
// =====================================================================================
// clk[0..3]: the replacement clock as a chain — stage s of a call reads clk[s] and writes
// clk[s+1], so no kernel reads a word that another block of the same kernel writes.
// flush != 0: the call pushes, so it also flushes the dirty victims collected so far (evict_)
// It also zeroes the scan states / work-list counters of the workspaces the call uses (a memset
// node between two kernels would cost their launch overlap).
struct ZeroList {
    KeyWorkspace::ZeroRange r[6];
    int n;
};
__global__ void __launch_bounds__(256) op_begin_kernel(CacheRegs *r, u64 *clk, int flush, ZeroList z) {
    pdl_enter();
#pragma unroll
    for (int k = 0; k < 6; k++) // (unrolled: a run-time index into the parameter array would go through local memory)
        if (k < z.n)
            for (u32 w = threadIdx.x; w < z.r[k].words; w += blockDim.x)
                z.r[k].p[w] = 0;
    if (threadIdx.x != 0)
        return;
    r->clock0 = r->clock;
    r->error = 0; // the previous call's failure has been reported in its own record
    clk[0] = r->clock;
    clk[1] = clk[2] = clk[3] = r->clock;
    r->U = r->M = r->alloc_base = 0;
    r->alloc_top0 = r->alloc_top1 = r->free_top;
    r->tail_done = 0;
    r->pulled = r->pushed = 0;
    r->pulled_remote = r->pushed_remote = 0;
    r->flushed = flush ? r->pending : 0;
    r->E = r->k_old = r->n_drop = r->need_min = r->nv = r->nc = 0;
    r->U2 = r->M2 = r->alloc_base2 = 0;
}

__device__ __forceinline__ void op_end_body(const CacheView &c, const u64 *clk, int last_stage, PerfRecord *rec,
                                            u32 kind, u32 num_all, int inserted_batch) {
    CacheRegs *r = c.regs;
    r->clock = clk[last_stage];
    if (inserted_batch) {
        u32 inserted = r->M - r->n_drop;
        r->size = r->size - r->nv + inserted;
    }
    rec->kind = kind;
    rec->num_all = num_all;
    if (kind == 0) {
        rec->num_unique = r->U;
        rec->num_miss = r->M;
        rec->num_evict = 0;
        rec->num_transfered = r->pulled;
    } else if (kind == 1) {
        rec->num_unique = r->U;
        rec->num_miss = r->M;
        rec->num_evict = r->flushed;
        rec->num_transfered = r->pushed + r->flushed;
    } else {
        rec->num_unique = r->U;
        rec->num_miss = r->M;
        rec->num_evict = r->flushed;
        rec->num_transfered = r->pulled;
    }
    rec->num_remote = kind == 0 ? r->pulled_remote : r->pushed_remote;
    rec->size = r->size;
    // failures of the exchange kernels (they run on their own stream, outside any call's begin / end
    // bracket) are reported by the first epilogue that sees them
    rec->error = max(r->error, atomicExch(&r->xerror, 0u));
    rec->ht_occupied = r->ht_occupied;
    rec->pending = r->pending;
    rec->limit_full = r->size == c.limit;
    rec->clock = r->clock;
    rec->floor = r->floor;
}

__global__ void op_end_kernel(CacheView c, const u64 *clk, int last_stage, PerfRecord *rec, u32 kind,
                              u32 num_all, int inserted_batch) {
    pdl_enter();
    op_end_body(c, clk, last_stage, rec, kind, num_all, inserted_batch);
}

// =====================================================================================
// batched insert, step one: the plan  (cache.cc:28-35 + policy insert())
// =====================================================================================
constexpr int kSelUnroll = 4; // slot priorities a thread loads before it uses the first
constexpr int kLogBlock = 512, kLogItems = 8;
constexpr int kLogTile = kLogBlock * kLogItems; // stamps one tile of the log walk covers

// The plan of a batched insert (block-wide; thread 0 decides, every thread helps zeroing): how many
// lines the serial loop would evict (E), where the new lines' stamps start, the bins / the log span
// of the victim selection.  `now` = the replacement clock after the batch's touches; M = regs->M.
// Runs as the tail of the resolve's last tile (a lookup: that thread has just settled M and the
// clock, and no other tile touches these words — one launch less in front of the victim selection)
// or as a kernel of its own (push_pull, single keys).
__device__ __forceinline__ void plan_insert_body(const CacheView &c, int bypass, u64 now, u64 *clk_out) {
    __shared__ u32 s_log_tiles;
    for (int b = threadIdx.x; b < 2 * kSelBins; b += blockDim.x)
        c.sel_hist[b] = 0;
    if (threadIdx.x == 0) {
        CacheRegs *r = c.regs;
        r->sel_done2 = 0;
        r->sel_above = 0;
        r->sel_fallback = 0;
        r->sel_cut_eff = r->sel_cut ? r->sel_cut : (u32)kSelBins;
        const u32 M = r->M, size_old = r->size, limit = c.limit;
        u32 E = 0;
        if (!bypass) {
            if (c.policy == HB_POLICY_LRU) { // insert, then evict while over limit (lru_cache.cc:9-25)
                u64 tot = (u64)size_old + M;
                E = tot > limit ? (u32)(tot - limit) : 0;
            } else { // evict before insert when full (lfu_cache.cc:9-20, lfuopt_cache.cc:9-26)
                u32 F = limit > size_old ? limit - size_old : 0;
                E = M > F ? M - F : 0;
            }
        }
        r->E = E;
        r->k_old = 0;
        r->n_drop = bypass ? M : 0;
        r->need_min = 0;
        r->nv = 0;
        r->nc = 0;
        r->sel_done = 0;
        r->min_use = 0xffffffffu;
        r->min_prio = ~0ull;
        r->ins_clock0 = now;
        *clk_out = now + M;
        // bins cover [floor, now): (stamp - floor) >> shift < kSelBins
        const u64 span = now > r->floor ? now - r->floor : 1; // stamp offsets 0 .. span-1
        const int bits = span > 1 ? 64 - __clzll((long long)(span - 1)) : 0;
        r->sel_shift = bits > kSelBits ? (u32)(bits - kSelBits) : 0;
        // LRU: every stamp of [floor, now) still has its own entry in the stamp log
        const bool use_log = c.policy == HB_POLICY_LRU && E > 0 && now > r->floor &&
                             now - r->floor <= (u64)c.log_mask + 1;
        r->sel_use_log = use_log ? 1u : 0u;
        r->sel_floor0 = r->floor;
        s_log_tiles = use_log ? (u32)((now - r->floor + kLogTile - 1) / kLogTile) : 0;
    }
    __syncthreads();
    // scan state of the log walk: ticket + one status word per tile
    for (u32 w = threadIdx.x; w < s_log_tiles + 2; w += blockDim.x)
        c.sel_scan[w] = 0;
}

__global__ void __launch_bounds__(256)
    plan_insert_kernel(CacheView c, int bypass, const u64 *clk_in, u64 *clk_out) {
    pdl_enter();
    plan_insert_body(c, bypass, *clk_in, clk_out);
}

// =====================================================================================
// resolve: policy lookup of every unique key (+ touch), ordered compaction of the misses
// =====================================================================================
// batch 0 writes U/M/alloc_base, batch 1 (push side of push_pull) writes U2/M2/alloc_base2.
constexpr int kResolveItems = 1; // measured: 4 per thread is slower (the probes of one thread serialise: 16 -> 19 us)

// Plan: what an update adds to the resolve of its batch — the segment-reduce work item of every
// unique (seg_plan_item: push decision, versions, mailbox header, hot lists), written by the thread
// that has just resolved that unique and still holds its slot: one kernel and one pass over the
// per-unique scalars less than a separate plan kernel.  NoSegPlan for lookups.
struct NoSegPlan {
    static constexpr bool enabled = false;
};
template <class F>
struct SegPlanArgs {
    static constexpr bool enabled = true;
    const u32 *seg_start, *perm;
    u32 thr;
    HotLists hl;
    F f;
};

template <class Plan>
__global__ void __launch_bounds__(kScanBlock)
    resolve_kernel(CacheView c, const u64 *uniq, const u32 *num_unique, i32 *uslot, u32 *miss_list,
                   int bypass, ScanState st, u32 ntiles, const u64 *clk_in, u64 *clk_out, int batch,
                   int dataless, int plan_insert, Plan plan) {
    pdl_enter();
    if constexpr (Plan::enabled)
        plan.f.kernel_begin();
    const u32 tile = take_ticket(st.ticket);
    const u32 U = *num_unique;
    const u64 base = *clk_in;
    // height of the free stack when this kernel starts (op_begin / the previous resolve of the call
    // left it there: the last tile of THIS kernel lowers regs->free_top while other tiles still run)
    const u32 top_in = batch == 0 ? c.regs->alloc_top0 : c.regs->alloc_top1;
    // kResolveItems consecutive uniques per thread
    const u32 i0 = (tile * kScanBlock + threadIdx.x) * kResolveItems;
    u64 key[kResolveItems];
    i32 sl[kResolveItems];
#pragma unroll
    for (int j = 0; j < kResolveItems; j++)
        key[j] = i0 + j < U ? uniq[i0 + j] : 0;
#pragma unroll
    for (int j = 0; j < kResolveItems; j++)
        sl[j] = (i0 + j < U && !bypass) ? ht_find(c.ht, c.ht_mask, key[j]) : -1;
    u32 miss = 0;
#pragma unroll
    for (int j = 0; j < kResolveItems; j++) {
        const u32 i = i0 + j;
        if (i >= U)
            continue;
        const i32 s = sl[j];
        if (key[j] >= c.table_len)
            atomicMax(&c.regs->error, (u32)E_KEY_RANGE);
        if (s >= 0) {
            const u64 stamp = base + i;
            c.stamp_log[stamp & c.log_mask] = (u32)s;
            switch (c.policy) {
            case HB_POLICY_LRU: // lru_cache.cc:27-39: move to the front
                c.slot_prio[s] = make_prio(0, stamp);
                break;
            case HB_POLICY_LFU: { // lfu_cache.cc:22-29, 52-69: front of the (use+1) list
                u32 use = c.slot_use[s] + 1;
                c.slot_use[s] = use;
                c.slot_prio[s] = make_prio(use, stamp);
                break;
            }
            default: // lfuopt_cache.cc:28-44
                if (c.slot_state[s] == S_CACHED) {
                    u32 use = c.slot_use[s];
                    if (use + 1 < kLfuOptUseMax) {
                        c.slot_use[s] = use + 1;
                        c.slot_prio[s] = make_prio(use + 1, stamp);
                    } else { // promoted to the permanent store
                        c.slot_state[s] = S_STORE;
                        c.slot_prio[s] = PRIO_NONE;
                        atomicAdd(&c.regs->store_size, 1u);
                    }
                }
                break;
            }
        } else {
            miss++;
        }
        uslot[i] = s;
    }
    ScanResult sr = grid_exclusive_scan<kScanBlock>(st, miss, tile);
    u32 pos = sr.excl;
    // Every miss gets a fresh line right here (cache.cc:70-76 with data, :146-151 dataless): miss
    // number j of the call takes the j-th slot from the top of the free stack — its rank is all a
    // thread needs, so no second kernel has to wait for the total.
#pragma unroll
    for (int j = 0; j < kResolveItems; j++)
        if (i0 + j < U && sl[j] < 0) {
            const u32 jpos = pos++;
            miss_list[jpos] = i0 + j;
            if (jpos < top_in) {
                const u32 s = c.free_stack[top_in - 1 - jpos];
                uslot[i0 + j] = (i32)s;
                c.slot_key[s] = key[j];
                c.slot_version[s] = -1; // embedding.h:35,42
                c.slot_updates[s] = 0;
                c.slot_prio[s] = PRIO_NONE;
                c.slot_use[s] = 0;
                c.slot_state[s] = S_TRANSIENT;
                c.slot_flags[s] = dataless ? F_DATALESS : 0;
            }
        }
    if (tile == ntiles - 1 && threadIdx.x == 0) {
        CacheRegs *r = c.regs;
        u32 M = sr.tile_prefix + sr.tile_total;
        if (M > top_in) {
            atomicMax(&r->error, (u32)E_NO_FREE_SLOT);
            M = top_in; // the misses beyond have no line (uslot -1); the call is reported as failed
        }
        const u32 new_top = top_in - M;
        r->free_top = new_top;
        r->alloc_top1 = new_top;
        r->slot_hw = max(r->slot_hw, c.capacity - new_top); // the stack hands out 0, 1, 2, ...
        if (batch == 0) {
            r->U = U;
            r->M = M;
        } else {
            r->U2 = U;
            r->M2 = M;
        }
        *clk_out = base + U;
    }
    // a lookup: the insert plan right here (this tile knows M and the clock; block-uniform branch)
    if (plan_insert && tile == ntiles - 1)
        plan_insert_body(c, bypass, base + U, clk_out + 1);
    if constexpr (Plan::enabled) {
        static_assert(kResolveItems == 1, "one unique per thread");
        // ticket order: item i = t * TK + l holds unique u = t + l * T (hb_rows.cuh); this thread owns
        // u = i0, or one of the padding items behind the last unique
        const u32 TK = plan.hl.ticket_rows;
        const u32 T = (U + TK - 1) / TK, total = T * TK;
        const u32 u = i0;
        const u32 i = (T && u < total) ? (u % T) * TK + u / T : total;
        seg_plan_item(i, u, U, total, plan.seg_start, plan.perm, plan.thr, plan.hl, plan.f);
        plan.f.kernel_end();
    }
}

// =====================================================================================
// sync with the owner shard: kSyncEmbedding (PSFhandle_embedding.cc:30-64) + client closure
// (hetu_client.cc:19-32): rows whose version is -1 or more than pull_bound behind are re-read.
// =====================================================================================
// A warp takes 32 consecutive uniques: the staleness test of all 32 is evaluated lane-parallel
// (slot, version, owner version), then the stale rows are copied ROWS at a time so that several
// 512 B row reads are in flight per warp.
template <int VEC, int ROWS>
__global__ void __launch_bounds__(kRowBlock)
    sync_kernel(CacheView c, const u64 *__restrict__ uniq, const i32 *__restrict__ uslot,
                i64 pull_bound, u64 applied_epoch) {
    pdl_enter();
    // multi-GPU, BSP: every owner has applied the pushes of the exchanges enqueued before this
    // lookup (the role of BarrierWorker, ParameterServerCommunicate.py:48-52)
    if (c.pv.world > 1)
        wait_flags(c, c.pv.ctrl + kMaxWorld, applied_epoch);
    using V = RowVec<VEC>;
    const unsigned lane = lane_id();
    const size_t warp_global = (size_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * kRowWarps;
    const size_t D = c.width, nvec = D / VEC;
    const u32 U = c.regs->U;
    u32 *cnt = block_counter(), *cnt_remote = block_counter2();
    if (threadIdx.x == 0)
        *cnt = *cnt_remote = 0;
    __syncthreads();
    u32 pulled = 0, pulled_remote = 0;
    for (size_t base = warp_global * 32; base < U; base += nwarps * 32) {
        const size_t i = base + lane;
        i32 s = -1;
        u64 trow = 0;
        i64 srv = 0;
        bool need = false, addup = false, live_grad = false;
        int owner = 0;
        if (i < U) {
            s = uslot[i];
            bool have;
            if (c.pv.world > 1) { // the owner's shard is read in place over NVLink
                owner = owner_of(c.pv, uniq[i], trow);
                have = uniq[i] < c.table_len;
            } else {
                trow = uniq[i] - c.row_begin;
                have = trow < c.nrows_local;
            }
            if (s >= 0 && have) {
                const i64 v = c.slot_version[s];
                srv = __ldg(&c.pv.ver[owner][trow]);
                need = v == -1 || srv - v > pull_bound;
                if (need) {
                    addup = c.slot_flags[s] & F_GRAD; // Line::addup (embedding.h:92-96)
                    live_grad = addup && c.slot_updates[s] != 0; // grad store is meaningful
                    c.slot_version[s] = srv;
                }
            }
        }
        unsigned m = __ballot_sync(FULL, need);
        pulled += __popc(m);
        pulled_remote += __popc(__ballot_sync(FULL, need && owner != c.pv.rank));
        if (VEC == 4 && need && (c.pv.world == 1 || owner == c.pv.rank)) // local rows: towards L2 now
            prefetch_l2(c.pv.rows[owner] + trow * D, (unsigned)(D * sizeof(float)));
        while (m) {
            int src[ROWS];
            i32 rs[ROWS];
            u64 rt[ROWS];
            const float *rbase[ROWS];
            bool ra[ROWS], rg[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
                src[r] = m ? __ffs(m) - 1 : -1;
                if (m)
                    m &= m - 1;
                const int from = src[r] < 0 ? 0 : src[r];
                rs[r] = __shfl_sync(FULL, s, from);
                rt[r] = __shfl_sync(FULL, trow, from);
                rbase[r] = c.pv.rows[__shfl_sync(FULL, owner, from)];
                ra[r] = __shfl_sync(FULL, addup, from);
                rg[r] = __shfl_sync(FULL, live_grad, from);
            }
            for (size_t k = lane; k < nvec; k += 32) {
                typename V::T x[ROWS], g[ROWS];
#pragma unroll
                for (int r = 0; r < ROWS; r++)
                    if (src[r] >= 0) {
                        // read-only for the whole kernel (pushes reach a shard only between the
                        // barriers of an update), and possibly a PEER's memory: the non-coherent
                        // path fetches whole lines over NVLink; a plain ld.global of peer memory
                        // ran 10x slower (660 us for 52k remote rows against 60 us)
                        x[r] = V::ld_nc(rbase[r] + rt[r] * D + k * VEC);
                        g[r] = rg[r] ? V::ld(c.grad + (size_t)rs[r] * D + k * VEC) : V::zero();
                    }
#pragma unroll
                for (int r = 0; r < ROWS; r++)
                    if (src[r] >= 0) {
                        if (ra[r])
                            x[r] = V::add(x[r], g[r]);
                        V::st_keep(c.data + (size_t)rs[r] * D + k * VEC, x[r]);
                    }
            }
        }
    }
    if (lane == 0 && pulled)
        atomicAdd(cnt, pulled);
    if (lane == 0 && pulled_remote)
        atomicAdd(cnt_remote, pulled_remote);
    __syncthreads();
    if (threadIdx.x == 0 && *cnt)
        atomicAdd(&c.regs->pulled, *cnt);
    if (threadIdx.x == 0 && *cnt_remote)
        atomicAdd(&c.regs->pulled_remote, *cnt_remote);
}

// The sync with the stale rows fetched by BULK asynchronous copies (cp.async.bulk, the 1-D TMA path):
// one elected lane issues one copy per row — 512 bytes (D = 128) or 2 KB (D = 512) per instruction —
// into a per-warp shared-memory stage and the copies complete on an mbarrier (complete_tx); the warp
// then moves the stage into the cache rows.  A bulk copy holds no register while it is in flight and is
// one fabric-level request stream per row instead of 32 lanes' 16-byte loads, which is what the pull
// over NVLink wants.  Rows of 16-byte multiples only.
template <int DUMMY>
__global__ void __launch_bounds__(kRowBlock)
    sync_bulk_kernel(CacheView c, const u64 *__restrict__ uniq, const i32 *__restrict__ uslot,
                     i64 pull_bound, u64 applied_epoch, int kBulkRows /* rows per round and warp */) {
    pdl_enter();
    if (c.pv.world > 1)
        wait_flags(c, c.pv.ctrl + kMaxWorld, applied_epoch);
    extern __shared__ __align__(128) float s_stage[]; // [kRowWarps][kBulkRows][D], then kRowWarps mbarriers
    using V = RowVec<4>;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const size_t warp_global = (size_t)blockIdx.x * kRowWarps + warp;
    const size_t nwarps = (size_t)gridDim.x * kRowWarps;
    const size_t D = c.width, nvec = D / 4;
    const unsigned row_bytes = (unsigned)(D * sizeof(float));
    const u32 U = c.regs->U;
    float *stage = s_stage + (size_t)warp * kBulkRows * D;
    u64 *bar = reinterpret_cast<u64 *>(s_stage + (size_t)kRowWarps * kBulkRows * D) + warp;
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(bar);
    if (lane == 0)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
    u32 *cnt = block_counter(), *cnt_remote = block_counter2();
    if (threadIdx.x == 0)
        *cnt = *cnt_remote = 0;
    __syncthreads(); // (also publishes the mbarrier inits)
    unsigned phase = 0;
    u32 pulled = 0, pulled_remote = 0;
    for (size_t base = warp_global * 32; base < U; base += nwarps * 32) {
        const size_t i = base + lane;
        i32 s = -1;
        u64 trow = 0;
        bool need = false, addup = false, live_grad = false;
        int owner = 0;
        if (i < U) {
            s = uslot[i];
            bool have;
            if (c.pv.world > 1) {
                owner = owner_of(c.pv, uniq[i], trow);
                have = uniq[i] < c.table_len;
            } else {
                trow = uniq[i] - c.row_begin;
                have = trow < c.nrows_local;
            }
            if (s >= 0 && have) {
                const i64 v = c.slot_version[s];
                const i64 srv = __ldg(&c.pv.ver[owner][trow]);
                need = v == -1 || srv - v > pull_bound;
                if (need) {
                    addup = c.slot_flags[s] & F_GRAD; // Line::addup (embedding.h:92-96)
                    live_grad = addup && c.slot_updates[s] != 0;
                    c.slot_version[s] = srv;
                }
            }
        }
        unsigned m = __ballot_sync(FULL, need);
        pulled += __popc(m);
        pulled_remote += __popc(__ballot_sync(FULL, need && owner != c.pv.rank));
        while (m) {
            const int nr = min(__popc(m), kBulkRows);
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a),
                             "r"(row_bytes * (unsigned)nr)
                             : "memory");
            unsigned round = 0;
            for (int r = 0; r < nr; r++) {
                const int from = __ffs(m) - 1;
                m &= m - 1;
                round |= 1u << from;
                const u64 rt = __shfl_sync(FULL, trow, from);
                const float *src = c.pv.rows[__shfl_sync(FULL, owner, from)] + rt * D;
                if (lane == 0) {
                    const unsigned d = (unsigned)__cvta_generic_to_shared(stage + (size_t)r * D);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                                 "l"(src), "r"(row_bytes), "r"(bar_a)
                                 : "memory");
                }
            }
            {   // every lane waits for the round's bytes
                unsigned done = 0;
                while (!done)
                    asm volatile("{\n\t.reg .pred p;\n\t"
                                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                                 "selp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done)
                                 : "r"(bar_a), "r"(phase)
                                 : "memory");
                phase ^= 1u;
            }
            int r = 0;
            while (round) {
                const int from = __ffs(round) - 1;
                round &= round - 1;
                const i32 rs = __shfl_sync(FULL, s, from);
                const bool ra = __shfl_sync(FULL, addup, from), rg = __shfl_sync(FULL, live_grad, from);
                const float *row = stage + (size_t)r * D;
                for (size_t k = lane; k < nvec; k += 32) {
                    float4 x = *reinterpret_cast<const float4 *>(row + k * 4);
                    if (ra) {
                        const float4 g = rg ? V::ld(c.grad + (size_t)rs * D + k * 4) : V::zero();
                        x = V::add(x, g);
                    }
                    V::st_keep(c.data + (size_t)rs * D + k * 4, x);
                }
                r++;
            }
            __syncwarp(); // the stage is free again (generic-proxy reads done before the next async writes)
        }
    }
    if (lane == 0 && pulled)
        atomicAdd(cnt, pulled);
    if (lane == 0 && pulled_remote)
        atomicAdd(cnt_remote, pulled_remote);
    __syncthreads();
    if (threadIdx.x == 0 && *cnt)
        atomicAdd(&c.regs->pulled, *cnt);
    if (threadIdx.x == 0 && *cnt_remote)
        atomicAdd(&c.regs->pulled_remote, *cnt_remote);
}

struct IndexFromSlots {
    const i32 *uslot;
    const u32 *inverse;
    __device__ long long operator()(size_t n) const {
        return (long long)uslot[inverse[n]];
    }
};

// =====================================================================================
// batched insert: select victims, evict, insert  (cache.cc:28-35 + policy insert())
// =====================================================================================
// LRU victim selection by walking the stamp log.  Every policy touch and insert takes the next
// tick of the replacement clock and records `stamp_log[t & mask] = slot`, so the resident lines in
// recency order are the entries of [floor, now) whose slot still carries that stamp.  The E oldest
// lines are found by walking the log from `floor`: tiles of kLogTile stamps, taken in order from a
// ticket, count their live entries and rank them with the grid scan; the tile in which the count
// reaches E writes the plan and stops the walk.  Work is proportional to the stamps issued since
// the victims were last touched (~4 stamps per victim in the WDL steady state) instead of two
// sweeps over the priorities of every slot of the cache.
__global__ void __launch_bounds__(kLogBlock) sel_log_kernel(CacheView c) {
    pdl_enter();
    CacheRegs *r = c.regs;
    if (!r->sel_use_log)
        return;
    __shared__ u32 s_tile;
    __shared__ u32 s_stop;
    const u32 E = r->E;
    const u64 floor = r->sel_floor0, now = r->ins_clock0;
    const u32 ntiles = (u32)((now - floor + kLogTile - 1) / kLogTile);
    ScanState st;
    st.ticket = reinterpret_cast<u32 *>(c.sel_scan);
    st.status = c.sel_scan + 1;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) {
            s_stop = *reinterpret_cast<volatile u32 *>(&r->sel_done);
            s_tile = s_stop ? 0u : atomicAdd(st.ticket, 1u);
        }
        __syncthreads();
        if (s_stop)
            break;
        const u32 tile = s_tile;
        if (tile >= ntiles)
            break;
        const u64 t0 = floor + (u64)tile * kLogTile + (u64)threadIdx.x * kLogItems;
        u32 slot[kLogItems];
        bool live[kLogItems];
#pragma unroll
        for (int j = 0; j < kLogItems; j++)
            slot[j] = t0 + j < now ? c.stamp_log[(t0 + j) & c.log_mask] : 0xffffffffu;
        u64 prio[kLogItems];
#pragma unroll
        for (int j = 0; j < kLogItems; j++)
            prio[j] = slot[j] < c.capacity ? c.slot_prio[slot[j]] : PRIO_NONE;
        u32 cnt = 0;
#pragma unroll
        for (int j = 0; j < kLogItems; j++) {
            live[j] = prio[j] == make_prio(0, t0 + j);
            cnt += live[j];
        }
        const ScanResult sr = grid_exclusive_scan<kLogBlock>(st, cnt, tile);
        u32 rank = sr.excl;
#pragma unroll
        for (int j = 0; j < kLogItems; j++) {
            if (live[j]) {
                if (rank < E)
                    c.victims[rank] = slot[j];
                if (rank == E - 1)
                    r->floor = t0 + j + 1; // every line below has just been selected
                rank++;
            }
        }
        const u32 incl = sr.tile_prefix + sr.tile_total;
        const bool completes = (sr.tile_prefix < E && incl >= E) || (tile == ntiles - 1 && incl < E);
        if (completes && threadIdx.x == 0) {
            const u32 k_old = min(E, incl);
            r->k_old = k_old;
            r->nv = k_old;
            r->n_drop = E - k_old; // the oldest new lines fall off the tail themselves
            r->need_min = 0;
            if (incl < E)
                r->floor = now;
            __threadfence();
            atomicExch(&r->sel_done, 1u);
        }
    }
}

// Block-wide: exclusive prefix of sh[0..kSelBins) in place; returns total.  blockDim.x is a
// multiple of 256 (the first 256 threads do the work, every thread must call it).
__device__ __forceinline__ u32 block_scan_bins(u32 *sh) {
    __shared__ u32 s_part[256];
    constexpr int PER = kSelBins / 256;
    const bool act = threadIdx.x < 256;
    const unsigned t256 = threadIdx.x & 255;
    u32 local[PER];
    u32 sum = 0;
    if (act) {
#pragma unroll
        for (int k = 0; k < PER; k++) {
            local[k] = sh[t256 * PER + k];
            sum += local[k];
        }
        s_part[t256] = sum;
    }
    __syncthreads();
    // Hillis-Steele over 256 partials
    for (int d = 1; d < 256; d <<= 1) {
        u32 t = (act && t256 >= (unsigned)d) ? s_part[t256 - d] : 0;
        __syncthreads();
        if (act)
            s_part[t256] += t;
        __syncthreads();
    }
    if (act) {
        u32 run = s_part[t256] - sum;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            sh[t256 * PER + k] = run;
            run += local[k];
        }
    }
    u32 total = s_part[255];
    __syncthreads();
    return total;
}

// Level-one histogram of the victim class over [floor, now).  The block that adds its bins last
// scans the finished histogram ONCE and leaves the selection plan in the registers: threshold bin,
// victims to take inside it, and the closed-form corner cases of DESIGN.md 4.3.
// level 0 counts only the bins below a window (`sel_cut_eff`, twice the previous call's threshold
// bin): the victims of a steady-state batch are the oldest few per cent of the class, so 97 % of
// the lines cost a compare instead of a shared-memory atomic on a handful of contended bins; the
// lines above the window are only counted.  If the threshold turns out to lie above the window
// (first call, a burst of evictions), level 1 — launched right after, a no-op otherwise — repeats
// the sweep over all bins.
__global__ void __launch_bounds__(1024) sel_hist_kernel(CacheView c, int level) {
    pdl_enter();
    CacheRegs *r = c.regs;
    const u32 E = r->E;
    if (E == 0 || r->sel_use_log)
        return;
    if (level == 1 && !r->sel_fallback)
        return;
    __shared__ u32 sh[kSelBins];
    __shared__ u32 s_bin, s_above;
    __shared__ bool s_last;
    u32 *const ghist = c.sel_hist + (level ? kSelBins : 0);
    u32 *const done = level ? &r->sel_done2 : &r->sel_done;
    const u32 cut = level ? (u32)kSelBins : r->sel_cut_eff;
    for (int b = threadIdx.x; b < kSelBins; b += blockDim.x)
        sh[b] = 0;
    if (threadIdx.x == 0)
        s_above = 0;
    __syncthreads();
    const u64 floor = r->floor;
    const u32 shift = r->sel_shift;
    const u64 cls = (u64)class_use_of(c.policy);
    const size_t hw = r->slot_hw;
    // kSelUnroll independent loads per thread before the first use: the sweep is a dependent-load
    // loop otherwise (one L2/HBM latency per 8 bytes per thread)
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    u32 above = 0;
    for (size_t s0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s0 < hw;
         s0 += stride * kSelUnroll) {
        u64 pr[kSelUnroll];
#pragma unroll
        for (int k = 0; k < kSelUnroll; k++) {
            const size_t s = s0 + (size_t)k * stride;
            pr[k] = s < hw ? c.slot_prio[s] : PRIO_NONE;
        }
#pragma unroll
        for (int k = 0; k < kSelUnroll; k++) {
            const u64 p = pr[k];
            if (p != PRIO_NONE && (p >> kStampBits) == cls) {
                const u64 bin = min(((p & kStampMask) - floor) >> shift, (u64)(kSelBins - 1));
                if (bin < cut)
                    atomicAdd(&sh[bin], 1u);
                else
                    above++;
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        above += __shfl_xor_sync(FULL, above, d);
    if (lane_id() == 0 && above)
        atomicAdd(&s_above, above);
    __syncthreads();
    for (int b = threadIdx.x; b < (int)cut; b += blockDim.x)
        if (sh[b])
            atomicAdd(&ghist[b], sh[b]);
    if (threadIdx.x == 0 && s_above)
        atomicAdd(&r->sel_above, s_above);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
        s_last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last)
        return;
    __threadfence();
    for (int b = threadIdx.x; b < kSelBins; b += blockDim.x)
        sh[b] = __ldcg(&ghist[b]);
    if (threadIdx.x == 0)
        s_bin = 0;
    __syncthreads();
    const u32 in_window = block_scan_bins(sh); // sh = exclusive prefix
    const u32 class_count = in_window + (level ? 0u : __ldcg(&r->sel_above));
    const u32 k_old = min(E, class_count);
    if (k_old > in_window) { // level 0 only: the threshold bin is above the window
        if (threadIdx.x == 0)
            r->sel_fallback = 1;
        return;
    }
    // threshold bin: the last bin whose exclusive prefix is < k_old (k_old > 0)
    if (k_old > 0) {
        for (int b = threadIdx.x; b < kSelBins; b += blockDim.x) {
            u32 ex = sh[b];
            u32 nxt = b + 1 < kSelBins ? sh[b + 1] : in_window;
            if (ex < k_old && nxt >= k_old)
                s_bin = b; // unique b: prefix is monotone and nxt > ex here
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const u32 bstar = s_bin;
        const u32 M = r->M, size_old = r->size;
        const u32 extra = E - k_old;
        u32 n_drop = 0, need_min = 0;
        if (extra > 0) {
            if (c.policy == HB_POLICY_LRU) {
                n_drop = extra; // the oldest new lines fall off the tail themselves
            } else {
                const u32 F = c.limit > size_old ? c.limit - size_old : 0;
                if (F + class_count == 0) {
                    const u32 evictable = size_old - r->store_size;
                    if (evictable > 0) { // one resident line goes, then new lines evict each other
                        need_min = 1;
                        n_drop = extra - 1;
                    } else { // LFUOpt: only the permanent store is populated -> nothing is cached
                        n_drop = M;
                    }
                } else {
                    n_drop = extra;
                }
            }
        }
        r->k_old = k_old;
        r->n_drop = n_drop;
        r->need_min = need_min;
        r->sel_bin = bstar;
        r->sel_rem = k_old > 0 ? k_old - sh[bstar] : 0;
        r->sel_cut = k_old > 0 ? min((u32)kSelBins, 2 * bstar + 64) : 0;
    }
}

// Second sweep over the slot priorities: everything below the threshold bin is a victim, the
// threshold bin's lines become candidates for sel_refine_kernel.  Both lists are gathered in
// shared memory and appended with ONE global atomic per block and flush (the victims of a
// steady-state batch are a few per cent of the slots, so a per-warp append would put tens of
// thousands of atomics on one address).
constexpr int kCollectCap = 2048;

__global__ void __launch_bounds__(256) sel_collect_kernel(CacheView c) {
    pdl_enter();
    CacheRegs *r = c.regs;
    if (r->E == 0 || r->k_old == 0 || r->sel_use_log)
        return;
    __shared__ u32 s_vic[kCollectCap];
    __shared__ u32 s_cslot[kCollectCap];
    __shared__ u64 s_cprio[kCollectCap];
    __shared__ u32 s_nv, s_nc, s_vbase, s_cbase;
    if (threadIdx.x == 0) {
        s_nv = 0;
        s_nc = 0;
    }
    __syncthreads();
    const u32 bstar = r->sel_bin;
    const u64 floor = r->floor;
    const u32 shift = r->sel_shift;
    const u64 cls = (u64)class_use_of(c.policy);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t hw = r->slot_hw;
    const size_t rounds = (hw + stride * kSelUnroll - 1) / (stride * kSelUnroll);
    constexpr u32 kPerRound = 256 * kSelUnroll; // most entries one round can add to a list
    static_assert(kPerRound * 2 <= kCollectCap, "a round must fit behind a flushed list");
    for (size_t it = 0; it <= rounds; it++) {
        if (it < rounds) {
            const size_t s0 = it * stride * kSelUnroll + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
            u64 pr[kSelUnroll];
#pragma unroll
            for (int k = 0; k < kSelUnroll; k++) {
                const size_t s = s0 + (size_t)k * stride;
                pr[k] = s < hw ? c.slot_prio[s] : PRIO_NONE;
            }
#pragma unroll
            for (int k = 0; k < kSelUnroll; k++) {
                const u64 p = pr[k];
                const size_t s = s0 + (size_t)k * stride;
                if (p != PRIO_NONE && (p >> kStampBits) == cls) {
                    const u64 bin = min(((p & kStampMask) - floor) >> shift, (u64)(kSelBins - 1));
                    if (bin < bstar) {
                        s_vic[atomicAdd(&s_nv, 1u)] = (u32)s;
                    } else if (bin == bstar) {
                        const u32 j = atomicAdd(&s_nc, 1u);
                        s_cprio[j] = p & kStampMask;
                        s_cslot[j] = (u32)s;
                    }
                }
            }
        }
        __syncthreads();
        // block-uniform: flush when the next round could overflow a list, and after the last round
        const u32 nv = s_nv, nc = s_nc;
        const bool last = it == rounds;
        if (last || nv + kPerRound > kCollectCap || nc + kPerRound > kCollectCap) {
            if (threadIdx.x == 0) {
                s_vbase = nv ? atomicAdd(&r->nv, nv) : 0;
                s_cbase = nc ? atomicAdd(&r->nc, nc) : 0;
            }
            __syncthreads();
            for (u32 k = threadIdx.x; k < nv; k += blockDim.x)
                c.victims[s_vbase + k] = s_vic[k];
            for (u32 k = threadIdx.x; k < nc; k += blockDim.x) {
                c.cand_prio[s_cbase + k] = s_cprio[k];
                c.cand_slot[s_cbase + k] = s_cslot[k];
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                s_nv = 0;
                s_nc = 0;
            }
        }
        __syncthreads();
    }
}

// One block finishes the selection inside the threshold bin: 12 more bits per round.
__global__ void __launch_bounds__(1024) sel_refine_kernel(CacheView c) {
    pdl_enter();
    CacheRegs *r = c.regs;
    if (r->E == 0 || r->k_old == 0 || r->sel_use_log)
        return;
    __shared__ u32 sh[kSelBins];
    __shared__ u32 s_part[1024];
    __shared__ u32 s_bin, s_next_n, s_vbase;
    u32 nc = r->nc, rem = r->sel_rem, shift = r->sel_shift;
    u64 lo = r->floor + ((u64)r->sel_bin << shift);
    int cur = 0;
    const size_t cap = c.capacity;
    while (rem > 0) {
        const u64 *cp = c.cand_prio + (size_t)cur * cap;
        const u32 *cs = c.cand_slot + (size_t)cur * cap;
        if (nc <= rem || shift == 0) { // everything left is a victim
            if (threadIdx.x == 0)
                s_vbase = atomicAdd(&r->nv, nc);
            __syncthreads();
            for (u32 k = threadIdx.x; k < nc; k += blockDim.x)
                c.victims[s_vbase + k] = cs[k];
            break;
        }
        const u32 nshift = shift > kSelBits ? shift - kSelBits : 0;
        for (int b = threadIdx.x; b < kSelBins; b += blockDim.x)
            sh[b] = 0;
        if (threadIdx.x == 0) {
            s_next_n = 0;
            s_bin = 0;
        }
        __syncthreads();
        for (u32 k = threadIdx.x; k < nc; k += blockDim.x)
            atomicAdd(&sh[min((cp[k] - lo) >> nshift, (u64)(kSelBins - 1))], 1u);
        __syncthreads();
        // exclusive prefix of the 4096 bins with 1024 threads (4 bins each)
        u32 local[4], sum = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            local[k] = sh[threadIdx.x * 4 + k];
            sum += local[k];
        }
        s_part[threadIdx.x] = sum;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            u32 t = threadIdx.x >= (unsigned)d ? s_part[threadIdx.x - d] : 0;
            __syncthreads();
            s_part[threadIdx.x] += t;
            __syncthreads();
        }
        u32 run = s_part[threadIdx.x] - sum;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            u32 ex = run, nxt = run + local[k];
            if (ex < rem && nxt >= rem)
                s_bin = threadIdx.x * 4 + k;
            sh[threadIdx.x * 4 + k] = ex;
            run = nxt;
        }
        __syncthreads();
        const u32 b2 = s_bin;
        const u32 below = sh[b2];
        if (threadIdx.x == 0)
            s_vbase = atomicAdd(&r->nv, below);
        __syncthreads();
        // victims below the bin, survivors of the bin go to the other candidate buffer
        u64 *np = c.cand_prio + (size_t)(cur ^ 1) * cap;
        u32 *ns = c.cand_slot + (size_t)(cur ^ 1) * cap;
        __shared__ u32 s_vcount;
        if (threadIdx.x == 0)
            s_vcount = 0;
        __syncthreads();
        for (u32 k0 = 0; k0 < nc; k0 += blockDim.x) {
            const u32 k = k0 + threadIdx.x;
            bool victim = false, keep = false;
            u64 p = 0;
            u32 s = 0;
            if (k < nc) {
                p = cp[k];
                s = cs[k];
                u32 bin = (u32)min((p - lo) >> nshift, (u64)(kSelBins - 1));
                victim = bin < b2;
                keep = bin == b2;
            }
            u32 vpos = warp_append(&s_vcount, victim);
            if (victim)
                c.victims[s_vbase + vpos] = s;
            u32 kpos = warp_append(&s_next_n, keep);
            if (keep) {
                np[kpos] = p;
                ns[kpos] = s;
            }
        }
        __syncthreads();
        rem -= below;
        nc = s_next_n;
        lo += (u64)b2 << nshift;
        shift = nshift;
        cur ^= 1;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        r->floor = lo; // every line of the class below `lo` has just been evicted
}

// LFU corner: the single resident line with the smallest (use, stamp).
__global__ void min_use_kernel(CacheView c) {
    pdl_enter();
    CacheRegs *r = c.regs;
    if (!r->need_min)
        return;
    u32 best = 0xffffffffu;
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < r->slot_hw;
         s += (size_t)gridDim.x * blockDim.x)
        if (c.slot_state[s] == S_CACHED)
            best = min(best, c.slot_use[s]);
    for (int d = 16; d > 0; d >>= 1)
        best = min(best, __shfl_xor_sync(FULL, best, d));
    if (lane_id() == 0 && best != 0xffffffffu)
        atomicMin(&r->min_use, best);
}
__global__ void min_prio_kernel(CacheView c) {
    pdl_enter();
    CacheRegs *r = c.regs;
    if (!r->need_min)
        return;
    const u32 mu = r->min_use;
    u64 best = ~0ull;
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < r->slot_hw;
         s += (size_t)gridDim.x * blockDim.x)
        if (c.slot_state[s] == S_CACHED && c.slot_use[s] == mu)
            best = min(best, c.slot_prio[s] & kStampMask);
    for (int d = 16; d > 0; d >>= 1)
        best = min(best, __shfl_xor_sync(FULL, best, d));
    if (lane_id() == 0 && best != ~0ull)
        atomicMin(&r->min_prio, best);
}
__global__ void min_pick_kernel(CacheView c) {
    pdl_enter();
    CacheRegs *r = c.regs;
    if (!r->need_min)
        return;
    const u32 mu = r->min_use;
    const u64 mp = r->min_prio;
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < r->slot_hw;
         s += (size_t)gridDim.x * blockDim.x)
        if (c.slot_state[s] == S_CACHED && c.slot_use[s] == mu &&
            (c.slot_prio[s] & kStampMask) == mp)
            c.victims[atomicAdd(&r->nv, 1u)] = (u32)s;
}

// Remove the victims from the index; dirty ones wait for the next push (evict_), clean ones
// are freed (lru_cache.cc:17-24).
__device__ __forceinline__ void evict_apply(const CacheView &c, u32 block, u32 nblocks) {
    CacheRegs *r = c.regs;
    const u32 nv = r->nv;
    const u32 stride = nblocks * blockDim.x;
    for (u32 v0 = block * blockDim.x; v0 < nv; v0 += stride) { // block-uniform trip count
        const u32 v = v0 + threadIdx.x;
        const bool act = v < nv;
        u32 s = 0;
        bool dirty = false;
        if (act) {
            s = c.victims[v];
            ht_erase(c.ht, c.ht_mask, c.slot_key[s]);
            c.slot_prio[s] = PRIO_NONE;
            dirty = c.slot_updates[s] != 0;
            c.slot_state[s] = dirty ? S_PENDING : S_FREE;
        }
        // one atomic per warp and list instead of one per victim on the same two words
        const u32 pp = warp_append(&r->pending, act && dirty);
        if (act && dirty) {
            if (pp < c.capacity)
                c.pending_list[pp] = s;
            else
                atomicMax(&r->error, (u32)E_EVICT_OVERFLOW);
        }
        const u32 fp = warp_append(&r->free_top, act && !dirty);
        if (act && !dirty)
            c.free_stack[fp] = s;
    }
}

__device__ __forceinline__ void insert_new(const CacheView &c, const i32 *uslot, const u32 *miss_list,
                                           u32 block, u32 nblocks) {
    CacheRegs *r = c.regs;
    const u32 M = r->M, n_drop = r->n_drop;
    const u64 clock0 = r->ins_clock0;
    const u32 use0 = c.policy == HB_POLICY_LFU ? 1u : 0u;
    u32 fresh = 0;
    for (u32 j = block * blockDim.x + threadIdx.x; j < M; j += nblocks * blockDim.x) {
        const u32 s = (u32)uslot[miss_list[j]];
        if (j < n_drop) { // served to the caller but never resident after the call
            c.slot_state[s] = S_FREE;
            c.free_stack[atomicAdd(&r->free_top, 1u)] = s;
            continue;
        }
        if (!ht_insert(c.ht, c.ht_mask, c.slot_key[s], s, fresh))
            atomicMax(&r->error, (u32)E_INDEX_FULL);
        c.slot_use[s] = use0;
        c.slot_prio[s] = make_prio(use0, clock0 + j);
        c.stamp_log[(clock0 + j) & c.log_mask] = s;
        c.slot_state[s] = S_CACHED;
    }
    add_occupied(&r->ht_occupied, fresh);
}

// One launch for both halves of the batched insert: the first `evict_blocks` blocks take the
// victims out of the index, the others put the new lines in.  The two may interleave on the
// open-addressing index: an insert only fills EMPTY or TOMB entries and an erase only turns its
// own key into TOMB, so no probe chain ever gains an EMPTY entry in front of a live key; the
// slots of the new lines were taken from the free stack before (resolve_kernel), the victims'
// slots return to it for later calls.
__global__ void __launch_bounds__(256)
    evict_insert_kernel(CacheView c, const i32 *uslot, const u32 *miss_list, u32 evict_blocks) {
    pdl_enter();
    if (blockIdx.x < evict_blocks)
        evict_apply(c, blockIdx.x, evict_blocks);
    else
        insert_new(c, uslot, miss_list, blockIdx.x - evict_blocks, gridDim.x - evict_blocks);
}

// =====================================================================================
// update: accumulate per unique row in occurrence order, fused with the push to the owner
// =====================================================================================
// SCALED / SPLIT are compile-time: the common call (gradient already scaled, occurrence order) runs a
// kernel that carries neither the multiply nor the two-level code — both cost registers in a kernel
// that sits at the 128-register limit of two CTAs per SM (measured: 80 -> 94 us with both compiled in).
template <int VEC, bool SCALED = false, bool SPLIT = false>
struct AccumulatePush {
    using V = RowVec<VEC>;
    struct Acc {
        typename V::T d, g, t; // cache row, pending gradient, owner row (read ahead for the push)
    };
    // what the data phase needs of a line (16 bytes, travels in the work item)
    enum : u8 { X_GRAD = 1, X_DATALESS = 2, X_PUSHED = 4, X_LOCAL = 8 };
    struct Ctx {
        i32 s;
        u32 trow; // row inside the owner's shard (a shard holds < 2^32 rows)
        u32 mpos; // multi-GPU: slot in the owner's mailbox (batch section)
        u8 bits;  // X_GRAD: the line holds a pending gradient to continue from
        u8 owner;
        u8 pad[2];
    };
    CacheView c;
    const u64 *uniq;
    const i32 *uslot;
    i64 push_bound;
    const u64 *push_keys; // non-null: Laia/Herald plan (cache.cc:286-301)
    u32 n_push;
    // push_pull (cache.cc:356-422) syncs and inserts BEFORE the post-push cleanup: a pushed line
    // keeps its update count and gradient until cleanup_pushed_kernel, so the sync sees it stale
    // and adds the gradient again, and an eviction in between counts it as dirty — as the
    // reference does.
    bool defer_cleanup;
    // every gradient value is multiplied by this first (exact fp32 product, then the exact add):
    // the -lr fold of ParameterServerCommunicate.py:24,58-59, which the reference does on the host
    // over the whole gradient before the push; 1 = the gradient arrives scaled (x * 1.0f == x)
    float scale;

    __device__ void kernel_begin() const {
        if (threadIdx.x == 0)
            *block_counter() = *block_counter2() = 0;
        __syncthreads();
    }
    __device__ void kernel_end() const {
        __syncthreads();
        if (threadIdx.x == 0 && *block_counter())
            atomicAdd(&c.regs->pushed, *block_counter());
        if (threadIdx.x == 0 && *block_counter2())
            atomicAdd(&c.regs->pushed_remote, *block_counter2());
    }
    __device__ bool in_plan(u64 key) const {
        u32 lo = 0, hi = n_push;
        while (lo < hi) {
            u32 mid = (lo + hi) >> 1;
            if (push_keys[mid] < key)
                lo = mid + 1;
            else
                hi = mid;
        }
        return lo < n_push && push_keys[lo] == key;
    }
    __device__ MailboxSection outbox(int owner) const {
        return mailbox_section(c.pv.out[owner], 0, c.pv.cap, c.width);
    }
    // One thread per unique key (plan kernel): the push decision and every per-line / per-row
    // scalar of the call — update count, version, owner version, mailbox header — are settled
    // here; the data kernel only moves rows.
    __device__ bool open(size_t u, u32 cnt, Ctx &x) const {
        x.s = uslot[u];
        const u64 key = uniq[u];
        x.owner = 0;
        x.mpos = 0;
        x.bits = 0;
        x.pad[0] = x.pad[1] = 0;
        u64 trow = 0;
        bool local = false;
        if (c.pv.world > 1) { // every line goes through the owner's mailbox, the local ones too:
                              // the owner applies all sources in rank order
            if (key >= c.table_len)
                return false;
            x.owner = (u8)owner_of(c.pv, key, trow);
            x.mpos = (u32)u - c.pv.lo[x.owner];
            if (x.mpos >= c.pv.cap) {
                atomicMax(&c.regs->error, (u32)E_MAILBOX);
                return false;
            }
            if (x.s < 0) { // no line (the call already failed): the slot must still read "not pushed"
                outbox(x.owner).upd[x.mpos] = 0;
                return false;
            }
        } else {
            if (x.s < 0)
                return false;
            trow = key - c.row_begin;
            local = trow < c.nrows_local;
        }
        x.trow = (u32)trow;
        const i32 upd0 = c.slot_updates[x.s];
        const u8 flags = c.slot_flags[x.s];
        const i64 version = c.slot_version[x.s];
        const i32 upd = upd0 + (i32)cnt;
        const bool dataless = flags & F_DATALESS;
        bool pushed;
        if (push_keys)
            pushed = !dataless && in_plan(key); // cache.cc:296
        else
            pushed = (i64)upd > push_bound || dataless; // cache.cc:157
        x.bits = (upd0 != 0 ? X_GRAD : 0) | (dataless ? X_DATALESS : 0) | (pushed ? X_PUSHED : 0) |
                 (local ? X_LOCAL : 0);
        if (!(flags & F_GRAD))
            c.slot_flags[x.s] = flags | F_GRAD;
        if (c.pv.world > 1) {
            const MailboxSection m = outbox(x.owner);
            m.key[x.mpos] = trow;
            m.upd[x.mpos] = pushed ? upd : 0; // 0 = slot not pushed this call
        }
        if (pushed) {
            if (local)
                c.tver[trow] += upd; // PSFhandle_embedding.cc:24
            atomicAdd(block_counter(), 1u);
            if (c.pv.world > 1 && x.owner != c.pv.rank)
                atomicAdd(block_counter2(), 1u);
        }
        if (defer_cleanup) {
            c.slot_updates[x.s] = upd;
        } else if (push_keys) { // cache.cc:308-314: every touched line, every call
            c.slot_version[x.s] = version + upd;
            c.slot_updates[x.s] = pushed ? 0 : upd;
        } else if (pushed && !dataless) { // cache.cc:171-177
            c.slot_version[x.s] = version + upd;
            c.slot_updates[x.s] = 0;
        } else {
            c.slot_updates[x.s] = upd;
        }
        return true;
    }
    __device__ Acc load(const Ctx &x, size_t k) const {
        Acc a;
        const size_t o = (size_t)x.s * c.width + k * VEC;
        a.d = (x.bits & X_DATALESS) ? V::zero() : V::ld_rmw(c.data + o);
        a.g = (x.bits & X_GRAD) ? V::ld_rmw(c.grad + o) : V::zero();
        a.t = (x.bits & (X_PUSHED | X_LOCAL)) == (X_PUSHED | X_LOCAL)
                  ? V::ld_rmw(c.trows + (size_t)x.trow * c.width + k * VEC)
                  : V::zero();
        return a;
    }
    __device__ Acc step(const Acc &a, const typename V::T &g) const {
        Acc r; // embedding.h:78-91: grad_ += g; data_ += g  (per occurrence, in order)
        if constexpr (SCALED) {
            const typename V::T gs = V::mul(g, scale);
            r.g = V::add(a.g, gs);
            r.d = V::add(a.d, gs);
        } else {
            r.g = V::add(a.g, g);
            r.d = V::add(a.d, g);
        }
        r.t = a.t;
        return r;
    }
    // two-level reduction of the very hot rows (hb_rows.cuh, opt-in by hb_cache_set_reduce_mode)
    static constexpr bool kSplit = SPLIT;
    __device__ float pre(float g) const { // one occurrence's contribution
        return SCALED ? __fmul_rn(g, scale) : g;
    }
    __device__ Acc step_pre(const Acc &a, const typename V::T &p) const { // a run sum, already scaled
        Acc r;
        r.g = V::add(a.g, p);
        r.d = V::add(a.d, p);
        r.t = a.t;
        return r;
    }
    __device__ void store(const Ctx &x, size_t k, const Acc &a) const {
        const size_t o = (size_t)x.s * c.width + k * VEC;
        const bool dataless = x.bits & X_DATALESS, pushed = x.bits & X_PUSHED;
        if (!dataless)
            V::st(c.data + o, a.d);
        if (pushed && (x.bits & X_LOCAL)) // PSFhandle_embedding.cc:25-26: row += pushed grad
            V::st(c.trows + (size_t)x.trow * c.width + k * VEC, V::add(a.t, a.g));
        else if (pushed && c.pv.world > 1) // deposit the pushed gradient at the owner (NVLink store)
            V::st(outbox(x.owner).grad + (size_t)x.mpos * c.width + k * VEC, a.g);
        if (!pushed || (defer_cleanup && !dataless))
            V::st(c.grad + o, a.g);
    }
};

// Flush the dirty victims collected since the last push (evict_, cache.cc:142-166): their whole
// pending gradient and update count go to the owner row; the slot is freed.
template <int VEC>
struct FlushPending {
    using V = RowVec<VEC>;
    CacheView c;
    u32 count;
    __device__ bool begin(size_t) const {
        return true;
    }
    __device__ void apply(size_t e, size_t k) const {
        const u32 s = c.pending_list[e];
        const u64 trow = c.slot_key[s] - c.row_begin;
        // updates == 0: a line that push_pull evicted right after pushing it; the reference
        // re-pushes its zeroed gradient, which changes nothing
        if (trow >= c.nrows_local || c.slot_updates[s] == 0)
            return;
        float *t = c.trows + trow * c.width + k * VEC;
        V::st(t, V::add(V::ld(t), V::ld(c.grad + (size_t)s * c.width + k * VEC)));
    }
    __device__ void end(size_t e) const {
        if (lane_id() != 0)
            return;
        const u32 s = c.pending_list[e];
        const u64 trow = c.slot_key[s] - c.row_begin;
        if (trow < c.nrows_local)
            c.tver[trow] += c.slot_updates[s];
        c.slot_updates[s] = 0;
        c.slot_state[s] = S_FREE;
        c.free_stack[atomicAdd(&c.regs->free_top, 1u)] = s;
    }
};


// ---- multi-GPU exchange -----------------------------------------------------------------
// lo[o] = first unique index whose key belongs to owner o: the sorted uniques split into
// contiguous per-owner slices exactly as PSAgent splits them with lower_bound (PSAgent.h:541-559)
// Also the first kernel of an update that touches peer memory: it waits until every owner has
// applied the previous exchange (`prev_epoch`), i.e. until the mailboxes may be overwritten.
__global__ void owner_bounds_kernel(CacheView c, const u64 *__restrict__ uniq,
                                    const u32 *__restrict__ num_unique, u64 prev_epoch) {
    pdl_enter();
    wait_flags(c, c.pv.ctrl + kMaxWorld, prev_epoch);
    const int o = threadIdx.x;
    if (o > c.pv.world)
        return;
    const u32 U = *num_unique;
    u32 lo = 0, hi = U;
    {
        // the last bound stops at the table length: keys beyond it (a failed call) have no owner
        const u64 target = o == c.pv.world ? c.table_len : shard_begin(c.pv, o);
        while (lo < hi) {
            const u32 mid = (lo + hi) >> 1;
            if (uniq[mid] < target)
                lo = mid + 1;
            else
                hi = mid;
        }
    }
    c.pv.lo[o] = lo;
    if (o < c.pv.world)
        c.pv.fl_count[o] = 0;
}

// Dirty victims go to their owner's "flush" section (one warp per line).
template <int VEC>
__global__ void __launch_bounds__(kRowBlock) flush_remote_kernel(CacheView c) {
    pdl_enter();
    using V = RowVec<VEC>;
    const unsigned lane = lane_id();
    const size_t warp_global = (size_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * kRowWarps;
    const u32 count = c.regs->flushed;
    const size_t D = c.width, nvec = D / VEC;
    for (size_t e = warp_global; e < count; e += nwarps) {
        const u32 s = c.pending_list[e];
        i32 upd = c.slot_updates[s];
        u64 trow;
        int owner = 0;
        if (c.slot_key[s] < c.table_len)
            owner = owner_of(c.pv, c.slot_key[s], trow);
        else
            upd = 0;
        // updates == 0: a line push_pull evicted right after pushing it; nothing to send
        u32 pos = 0;
        if (lane == 0 && upd != 0)
            pos = atomicAdd(&c.pv.fl_count[owner], 1u);
        pos = __shfl_sync(FULL, pos, 0);
        if (upd != 0) {
            if (pos < c.pv.cap) {
                const MailboxSection m = mailbox_section(c.pv.out[owner], 1, c.pv.cap, D);
                for (size_t k = lane; k < nvec; k += 32)
                    V::st(m.grad + (size_t)pos * D + k * VEC, V::ld(c.grad + (size_t)s * D + k * VEC));
                if (lane == 0) {
                    m.key[pos] = trow;
                    m.upd[pos] = upd;
                }
            } else if (lane == 0) {
                atomicMax(&c.regs->error, (u32)E_MAILBOX);
            }
        }
        if (lane == 0) {
            c.slot_updates[s] = 0;
            c.slot_state[s] = S_FREE;
            c.free_stack[atomicAdd(&c.regs->free_top, 1u)] = s;
        }
    }
}

// Checkpoint support: push every RESIDENT dirty line (updates != 0) to its owner and mark it
// clean, exactly what a push does to a line over the bound (cache.cc:171-177,
// PSFhandle_embedding.cc:24-26).  The reference has no such call — ParamSave writes the server
// table while the workers' pending updates stay in their caches.  A warp takes 32 slots, tests
// them lane-parallel and walks the dirty ones.  Multi-GPU: ranks take turns (host barriers), the
// owner row is updated in place through the peer mapping.
template <int VEC>
__global__ void __launch_bounds__(kRowBlock) flush_resident_kernel(CacheView c) {
    pdl_enter();
    using V = RowVec<VEC>;
    const unsigned lane = lane_id();
    const size_t warp_global = (size_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * kRowWarps;
    const size_t D = c.width, nvec = D / VEC;
    const size_t hw = c.regs->slot_hw;
    u32 flushed = 0;
    for (size_t base = warp_global * 32; base < hw; base += nwarps * 32) {
        const size_t s = base + lane;
        i32 upd = 0;
        u64 key = 0;
        if (s < hw) {
            const u8 st = c.slot_state[s];
            if (st == S_CACHED || st == S_STORE) {
                upd = c.slot_updates[s];
                key = c.slot_key[s];
                if (key >= c.table_len)
                    upd = 0;
            }
        }
        unsigned m = __ballot_sync(FULL, upd != 0);
        while (m) {
            const int from = __ffs(m) - 1;
            m &= m - 1;
            const u64 k = __shfl_sync(FULL, key, from);
            const i32 u = __shfl_sync(FULL, upd, from);
            const size_t slot = base + from;
            float *rows = c.trows;
            i64 *ver = c.tver;
            u64 trow = k - c.row_begin;
            if (c.pv.world > 1) {
                const int owner = owner_of(c.pv, k, trow);
                rows = const_cast<float *>(c.pv.rows[owner]);
                ver = const_cast<i64 *>(c.pv.ver[owner]);
            } else if (trow >= c.nrows_local) {
                continue;
            }
            for (size_t q = lane; q < nvec; q += 32) {
                float *dst = rows + trow * D + q * VEC;
                V::st(dst, V::add(V::ld(dst), V::ld(c.grad + slot * D + q * VEC)));
            }
            if (lane == 0) {
                ver[trow] += u;
                c.slot_version[slot] += u;
                c.slot_updates[slot] = 0;
                flushed++;
            }
        }
    }
    if (lane == 0 && flushed)
        atomicAdd(&c.regs->pushed, flushed);
}

// End of the sending half of an exchange: tell every owner how many slots of its two sections this
// rank filled, then raise this rank's `ready` flag at every owner (one launch, after the kernels
// that wrote the mailboxes).
__global__ void exchange_arrive_kernel(CacheView c, u64 epoch) {
    pdl_enter();
    const int o = threadIdx.x;
    if (o >= c.pv.world)
        return;
    u32 *hdr = reinterpret_cast<u32 *>(c.pv.out[o]);
    hdr[0] = c.pv.lo[o + 1] - c.pv.lo[o];
    hdr[1] = min(c.pv.fl_count[o], c.pv.cap);
    __threadfence_system(); // the mailbox stores of the earlier kernels and the header above
    raise_flag(c.pv.ctrl_peer[o] + c.pv.rank, epoch);
}

// Owner side (PSFhandle_embedding.cc:5-28): row += pushed grad; ver += updates.  Several sources
// push the same hot rows and fp32 adds do not commute, so the adds on one row must happen in a
// fixed order: source rank, batch section before flush section, slot — the order a serial server
// fed by the ranks in turn would apply them in.  Two kernels over ALL sources and sections:
//   link   one thread per mailbox entry: the pushed ones are chained per row through an atomic
//          exchange on head[row]; the entry that finds the row unclaimed will apply it;
//   apply  the claiming entry's warp walks the row's chain (<= 2 x world entries unless one source
//          flushed the same key twice), orders it, and does one read-modify-write of the row with
//          the gradients added in that order; the last CTA raises `applied` at every peer.
// An entry id is list * cap + slot with list = 2 * src + section, so ascending ids are the order.
constexpr u32 kNotPushed = 0xffffffffu;
constexpr int kMaxChain = 24; // 2 x 8 ranks + room for a key one source flushed more than once

__global__ void __launch_bounds__(256) link_mailbox_kernel(CacheView c, u64 epoch) {
    pdl_enter();
    wait_flags(c, c.pv.ctrl, epoch); // every source has completed its mailbox here
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
        reinterpret_cast<u32 *>(c.pv.ctrl + 2 * kMaxWorld)[0] = 0; // the apply kernel's CTA counter
    const int list = blockIdx.y, src = list >> 1, section = list & 1;
    char *region = c.pv.in + (size_t)src * c.pv.region_bytes;
    const u32 count = min(reinterpret_cast<const u32 *>(region)[section], c.pv.cap);
    const MailboxSection mb = mailbox_section(region, section, c.pv.cap, c.width);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const u32 e = (u32)list * c.pv.cap + i;
        const u64 trow = mb.key[i];
        u32 link = kNotPushed;
        if (mb.upd[i] != 0 && trow < c.nrows_local)
            link = atomicExch(&c.pv.head[trow], e + 1);
        c.pv.next[e] = link;
    }
}

template <int VEC, int ROWS>
__global__ void __launch_bounds__(kRowBlock) apply_linked_kernel(CacheView c, u64 epoch) {
    pdl_enter();
    using V = RowVec<VEC>;
    __shared__ u32 s_chain[kRowWarps][kMaxChain][32]; // [warp][position][lane]: a lane's chain, ascending ids
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const size_t warp_global = (size_t)blockIdx.x * kRowWarps + warp;
    const size_t nwarps = (size_t)gridDim.x * kRowWarps;
    const size_t D = c.width, nvec = D / VEC;
    const u32 cap = c.pv.cap;
    const int list = blockIdx.y, src0 = list >> 1, section0 = list & 1;
    char *region0 = c.pv.in + (size_t)src0 * c.pv.region_bytes;
    const u32 count = min(reinterpret_cast<const u32 *>(region0)[section0], cap);
    const MailboxSection mb0 = mailbox_section(region0, section0, cap, D);
    auto section_of = [&](u32 e, u32 &slot) {
        const u32 l = e / cap;
        slot = e - l * cap;
        return mailbox_section(c.pv.in + (size_t)(l >> 1) * c.pv.region_bytes, (int)(l & 1), cap, D);
    };
    for (size_t base = warp_global * 32; base < count; base += nwarps * 32) {
        const size_t i = base + lane;
        // the entry that found its row unclaimed applies the row: every lane walks the chain of its
        // own row (a handful of dependent loads), keeping the ids in ascending order
        const bool mine = i < count && c.pv.next[(size_t)list * cap + i] == 0;
        u64 my_row = 0;
        u32 n = 0;
        if (mine) {
            my_row = mb0.key[i];
            u32 e1 = c.pv.head[my_row];
            i64 upd_sum = 0;
            while (e1 != 0) {
                const u32 e = e1 - 1;
                if (n == (u32)kMaxChain) { // absurdly long chain: reported, the excess is not applied
                    atomicMax(&c.regs->xerror, (u32)E_MAILBOX);
                    break;
                }
                u32 j = n;
                while (j > 0 && s_chain[warp][j - 1][lane] > e) {
                    s_chain[warp][j][lane] = s_chain[warp][j - 1][lane];
                    j--;
                }
                s_chain[warp][j][lane] = e;
                n++;
                u32 slot;
                const MailboxSection mb = section_of(e, slot);
                upd_sum += mb.upd[slot];
                e1 = c.pv.next[e];
            }
            c.pv.head[my_row] = 0;
            c.tver[my_row] += upd_sum;
        }
        __syncwarp();
        unsigned m = __ballot_sync(FULL, mine);
        while (m) {
            int from[ROWS];
            u64 rt[ROWS];
            u32 rn[ROWS];
            u32 max_n = 0;
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
                from[r] = m ? __ffs(m) - 1 : -1;
                if (m)
                    m &= m - 1;
                const int f = from[r] < 0 ? 0 : from[r];
                rt[r] = __shfl_sync(FULL, my_row, f);
                rn[r] = from[r] < 0 ? 0u : __shfl_sync(FULL, n, f);
                max_n = max(max_n, rn[r]);
            }
            for (size_t k = lane; k < nvec; k += 32) {
                typename V::T acc[ROWS], g[ROWS];
#pragma unroll
                for (int r = 0; r < ROWS; r++)
                    if (from[r] >= 0)
                        acc[r] = V::ld(c.trows + rt[r] * D + k * VEC);
                for (u32 j = 0; j < max_n; j++) {
#pragma unroll
                    for (int r = 0; r < ROWS; r++)
                        if (j < rn[r]) {
                            u32 slot;
                            const MailboxSection mb = section_of(s_chain[warp][j][from[r]], slot);
                            g[r] = V::ld(mb.grad + (size_t)slot * D + k * VEC);
                        }
#pragma unroll
                    for (int r = 0; r < ROWS; r++)
                        if (j < rn[r])
                            acc[r] = V::add(acc[r], g[r]);
                }
#pragma unroll
                for (int r = 0; r < ROWS; r++)
                    if (from[r] >= 0)
                        V::st(c.trows + rt[r] * D + k * VEC, acc[r]);
            }
        }
        __syncwarp();
    }
    // the last CTA of the grid tells every peer that this shard is up to date
    __shared__ u32 s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 *done = reinterpret_cast<u32 *>(c.pv.ctrl + 2 * kMaxWorld);
        s_last = atomicAdd(done, 1u) == gridDim.x * gridDim.y - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x < (unsigned)c.pv.world) {
        __threadfence_system();
        raise_flag(c.pv.ctrl_peer[threadIdx.x] + kMaxWorld + c.pv.rank, epoch);
    }
}

// dataless lines are dropped after the push (never inserted)
__global__ void free_transient_kernel(CacheView c, const i32 *uslot, const u32 *miss_list, int batch,
                                      int flushed) {
    pdl_enter();
    CacheRegs *r = c.regs;
    if (flushed && blockIdx.x == 0 && threadIdx.x == 0)
        r->pending = 0; // the pending victims have just been pushed
    const u32 M = batch == 0 ? r->M : r->M2;
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
        const u32 s = (u32)uslot[miss_list[j]];
        c.slot_state[s] = S_FREE;
        c.slot_updates[s] = 0;
        c.slot_flags[s] = 0;
        c.free_stack[atomicAdd(&r->free_top, 1u)] = s;
    }
}

// The tail of a single-GPU update in ONE launch: flush of the dirty victims collected since the last
// push (FlushPending), release of the dataless lines of this call's misses (free_transient), and —
// by whichever CTA finishes last — the call's epilogue (op_end): three kernels and two launch gaps
// less on the step's critical path.
template <int VEC>
__global__ void __launch_bounds__(kRowBlock)
    update_tail_kernel(CacheView c, const i32 *uslot, const u32 *miss_list, const u64 *clk, int last_stage,
                       PerfRecord *rec, u32 num_all) {
    pdl_enter();
    CacheRegs *r = c.regs;
    const unsigned lane = lane_id();
    {   // pending victims: one warp per line
        FlushPending<VEC> f{c, 0};
        const size_t warp_global = (size_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
        const size_t nwarps = (size_t)gridDim.x * kRowWarps;
        const size_t nvec = c.width / VEC;
        const size_t R = r->flushed;
        for (size_t e = warp_global; e < R; e += nwarps) {
            for (size_t k = lane; k < nvec; k += 32)
                f.apply(e, k);
            __syncwarp();
            f.end(e);
        }
    }
    const u32 M = r->M;
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
        const u32 s = (u32)uslot[miss_list[j]];
        c.slot_state[s] = S_FREE;
        c.slot_updates[s] = 0;
        c.slot_flags[s] = 0;
        c.free_stack[atomicAdd(&r->free_top, 1u)] = s;
    }
    __shared__ u32 s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
        s_last = atomicAdd(&r->tail_done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        r->pending = 0; // the pending victims have just been pushed
        op_end_body(c, clk, last_stage, rec, 1, num_all, 0);
    }
}

// Deferred cleanup of push_pull (cache.cc:413-421): version += updates; zeroGrad — on every line
// of the push batch that was pushed and holds data, wherever it is now (resident or evicted).
__global__ void cleanup_pushed_kernel(CacheView c, const i32 *uslot, i64 push_bound) {
    pdl_enter();
    const u32 U = c.regs->U2;
    for (u32 u = blockIdx.x * blockDim.x + threadIdx.x; u < U; u += gridDim.x * blockDim.x) {
        const i32 s = uslot[u];
        if (s < 0 || c.slot_state[s] == S_FREE) // dataless lines were dropped after the push
            continue;
        const i32 upd = c.slot_updates[s];
        if ((i64)upd > push_bound) {
            c.slot_version[s] += upd;
            c.slot_updates[s] = 0;
        }
    }
}

__global__ void convert_keys_kernel(const float *in, u64 *out, size_t n) {
    pdl_enter();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = key_from_f32(in[i]);
}

// =====================================================================================
// maintenance / debug kernels
// =====================================================================================
__global__ void init_slots_kernel(CacheView c) {
    pdl_enter();
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < c.capacity;
         s += (size_t)gridDim.x * blockDim.x) {
        c.slot_prio[s] = PRIO_NONE;
        c.slot_state[s] = S_FREE;
        c.slot_flags[s] = 0;
        c.slot_updates[s] = 0;
        c.slot_use[s] = 0;
        c.slot_version[s] = -1;
        c.slot_key[s] = HT_EMPTY;
        c.free_stack[s] = (u32)(c.capacity - 1 - s); // slot 0 is handed out first
    }
}

// new slots [old_cap, c.capacity) of a grown row store: free, at the BOTTOM of the free stack (the
// stack keeps handing out the lowest never-used slot first, which slot_hw relies on); the old
// stack's `top` entries move up by the number of new slots
__global__ void grow_slots_kernel(CacheView c, u32 old_cap, const u32 *old_stack, u32 top) {
    pdl_enter();
    const u32 delta = c.capacity - old_cap;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < c.capacity;
         i += (size_t)gridDim.x * blockDim.x) {
        if (i >= old_cap) {
            const size_t s = i;
            c.slot_prio[s] = PRIO_NONE;
            c.slot_state[s] = S_FREE;
            c.slot_flags[s] = 0;
            c.slot_updates[s] = 0;
            c.slot_use[s] = 0;
            c.slot_version[s] = -1;
            c.slot_key[s] = HT_EMPTY;
        }
        if (i < delta)
            c.free_stack[i] = (u32)(c.capacity - 1 - i);
        else if (i - delta < top)
            c.free_stack[i] = old_stack[i - delta];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        c.regs->free_top = top + delta;
}

__global__ void rebuild_index_kernel(CacheView c) {
    pdl_enter();
    u32 fresh = 0;
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < c.capacity;
         s += (size_t)gridDim.x * blockDim.x) {
        u8 stt = c.slot_state[s];
        if (stt == S_CACHED || stt == S_STORE)
            if (!ht_insert(c.ht, c.ht_mask, c.slot_key[s], (u32)s, fresh))
                atomicMax(&c.regs->error, (u32)E_INDEX_FULL);
    }
    add_occupied(&c.regs->ht_occupied, fresh);
}

__global__ void collect_keys_kernel(CacheView c, u64 *out, u32 *count) {
    pdl_enter();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t rounds = (c.capacity + stride - 1) / stride;
    for (size_t it = 0; it < rounds; it++) {
        const size_t s = it * stride + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        bool in = false;
        if (s < c.capacity) {
            u8 stt = c.slot_state[s];
            in = stt == S_CACHED || stt == S_STORE;
        }
        u32 pos = warp_append(count, in);
        if (in)
            out[pos] = c.slot_key[s];
    }
}

// ---- Laia / Herald scoring against the REAL cache index ----------------------------
// The reference scores a sample for a worker by counting its embedding ids in a host-side SIMULATION
// of that worker's cache (MiniLRUCache snapshots: laia_scheduler.cc:171-210, topk_scheduler.cc:
// 405-428).  With the cache in HBM the worker can answer from the index itself: one warp takes one
// sample row, lane j probes table order[j] (the first `top_k` tables of the planner's order), a
// ballot counts the resident ids.  `fresh` != 0 counts only lines a lookup would NOT have to
// re-pull (version within pull_bound of the owner's: the snapshots' "valid" bit).  Read-only.
template <int KIND>
__global__ void __launch_bounds__(256)
    score_samples_kernel(CacheView c, const void *__restrict__ ids, size_t num_samples, u32 num_tables,
                         const u32 *__restrict__ order, u32 top_k, int fresh, i64 pull_bound,
                         u32 *__restrict__ scores) {
    pdl_enter();
    const unsigned lane = lane_id();
    const size_t warp_global = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
    for (size_t i = warp_global; i < num_samples; i += nwarps) {
        u32 count = 0;
        for (u32 j0 = 0; j0 < top_k; j0 += 32) {
            const u32 j = j0 + lane;
            bool hit = false;
            if (j < top_k) {
                const u32 t = order ? order[j] : j;
                const size_t e = i * num_tables + t;
                const u64 key = KIND == HB_KEYS_F32 ? key_from_f32(reinterpret_cast<const float *>(ids)[e])
                                                    : reinterpret_cast<const u64 *>(ids)[e];
                const i32 sl = ht_find(c.ht, c.ht_mask, key);
                hit = sl >= 0;
                if (hit && fresh) {
                    const i64 v = c.slot_version[sl];
                    u64 trow = key - c.row_begin;
                    int owner = 0;
                    if (c.pv.world > 1)
                        owner = owner_of(c.pv, key, trow);
                    hit = v != -1 && key < c.table_len && __ldg(&c.pv.ver[owner][trow]) - v <= pull_bound;
                }
            }
            count += __popc(__ballot_sync(FULL, hit));
        }
        if (lane == 0)
            scores[i] = count;
    }
}

// resident[i] = 1 if keys[i] has a line in the index (CacheBase::count, python_api.cc:56), else 0
template <int KIND>
__global__ void __launch_bounds__(256)
    probe_keys_kernel(CacheView c, const void *__restrict__ keys, size_t n, u8 *__restrict__ resident) {
    pdl_enter();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const u64 key = KIND == HB_KEYS_F32 ? key_from_f32(reinterpret_cast<const float *>(keys)[i])
                                            : reinterpret_cast<const u64 *>(keys)[i];
        resident[i] = ht_find(c.ht, c.ht_mask, key) >= 0 ? 1 : 0;
    }
}

struct PeekResult {
    i32 slot;
    i32 updates;
    i64 version;
    u32 flags;
    u32 state;
};
__global__ void peek_kernel(CacheView c, u64 key, PeekResult *out) {
    pdl_enter();
    i32 s = ht_find(c.ht, c.ht_mask, key);
    out->slot = s;
    if (s >= 0) {
        out->updates = c.slot_updates[s];
        out->version = c.slot_version[s];
        out->flags = c.slot_flags[s];
        out->state = c.slot_state[s];
    }
}

// single-line insert(Embedding) support: overwrite or stage one line
__global__ void set_line_kernel(CacheView c, i32 slot, i64 version, const float *data) {
    pdl_enter();
    for (u32 k = threadIdx.x; k < c.width; k += blockDim.x)
        c.data[(size_t)slot * c.width + k] = data[k];
    if (threadIdx.x == 0) {
        c.slot_version[slot] = version;
        c.slot_updates[slot] = 0;
        c.slot_flags[slot] = 0;
    }
}
// policy effect of re-inserting a resident key (lru_cache.cc:11-16, lfu_cache.cc:16-19,
// lfuopt_cache.cc:10-17)
__global__ void reinsert_touch_kernel(CacheView c, i32 s) {
    pdl_enter();
    CacheRegs *r = c.regs;
    const u64 stamp = r->clock;
    c.stamp_log[stamp & c.log_mask] = (u32)s;
    if (c.policy == HB_POLICY_LRU) {
        c.slot_prio[s] = make_prio(0, stamp);
        r->clock = stamp + 1;
    } else if (c.policy == HB_POLICY_LFU) {
        u32 use = c.slot_use[s] + 1;
        c.slot_use[s] = use;
        c.slot_prio[s] = make_prio(use, stamp);
        r->clock = stamp + 1;
    }
}
__global__ void single_key_kernel(u64 *uniq, u32 *num_unique, u64 key) {
    pdl_enter();
    uniq[0] = key;
    *num_unique = 1;
}
__global__ void read_slot_kernel(const i32 *uslot, i32 *out) {
    pdl_enter();
    *out = uslot[0];
}

// ---- sparse read of an owner shard: out[i,:] = rows[keys[i] - row_begin,:], ver likewise ----
struct IndexFromShardKeys {
    const u64 *keys;
    u64 row_begin, nrows;
    __device__ long long operator()(size_t n) const {
        const u64 r = keys[n] - row_begin; // wraps for keys below the shard
        return r < nrows ? (long long)r : -1;
    }
};
__global__ void gather_versions_kernel(const i64 *ver, const u64 *keys, size_t n, u64 row_begin,
                                       u64 nrows, i64 *out) {
    pdl_enter();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const u64 r = keys[i] - row_begin;
        out[i] = r < nrows ? ver[r] : -1;
    }
}

// ---- table init: counter-based generator (splitmix64 of (seed, element index)) -----------
__host__ __device__ __forceinline__ u64 splitmix64(u64 x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ float u01(u64 bits) { // (0,1]
    return ((float)(bits >> 40) + 1.0f) * (1.0f / 16777216.0f);
}
__global__ void table_init_kernel(float *rows, size_t nelem, size_t elem_begin, int init_type,
                                  float a, float b, u64 seed) {
    pdl_enter();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nelem;
         i += (size_t)gridDim.x * blockDim.x) {
        const u64 g = elem_begin + i;
        float v;
        if (init_type == 0) {
            v = a;
        } else if (init_type == 1) {
            float u = u01(splitmix64(seed ^ (g * 2))) - (1.0f / 33554432.0f);
            v = a + (b - a) * u;
        } else {
            u64 ctr = 0;
            while (true) { // Box-Muller; truncated normal redraws outside 2 sigma
                float u1 = u01(splitmix64(seed ^ (g * 2) ^ (ctr << 56)));
                float u2 = u01(splitmix64(seed ^ (g * 2 + 1) ^ (ctr << 56)));
                float z = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
                v = a + b * z;
                if (init_type != 3 || fabsf(z) <= 2.0f)
                    break;
                ctr++;
            }
        }
        rows[i] = v;
    }
}

// =====================================================================================
// host side
// =====================================================================================
std::mutex g_tables_mtx;
std::map<int, hb_table *> g_tables;

bool is_device_ptr(const void *p) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

template <typename T>
void dmalloc(T *&p, size_t count) {
    HB_CUDA(cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T)));
}
// Memory other ranks map through CUDA IPC (shards, versions, mailboxes): the size is rounded up to
// a multiple of 2 MiB.  Measured on 2 x B200 (scripts/ipcbench.cu, profiles/r01_ipcbench.txt): a
// peer's mapping of an 8.64 GB cudaMalloc whose size is NOT a multiple of 2 MiB serves random 512 B
// row reads at 38 GB/s (small pages on the importer's side), the same allocation rounded up at
// 431 GB/s.
template <typename T>
void dmalloc_shared(T *&p, size_t count) {
    const size_t two_mb = (size_t)2 << 20;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    bytes = (bytes + two_mb - 1) / two_mb * two_mb;
    HB_CUDA(cudaMalloc((void **)&p, bytes));
}
template <typename T>
void dfree(T *&p) {
    if (p)
        cudaFree(p);
    p = nullptr;
}

inline int lin_grid(size_t n, int block = 256) {
    size_t b = (n + block - 1) / block;
    return (int)std::max<size_t>(1, std::min<size_t>(b, (size_t)sm_count() * 16));
}

struct Guard { // select the cache's device for the duration of a call
    int prev = 0;
    explicit Guard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev)
            cudaSetDevice(dev);
        else
            prev = -1;
    }
    ~Guard() {
        if (prev >= 0)
            cudaSetDevice(prev);
    }
};

// (Re)allocate this rank's mailboxes for `rows` slots per section and map every owner's.
// Collective: every rank calls it with the same `rows`, in the same order.
void setup_mailbox(hb_cache *c, size_t rows) {
    PeerView &pv = c->view.pv;
    if (c->mailbox) {
        HB_CUDA(cudaDeviceSynchronize());
        HB_CHECK(hb_comm_barrier() == 0, "barrier failed");
        ipc_unshare(reinterpret_cast<void **>(c->peer_mailbox));
        cudaFree(c->mailbox);
        c->mailbox = nullptr;
    }
    pv.cap = (u32)rows;
    pv.region_bytes = mailbox_region_bytes(rows, c->width);
    const size_t bytes = kMailboxCtrlBytes + pv.region_bytes * pv.world;
    dmalloc_shared(c->mailbox, bytes);
    HB_CUDA(cudaMemset(c->mailbox, 0, bytes));
    c->mailbox_cap = rows;
    c->xrows_upper = rows;
    c->xepoch = 0; // the flags of the new control block start at zero on every rank
    ipc_share(c->mailbox, reinterpret_cast<void **>(c->peer_mailbox));
    pv.ctrl = reinterpret_cast<u64 *>(c->mailbox);
    pv.in = c->mailbox + kMailboxCtrlBytes;
    for (int o = 0; o < pv.world; o++) {
        pv.ctrl_peer[o] = reinterpret_cast<u64 *>(c->peer_mailbox[o]);
        pv.out[o] = c->peer_mailbox[o] + kMailboxCtrlBytes + (size_t)pv.rank * pv.region_bytes;
    }
    dfree(pv.next);
    dmalloc(pv.next, (size_t)2 * pv.world * rows);
    if (!pv.head) {
        dmalloc(pv.head, (size_t)c->view.nrows_local);
        HB_CUDA(cudaMemset(pv.head, 0, std::max<size_t>(c->view.nrows_local, 1) * sizeof(u32)));
    }
    HB_CHECK(hb_comm_barrier() == 0, "barrier failed"); // every mailbox is zeroed and mapped
}

u64 *clk_of(hb_cache *c) {
    // four u64 right behind the perf record in the same device allocation
    return reinterpret_cast<u64 *>(reinterpret_cast<char *>(c->dev_record) + 64);
}

void sync_all(hb_cache *c) {
    if (c->xstream)
        HB_CUDA(cudaStreamSynchronize(c->xstream));
    HB_CUDA(cudaStreamSynchronize(c->side));
    HB_CUDA(cudaStreamSynchronize(c->side2));
    HB_CUDA(cudaStreamSynchronize(c->h2d));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    HB_CUDA(cudaStreamSynchronize(c->d2h));
}

void ensure_batch(hb_cache *c, size_t n) {
    if (n <= c->batch_cap)
        return;
    sync_all(c);
    size_t cap = std::max<size_t>(n, 4096);
    for (int b = 0; b < 2; b++) {
        c->ws[b].reserve(cap);
        dfree(c->uslot[b]);
        dfree(c->miss_list[b]);
        dmalloc(c->uslot[b], c->ws[b].cap);
        dmalloc(c->miss_list[b], c->ws[b].cap);
    }
    c->batch_cap = c->ws[0].cap;
}

// Both key staging buffers hold at least n keys (called once per call, before any staging, so a
// later batch of the same call never reallocates under an earlier one).
void ensure_keys_stage(hb_cache *c, size_t n) {
    if (n * 8 <= c->keys_stage_cap)
        return;
    sync_all(c);
    size_t cap = std::max<size_t>(n, 4096) * 8;
    for (int b = 0; b < 2; b++) {
        if (c->keys_stage[b])
            cudaFree(c->keys_stage[b]);
        c->keys_stage[b] = nullptr;
        HB_CUDA(cudaMalloc(&c->keys_stage[b], cap));
    }
    c->keys_stage_cap = cap;
}

// keys as given by the caller -> device pointer (staged when they live in host memory).  Only the
// sort reads the raw keys, so the copy goes to the side stream, in front of it.
const void *stage_keys(hb_cache *c, const void *keys, int kind, size_t n, int which) {
    if (n == 0 || is_device_ptr(keys))
        return keys;
    size_t bytes = n * (kind == HB_KEYS_F32 ? 4 : 8);
    HB_CHECK(bytes <= c->keys_stage_cap, "key staging buffer not reserved");
    HB_CUDA(cudaMemcpyAsync(c->keys_stage[which], keys, bytes, cudaMemcpyHostToDevice, c->side));
    return c->keys_stage[which];
}

float *rows_stage(hb_cache *c, size_t n, int which) {
    size_t need = n * c->width;
    if (need > c->rows_stage_cap[which]) {
        sync_all(c);
        dfree(c->rows_stage[which]);
        dmalloc(c->rows_stage[which], need);
        c->rows_stage_cap[which] = need;
    }
    return c->rows_stage[which];
}

void maybe_rebuild_index(hb_cache *c, size_t incoming) {
    c->occ_upper += incoming;
    c->incoming_ring[c->calls % hb_cache::kRing] = incoming;
    if (c->occ_upper * 2 <= c->ht_size)
        return;
    // the records of finished calls carry the real occupancy: refresh the bound from the newest
    // one that has completed before paying for a synchronisation
    for (uint64_t back = 1; back <= std::min<uint64_t>(c->calls, 64); back++) {
        const uint64_t call = c->calls - back;
        const int idx = (int)(call % hb_cache::kRing);
        if (cudaEventQuery(c->ev_end[idx]) != cudaSuccess)
            continue;
        size_t since = incoming;
        for (uint64_t k = call + 1; k < c->calls; k++)
            since += c->incoming_ring[k % hb_cache::kRing];
        c->occ_upper = std::min(c->occ_upper, (size_t)c->ring[idx].ht_occupied + since);
        break;
    }
    (void)cudaGetLastError(); // cudaErrorNotReady of the queries is not an error
    if (c->occ_upper * 2 <= c->ht_size)
        return;
    HB_CUDA(cudaStreamSynchronize(c->stream));
    CacheRegs regs;
    HB_CUDA(cudaMemcpy(&regs, c->view.regs, sizeof(regs), cudaMemcpyDeviceToHost));
    if ((size_t)regs.ht_occupied + incoming > c->ht_size / 2) {
        // tombstones have piled up: clear and re-insert the resident lines
        HB_CUDA(cudaMemsetAsync(c->view.ht, 0xff, c->ht_size * sizeof(HtEntry), c->stream));
        HB_CUDA(cudaMemsetAsync(&c->view.regs->ht_occupied, 0, sizeof(u32), c->stream));
        HB_LAUNCH(rebuild_index_kernel, lin_grid(c->view.capacity), 256, 0, c->stream, c->view);
        HB_LAUNCHED();
        c->occ_upper = regs.size + incoming;
    } else {
        c->occ_upper = regs.ht_occupied + incoming;
    }
}

// Grow the row store to `new_cap` slots (rare: the slack is sized by hb_cache_reserve / the first
// calls).  Synchronises; every slot array is reallocated and copied.
template <typename T>
void regrow(T *&p, size_t old_count, size_t new_count, bool keep) {
    T *q = nullptr;
    dmalloc(q, new_count);
    if (keep && p && old_count)
        HB_CUDA(cudaMemcpy(q, p, old_count * sizeof(T), cudaMemcpyDeviceToDevice));
    dfree(p);
    p = q;
}

void grow_store(hb_cache *c, size_t new_cap) {
    CacheView &v = c->view;
    const size_t old_cap = v.capacity;
    HB_CHECK(new_cap > old_cap && new_cap < (1ull << 31), "row store cannot grow that far");
    sync_all(c);
    CacheRegs regs;
    HB_CUDA(cudaMemcpy(&regs, v.regs, sizeof(regs), cudaMemcpyDeviceToHost));
    regrow(v.slot_key, old_cap, new_cap, true);
    regrow(v.slot_version, old_cap, new_cap, true);
    regrow(v.slot_updates, old_cap, new_cap, true);
    regrow(v.slot_prio, old_cap, new_cap, true);
    regrow(v.slot_use, old_cap, new_cap, true);
    regrow(v.slot_state, old_cap, new_cap, true);
    regrow(v.slot_flags, old_cap, new_cap, true);
    regrow(v.data, old_cap * c->width, new_cap * c->width, true);
    regrow(v.grad, old_cap * c->width, new_cap * c->width, true);
    regrow(v.pending_list, old_cap, new_cap, true);
    regrow(v.victims, old_cap, new_cap, false);
    regrow(v.cand_prio, 2 * old_cap, 2 * new_cap, false);
    regrow(v.cand_slot, 2 * old_cap, 2 * new_cap, false);
    u32 *old_stack = v.free_stack;
    v.free_stack = nullptr;
    dmalloc(v.free_stack, new_cap);
    v.capacity = (u32)new_cap;
    HB_LAUNCH(grow_slots_kernel, lin_grid(new_cap), 256, 0, c->stream, v, (u32)old_cap, old_stack, regs.free_top);
    HB_LAUNCHED();
    HB_CUDA(cudaStreamSynchronize(c->stream));
    dfree(old_stack);
    c->slack = new_cap - c->limit;
}

// The calls enqueued so far may hold up to `pending_upper` dirty victims in slots (evict_, flushed
// by the next pushing call) and the new call needs up to `n` fresh lines: make sure the store's
// slack covers both.  The reference's transient lines and evict_ vector are heap-backed and
// unbounded (cache.cc:140-166); here the store grows instead (a synchronising reallocation, which
// hb_cache_reserve avoids by sizing the slack up front).
void ensure_slack(hb_cache *c, size_t n) {
    if (c->pending_upper + n <= c->slack)
        return;
    // tighten the bound from the newest finished call's record: pending then + keys looked up since
    for (uint64_t back = 1; back <= std::min<uint64_t>(c->calls, 64); back++) {
        const uint64_t call = c->calls - back;
        const int idx = (int)(call % hb_cache::kRing);
        if (cudaEventQuery(c->ev_end[idx]) != cudaSuccess)
            continue;
        size_t bound = c->ring[idx].pending;
        for (uint64_t k = call + 1; k < c->calls; k++)
            bound += c->incoming_ring[k % hb_cache::kRing];
        c->pending_upper = std::min(c->pending_upper, bound);
        break;
    }
    (void)cudaGetLastError();
    if (c->pending_upper + n <= c->slack)
        return;
    sync_all(c);
    if (c->calls)
        c->pending_upper = c->ring[(c->calls - 1) % hb_cache::kRing].pending;
    if (c->pending_upper + n <= c->slack)
        return;
    grow_store(c, c->limit + 2 * (c->pending_upper + n));
}

// phase boundary k of the running call (only when perf is enabled: cache.cc:89-106 timings)
void mark(hb_cache *c, int k) {
    // an event between two kernels costs their launch overlap (PDL): with sampling, only every
    // perf_every-th update/lookup pair carries the phase events (the counters are always recorded)
    if (!c->perf_phases || (c->perf_every > 1 && (c->calls / 2) % c->perf_every != 0))
        return;
    int idx = (int)(c->calls % hb_cache::kRing);
    HB_CUDA(cudaEventRecord(c->ev_phase[idx * hb_cache::kPhases + k], c->stream));
    c->phase_mask[idx] |= 1u << k;
}

// Side stream: sort + unique of one key batch into workspace `wsi`.  `check`: the batch may be the
// one the workspace already holds sorted (Hetu's BSP loop updates the batch it looked up one call
// earlier, ParameterServerCommunicate.py:48-52) — then an exact device-side comparison runs first
// and the sort kernels return at once when it finds no difference.
void presort(hb_cache *c, const void *dev_keys, int kind, size_t n, int wsi, bool check) {
    KeyWorkspace &ws = c->ws[wsi];
    cudaStream_t sd = c->side;
    c->cur_ticks += n;
    ws.reset_side(sd);
    const u32 *same = check ? check_same_keys(ws, dev_keys, kind, n, sd) : nullptr;
    // the kernels below rewrite the workspace: its main-stream readers enqueued so far must be done
    // (the comparison above only reads it, and what it reads was written on this stream)
    if (c->ev_ws_free[wsi])
        HB_CUDA(cudaStreamWaitEvent(sd, c->ev_ws_free[wsi], 0));
    SortedKeys sk{nullptr, nullptr};
    if (n)
        sk = radix_sort_keys(ws, dev_keys, kind, n, c->key_bits, sd, same);
    c->sorted[wsi] = sk;
    unique_from_sorted(ws, sk, n, sd, same);
    ws.sorted_valid = n > 0;
    ws.sorted_n = n;
    HB_CUDA(cudaEventRecord(c->ev_sorted[wsi], sd));
}

// Call boundary on the main stream.  Everything that is not a kernel (memsets, event records and
// waits) is gathered here: between two kernels it would cost their launch overlap (PDL).
// ws_a / ws_b: workspaces the call's kernels read (their presort must have finished); -1 = none.
void begin_call(hb_cache *c, bool flush, size_t n, int ws_a, int ws_b = -1) {
    int idx = (int)(c->calls % hb_cache::kRing);
    c->phase_mask[idx] = 0;
    c->dl_of_call[idx] = 0;
    ZeroList z;
    z.n = 0;
    for (int w : {ws_a, ws_b})
        if (w >= 0) {
            z.n += c->ws[w].main_ranges(n, 2, z.r + z.n);
            HB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_sorted[w], 0));
        }
    // the call's duration is only measured with perf enabled (python_api.cc:40-41), and with perf
    // sampling only on the sampled calls
    c->timed[idx] = c->perf_phases && !(c->perf_every > 1 && (c->calls / 2) % c->perf_every != 0);
    if (c->timed[idx])
        HB_CUDA(cudaEventRecord(c->ev_begin[idx], c->stream));
    HB_LAUNCH(op_begin_kernel, 1, 256, 0, c->stream, c->view.regs, clk_of(c), flush ? 1 : 0, z);
    HB_LAUNCHED();
}

// the main-stream readers of workspace `wsi` enqueued so far end here
void release_ws(hb_cache *c, int wsi) {
    HB_CUDA(cudaEventRecord(c->ev_ws_rel[wsi], c->stream));
    c->ev_ws_free[wsi] = c->ev_ws_rel[wsi];
}
// ... the same when the call has just ended: its end event stands for the release
void release_ws_at_end(hb_cache *c, int wsi) {
    c->ev_ws_free[wsi] = c->ev_last_end;
}

void end_call(hb_cache *c, int last_stage, u32 kind, size_t n, bool inserted, bool epilogue_done = false) {
    int idx = (int)(c->calls % hb_cache::kRing);
    if (c->x_pending == 2) { // the exchange forked by the PREVIOUS call: long finished, joined for the record
        HB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_x_done, 0));
        c->x_pending = 0;
    } else if (c->x_pending == 1) {
        c->x_pending = 2;
    }
    // the record goes straight into the host's ring (mapped pinned memory): no copy node
    if (!epilogue_done) { // (a single-GPU update's tail kernel has already written it)
        HB_LAUNCH(op_end_kernel, 1, 1, 0, c->stream, c->view, clk_of(c), last_stage, c->ring_dev + idx, kind,
                  (u32)n, inserted ? 1 : 0);
        HB_LAUNCHED();
    }
    HB_CUDA(cudaEventRecord(c->ev_end[idx], c->stream));
    c->ev_last_end = c->ev_end[idx];
    c->ticks_ring[idx] = c->cur_ticks + 4; // + the single-line paths (reinsert touch)
    c->cur_ticks = 0;
    c->calls++;
    c->incoming_ring[c->calls % hb_cache::kRing] = 0;
}

// resolve (+ alloc) of the batch workspace `wsi` holds; `batch` = which counter set of the call
template <class Plan>
void resolve_batch(hb_cache *c, size_t n, int batch, int wsi, bool dataless, int clk_stage, bool marks,
                   Plan plan, bool plan_insert = false) {
    KeyWorkspace &ws = c->ws[wsi];
    cudaStream_t st = c->stream;
    if (marks)
        mark(c, 0);
    // (a planning resolve also writes the padding items behind the last unique: up to ticket_rows more)
    const size_t cover = Plan::enabled ? n + ticket_rows() : n;
    u32 ntiles = (u32)std::max(1, ceil_div(cover, kScanBlock * kResolveItems));
    u64 *clk = clk_of(c);
    HB_LAUNCH(resolve_kernel<Plan>, ntiles, kScanBlock, 0, st, c->view, ws.uniq, ws.num_unique, c->uslot[batch],
                                                  c->miss_list[batch], c->bypass ? 1 : 0,
                                                  ws.next_scan(), ntiles, clk + clk_stage,
                                                  clk + clk_stage + 1, batch, dataless ? 1 : 0,
                                                  plan_insert ? 1 : 0, plan);
    HB_LAUNCHED();
    if (marks)
        mark(c, 1);
}
// plan_insert: the resolve's last tile also writes the plan of the insert that follows (run_insert
// with planned = true)
void resolve_batch(hb_cache *c, size_t n, int batch, int wsi, bool dataless, int clk_stage,
                   bool marks = true, bool plan_insert = false) {
    resolve_batch(c, n, batch, wsi, dataless, clk_stage, marks, NoSegPlan{}, plan_insert);
}

bool vec4(const hb_cache *c, const void *user_rows) {
    return c->width % 4 == 0 && reinterpret_cast<uintptr_t>(user_rows) % 16 == 0;
}

void run_sync(hb_cache *c, size_t n, int wsi) {
    if (!n)
        return;
    int grid = row_grid((n + 31) / 32);
    // bulk-copy variant (default for rows of 16-byte multiples): $HERALD_SYNC_BULK = 0 selects the
    // register-staged kernel
    static const int force_bulk = [] {
        const char *e = getenv("HERALD_SYNC_BULK");
        return e ? atoi(e) : -1;
    }();
    static const int bulk_rows_env = [] {
        const char *e = getenv("HERALD_BULK_ROWS");
        return e ? std::max(1, std::min(32, atoi(e))) : 8;
    }();
    // rows per round: as configured, but a CTA's stage stays within 64 KB (D = 512: 4 rows)
    const int bulk_rows = (int)std::max<size_t>(1, std::min<size_t>(bulk_rows_env, (64 * 1024) / (kRowWarps * c->width * sizeof(float))));
    const size_t bulk_smem = (size_t)kRowWarps * bulk_rows * c->width * sizeof(float) + kRowWarps * 8;
    const bool use_bulk = (force_bulk >= 0 ? force_bulk != 0 : true) && c->width % 4 == 0 &&
                          c->width * sizeof(float) <= 8192;
    if (use_bulk) {
        static size_t attr_smem = 0;
        if (bulk_smem > attr_smem) {
            HB_CUDA(cudaFuncSetAttribute(sync_bulk_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem));
            attr_smem = bulk_smem;
        }
        HB_LAUNCH(sync_bulk_kernel<0>, grid, kRowBlock, bulk_smem, c->stream, c->view, c->ws[wsi].uniq, c->uslot[0],
                  c->pull_bound, (u64)c->xepoch, bulk_rows);
    } else if (c->width % 4 == 0)
        HB_LAUNCH((sync_kernel<4, 4>), grid, kRowBlock, 0, c->stream, c->view, c->ws[wsi].uniq, c->uslot[0],
                                                             c->pull_bound, (u64)c->xepoch);
    else
        HB_LAUNCH((sync_kernel<1, 4>), grid, kRowBlock, 0, c->stream, c->view, c->ws[wsi].uniq, c->uslot[0],
                                                             c->pull_bound, (u64)c->xepoch);
    HB_LAUNCHED();
}

void run_gather(hb_cache *c, size_t n, int wsi, float *dev_dest) {
    if (!n)
        return;
    IndexFromSlots idx{c->uslot[0], c->ws[wsi].inverse};
    int grid = row_grid((n + 3) / 4);
    // bulk-copy (TMA) variant for wide rows: measured at D = 512 (2 KB rows) 183 -> 136 us, at D = 128
    // (512 B rows) 52 -> 53 us; $HERALD_GATHER_BULK = 0 / 1 forces either
    static const int force_gather_bulk = [] {
        const char *e = getenv("HERALD_GATHER_BULK");
        return e ? atoi(e) : -1;
    }();
    const bool gather_bulk = force_gather_bulk >= 0 ? force_gather_bulk != 0 : c->width * sizeof(float) >= 1024;
    if (gather_bulk && vec4(c, dev_dest) && c->width * sizeof(float) <= 4096) {
        const int R = (int)std::max<size_t>(1, std::min<size_t>(8, 4096 / (c->width * sizeof(float))));
        const size_t smem = (size_t)kRowWarps * 2 * R * c->width * sizeof(float) + kRowWarps * 2 * 8;
        static size_t attr_smem = 0;
        if (smem > attr_smem) {
            HB_CUDA(cudaFuncSetAttribute(gather_bulk_kernel<IndexFromSlots>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_smem = smem;
        }
        grid = std::min(row_grid((n + 31) / 32), sm_count() * 3);
        HB_LAUNCH(gather_bulk_kernel<IndexFromSlots>, grid, kRowBlock, smem, c->stream, c->view.data, dev_dest, n, c->width, idx, R);
    } else if (vec4(c, dev_dest))
        HB_LAUNCH((gather_rows_kernel<4, 4, IndexFromSlots>), grid, kRowBlock, 0, c->stream, c->view.data, dev_dest, n, c->width, idx);
    else
        HB_LAUNCH((gather_rows_kernel<1, 4, IndexFromSlots>), grid, kRowBlock, 0, c->stream, c->view.data, dev_dest, n, c->width, idx);
    HB_LAUNCHED();
}

// planned: the resolve of this batch has already written the plan (resolve_batch, plan_insert)
void run_insert(hb_cache *c, size_t n, int clk_stage, cudaStream_t st, bool planned = false) {
    u64 *clk = clk_of(c);
    if (!planned) {
        HB_LAUNCH(plan_insert_kernel, 1, 256, 0, st, c->view, c->bypass ? 1 : 0, clk + clk_stage,
                                              clk + clk_stage + 1);
        HB_LAUNCHED();
    }
    if (!n)
        return;
    int sgrid = lin_grid(c->view.capacity);
    c->cur_ticks += n;
    // LRU walks the stamp log (sel_log_kernel) as long as [floor, now) fits the log; the device
    // decides (plan_insert_kernel), the histogram kernels below are its fallback.  They are not
    // even launched when the host can prove the span fits: newest finished call's clock - floor
    // plus an upper bound of the ticks taken since.
    bool log_certain = false;
    if (c->policy == HB_POLICY_LRU) {
        for (uint64_t back = 1; back <= std::min<uint64_t>(c->calls, 64); back++) {
            const uint64_t call = c->calls - back;
            const int idx = (int)(call % hb_cache::kRing);
            if (cudaEventQuery(c->ev_end[idx]) != cudaSuccess)
                continue;
            const PerfRecord &rec = c->ring[idx];
            uint64_t span = rec.clock - rec.floor + c->cur_ticks;
            for (uint64_t k = call + 1; k < c->calls; k++)
                span += c->ticks_ring[k % hb_cache::kRing];
            log_certain = span <= (uint64_t)c->view.log_mask + 1;
            break;
        }
        (void)cudaGetLastError(); // cudaErrorNotReady of the queries is not an error
        HB_LAUNCH(sel_log_kernel, sm_count(), kLogBlock, 0, st, c->view);
        HB_LAUNCHED();
    }
    if (!log_certain) {
        // few fat CTAs: every CTA ends with one global atomic per non-empty bin of its histogram
        const int hgrid = sm_count() * 2;
        HB_LAUNCH(sel_hist_kernel, hgrid, 1024, 0, st, c->view, 0);
        HB_LAUNCHED();
        HB_LAUNCH(sel_hist_kernel, hgrid, 1024, 0, st, c->view, 1);
        HB_LAUNCHED();
        HB_LAUNCH(sel_collect_kernel, sgrid, 256, 0, st, c->view);
        HB_LAUNCHED();
        HB_LAUNCH(sel_refine_kernel, 1, 1024, 0, st, c->view);
        HB_LAUNCHED();
    }
    if (c->policy != HB_POLICY_LRU) {
        HB_LAUNCH(min_use_kernel, sgrid, 256, 0, st, c->view);
        HB_LAUNCHED();
        HB_LAUNCH(min_prio_kernel, sgrid, 256, 0, st, c->view);
        HB_LAUNCHED();
        HB_LAUNCH(min_pick_kernel, sgrid, 256, 0, st, c->view);
        HB_LAUNCHED();
    }
    const u32 half = (u32)std::max(1, lin_grid(n) / 2);
    HB_LAUNCH(evict_insert_kernel, 2 * half, 256, 0, st, c->view, c->uslot[0], c->miss_list[0], half);
    HB_LAUNCHED();
}

// Multi-GPU: the pushes of this call sit in the owners' mailboxes.  Three launches, no kernel of
// their own for the barriers: arrive (counts + `ready` flags) -> link (waits for every source's
// `ready`, chains the pushed entries per row) -> apply (one ordered read-modify-write per row; its
// last CTA raises `applied` at every peer).  The next kernel that touches peer memory — the sync
// of the following lookup, or the next update's owner_bounds — waits for `applied`, so the skew
// between the ranks is absorbed by whatever runs in between.  Nothing synchronises with the host.
// The three launches run on their own stream (`xstream`), forked after the kernels that filled the
// mailboxes: nothing on the main stream reads what they write before it has seen the `applied` flags
// (device side), so the update's tail and the next lookup's resolve proceed next to the apply; the
// main stream joins it at the end of the NEXT call (by then the sync kernel has long waited for it),
// which keeps every exchange inside the main stream's timeline one call later.
void exchange_pushes(hb_cache *c) {
    HB_CUDA(cudaEventRecord(c->ev_x_fork, c->stream));
    HB_CUDA(cudaStreamWaitEvent(c->xstream, c->ev_x_fork, 0));
    cudaStream_t st = c->xstream;
    const int world = c->view.pv.world;
    const u64 epoch = ++c->xepoch;
    HB_LAUNCH(exchange_arrive_kernel, 1, 32, 0, st, c->view, epoch);
    HB_LAUNCHED();
    const size_t per_list = std::max<size_t>(c->xrows_upper, 1);
    const int lgrid = (int)std::max<size_t>(1, std::min<size_t>((per_list + 255) / 256, (size_t)sm_count()));
    HB_LAUNCH(link_mailbox_kernel, dim3(lgrid, 2 * world), 256, 0, st, c->view, epoch);
    HB_LAUNCHED();
    const size_t groups = (per_list + 31) / 32;
    // full width: the apply is on the critical path of the next lookup's sync (it waits for the
    // `applied` flags); a small low-priority grid that "leaves room" was measured: 0.352 -> 0.418 ms
    const int agrid = (int)std::max<size_t>(1, std::min<size_t>((groups + kRowWarps - 1) / kRowWarps,
                                                               (size_t)sm_count() * 8 / (2 * world) + 1));
    if (c->width % 4 == 0)
        HB_LAUNCH((apply_linked_kernel<4, 4>), dim3(agrid, 2 * world), kRowBlock, 0, st, c->view, epoch);
    else
        HB_LAUNCH((apply_linked_kernel<1, 4>), dim3(agrid, 2 * world), kRowBlock, 0, st, c->view, epoch);
    HB_LAUNCHED();
    HB_CUDA(cudaEventRecord(c->ev_x_done, st));
    c->x_pending = 1; // joined by the main stream at the end of the next call (end_call)
}

// accumulate + push of batch `batch`, then flush of pending victims, then drop dataless lines
// the accumulate functor of one update (any of its compile-time variants: same fields)
template <int VEC, bool SC, bool SP>
AccumulatePush<VEC, SC, SP> make_accumulate(hb_cache *c, int batch, int wsi, const u64 *dev_push_keys, size_t n_push,
                                            bool use_plan, bool defer_cleanup) {
    KeyWorkspace &ws = c->ws[wsi];
    const u64 *plan = use_plan ? dev_push_keys : nullptr;
    u32 plan_n = (u32)n_push;
    if (use_plan && !dev_push_keys) { // empty plan: nothing is pushed
        plan = ws.uniq;
        plan_n = 0;
    }
    return AccumulatePush<VEC, SC, SP>{c->view, ws.uniq, c->uslot[batch], c->push_bound, plan, plan_n,
                                       defer_cleanup, c->grad_scale};
}

void run_owner_bounds(hb_cache *c, int wsi) {
    if (c->view.pv.world > 1) {
        KeyWorkspace &ws = c->ws[wsi];
        HB_LAUNCH(owner_bounds_kernel, 1, 32, 0, c->stream, c->view, ws.uniq, ws.num_unique, (u64)c->xepoch);
        HB_LAUNCHED();
    }
}

// pre: the batch's work items were planned inside its resolve (do_update) with this setup
void run_accumulate(hb_cache *c, size_t n, int batch, int wsi, const float *dev_grads,
                    const u64 *dev_push_keys, size_t n_push, bool use_plan, bool defer_cleanup = false,
                    bool fuse_tail = false, const SegSetup *pre = nullptr) {
    cudaStream_t st = c->stream;
    KeyWorkspace &ws = c->ws[wsi];
    if (!pre)
        run_owner_bounds(c, wsi);
    if (n) {
        const u32 *p = c->sorted[wsi].perm;
        auto go = [&](auto scaled, auto split) {
            constexpr bool SC = decltype(scaled)::value, SP = decltype(split)::value;
            auto f1 = make_accumulate<1, SC, SP>(c, batch, wsi, dev_push_keys, n_push, use_plan, defer_cleanup);
            auto f4 = make_accumulate<4, SC, SP>(c, batch, wsi, dev_push_keys, n_push, use_plan, defer_cleanup);
            run_segment_reduce(ws, p, dev_grads, c->width, n, vec4(c, dev_grads), c->hot_threshold, st,
                               f1, f4, [&] {
                                   if (batch == 0)
                                       mark(c, 3);
                               }, SP, pre);
        };
        const bool scaled = c->grad_scale != 1.0f; // x * 1.0f == x: the unscaled kernel is exact for it
        if (scaled && c->reduce_split)
            go(std::true_type{}, std::true_type{});
        else if (scaled)
            go(std::true_type{}, std::false_type{});
        else if (c->reduce_split)
            go(std::false_type{}, std::true_type{});
        else
            go(std::false_type{}, std::false_type{});
    }
    if (batch == 0)
        mark(c, 2);
    // pending victims (count is device-side; bound the grid with the host's upper bound)
    size_t pend = std::min<size_t>(c->pending_upper, c->view.capacity);
    if (c->view.pv.world > 1) {
        if (pend) {
            int grid = row_grid(pend);
            if (c->width % 4 == 0)
                HB_LAUNCH(flush_remote_kernel<4>, grid, kRowBlock, 0, st, c->view);
            else
                HB_LAUNCH(flush_remote_kernel<1>, grid, kRowBlock, 0, st, c->view);
            HB_LAUNCHED();
        }
        exchange_pushes(c);
    } else if (fuse_tail) {
        // flush + release of the dataless lines + the call's epilogue in one launch
        const int idx = (int)(c->calls % hb_cache::kRing);
        const int grid = std::max(row_grid(c->may_hold_dirty ? pend : 0), std::max(1, lin_grid(n) / 4));
        if (c->width % 4 == 0)
            HB_LAUNCH(update_tail_kernel<4>, grid, kRowBlock, 0, st, c->view, c->uslot[batch], c->miss_list[batch],
                      clk_of(c), 1, c->ring_dev + idx, (u32)n);
        else
            HB_LAUNCH(update_tail_kernel<1>, grid, kRowBlock, 0, st, c->view, c->uslot[batch], c->miss_list[batch],
                      clk_of(c), 1, c->ring_dev + idx, (u32)n);
        HB_LAUNCHED();
        c->pending_upper = 0;
        return;
    } else if (pend) {
        int grid = row_grid(pend);
        if (c->width % 4 == 0) {
            FlushPending<4> f{c->view, 0};
            HB_LAUNCH((foreach_row_kernel<4, FlushPending<4>>), grid, kRowBlock, 0, st, 0, &c->view.regs->flushed, c->width, f);
        } else {
            FlushPending<1> f{c->view, 0};
            HB_LAUNCH((foreach_row_kernel<1, FlushPending<1>>), grid, kRowBlock, 0, st, 0, &c->view.regs->flushed, c->width, f);
        }
        HB_LAUNCHED();
    }
    c->pending_upper = 0;
    HB_LAUNCH(free_transient_kernel, lin_grid(n), 256, 0, st, c->view, c->uslot[batch], c->miss_list[batch],
                                                       batch, 1);
    HB_LAUNCHED();
}

const u64 *stage_push_keys(hb_cache *c, const void *push_keys, int kind, size_t n_push) {
    if (!n_push)
        return nullptr;
    cudaStream_t st = c->stream;
    size_t need = n_push * 8 * 2;
    if (need > c->push_keys_stage_cap) {
        HB_CUDA(cudaStreamSynchronize(st));
        if (c->push_keys_stage)
            cudaFree(c->push_keys_stage);
        HB_CUDA(cudaMalloc(&c->push_keys_stage, need));
        c->push_keys_stage_cap = need;
    }
    u64 *out = reinterpret_cast<u64 *>(c->push_keys_stage);
    char *raw = reinterpret_cast<char *>(c->push_keys_stage) + n_push * 8;
    const void *dev = push_keys;
    if (!is_device_ptr(push_keys)) {
        HB_CUDA(cudaMemcpyAsync(raw, push_keys, n_push * (kind == HB_KEYS_F32 ? 4 : 8),
                                cudaMemcpyHostToDevice, st));
        dev = raw;
    }
    if (kind == HB_KEYS_F32) {
        HB_LAUNCH(convert_keys_kernel, ceil_div(n_push, 256), 256, 0, st, (const float *)dev, out, n_push);
        HB_LAUNCHED();
        return out;
    }
    return reinterpret_cast<const u64 *>(dev);
}

// gradients as given by the caller -> device pointer.  Host gradients travel on the upload stream
// (behind the previous consumer of the staging buffer); the main stream waits for them at the call
// boundary (begin_call follows).
const float *stage_grads(hb_cache *c, const float *grads, size_t n) {
    if (!n || is_device_ptr(grads))
        return grads;
    float *stage = rows_stage(c, n, 2);
    HB_CUDA(cudaStreamWaitEvent(c->h2d, c->ev_grads_free, 0));
    HB_CUDA(cudaMemcpyAsync(stage, grads, n * c->width * sizeof(float), cudaMemcpyHostToDevice, c->h2d));
    HB_CUDA(cudaEventRecord(c->ev_up, c->h2d));
    HB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_up, 0));
    return stage;
}

void do_update(hb_cache *c, const void *keys, int kind, size_t n, const float *grads,
               const void *push_keys, int push_kind, size_t n_push, bool use_plan) {
    Guard g(c->device);
    ensure_batch(c, n);
    ensure_keys_stage(c, n);
    ensure_slack(c, n);
    if (use_plan || c->push_bound > 0)
        c->may_hold_dirty = true; // lines outside the plan / under the bound keep their gradient
    const int w = c->cur; // the workspace of the most recent lookup: usually this very batch
    const void *dkeys = stage_keys(c, keys, kind, n, 0);
    presort(c, dkeys, kind, n, w, /*check=*/true);
    const float *dgrads = stage_grads(c, grads, n);
    const u64 *dpush = use_plan ? stage_push_keys(c, push_keys, push_kind, n_push) : nullptr;
    begin_call(c, /*flush=*/true, n, w);
    const bool fuse_tail = c->view.pv.world == 1;
    if (n) {
        // the work items of the segment reduce are planned inside the resolve (multi-GPU: the owner
        // ranges of the sorted uniques first, the plan writes mailbox headers)
        run_owner_bounds(c, w);
        const SegSetup su = seg_setup(c->ws[w], c->width, n, c->hot_threshold, c->reduce_split, c->stream);
        SegPlanArgs<AccumulatePush<4>> plan{c->ws[w].seg_start, c->sorted[w].perm, su.thr, su.hl,
                                            make_accumulate<4, false, false>(c, 0, w, dpush, n_push, use_plan, false)};
        resolve_batch(c, n, 0, w, /*dataless=*/true, 0, true, plan);
        run_accumulate(c, n, 0, w, dgrads, dpush, n_push, use_plan, false, fuse_tail, &su);
    } else {
        resolve_batch(c, n, 0, w, /*dataless=*/true, 0);
        run_accumulate(c, n, 0, w, dgrads, dpush, n_push, use_plan, false, fuse_tail);
    }
    end_call(c, 1, 1, n, false, fuse_tail);
    release_ws_at_end(c, w);
    if (dgrads != grads)
        HB_CUDA(cudaEventRecord(c->ev_grads_free, c->stream));
}

} // namespace
} // namespace hb

using namespace hb;

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

int hb_table_create(int node_id, size_t length, size_t width, int device, hb_table **out) {
    HB_API_BEGIN();
    HB_CHECK(length > 0 && width > 0, "empty table");
    std::lock_guard<std::mutex> lock(g_tables_mtx);
    HB_CHECK(!g_tables.count(node_id), "table id already registered");
    Guard g(device);
    auto *t = new hb_table();
    t->node_id = node_id;
    t->device = device;
    t->length = length;
    t->width = width;
    int rank = 0, world = 1;
    hb_comm_rank(&rank, &world);
    // AveragePartitioner: len/S rows each, the first len%S shards one more (partitioner.h:46-57)
    size_t per = length / world, rem = length % world;
    t->row_begin = (size_t)rank * per + std::min<size_t>(rank, rem);
    t->nrows = per + ((size_t)rank < rem ? 1 : 0);
    HB_CHECK(t->nrows < (1ull << 32), "a shard holds fewer than 2^32 rows");
    dmalloc_shared(t->rows, t->nrows * width);
    dmalloc_shared(t->ver, t->nrows);
    HB_CUDA(cudaMemset(t->rows, 0, std::max<size_t>(t->nrows * width, 1) * sizeof(float)));
    HB_CUDA(cudaMemset(t->ver, 0, std::max<size_t>(t->nrows, 1) * sizeof(i64)));
    t->rank = rank;
    t->world = world;
    ipc_share(t->rows, reinterpret_cast<void **>(t->peer_rows));
    ipc_share(t->ver, reinterpret_cast<void **>(t->peer_ver));
    g_tables[node_id] = t;
    if (out)
        *out = t;
    HB_API_END();
}

int hb_table_get(int node_id, hb_table **out) {
    HB_API_BEGIN();
    std::lock_guard<std::mutex> lock(g_tables_mtx);
    auto it = g_tables.find(node_id);
    HB_CHECK(it != g_tables.end(), "no table with this node_id (InitTensor first)");
    *out = it->second;
    HB_API_END();
}

int hb_table_destroy(hb_table *t) {
    HB_API_BEGIN();
    if (t) {
        std::lock_guard<std::mutex> lock(g_tables_mtx);
        g_tables.erase(t->node_id);
        Guard g(t->device);
        HB_CUDA(cudaDeviceSynchronize());
        if (t->world > 1) {
            hb_comm_barrier(); // every rank is done with every shard
            ipc_unshare(reinterpret_cast<void **>(t->peer_rows));
            ipc_unshare(reinterpret_cast<void **>(t->peer_ver));
        }
        dfree(t->rows);
        dfree(t->ver);
        delete t;
    }
    HB_API_END();
}

int hb_table_init(hb_table *t, int init_type, double a, double b, unsigned long long seed) {
    HB_API_BEGIN();
    HB_CHECK(init_type >= 0 && init_type <= 3, "unknown init_type");
    Guard g(t->device);
    size_t nelem = t->nrows * t->width;
    if (nelem) {
        HB_LAUNCH(table_init_kernel, lin_grid(nelem), 256, 0, 0, t->rows, nelem, t->row_begin * t->width, init_type,
                                                    (float)a, (float)b, splitmix64(seed));
        HB_LAUNCHED();
        HB_CUDA(cudaDeviceSynchronize());
    }
    HB_API_END();
}

static void clip_range(const hb_table *t, size_t row_begin, size_t nrows, size_t &lo, size_t &hi) {
    lo = std::max(row_begin, t->row_begin);
    hi = std::min(row_begin + nrows, t->row_begin + t->nrows);
}

int hb_table_load_rows(hb_table *t, size_t row_begin, size_t nrows, const float *rows) {
    HB_API_BEGIN();
    HB_CHECK(row_begin + nrows <= t->length, "row range outside the table");
    Guard g(t->device);
    size_t lo, hi;
    clip_range(t, row_begin, nrows, lo, hi);
    if (lo < hi)
        HB_CUDA(cudaMemcpy(t->rows + (lo - t->row_begin) * t->width, rows + (lo - row_begin) * t->width,
                           (hi - lo) * t->width * sizeof(float), cudaMemcpyDefault));
    HB_API_END();
}

int hb_table_read_rows(hb_table *t, size_t row_begin, size_t nrows, float *rows) {
    HB_API_BEGIN();
    HB_CHECK(row_begin + nrows <= t->length, "row range outside the table");
    Guard g(t->device);
    HB_CUDA(cudaDeviceSynchronize());
    size_t lo, hi;
    clip_range(t, row_begin, nrows, lo, hi);
    if (lo < hi)
        HB_CUDA(cudaMemcpy(rows + (lo - row_begin) * t->width, t->rows + (lo - t->row_begin) * t->width,
                           (hi - lo) * t->width * sizeof(float), cudaMemcpyDefault));
    HB_API_END();
}

int hb_table_read_versions(hb_table *t, size_t row_begin, size_t nrows, int64_t *versions) {
    HB_API_BEGIN();
    HB_CHECK(row_begin + nrows <= t->length, "row range outside the table");
    Guard g(t->device);
    HB_CUDA(cudaDeviceSynchronize());
    size_t lo, hi;
    clip_range(t, row_begin, nrows, lo, hi);
    if (lo < hi)
        HB_CUDA(cudaMemcpy(versions + (lo - row_begin), t->ver + (lo - t->row_begin),
                           (hi - lo) * sizeof(i64), cudaMemcpyDefault));
    HB_API_END();
}

int hb_table_read_rows_at(hb_table *t, const uint64_t *keys, size_t n, float *rows, int64_t *versions) {
    HB_API_BEGIN();
    Guard g(t->device);
    HB_CUDA(cudaDeviceSynchronize());
    if (n) {
        u64 *dkeys = nullptr;
        float *drows = nullptr;
        i64 *dver = nullptr;
        dmalloc(dkeys, n);
        HB_CUDA(cudaMemcpy(dkeys, keys, n * sizeof(u64), cudaMemcpyDefault));
        if (rows) {
            dmalloc(drows, n * t->width);
            IndexFromShardKeys idx{dkeys, t->row_begin, t->nrows};
            int grid = row_grid((n + 3) / 4);
            if (t->width % 4 == 0)
                HB_LAUNCH((gather_rows_kernel<4, 4, IndexFromShardKeys>), grid, kRowBlock, 0, 0, t->rows, drows, n, t->width, idx);
            else
                HB_LAUNCH((gather_rows_kernel<1, 4, IndexFromShardKeys>), grid, kRowBlock, 0, 0, t->rows, drows, n, t->width, idx);
            HB_LAUNCHED();
            HB_CUDA(cudaMemcpy(rows, drows, n * t->width * sizeof(float), cudaMemcpyDefault));
        }
        if (versions) {
            dmalloc(dver, n);
            HB_LAUNCH(gather_versions_kernel, lin_grid(n), 256, 0, 0, t->ver, dkeys, n, (u64)t->row_begin,
                      (u64)t->nrows, dver);
            HB_LAUNCHED();
            HB_CUDA(cudaMemcpy(versions, dver, n * sizeof(i64), cudaMemcpyDefault));
        }
        dfree(dkeys);
        dfree(drows);
        dfree(dver);
    }
    HB_API_END();
}

int hb_table_shard(hb_table *t, size_t *row_begin, size_t *nrows, float **dev_rows,
                   int64_t **dev_versions) {
    HB_API_BEGIN();
    if (row_begin)
        *row_begin = t->row_begin;
    if (nrows)
        *nrows = t->nrows;
    if (dev_rows)
        *dev_rows = t->rows;
    if (dev_versions)
        *dev_versions = reinterpret_cast<int64_t *>(t->ver);
    HB_API_END();
}

// ---------------------------------------------------------------------------------------
int hb_cache_create(int policy, size_t limit, size_t length, size_t width, int node_id,
                    hb_cache **out) {
    HB_API_BEGIN();
    HB_CHECK(policy >= HB_POLICY_LRU && policy <= HB_POLICY_LFUOPT, "unknown policy");
    HB_CHECK(limit < (1ull << 31), "limit too large");
    hb_table *t = nullptr;
    HB_CHECK(hb_table_get(node_id, &t) == 0, "no table with this node_id (InitTensor first)");
    HB_CHECK(t->width == width, "cache width differs from the table's");
    Guard g(t->device);
    auto *c = new hb_cache();
    c->policy = policy;
    c->limit = limit;
    c->length = length;
    c->width = width;
    c->node_id = node_id;
    c->device = t->device;
    c->table = t;
    c->key_bits = bits_for(std::max<size_t>(length, t->length));
    c->hot_threshold = default_hot_threshold();
    if (const char *e = getenv("HERALD_REDUCE")) // process-wide default of hb_cache_set_reduce_mode
        c->reduce_split = std::string(e) == "split" || std::string(e) == "1";
    HB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    HB_CUDA(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    HB_CUDA(cudaStreamCreateWithFlags(&c->side2, cudaStreamNonBlocking));
    HB_CUDA(cudaStreamCreateWithFlags(&c->xstream, cudaStreamNonBlocking));
    HB_CUDA(cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking));
    HB_CUDA(cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking));
    for (cudaEvent_t *e : {&c->ev_ws_rel[0], &c->ev_ws_rel[1], &c->ev_sorted[0], &c->ev_sorted[1],
                           &c->ev_up, &c->ev_grads_free, &c->ev_gathered[0], &c->ev_gathered[1],
                           &c->ev_dl[0], &c->ev_dl[1], &c->ev_producer, &c->ev_fork, &c->ev_join,
                           &c->ev_x_fork, &c->ev_x_done})
        HB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    // row store: limit resident lines + slack for the running call's fresh lines and for dirty
    // victims waiting for the next push
    c->slack = std::max<size_t>(1 << 16, std::min<size_t>(limit, 1 << 22));
    size_t cap = limit + c->slack;
    CacheView &v = c->view;
    v.capacity = (u32)cap;
    v.limit = (u32)limit;
    v.width = (u32)width;
    v.policy = policy;
    size_t hs = 1024;
    // >= 4 entries per resident line and a rebuild when half of the entries are used (lines +
    // tombstones): an unsuccessful probe (every miss of a batch) walks (1 + 1/(1-a)^2)/2 entries,
    // 2.5 at a = 1/2 against 8.5 at 3/4; 16 B per entry is cheap next to the rows
    while (hs < 4 * std::max<size_t>(limit, 1))
        hs <<= 1;
    c->ht_size = hs;
    v.ht_mask = (u32)(hs - 1);
    dmalloc(v.slot_key, cap);
    dmalloc(v.slot_version, cap);
    dmalloc(v.slot_updates, cap);
    dmalloc(v.slot_prio, cap);
    dmalloc(v.slot_use, cap);
    dmalloc(v.slot_state, cap);
    dmalloc(v.slot_flags, cap);
    dmalloc(v.data, cap * width);
    dmalloc(v.grad, cap * width);
    dmalloc(v.ht, hs);
    dmalloc(v.free_stack, cap);
    dmalloc(v.pending_list, cap);
    dmalloc(v.victims, cap);
    dmalloc(v.cand_prio, 2 * cap);
    dmalloc(v.cand_slot, 2 * cap);
    dmalloc(v.sel_hist, 2 * kSelBins);
    {   // stamp log: >= 4 entries per slot (the span of live stamps of an LRU cache in steady state
        // is a few ticks per resident line)
        size_t ls = 1 << 16;
        while (ls < 4 * cap)
            ls <<= 1;
        if (const char *e = getenv("HERALD_STAMP_LOG_SLOTS")) { // tests: a tiny log forces the fallback
            ls = 16;
            while (ls < (size_t)std::max(atoi(e), 16))
                ls <<= 1;
        }
        v.log_mask = (u32)(ls - 1);
        dmalloc(v.stamp_log, ls);
        HB_CUDA(cudaMemset(v.stamp_log, 0xff, ls * sizeof(u32)));
        dmalloc(v.sel_scan, ls / kLogTile + 4);
    }
    dmalloc(v.regs, 1);
    v.trows = t->rows;
    v.tver = t->ver;
    v.row_begin = t->row_begin;
    v.nrows_local = t->nrows;
    v.table_len = t->length;
    PeerView &pv = v.pv;
    std::memset(&pv, 0, sizeof(pv));
    pv.world = t->world;
    pv.rank = t->rank;
    pv.per = t->length / t->world;
    pv.rem = t->length % t->world;
    for (int r = 0; r < kMaxWorld; r++) {
        pv.rows[r] = t->peer_rows[r];
        pv.ver[r] = t->peer_ver[r];
    }
    if (t->world == 1) {
        pv.rows[0] = t->rows;
        pv.ver[0] = t->ver;
    }
    pv.timeout_ns = 120ull * 1000000000ull;
    if (const char *e = getenv("HERALD_PEER_TIMEOUT_S"))
        pv.timeout_ns = std::max<u64>(1, strtoull(e, nullptr, 10)) * 1000000000ull;
    dmalloc(pv.lo, kMaxWorld + 1);
    dmalloc(pv.fl_count, kMaxWorld);
    HB_CUDA(cudaMemset(pv.lo, 0, (kMaxWorld + 1) * sizeof(u32)));
    HB_CUDA(cudaMemset(pv.fl_count, 0, kMaxWorld * sizeof(u32)));
    if (t->world > 1) {
        size_t rows = 1 << 18;
        if (const char *e = getenv("HERALD_MAILBOX_ROWS"))
            rows = std::max<size_t>(1024, strtoull(e, nullptr, 10));
        setup_mailbox(c, rows);
    }
    HB_CUDA(cudaMemset(v.ht, 0xff, hs * sizeof(HtEntry)));
    CacheRegs regs;
    std::memset(&regs, 0, sizeof(regs));
    regs.free_top = (u32)cap;
    HB_CUDA(cudaMemcpy(v.regs, &regs, sizeof(regs), cudaMemcpyHostToDevice));
    HB_LAUNCH(init_slots_kernel, lin_grid(cap), 256, 0, 0, v);
    HB_LAUNCHED();
    char *rec = nullptr;
    HB_CUDA(cudaMalloc((void **)&rec, 128));
    HB_CUDA(cudaMemset(rec, 0, 128));
    c->dev_record = reinterpret_cast<PerfRecord *>(rec);
    HB_CUDA(cudaHostAlloc((void **)&c->ring, sizeof(PerfRecord) * hb_cache::kRing,
                          cudaHostAllocMapped | cudaHostAllocPortable));
    std::memset(c->ring, 0, sizeof(PerfRecord) * hb_cache::kRing);
    HB_CUDA(cudaHostGetDevicePointer((void **)&c->ring_dev, c->ring, 0));
    c->ev_begin.resize(hb_cache::kRing);
    c->ev_end.resize(hb_cache::kRing);
    for (int i = 0; i < hb_cache::kRing; i++) {
        HB_CUDA(cudaEventCreate(&c->ev_begin[i]));
        HB_CUDA(cudaEventCreate(&c->ev_end[i]));
    }
    c->ev_phase.resize((size_t)hb_cache::kRing * hb_cache::kPhases);
    for (auto &e : c->ev_phase)
        HB_CUDA(cudaEventCreate(&e));
    c->phase_mask.assign(hb_cache::kRing, 0);
    HB_CUDA(cudaDeviceSynchronize());
    *out = c;
    HB_API_END();
}

int hb_cache_destroy(hb_cache *c) {
    HB_API_BEGIN();
    if (c) {
        Guard g(c->device);
        sync_all(c);
        CacheView &v = c->view;
        dfree(v.slot_key);
        dfree(v.slot_version);
        dfree(v.slot_updates);
        dfree(v.slot_prio);
        dfree(v.slot_use);
        dfree(v.slot_state);
        dfree(v.slot_flags);
        dfree(v.data);
        dfree(v.grad);
        dfree(v.ht);
        dfree(v.free_stack);
        dfree(v.pending_list);
        dfree(v.victims);
        dfree(v.cand_prio);
        dfree(v.cand_slot);
        dfree(v.sel_hist);
        dfree(v.stamp_log);
        dfree(v.sel_scan);
        dfree(v.regs);
        dfree(v.pv.lo);
        dfree(v.pv.fl_count);
        dfree(v.pv.head);
        dfree(v.pv.next);
        if (c->mailbox) {
            hb_comm_barrier(); // nobody still writes into this rank's mailboxes
            ipc_unshare(reinterpret_cast<void **>(c->peer_mailbox));
            cudaFree(c->mailbox);
        }
        for (int b = 0; b < 2; b++) {
            c->ws[b].release();
            dfree(c->uslot[b]);
            dfree(c->miss_list[b]);
            if (c->keys_stage[b])
                cudaFree(c->keys_stage[b]);
        }
        for (int b = 0; b < 3; b++)
            dfree(c->rows_stage[b]);
        if (c->push_keys_stage)
            cudaFree(c->push_keys_stage);
        if (c->score_scratch)
            cudaFree(c->score_scratch);
        cudaFree(c->dev_record);
        cudaFreeHost(c->ring);
        for (auto &e : c->ev_begin)
            cudaEventDestroy(e);
        for (auto &e : c->ev_end)
            cudaEventDestroy(e);
        for (auto &e : c->ev_phase)
            cudaEventDestroy(e);
        for (cudaEvent_t e : {c->ev_ws_rel[0], c->ev_ws_rel[1], c->ev_sorted[0], c->ev_sorted[1],
                              c->ev_up, c->ev_grads_free, c->ev_gathered[0], c->ev_gathered[1],
                              c->ev_dl[0], c->ev_dl[1], c->ev_producer, c->ev_fork, c->ev_join,
                              c->ev_x_fork, c->ev_x_done})
            cudaEventDestroy(e);
        cudaStreamDestroy(c->stream);
        cudaStreamDestroy(c->side);
        cudaStreamDestroy(c->side2);
        cudaStreamDestroy(c->xstream);
        cudaStreamDestroy(c->h2d);
        cudaStreamDestroy(c->d2h);
        delete c;
    }
    HB_API_END();
}

int hb_cache_set_bounds(hb_cache *c, int64_t pull_bound, int64_t push_bound) {
    HB_API_BEGIN();
    c->pull_bound = pull_bound;
    c->push_bound = push_bound;
    HB_API_END();
}

int hb_cache_get_bounds(hb_cache *c, int64_t *pull_bound, int64_t *push_bound) {
    HB_API_BEGIN();
    *pull_bound = c->pull_bound;
    *push_bound = c->push_bound;
    HB_API_END();
}

int hb_cache_set_grad_scale(hb_cache *c, float scale) {
    HB_API_BEGIN();
    c->grad_scale = scale;
    HB_API_END();
}

int hb_cache_set_reduce_mode(hb_cache *c, int mode) {
    HB_API_BEGIN();
    HB_CHECK(mode == 0 || mode == 1, "reduce mode: 0 = occurrence order, 1 = two-level for very hot rows");
    c->reduce_split = mode == 1;
    HB_API_END();
}

int hb_cache_get_reduce_mode(hb_cache *c, int *mode) {
    HB_API_BEGIN();
    *mode = c->reduce_split ? 1 : 0;
    HB_API_END();
}

int hb_cache_get_grad_scale(hb_cache *c, float *scale) {
    HB_API_BEGIN();
    *scale = c->grad_scale;
    HB_API_END();
}

int hb_cache_after_stream(hb_cache *c, void *producer_stream) {
    HB_API_BEGIN();
    Guard g(c->device);
    cudaStream_t ps = (cudaStream_t)producer_stream;
    HB_CUDA(cudaEventRecord(c->ev_producer, ps));
    // keys are read by the sort (side), gradients by the accumulate kernel (main) or staged by a
    // device-to-device-free path: both streams order themselves behind the producer
    HB_CUDA(cudaStreamWaitEvent(c->side, c->ev_producer, 0));
    HB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_producer, 0));
    HB_API_END();
}

int hb_cache_set_bypass(hb_cache *c, int on) {
    HB_API_BEGIN();
    c->bypass = on != 0;
    HB_API_END();
}

int hb_cache_set_perf(hb_cache *c, int on) {
    HB_API_BEGIN();
    c->perf_phases = on != 0;
    HB_API_END();
}

int hb_cache_set_perf_sampling(hb_cache *c, unsigned every) {
    HB_API_BEGIN();
    c->perf_every = every ? every : 1;
    HB_API_END();
}

int hb_cache_reserve(hb_cache *c, size_t max_keys) {
    HB_API_BEGIN();
    Guard g(c->device);
    ensure_batch(c, max_keys);
    // transient lines of one call + dirty victims of a few lookups without a push in between
    if (3 * max_keys > c->slack)
        grow_store(c, c->limit + 3 * max_keys);
    if (c->view.pv.world > 1 && max_keys > c->mailbox_cap) {
        sync_all(c);
        setup_mailbox(c, max_keys); // collective
    }
    HB_API_END();
}

int hb_cache_stream(hb_cache *c, void **stream) {
    HB_API_BEGIN();
    *stream = (void *)c->stream;
    HB_API_END();
}

// dest of a lookup: the caller's device buffer, or one of two device staging buffers from which
// the download stream copies to the caller's host buffer.  *k = staging buffer (-1: none); the main
// stream is made to wait until the previous download out of that buffer has finished.
static float *stage_dest(hb_cache *c, float *dest, size_t n, int *k) {
    *k = -1;
    if (!n || is_device_ptr(dest))
        return dest;
    *k = c->dl_next;
    c->dl_next ^= 1;
    float *d = rows_stage(c, n, *k);
    HB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_dl[*k], 0));
    return d;
}

// after the gather: hand the staged rows to the download stream
static void download_dest(hb_cache *c, float *dest, const float *ddest, size_t n, int k) {
    if (k < 0)
        return;
    HB_CUDA(cudaEventRecord(c->ev_gathered[k], c->stream));
    HB_CUDA(cudaStreamWaitEvent(c->d2h, c->ev_gathered[k], 0));
    HB_CUDA(cudaMemcpyAsync(dest, ddest, n * c->width * sizeof(float), cudaMemcpyDeviceToHost, c->d2h));
    HB_CUDA(cudaEventRecord(c->ev_dl[k], c->d2h));
    c->dl_of_call[c->calls % hb_cache::kRing] = k + 1;
    c->dl_seq[k] = c->calls + 1;
}

int hb_cache_lookup(hb_cache *c, const void *keys, int key_kind, size_t n, float *dest) {
    HB_API_BEGIN();
    Guard g(c->device);
    HB_CHECK(n < (1ull << 31), "too many keys in one call");
    ensure_batch(c, n);
    ensure_keys_stage(c, n);
    ensure_slack(c, n);
    maybe_rebuild_index(c, n);
    // a lookup sorts into the workspace the previous lookup did NOT use: that one usually still
    // serves the update of its batch (same keys, no second sort)
    const int w = c->cur ^ 1;
    const void *dkeys = stage_keys(c, keys, key_kind, n, 0);
    presort(c, dkeys, key_kind, n, w, /*check=*/false);
    c->cur = w;
    int k;
    float *ddest = stage_dest(c, dest, n, &k);
    begin_call(c, false, n, w);
    resolve_batch(c, n, 0, w, /*dataless=*/false, 0, true, /*plan_insert=*/true);
    // The insert phase (victim selection, evictions, index inserts) needs the resolve's results
    // only: it works on slot scalars and the index, never on row data, and its victims are never
    // lines of this batch — so it runs on its own stream NEXT TO sync + gather, which move the rows.
    const bool fork = n >= 4096; // (a small call is launch-bound: keep it on one stream)
    if (fork) {
        HB_CUDA(cudaEventRecord(c->ev_fork, c->stream));
        HB_CUDA(cudaStreamWaitEvent(c->side2, c->ev_fork, 0));
        run_insert(c, n, 1, c->side2, /*planned=*/true);
        HB_CUDA(cudaEventRecord(c->ev_join, c->side2));
    }
    run_sync(c, n, w);
    mark(c, 2);
    run_gather(c, n, w, ddest);
    mark(c, 3);
    release_ws(c, w); // the insert phase works on slots, not on the workspace
    download_dest(c, dest, ddest, n, k);
    if (fork)
        HB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    else
        run_insert(c, n, 1, c->stream, /*planned=*/true);
    c->pending_upper += n;
    end_call(c, 2, 0, n, true);
    HB_API_END();
}

int hb_cache_update(hb_cache *c, const void *keys, int key_kind, size_t n, const float *grads) {
    HB_API_BEGIN();
    HB_CHECK(n < (1ull << 31), "too many keys in one call");
    do_update(c, keys, key_kind, n, grads, nullptr, 0, 0, false);
    HB_API_END();
}

int hb_cache_update_with_push_keys(hb_cache *c, const void *keys, int key_kind, size_t n,
                                   const void *push_keys, int push_key_kind, size_t n_push,
                                   const float *grads) {
    HB_API_BEGIN();
    HB_CHECK(n < (1ull << 31), "too many keys in one call");
    do_update(c, keys, key_kind, n, grads, push_keys, push_key_kind, n_push, true);
    HB_API_END();
}

int hb_cache_push_pull(hb_cache *c, const void *pull_keys, int pull_kind, size_t n_pull, float *dest,
                       const void *push_keys, int push_kind, size_t n_push, const float *grads) {
    HB_API_BEGIN();
    Guard g(c->device);
    HB_CHECK(n_pull < (1ull << 31) && n_push < (1ull << 31), "too many keys in one call");
    ensure_batch(c, std::max(n_pull, n_push));
    ensure_keys_stage(c, std::max(n_pull, n_push));
    ensure_slack(c, n_pull + n_push);
    maybe_rebuild_index(c, n_pull);
    cudaStream_t st = c->stream;
    // the push batch is usually the previous call's pull batch (ASP prefetch,
    // ParameterServerCommunicate.py:36-38): it is checked against that workspace; the pull batch
    // goes to the other one
    c->may_hold_dirty = true; // (deferred cleanup: a pushed line counts as dirty until the call's end)
    const int wpush = c->cur, wpull = c->cur ^ 1;
    const void *dpush = stage_keys(c, push_keys, push_kind, n_push, 1);
    presort(c, dpush, push_kind, n_push, wpush, /*check=*/true);
    const void *dpull = stage_keys(c, pull_keys, pull_kind, n_pull, 0);
    presort(c, dpull, pull_kind, n_pull, wpull, /*check=*/false);
    c->cur = wpull;
    int k;
    float *ddest = stage_dest(c, dest, n_pull, &k);
    const float *dgrads = stage_grads(c, grads, n_push);
    begin_call(c, /*flush=*/true, std::max(n_pull, n_push), wpull, wpush);
    // cache.cc:360-391: pull-side lookup, then push-side lookup + accumulate
    resolve_batch(c, n_pull, 0, wpull, /*dataless=*/false, 0, false);
    resolve_batch(c, n_push, 1, wpush, /*dataless=*/true, 1, false);
    // server order (PSFhandle_embedding.cc:66-79): push first, then sync
    run_accumulate(c, n_push, 1, wpush, dgrads, nullptr, 0, false, /*defer_cleanup=*/true);
    run_sync(c, n_pull, wpull);
    run_gather(c, n_pull, wpull, ddest);
    release_ws(c, wpull);
    download_dest(c, dest, ddest, n_pull, k);
    run_insert(c, n_pull, 2, st);
    if (n_push) {
        HB_LAUNCH(cleanup_pushed_kernel, lin_grid(n_push), 256, 0, st, c->view, c->uslot[1], c->push_bound);
        HB_LAUNCHED();
    }
    c->pending_upper += n_pull;
    end_call(c, 3, 2, n_pull, true);
    if (dgrads != grads)
        HB_CUDA(cudaEventRecord(c->ev_grads_free, c->stream));
    HB_API_END();
}

int hb_cache_flush(hb_cache *c) {
    HB_API_BEGIN();
    // 1. the dirty victims waiting in evict_ (an update of zero keys flushes them; collective)
    do_update(c, nullptr, HB_KEYS_U64, 0, nullptr, nullptr, 0, 0, false);
    // 2. the resident dirty lines, one rank at a time so that the adds on a row happen in rank order
    Guard g(c->device);
    const int world = c->view.pv.world, rank = c->view.pv.rank;
    for (int turn = 0; turn < world; turn++) {
        if (world > 1) {
            sync_all(c); // (the exchange stream too: the resident flush writes owner rows in place)
            HB_CHECK(hb_comm_barrier() == 0, "barrier failed");
        }
        if (turn != rank)
            continue;
        int grid = row_grid((c->view.capacity + 31) / 32);
        if (c->width % 4 == 0)
            HB_LAUNCH(flush_resident_kernel<4>, grid, kRowBlock, 0, c->stream, c->view);
        else
            HB_LAUNCH(flush_resident_kernel<1>, grid, kRowBlock, 0, c->stream, c->view);
        HB_LAUNCHED();
    }
    HB_CUDA(cudaStreamSynchronize(c->stream));
    if (world > 1)
        HB_CHECK(hb_comm_barrier() == 0, "barrier failed");
    HB_API_END();
}

// ---- checkpoint of the owner shard --------------------------------------------------------
// File layout of the reference (ps-lite/include/ps/worker/PSAgent.h:447-476,
// ps-lite/include/ps/server/PSFHandle.h:401-439): "<dir>/<node_id>_<partition>.dat", the
// partition's rows as raw row-major float32.  Partition index = rank (AveragePartitioner order).
// Extension: "<dir>/<node_id>_<partition>.ver" holds the int64 row versions, which the reference
// does not save (a reloaded reference table restarts every version at its in-memory value);
// hb_table_load restores them when the file exists.
static std::string shard_path(const hb_table *t, const char *dir, const char *ext) {
    return std::string(dir) + "/" + std::to_string(t->node_id) + "_" + std::to_string(t->rank) + ext;
}

static void stream_file(void *dev, size_t bytes, const std::string &path, bool save) {
    FILE *f = fopen(path.c_str(), save ? "wb" : "rb");
    if (!f)
        throw Error("cannot open " + path);
    const size_t chunk = (size_t)64 << 20;
    void *host = nullptr;
    if (cudaHostAlloc(&host, chunk, cudaHostAllocDefault) != cudaSuccess) {
        fclose(f);
        throw Error("cannot allocate the pinned staging buffer");
    }
    bool ok = true;
    for (size_t off = 0; off < bytes && ok; off += chunk) {
        const size_t len = std::min(chunk, bytes - off);
        if (save) {
            ok = cudaMemcpy(host, (char *)dev + off, len, cudaMemcpyDeviceToHost) == cudaSuccess &&
                 fwrite(host, 1, len, f) == len;
        } else {
            ok = fread(host, 1, len, f) == len &&
                 cudaMemcpy((char *)dev + off, host, len, cudaMemcpyHostToDevice) == cudaSuccess;
        }
    }
    cudaFreeHost(host);
    ok = (fclose(f) == 0) && ok;
    if (!ok)
        throw Error(std::string(save ? "short write to " : "short read from ") + path);
}

int hb_table_save(hb_table *t, const char *dir) {
    HB_API_BEGIN();
    HB_CHECK(t && dir, "null argument");
    Guard g(t->device);
    HB_CUDA(cudaDeviceSynchronize());
    stream_file(t->rows, t->nrows * t->width * sizeof(float), shard_path(t, dir, ".dat"), true);
    stream_file(t->ver, t->nrows * sizeof(i64), shard_path(t, dir, ".ver"), true);
    HB_API_END();
}

int hb_table_load(hb_table *t, const char *dir) {
    HB_API_BEGIN();
    HB_CHECK(t && dir, "null argument");
    Guard g(t->device);
    HB_CUDA(cudaDeviceSynchronize());
    stream_file(t->rows, t->nrows * t->width * sizeof(float), shard_path(t, dir, ".dat"), false);
    const std::string vpath = shard_path(t, dir, ".ver");
    if (FILE *f = fopen(vpath.c_str(), "rb")) {
        fclose(f);
        stream_file(t->ver, t->nrows * sizeof(i64), vpath, false);
    }
    HB_CUDA(cudaDeviceSynchronize());
    HB_API_END();
}

static void fill_perf(hb_cache *c, uint64_t call, hb_perf *perf) {
    int idx = (int)(call % hb_cache::kRing);
    const PerfRecord &r = c->ring[idx];
    std::memset(perf, 0, sizeof(*perf));
    perf->num_all = r.num_all;
    perf->num_unique = r.num_unique;
    perf->num_miss = r.num_miss;
    perf->num_evict = r.num_evict;
    perf->num_transfered = r.num_transfered;
    perf->num_remote = r.num_remote;
    perf->is_full = r.limit_full;
    perf->size = r.size;
    perf->error = r.error;
    float ms = 0.f;
    if (c->timed[idx]) {
        if (cudaEventElapsedTime(&ms, c->ev_begin[idx], c->ev_end[idx]) == cudaSuccess)
            perf->time_ms = ms;
        else
            cudaGetLastError();
    }
    // phase split (reference perf dict: sort/lookup/transfer/copy/insert, cache.cc:99-104, :189-193)
    const uint32_t mask = c->phase_mask[idx];
    auto span = [&](cudaEvent_t a, cudaEvent_t b) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, a, b) != cudaSuccess) {
            cudaGetLastError();
            t = 0.f;
        }
        return t;
    };
    cudaEvent_t *ph = &c->ev_phase[idx * hb_cache::kPhases];
    if ((mask & 3u) == 3u) {
        perf->sort_ms = span(c->ev_begin[idx], ph[0]);
        perf->lookup_ms = span(ph[0], ph[1]);
        if (r.kind == 0 && (mask & 12u) == 12u) {
            perf->transfer_ms = span(ph[1], ph[2]); // sync with the owner
            perf->copy_ms = span(ph[2], ph[3]);     // gather into dest
            perf->insert_ms = span(ph[3], c->ev_end[idx]);
        } else if (r.kind == 1 && (mask & 4u)) {
            perf->copy_ms = span(ph[1], ph[2]);     // accumulate (+ fused push): plan + data kernel
            perf->transfer_ms = span(ph[2], c->ev_end[idx]); // flush of evicted lines + cleanup
            if (mask & 8u)
                perf->kernel_ms = span(ph[3], ph[2]); // the data kernel alone (segment_reduce)
        }
    }
}

// Report the device-side failures of the calls [c->err_checked, upto] that have not been reported
// yet.  Every call's record carries its own error word (op_begin clears the device's), so a failed
// call is reported once and the cache stays usable.
static void check_errors(hb_cache *c, uint64_t upto) {
    static const char *names[] = {"", "row store slack exhausted (too many transient/pending lines)",
                                  "cache index full", "key outside the table",
                                  "pending-eviction list overflow",
                                  "owner mailbox too small (hb_cache_reserve / HERALD_MAILBOX_ROWS)",
                                  "a peer did not reach the exchange barrier"};
    uint64_t first = c->err_checked;
    if (upto + 1 > first + hb_cache::kRing)
        first = upto + 1 - hb_cache::kRing;
    u32 err = 0;
    for (uint64_t call = first; call <= upto; call++)
        err = std::max(err, c->ring[call % hb_cache::kRing].error);
    c->err_checked = upto + 1;
    if (err)
        throw Error(std::string("device-side cache failure: ") + names[std::min<u32>(err, 6)]);
}

int hb_cache_wait(hb_cache *c, hb_perf *perf) {
    HB_API_BEGIN();
    Guard g(c->device);
    HB_CUDA(cudaStreamSynchronize(c->stream));
    HB_CUDA(cudaStreamSynchronize(c->d2h));
    HB_CUDA(cudaStreamSynchronize(c->xstream));
    if (c->calls) {
        int idx = (int)((c->calls - 1) % hb_cache::kRing);
        const PerfRecord &r = c->ring[idx];
        c->occ_upper = r.ht_occupied;
        c->pending_upper = std::max<size_t>(c->pending_upper, r.pending);
        if (perf)
            fill_perf(c, c->calls - 1, perf);
        check_errors(c, c->calls - 1);
    } else if (perf) {
        std::memset(perf, 0, sizeof(*perf));
    }
    HB_API_END();
}

int hb_cache_last_call(hb_cache *c, uint64_t *seq) {
    HB_API_BEGIN();
    HB_CHECK(c->calls > 0, "no call has been enqueued");
    *seq = c->calls - 1;
    HB_API_END();
}

int hb_cache_wait_call(hb_cache *c, uint64_t seq, hb_perf *perf) {
    HB_API_BEGIN();
    Guard g(c->device);
    HB_CHECK(seq < c->calls, "no such call");
    HB_CHECK(seq + hb_cache::kRing > c->calls, "call too old: its record has been overwritten");
    const int idx = (int)(seq % hb_cache::kRing);
    HB_CUDA(cudaEventSynchronize(c->ev_end[idx]));
    if (const int k1 = c->dl_of_call[idx]) {
        // the download stream is in order: an event recorded for a later call covers this one
        HB_CUDA(cudaEventSynchronize(c->ev_dl[k1 - 1]));
    }
    if (seq + 1 == c->calls) {
        const PerfRecord &r = c->ring[idx];
        c->occ_upper = r.ht_occupied;
        c->pending_upper = std::max<size_t>(c->pending_upper, r.pending);
    }
    if (perf)
        fill_perf(c, seq, perf);
    check_errors(c, seq);
    HB_API_END();
}

int hb_cache_perf_range(hb_cache *c, uint64_t first, int count, hb_perf *out, int *kinds) {
    HB_API_BEGIN();
    Guard g(c->device);
    HB_CHECK(count >= 0 && first + (uint64_t)count <= c->calls, "no such calls");
    HB_CHECK(first + hb_cache::kRing >= c->calls, "calls too old: their records have been overwritten");
    if (count)
        HB_CUDA(cudaEventSynchronize(c->ev_end[(first + count - 1) % hb_cache::kRing]));
    for (int k = 0; k < count; k++) {
        fill_perf(c, first + k, &out[k]);
        if (kinds)
            kinds[k] = (int)c->ring[(first + k) % hb_cache::kRing].kind;
    }
    HB_API_END();
}

int hb_cache_perf_history(hb_cache *c, hb_perf *out, int *kinds, int max, int *written) {
    HB_API_BEGIN();
    Guard g(c->device);
    HB_CUDA(cudaStreamSynchronize(c->stream));
    uint64_t avail = std::min<uint64_t>(c->calls, hb_cache::kRing);
    uint64_t take = std::min<uint64_t>(avail, (uint64_t)std::max(max, 0));
    for (uint64_t k = 0; k < take; k++) {
        uint64_t call = c->calls - take + k;
        fill_perf(c, call, &out[k]);
        if (kinds)
            kinds[k] = (int)c->ring[call % hb_cache::kRing].kind;
    }
    *written = (int)take;
    HB_API_END();
}

int hb_cache_size(hb_cache *c, size_t *size) {
    HB_API_BEGIN();
    Guard g(c->device);
    HB_CUDA(cudaStreamSynchronize(c->stream));
    CacheRegs regs;
    HB_CUDA(cudaMemcpy(&regs, c->view.regs, sizeof(regs), cudaMemcpyDeviceToHost));
    *size = regs.size;
    HB_API_END();
}

static PeekResult peek(hb_cache *c, uint64_t key) {
    PeekResult *d = nullptr, h;
    HB_CUDA(cudaMalloc((void **)&d, sizeof(PeekResult)));
    HB_LAUNCH(peek_kernel, 1, 1, 0, c->stream, c->view, key, d);
    g_launches++;
    cudaError_t e = cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    HB_CUDA(e);
    return h;
}

int hb_cache_count(hb_cache *c, uint64_t key, int *count) {
    HB_API_BEGIN();
    Guard g(c->device);
    *count = peek(c, key).slot >= 0 ? 1 : 0;
    HB_API_END();
}

int hb_cache_keys(hb_cache *c, uint64_t *keys, size_t capacity, size_t *n) {
    HB_API_BEGIN();
    Guard g(c->device);
    u64 *dkeys = nullptr;
    u32 *dcount = nullptr;
    dmalloc(dkeys, c->view.capacity);
    dmalloc(dcount, 1);
    HB_CUDA(cudaMemsetAsync(dcount, 0, sizeof(u32), c->stream));
    HB_LAUNCH(collect_keys_kernel, lin_grid(c->view.capacity), 256, 0, c->stream, c->view, dkeys, dcount);
    HB_LAUNCHED();
    u32 count = 0;
    HB_CUDA(cudaMemcpyAsync(&count, dcount, sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    *n = count;
    size_t take = std::min<size_t>(count, capacity);
    std::vector<u64> host(count);
    if (count)
        HB_CUDA(cudaMemcpy(host.data(), dkeys, count * sizeof(u64), cudaMemcpyDeviceToHost));
    std::sort(host.begin(), host.end()); // LRUCache::PyAPI_keys sorts (lru_cache.cc:41-48)
    for (size_t i = 0; i < take; i++)
        keys[i] = host[i];
    dfree(dkeys);
    dfree(dcount);
    HB_API_END();
}

int hb_cache_peek(hb_cache *c, uint64_t key, int *found, int64_t *version, int64_t *updates,
                  float *data, float *grad) {
    HB_API_BEGIN();
    Guard g(c->device);
    PeekResult r = peek(c, key);
    *found = r.slot >= 0;
    if (r.slot >= 0) {
        if (version)
            *version = r.version;
        if (updates)
            *updates = r.updates;
        if (data)
            HB_CUDA(cudaMemcpy(data, c->view.data + (size_t)r.slot * c->width, c->width * sizeof(float),
                               cudaMemcpyDeviceToHost));
        if (grad) {
            if (r.updates != 0)
                HB_CUDA(cudaMemcpy(grad, c->view.grad + (size_t)r.slot * c->width,
                                   c->width * sizeof(float), cudaMemcpyDeviceToHost));
            else
                std::memset(grad, 0, c->width * sizeof(float)); // logically zero after a push
        }
    }
    HB_API_END();
}

// scratch device buffer of the scoring calls (grown on demand)
static void *score_scratch(hb_cache *c, size_t bytes) {
    if (bytes > c->score_scratch_cap) {
        sync_all(c);
        if (c->score_scratch)
            cudaFree(c->score_scratch);
        c->score_scratch = nullptr;
        HB_CUDA(cudaMalloc(&c->score_scratch, bytes));
        c->score_scratch_cap = bytes;
    }
    return c->score_scratch;
}

int hb_cache_score(hb_cache *c, const void *sample_ids, int key_kind, size_t num_samples, size_t num_tables,
                   const uint32_t *table_order, size_t top_k, int fresh, uint32_t *scores) {
    HB_API_BEGIN();
    Guard g(c->device);
    HB_CHECK(num_tables > 0 && top_k <= num_tables, "top_k must not exceed the number of tables");
    if (num_samples) {
        const size_t n = num_samples * num_tables, kb = key_kind == HB_KEYS_F32 ? 4 : 8;
        const bool ids_host = !is_device_ptr(sample_ids), out_host = !is_device_ptr(scores);
        // scratch layout: [order u32 x T][scores u32 x S (host callers)][ids (host callers)]
        const size_t off_scores = (num_tables * 4 + 15) & ~(size_t)15;
        const size_t off_ids = off_scores + ((num_samples * 4 + 15) & ~(size_t)15);
        char *scr = static_cast<char *>(score_scratch(c, off_ids + (ids_host ? n * kb : 0)));
        cudaStream_t st = c->stream;
        u32 *dorder = nullptr;
        if (table_order) {
            dorder = reinterpret_cast<u32 *>(scr);
            HB_CUDA(cudaMemcpyAsync(dorder, table_order, top_k * 4, cudaMemcpyHostToDevice, st));
        }
        const void *dids = sample_ids;
        if (ids_host) {
            HB_CUDA(cudaMemcpyAsync(scr + off_ids, sample_ids, n * kb, cudaMemcpyHostToDevice, st));
            dids = scr + off_ids;
        }
        u32 *dscores = out_host ? reinterpret_cast<u32 *>(scr + off_scores) : scores;
        const int grid = (int)std::max<size_t>(1, std::min<size_t>((num_samples + 7) / 8, (size_t)sm_count() * 8));
        if (key_kind == HB_KEYS_F32)
            HB_LAUNCH(score_samples_kernel<HB_KEYS_F32>, grid, 256, 0, st, c->view, dids, num_samples,
                      (u32)num_tables, dorder, (u32)top_k, fresh, c->pull_bound, dscores);
        else
            HB_LAUNCH(score_samples_kernel<HB_KEYS_U64>, grid, 256, 0, st, c->view, dids, num_samples,
                      (u32)num_tables, dorder, (u32)top_k, fresh, c->pull_bound, dscores);
        HB_LAUNCHED();
        if (out_host) {
            HB_CUDA(cudaMemcpyAsync(scores, dscores, num_samples * 4, cudaMemcpyDeviceToHost, st));
            HB_CUDA(cudaStreamSynchronize(st));
        }
    }
    HB_API_END();
}

int hb_cache_probe(hb_cache *c, const void *keys, int key_kind, size_t n, uint8_t *resident) {
    HB_API_BEGIN();
    Guard g(c->device);
    if (n) {
        const size_t kb = key_kind == HB_KEYS_F32 ? 4 : 8;
        const bool keys_host = !is_device_ptr(keys), out_host = !is_device_ptr(resident);
        const size_t off_keys = (n + 15) & ~(size_t)15;
        char *scr = static_cast<char *>(score_scratch(c, off_keys + (keys_host ? n * kb : 0)));
        cudaStream_t st = c->stream;
        const void *dkeys = keys;
        if (keys_host) {
            HB_CUDA(cudaMemcpyAsync(scr + off_keys, keys, n * kb, cudaMemcpyHostToDevice, st));
            dkeys = scr + off_keys;
        }
        u8 *dres = out_host ? reinterpret_cast<u8 *>(scr) : resident;
        if (key_kind == HB_KEYS_F32)
            HB_LAUNCH(probe_keys_kernel<HB_KEYS_F32>, lin_grid(n), 256, 0, st, c->view, dkeys, n, dres);
        else
            HB_LAUNCH(probe_keys_kernel<HB_KEYS_U64>, lin_grid(n), 256, 0, st, c->view, dkeys, n, dres);
        HB_LAUNCHED();
        if (out_host) {
            HB_CUDA(cudaMemcpyAsync(resident, dres, n, cudaMemcpyDeviceToHost, st));
            HB_CUDA(cudaStreamSynchronize(st));
        }
    }
    HB_API_END();
}

int hb_cache_touch(hb_cache *c, uint64_t key, int *found, int64_t *version, float *data) {
    HB_API_BEGIN();
    Guard g(c->device);
    ensure_batch(c, 1);
    sync_all(c);
    cudaStream_t st = c->stream;
    KeyWorkspace &ws = c->ws[c->cur];
    // a batched lookup of one key without the insert/sync half: CacheBase::lookup via python
    begin_call(c, false, 1, c->cur);
    ws.sorted_valid = false;
    HB_LAUNCH(single_key_kernel, 1, 1, 0, st, ws.uniq, ws.num_unique, key);
    HB_LAUNCHED();
    u64 *clk = clk_of(c);
    HB_LAUNCH(resolve_kernel<NoSegPlan>, 1, kScanBlock, 0, st, c->view, ws.uniq, ws.num_unique, c->uslot[0],
                                             c->miss_list[0], c->bypass ? 1 : 0, ws.next_scan(), 1,
                                             clk, clk + 1, 0, 0, 0, NoSegPlan{});
    HB_LAUNCHED();
    // a miss reserved a slot for a fresh line; materialise and hand it back (lookup() alone
    // allocates nothing)
    HB_LAUNCH(free_transient_kernel, 1, 32, 0, st, c->view, c->uslot[0], c->miss_list[0], 0, 0);
    HB_LAUNCHED();
    end_call(c, 1, 0, 1, false);
    HB_CUDA(cudaStreamSynchronize(st));
    PeekResult r = peek(c, key);
    *found = r.slot >= 0;
    if (r.slot >= 0) {
        if (version)
            *version = r.version;
        if (data)
            HB_CUDA(cudaMemcpy(data, c->view.data + (size_t)r.slot * c->width, c->width * sizeof(float),
                               cudaMemcpyDeviceToHost));
    }
    HB_API_END();
}

int hb_cache_insert(hb_cache *c, uint64_t key, int64_t version, const float *data) {
    HB_API_BEGIN();
    Guard g(c->device);
    ensure_batch(c, 1);
    maybe_rebuild_index(c, 1);
    sync_all(c);
    cudaStream_t st = c->stream;
    float *ddata = rows_stage(c, 1, 2);
    HB_CUDA(cudaMemcpyAsync(ddata, data, c->width * sizeof(float), cudaMemcpyDefault, st));
    PeekResult r = peek(c, key);
    if (r.slot >= 0) {
        HB_LAUNCH(set_line_kernel, 1, 128, 0, st, c->view, r.slot, version, ddata);
        HB_LAUNCHED();
        if (r.state == S_CACHED) {
            HB_LAUNCH(reinsert_touch_kernel, 1, 1, 0, st, c->view, r.slot);
            HB_LAUNCHED();
        }
    } else {
        KeyWorkspace &ws = c->ws[c->cur];
        begin_call(c, false, 1, c->cur);
        ws.sorted_valid = false;
        HB_LAUNCH(single_key_kernel, 1, 1, 0, st, ws.uniq, ws.num_unique, key);
        HB_LAUNCHED();
        u64 *clk = clk_of(c);
        // bypass=1: resolve as a miss without touching anything, which allocates the fresh line
        HB_LAUNCH(resolve_kernel<NoSegPlan>, 1, kScanBlock, 0, st, c->view, ws.uniq, ws.num_unique, c->uslot[0],
                                                 c->miss_list[0], 1, ws.next_scan(), 1, clk, clk, 0, 0, 0, NoSegPlan{});
        HB_LAUNCHED();
        i32 *dslot = nullptr, hslot = -1;
        dmalloc(dslot, 1);
        HB_LAUNCH(read_slot_kernel, 1, 1, 0, st, c->uslot[0], dslot);
        HB_LAUNCHED();
        HB_CUDA(cudaMemcpyAsync(&hslot, dslot, sizeof(i32), cudaMemcpyDeviceToHost, st));
        HB_CUDA(cudaStreamSynchronize(st));
        dfree(dslot);
        HB_CHECK(hslot >= 0, "no free slot for insert");
        HB_LAUNCH(set_line_kernel, 1, 128, 0, st, c->view, hslot, version, ddata);
        HB_LAUNCHED();
        run_insert(c, 1, 0, st);
        c->pending_upper += 1;
        end_call(c, 1, 0, 1, true);
    }
    HB_CUDA(cudaStreamSynchronize(st));
    HB_API_END();
}

} // extern "C"
