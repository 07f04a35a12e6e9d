// Stable LSD radix sort + sorted-unique (see hb_sort.cuh for what it replaces).
//
// Sort: 9-bit digits (3 passes for ids < 2^27, i.e. the 33.7M-row Criteo table and the 1e8-row
// sweep).  One kernel counts the digits of every pass; then one kernel per pass ranks a tile,
// finds its digit offsets by chained look-back over the earlier tiles and scatters (stable:
// "onesweep").  Ranking inside a tile is warp-cooperative: nine ballots group the lanes of a
// warp that hold the same digit, the lowest lane of each group bumps the warp's digit counter in
// shared memory, so there are no shared-memory atomics and hot (Zipf) keys do not serialise.
// HBM traffic per pass: 1 read + 1 write of 12 B/key; at N = 212,992 everything is L2-resident.
#include <algorithm>

#include "hb_sort.cuh"

namespace hb {

namespace {

constexpr int RB = kSortRadixBits;
constexpr int RADIX = kSortRadix;
constexpr int SORT_THREADS = 512; // 52 tiles at N = 212,992: the look-back chain is what a pass costs
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS; // 2048 keys per block

template <int KIND>
__device__ __forceinline__ u64 load_key(const void *in, size_t e) {
    if (KIND == HB_KEYS_F32)
        return key_from_f32(reinterpret_cast<const float *>(in)[e]);
    return reinterpret_cast<const u64 *>(in)[e];
}

// Each warp walks its SORT_ITEMS*32 consecutive keys in order, 32 at a time.  On return
// s_cnt[w][d] = number of keys with digit d in warp w's span, rank[r] = how many keys with the
// same digit precede this one inside the warp's span (stable).
__device__ __forceinline__ void rank_tile(int shift, u32 (*s_cnt)[RADIX], const u64 (&key)[SORT_ITEMS],
                                          u32 (&rank)[SORT_ITEMS], const bool (&valid)[SORT_ITEMS]) {
    const unsigned warp = threadIdx.x >> 5;
    for (int b = threadIdx.x; b < SORT_WARPS * RADIX; b += SORT_THREADS)
        (&s_cnt[0][0])[b] = 0;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        unsigned vm = __ballot_sync(FULL, valid[r]);
        if (valid[r]) {
            u32 d = (u32)(key[r] >> shift) & (RADIX - 1);
            // lanes holding the same digit: RB ballots
            unsigned peers = vm;
#pragma unroll
            for (int bit = 0; bit < RB; bit++) {
                const bool one = (d >> bit) & 1u;
                const unsigned bal = __ballot_sync(vm, one);
                peers &= one ? bal : ~bal;
            }
            u32 base = s_cnt[warp][d];
            __syncwarp(vm);
            if ((peers & lanemask_lt()) == 0) // lowest lane of the group
                s_cnt[warp][d] = base + __popc(peers);
            rank[r] = base + __popc(peers & lanemask_lt());
        }
        __syncwarp();
    }
    __syncthreads();
}

// Digit histograms of EVERY pass in one read of the keys (the totals do not depend on the order
// the keys are in).  totals[p][d] must be zero on entry (they live in the scan arena).
template <int KIND>
__global__ void __launch_bounds__(SORT_THREADS)
    sort_hist_all_kernel(const void *kin, size_t n, int passes, u32 *totals, const u32 *mismatch) {
    pdl_enter();
    __shared__ u32 sh[kMaxSortPasses * RADIX];
    if (mismatch && *mismatch == 0)
        return; // same keys as the batch this workspace already holds sorted
    for (int b = threadIdx.x; b < passes * RADIX; b += SORT_THREADS)
        sh[b] = 0;
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * SORT_THREADS;
    for (size_t e = (size_t)blockIdx.x * SORT_THREADS + threadIdx.x; e < n; e += stride) {
        const u64 key = load_key<KIND>(kin, e);
        for (int p = 0; p < passes; p++)
            atomicAdd(&sh[p * RADIX + ((u32)(key >> (p * RB)) & (RADIX - 1))], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < passes * RADIX; b += SORT_THREADS)
        if (sh[b])
            atomicAdd(&totals[b], sh[b]);
}

// status word of one (tile, digit): epoch << 34 | flag << 32 | count.  flag 1 = the tile's own
// count, 2 = inclusive prefix over tiles 0..tile.  A word whose epoch is not the running pass's
// counts as "not written yet", so the status array is never cleared.
__device__ __forceinline__ u64 pack_status(u32 epoch, u32 flag, u32 value) {
    return ((u64)epoch << 34) | ((u64)flag << 32) | value;
}

// One pass of the stable LSD radix sort in a single kernel: rank inside the tile, chained
// look-back over the earlier tiles per digit, scatter.  Tiles are numbered by an atomic ticket,
// so a tile only ever waits for tiles that are already running.
// FIRST: keys come from the caller's array (KIND) and the payload is the position.  PIN / POUT:
// the input / output of the pass is one packed word per key, key << 32 | original index (used
// when the keys fit 32 bits: one scattered store per key instead of two).
template <int KIND, bool FIRST, bool PIN, bool POUT>
__global__ void __launch_bounds__(SORT_THREADS, 1)
    sort_pass_kernel(const void *kin, const u32 *vin, u64 *kout, u32 *vout, size_t n, int shift,
                     const u32 *__restrict__ totals, u64 *status, u32 *counts, u32 *ticket, u32 epoch,
                     const u32 *mismatch) {
    pdl_enter();
    __shared__ u32 s_cnt[SORT_WARPS][RADIX];
    if (mismatch && *mismatch == 0)
        return;
    __shared__ u32 s_off[RADIX];
    __shared__ u32 s_wsum[SORT_WARPS];
    u64 key[SORT_ITEMS];
    u32 rank[SORT_ITEMS];
    bool valid[SORT_ITEMS];
    const u32 tile = take_ticket(ticket);
    const size_t tile_base = (size_t)tile * SORT_TILE;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const size_t warp_base = tile_base + (size_t)warp * (SORT_ITEMS * 32);
    u32 idx[SORT_ITEMS];
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        const size_t e = warp_base + r * 32 + lane;
        valid[r] = e < n;
        key[r] = 0;
        idx[r] = (u32)e;
        if (valid[r]) {
            if (FIRST) {
                key[r] = load_key<KIND>(kin, e);
            } else if (PIN) {
                const u64 w = reinterpret_cast<const u64 *>(kin)[e];
                key[r] = w >> 32;
                idx[r] = (u32)w;
            } else {
                key[r] = reinterpret_cast<const u64 *>(kin)[e];
                idx[r] = vin[e];
            }
        }
    }
    rank_tile(shift, s_cnt, key, rank, valid);
    constexpr int DPT = RADIX / SORT_THREADS; // consecutive digits per thread
    // With few tiles (a WDL batch: 52) every tile publishes its own digit counts once and sums the
    // counts of ALL its predecessors with independent loads: one L2 round trip, where the chained
    // look-back below pays one per 16 predecessors (the tiles of a small sort start together, so
    // nobody finds a finished prefix to stop at).  tag = epoch folded to 19 bits, never 0.
    const bool direct = gridDim.x <= (unsigned)kSortDirectTiles;
    const u32 tag = (epoch % 0x7fffeu) + 1u;
    u32 cnt[DPT], tot[DPT];
    u32 tsum = 0;
#pragma unroll
    for (int k = 0; k < DPT; k++) {
        const int b = threadIdx.x * DPT + k;
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) { // exclusive prefix over the tile's warps
            u32 t = s_cnt[w][b];
            s_cnt[w][b] = run;
            run += t;
        }
        cnt[k] = run;
        if (direct)
            *reinterpret_cast<volatile u32 *>(&counts[(size_t)tile * RADIX + b]) = (tag << 13) | run;
        else
            *reinterpret_cast<volatile u64 *>(&status[(size_t)tile * RADIX + b]) =
                pack_status(epoch, tile == 0 ? 2u : 1u, run);
        tot[k] = totals[b];
        tsum += tot[k];
    }
    // exclusive scan of the digit totals over the block -> first output position of each digit
    u32 incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(FULL, incl, d);
        if (lane >= (unsigned)d)
            incl += t;
    }
    if (lane == 31)
        s_wsum[warp] = incl;
    __syncthreads();
    u32 base = incl - tsum;
    for (unsigned w = 0; w < warp; w++)
        base += s_wsum[w];
    // digit offsets of this tile = first position of the digit + its count in the earlier tiles
#pragma unroll
    for (int k = 0; k < DPT; k++) {
        const int b = threadIdx.x * DPT + k;
        u32 excl = 0;
        if (direct) {
            u32 sv[kSortDirectTiles];
#pragma unroll
            for (int i = 0; i < kSortDirectTiles; i++)
                sv[i] = i < (int)tile ? *reinterpret_cast<volatile u32 *>(&counts[(size_t)i * RADIX + b])
                                      : (tag << 13);
#pragma unroll
            for (int i = 0; i < kSortDirectTiles; i++) {
                if (i < (int)tile) {
                    while ((sv[i] >> 13) != tag)
                        sv[i] = *reinterpret_cast<volatile u32 *>(&counts[(size_t)i * RADIX + b]);
                    excl += sv[i] & 0x1fffu;
                }
            }
        } else if (tile > 0) {
            // kLook predecessors per round trip (independent loads); stop at the nearest one
            // that already holds a prefix
            constexpr int kLook = 16;
            int look = (int)tile - 1;
            bool done = false;
            while (!done) {
                u64 sv[kLook];
#pragma unroll
                for (int i = 0; i < kLook; i++) {
                    const int idx = look - i;
                    sv[i] = idx >= 0 ? *reinterpret_cast<volatile u64 *>(&status[(size_t)idx * RADIX + b])
                                     : pack_status(epoch, 2u, 0);
                }
#pragma unroll
                for (int i = 0; i < kLook; i++) {
                    const int idx = look - i;
                    if (!done) {
                        while ((u32)(sv[i] >> 34) != epoch || ((sv[i] >> 32) & 3u) == 0)
                            sv[i] = *reinterpret_cast<volatile u64 *>(&status[(size_t)idx * RADIX + b]);
                        excl += (u32)sv[i];
                        done = ((sv[i] >> 32) & 3u) == 2u;
                    }
                }
                look -= kLook;
            }
            *reinterpret_cast<volatile u64 *>(&status[(size_t)tile * RADIX + b]) =
                pack_status(epoch, 2u, excl + cnt[k]);
        }
        s_off[b] = base + excl;
        base += tot[k];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        if (valid[r]) {
            u32 d = (u32)(key[r] >> shift) & (RADIX - 1);
            u32 pos = s_off[d] + s_cnt[warp][d] + rank[r];
            if (POUT) {
                kout[pos] = (key[r] << 32) | idx[r];
            } else {
                kout[pos] = key[r];
                vout[pos] = idx[r];
            }
        }
    }
}

// ---- unique: flags + single-pass scan --------------------------------------------------
constexpr int UNIQ_ITEMS = 4;
constexpr int UNIQ_TILE = kScanBlock * UNIQ_ITEMS;

__global__ void __launch_bounds__(kScanBlock)
    unique_kernel(const u64 *sk, const u32 *perm, size_t n, u64 *uniq, u32 *inverse,
                  u32 *seg_start, u32 *num_unique, ScanState st, u32 ntiles, const u32 *mismatch) {
    pdl_enter();
    if (mismatch && *mismatch == 0)
        return;
    const u32 tile = take_ticket(st.ticket);
    const size_t base = (size_t)tile * UNIQ_TILE + (size_t)threadIdx.x * UNIQ_ITEMS;
    u64 k[UNIQ_ITEMS];
    bool head[UNIQ_ITEMS];
    u64 prev = 0;
    if (base > 0 && base < n)
        prev = sk[base - 1];
    u32 cnt = 0;
#pragma unroll
    for (int j = 0; j < UNIQ_ITEMS; j++) {
        size_t p = base + j;
        head[j] = false;
        if (p < n) {
            k[j] = sk[p];
            head[j] = (p == 0) || (k[j] != prev);
            prev = k[j];
            cnt += head[j];
        }
    }
    ScanResult sr = grid_exclusive_scan<kScanBlock>(st, cnt, tile);
    u32 r = sr.excl; // number of heads strictly before this thread's first element
#pragma unroll
    for (int j = 0; j < UNIQ_ITEMS; j++) {
        size_t p = base + j;
        if (p < n) {
            if (head[j]) {
                uniq[r] = k[j];
                seg_start[r] = (u32)p;
                r++;
            }
            inverse[perm[p]] = r - 1;
        }
    }
    if (tile == ntiles - 1 && threadIdx.x == 0) {
        u32 total = sr.tile_prefix + sr.tile_total;
        *num_unique = total;
        seg_start[total] = (u32)n;
    }
}

// Does the new batch hold the same key sequence as the one this workspace has sorted?  Exact:
// key(new[i]) == uniq[inverse[i]] for every i.  *mismatch (zeroed with the scan arena) counts
// the positions that differ.
template <int KIND>
__global__ void __launch_bounds__(256)
    same_keys_kernel(const void *kin, size_t n, const u64 *__restrict__ uniq,
                     const u32 *__restrict__ inverse, const u32 *__restrict__ num_unique,
                     u32 *mismatch) {
    pdl_enter();
    const u32 U = *num_unique;
    bool bad = false;
    // kSame elements per thread, each level of the dependent loads (inverse -> uniq) issued for
    // all of them before the first use
    constexpr int kSame = 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * kSame) {
        u32 r[kSame];
        u64 k[kSame], q[kSame];
#pragma unroll
        for (int j = 0; j < kSame; j++) {
            const size_t i = i0 + j * stride;
            r[j] = i < n ? inverse[i] : 0;
            k[j] = i < n ? load_key<KIND>(kin, i) : 0;
        }
#pragma unroll
        for (int j = 0; j < kSame; j++)
            q[j] = (i0 + j * stride < n && r[j] < U) ? uniq[r[j]] : ~0ull;
#pragma unroll
        for (int j = 0; j < kSame; j++)
            if (i0 + j * stride < n)
                bad |= r[j] >= U || q[j] != k[j];
    }
    if (__any_sync(FULL, bad) && lane_id() == 0)
        atomicAdd(mismatch, 1u);
}

__global__ void set_zero_unique(u32 *num_unique, u32 *seg_start) {
    pdl_enter();
    *num_unique = 0;
    seg_start[0] = 0;
}

template <typename T>
void dev_alloc(T *&p, size_t count) {
    HB_CUDA(cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T)));
}
template <typename T>
void dev_free(T *&p) {
    if (p)
        cudaFree(p);
    p = nullptr;
}

} // namespace

const u32 *check_same_keys(KeyWorkspace &ws, const void *keys_in, int key_kind, size_t n,
                           cudaStream_t st) {
    if (!ws.sorted_valid || n == 0 || n != ws.sorted_n)
        return nullptr;
    u32 *mismatch = ws.reuse_mismatch();
    int grid = std::min(ceil_div(n, 256 * 4), sm_count() * 4);
    if (key_kind == HB_KEYS_F32)
        HB_LAUNCH(same_keys_kernel<HB_KEYS_F32>, grid, 256, 0, st, keys_in, n, ws.uniq, ws.inverse,
                                                           ws.num_unique, mismatch);
    else
        HB_LAUNCH(same_keys_kernel<HB_KEYS_U64>, grid, 256, 0, st, keys_in, n, ws.uniq, ws.inverse,
                                                           ws.num_unique, mismatch);
    HB_LAUNCHED();
    return mismatch;
}

int bits_for(u64 max_key_exclusive) {
    int bits = 1;
    while (bits < 64 && (1ull << bits) < max_key_exclusive)
        bits++;
    return bits;
}

void KeyWorkspace::reserve(size_t n) {
    if (n <= cap)
        return;
    HB_CUDA(cudaDeviceSynchronize()); // buffers may be in flight
    release();
    size_t c = std::max<size_t>(n, 4096);
    c = (c + 4095) / 4096 * 4096;
    cap = c; // (the arena sizes below depend on it)
    for (int i = 0; i < 2; i++) {
        dev_alloc(keys[i], c);
        dev_alloc(vals[i], c);
    }
    nblk_cap = (c + SORT_TILE - 1) / SORT_TILE;
    dev_alloc(sort_status, (size_t)RADIX * nblk_cap);
    HB_CUDA(cudaMemset(sort_status, 0, (size_t)RADIX * nblk_cap * sizeof(u64)));
    dev_alloc(sort_counts, (size_t)RADIX * kSortDirectTiles);
    HB_CUDA(cudaMemset(sort_counts, 0, (size_t)RADIX * kSortDirectTiles * sizeof(u32)));
    dev_alloc(uniq, c);
    dev_alloc(inverse, c);
    dev_alloc(seg_start, c + 1);
    dev_alloc(num_unique, 1);
    ntile_cap = (c + kScanBlock - 1) / kScanBlock + 1;
    dev_alloc(scan_arena, arena_words());
    HB_CUDA(cudaMemset(scan_arena, 0, arena_words() * sizeof(u64)));
    dev_alloc(side_arena, side_words());
    HB_CUDA(cudaMemset(side_arena, 0, side_words() * sizeof(u64)));
    sorted_valid = false;
    dev_alloc(hot_a, c);
    dev_alloc(hot_b, c);
    dev_alloc(medium, c);
    {
        char *p = nullptr;
        dev_alloc(p, (c + 32) * 32); // kSegItemBytes per work item
        seg_items = p;
    }
    cap = c;
}

void KeyWorkspace::release() {
    for (int i = 0; i < 2; i++) {
        dev_free(keys[i]);
        dev_free(vals[i]);
    }
    dev_free(sort_status);
    dev_free(sort_counts);
    dev_free(uniq);
    dev_free(inverse);
    dev_free(seg_start);
    dev_free(num_unique);
    dev_free(scan_arena);
    dev_free(side_arena);
    dev_free(hot_a);
    dev_free(hot_b);
    dev_free(split_partials);
    split_partials_cap = 0;
    dev_free(medium);
    {
        char *p = reinterpret_cast<char *>(seg_items);
        dev_free(p);
        seg_items = nullptr;
    }
    cap = 0;
}

void KeyWorkspace::reset_main(cudaStream_t st) {
    HB_CUDA(cudaMemsetAsync(scan_arena, 0, arena_words() * sizeof(u64), st));
    scan_next = 0;
}

void KeyWorkspace::reset_side(cudaStream_t st) {
    HB_CUDA(cudaMemsetAsync(side_arena, 0, side_words() * sizeof(u64), st));
}

ScanState KeyWorkspace::side_scan() const {
    ScanState s;
    s.ticket = reinterpret_cast<u32 *>(side_arena);
    s.status = side_arena + 1;
    return s;
}

ScanState KeyWorkspace::next_scan() {
    HB_CHECK(scan_next < kScanSlots, "scan arena exhausted");
    u64 *slot = scan_arena + (size_t)scan_next * scan_slot_words();
    scan_next++;
    ScanState s;
    s.ticket = reinterpret_cast<u32 *>(slot);
    s.status = slot + 1;
    return s;
}

SortedKeys radix_sort_keys(KeyWorkspace &ws, const void *keys_in, int key_kind, size_t n,
                           int key_bits, cudaStream_t st, const u32 *mismatch) {
    HB_CHECK(n <= ws.cap, "sort workspace too small");
    HB_CHECK(n < (1ull << 32), "too many keys");
    int passes = std::min(kMaxSortPasses, std::max(1, (key_bits + RB - 1) / RB));
    int nblk = ceil_div(n, SORT_TILE);
    u32 *totals = ws.sort_totals();
    int hgrid = std::max(1, std::min(ceil_div(n, 1024), sm_count() * 2));
    if (key_kind == HB_KEYS_F32)
        HB_LAUNCH(sort_hist_all_kernel<HB_KEYS_F32>, hgrid, SORT_THREADS, 0, st, keys_in, n, passes, totals, mismatch);
    else
        HB_LAUNCH(sort_hist_all_kernel<HB_KEYS_U64>, hgrid, SORT_THREADS, 0, st, keys_in, n, passes, totals, mismatch);
    HB_LAUNCHED();
    const void *kin = keys_in;
    const u32 *vin = nullptr;
    int out = 0;
    for (int p = 0; p < passes; p++) {
        int shift = p * RB;
        bool first = p == 0;
        bool f32 = first && key_kind == HB_KEYS_F32;
        u32 epoch = ws.next_sort_epoch();
        if (nblk <= kSortDirectTiles && epoch % 0x7fffeu == 0) // the 19-bit tag starts over
            HB_CUDA(cudaMemsetAsync(ws.sort_counts, 0, (size_t)RADIX * kSortDirectTiles * sizeof(u32), st));
        const u32 *tp = totals + (size_t)p * RADIX;
        u32 *ticket = ws.sort_tickets() + p;
        // keys below 2^32 travel packed with their index between the passes
        const bool packed = key_bits <= 32 && passes > 1;
        const bool pin = packed && !first, pout = packed && p + 1 < passes;
#define HB_SORT_PASS(KIND, F, PI, PO)                                                             \
    HB_LAUNCH((sort_pass_kernel<KIND, F, PI, PO>), nblk, SORT_THREADS, 0, st, \
        kin, vin, ws.keys[out], ws.vals[out], n, shift, tp, ws.sort_status, ws.sort_counts, ticket, epoch, mismatch)
        if (f32 && pout)
            HB_SORT_PASS(HB_KEYS_F32, true, false, true);
        else if (f32)
            HB_SORT_PASS(HB_KEYS_F32, true, false, false);
        else if (first && pout)
            HB_SORT_PASS(HB_KEYS_U64, true, false, true);
        else if (first)
            HB_SORT_PASS(HB_KEYS_U64, true, false, false);
        else if (pin && pout)
            HB_SORT_PASS(HB_KEYS_U64, false, true, true);
        else if (pin)
            HB_SORT_PASS(HB_KEYS_U64, false, true, false);
        else
            HB_SORT_PASS(HB_KEYS_U64, false, false, false);
#undef HB_SORT_PASS
        HB_LAUNCHED();
        kin = ws.keys[out];
        vin = ws.vals[out];
        out ^= 1;
    }
    SortedKeys sk;
    sk.keys = reinterpret_cast<const u64 *>(kin);
    sk.perm = vin;
    return sk;
}

void unique_from_sorted(KeyWorkspace &ws, const SortedKeys &sk, size_t n, cudaStream_t st,
                        const u32 *mismatch) {
    if (n == 0) {
        HB_LAUNCH(set_zero_unique, 1, 1, 0, st, ws.num_unique, ws.seg_start);
        HB_LAUNCHED();
        return;
    }
    u32 ntiles = (u32)ceil_div(n, UNIQ_TILE);
    HB_LAUNCH(unique_kernel, ntiles, kScanBlock, 0, st, sk.keys, sk.perm, n, ws.uniq, ws.inverse,
                                                 ws.seg_start, ws.num_unique, ws.side_scan(),
                                                 ntiles, mismatch);
    HB_LAUNCHED();
}

} // namespace hb
