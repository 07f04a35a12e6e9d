// Stable LSD radix sort + sorted-unique (see hb_sort.cuh for what it replaces).
//
// Sort: 9-bit digits (3 passes for ids < 2^27, i.e. the 33.7M-row Criteo table and the 1e8-row
// sweep).  Per pass: tile histogram -> one-block scan of the [digit][tile] matrix -> stable
// scatter.  Ranking inside a tile is warp-cooperative: __match_any_sync groups the lanes of a
// warp that hold the same digit, the lowest lane of each group bumps the warp's digit counter in
// shared memory, so there are no shared-memory atomics and hot (Zipf) keys do not serialise.
// HBM traffic per pass: 2 reads + 1 write of 12 B/key; at N = 212,992 everything is L2-resident.
#include <algorithm>

#include "hb_sort.cuh"

namespace hb {

namespace {

constexpr int RB = 9;
constexpr int RADIX = 1 << RB;
constexpr int SORT_THREADS = 256;
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
template <int KIND>
__device__ __forceinline__ void rank_tile(const void *kin, size_t n, int shift, size_t tile_base,
                                          u32 (*s_cnt)[RADIX], u64 (&key)[SORT_ITEMS],
                                          u32 (&rank)[SORT_ITEMS], bool (&valid)[SORT_ITEMS]) {
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    for (int b = threadIdx.x; b < SORT_WARPS * RADIX; b += SORT_THREADS)
        (&s_cnt[0][0])[b] = 0;
    __syncthreads();
    const size_t warp_base = tile_base + (size_t)warp * (SORT_ITEMS * 32);
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        size_t e = warp_base + r * 32 + lane;
        valid[r] = e < n;
        key[r] = valid[r] ? load_key<KIND>(kin, e) : 0;
    }
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        unsigned vm = __ballot_sync(FULL, valid[r]);
        if (valid[r]) {
            u32 d = (u32)(key[r] >> shift) & (RADIX - 1);
            unsigned peers = __match_any_sync(vm, d);
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

template <int KIND>
__global__ void __launch_bounds__(SORT_THREADS)
    sort_hist_kernel(const void *kin, size_t n, int shift, u32 *blk_hist, int nblk) {
    __shared__ u32 s_cnt[SORT_WARPS][RADIX];
    u64 key[SORT_ITEMS];
    u32 rank[SORT_ITEMS];
    bool valid[SORT_ITEMS];
    rank_tile<KIND>(kin, n, shift, (size_t)blockIdx.x * SORT_TILE, s_cnt, key, rank, valid);
    for (int b = threadIdx.x; b < RADIX; b += SORT_THREADS) {
        u32 sum = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++)
            sum += s_cnt[w][b];
        blk_hist[(size_t)b * nblk + blockIdx.x] = sum;
    }
}

// In-place exclusive scan of `total` counters by one block.
__global__ void __launch_bounds__(1024) sort_scan_kernel(u32 *data, size_t total) {
    __shared__ u32 s_warp[32];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    size_t chunk = (total + 1023) / 1024;
    size_t lo = (size_t)threadIdx.x * chunk, hi = min(lo + chunk, total);
    u32 sum = 0;
    for (size_t i = lo; i < hi; i++)
        sum += data[i];
    u32 incl = sum;
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
        u32 w = s_warp[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 t = __shfl_up_sync(FULL, wi, d);
            if (lane >= (unsigned)d)
                wi += t;
        }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    u32 run = s_warp[warp] + incl - sum;
    for (size_t i = lo; i < hi; i++) {
        u32 v = data[i];
        data[i] = run;
        run += v;
    }
}

template <int KIND, bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS)
    sort_scatter_kernel(const void *kin, const u32 *vin, u64 *kout, u32 *vout, size_t n, int shift,
                        const u32 *offsets, int nblk) {
    __shared__ u32 s_cnt[SORT_WARPS][RADIX];
    u64 key[SORT_ITEMS];
    u32 rank[SORT_ITEMS];
    bool valid[SORT_ITEMS];
    const size_t tile_base = (size_t)blockIdx.x * SORT_TILE;
    rank_tile<KIND>(kin, n, shift, tile_base, s_cnt, key, rank, valid);
    // per digit: global offset of this tile, then exclusive prefix over the tile's warps
    for (int b = threadIdx.x; b < RADIX; b += SORT_THREADS) {
        u32 run = offsets[(size_t)b * nblk + blockIdx.x];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            u32 t = s_cnt[w][b];
            s_cnt[w][b] = run;
            run += t;
        }
    }
    __syncthreads();
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const size_t warp_base = tile_base + (size_t)warp * (SORT_ITEMS * 32);
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        if (valid[r]) {
            size_t e = warp_base + r * 32 + lane;
            u32 d = (u32)(key[r] >> shift) & (RADIX - 1);
            u32 pos = s_cnt[warp][d] + rank[r];
            kout[pos] = key[r];
            vout[pos] = FIRST ? (u32)e : vin[e];
        }
    }
}

// ---- unique: flags + single-pass scan --------------------------------------------------
constexpr int UNIQ_ITEMS = 4;
constexpr int UNIQ_TILE = kScanBlock * UNIQ_ITEMS;

__global__ void __launch_bounds__(kScanBlock)
    unique_kernel(const u64 *sk, const u32 *perm, size_t n, u64 *uniq, u32 *inverse,
                  u32 *seg_start, u32 *num_unique, ScanState st, u32 ntiles) {
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

__global__ void set_zero_unique(u32 *num_unique, u32 *seg_start) {
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
    for (int i = 0; i < 2; i++) {
        dev_alloc(keys[i], c);
        dev_alloc(vals[i], c);
    }
    nblk_cap = (c + SORT_TILE - 1) / SORT_TILE;
    dev_alloc(blk_hist, (size_t)RADIX * nblk_cap);
    dev_alloc(uniq, c);
    dev_alloc(inverse, c);
    dev_alloc(seg_start, c + 1);
    dev_alloc(num_unique, 1);
    ntile_cap = (c + kScanBlock - 1) / kScanBlock + 1;
    dev_alloc(scan_arena, arena_words());
    HB_CUDA(cudaMemset(scan_arena, 0, arena_words() * sizeof(u64)));
    dev_alloc(hot_a, c);
    dev_alloc(hot_b, c);
    cap = c;
}

void KeyWorkspace::release() {
    for (int i = 0; i < 2; i++) {
        dev_free(keys[i]);
        dev_free(vals[i]);
    }
    dev_free(blk_hist);
    dev_free(uniq);
    dev_free(inverse);
    dev_free(seg_start);
    dev_free(num_unique);
    dev_free(scan_arena);
    dev_free(hot_a);
    dev_free(hot_b);
    cap = 0;
}

void KeyWorkspace::reset_scans(cudaStream_t st) {
    HB_CUDA(cudaMemsetAsync(scan_arena, 0, arena_words() * sizeof(u64), st));
    scan_next = 0;
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
                           int key_bits, cudaStream_t st) {
    HB_CHECK(n <= ws.cap, "sort workspace too small");
    HB_CHECK(n < (1ull << 32), "too many keys");
    int passes = std::max(1, (key_bits + RB - 1) / RB);
    int nblk = ceil_div(n, SORT_TILE);
    const void *kin = keys_in;
    const u32 *vin = nullptr;
    int out = 0;
    for (int p = 0; p < passes; p++) {
        int shift = p * RB;
        bool first = p == 0;
        bool f32 = first && key_kind == HB_KEYS_F32;
        if (f32)
            sort_hist_kernel<HB_KEYS_F32><<<nblk, SORT_THREADS, 0, st>>>(kin, n, shift,
                                                                         ws.blk_hist, nblk);
        else
            sort_hist_kernel<HB_KEYS_U64><<<nblk, SORT_THREADS, 0, st>>>(kin, n, shift,
                                                                         ws.blk_hist, nblk);
        HB_LAUNCHED();
        sort_scan_kernel<<<1, 1024, 0, st>>>(ws.blk_hist, (size_t)RADIX * nblk);
        HB_LAUNCHED();
        if (f32)
            sort_scatter_kernel<HB_KEYS_F32, true><<<nblk, SORT_THREADS, 0, st>>>(
                kin, vin, ws.keys[out], ws.vals[out], n, shift, ws.blk_hist, nblk);
        else if (first)
            sort_scatter_kernel<HB_KEYS_U64, true><<<nblk, SORT_THREADS, 0, st>>>(
                kin, vin, ws.keys[out], ws.vals[out], n, shift, ws.blk_hist, nblk);
        else
            sort_scatter_kernel<HB_KEYS_U64, false><<<nblk, SORT_THREADS, 0, st>>>(
                kin, vin, ws.keys[out], ws.vals[out], n, shift, ws.blk_hist, nblk);
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

void unique_from_sorted(KeyWorkspace &ws, const SortedKeys &sk, size_t n, cudaStream_t st) {
    if (n == 0) {
        set_zero_unique<<<1, 1, 0, st>>>(ws.num_unique, ws.seg_start);
        HB_LAUNCHED();
        return;
    }
    u32 ntiles = (u32)ceil_div(n, UNIQ_TILE);
    unique_kernel<<<ntiles, kScanBlock, 0, st>>>(sk.keys, sk.perm, n, ws.uniq, ws.inverse,
                                                 ws.seg_start, ws.num_unique, ws.next_scan(),
                                                 ntiles);
    HB_LAUNCHED();
}

} // namespace hb
