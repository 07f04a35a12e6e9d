// Device-side sorted-unique machinery: stable LSD radix sort of (key, index) pairs, then a
// single-pass scan that emits ascending unique keys, the inverse map and the occurrence
// segments.  Replaces the host-side dedup of both reference variants:
//   Unique<T> (argsort + scan)           src/hetu_cache/include/unqiue_tools.h:9-48
//   np.unique(return_inverse=True)       python/hetu/ndarray.py:532-536
// Both produce ASCENDING unique ids; the stable sort additionally yields, for every unique
// key, its occurrences in ascending original index, which is the order the reference
// accumulates gradients in (src/hetu_cache/src/cache.cc:145-154).
#pragma once

#include <algorithm>

#include "hb_common.cuh"

namespace hb {

constexpr int kSortDirectTiles = 56; // passes of at most this many tiles sum their predecessors' counts directly
constexpr int kScanSlots = 12;  // independent grid-scan states per workspace
constexpr int kScanBlock = 256; // threads per scan tile
constexpr int kSortRadixBits = 9;
constexpr int kSortRadix = 1 << kSortRadixBits;
constexpr int kMaxSortPasses = 8; // 64-bit keys / 9 bits, rounded up

// Per-call scratch, sized for `cap` keys.  All device memory.
struct KeyWorkspace {
    size_t cap = 0;
    u64 *keys[2] = {nullptr, nullptr}; // ping-pong sort buffers (keys)
    u32 *vals[2] = {nullptr, nullptr}; // ping-pong sort buffers (original index)
    u64 *sort_status = nullptr;        // [nblk][RADIX] look-back status words (epoch-tagged)
    u32 *sort_counts = nullptr;        // [kSortDirectTiles][RADIX] per-tile digit counts (tag << 13 | count)
    size_t nblk_cap = 0;
    u32 sort_epoch = 0;                // one epoch per sort pass ever run on this workspace
    // results of unique_from_sorted
    u64 *uniq = nullptr;      // [cap]   ascending unique keys
    u32 *inverse = nullptr;   // [cap]   rank of keys[i] in uniq
    u32 *seg_start = nullptr; // [cap+1] first sorted position of each unique key
    u32 *num_unique = nullptr; // device scalar
    // segment-reduce work lists and work items (see hb_rows.cuh); the control words live behind the
    // main scan arena so that reset_main() zeroes them with the same memset
    u32 *hot_a = nullptr, *hot_b = nullptr, *medium = nullptr; // [cap] item indices
    void *seg_items = nullptr; // [cap + 32] x 32 B, ticket order
    float *split_partials = nullptr; // [split_rows_cap()][kSplitTiles][D] tile sums of the two-level reduction
    size_t split_partials_cap = 0;   // floats
    // main arena (zeroed by reset_main): kScanSlots x (ticket + status[ntile_cap]), then 4 control
    // words (work lists of the segment reduce)
    u64 *scan_arena = nullptr;
    // side arena (zeroed by reset_side): everything the sort + unique kernels count in — the unique
    // scan's slot, the sort's digit totals [kMaxSortPasses][RADIX], its per-pass tile tickets and
    // the same-keys mismatch counter.  Separate from the main arena because the cache runs sort +
    // unique on a side stream while the main stream still works on the previous batch.
    u64 *side_arena = nullptr;
    size_t ntile_cap = 0;
    int scan_next = 0;

    void reserve(size_t n);
    void release();
    size_t scan_slot_words() const {
        return ntile_cap + 1;
    }
    size_t arena_words() const {
        return (size_t)kScanSlots * scan_slot_words() + 4 + split_rows_cap() / 2 + 1;
    }
    // rows that can take the two-level reduction in one call: more than kVeryHot (1024) occurrences each
    size_t split_rows_cap() const {
        return cap / 1024 + 2;
    }
    u32 *split_done() const { // [split_rows_cap()] tiles finished per very hot row (zeroed by reset_main)
        return hot_ctrl() + 8;
    }
    size_t side_words() const {
        return scan_slot_words() + (kMaxSortPasses * kSortRadix) / 2 + kMaxSortPasses / 2 + 1;
    }
    u32 *hot_ctrl() const { // 8 x u32
        return reinterpret_cast<u32 *>(scan_arena + (size_t)kScanSlots * scan_slot_words());
    }
    u32 *sort_totals() const {
        return reinterpret_cast<u32 *>(side_arena + scan_slot_words());
    }
    u32 *sort_tickets() const {
        return sort_totals() + kMaxSortPasses * kSortRadix;
    }
    u32 *reuse_mismatch() const {
        return sort_tickets() + kMaxSortPasses;
    }
    // host-side record of what uniq / inverse / seg_start / the sorted buffers currently hold
    bool sorted_valid = false;
    size_t sorted_n = 0;
    u32 next_sort_epoch() {
        sort_epoch = (sort_epoch + 1) & 0x3fffffffu;
        if (sort_epoch == 0)
            sort_epoch = 1;
        return sort_epoch;
    }
    // The words of the main arena a call of `n` keys with at most `slots` grid scans can touch, as
    // ranges a kernel zeroes (the cache's op_begin does, instead of a memset node between kernels).
    struct ZeroRange {
        u64 *p;
        u32 words;
    };
    int main_ranges(size_t n, int slots, ZeroRange *out) {
        const u32 tiles = (u32)std::min<size_t>(ntile_cap, n / kScanBlock + 2);
        int k = 0;
        for (int s = 0; s < slots; s++)
            out[k++] = ZeroRange{scan_arena + (size_t)s * scan_slot_words(), tiles + 1};
        out[k++] = ZeroRange{scan_arena + (size_t)kScanSlots * scan_slot_words(),
                             (u32)(4 + std::min<size_t>(split_rows_cap(), n / 1024 + 2) / 2 + 1)};
        scan_next = 0;
        return k;
    }
    // zero the scan slots / the sort's counters — once per op, before the kernels that use them
    void reset_main(cudaStream_t st);
    void reset_side(cudaStream_t st);
    void reset_scans(cudaStream_t st) { // both, for single-stream callers
        reset_main(st);
        reset_side(st);
    }
    ScanState next_scan();
    ScanState side_scan() const; // the unique kernel's scan state
};

struct SortedKeys {
    const u64 *keys; // ascending
    const u32 *perm; // perm[p] = original index of sorted position p (stable)
};

// keys_in: n keys of `key_kind` (HB_KEYS_U64 / HB_KEYS_F32) in device memory.
// key_bits: keys are < 2^key_bits (fewer bits = fewer passes).
// `mismatch` (from check_same_keys, may be null): the kernels do nothing when *mismatch == 0,
// i.e. when the workspace already holds this very batch sorted.
SortedKeys radix_sort_keys(KeyWorkspace &ws, const void *keys_in, int key_kind, size_t n,
                           int key_bits, cudaStream_t st, const u32 *mismatch = nullptr);

// Hetu's BSP loop updates the batch it looked up one call earlier (python/hetu/gpu_ops/
// ParameterServerCommunicate.py:48-52), so the second sort of a batch can be skipped.  Enqueues
// an exact device-side comparison of the new keys with the batch the workspace holds and returns
// the device counter the sort kernels test; null when there is nothing to compare with.
const u32 *check_same_keys(KeyWorkspace &ws, const void *keys_in, int key_kind, size_t n,
                           cudaStream_t st);

// Fills ws.uniq / ws.inverse / ws.seg_start / ws.num_unique from a sorted sequence.
void unique_from_sorted(KeyWorkspace &ws, const SortedKeys &sk, size_t n, cudaStream_t st,
                        const u32 *mismatch = nullptr);

int bits_for(u64 max_key_exclusive);

} // namespace hb
