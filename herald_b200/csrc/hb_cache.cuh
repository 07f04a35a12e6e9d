// Device-resident worker cache + owner table shard (types shared by hb_cache.cu / hb_comm.cu).
//
// Replaces, on one B200:
//   worker side  src/hetu_cache (LRU/LFU/LFUOpt + CacheBase)                      -> hb_cache
//   owner side   ps-lite CacheTable + PSFhandle_embedding.cc handlers             -> hb_table
// Both live in HBM; a "pull" or "push" between them is a row copy / row add inside one kernel.
#pragma once

#include <vector>

#include "hb_comm.cuh"
#include "hb_common.cuh"
#include "hb_sort.cuh"

struct hb_table {
    int node_id = 0;
    int device = 0;
    size_t length = 0, width = 0; // global table geometry
    size_t row_begin = 0, nrows = 0; // rows owned by this rank
    float *rows = nullptr;           // [nrows, width]
    hb::i64 *ver = nullptr;          // [nrows]  (ps-lite/include/ps/server/param.h:124)
    // every rank's shard mapped into this process (CUDA IPC over NVLink); [rank] = own shard
    float *peer_rows[hb::kMaxWorld] = {};
    hb::i64 *peer_ver[hb::kMaxWorld] = {};
    int rank = 0, world = 1;
};

namespace hb {

// ---- open-addressing index: key -> slot ------------------------------------------------
struct __align__(16) HtEntry {
    u64 key;
    u32 slot;
    u32 pad;
};
constexpr u64 HT_EMPTY = ~0ull;
constexpr u64 HT_TOMB = ~0ull - 1;

// slot states
enum : u8 {
    S_FREE = 0,
    S_CACHED = 1,    // in the index, evictable
    S_STORE = 2,     // in the index, LFUOpt permanent store (never evicted)
    S_TRANSIENT = 3, // line of the running call, not (yet) in the index
    S_PENDING = 4    // evicted while dirty: waits for the next push (CacheBase::evict_)
};
// slot flags
enum : u8 {
    F_GRAD = 1,    // Line::grad_ has been allocated (embedding.h:120-123): addup() adds it
    F_DATALESS = 2 // Line built with init_data=false by an update miss (cache.cc:146-151)
};

// replacement priority word: (min(use, kUseSat) << 52) | stamp, ~0 for slots that cannot be evicted
constexpr int kStampBits = 52;
constexpr u64 kStampMask = (1ull << kStampBits) - 1;
constexpr u32 kUseSat = 4095;
constexpr u64 PRIO_NONE = ~0ull;
constexpr int kSelBins = 4096; // radix of the victim selection
constexpr int kSelBits = 12;
constexpr u32 kLfuOptUseMax = 10; // src/hetu_cache/include/lfuopt_cache.h:25

// device error codes (hb_perf.error)
enum : u32 {
    E_NONE = 0,
    E_NO_FREE_SLOT = 1,  // transient + pending lines exceeded the slack of the row store
    E_INDEX_FULL = 2,    // open-addressing index could not place a key
    E_KEY_RANGE = 3,     // key >= table length
    E_EVICT_OVERFLOW = 4, // pending-eviction list overflow
    E_MAILBOX = 5,        // more pushed lines for one owner than its mailbox holds
    E_BARRIER = 6         // a peer did not reach the exchange barrier
};

// ---- multi-GPU: owner shards and push mailboxes of the whole group, mapped over NVLink ----
// Rank w's cache reads missing / stale rows straight out of the owner's shard (sync) and
// deposits the lines it pushes into a mailbox inside the owner's memory; after a barrier the
// owner applies the mailboxes in source-rank order (deterministic), see DESIGN.md section 6.
// A mailbox region (one per source rank, inside each owner) holds two sections, "batch" (the
// lines of the update call, one slot per unique key of the owner's range, upd == 0 = not
// pushed) and "flush" (dirty victims), each {key[cap] (row inside the shard), upd[cap],
// grad[cap, D]}, behind a 16-byte header {cnt_batch, cnt_flush}.
struct PeerView {
    int world, rank;
    u64 per, rem;                 // AveragePartitioner: len / G rows each, the first len % G one more
    const float *rows[kMaxWorld]; // owner shards (rows[rank] is local)
    const i64 *ver[kMaxWorld];
    char *out[kMaxWorld]; // this rank's region inside owner o
    char *in;             // this rank's mailboxes as an owner: world regions, indexed by source
    size_t region_bytes;
    u32 cap;       // slots per section
    u32 *lo;       // [world + 1] first unique index of each owner's key range (per call)
    u32 *fl_count; // [world] flush slots taken per owner (per call)
    // Exchange control block of THIS cache (one per rank, in front of its mailboxes, mapped into
    // every peer): ready[src] = last exchange whose mailbox contents source `src` has completed
    // here; applied[owner] = last exchange owner `owner` has applied to its shard.  Epochs count
    // the exchanges of this cache, so several caches never share a flag.
    u64 *ctrl;               // local: ready[kMaxWorld], applied[kMaxWorld], then 2 u32 scratch counters
    u64 *ctrl_peer[kMaxWorld];
    u32 *head; // [nrows_local] newest mailbox entry (+1) of each local row during an apply, 0 otherwise
    u32 *next; // [2 * world * cap] entry -> the entry linked before it (+1), 0 = first, ~0 = not pushed
    u64 timeout_ns; // a peer that does not show up within this fails the call (E_BARRIER)
};
constexpr size_t kMailboxCtrlBytes = 4096;

struct MailboxSection {
    u64 *key;
    i32 *upd;
    float *grad;
};

__host__ __device__ inline size_t align16(size_t x) {
    return (x + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t mailbox_section_bytes(size_t cap, size_t D) {
    return align16(cap * 8) + align16(cap * 4) + align16(cap * D * 4);
}
__host__ __device__ inline size_t mailbox_region_bytes(size_t cap, size_t D) {
    return 16 + 2 * mailbox_section_bytes(cap, D);
}
__host__ __device__ inline MailboxSection mailbox_section(char *region, int section, size_t cap,
                                                          size_t D) {
    char *p = region + 16 + (size_t)section * mailbox_section_bytes(cap, D);
    MailboxSection m;
    m.key = reinterpret_cast<u64 *>(p);
    m.upd = reinterpret_cast<i32 *>(p + align16(cap * 8));
    m.grad = reinterpret_cast<float *>(p + align16(cap * 8) + align16(cap * 4));
    return m;
}

// Mutable scalars of one cache, resident in device memory (one cache line or two).
struct CacheRegs {
    // persistent
    u32 size;        // lines in the index (evictable + store)
    u32 store_size;  // LFUOpt permanent lines
    u32 free_top;    // height of the free-slot stack
    u32 pending;     // dirty victims awaiting a push
    u32 ht_occupied; // non-EMPTY index entries (live + tombstones)
    u32 error;
    u32 slot_hw;     // slots [0, slot_hw) have been handed out at least once (scan bound)
    u32 pad0;
    u64 clock; // replacement clock: one tick per policy touch / insert
    u64 floor; // lower bound of the stamps of the policy's victim class
    // per call (zeroed by op_begin)
    u64 clock0;     // clock at call start
    u32 U;          // unique keys of the batch being resolved
    u32 M;          // misses among them
    u32 alloc_base; // (unused)
    u32 alloc_top0, alloc_top1; // free-stack height at the start of the call / after its first resolve:
                                // miss j of a resolve takes free_stack[top - 1 - j]
    u32 pulled;     // rows transferred by the sync
    u32 pushed;     // lines pushed by the update
    u32 flushed;    // pending victims flushed by the update
    // insert planning
    u32 E;         // evictions required by the policy
    u32 k_old;     // of which: resident lines selected by stamp
    u32 n_drop;    // new lines that do not survive the call (first n_drop misses)
    u32 need_min;  // LFU corner: evict the global minimum (use, stamp) line
    u32 nv;        // victims collected so far
    u32 nc;        // boundary candidates collected
    u32 sel_shift; // bin shift of the first selection level
    u32 sel_bin;   // threshold bin
    u32 sel_rem;   // victims still to take from the threshold bin
    u32 min_use;   // scratch of the global-minimum search
    u64 min_prio;
    u64 ins_clock0; // stamps of inserted lines start here
    // second batch of a push_pull call
    u32 U2, M2, alloc_base2;
    u32 sel_done; // blocks of sel_hist_kernel that have added their bins (zeroed by plan_insert)
    // windowed first level of the victim selection (see sel_hist_kernel)
    u32 sel_done2;    // same for the full-range fallback pass
    u32 sel_cut;      // persistent: bins the next call's window covers (0 = not known yet)
    u32 sel_cut_eff;  // window of the running call
    u32 sel_above;    // class lines at or above the window
    u32 sel_fallback; // the threshold was not inside the window: run the full-range pass
    u32 sel_use_log;  // LRU: this call selects its victims by walking the stamp log
    u64 sel_floor0;   // floor at the start of the selection (the log walk moves `floor` itself)
    // multi-GPU traffic of the call (zeroed by op_begin): rows pulled from / lines pushed to a PEER
    u32 pulled_remote, pushed_remote;
    u32 tail_done; // CTAs of update_tail_kernel that have finished (zeroed by op_begin)
    u32 xerror;    // failure raised by an exchange kernel; moved into the next call record
};

// Everything a kernel needs to address the cache (passed by value).
struct CacheView {
    // geometry
    u32 capacity; // slots in the row store (limit + slack)
    u32 limit;
    u32 width;
    u32 ht_mask;
    int policy;
    // slot arrays
    u64 *slot_key;
    i64 *slot_version;
    i32 *slot_updates;
    u64 *slot_prio;
    u32 *slot_use;
    u8 *slot_state;
    u8 *slot_flags;
    float *data;
    float *grad;
    HtEntry *ht;
    u32 *free_stack;
    u32 *pending_list;
    u32 *victims;    // [capacity]
    u64 *cand_prio;  // [2][capacity] boundary candidates (ping-pong)
    u32 *cand_slot;  // [2][capacity]
    u32 *sel_hist;   // [2][kSelBins]: windowed pass, full-range fallback
    u32 *stamp_log;  // [log_mask + 1]: slot that was given stamp t, at t & log_mask (see sel_log_kernel)
    u64 *sel_scan;   // ticket + tile status words of the log walk's grid scan
    u32 log_mask;
    CacheRegs *regs;
    // owner shard (same GPU)
    float *trows;
    i64 *tver;
    u64 row_begin, nrows_local, table_len;
    PeerView pv; // pv.world == 1: single GPU, trows / tver are the whole table
};

#ifdef __CUDACC__
// owner of a key and its row inside the owner's shard (ps-lite/include/ps/partitioner.h:46-57)
__device__ __forceinline__ int owner_of(const PeerView &pv, u64 key, u64 &trow) {
    const u64 cut = pv.rem * (pv.per + 1);
    if (key < cut) {
        const u64 o = key / (pv.per + 1);
        trow = key - o * (pv.per + 1);
        return (int)o;
    }
    const u64 o = pv.per ? (key - cut) / pv.per : 0;
    trow = key - cut - o * pv.per;
    return (int)(pv.rem + o);
}
__device__ __forceinline__ u64 shard_begin(const PeerView &pv, int o) {
    const u64 uo = (u64)o;
    return uo < pv.rem ? uo * (pv.per + 1) : pv.rem * (pv.per + 1) + (uo - pv.rem) * pv.per;
}
#endif

// counters copied to pinned host memory at the end of every call
struct PerfRecord {
    u32 kind; // 0 pull, 1 push, 2 push_pull
    u32 num_all, num_unique, num_miss, num_evict, num_transfered;
    u32 size, error, ht_occupied, pending, limit_full;
    u32 num_remote; // Pull: rows read from a peer's shard; Push: lines deposited in a peer's mailbox
    u64 clock, floor; // replacement clock and victim-class floor after the call
};

} // namespace hb

struct hb_cache {
    int policy = 0;
    size_t limit = 0, length = 0, width = 0;
    int node_id = 0;
    int device = 0;
    hb::i64 pull_bound = 5, push_bound = 5;
    bool bypass = false;
    hb_table *table = nullptr;
    // Streams.  `stream` (main) orders every kernel that touches the cache; `side` runs the sort +
    // unique of a batch (it depends on the keys only) while the main stream still works on the
    // previous call; `h2d` / `d2h` move host callers' gradients in and gathered rows out, so that
    // an upload, a download and the kernels of a third call can overlap (PCIe is full duplex).
    cudaStream_t stream = nullptr, side = nullptr, h2d = nullptr, d2h = nullptr;
    cudaStream_t xstream = nullptr;             // multi-GPU: arrive / link / apply of an update's exchange
    cudaEvent_t ev_x_fork = nullptr, ev_x_done = nullptr;
    int x_pending = 0;                          // 1: forked by the running call, 2: by the previous one
    cudaStream_t side2 = nullptr;               // a lookup's insert phase, next to its sync + gather
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int cur = 0;                                // workspace that holds the most recent lookup batch
    cudaEvent_t ev_ws_free[2] = {nullptr, nullptr}; // (handle) main: the readers of ws[i] enqueued so far are done
    cudaEvent_t ev_ws_rel[2] = {nullptr, nullptr};  // owned events ev_ws_free may point at
    cudaEvent_t ev_last_end = nullptr;              // end event of the call enqueued last
    cudaEvent_t ev_sorted[2] = {nullptr, nullptr};  // side: ws[i] holds uniq / inverse / segments
    cudaEvent_t ev_up = nullptr;                // h2d: the gradients of the running update have arrived
    cudaEvent_t ev_grads_free = nullptr;        // main: the gradient staging buffer has been consumed
    cudaEvent_t ev_gathered[2] = {nullptr, nullptr}; // main: dest staging buffer k is complete
    cudaEvent_t ev_dl[2] = {nullptr, nullptr};       // d2h: download out of staging buffer k is done
    cudaEvent_t ev_producer = nullptr;          // a caller's stream: its device buffers are ready
    bool reduce_split = false;                  // two-level reduction of rows with > 1024 occurrences
    float grad_scale = 1.0f;                    // gradients are multiplied by this (the -lr fold)
    int dl_next = 0;                            // dest staging buffer of the next host-dest lookup
    int dl_of_call[1024] = {};                  // [kRing] 1 + staging buffer a call downloads from, 0 = none
    uint64_t dl_seq[2] = {0, 0};                // call whose download ev_dl[k] stands for (+1)
    uint64_t err_checked = 0;                   // calls whose error word has been reported
    hb::CacheView view{};
    size_t ht_size = 0;
    size_t slack = 0;
    hb::KeyWorkspace ws[2];
    hb::SortedKeys sorted[2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    // per-batch resolve results (sized with the workspaces)
    hb::i32 *uslot[2] = {nullptr, nullptr};
    hb::u32 *miss_list[2] = {nullptr, nullptr};
    size_t batch_cap = 0;
    // staging for host-memory callers
    void *keys_stage[2] = {nullptr, nullptr};
    size_t keys_stage_cap = 0;
    float *rows_stage[3] = {nullptr, nullptr, nullptr}; // [0], [1]: gathered rows (ping-pong); [2]: gradients
    size_t rows_stage_cap[3] = {0, 0, 0};
    void *push_keys_stage = nullptr;
    size_t push_keys_stage_cap = 0;
    void *score_scratch = nullptr; // hb_cache_score / hb_cache_probe staging
    size_t score_scratch_cap = 0;
    // perf ring (pinned host) + events
    hb::PerfRecord *ring = nullptr;      // [kRing] mapped pinned memory: op_end writes the record there
    hb::PerfRecord *ring_dev = nullptr;  // the same ring as the device sees it
    bool timed[1024] = {};               // [kRing] the call carries its begin event
    hb::PerfRecord *dev_record = nullptr;
    static constexpr int kRing = 1024;
    uint64_t calls = 0;
    std::vector<cudaEvent_t> ev_begin, ev_end;
    static constexpr int kPhases = 4;
    std::vector<cudaEvent_t> ev_phase; // [kRing][kPhases] phase boundaries (perf enabled only)
    std::vector<uint32_t> phase_mask;
    bool perf_phases = false;
    unsigned perf_every = 1; // phase events on every perf_every-th pair of calls
    // host-side upper bound of index occupancy (refreshed at wait)
    size_t occ_upper = 0;
    size_t incoming_ring[kRing] = {}; // keys each call could add to the index (per call of the ring)
    size_t ticks_ring[kRing] = {};    // most replacement-clock ticks each call could take
    size_t cur_ticks = 0;             // ... the call being enqueued
    size_t pending_upper = 0;
    // false while no call can have left a line dirty (every update so far pushed every line it touched:
    // plain updates with push_bound <= 0): the update's tail is then sized for its other work only (its
    // loops are grid-stride, the count of pending victims is read on the device either way)
    bool may_hold_dirty = false;
    int key_bits = 64;
    hb::u32 hot_threshold = 64; // segments longer than this take the column-split path
    // multi-GPU
    char *mailbox = nullptr;              // this rank's regions as an owner
    char *peer_mailbox[hb::kMaxWorld] = {}; // every owner's regions, mapped
    size_t mailbox_cap = 0;
    size_t xrows_upper = 0;   // most entries one mailbox section can hold (sizes the apply grids)
    uint64_t xepoch = 0;      // exchanges of THIS cache enqueued so far (same on every rank)
};
