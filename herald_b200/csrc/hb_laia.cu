// Herald's "Laia" embedding scheduler: which worker trains which sample of
// the next global batch, and which cached rows each worker must push first.
//
// Host code, as in the reference (laia/src/laia_scheduler.cc; the older Cython version is
// python/hetu/laia/laia.pyx): every worker runs the same deterministic planner over simulated
// per-worker LRU caches ("snapshots", laia/include/mini_lru_cache.h) and keeps the part for its own
// rank.  Restated from scratch:
//   * MiniLru: the snapshot — open-addressing index + array-linked recency list + valid bit, no
//     allocation after construction (the reference: std::list + std::unordered_map per key);
//   * scoring (laia_scheduler.cc:194-231): samples are split over threads; a sample's score for
//     worker z is the number of its embeddings that are valid in z's snapshot;
//   * greedy assignment (:233-254): serial, sample by sample, highest score among the workers that
//     still have room, ties broken by the rotating order (j + batch_id) % W;
//   * communication plan (:256-270): worker w pushes the rows it holds valid that samples assigned
//     to OTHER workers will touch; emitted as ascending unique keys — exactly the `push_keys` the
//     cache's update_with_push_keys consumes (cache.cc:286-301);
//   * snapshot update (:146-161): plan keys are outdated, then the worker's own unique keys are
//     touched in ascending order; one thread per worker.
// The epoch/batch sequence, including the extra batch of the last epoch and the {0} terminator of
// the Python wire format (:126-135, :166-168), is reproduced by hb_laia_next.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "hb_common.cuh"

namespace hb {
namespace {

class MiniLru {
  public:
    explicit MiniLru(size_t capacity) : cap_(capacity) {
        size_t hs = 16;
        while (hs < 2 * (capacity + 2))
            hs <<= 1;
        mask_ = hs - 1;
        table_.assign(hs, kEmpty);
        const size_t nodes = capacity + 2; // one transient node beyond the capacity
        tagged_ = nodes < 0x00ffffffu;     // (0xff tag | 0xffffff node) would read as kEmpty
        node_.resize(nodes);
        free_.reserve(nodes);
        for (size_t i = nodes; i-- > 0;)
            free_.push_back((u32)i);
        head_ = tail_ = kNil;
    }

    bool check(u64 key) const { // mini_lru_cache.h:55-63
        return check(key, hash(key));
    }
    // the same with the key's hash supplied (scoring probes W snapshots with one key)
    bool check(u64 key, size_t hashed) const {
        const u32 n = find(key, hashed);
        return n != kNil && node_[n].valid;
    }
    static size_t hash_of(u64 key) {
        return hash(key);
    }
    void prefetch_entry_h(size_t hashed) const {
        __builtin_prefetch(&table_[hashed & mask_]);
    }
    void prefetch_node_h(size_t hashed) const {
        const u32 e = table_[hashed & mask_];
        if (e != kEmpty && (!tagged_ || (e & 0xff000000u) == tag_of(hashed)))
            __builtin_prefetch(&node_[node_of(e)]);
    }

    // -1 hit, -2 stale hit, 0 miss, 1 miss that evicted a valid line (mini_lru_cache.h:69-105)
    int get(u64 key) {
        const u32 n = find(key);
        if (n == kNil)
            return insert(key);
        const int res = node_[n].valid ? -1 : -2;
        unlink(n);
        push_front(n);
        node_[n].valid = 1;
        return res;
    }

    int insert(u64 key) { // key must be absent
        const u32 n = free_.back();
        free_.pop_back();
        node_[n].key = key;
        node_[n].valid = 1;
        push_front(n);
        index_insert(key, n);
        size_++;
        if (size_ > cap_) {
            const u32 v = tail_;
            const int res = node_[v].valid ? 1 : 0;
            unlink(v);
            index_erase(node_[v].key);
            free_.push_back(v);
            size_--;
            return res;
        }
        return 0;
    }

    void outdate(u64 key) { // mini_lru_cache.h:118-125
        const u32 n = find(key);
        if (n != kNil)
            node_[n].valid = 0;
    }

    void evict(u64 key) { // mini_lru_cache.h:107-116
        const u32 n = find(key);
        if (n == kNil)
            return;
        unlink(n);
        index_erase(key);
        free_.push_back(n);
        size_--;
    }

    void valid_keys(std::vector<u64> &out) const { // get_keys(): valid keys, ascending
        out.clear();
        for (u32 n = head_; n != kNil; n = node_[n].next)
            if (node_[n].valid)
                out.push_back(node_[n].key);
        std::sort(out.begin(), out.end());
    }

    // A snapshot of a 3.4 M-line cache does not fit any CPU cache and every operation is a chain
    // of dependent misses (index entry -> node -> list neighbours).  The planner knows the keys it
    // is about to touch, so it runs these three a few keys ahead of the operation itself:
    void prefetch_entry(u64 key) const { // the index entry of the key's home position
        __builtin_prefetch(&table_[hash(key) & mask_]);
    }
    void prefetch_node(u64 key) const { // needs the index entry: the node it points to
        prefetch_node_h(hash(key));
    }
    void prefetch_links(u64 key) const { // needs the node: its neighbours in the recency list
        const size_t hashed = hash(key);
        const u32 e = table_[hashed & mask_];
        if (e == kEmpty || (tagged_ && (e & 0xff000000u) != tag_of(hashed)))
            return;
        const u32 n = node_of(e);
        const u32 p = node_[n].prev, x = node_[n].next;
        if (p != kNil)
            __builtin_prefetch(&node_[p]);
        if (x != kNil)
            __builtin_prefetch(&node_[x]);
    }

  private:
    static constexpr u32 kNil = 0xffffffffu, kEmpty = 0xffffffffu;
    struct Node { // one cache line touch per node
        u64 key;
        u32 prev, next;
        u32 valid;
        u32 pad;
    };
    static size_t hash(u64 k) {
        k ^= k >> 33;
        k *= 0xff51afd7ed558ccdull;
        k ^= k >> 33;
        return (size_t)k;
    }
    // An index entry is (tag << 24 | node): 8 bits of the key's hash next to the node number, so a
    // probe that lands on another key's entry is rejected without touching that key's node (a
    // cache miss in a 3.4 M-line snapshot).  Needs node numbers below 2^24 - 1; larger snapshots
    // use untagged entries (tag_shift_ = 32: tag and node mask degenerate).
    u32 tag_of(size_t hashed) const {
        return tagged_ ? (u32)((hashed >> 40) & 0xffu) << 24 : 0u;
    }
    u32 node_of(u32 entry) const {
        return tagged_ ? (entry & 0x00ffffffu) : entry;
    }
    u32 find(u64 key) const {
        return find(key, hash(key));
    }
    u32 find(u64 key, size_t hashed) const {
        const u32 tag = tag_of(hashed);
        for (size_t h = hashed & mask_;; h = (h + 1) & mask_) {
            const u32 e = table_[h];
            if (e == kEmpty)
                return kNil;
            if (tagged_ && (e & 0xff000000u) != tag)
                continue;
            const u32 n = node_of(e);
            if (node_[n].key == key)
                return n;
        }
    }
    void index_insert(u64 key, u32 n) {
        const size_t hashed = hash(key);
        size_t h = hashed & mask_;
        while (table_[h] != kEmpty)
            h = (h + 1) & mask_;
        table_[h] = tag_of(hashed) | n;
    }
    void index_erase(u64 key) { // linear probing with backward shift: no tombstones
        size_t h = hash(key) & mask_;
        while (node_[node_of(table_[h])].key != key)
            h = (h + 1) & mask_;
        size_t hole = h;
        for (size_t j = (h + 1) & mask_;; j = (j + 1) & mask_) {
            const u32 e = table_[j];
            if (e == kEmpty)
                break;
            const size_t home = hash(node_[node_of(e)].key) & mask_;
            // e may move into the hole if its home is not in the (cyclic) interval (hole, j]
            const bool between = hole <= j ? (home > hole && home <= j) : (home > hole || home <= j);
            if (!between) {
                table_[hole] = e;
                hole = j;
            }
        }
        table_[hole] = kEmpty;
    }
    void unlink(u32 n) {
        const u32 p = node_[n].prev, x = node_[n].next;
        if (p != kNil)
            node_[p].next = x;
        else
            head_ = x;
        if (x != kNil)
            node_[x].prev = p;
        else
            tail_ = p;
    }
    void push_front(u32 n) {
        node_[n].prev = kNil;
        node_[n].next = head_;
        if (head_ != kNil)
            node_[head_].prev = n;
        head_ = n;
        if (tail_ == kNil)
            tail_ = n;
    }

    size_t cap_, mask_ = 0, size_ = 0;
    bool tagged_ = false;
    std::vector<u32> table_, free_;
    std::vector<Node> node_;
    u32 head_, tail_;
};

// v := ascending unique values of v.  LSD radix sort, 11-bit digits over the significant bits
// (embedding ids: 3 passes for a 33.7 M-row table), then one scan — the plan lists hold every
// occurrence of every shared row (~0.5 M entries, mostly repeats of hot rows) and a comparison sort
// of them was 40 % of the planner.
void radix_sort_unique(std::vector<u64> &v, std::vector<u64> &tmp) {
    const size_t n = v.size();
    if (n < 2)
        return;
    if (n < 2048) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        return;
    }
    u64 all = 0;
    for (u64 x : v)
        all |= x;
    int bits = 0;
    while (bits < 64 && (all >> bits) != 0)
        bits++;
    constexpr int RB = 11;
    constexpr size_t R = (size_t)1 << RB;
    tmp.resize(n);
    u64 *src = v.data(), *dst = tmp.data();
    size_t count[R];
    for (int shift = 0; shift < bits; shift += RB) {
        std::memset(count, 0, sizeof(count));
        for (size_t i = 0; i < n; i++)
            count[(src[i] >> shift) & (R - 1)]++;
        size_t run = 0;
        for (size_t d = 0; d < R; d++) {
            const size_t c = count[d];
            count[d] = run;
            run += c;
        }
        for (size_t i = 0; i < n; i++)
            dst[count[(src[i] >> shift) & (R - 1)]++] = src[i];
        std::swap(src, dst);
    }
    if (src != v.data())
        std::memcpy(v.data(), src, n * sizeof(u64));
    v.erase(std::unique(v.begin(), v.end()), v.end());
}

// out := the ids whose bit is set, ascending; the bitmap is all zero again afterwards
void drain_bitmap(std::vector<u64> &bm, std::vector<u64> &out) {
    out.clear();
    for (size_t wi = 0; wi < bm.size(); wi++) {
        u64 w = bm[wi];
        if (!w)
            continue;
        bm[wi] = 0;
        while (w) {
            out.push_back(((u64)wi << 6) | (u64)__builtin_ctzll(w));
            w &= w - 1;
        }
    }
}

// f(i) for i in [0, n) with the three prefetch stages of `lru` running 24 / 16 / 8 keys ahead
template <class F>
void for_keys_prefetched(const MiniLru &lru, const std::vector<u64> &keys, F f) {
    constexpr size_t kA = 24, kB = 16, kC = 8;
    const size_t n = keys.size();
    for (size_t i = 0; i < std::min(n, kA); i++)
        lru.prefetch_entry(keys[i]);
    for (size_t i = 0; i < std::min(n, kB); i++)
        lru.prefetch_node(keys[i]);
    for (size_t i = 0; i < std::min(n, kC); i++)
        lru.prefetch_links(keys[i]);
    for (size_t i = 0; i < n; i++) {
        if (i + kA < n)
            lru.prefetch_entry(keys[i + kA]);
        if (i + kB < n)
            lru.prefetch_node(keys[i + kB]);
        if (i + kC < n)
            lru.prefetch_links(keys[i + kC]);
        f(keys[i]);
    }
}

// run f(t) for t in [0, n) on up to `threads` std::threads (the calling thread takes part)
template <class F>
void parallel_for(size_t n, size_t threads, F f) {
    threads = std::max<size_t>(1, std::min(threads, n));
    if (threads == 1) {
        for (size_t t = 0; t < n; t++)
            f(t);
        return;
    }
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (size_t t = next.fetch_add(1); t < n; t = next.fetch_add(1))
            f(t);
    };
    std::vector<std::thread> pool;
    for (size_t i = 1; i < threads; i++)
        pool.emplace_back(work);
    work();
    for (auto &th : pool)
        th.join();
}

} // namespace
} // namespace hb

using namespace hb;

constexpr u64 kBitmapIds = 1ull << 31; // 256 MB of bits per worker at most; larger ids take the radix sort

struct hb_laia {
    std::vector<u64> embs; // [num_sample][num_table]
    size_t num_sample = 0, num_table = 0, mini = 0, W = 0, rank = 0, batch_size = 0;
    size_t epoch_num = 0, batch_num = 0, epoch_id = 0, batch_id = 0, threads = 1;
    bool in_epoch = false, finished = false;
    std::vector<MiniLru> snaps;
    // the plan of the most recent batch, all workers
    std::vector<std::vector<u64>> plans; // ascending unique keys
    std::vector<u64> dist;               // [W][mini] sample positions
    // scratch
    std::vector<u32> scores;             // [batch][W]
    std::vector<u32> assigned;           // [batch] worker of each sample
    std::vector<u64> holders;            // [batch][T] bit z: the embedding is valid in worker z's snapshot
                                         // (W <= 64; the reference's sample_emb_dep_, as a bit set)
    // TopkScheduler mode (laia/src/topk_scheduler.cc): scores count only the first `top_k` tables of a
    // dataset-specific order; samples and worker slots are split over `parts` logical threads, each
    // with its own workload counters (the reference's num_thread_: the result depends on it)
    bool topk = false;
    size_t top_k = 0, parts = 1;
    std::vector<u32> table_order;
    u64 max_key = 0;                     // largest embedding id of the sample set
    std::vector<std::vector<u64>> seen;  // per worker: one bit per id (ids below kBitmapIds), all zero
                                         // between uses — ascending unique keys without a sort
};

namespace {

bool laia_advance(hb_laia *s) { // laia_scheduler.cc:126-135, 166
    while (true) {
        if (!s->in_epoch) {
            if (s->epoch_id >= s->epoch_num)
                return false;
            s->epoch_id++;
            s->batch_id = 0;
            if (s->epoch_id == s->epoch_num)
                s->batch_num += 1; // one more allocation for the cache prefetch of the last epoch
            s->in_epoch = true;
        }
        if (s->batch_id < s->batch_num)
            return true;
        s->in_epoch = false;
    }
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void laia_plan_batch(hb_laia *s) {
    static const bool timing = getenv("HERALD_LAIA_TIMING") != nullptr; // per-phase ms on stderr
    const double t0 = timing ? now_ms() : 0;
    const size_t W = s->W, B = s->batch_size, T = s->num_table, S = s->num_sample;
    const size_t start = (s->batch_id * B) % S;
    auto pos_of = [&](size_t i) { return (start + i) % S; };
    // scoring: chunks of samples over the threads
    const bool use_bits = W <= 64;
    const size_t chunks = std::min<size_t>(B, s->threads * 4);
    parallel_for(chunks, s->threads, [&](size_t c) {
        const size_t lo = B * c / chunks, hi = B * (c + 1) / chunks;
        // the index entries of sample i + 2 and the nodes of sample i + 1 are prefetched while
        // sample i is scored: W x T dependent-miss chains per sample otherwise
        constexpr size_t kMemo = 4096;
        struct Memo {
            u64 key, bits;
        };
        std::vector<Memo> memo;
        if (use_bits)
            memo.assign(kMemo, Memo{~0ull, 0});
        auto prefetch = [&](size_t i, bool nodes) {
            const u64 *e = &s->embs[pos_of(i) * T];
            for (size_t j = 0; j < T; j++) {
                const size_t hk = MiniLru::hash_of(e[j]);
                if (use_bits && memo[hk & (kMemo - 1)].key == e[j])
                    continue; // a repeat: answered from the memo
                for (size_t z = 0; z < W; z++) {
                    if (nodes)
                        s->snaps[z].prefetch_node_h(hk);
                    else
                        s->snaps[z].prefetch_entry_h(hk);
                }
            }
        };
        // (Zipf ids repeat: the small direct-mapped memo id -> holder bits declared above answers the
        // repeats of this chunk without probing W snapshots again; they do not change during scoring)
        if (lo < hi)
            prefetch(lo, false);
        if (lo + 1 < hi)
            prefetch(lo + 1, false);
        if (lo < hi)
            prefetch(lo, true);
        for (size_t i = lo; i < hi; i++) {
            if (i + 2 < hi)
                prefetch(i + 2, false);
            if (i + 1 < hi)
                prefetch(i + 1, true);
            const u64 *e = &s->embs[pos_of(i) * T];
            u32 *sc = &s->scores[i * W];
            for (size_t z = 0; z < W; z++)
                sc[z] = 0;
            for (size_t j = 0; j < T; j++) {
                const size_t hk = MiniLru::hash_of(e[j]);
                if (use_bits) {
                    Memo &m = memo[hk & (kMemo - 1)];
                    if (m.key != e[j]) {
                        u64 bits = 0;
                        for (size_t z = 0; z < W; z++)
                            if (s->snaps[z].check(e[j], hk))
                                bits |= 1ull << z;
                        m.key = e[j];
                        m.bits = bits;
                    }
                    s->holders[i * T + j] = m.bits;
                    for (u64 b = m.bits; b; b &= b - 1)
                        sc[__builtin_ctzll(b)]++;
                } else {
                    for (size_t z = 0; z < W; z++)
                        if (s->snaps[z].check(e[j], hk))
                            sc[z]++;
                }
            }
        }
    });
    const double t1 = timing ? now_ms() : 0;
    // greedy assignment (serial: every choice depends on the workloads so far)
    std::vector<size_t> workload(W, 0);
    for (size_t i = 0; i < B; i++) {
        long best = -1;
        size_t best_w = 0;
        for (size_t j = 0; j < W; j++) {
            const size_t w = (j + s->batch_id) % W;
            const long score = (long)s->scores[i * W + w];
            if (workload[w] < s->mini && best < score) {
                best = score;
                best_w = w;
            }
        }
        s->dist[best_w * s->mini + workload[best_w]] = pos_of(i);
        s->assigned[i] = (u32)best_w;
        workload[best_w]++;
    }
    const double t2 = timing ? now_ms() : 0;
    // communication plan + snapshot update, one worker per thread
    double sub[5] = {0, 0, 0, 0, 0}; // worker 0's share of the phase (timing only)
    parallel_for(W, s->threads, [&](size_t w) {
        const bool tw = timing && w == 0;
        double c0 = tw ? now_ms() : 0;
        std::vector<u64> &plan = s->plans[w];
        plan.clear();
        MiniLru &snap = s->snaps[w];
        std::vector<u64> &bm = s->seen[w];
        const bool bitmap = !bm.empty();
        auto add = [&](u64 k) {
            if (bitmap)
                bm[k >> 6] |= 1ull << (k & 63);
            else
                plan.push_back(k);
        };
        for (size_t i = 0; i < B; i++) {
            if (s->assigned[i] == w)
                continue;
            const u64 *e = &s->embs[pos_of(i) * T];
            if (use_bits) { // what scoring saw (the snapshots have not changed since)
                const u64 *h = &s->holders[i * T];
                for (size_t j = 0; j < T; j++)
                    if ((h[j] >> w) & 1ull)
                        add(e[j]);
            } else {
                for (size_t j = 0; j < T; j++)
                    if (snap.check(e[j]))
                        add(e[j]);
            }
        }
        if (tw) { sub[0] = now_ms() - c0; c0 = now_ms(); }
        std::vector<u64> scratch;
        if (bitmap)
            drain_bitmap(bm, plan);
        else
            radix_sort_unique(plan, scratch);
        if (tw) { sub[1] = now_ms() - c0; c0 = now_ms(); }
        for_keys_prefetched(snap, plan, [&](u64 k) { snap.outdate(k); });
        if (tw) { sub[2] = now_ms() - c0; c0 = now_ms(); }
        std::vector<u64> uniq;
        uniq.reserve(s->mini * T);
        for (size_t j = 0; j < s->mini; j++) {
            const u64 *e = &s->embs[s->dist[w * s->mini + j] * T];
            if (bitmap)
                for (size_t t = 0; t < T; t++)
                    bm[e[t] >> 6] |= 1ull << (e[t] & 63);
            else
                uniq.insert(uniq.end(), e, e + T);
        }
        if (bitmap)
            drain_bitmap(bm, uniq);
        else
            radix_sort_unique(uniq, scratch);
        if (tw) { sub[3] = now_ms() - c0; c0 = now_ms(); }
        for_keys_prefetched(snap, uniq, [&](u64 k) { snap.get(k); });
        if (tw) sub[4] = now_ms() - c0;
    });
    if (timing)
        fprintf(stderr,
                "laia batch %zu: score %.2f ms, assign %.2f ms, plan + snapshots %.2f ms (worker 0: collect %.2f, "
                "sort %.2f, outdate %.2f, own keys %.2f, touch %.2f)\n",
                s->batch_id, t1 - t0, t2 - t1, now_ms() - t2, sub[0], sub[1], sub[2], sub[3], sub[4]);
    s->batch_id++;
}

// ---- TopkScheduler::get_dist + the snapshot update of TopkScheduler::launch ----------------------
// laia/src/topk_scheduler.cc:362-502 and :291-318.  Restated, not translated:
//  * scoring: a sample's score for worker z = how many of its first top_k tables (in the dataset's
//    pre-profiled order) hold an id that is valid in z's snapshot; the candidate is the worker that
//    FIRST reached the final maximum while the tables were walked in that order (:413-424);
//  * assignment: logical thread t owns samples [lo, hi) and, of every worker's mini-batch, the slots
//    [wstart, wend); it walks the workers from the candidate on and takes the best-scoring one that
//    still has room in its slots, stopping at once if that is the candidate itself (:430-452);
//  * plan of worker i: the sorted unique ids of its samples, walked cell by cell with an erase of
//    every id that is not valid in i's snapshot — the reference erases from the flat_set INSIDE the
//    range-for over it (:478-482), so with pointer iterators and a cached end the cell after an
//    erased one is never examined and the stale copies at the tail are; that is part of the result
//    (an empty snapshot keeps every other id) and is reproduced here cell by cell;
//  * snapshots: the plan's ids are outdated, then every unique id of the worker's samples is
//    touched in ascending order (:297-311).
void topk_plan_batch(hb_laia *s) {
    const size_t W = s->W, B = s->batch_size, T = s->num_table, S = s->num_sample, mini = s->mini;
    const size_t P = s->parts;
    const size_t start = (s->batch_id * B) % S;
    auto pos_of = [&](size_t i) { return (start + i) % S; };
    std::fill(s->dist.begin(), s->dist.end(), 0); // dist.reset(0)
    parallel_for(P, s->threads, [&](size_t t) {
        const size_t x = B / P, y = B % P;
        const size_t lo = t == 0 ? 0 : y + t * x, hi = t == 0 ? x + y : lo + x;
        const size_t wx = mini / P, wy = mini % P;
        const size_t wstart = t == 0 ? 0 : wy + t * wx, wend = t == 0 ? wx + wy : wstart + wx;
        const size_t room = wend - wstart;
        std::vector<size_t> load(W, 0);
        std::vector<u64> score(W);
        for (size_t i = lo; i < hi; i++) {
            const u64 *e = &s->embs[pos_of(i) * T];
            std::fill(score.begin(), score.end(), 0);
            u64 top = 0;
            size_t candidate = 0;
            for (size_t k = 0; k < s->top_k; k++) {
                const u64 emb = e[s->table_order[k]];
                const size_t hk = MiniLru::hash_of(emb);
                for (size_t z = 0; z < W; z++)
                    if (s->snaps[z].check(emb, hk) && ++score[z] > top) {
                        top = score[z];
                        candidate = z;
                    }
            }
            long best = -1;
            size_t best_w = W;
            for (size_t j = 0; j < W; j++) {
                const size_t w = (j + candidate) % W;
                if (best < (long)score[w] && load[w] < room) {
                    best = (long)score[w];
                    best_w = w;
                    if (w == candidate)
                        break;
                }
            }
            // (create() guarantees room for every sample: the reference indexes dist[-1] otherwise)
            s->dist[best_w * mini + wstart + load[best_w]] = pos_of(i);
            s->assigned[i] = (u32)best_w;
            load[best_w]++;
        }
    });
    parallel_for(W, s->threads, [&](size_t w) {
        MiniLru &snap = s->snaps[w];
        std::vector<u64> scratch, cells;
        for (size_t i = 0; i < B; i++)
            if (s->assigned[i] == w) {
                const u64 *e = &s->embs[pos_of(i) * T];
                cells.insert(cells.end(), e, e + T);
            }
        radix_sort_unique(cells, scratch);
        std::vector<u64> uniq = cells; // the worker's unique ids, for the snapshot update below
        // the erase-while-iterating walk over the ORIGINAL cells
        size_t live = cells.size();
        const size_t n0 = cells.size();
        for (size_t p = 0; p < n0; p++) {
            const u64 key = cells[p];
            if (snap.check(key))
                continue;
            u64 *q = std::lower_bound(cells.data(), cells.data() + live, key);
            if (q != cells.data() + live && *q == key) {
                std::memmove(q, q + 1, (size_t)(cells.data() + live - (q + 1)) * sizeof(u64));
                live--;
            }
        }
        cells.resize(live);
        s->plans[w].swap(cells);
        for (u64 k : s->plans[w])
            snap.outdate(k);
        for_keys_prefetched(snap, uniq, [&](u64 k) { snap.get(k); });
    });
    s->batch_id++;
}

// pre-profiled table orders of the reference (topk_scheduler.cc:150-165)
bool topk_table_order(const std::string &dataset, std::vector<u32> &order) {
    if (dataset == "criteo")
        order = {9, 13, 22, 20, 12, 21, 17, 14, 24, 3, 5, 10, 16, 15, 19, 2, 4, 11, 7, 25, 23, 18, 8, 1, 0, 6};
    else if (dataset == "avazu")
        order = {1, 2, 4, 5, 15, 7, 6, 16, 12, 0, 17, 8, 14, 10, 9, 11, 13, 3};
    else if (dataset == "movie")
        order = {0, 1};
    else if (dataset == "criteosearch")
        order = {0, 11, 3, 4, 5, 14, 1, 6, 2, 13, 16, 9, 8, 10, 12, 7, 15};
    else
        return false;
    return true;
}

} // namespace

extern "C" {

int hb_laia_create_topk(hb_laia **out, const uint64_t *sample_embs, size_t num_sample, size_t num_table,
                        size_t epoch_num, size_t mini_batch_size, size_t batch_num, size_t nrank, size_t rank,
                        size_t cache_size, size_t num_threads, const char *dataset, size_t top_k_table) {
    HB_API_BEGIN();
    std::vector<u32> order;
    HB_CHECK(dataset && topk_table_order(dataset, order), "dataset not supported"); // topk_scheduler.cc:162-165
    for (u32 j : order)
        HB_CHECK(j < num_table, "the dataset's table order names a table the samples do not have");
    const size_t parts = std::max<size_t>(1, num_threads);
    // the reference writes outside dist[] when a logical thread has more samples than slots
    HB_CHECK(mini_batch_size % parts == 0,
             "mini_batch_size must be a multiple of num_threads (the reference's per-thread slot split)");
    HB_CHECK(hb_laia_create(out, sample_embs, num_sample, num_table, epoch_num, mini_batch_size, batch_num, nrank,
                            rank, cache_size, num_threads) == 0, "create failed");
    hb_laia *s = *out;
    s->topk = true;
    s->parts = parts;
    s->table_order = order;
    if (!top_k_table)
        top_k_table = num_table; // topk_scheduler.cc:121-123
    s->top_k = std::min(top_k_table, order.size());
    HB_API_END();
}

int hb_laia_create(hb_laia **out, const uint64_t *sample_embs, size_t num_sample, size_t num_table,
                   size_t epoch_num, size_t mini_batch_size, size_t batch_num, size_t nrank, size_t rank,
                   size_t cache_size, size_t num_threads) {
    HB_API_BEGIN();
    HB_CHECK(out && sample_embs, "null argument");
    HB_CHECK(num_sample > 0 && num_table > 0 && mini_batch_size > 0 && nrank > 0, "empty problem");
    HB_CHECK(rank < nrank, "rank out of range");
    HB_CHECK(mini_batch_size * nrank <= num_sample,
             "a global batch must not hold a sample twice (mini_batch_size * nrank <= num_sample)");
    auto *s = new hb_laia();
    s->embs.assign(sample_embs, sample_embs + num_sample * num_table);
    s->num_sample = num_sample;
    s->num_table = num_table;
    s->mini = mini_batch_size;
    s->W = nrank;
    s->rank = rank;
    s->batch_size = mini_batch_size * nrank;
    s->epoch_num = epoch_num;
    s->batch_num = batch_num;
    s->threads = std::max<size_t>(1, num_threads);
    s->snaps.reserve(nrank);
    for (size_t w = 0; w < nrank; w++)
        s->snaps.emplace_back(cache_size);
    s->plans.resize(nrank);
    s->dist.assign(nrank * mini_batch_size, 0);
    s->scores.assign(s->batch_size * nrank, 0);
    s->assigned.assign(s->batch_size, 0);
    if (nrank <= 64)
        s->holders.assign(s->batch_size * num_table, 0);
    for (u64 k : s->embs)
        s->max_key = std::max(s->max_key, k);
    s->seen.resize(nrank);
    if (s->max_key < kBitmapIds)
        for (auto &bm : s->seen)
            bm.assign((size_t)(s->max_key >> 6) + 1, 0);
    *out = s;
    HB_API_END();
}

int hb_laia_destroy(hb_laia *s) {
    delete s;
    return 0;
}

int hb_laia_next(hb_laia *s, int *done) {
    HB_API_BEGIN();
    HB_CHECK(s && done, "null argument");
    if (s->finished || !laia_advance(s)) {
        s->finished = true;
        *done = 1;
    } else {
        if (s->topk)
            topk_plan_batch(s);
        else
            laia_plan_batch(s);
        *done = 0;
    }
    HB_API_END();
}

int hb_laia_plan_size(hb_laia *s, size_t worker, size_t *n) {
    HB_API_BEGIN();
    HB_CHECK(s && n && worker < s->W, "bad argument");
    *n = s->plans[worker].size();
    HB_API_END();
}

int hb_laia_plan(hb_laia *s, size_t worker, uint64_t *keys, size_t cap) {
    HB_API_BEGIN();
    HB_CHECK(s && worker < s->W, "bad argument");
    HB_CHECK(cap >= s->plans[worker].size(), "plan buffer too small");
    if (!s->plans[worker].empty())
        std::memcpy(keys, s->plans[worker].data(), s->plans[worker].size() * sizeof(u64));
    HB_API_END();
}

int hb_laia_dist(hb_laia *s, size_t worker, uint64_t *sample_idx) {
    HB_API_BEGIN();
    HB_CHECK(s && sample_idx && worker < s->W, "bad argument");
    std::memcpy(sample_idx, &s->dist[worker * s->mini], s->mini * sizeof(u64));
    HB_API_END();
}

int hb_laia_snapshot_keys(hb_laia *s, size_t worker, uint64_t *keys, size_t cap, size_t *n) {
    HB_API_BEGIN();
    HB_CHECK(s && n && worker < s->W, "bad argument");
    std::vector<u64> k;
    s->snaps[worker].valid_keys(k);
    *n = k.size();
    if (keys) {
        HB_CHECK(cap >= k.size(), "key buffer too small");
        if (!k.empty())
            std::memcpy(keys, k.data(), k.size() * sizeof(u64));
    }
    HB_API_END();
}

/* ---- shared-memory message ring between the planning process and a local worker -----------------
 * laia/include/share_mem.h:39-160 + ring_buffer.h: one POSIX shared-memory object per local worker
 * ("laia_cache_<local rank>"), a single-producer single-consumer ring of uint64 words; a message is
 * its length followed by its words (share_mem.h:126-139).  Indices are monotone word counts. */
struct hb_shmring {
    std::string name;
    bool creator = false;
    size_t words = 0; // capacity of the data region, a power of two
    u64 *base = nullptr;
    size_t map_bytes = 0;
    std::atomic<u64> *rd = nullptr, *wr = nullptr;
};

int hb_shmring_open(hb_shmring **out, const char *name, int create, size_t data_bytes) {
    HB_API_BEGIN();
    HB_CHECK(out && name, "null argument");
    auto *r = new hb_shmring();
    r->name = std::string("/") + name;
    r->creator = create != 0;
    if (create)
        shm_unlink(r->name.c_str()); // a ring left behind by a crashed run must not be read as this one
    int fd = shm_open(r->name.c_str(), create ? (O_RDWR | O_CREAT | O_EXCL) : O_RDWR, 0644);
    if (fd < 0) {
        delete r;
        throw Error(std::string("shm_open ") + name + ": " + strerror(errno));
    }
    size_t words = 1024;
    if (create) {
        while (words * 8 < data_bytes)
            words <<= 1; // share_mem.h:60-66: rounded up to a power of two
        r->map_bytes = words * 8 + 64;
        if (ftruncate(fd, (off_t)r->map_bytes) != 0) {
            close(fd);
            delete r;
            throw Error(std::string("ftruncate: ") + strerror(errno));
        }
    } else {
        struct stat st;
        fstat(fd, &st);
        r->map_bytes = (size_t)st.st_size;
        HB_CHECK(r->map_bytes > 64, "shared ring not sized yet (the creating rank starts first)");
        words = (r->map_bytes - 64) / 8;
    }
    void *p = mmap(nullptr, r->map_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) {
        delete r;
        throw Error(std::string("mmap: ") + strerror(errno));
    }
    r->words = words;
    r->base = static_cast<u64 *>(p);
    r->rd = reinterpret_cast<std::atomic<u64> *>(r->base + words);
    r->wr = reinterpret_cast<std::atomic<u64> *>(r->base + words + 1);
    if (create) {
        r->rd->store(0);
        r->wr->store(0);
    }
    *out = r;
    HB_API_END();
}

int hb_shmring_close(hb_shmring *r) {
    if (r) {
        if (r->base)
            munmap(r->base, r->map_bytes);
        if (r->creator)
            shm_unlink(r->name.c_str());
        delete r;
    }
    return 0;
}

/* -> n on success, -1 when the ring has no room for the message (share_mem.h:126-139) */
int hb_shmring_send(hb_shmring *r, const uint64_t *data, size_t n, long long *sent) {
    HB_API_BEGIN();
    const u64 wr = r->wr->load(std::memory_order_relaxed), rd = r->rd->load(std::memory_order_acquire);
    if (r->words - (wr - rd) < n + 1) {
        *sent = -1;
    } else {
        const size_t mask = r->words - 1;
        r->base[wr & mask] = n;
        for (size_t i = 0; i < n; i++)
            r->base[(wr + 1 + i) & mask] = data[i];
        r->wr->store(wr + 1 + n, std::memory_order_release);
        *sent = (long long)n;
    }
    HB_API_END();
}

/* Length of the next message (-1: ring empty); with `data` non-null and cap >= length, also
 * consumes it. */
int hb_shmring_recv(hb_shmring *r, uint64_t *data, size_t cap, long long *n) {
    HB_API_BEGIN();
    const u64 rd = r->rd->load(std::memory_order_relaxed), wr = r->wr->load(std::memory_order_acquire);
    if (wr == rd) {
        *n = -1;
    } else {
        const size_t mask = r->words - 1;
        const u64 len = r->base[rd & mask];
        *n = (long long)len;
        if (data && cap >= len) {
            for (size_t i = 0; i < len; i++)
                data[i] = r->base[(rd + 1 + i) & mask];
            r->rd->store(rd + 1 + len, std::memory_order_release);
        }
    }
    HB_API_END();
}

int hb_shmring_used(hb_shmring *r, size_t *words) {
    HB_API_BEGIN();
    *words = (size_t)(r->wr->load(std::memory_order_acquire) - r->rd->load(std::memory_order_acquire));
    HB_API_END();
}

/* the snapshot cache on its own (a18: return codes of get) */
struct hb_minilru {
    MiniLru lru;
    explicit hb_minilru(size_t cap) : lru(cap) {}
};

int hb_minilru_create(hb_minilru **out, size_t capacity) {
    HB_API_BEGIN();
    HB_CHECK(out, "null argument");
    *out = new hb_minilru(capacity);
    HB_API_END();
}
int hb_minilru_destroy(hb_minilru *m) {
    delete m;
    return 0;
}
int hb_minilru_get(hb_minilru *m, uint64_t key) {
    return m->lru.get(key);
}
int hb_minilru_check(hb_minilru *m, uint64_t key) {
    return m->lru.check(key) ? 1 : 0;
}
int hb_minilru_outdate(hb_minilru *m, uint64_t key) {
    m->lru.outdate(key);
    return 0;
}
int hb_minilru_evict(hb_minilru *m, uint64_t key) {
    m->lru.evict(key);
    return 0;
}
int hb_minilru_keys(hb_minilru *m, uint64_t *keys, size_t cap, size_t *n) {
    HB_API_BEGIN();
    std::vector<u64> k;
    m->lru.valid_keys(k);
    *n = k.size();
    if (keys) {
        HB_CHECK(cap >= k.size(), "key buffer too small");
        if (!k.empty())
            std::memcpy(keys, k.data(), k.size() * sizeof(u64));
    }
    HB_API_END();
}

} // extern "C"
