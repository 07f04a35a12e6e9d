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
        key_.resize(nodes);
        prev_.resize(nodes);
        next_.resize(nodes);
        valid_.resize(nodes);
        free_.reserve(nodes);
        for (size_t i = nodes; i-- > 0;)
            free_.push_back((u32)i);
        head_ = tail_ = kNil;
    }

    bool check(u64 key) const { // mini_lru_cache.h:55-63
        const u32 n = find(key);
        return n != kNil && valid_[n];
    }

    // -1 hit, -2 stale hit, 0 miss, 1 miss that evicted a valid line (mini_lru_cache.h:69-105)
    int get(u64 key) {
        const u32 n = find(key);
        if (n == kNil)
            return insert(key);
        const int res = valid_[n] ? -1 : -2;
        unlink(n);
        push_front(n);
        valid_[n] = 1;
        return res;
    }

    int insert(u64 key) { // key must be absent
        const u32 n = free_.back();
        free_.pop_back();
        key_[n] = key;
        valid_[n] = 1;
        push_front(n);
        index_insert(key, n);
        size_++;
        if (size_ > cap_) {
            const u32 v = tail_;
            const int res = valid_[v] ? 1 : 0;
            unlink(v);
            index_erase(key_[v]);
            free_.push_back(v);
            size_--;
            return res;
        }
        return 0;
    }

    void outdate(u64 key) { // mini_lru_cache.h:118-125
        const u32 n = find(key);
        if (n != kNil)
            valid_[n] = 0;
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
        for (u32 n = head_; n != kNil; n = next_[n])
            if (valid_[n])
                out.push_back(key_[n]);
        std::sort(out.begin(), out.end());
    }

  private:
    static constexpr u32 kNil = 0xffffffffu, kEmpty = 0xffffffffu;
    static size_t hash(u64 k) {
        k ^= k >> 33;
        k *= 0xff51afd7ed558ccdull;
        k ^= k >> 33;
        return (size_t)k;
    }
    u32 find(u64 key) const {
        for (size_t h = hash(key) & mask_;; h = (h + 1) & mask_) {
            const u32 n = table_[h];
            if (n == kEmpty)
                return kNil;
            if (key_[n] == key)
                return n;
        }
    }
    void index_insert(u64 key, u32 n) {
        size_t h = hash(key) & mask_;
        while (table_[h] != kEmpty)
            h = (h + 1) & mask_;
        table_[h] = n;
    }
    void index_erase(u64 key) { // linear probing with backward shift: no tombstones
        size_t h = hash(key) & mask_;
        while (key_[table_[h]] != key)
            h = (h + 1) & mask_;
        size_t hole = h;
        for (size_t j = (h + 1) & mask_;; j = (j + 1) & mask_) {
            const u32 n = table_[j];
            if (n == kEmpty)
                break;
            const size_t home = hash(key_[n]) & mask_;
            // n may move into the hole if its home is not in the (cyclic) interval (hole, j]
            const bool between = hole <= j ? (home > hole && home <= j) : (home > hole || home <= j);
            if (!between) {
                table_[hole] = n;
                hole = j;
            }
        }
        table_[hole] = kEmpty;
    }
    void unlink(u32 n) {
        const u32 p = prev_[n], x = next_[n];
        if (p != kNil)
            next_[p] = x;
        else
            head_ = x;
        if (x != kNil)
            prev_[x] = p;
        else
            tail_ = p;
    }
    void push_front(u32 n) {
        prev_[n] = kNil;
        next_[n] = head_;
        if (head_ != kNil)
            prev_[head_] = n;
        head_ = n;
        if (tail_ == kNil)
            tail_ = n;
    }

    size_t cap_, mask_ = 0, size_ = 0;
    std::vector<u32> table_, prev_, next_, free_;
    std::vector<u64> key_;
    std::vector<u8> valid_;
    u32 head_, tail_;
};

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
        for (size_t i = lo; i < hi; i++) {
            const u64 *e = &s->embs[pos_of(i) * T];
            u32 *sc = &s->scores[i * W];
            for (size_t z = 0; z < W; z++)
                sc[z] = 0;
            for (size_t j = 0; j < T; j++) {
                u64 bits = 0;
                for (size_t z = 0; z < W; z++)
                    if (s->snaps[z].check(e[j])) {
                        sc[z]++;
                        bits |= 1ull << (z & 63);
                    }
                if (use_bits)
                    s->holders[i * T + j] = bits;
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
    parallel_for(W, s->threads, [&](size_t w) {
        std::vector<u64> &plan = s->plans[w];
        plan.clear();
        MiniLru &snap = s->snaps[w];
        for (size_t i = 0; i < B; i++) {
            if (s->assigned[i] == w)
                continue;
            const u64 *e = &s->embs[pos_of(i) * T];
            if (use_bits) { // what scoring saw (the snapshots have not changed since)
                const u64 *h = &s->holders[i * T];
                for (size_t j = 0; j < T; j++)
                    if ((h[j] >> w) & 1ull)
                        plan.push_back(e[j]);
            } else {
                for (size_t j = 0; j < T; j++)
                    if (snap.check(e[j]))
                        plan.push_back(e[j]);
            }
        }
        std::sort(plan.begin(), plan.end());
        plan.erase(std::unique(plan.begin(), plan.end()), plan.end());
        for (u64 k : plan)
            snap.outdate(k);
        std::vector<u64> uniq;
        uniq.reserve(s->mini * T);
        for (size_t j = 0; j < s->mini; j++) {
            const u64 *e = &s->embs[s->dist[w * s->mini + j] * T];
            uniq.insert(uniq.end(), e, e + T);
        }
        std::sort(uniq.begin(), uniq.end());
        uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
        for (u64 k : uniq)
            snap.get(k);
    });
    if (timing)
        fprintf(stderr, "laia batch %zu: score %.2f ms, assign %.2f ms, plan + snapshots %.2f ms\n", s->batch_id,
                t1 - t0, t2 - t1, now_ms() - t2);
    s->batch_id++;
}

} // namespace

extern "C" {

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
