// TEST INFRASTRUCTURE — CPU oracle ("port"), never linked into or called by the
// product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.
//
// A from-scratch restatement of the reference's embedding hot path:
//   worker cache (hetu_cache)  +  parameter-server handlers for cache tables.
// Parity status: PINNED — tests/test_oracle_port.py drives this port and the
// compiled reference (oracle/_ref, built from /root/reference by oracle/Makefile)
// with identical random call sequences and requires bit-identical rows, versions,
// key sets and perf counters; tests/golden/*.npz hold reference-generated vectors.
//
// Each block cites the reference file:line it restates (paths relative to the
// reference root).  The structure is deliberately different from the reference:
// replacement order is kept as a (use-count, stamp) priority in an ordered set
// instead of linked lists, and lines are plain structs in one hash map.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <unordered_map>
#include <vector>

namespace {

using u64 = uint64_t;
using i64 = int64_t;

// ---------------------------------------------------------------------------
// Owner side: authoritative rows + per-row version.
// ps-lite/include/ps/server/param.h:119-138 (CacheTable: rows + ver[], zero-init)
// ---------------------------------------------------------------------------
struct Server {
    size_t len, width;
    std::vector<float> rows;
    std::vector<i64> ver;
    Server(size_t l, size_t w) : len(l), width(w), rows(l * w, 0.f), ver(l, 0) {}

    // ps-lite/src/PSFhandle_embedding.cc:5-28 (kPushEmbedding): serial, in request order
    void push(u64 key, const float *grad, i64 updates) {
        ver[key] += updates;
        float *r = &rows[key * width];
        for (size_t j = 0; j < width; j++)
            r[j] += grad[j];
    }
    // ps-lite/src/PSFhandle_embedding.cc:30-64 (kSyncEmbedding) staleness predicate
    bool stale(u64 key, i64 client_ver, i64 bound) const {
        return client_ver == -1 || ver[key] - client_ver > bound;
    }
};

// ---------------------------------------------------------------------------
// One cache line.  src/hetu_cache/include/embedding.h:19-149
// ---------------------------------------------------------------------------
struct Line {
    u64 key;
    i64 version = -1; // embedding.h:35,42
    i64 updates = 0;
    bool has_data;
    bool has_grad = false; // grad_ lazily allocated (embedding.h:120-123)
    std::vector<float> data, grad;
    // replacement-order state
    i64 use = 0;
    u64 stamp = 0;

    Line(u64 k, size_t w, bool with_data) : key(k), has_data(with_data) {
        if (with_data)
            data.assign(w, 0.f);
        grad.assign(w, 0.f);
    }
    // embedding.h:78-91
    void accumulate(const float *g, size_t w) {
        has_grad = true;
        if (!has_data) {
            for (size_t j = 0; j < w; j++)
                grad[j] += g[j];
        } else {
            for (size_t j = 0; j < w; j++) {
                grad[j] += g[j];
                data[j] += g[j];
            }
        }
        updates++;
    }
    // embedding.h:92-96
    void addup(size_t w) {
        if (has_grad)
            for (size_t j = 0; j < w; j++)
                data[j] += grad[j];
    }
    // embedding.h:112-118
    void zero_grad(size_t w) {
        has_grad = true;
        std::fill(grad.begin(), grad.end(), 0.f);
        updates = 0;
    }
};
using LinePtr = std::shared_ptr<Line>;

enum Policy { kLRU = 0, kLFU = 1, kLFUOpt = 2 };
constexpr i64 kUseCntMax = 10; // src/hetu_cache/include/lfuopt_cache.h:25

// ---------------------------------------------------------------------------
// Sorted unique + inverse.  src/hetu_cache/include/unqiue_tools.h:9-48
// ---------------------------------------------------------------------------
void sorted_unique(const u64 *keys, size_t n, std::vector<u64> &uniq,
                   std::vector<size_t> &inverse) {
    std::vector<size_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [keys](size_t a, size_t b) { return keys[a] < keys[b]; });
    uniq.clear();
    inverse.assign(n, 0);
    for (size_t p = 0; p < n; p++) {
        if (p == 0 || keys[order[p]] != keys[order[p - 1]])
            uniq.push_back(keys[order[p]]);
        inverse[order[p]] = uniq.size() - 1;
    }
}

struct Perf {
    i64 num_all = 0, num_unique = 0, num_miss = 0, num_evict = 0,
        num_transfered = 0, is_full = 0;
};

// ---------------------------------------------------------------------------
// Worker cache: policy + batched entry points.
// ---------------------------------------------------------------------------
struct Cache {
    Server *srv;
    Policy policy;
    size_t limit, width;
    i64 pull_bound = 5, push_bound = 5; // src/hetu_cache/include/cache.h:26-27
    bool bypass = false;
    u64 clock = 0;

    std::unordered_map<u64, LinePtr> lines;            // evictable lines
    std::set<std::tuple<i64, u64, u64>> order;         // (use, stamp, key) ascending = next victim first
    std::unordered_map<u64, LinePtr> store;            // LFUOpt permanent store
    std::vector<LinePtr> evicted;                      // dirty victims awaiting the next push

    size_t size() const { return lines.size() + store.size(); }

    void touch(const LinePtr &l, i64 new_use) {
        order.erase({l->use, l->stamp, l->key});
        l->use = new_use;
        l->stamp = ++clock;
        order.insert({l->use, l->stamp, l->key});
    }
    void add(const LinePtr &l, i64 use) {
        l->use = use;
        l->stamp = ++clock;
        lines[l->key] = l;
        order.insert({l->use, l->stamp, l->key});
    }
    void evict_one() {
        auto it = order.begin();
        u64 k = std::get<2>(*it);
        LinePtr victim = lines[k];
        order.erase(it);
        lines.erase(k);
        if (victim->updates != 0) // lru_cache.cc:19-23, lfu_cache.cc:31-41, lfuopt_cache.cc:52-63
            evicted.push_back(victim);
    }

    // policy lookup: lru_cache.cc:27-39, lfu_cache.cc:22-29 (+52-69), lfuopt_cache.cc:28-44
    LinePtr lookup(u64 k) {
        if (policy == kLFUOpt) {
            auto s = store.find(k);
            if (s != store.end())
                return s->second;
        }
        auto it = lines.find(k);
        if (it == lines.end())
            return nullptr;
        LinePtr l = it->second;
        switch (policy) {
        case kLRU:
            touch(l, 0);
            break;
        case kLFU:
            touch(l, l->use + 1);
            break;
        case kLFUOpt:
            if (l->use + 1 < kUseCntMax) {
                touch(l, l->use + 1);
            } else { // promoted to the never-evicted store
                order.erase({l->use, l->stamp, l->key});
                lines.erase(k);
                store[k] = l;
            }
            break;
        }
        return l;
    }

    // policy insert: lru_cache.cc:9-25, lfu_cache.cc:9-20, lfuopt_cache.cc:9-26
    void insert(const LinePtr &l) {
        switch (policy) {
        case kLRU: {
            auto it = lines.find(l->key);
            if (it != lines.end()) {
                order.erase({it->second->use, it->second->stamp, it->second->key});
                lines.erase(it);
            }
            add(l, 0);
            if (lines.size() > limit)
                evict_one();
            break;
        }
        case kLFU: {
            auto it = lines.find(l->key);
            if (it == lines.end()) {
                if (lines.size() == limit)
                    evict_one();
                add(l, 1);
            } else { // re-insert of a present key counts as a use
                i64 use = it->second->use;
                order.erase({use, it->second->stamp, l->key});
                lines.erase(it);
                add(l, use + 1);
            }
            break;
        }
        case kLFUOpt: {
            if (store.count(l->key)) {
                store[l->key] = l;
                return;
            }
            auto it = lines.find(l->key);
            if (it != lines.end()) { // pointer replaced, position kept
                l->use = it->second->use;
                l->stamp = it->second->stamp;
                it->second = l;
            } else {
                if (size() == limit) {
                    if (!lines.empty())
                        evict_one();
                    else
                        return; // only the permanent store is populated: do not cache
                }
                add(l, 0);
            }
            break;
        }
        }
    }

    // cache.cc:15-26
    std::vector<LinePtr> batched_lookup(const std::vector<u64> &uniq) {
        std::vector<LinePtr> out(uniq.size());
        if (bypass)
            return out;
        for (size_t i = 0; i < uniq.size(); i++)
            out[i] = lookup(uniq[i]);
        return out;
    }
    // cache.cc:28-35
    void batched_insert(std::vector<LinePtr> &ls) {
        if (bypass)
            return;
        for (auto &l : ls)
            insert(l);
    }

    // hetu_client.cc:6-37 + server kSyncEmbedding; returns #rows transferred
    size_t sync(std::vector<LinePtr> &embeds) {
        size_t pulled = 0;
        for (auto &l : embeds) {
            if (srv->stale(l->key, l->version, pull_bound)) {
                l->version = srv->ver[l->key];
                std::copy(&srv->rows[l->key * width], &srv->rows[(l->key + 1) * width],
                          l->data.begin());
                l->addup(width);
                pulled++;
            }
        }
        return pulled;
    }
    // hetu_client.cc:39-55 + server kPushEmbedding
    void push(std::vector<LinePtr> &ls) {
        for (auto &l : ls) {
            l->has_grad = true; // grad() allocates (embedding.h:58-61)
            srv->push(l->key, l->grad.data(), l->updates);
        }
    }

    // cache.cc:60-107
    void embedding_lookup(const u64 *keys, size_t n, float *dest, Perf *perf) {
        std::vector<u64> uniq;
        std::vector<size_t> inv;
        sorted_unique(keys, n, uniq, inv);
        auto embeds = batched_lookup(uniq);
        std::vector<LinePtr> fresh;
        for (size_t i = 0; i < uniq.size(); i++)
            if (!embeds[i]) {
                embeds[i] = std::make_shared<Line>(uniq[i], width, true);
                fresh.push_back(embeds[i]);
            }
        size_t pulled = sync(embeds);
        for (size_t r = 0; r < n; r++)
            std::copy(embeds[inv[r]]->data.begin(), embeds[inv[r]]->data.end(),
                      dest + r * width);
        batched_insert(fresh);
        if (perf) {
            perf->num_all = n;
            perf->num_unique = uniq.size();
            perf->num_miss = fresh.size();
            perf->num_transfered = pulled;
            perf->is_full = size() == limit;
        }
    }

    // shared front half of cache.cc:132-154 / 248-270 / 372-391
    void accumulate_all(const u64 *keys, size_t n, const float *grads,
                        std::vector<u64> &uniq, std::vector<LinePtr> &embeds,
                        size_t &miss_cnt) {
        std::vector<size_t> inv;
        sorted_unique(keys, n, uniq, inv);
        embeds = batched_lookup(uniq);
        miss_cnt = 0;
        for (size_t r = 0; r < n; r++) {
            size_t i = inv[r];
            if (!embeds[i]) {
                embeds[i] = std::make_shared<Line>(uniq[i], width, false); // dataless
                miss_cnt++;
            }
            embeds[i]->accumulate(grads + r * width, width);
        }
    }

    // cache.cc:132-196
    void embedding_update(const u64 *keys, size_t n, const float *grads, Perf *perf) {
        std::vector<u64> uniq;
        std::vector<LinePtr> embeds;
        size_t miss_cnt;
        std::vector<LinePtr> ev = std::move(evicted);
        evicted.clear();
        // note: reference moves evict_ after batchedLookup; lookups never append to it
        accumulate_all(keys, n, grads, uniq, embeds, miss_cnt);
        std::vector<LinePtr> should_push;
        size_t e = 0;
        for (size_t i = 0; i < uniq.size(); i++) {
            if (embeds[i]->updates > push_bound || !embeds[i]->has_data) {
                while (e < ev.size() && ev[e]->key < embeds[i]->key)
                    should_push.push_back(ev[e++]);
                should_push.push_back(embeds[i]);
            }
        }
        while (e < ev.size())
            should_push.push_back(ev[e++]);
        push(should_push);
        for (size_t i = 0; i < uniq.size(); i++)
            if (embeds[i]->updates > push_bound && embeds[i]->has_data) {
                embeds[i]->version += embeds[i]->updates;
                embeds[i]->zero_grad(width);
            }
        if (perf) {
            perf->num_all = n;
            perf->num_unique = uniq.size();
            perf->num_evict = ev.size();
            perf->num_miss = miss_cnt;
            perf->num_transfered = should_push.size();
            perf->is_full = size() == limit;
        }
    }

    // cache.cc:248-334 (Laia/Herald plan: push only keys named by the plan)
    void embedding_update_push_keys(const u64 *keys, size_t n, const u64 *push_keys,
                                    size_t n_push, const float *grads, Perf *perf) {
        std::vector<u64> uniq;
        std::vector<LinePtr> embeds;
        size_t miss_cnt;
        std::vector<LinePtr> ev = std::move(evicted);
        evicted.clear();
        accumulate_all(keys, n, grads, uniq, embeds, miss_cnt);
        std::vector<LinePtr> should_push;
        std::vector<size_t> pushed_idx;
        size_t e = 0, p = 0;
        for (size_t i = 0; i < uniq.size(); i++) {
            while (e < ev.size() && ev[e]->key < embeds[i]->key)
                should_push.push_back(ev[e++]);
            while (p < n_push && push_keys[p] < embeds[i]->key)
                p++;
            if (p < n_push && push_keys[p] == embeds[i]->key && embeds[i]->has_data) {
                should_push.push_back(embeds[i]);
                pushed_idx.push_back(i);
            }
        }
        while (e < ev.size())
            should_push.push_back(ev[e++]);
        push(should_push);
        for (size_t i = 0, j = 0; i < uniq.size(); i++) {
            embeds[i]->version += embeds[i]->updates; // every touched line, every call
            if (j < pushed_idx.size() && pushed_idx[j] == i) {
                embeds[i]->zero_grad(width);
                j++;
            }
        }
        if (perf) {
            perf->num_all = n;
            perf->num_unique = uniq.size();
            perf->num_evict = ev.size();
            perf->num_miss = miss_cnt;
            perf->num_transfered = should_push.size();
            perf->is_full = size() == limit;
        }
    }

    // cache.cc:356-422 (ASP prefetch: push current batch, pull next batch, one round trip)
    void embedding_push_pull(const u64 *pull_keys, size_t n_pull, float *dest,
                             const u64 *push_keys, size_t n_push, const float *grads) {
        std::vector<u64> uniq;
        std::vector<size_t> inv;
        sorted_unique(pull_keys, n_pull, uniq, inv);
        auto embeds = batched_lookup(uniq);
        std::vector<LinePtr> fresh;
        for (size_t i = 0; i < uniq.size(); i++)
            if (!embeds[i]) {
                embeds[i] = std::make_shared<Line>(uniq[i], width, true);
                fresh.push_back(embeds[i]);
            }
        std::vector<u64> puniq;
        std::vector<LinePtr> pembeds;
        size_t miss_cnt;
        // reference order: push-side batchedLookup, then evict_ is taken
        accumulate_all_deferred_evict(push_keys, n_push, grads, puniq, pembeds, miss_cnt);
        std::vector<LinePtr> ev = std::move(evicted);
        evicted.clear();
        std::vector<LinePtr> should_push;
        size_t e = 0;
        for (size_t i = 0; i < puniq.size(); i++) {
            if (pembeds[i]->updates > push_bound || !pembeds[i]->has_data) {
                while (e < ev.size() && ev[e]->key < pembeds[i]->key)
                    should_push.push_back(ev[e++]);
                should_push.push_back(pembeds[i]);
            }
        }
        while (e < ev.size())
            should_push.push_back(ev[e++]);
        // server: push first, then sync (PSFhandle_embedding.cc:66-79)
        push(should_push);
        sync(embeds);
        for (size_t r = 0; r < n_pull; r++)
            std::copy(embeds[inv[r]]->data.begin(), embeds[inv[r]]->data.end(),
                      dest + r * width);
        batched_insert(fresh);
        for (size_t i = 0; i < puniq.size(); i++)
            if (pembeds[i]->updates > push_bound && pembeds[i]->has_data) {
                pembeds[i]->version += pembeds[i]->updates;
                pembeds[i]->zero_grad(width);
            }
    }
    void accumulate_all_deferred_evict(const u64 *keys, size_t n, const float *grads,
                                       std::vector<u64> &uniq,
                                       std::vector<LinePtr> &embeds, size_t &miss_cnt) {
        accumulate_all(keys, n, grads, uniq, embeds, miss_cnt);
    }

    LinePtr find(u64 k) const {
        auto s = store.find(k);
        if (s != store.end())
            return s->second;
        auto it = lines.find(k);
        return it == lines.end() ? nullptr : it->second;
    }
};

void fill_perf(const Perf &p, i64 *out) {
    if (!out)
        return;
    out[0] = p.num_all;
    out[1] = p.num_unique;
    out[2] = p.num_miss;
    out[3] = p.num_evict;
    out[4] = p.num_transfered;
    out[5] = p.is_full;
}

} // namespace

extern "C" {

void *hp_server_create(size_t len, size_t width) {
    return new Server(len, width);
}
void hp_server_destroy(void *s) {
    delete static_cast<Server *>(s);
}
void hp_server_load(void *s, const float *rows) {
    auto *srv = static_cast<Server *>(s);
    std::copy(rows, rows + srv->len * srv->width, srv->rows.begin());
}
void hp_server_read(void *s, float *rows, i64 *ver) {
    auto *srv = static_cast<Server *>(s);
    if (rows)
        std::copy(srv->rows.begin(), srv->rows.end(), rows);
    if (ver)
        std::copy(srv->ver.begin(), srv->ver.end(), ver);
}
float *hp_server_rows_ptr(void *s) {
    return static_cast<Server *>(s)->rows.data();
}

void *hp_cache_create(void *server, int policy, size_t limit, size_t width) {
    auto *c = new Cache();
    c->srv = static_cast<Server *>(server);
    c->policy = static_cast<Policy>(policy);
    c->limit = limit;
    c->width = width;
    return c;
}
void hp_cache_destroy(void *c) {
    delete static_cast<Cache *>(c);
}
void hp_cache_set_bounds(void *c, i64 pull_bound, i64 push_bound) {
    static_cast<Cache *>(c)->pull_bound = pull_bound;
    static_cast<Cache *>(c)->push_bound = push_bound;
}
void hp_cache_set_bypass(void *c, int on) {
    static_cast<Cache *>(c)->bypass = on != 0;
}
void hp_cache_lookup(void *c, const u64 *keys, size_t n, float *dest, i64 *perf) {
    Perf p;
    static_cast<Cache *>(c)->embedding_lookup(keys, n, dest, &p);
    fill_perf(p, perf);
}
void hp_cache_update(void *c, const u64 *keys, size_t n, const float *grads, i64 *perf) {
    Perf p;
    static_cast<Cache *>(c)->embedding_update(keys, n, grads, &p);
    fill_perf(p, perf);
}
void hp_cache_update_push_keys(void *c, const u64 *keys, size_t n, const u64 *push_keys,
                               size_t n_push, const float *grads, i64 *perf) {
    Perf p;
    static_cast<Cache *>(c)->embedding_update_push_keys(keys, n, push_keys, n_push, grads, &p);
    fill_perf(p, perf);
}
void hp_cache_push_pull(void *c, const u64 *pull_keys, size_t n_pull, float *dest,
                        const u64 *push_keys, size_t n_push, const float *grads) {
    static_cast<Cache *>(c)->embedding_push_pull(pull_keys, n_pull, dest, push_keys, n_push,
                                                 grads);
}
size_t hp_cache_size(void *c) {
    return static_cast<Cache *>(c)->size();
}
size_t hp_cache_num_evicted_pending(void *c) {
    return static_cast<Cache *>(c)->evicted.size();
}
// sorted keys, as LRUCache::PyAPI_keys (lru_cache.cc:41-48)
void hp_cache_keys(void *c, u64 *out) {
    auto *cache = static_cast<Cache *>(c);
    size_t n = 0;
    for (auto &kv : cache->store)
        out[n++] = kv.first;
    for (auto &kv : cache->lines)
        out[n++] = kv.first;
    std::sort(out, out + n);
}
// debug single-line read; returns 1 if present
int hp_cache_line(void *c, u64 key, i64 *version, i64 *updates, float *data, float *grad) {
    auto *cache = static_cast<Cache *>(c);
    LinePtr l = cache->find(key);
    if (!l)
        return 0;
    if (version)
        *version = l->version;
    if (updates)
        *updates = l->updates;
    if (data && l->has_data)
        std::copy(l->data.begin(), l->data.end(), data);
    if (grad)
        std::copy(l->grad.begin(), l->grad.end(), grad);
    return 1;
}
// debug single-key policy calls (python_api.cc:56-60 `lookup` / `insert`)
int hp_cache_touch(void *c, u64 key) {
    return static_cast<Cache *>(c)->lookup(key) != nullptr;
}
void hp_cache_insert_line(void *c, u64 key, i64 version, const float *data) {
    auto *cache = static_cast<Cache *>(c);
    auto l = std::make_shared<Line>(key, cache->width, true);
    std::copy(data, data + cache->width, l->data.begin());
    l->version = version;
    cache->insert(l);
}

void hp_unique(const u64 *keys, size_t n, u64 *uniq, u64 *inverse, size_t *num_unique) {
    std::vector<u64> u;
    std::vector<size_t> inv;
    sorted_unique(keys, n, u, inv);
    std::copy(u.begin(), u.end(), uniq);
    for (size_t i = 0; i < n; i++)
        inverse[i] = inv[i];
    *num_unique = u.size();
}

} // extern "C"
