// TEST INFRASTRUCTURE — not product code.
//
// In-process replacement for the three ps-lite transport functions that the
// reference worker cache calls (declared in the reference at
// ps-lite/include/ps/worker/hetu_binding.h:14-28).  The reference sources
// (src/hetu_cache/src/*.cc, ps-lite/src/PSFhandle_embedding.cc,
// ps-lite/src/thread_pool.cc) are compiled UNMODIFIED from /root/reference by
// oracle/Makefile; this file only supplies the "network": a request is served
// by calling the reference server handler directly.
//
// Row-range partitioning over S servers follows PSAgent::syncEmbedding /
// pushEmbedding / pushSyncEmbedding (ps-lite/include/ps/worker/PSAgent.h:537-627)
// with the AveragePartitioner split (ps-lite/include/ps/partitioner.h:46-57).
#include "ps/server/PSFHandle.h"
#include "ps/worker/hetu_binding.h"

#include <algorithm>
#include <map>
#include <memory>
#include <vector>

namespace ps {

namespace {
using Handler = PSHandler<PsfGroup::kParameterServer>;

struct TableMeta {
    size_t len = 0, width = 0;
    std::vector<size_t> part; // rows owned by each server
};

int g_nserver = 1;
std::vector<std::unique_ptr<Handler>> g_servers;
std::map<int, TableMeta> g_meta;

Handler &server(int s) {
    while ((int)g_servers.size() <= s)
        g_servers.emplace_back(new Handler());
    return *g_servers[s];
}

template <typename Fn>
void for_each_range(const TableMeta &meta, const SArray<uint64_t> &rows, Fn fn) {
    size_t start = 0, end = 0, cur_len = 0;
    for (size_t s = 0; s < meta.part.size(); s++) {
        start = end;
        end = std::lower_bound(rows.begin() + start, rows.end(),
                               cur_len + meta.part[s])
              - rows.begin();
        fn(s, start, end, cur_len);
        cur_len += meta.part[s];
    }
}

SArray<size_t> rebased(const SArray<uint64_t> &rows, size_t start, size_t end,
                       size_t base) {
    SArray<size_t> out(end - start);
    for (size_t i = start; i < end; i++)
        out[i - start] = rows[i] - base;
    return out;
}
} // namespace

const char *getPSFunctionName(const PsfType &) {
    return "oracle-shim";
}

void debug() {
}

void syncEmbedding(int node_id, const SArray<uint64_t> &keys,
                   const SArray<version_t> &ver, version_t bound,
                   PSFData<kSyncEmbedding>::Closure closure) {
    const TableMeta &meta = g_meta.at(node_id);
    for_each_range(meta, keys, [&](size_t s, size_t start, size_t end, size_t base) {
        if (start == end)
            return;
        PSFData<kSyncEmbedding>::Request req((Key)node_id,
                                             rebased(keys, start, end, base),
                                             ver.segment(start, end), bound);
        PSFData<kSyncEmbedding>::Response resp;
        server(s).serve(req, resp);
        closure(resp, start);
    });
}

void PushEmbedding(int node_id, const SArray<uint64_t> &keys,
                   const SArray<float> &data,
                   const SArray<version_t> &updates) {
    const TableMeta &meta = g_meta.at(node_id);
    for_each_range(meta, keys, [&](size_t s, size_t start, size_t end, size_t base) {
        if (start == end)
            return;
        PSFData<kPushEmbedding>::Request req(
            (Key)node_id, rebased(keys, start, end, base),
            data.segment(start * meta.width, end * meta.width),
            updates.segment(start, end));
        PSFData<kPushEmbedding>::Response resp;
        server(s).serve(req, resp);
    });
}

void PushSyncEmbedding(int node_id, const SArray<uint64_t> &keys,
                       const SArray<version_t> &ver, version_t bound,
                       PSFData<kSyncEmbedding>::Closure closure,
                       const SArray<uint64_t> &push_keys,
                       const SArray<float> &data,
                       const SArray<version_t> &updates) {
    const TableMeta &meta = g_meta.at(node_id);
    size_t start = 0, end = 0, pstart = 0, pend = 0, cur_len = 0;
    for (size_t s = 0; s < meta.part.size(); s++) {
        start = end;
        pstart = pend;
        end = std::lower_bound(keys.begin() + start, keys.end(),
                               cur_len + meta.part[s])
              - keys.begin();
        pend = std::lower_bound(push_keys.begin() + pstart, push_keys.end(),
                                cur_len + meta.part[s])
               - push_keys.begin();
        if (!(start == end && pstart == pend)) {
            PSFData<kPushSyncEmbedding>::Request req(
                (Key)node_id, rebased(keys, start, end, cur_len),
                ver.segment(start, end), bound,
                rebased(push_keys, pstart, pend, cur_len),
                data.segment(pstart * meta.width, pend * meta.width),
                updates.segment(pstart, pend));
            PSFData<kPushSyncEmbedding>::Response resp;
            server(s).serve(req, resp);
            closure(resp, start);
        }
        cur_len += meta.part[s];
    }
}

} // namespace ps

// ---- helpers reached from Python via ctypes.CDLL(hetu_cache.__file__) ----
extern "C" {

void oracle_set_servers(int nserver) {
    ps::g_nserver = nserver < 1 ? 1 : nserver;
}

// ParamInit of a kCacheTable on every server partition.
void oracle_init_table(int id, size_t len, size_t width, int init_type,
                       double a, double b, unsigned long long seed) {
    ps::TableMeta meta;
    meta.len = len;
    meta.width = width;
    size_t S = ps::g_nserver;
    for (size_t s = 0; s < S; s++)
        meta.part.push_back(len / S + (s < len % S));
    ps::g_meta[id] = meta;
    for (size_t s = 0; s < S; s++) {
        ::SArray<float> lrs(1);
        lrs[0] = 0.1f;
        ps::PSFData<ps::ParamInit>::Request req(
            (ps::Key)id, (int)ps::kCacheTable, meta.part[s], width, init_type,
            a, b, seed + s, (int)ps::SGD, lrs);
        ps::PSFData<ps::ParamInit>::Response resp;
        ps::server(s).serve(req, resp);
    }
}

// Zero-initialised table + additive DensePush == load exact rows.
void oracle_load_rows(int id, size_t len, size_t width, const float *rows) {
    oracle_init_table(id, len, width, (int)ps::Constant, 0.0, 0.0, 0);
    const ps::TableMeta &meta = ps::g_meta.at(id);
    size_t base = 0;
    for (size_t s = 0; s < meta.part.size(); s++) {
        size_t n = meta.part[s] * width;
        ::SArray<float> vals(n);
        std::copy(rows + base * width, rows + base * width + n, vals.begin());
        ps::PSFData<ps::DensePush>::Request req((ps::Key)id, n, vals);
        ps::PSFData<ps::DensePush>::Response resp;
        ps::server(s).serve(req, resp);
        base += meta.part[s];
    }
}

void oracle_read_rows(int id, float *out) {
    const ps::TableMeta &meta = ps::g_meta.at(id);
    size_t base = 0;
    for (size_t s = 0; s < meta.part.size(); s++) {
        size_t n = meta.part[s] * meta.width;
        ps::PSFData<ps::DensePull>::Request req((ps::Key)id, n);
        ps::PSFData<ps::DensePull>::Response resp;
        ps::server(s).serve(req, resp);
        auto &v = std::get<0>(resp);
        std::copy(v.begin(), v.end(), out + base * meta.width);
        base += meta.part[s];
    }
}

// Server versions, read through the sync protocol with "never synced" clients.
void oracle_read_versions(int id, int64_t *out) {
    const ps::TableMeta &meta = ps::g_meta.at(id);
    size_t base = 0;
    for (size_t s = 0; s < meta.part.size(); s++) {
        size_t n = meta.part[s];
        ::SArray<size_t> rows(n);
        ::SArray<ps::version_t> ver(n);
        for (size_t i = 0; i < n; i++) {
            rows[i] = i;
            ver[i] = -1;
        }
        ps::PSFData<ps::kSyncEmbedding>::Request req((ps::Key)id, rows, ver, 0);
        ps::PSFData<ps::kSyncEmbedding>::Response resp;
        ps::server(s).serve(req, resp);
        auto &rv = std::get<1>(resp);
        std::copy(rv.begin(), rv.end(), out + base);
        base += n;
    }
}

// Rows / versions of SELECTED keys (ascending, global row ids) through the sync protocol with
// "never synced" clients: what a fresh worker would be sent.  Nothing on the server changes.
void oracle_read_rows_at(int id, const uint64_t *keys, size_t n, float *out_rows,
                         int64_t *out_ver) {
    const ps::TableMeta &meta = ps::g_meta.at(id);
    ::SArray<uint64_t> k(n);
    for (size_t i = 0; i < n; i++)
        k[i] = keys[i];
    ps::for_each_range(meta, k, [&](size_t s, size_t start, size_t end, size_t base) {
        if (start == end)
            return;
        ::SArray<ps::version_t> ver(end - start);
        for (size_t i = 0; i < end - start; i++)
            ver[i] = -1;
        ps::PSFData<ps::kSyncEmbedding>::Request req((ps::Key)id, ps::rebased(k, start, end, base),
                                                     ver, 0);
        ps::PSFData<ps::kSyncEmbedding>::Response resp;
        ps::server(s).serve(req, resp);
        auto &idx = std::get<0>(resp);
        auto &rv = std::get<1>(resp);
        auto &rows = std::get<2>(resp);
        for (size_t j = 0; j < idx.size(); j++) {
            const size_t pos = start + idx[j];
            if (out_ver)
                out_ver[pos] = rv[j];
            if (out_rows)
                std::copy(rows.begin() + j * meta.width, rows.begin() + (j + 1) * meta.width,
                          out_rows + pos * meta.width);
        }
    });
}

// rows[keys[i], :] += vals[i, :] (SparsePush).  On a zero-initialised table this LOADS exact rows
// at the given keys without materialising the whole table on the Python side.
void oracle_add_rows_at(int id, const uint64_t *keys, size_t n, const float *vals) {
    const ps::TableMeta &meta = ps::g_meta.at(id);
    ::SArray<uint64_t> k(n);
    for (size_t i = 0; i < n; i++)
        k[i] = keys[i];
    ps::for_each_range(meta, k, [&](size_t s, size_t start, size_t end, size_t base) {
        if (start == end)
            return;
        ::SArray<float> v((end - start) * meta.width);
        std::copy(vals + start * meta.width, vals + end * meta.width, v.begin());
        ps::PSFData<ps::SparsePush>::Request req((ps::Key)id, ps::rebased(k, start, end, base), v);
        ps::PSFData<ps::SparsePush>::Response resp;
        ps::server(s).serve(req, resp);
    });
}

void oracle_clear_table(int id) {
    auto it = ps::g_meta.find(id);
    if (it == ps::g_meta.end())
        return;
    for (size_t s = 0; s < it->second.part.size(); s++) {
        ps::PSFData<ps::ParamClear>::Request req((ps::Key)id);
        ps::PSFData<ps::ParamClear>::Response resp;
        ps::server(s).serve(req, resp);
    }
    ps::g_meta.erase(it);
}

} // extern "C"
