// Host-side driver of the deterministic segment reduce: cold segments by one warp per unique
// row (128-bit lanes), hot segments column-split across CTAs (hb_rows.cuh).
#pragma once

#include <cstdlib>
#include <functional>

#include "hb_rows.cuh"
#include "hb_sort.cuh"

namespace hb {

constexpr u32 kDefaultHotThreshold = 64;

// process-wide default (HBSetHotThreshold / HERALD_HOT_THRESHOLD); tests lower it to push
// ordinary-sized batches through the hot path
u32 default_hot_threshold();

template <class K>
int persistent_grid(K kernel, size_t smem) {
    int per_sm = 0;
    HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kRowBlock, smem));
    return sm_count() * std::max(per_sm, 1);
}

// segments a warp keeps in flight in the cold phase (2 or 4; $HERALD_SEG_ROWS for tuning)
inline int seg_rows() {
    static const int rows = [] {
        const char *e = getenv("HERALD_SEG_ROWS");
        int r = e ? atoi(e) : 4;
        return r == 2 ? 2 : 4;
    }();
    return rows;
}

inline u32 seg_mode() {
    static const u32 mode = [] {
        const char *e = getenv("HERALD_SEG_MODE");
        return e ? (u32)atoi(e) : 0u;
    }();
    return mode;
}

template <int VEC, int ROWS, class FV, class F1>
void launch_segment_reduce(KeyWorkspace &ws, const u32 *perm, const float *vals, size_t D, size_t n,
                           bool hot, u32 thr, const HotLists &hl, cudaStream_t st, FV fv, F1 f1) {
    auto k = segment_reduce_kernel<VEC, ROWS, FV, F1>;
    const size_t smem = hot_smem_bytes(hot_stages());
    static const int full = persistent_grid(k, smem);
    // one warp per cold ticket (upper bound) or medium row
    const size_t units = (n + ticket_rows() - 1) / ticket_rows() + n / (medium_threshold() + 1);
    int grid = (int)std::max<size_t>(1, std::min<size_t>(full, (units + kRowWarps - 1) / kRowWarps));
    HB_LAUNCH(k, hot ? full : grid, kRowBlock, smem, st, perm, ws.num_unique, vals, D, thr, hl, fv, f1);
    HB_LAUNCHED();
}

// Everything the plan and the data kernel of one segment reduce share.
struct SegSetup {
    HotLists hl;
    u32 thr;  // hot threshold as the kernels see it (0xffffffff: no hot path)
    bool hot;
};

// `n` = number of values (upper bound of any segment length); split: two-level reduction wanted and
// the functor supports it.
inline SegSetup seg_setup(KeyWorkspace &ws, size_t D, size_t n, u32 hot_threshold, bool split, cudaStream_t st) {
    SegSetup s;
    s.hot = n > hot_threshold;
    s.thr = s.hot ? hot_threshold : 0xffffffffu;
    split = split && s.hot;
    if (split) { // run sums of the very hot rows: at most n / kVeryHot rows qualify
        const size_t need = std::min(ws.split_rows_cap(), n / kVeryHot + 1) * kSplitTiles * D;
        if (need > ws.split_partials_cap) {
            HB_CUDA(cudaStreamSynchronize(st));
            if (ws.split_partials)
                cudaFree(ws.split_partials);
            ws.split_partials = nullptr;
            HB_CUDA(cudaMalloc((void **)&ws.split_partials, need * sizeof(float)));
            ws.split_partials_cap = need;
        }
    }
    s.hl = HotLists{ws.hot_a, ws.hot_b, ws.medium, ws.hot_ctrl(), ws.seg_items, seg_trace_buffer(),
                    hot_stages(), ticket_rows(), medium_threshold(), seg_mode(),
                    split ? 1u : 0u, ws.split_partials, ws.split_done()};
    return s;
}

// f1 / f4: the same functor instantiated for VEC = 1 and VEC = 4; v4 selects the 128-bit cold
// path.  `planned`: the work items and lists have already been written (by a caller that plans
// inside another kernel: seg_plan_item) with the SegSetup passed in `pre`.
template <class F1, class F4>
void run_segment_reduce(KeyWorkspace &ws, const u32 *perm, const float *vals, size_t D, size_t n,
                        bool v4, u32 hot_threshold, cudaStream_t st, F1 f1, F4 f4,
                        const std::function<void()> &after_plan = nullptr, bool split = false,
                        const SegSetup *pre = nullptr) {
    if (n == 0)
        return;
    const SegSetup su = pre ? *pre : seg_setup(ws, D, n, hot_threshold, split && can_split<F1>::value, st);
    if (!pre) {   // open every unique row and lay the work items out in ticket order
        const size_t total = n + ticket_rows();
        int g = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 8);
        HB_LAUNCH(seg_plan_kernel<F4>, std::max(g, 1), 256, 0, st, ws.seg_start, perm, ws.num_unique, su.thr,
                                                        su.hl, f4);
        HB_LAUNCHED();
    }
    if (after_plan)
        after_plan(); // (timing mark between the plan kernel and the data kernel)
    // (four rows in flight per warp in the cold phase; the two-row variants were tuning aids)
    if (v4)
        launch_segment_reduce<4, 4>(ws, perm, vals, D, n, su.hot, su.thr, su.hl, st, f4, f1);
    else
        launch_segment_reduce<1, 4>(ws, perm, vals, D, n, su.hot, su.thr, su.hl, st, f1, f1);
}

} // namespace hb
