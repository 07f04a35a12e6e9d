// Host-side driver of the deterministic segment reduce: cold segments by one warp per unique
// row (128-bit lanes), hot segments column-split across CTAs (hb_rows.cuh).
#pragma once

#include <cstdlib>

#include "hb_rows.cuh"
#include "hb_sort.cuh"

namespace hb {

constexpr u32 kDefaultHotThreshold = 64;

// process-wide default (HBSetHotThreshold / HERALD_HOT_THRESHOLD); tests lower it to push
// ordinary-sized batches through the hot path
u32 default_hot_threshold();

// f1 / f4: the same functor instantiated for VEC = 1 and VEC = 4; v4 selects the 128-bit cold
// path.  `n` = number of values (upper bound of any segment length).
template <class F1, class F4>
void run_segment_reduce(KeyWorkspace &ws, const u32 *perm, const float *vals, size_t D, size_t n,
                        bool v4, u32 hot_threshold, cudaStream_t st, F1 f1, F4 f4) {
    if (n == 0)
        return;
    const bool hot = n > hot_threshold;
    HotLists hl{ws.hot_a, ws.hot_b, ws.hot_ctrl()};
    if (hot) {
        int g = (int)std::min<size_t>((n + 255) / 256, (size_t)sm_count() * 4);
        build_hot_lists_kernel<<<std::max(g, 1), 256, 0, st>>>(ws.seg_start, ws.num_unique,
                                                             hot_threshold, hl);
        HB_LAUNCHED();
    }
    const u32 thr = hot ? hot_threshold : 0xffffffffu;
    int grid = row_grid(n);
    if (v4)
        segment_rows_kernel<4, F4><<<grid, kRowBlock, 0, st>>>(ws.seg_start, perm, ws.num_unique,
                                                               vals, D, thr, f4);
    else
        segment_rows_kernel<1, F1><<<grid, kRowBlock, 0, st>>>(ws.seg_start, perm, ws.num_unique,
                                                               vals, D, thr, f1);
    HB_LAUNCHED();
    if (hot) {
        segment_hot_kernel<F1><<<sm_count() * 4, kRowBlock, 0, st>>>(ws.seg_start, perm, vals, D, hl,
                                                                     f1);
        HB_LAUNCHED();
        segment_hot_finish_kernel<F1><<<32, kRowBlock, 0, st>>>(ws.seg_start, hl, f1);
        HB_LAUNCHED();
    }
}

} // namespace hb
