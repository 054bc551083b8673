// Bucket scatter of ONE (view, Gaussian) pair, split into phases so that it can run either as its own kernel
// (bucket_scatter_kernel, bucketbin.cu: count-only probe and the unfused path) or inside the projection kernel
// (preprocess_map_kernel<.., true>, preprocess.cu), where the latency of the cursor atomics hides behind the SH
// evaluation.  See bucketbin.cu for the scheme.
#pragma once
#include "internal.h"

namespace gsevt {
namespace bscatter {

constexpr uint32_t SMALL = 4;   // rects meeting up to 4 buckets take the straight-line path
constexpr uint32_t CSTEP = GSEVT_BK_SUB * GSEVT_BK_CURSOR_STRIDE;

// Which of a bucket's 2 x 2 tiles a rect covers, for the bucket at offset (dx, dy) inside the rect's bucket range:
// ax = 2 * bx0 - x0 (0 or -1), wx = x1 - x0 (tiles), same in y.  bit ky * 2 + kx (== sortcore::cover_mask4).
__device__ __forceinline__ uint32_t cover_bits(int a0, uint32_t w, uint32_t d) {
    const uint32_t t = (uint32_t)(a0 + 2 * (int)d);
    return (t < w ? 1u : 0u) | (t + 1u < w ? 2u : 0u);
}
template <int S>
__device__ __forceinline__ uint32_t key_low(uint32_t id, int ax, uint32_t wx, int ay, uint32_t wy, uint32_t dx, uint32_t dy) {
    if constexpr (S == 0) return id << 4;
    const uint32_t cx = cover_bits(ax, wx, dx), cy = cover_bits(ay, wy, dy);
    return (id << 4) | ((cy & 1u) ? cx : 0u) | ((cy & 2u) ? cx << 2 : 0u);
}

struct Pair {
    uint32_t cnt, w, b00, wx, wy;   // buckets met (0: not visible), bucket-range width, first bucket, rect size in tiles
    int ax, ay;
    uint32_t slot[SMALL];           // cursor values of the straight-line path
};

// rect = x0 | y0 << 8 | x1 << 16 | y1 << 24 in tiles, 0 = not visible
template <int S>
__device__ __forceinline__ void prepare(const BucketArgs& a, uint32_t view, uint32_t rect, Pair& p) {
    const int x0 = (int)(rect & 255u), y0 = (int)(rect >> 8 & 255u), x1 = (int)(rect >> 16 & 255u), y1 = (int)(rect >> 24);
    const int bx0 = x0 >> S, by0 = y0 >> S;                         // rect == 0: w = h = 0 below
    p.w = rect ? (uint32_t)(((x1 - 1) >> S) + 1 - bx0) : 0u;
    const uint32_t h = rect ? (uint32_t)(((y1 - 1) >> S) + 1 - by0) : 0u;
    p.cnt = p.w * h;
    p.ax = 2 * bx0 - x0; p.ay = 2 * by0 - y0;
    p.wx = (uint32_t)(x1 - x0); p.wy = (uint32_t)(y1 - y0);
    p.b00 = view * (uint32_t)a.nb + (uint32_t)(by0 - a.by_origin) * (uint32_t)a.nbx + (uint32_t)bx0;
}

__device__ __forceinline__ void bucket_offset(const Pair& p, uint32_t t, uint32_t& dx, uint32_t& dy) {
    dy = t == 0 ? 0u : (t >= p.w ? 1u : 0u) + (t >= 2u * p.w ? 1u : 0u) + (t >= 3u * p.w ? 1u : 0u);
    dx = t - dy * p.w;
}

// straight-line path, phase 1: all cursor atomics in flight
__device__ __forceinline__ void issue_small(const BucketArgs& a, uint32_t* cur0, Pair& p) {
    if (p.cnt == 0 || p.cnt > SMALL) return;
#pragma unroll
    for (uint32_t t = 0; t < SMALL; t++) {
        if (t < p.cnt) {
            uint32_t dx, dy;
            bucket_offset(p, t, dx, dy);
            p.slot[t] = atomicAdd(cur0 + (size_t)(p.b00 + dy * (uint32_t)a.nbx + dx) * CSTEP, 1u);
        }
    }
}

// straight-line path, phase 2: the keys
template <int S>
__device__ __forceinline__ void finish_small(const BucketArgs& a, uint32_t sub, const Pair& p, uint32_t depth, uint32_t id) {
    if (p.cnt == 0 || p.cnt > SMALL) return;
#pragma unroll
    for (uint32_t t = 0; t < SMALL; t++) {
        if (t < p.cnt) {
            uint32_t dx, dy;
            bucket_offset(p, t, dx, dy);
            const uint32_t bb = p.b00 + dy * (uint32_t)a.nbx + dx;
            const uint32_t subcap = __ldg(a.bk_cap + bb) / GSEVT_BK_SUB;
            if (p.slot[t] < subcap)
                a.keys[__ldg(a.bk_start + bb) + sub * subcap + p.slot[t]] = ((uint64_t)depth << 32) | key_low<S>(id, p.ax, p.wx, p.ay, p.wy, dx, dy);
            else *a.overflow = 1;
        }
    }
}

// large rects: the whole warp walks one pair's buckets, 32 per step, so that no lane loops over a screen-filling
// Gaussian alone.  Every lane of the warp must call this (lanes without a large rect pass cnt <= SMALL).  `sub` is the
// OWNER's sub-segment (it may differ from lane to lane when the pairs come from a list).
template <bool COUNT_ONLY, int S>
__device__ __forceinline__ void big_rects(const BucketArgs& a, uint32_t sub_mine, const Pair& p, uint32_t depth, uint32_t id) {
    unsigned bigs = __ballot_sync(0xffffffffu, p.cnt > SMALL);
    const uint32_t lane = threadIdx.x & 31u;
    while (bigs) {
        const int src = __ffs(bigs) - 1;
        bigs &= bigs - 1;
        const uint32_t sub = __shfl_sync(0xffffffffu, sub_mine, src);
        uint32_t* const cur0 = a.cursor + (size_t)sub * GSEVT_BK_CURSOR_STRIDE;
        const uint32_t b_cnt = __shfl_sync(0xffffffffu, p.cnt, src), b_w = __shfl_sync(0xffffffffu, p.w, src);
        const uint32_t b_b00 = __shfl_sync(0xffffffffu, p.b00, src), b_wx = __shfl_sync(0xffffffffu, p.wx, src), b_wy = __shfl_sync(0xffffffffu, p.wy, src);
        const int b_ax = __shfl_sync(0xffffffffu, p.ax, src), b_ay = __shfl_sync(0xffffffffu, p.ay, src);
        const uint32_t b_depth = __shfl_sync(0xffffffffu, depth, src), b_id = __shfl_sync(0xffffffffu, id, src);
        for (uint32_t t = lane; t < b_cnt; t += 32u) {
            const uint32_t dy = t / b_w, dx = t - dy * b_w;
            const uint32_t bb = b_b00 + dy * (uint32_t)a.nbx + dx;
            const uint32_t slot = atomicAdd(cur0 + (size_t)bb * CSTEP, 1u);
            if constexpr (!COUNT_ONLY) {
                const uint32_t subcap = __ldg(a.bk_cap + bb) / GSEVT_BK_SUB;
                if (slot < subcap) a.keys[__ldg(a.bk_start + bb) + sub * subcap + slot] = ((uint64_t)b_depth << 32) | key_low<S>(b_id, b_ax, b_wx, b_ay, b_wy, dx, dy);
                else *a.overflow = 1;
            }
        }
    }
}

}  // namespace bscatter
}  // namespace gsevt
