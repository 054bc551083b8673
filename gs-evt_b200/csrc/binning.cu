// Binning: prefix sum of tile counts, (tile, depth) key emission, key sort, tile ranges.
// Restates rasterizer_impl.cu:70-138,280-321.  Integer work, bit-exact by construction:
//   key   = (tile_id << 32) | float_bits(depth)          (duplicateWithKeys, :70-111)
//   order = stable ascending sort on bits [0, 32+bit)     (cub::DeviceRadixSort, :306-311)
//   range = [first, last+1) of each tile's run, (0,0) if untouched (identifyTileRanges, :116-138)
// The scan and the radix sort are CUB device primitives from the CUDA toolkit (the same third-party
// library the reference calls); everything else is ours.  HBM-bound.
#include "internal.h"
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

namespace gsevt {

uint32_t higher_msb(uint32_t n) {  // getHigherMsb, rasterizer_impl.cu:35-50
    uint32_t msb = sizeof(n) * 4;
    uint32_t step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

size_t scan_temp_bytes(int n) {
    size_t bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, n > 0 ? n : 1);
    return bytes;
}
size_t sort_temp_bytes(int n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, n > 0 ? n : 1);
    return bytes;
}
void launch_scan(void* temp, size_t temp_bytes, const uint32_t* in, uint32_t* out, int n, cudaStream_t s) {
    if (n <= 0) return;
    cub::DeviceScan::InclusiveSum(temp, temp_bytes, in, out, n, s);
}
void launch_sort_pairs(void* temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
                       const uint32_t* vals_in, uint32_t* vals_out, int n, int end_bit, cudaStream_t s) {
    if (n <= 0) return;
    cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, n, 0, end_bit, s);
}

// One thread per (view, Gaussian); a visible Gaussian writes its rect's keys in row-major tile order
// starting at offsets[i-1] (rasterizer_impl.cu:85-109).  With nviews == 2 both views share one
// instance list: view v uses tile ids v*tiles + t, so one sort orders both.
// Threads beyond nviews*P fill unused capacity with sentinel keys (engine only).
__global__ void __launch_bounds__(256) emit_keys_kernel(int P, int nviews, const ViewParams* __restrict__ views,
                                                        const float4* __restrict__ rec, const int* __restrict__ radii,
                                                        const uint32_t* __restrict__ offsets,
                                                        uint64_t* __restrict__ keys, uint32_t* __restrict__ values,
                                                        int cap, int* __restrict__ overflow,
                                                        const EngineCtl* __restrict__ ctl) {
    if (ctl && ctl->level_done) return;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = nviews * P;
    const uint32_t total = offsets[n - 1];
    if (cap > 0) {
        if (gid == 0 && total > (uint32_t)cap) *overflow = 1;
        // sentinel fill: grid also covers [0, cap)
        if (gid < cap && (uint32_t)gid >= total) {
            keys[gid] = ~0ull;
            values[gid] = 0u;
        }
    }
    if (gid >= n) return;
    const int r = radii[gid];
    if (r <= 0) return;
    const int v = gid / P;
    const int gx = views[v].grid_x, gy = views[v].grid_y;
    const float4 r0 = __ldg(rec + 2 * (size_t)gid);
    const float4 r1 = __ldg(rec + 2 * (size_t)gid + 1);
    int x0, y0, x1, y1;
    tile_rect(r0.x, r0.y, r, gx, gy, x0, y0, x1, y1);
    uint32_t off = gid == 0 ? 0u : offsets[gid - 1];
    const uint32_t depth_bits = __float_as_uint(r1.w);
    const uint32_t tile_base = (uint32_t)v * (uint32_t)(gx * gy);
    const uint32_t idx = (uint32_t)(gid - v * P);
    for (int y = y0; y < y1; y++) {
        for (int x = x0; x < x1; x++) {
            if (cap > 0 && off >= (uint32_t)cap) return;
            const uint64_t key = ((uint64_t)(tile_base + (uint32_t)(y * gx + x)) << 32) | depth_bits;
            keys[off] = key;
            values[off] = idx;
            off++;
        }
    }
}

void launch_emit_keys(int P, int nviews, const ViewParams* views, const float4* rec, const int* radii,
                      const uint32_t* offsets, uint64_t* keys, uint32_t* values, int cap, int* overflow,
                      const EngineCtl* ctl, cudaStream_t s) {
    const int n = nviews * P;
    const int threads = n > cap ? n : cap;
    if (threads <= 0) return;
    emit_keys_kernel<<<(threads + 255) / 256, 256, 0, s>>>(P, nviews, views, rec, radii, offsets, keys, values, cap,
                                                           overflow, ctl);
}

__global__ void __launch_bounds__(256) identify_ranges_kernel(const uint64_t* __restrict__ keys,
                                                              uint2* __restrict__ ranges, int n_host,
                                                              const uint32_t* __restrict__ n_dev, int cap) {
    const uint32_t L = n_host >= 0 ? (uint32_t)n_host : min(*n_dev, (uint32_t)cap);
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L) return;
    const uint32_t cur = (uint32_t)(keys[idx] >> 32);
    if (idx == 0)
        ranges[cur].x = 0;
    else {
        const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
        if (cur != prev) {
            ranges[prev].y = idx;
            ranges[cur].x = idx;
        }
    }
    if (idx == L - 1) ranges[cur].y = L;
}

void launch_identify_ranges(const uint64_t* keys, uint2* ranges, int ntiles_total, int n_host, const uint32_t* n_dev,
                            int cap, cudaStream_t s) {
    cudaMemsetAsync(ranges, 0, (size_t)ntiles_total * sizeof(uint2), s);
    const int threads = n_host >= 0 ? n_host : cap;
    if (threads <= 0) return;
    identify_ranges_kernel<<<(threads + 255) / 256, 256, 0, s>>>(keys, ranges, n_host, n_dev, cap);
}

// Operator path: gather the caller's device-side matrices into one ViewParams block.
__global__ void build_view_params_kernel(ViewParams* out, const float* view, const float* proj, const float* proj_raw,
                                         const float* campos, const float* vel, const float* vel_inv, const float* bg,
                                         float tanfovx, float tanfovy, int W, int H, float delta_time) {
    const int t = threadIdx.x;
    if (t < 16) {
        out->view[t] = view[t];
        out->proj[t] = proj[t];
        out->vel[t] = vel ? vel[t] : (t % 5 == 0 ? 1.0f : 0.0f);
        out->vel_inv[t] = vel_inv ? vel_inv[t] : (t % 5 == 0 ? 1.0f : 0.0f);
    }
    if (t < 3) {
        out->campos[t] = campos[t];
        out->bg[t] = bg ? bg[t] : 0.0f;
    }
    if (t == 0) {
        out->tanfovx = tanfovx;
        out->tanfovy = tanfovy;
        // rasterizer_impl.cu:225-226: focal = size / (2 * tan)
        out->focal_x = W / (2.0f * tanfovx);
        out->focal_y = H / (2.0f * tanfovy);
        out->W = W;
        out->H = H;
        out->grid_x = (W + GSEVT_TILE - 1) / GSEVT_TILE;
        out->grid_y = (H + GSEVT_TILE - 1) / GSEVT_TILE;
        out->proj_a = proj_raw ? proj_raw[0] : 0.0f;
        out->proj_b = proj_raw ? proj_raw[5] : 0.0f;
        out->proj_e = proj_raw ? proj_raw[11] : 0.0f;
        out->delta_time = delta_time;
        out->pad_ = 0.0f;
    }
}

void launch_build_view_params(ViewParams* out, const float* view, const float* proj, const float* proj_raw,
                              const float* campos, const float* vel, const float* vel_inv, const float* bg,
                              float tanfovx, float tanfovy, int W, int H, float delta_time, cudaStream_t s) {
    build_view_params_kernel<<<1, 32, 0, s>>>(out, view, proj, proj_raw, campos, vel, vel_inv, bg, tanfovx, tanfovy, W,
                                              H, delta_time);
}

// ------------------------------------------------------------------------------------------------
// Engine path: depth sort of the visible pairs, then the per-tile lists.
//
// The reference's order inside a tile is (depth bits, Gaussian index) — the second by stability of the
// radix sort over the emission order (rasterizer_impl.cu:85-109,306-311).  The same total order is
// produced with far less traffic by
//   1. sorting the VISIBLE (view, Gaussian) pairs — compacted in index order by preprocess.cu — ONCE by
//      view << 31 | depth bits (stable: ties keep index order; sentinel keys 0xFFFFFFFF fill the slack),
//   2. a stable partition of the instance sequence by tile id: tilebin.cu (counting, the default), or — for
//      grids with more than 2048 tiles per view and strip — the kernels of this file: tile instances emitted in
//      that order with a 16-bit key = tile id (+ tiles per view for view 1), a stable CUB sort on the tile id alone
//      (<= 13 bits: two 8-bit passes instead of six over 64-bit keys), and a range scan.
// Either way point lists and tile ranges are identical to the reference's; gsevt_engine_binning() rebuilds the
// 64-bit keys for the tests.
// ------------------------------------------------------------------------------------------------
size_t sort32_temp_bytes(int n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (const unsigned long long*)nullptr, (unsigned long long*)nullptr, n > 0 ? n : 1);
    return bytes;
}
size_t sort16_temp_bytes(int n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint16_t*)nullptr, (uint16_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, n > 0 ? n : 1);
    return bytes;
}
// depth sort: key = depth bits, value = packed {tile rect (high 32) | pair id (low 32)} so that everything the
// emission needs travels with the sort and is read back coalesced
void launch_sort_pairs32(void* temp, size_t temp_bytes, const uint32_t* keys_in, uint32_t* keys_out,
                         const uint64_t* vals_in, uint64_t* vals_out, int n, cudaStream_t s) {
    if (n <= 0) return;
    cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, (const unsigned long long*)vals_in,
                                    (unsigned long long*)vals_out, n, 0, 32, s);
}
void launch_sort_pairs16(void* temp, size_t temp_bytes, const uint16_t* keys_in, uint16_t* keys_out,
                         const uint32_t* vals_in, uint32_t* vals_out, int n, int end_bit, cudaStream_t s) {
    if (n <= 0) return;
    cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, n, 0, end_bit, s);
}

namespace {
__host__ __device__ __forceinline__ uint32_t rect_area(uint32_t r) {
    return ((r >> 16 & 255u) - (r & 255u)) * ((r >> 24) - (r >> 8 & 255u));
}
struct PairArea {
    const uint64_t* pairs;
    const uint32_t* n_live;   // entries at and past *n_live are padding (whatever they hold): area 0
    __host__ __device__ __forceinline__ uint32_t operator()(int i) const {
        return (uint32_t)i < *n_live ? rect_area((uint32_t)(pairs[i] >> 32)) : 0u;
    }
};
}  // namespace
size_t scan_gather_temp_bytes(int n) {
    size_t bytes = 0;
    auto it = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), PairArea{nullptr, nullptr});
    cub::DeviceScan::InclusiveSum(nullptr, bytes, it, (uint32_t*)nullptr, n > 0 ? n : 1);
    return bytes;
}
// offsets[i] = inclusive sum of the tile-rect areas of pairs[0..i] (pairs in emission order)
void launch_scan_gather(void* temp, size_t temp_bytes, const uint64_t* pairs, const uint32_t* n_live, uint32_t* offsets, int n,
                        cudaStream_t s) {
    if (n <= 0) return;
    auto it = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), PairArea{pairs, n_live});
    cub::DeviceScan::InclusiveSum(temp, temp_bytes, it, offsets, n, s);
}

// Tile-instance emission, output-centric.  A CTA takes 256 consecutive pairs of the depth-sorted list (coalesced
// 8-byte loads, nothing gathered), whose instances occupy ONE contiguous output range; its threads then walk that
// range with unit stride — each output slot finds its pair by binary search over the 256 offsets in shared memory
// — so the (tile, id) stores are fully coalesced and the work is balanced whatever the rect sizes are.
// (The reference emits one serial, divergent loop per Gaussian: rasterizer_impl.cu:85-109.)
// Threads past the pair list fill the unused capacity [total, cap) with sentinel keys.
__global__ void __launch_bounds__(256) emit_tiles_kernel(int P, int n, int gx, int tiles_per_view, const uint64_t* __restrict__ pairs,
                                                         const uint32_t* __restrict__ offsets, uint16_t* __restrict__ keys,
                                                         uint32_t* __restrict__ values, int cap, int* __restrict__ overflow,
                                                         const EngineCtl* __restrict__ ctl) {
    if (ctl && ctl->level_done) return;
    __shared__ uint32_t s_off[257];     // exclusive offsets of the CTA's pairs, [256] = end
    __shared__ uint32_t s_rect[256];
    __shared__ uint32_t s_gid[256];
    __shared__ float s_rcpw[256];
    const int i = blockIdx.x * 256 + threadIdx.x;   // n = pairs in the (depth-sorted, compacted) list incl. sentinel slack
    const uint32_t total = offsets[n - 1];
    if (i == 0 && total > (uint32_t)cap) *overflow = 1;
    if (i < cap && (uint32_t)i >= total) {
        keys[i] = 0xFFFFu;
        values[i] = 0u;
    }
    if (blockIdx.x * 256 >= n) return;
    uint32_t rect = 0, gid = 0, incl = total;
    if (i < n) {
        const uint64_t pr = pairs[i];
        rect = (uint32_t)(pr >> 32);
        gid = (uint32_t)pr;
        incl = offsets[i];
    }
    const uint32_t area = rect_area(rect);
    s_off[threadIdx.x] = incl - area;
    s_rect[threadIdx.x] = rect;
    s_gid[threadIdx.x] = gid;
    const uint32_t w = (rect >> 16 & 255u) - (rect & 255u);
    s_rcpw[threadIdx.x] = w ? 1.0f / (float)w : 0.0f;
    if (threadIdx.x == 255) s_off[256] = incl;
    __syncthreads();
    const uint32_t begin = s_off[0];
    const uint32_t end = min(s_off[256], (uint32_t)cap);
    for (uint32_t o = begin + threadIdx.x; o < end; o += 256) {
        // last pair g with s_off[g] <= o (zero-area pairs share their successor's offset and are skipped by this)
        int lo = 0, hi = 255;
#pragma unroll
        for (int step = 0; step < 8; step++) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_off[mid] <= o) lo = mid; else hi = mid - 1;
        }
        const uint32_t r = s_rect[lo];
        const uint32_t g = s_gid[lo];
        const uint32_t t = o - s_off[lo];
        const uint32_t x0 = r & 255u, y0 = r >> 8 & 255u, wd = (r >> 16 & 255u) - x0;
        const uint32_t ty = (uint32_t)(((float)t + 0.5f) * s_rcpw[lo]);   // exact: t < 65536, wd <= 255
        const uint32_t tx = t - ty * wd;
        const uint32_t v = g >= (uint32_t)P ? 1u : 0u;
        keys[o] = (uint16_t)(v * (uint32_t)tiles_per_view + (y0 + ty) * (uint32_t)gx + x0 + tx);
        values[o] = g - v * (uint32_t)P;
    }
}
void launch_emit_tiles(int P, int n_pairs, int grid_x, int tiles_per_view, const uint64_t* pairs, const uint32_t* offsets,
                       uint16_t* keys, uint32_t* values, int cap, int* overflow, const EngineCtl* ctl, cudaStream_t s) {
    const int threads = n_pairs > cap ? n_pairs : cap;
    if (threads <= 0 || n_pairs <= 0) return;
    emit_tiles_kernel<<<(threads + 255) / 256, 256, 0, s>>>(P, n_pairs, grid_x, tiles_per_view, pairs, offsets, keys, values, cap,
                                                           overflow, ctl);
}

// Eight keys per thread (one 16-byte load + the key before them).
__global__ void __launch_bounds__(256) identify_ranges16_kernel(const uint16_t* __restrict__ keys, uint2* __restrict__ ranges,
                                                                const uint32_t* __restrict__ n_dev, int cap) {
    const uint32_t L = min(*n_dev, (uint32_t)cap);
    const uint32_t first = (blockIdx.x * blockDim.x + threadIdx.x) * 8u;
    if (first >= L) return;
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(keys + first));   // buffers are allocated in multiples of 256 B
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t prev = first == 0 ? 0xFFFFFFFFu : (uint32_t)__ldg(keys + first - 1);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t idx = first + k;
        if (idx >= L) break;
        const uint32_t cur = (w[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
        if (idx == 0) {
            ranges[cur].x = 0;
        } else if (cur != prev) {
            ranges[prev].y = idx;
            ranges[cur].x = idx;
        }
        if (idx == L - 1) ranges[cur].y = L;
        prev = cur;
    }
}
void launch_identify_ranges16(const uint16_t* keys, uint2* ranges, int ntiles_total, const uint32_t* n_dev, int cap,
                              cudaStream_t s) {
    cudaMemsetAsync(ranges, 0, (size_t)ntiles_total * sizeof(uint2), s);
    if (cap <= 0) return;
    const int threads = (cap + 7) / 8;
    identify_ranges16_kernel<<<(threads + 255) / 256, 256, 0, s>>>(keys, ranges, n_dev, cap);
}

// Screen-tile split: tile instances per tile row (both views), the cost model the strips are balanced on.
__global__ void __launch_bounds__(256) row_histogram_kernel(int n, const uint64_t* __restrict__ pairs, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const uint32_t r = (uint32_t)(pairs[i] >> 32);
        const uint32_t w = (r >> 16 & 255u) - (r & 255u);
        if (w == 0) continue;
        for (uint32_t y = r >> 8 & 255u; y < r >> 24; y++) atomicAdd(&s_h[y], w);
    }
    __syncthreads();
    if (s_h[threadIdx.x]) atomicAdd(hist + threadIdx.x, s_h[threadIdx.x]);
}
void launch_row_histogram(int n_pairs, const uint64_t* pairs, uint32_t* hist256, cudaStream_t s) {
    cudaMemsetAsync(hist256, 0, 256 * sizeof(uint32_t), s);
    if (n_pairs <= 0) return;
    int blocks = (n_pairs + 255) / 256;
    if (blocks > 1184) blocks = 1184;   // 8 CTAs per SM
    row_histogram_kernel<<<blocks, 256, 0, s>>>(n_pairs, pairs, hist256);
}

// Parity-test helper: rebuild the reference's 64-bit keys of one view from the engine's sorted lists.
__global__ void rebuild_keys_kernel(const uint16_t* __restrict__ tile_keys, const uint32_t* __restrict__ vals,
                                    const float4* __restrict__ rec_view, uint32_t tile_base, uint32_t first, uint32_t count,
                                    uint64_t* __restrict__ keys_out, uint32_t* __restrict__ list_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t id = vals[first + i];
    const uint32_t tile = (uint32_t)tile_keys[first + i] - tile_base;
    keys_out[i] = ((uint64_t)tile << 32) | __float_as_uint(rec_view[2 * (size_t)id + 1].w);
    list_out[i] = id;
}
void launch_rebuild_keys(const uint16_t* tile_keys, const uint32_t* vals, const float4* rec_view, uint32_t tile_base,
                         uint32_t first, uint32_t count, uint64_t* keys_out, uint32_t* list_out, cudaStream_t s) {
    if (count) rebuild_keys_kernel<<<(count + 255) / 256, 256, 0, s>>>(tile_keys, vals, rec_view, tile_base, first, count, keys_out, list_out);
}

}  // namespace gsevt
