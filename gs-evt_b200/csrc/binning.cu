// Binning of the OPERATOR path (the drop-in diff_gaussian_rasterization package): prefix sum of tile counts, (tile, depth)
// key emission, key sort, tile ranges — the reference's own steps on the reference's own buffer layout, because the
// parity tests compare these work buffers with the reference's byte for byte.  The tracking engine does not come through
// here: its binning is bucketbin.cu (no library kernels).
// Restates rasterizer_impl.cu:70-138,280-321.  Integer work, bit-exact by construction:
//   key   = (tile_id << 32) | float_bits(depth)          (duplicateWithKeys, :70-111)
//   order = stable ascending sort on bits [0, 32+bit)     (cub::DeviceRadixSort, :306-311)
//   range = [first, last+1) of each tile's run, (0,0) if untouched (identifyTileRanges, :116-138)
// The scan and the radix sort are CUB device primitives from the CUDA toolkit (the same third-party
// library the reference calls); everything else is ours.  HBM-bound.
#include "internal.h"
#include <cub/cub.cuh>

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

}  // namespace gsevt
