// Tile blending, forward and backward.
//
// Forward restates renderCUDA (dgr/cuda_rasterizer/forward.cu:263-392); backward restates renderCUDA
// (dgr/cuda_rasterizer/backward.cu:679-903).  One CTA = one 16x16 tile, one thread = one pixel, the
// tile's depth-sorted instance list is staged through shared memory in batches.
//
// What is different from the reference (results identical):
//  * each warp owns a compact 8x4 pixel block and skips an instance with one warp-uniform bounding-box
//    test against the ellipse {alpha >= 1/255} (computed once per instance by the thread that stages it);
//  * one 32-byte record per Gaussian {x, y, A, B | C, opacity, gray, depth} = one DRAM/L2 sector per
//    gather instead of three separate arrays; colours ride along (RGB in a second 16-byte record);
//  * backward starts at the last instance any pixel of the tile actually blended (max n_contrib) instead
//    of the end of the list, reduces each instance's gradient with warp shuffles, parks per-warp partial
//    sums in shared memory and issues one vector atomic per instance per batch — no block barrier inside
//    the per-instance loop (the reference has ~12);
//  * the engine variant renders one grayscale channel for BOTH views in one launch (blockIdx.z = view)
//    and its backward derives dL/dpixel on the fly from the normalised event loss (frame.py:86-92,
//    tracker.py:93-103) instead of reading upstream gradient images.
// The per-pixel arithmetic (power, expf, alpha, transmittance test, accumulation) follows the
// reference's rounding order so n_contrib / final_T are bit-identical given identical inputs.
#include "internal.h"

namespace gsevt {

namespace {

constexpr float kAlphaMin = 1.0f / 255.0f;

// Half extents (plus the warp block's own half size) of the bounding box of {Q(d) <= tau}; +inf when the
// conic is degenerate, -1 when the instance can never reach alpha >= 1/255.
__device__ __forceinline__ float2 cull_extent(float A, float B, float C, float o) {
    const float tau = 2.0f * __logf(255.0f * o) * 1.01f + 0.05f;  // conservative
    const float det = A * C - B * B;
    if (!(tau > 0.0f)) return make_float2(-1.0f, -1.0f);
    if (!(det > 0.0f) || !(A > 0.0f) || !(C > 0.0f)) return make_float2(3.0e38f, 3.0e38f);
    const float inv = tau / det;
    return make_float2(sqrtf(inv * C) * 1.001f + 3.5f + 0.05f, sqrtf(inv * A) * 1.001f + 1.5f + 0.05f);
}

// alpha and its ingredients in the reference's rounding order (forward.cu:342-353).
__device__ __forceinline__ float eval_power(float dx, float dy, float A, float B, float C) {
    const float q = __fmaf_rn(dx, __fmul_rn(dx, A), __fmul_rn(dy, __fmul_rn(dy, C)));
    return __fmaf_rn(q, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, B)));
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Forward
// ------------------------------------------------------------------------------------------------
template <int C, bool OPERATOR>
__global__ void __launch_bounds__(256) blend_fwd_kernel(BlendFwdArgs a) {
    if (a.ctl && a.ctl->level_done) return;
    __shared__ float4 s_r0[2][256];
    __shared__ float4 s_r1[2][256];
    __shared__ float4 s_cull[2][256];   // {x, y, half extent x, half extent y} of the alpha >= 1/255 ellipse's box
    __shared__ float4 s_rgb[OPERATOR ? 2 : 1][OPERATOR ? 256 : 1];
    __shared__ int s_id[OPERATOR ? 2 : 1][OPERATOR ? 256 : 1];

    const int view = blockIdx.z;
    const int tiles = a.grid_x * a.grid_y;
    const int HW = a.W * a.H;
    const uint2 range = a.ranges[view * tiles + blockIdx.y * a.grid_x + blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bx = blockIdx.x * GSEVT_TILE + (warp & 1) * 8, by = blockIdx.y * GSEVT_TILE + (warp >> 1) * 4;
    const int pixx = bx + (lane & 7), pixy = by + (lane >> 3);
    const float cxw = (float)bx + 3.5f, cyw = (float)by + 1.5f;
    const float pxf = (float)pixx, pyf = (float)pixy;
    const bool inside = pixx < a.W && pixy < a.H;
    bool done = !inside;

    const float4* __restrict__ rec = a.rec + 2 * (size_t)view * a.view_stride_gauss;
    const int todo = (int)(range.y - range.x);
    const int rounds = (todo + 255) / 256;

    float T = 1.0f;
    uint32_t last_contributor = 0;
    float acc[C];
#pragma unroll
    for (int ch = 0; ch < C; ch++) acc[ch] = 0.0f;
    float D = 0.0f;

    // double-buffered staging: ONE block barrier per round (it also counts the finished pixels)
    auto stage = [&](int i, int buf) {
        const int prog = i * 256 + threadIdx.x;
        if (prog < todo) {
            const uint32_t id = __ldg(a.point_list + range.x + prog);
            const float4 r0 = __ldg(rec + 2 * (size_t)id);
            const float4 r1 = __ldg(rec + 2 * (size_t)id + 1);
            s_r0[buf][threadIdx.x] = r0;
            s_r1[buf][threadIdx.x] = r1;
            const float2 ext = cull_extent(r0.z, r0.w, r1.x, r1.y);
            s_cull[buf][threadIdx.x] = make_float4(r0.x, r0.y, ext.x, ext.y);
            if constexpr (OPERATOR) {
                s_rgb[buf][threadIdx.x] = __ldg(a.rgb4 + id);
                s_id[buf][threadIdx.x] = (int)id;
            }
        }
    };
    if (rounds > 0) stage(0, 0);
    for (int i = 0; i < rounds; i++) {
        const int buf = i & 1;
        if (__syncthreads_count(done) == 256) break;
        if (i + 1 < rounds) stage(i + 1, buf ^ 1);
        const int base = i * 256;
        const int cnt = min(256, todo - base);
        // 32 staged instances at a time: lane l tests instance j0+l against this warp's 8x4 pixel block, the
        // ballot is the ordered hit list, and only hits are evaluated per pixel (list order is preserved).
        for (int j0 = 0; j0 < cnt; j0 += 32) {
            if (__all_sync(0xffffffffu, done)) break;
            bool hit = false;
            if (j0 + lane < cnt) {
                const float4 c = s_cull[buf][j0 + lane];
                hit = fabsf(c.x - cxw) <= c.z && fabsf(c.y - cyw) <= c.w;
            }
            unsigned hits = __ballot_sync(0xffffffffu, hit);
            while (hits) {
                const int j = j0 + __ffs(hits) - 1;
                hits &= hits - 1;
                if (done) continue;
                const float4 r0 = s_r0[buf][j];
                const float dx = __fadd_rn(r0.x, -pxf), dy = __fadd_rn(r0.y, -pyf);
                const float4 r1 = s_r1[buf][j];
                const float power = eval_power(dx, dy, r0.z, r0.w, r1.x);
                if (power > 0.0f) continue;
                const float alpha = fminf(0.99f, __fmul_rn(r1.y, expf(power)));
                if (alpha < kAlphaMin) continue;
                const float test_T = __fmul_rn(T, __fadd_rn(1.0f, -alpha));
                if (test_T < 0.0001f) {
                    done = true;
                    continue;
                }
                if constexpr (OPERATOR) {
                    const float4 col = s_rgb[buf][j];
                    acc[0] = __fmaf_rn(T, __fmul_rn(alpha, col.x), acc[0]);
                    acc[1] = __fmaf_rn(T, __fmul_rn(alpha, col.y), acc[1]);
                    acc[2] = __fmaf_rn(T, __fmul_rn(alpha, col.z), acc[2]);
                    D = __fmaf_rn(T, __fmul_rn(alpha, r1.w), D);
                    if (a.n_touched && test_T > 0.5f) atomicAdd(a.n_touched + s_id[buf][j], 1);
                } else {
                    acc[0] = __fmaf_rn(T, __fmul_rn(alpha, r1.z), acc[0]);
                }
                T = test_T;
                last_contributor = (uint32_t)(base + j + 1);
            }
        }
    }

    if (inside) {
        const size_t pix = (size_t)pixy * a.W + pixx;
        a.final_T[(size_t)view * HW + pix] = T;
        a.n_contrib[(size_t)view * HW + pix] = last_contributor;
        if constexpr (OPERATOR) {
#pragma unroll
            for (int ch = 0; ch < C; ch++) a.out_color[(size_t)ch * HW + pix] = __fmaf_rn(__ldg(a.bg + ch), T, acc[ch]);
            a.out_depth[pix] = D;
            a.out_opacity[pix] = __fadd_rn(1.0f, -T);
        } else {
            const ViewParams& vp = a.views[view];
            const float bgg = GSEVT_GRAY_R * vp.bg[0] + GSEVT_GRAY_G * vp.bg[1] + GSEVT_GRAY_B * vp.bg[2];
            a.out_color[(size_t)view * HW + pix] = __fmaf_rn(bgg, T, acc[0]);
        }
    }
}

void launch_blend_fwd_rgb(const BlendFwdArgs& a, cudaStream_t s) {
    dim3 grid(a.grid_x, a.grid_y, a.nviews);
    blend_fwd_kernel<3, true><<<grid, 256, 0, s>>>(a);
}
void launch_blend_fwd_gray(const BlendFwdArgs& a, cudaStream_t s) {
    dim3 grid(a.grid_x, a.grid_y, a.nviews);
    blend_fwd_kernel<1, false><<<grid, 256, 0, s>>>(a);
}

// ------------------------------------------------------------------------------------------------
// Backward
// ------------------------------------------------------------------------------------------------
// Values reduced per instance.  Engine: dmx dmy dA dB dC dgray (6, padded to 8 for the reduction).
// Operator: dmx dmy dA dB dC dop dc0 ddepth | dc1 dc2 (10, padded to 16).
template <bool OPERATOR> struct BwdCfg;
template <> struct BwdCfg<false> { static constexpr int K = 6, KR = 8, BATCH = 256; };
template <> struct BwdCfg<true> { static constexpr int K = 10, KR = 16, BATCH = 256; };

// Sum of each of N (power of two <= 32) per-lane values over the warp.  Returns, on every lane, the total of
// component lane / (32 / N).  Halving exchange: at each step a lane keeps one half of its values and trades
// the other half with its partner, so N values cost N - 1 + log2(32 / N) shuffles instead of 5 N.
template <int N>
__device__ __forceinline__ float warp_reduce_scatter(float (&v)[N], int lane) {
    int off = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1) {
        const int half = n >> 1;
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; i++) {
            const float send = hi ? v[i] : v[half + i];
            const float keep = hi ? v[half + i] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
        off >>= 1;
    }
    float r = v[0];
#pragma unroll
    for (; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
    return r;
}

// Back-to-front traversal.  Per staged batch (double-buffered: one block barrier per batch):
//   * lane l tests instance j0+l against the warp's 8x4 pixel block and against the deepest position any of the
//     warp's pixels blended; the ballot is the ordered hit list;
//   * per hit: recompute alpha, rebuild T and the colour behind, per-lane gradient terms;
//   * reduce-scatter over the warp and ONE red.add instruction per warp and hit: the lanes that own a component
//     add it straight into the Gaussian's 32-byte accumulator record (same sector, one L2 request).
// No shared-memory parking, no per-instance barrier; the reference has ~12 barriers and a 256-thread tree
// reduction per Gaussian-tile instance (backward.cu:783-900).
template <int C, bool OPERATOR>
__global__ void __launch_bounds__(256) blend_bwd_kernel(BlendBwdArgs a) {
    if (a.ctl && a.ctl->level_done) return;
    constexpr int K = BwdCfg<OPERATOR>::K, KR = BwdCfg<OPERATOR>::KR, BATCH = BwdCfg<OPERATOR>::BATCH;
    static_assert(BATCH == 256, "one staged instance per thread");
    __shared__ float4 s_r0[2][BATCH];
    __shared__ float4 s_r1[2][BATCH];
    __shared__ float4 s_cull[2][BATCH];
    __shared__ float4 s_rgb[OPERATOR ? 2 : 1][OPERATOR ? BATCH : 1];
    __shared__ uint32_t s_id[2][BATCH];
    __shared__ uint32_t s_max[8];

    const int view = blockIdx.z;
    const int tiles = a.grid_x * a.grid_y;
    const int HW = a.W * a.H;
    const uint2 range = a.ranges[view * tiles + blockIdx.y * a.grid_x + blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bx = blockIdx.x * GSEVT_TILE + (warp & 1) * 8, by = blockIdx.y * GSEVT_TILE + (warp >> 1) * 4;
    const int pixx = bx + (lane & 7), pixy = by + (lane >> 3);
    const float cxw = (float)bx + 3.5f, cyw = (float)by + 1.5f;
    const float pxf = (float)pixx, pyf = (float)pixy;
    const bool inside = pixx < a.W && pixy < a.H;
    const size_t pix = (size_t)pixy * a.W + pixx;

    const float T_final = inside ? a.final_T[(size_t)view * HW + pix] : 0.0f;
    const uint32_t last_contributor = inside ? a.n_contrib[(size_t)view * HW + pix] : 0u;

    float dpix[C];
    float dpix_depth = 0.0f;
    float bg_dot = 0.0f;
#pragma unroll
    for (int ch = 0; ch < C; ch++) dpix[ch] = 0.0f;
    if (inside) {
        if constexpr (OPERATOR) {
#pragma unroll
            for (int ch = 0; ch < C; ch++) {
                dpix[ch] = a.dL_dpix[(size_t)ch * HW + pix];
                bg_dot += __ldg(a.bg + ch) * dpix[ch];
            }
            if (a.dL_dpix_depth) dpix_depth = a.dL_dpix_depth[pix];
        } else {
            // d = gray_next - gray_last; dL/dd = alpha*d - beta*E_eff (see DESIGN.md "Loss")
            const float d = a.gray[(size_t)HW + pix] - a.gray[pix];
            const float E = a.event_frame[pix];
            const float Eeff = a.ctl->loss_signed ? E : (d > 0.0f ? fabsf(E) : (d < 0.0f ? -fabsf(E) : 0.0f));
            const float g = a.ctl->loss_alpha * d - a.ctl->loss_beta * Eeff;
            dpix[0] = view == 1 ? g : -g;
            const ViewParams& vp = a.views[view];
            bg_dot = (GSEVT_GRAY_R * vp.bg[0] + GSEVT_GRAY_G * vp.bg[1] + GSEVT_GRAY_B * vp.bg[2]) * dpix[0];
        }
    }

    // Deepest list position any pixel of this warp / this tile blended.
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, last_contributor);
    if (lane == 0) s_max[warp] = wmax;
    __syncthreads();
    uint32_t max_lc = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) max_lc = max(max_lc, s_max[w]);
    if (max_lc == 0) return;

    const float4* __restrict__ rec = a.rec + 2 * (size_t)view * a.view_stride_gauss;
    float* __restrict__ grad_f = reinterpret_cast<float*>(a.grad8 + 2 * (size_t)view * a.view_stride_gauss);
    const float ddelx_dx = 0.5f * a.W, ddely_dy = 0.5f * a.H;

    float T = T_final;
    float accum_rec[C], last_color[C];
#pragma unroll
    for (int ch = 0; ch < C; ch++) accum_rec[ch] = last_color[ch] = 0.0f;
    float accum_rec_depth = 0.0f, last_depth = 0.0f, last_alpha = 0.0f;

    // which component this lane owns after the reduce-scatter, and whether it adds it to memory
    constexpr int LPC = 32 / KR;                       // lanes per component
    const int comp = lane / LPC;
    const bool owner = (lane % LPC) == 0 && comp < K;

    auto stage = [&](int b, int buf) {
        const int hi = (int)max_lc - 1 - b * BATCH;    // list position of batch entry 0
        if ((int)threadIdx.x <= hi) {
            const uint32_t id = __ldg(a.point_list + range.x + (uint32_t)(hi - (int)threadIdx.x));
            const float4 r0 = __ldg(rec + 2 * (size_t)id);
            const float4 r1 = __ldg(rec + 2 * (size_t)id + 1);
            s_r0[buf][threadIdx.x] = r0;
            s_r1[buf][threadIdx.x] = r1;
            const float2 ext = cull_extent(r0.z, r0.w, r1.x, r1.y);
            s_cull[buf][threadIdx.x] = make_float4(r0.x, r0.y, ext.x, ext.y);
            s_id[buf][threadIdx.x] = id;
            if constexpr (OPERATOR) s_rgb[buf][threadIdx.x] = __ldg(a.rgb4 + id);
        }
    };

    const int nbatches = ((int)max_lc + BATCH - 1) / BATCH;
    stage(0, 0);
    for (int b = 0; b < nbatches; b++) {
        const int buf = b & 1;
        __syncthreads();                                   // batch b staged; everyone is done with buffer buf^1
        if (b + 1 < nbatches) stage(b + 1, buf ^ 1);
        const int hi = (int)max_lc - 1 - b * BATCH;
        const int cnt = min(BATCH, hi + 1);
        // entries of this batch with list position >= wmax cannot touch any pixel of this warp
        const int jskip = max(0, hi + 1 - (int)wmax);      // first batch entry with pos < wmax
        for (int j0 = jskip & ~31; j0 < cnt; j0 += 32) {
            bool hit = false;
            const int jl = j0 + lane;
            if (jl < cnt && jl >= jskip) {
                const float4 c = s_cull[buf][jl];
                hit = fabsf(c.x - cxw) <= c.z && fabsf(c.y - cyw) <= c.w;
            }
            unsigned hits = __ballot_sync(0xffffffffu, hit);
            while (hits) {
                const int j = j0 + __ffs(hits) - 1;
                hits &= hits - 1;
                const float4 r0 = s_r0[buf][j];
                const uint32_t pos = (uint32_t)(hi - j);
                const float4 r1 = s_r1[buf][j];
                const float dx = __fadd_rn(r0.x, -pxf), dy = __fadd_rn(r0.y, -pyf);
                const float power = eval_power(dx, dy, r0.z, r0.w, r1.x);
                const float G = expf(power);
                const float alpha = fminf(0.99f, __fmul_rn(r1.y, G));
                const bool skip = pos >= last_contributor || power > 0.0f || alpha < kAlphaMin;   // !inside => last_contributor == 0
                if (__all_sync(0xffffffffu, skip)) continue;

                float v[KR];
#pragma unroll
                for (int k = 0; k < KR; k++) v[k] = 0.0f;
                if (!skip) {
                    const float inv = __frcp_rn(1.0f - alpha);
                    T = T * inv;
                    const float w = alpha * T;
                    float dL_dalpha = 0.0f;
                    if constexpr (OPERATOR) {
                        const float4 col = s_rgb[buf][j];
                        const float c3[3] = {col.x, col.y, col.z};
#pragma unroll
                        for (int ch = 0; ch < C; ch++) {
                            accum_rec[ch] = last_alpha * last_color[ch] + (1.0f - last_alpha) * accum_rec[ch];
                            last_color[ch] = c3[ch];
                            dL_dalpha += (c3[ch] - accum_rec[ch]) * dpix[ch];
                        }
                        v[6] = w * dpix[0];
                        v[8] = w * dpix[1];
                        v[9] = w * dpix[2];
                        accum_rec_depth = last_alpha * last_depth + (1.0f - last_alpha) * accum_rec_depth;
                        last_depth = r1.w;
                        dL_dalpha += (r1.w - accum_rec_depth) * dpix_depth;
                        v[7] = w * dpix_depth;
                    } else {
                        accum_rec[0] = last_alpha * last_color[0] + (1.0f - last_alpha) * accum_rec[0];
                        last_color[0] = r1.z;
                        dL_dalpha += (r1.z - accum_rec[0]) * dpix[0];
                        v[5] = w * dpix[0];
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final * inv) * bg_dot;
                    const float dL_dG = r1.y * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * r0.z - gdy * r0.w;
                    const float dG_ddely = -gdy * r1.x - gdx * r0.w;
                    v[0] = dL_dG * dG_ddelx * ddelx_dx;
                    v[1] = dL_dG * dG_ddely * ddely_dy;
                    v[2] = -0.5f * gdx * dx * dL_dG;
                    v[3] = -0.5f * gdx * dy * dL_dG;
                    v[4] = -0.5f * gdy * dy * dL_dG;
                    if constexpr (OPERATOR) v[5] = G * dL_dalpha;
                }
                const float mine = warp_reduce_scatter<KR>(v, lane);
                if (owner) {
                    const uint32_t id = s_id[buf][j];
                    if constexpr (OPERATOR) {
                        float* dst = comp < 8 ? grad_f + 8 * (size_t)id + comp : reinterpret_cast<float*>(a.gradc + id) + (comp - 8);
                        atomicAdd(dst, mine);
                    } else {
                        atomicAdd(grad_f + 8 * (size_t)id + comp, mine);
                    }
                }
            }
        }
    }
}

void launch_blend_bwd_rgb(const BlendBwdArgs& a, cudaStream_t s) {
    dim3 grid(a.grid_x, a.grid_y, a.nviews);
    blend_bwd_kernel<3, true><<<grid, 256, 0, s>>>(a);
}
void launch_blend_bwd_gray(const BlendBwdArgs& a, cudaStream_t s) {
    dim3 grid(a.grid_x, a.grid_y, a.nviews);
    blend_bwd_kernel<1, false><<<grid, 256, 0, s>>>(a);
}

}  // namespace gsevt
