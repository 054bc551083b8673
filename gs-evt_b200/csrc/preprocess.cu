// Projection kernels: per-Gaussian cull / project / EWA covariance / conic / radius / tile rect /
// SH -> colour.  Restates preprocessCUDA (dgr/cuda_rasterizer/forward.cu:157-258) with the rounding
// contract of common.cuh so that depth bits, pixel centres, radii and tile rects are bit-identical.
//
//  * preprocess_aos_kernel  — operator path: caller's AoS tensors, one view.
//  * preprocess_map_kernel  — engine path: packed frozen map (SoA, planar SH, precomputed cov3D),
//                             BOTH views per thread so the 232 B/Gaussian map read happens once.
//  * strip_pretest_kernel + preprocess_map_list_kernel — engine path under the screen-tile split: a conservative 20-byte
//                             test of the whole map against this rank's strip of tile rows, then the exact projection of
//                             the survivors only (full warps, per-Gaussian SH copy) — see the comment above them.
// HBM-bound: grid = ceil(P/256) x 256 threads, all per-Gaussian loads coalesced (AoS inputs are
// staged through shared memory in 16-byte vectors; planar SH is read as 48 coalesced lines per warp).
#include <cstdlib>
#include <cstring>
#include "internal.h"

namespace gsevt {

struct ProjOut {
    int radius;
    int tiles;
    uint32_t rect;   // x0 | y0 << 8 | x1 << 16 | y1 << 24 (tile units; grids up to 255 x 255)
    float mx, my, depth;
    float A, B, C;
};

// Everything of preprocessCUDA between the frustum test and the colour evaluation.
__device__ __forceinline__ bool project_geometry(const ViewParams& vp, float px, float py, float pz,
                                                 const float* __restrict__ cov3D, ProjOut& o) {
    o.radius = 0;
    o.tiles = 0;
    o.rect = 0;
    const float depth = affine3r(vp.view[2], vp.view[6], vp.view[10], vp.view[14], px, py, pz);
    if (!(depth > 0.2f)) return false;  // in_frustum (auxiliary.h:154): p_view.z <= 0.2 is culled
    const float hx = affine3r(vp.proj[0], vp.proj[4], vp.proj[8], vp.proj[12], px, py, pz);
    const float hy = affine3r(vp.proj[1], vp.proj[5], vp.proj[9], vp.proj[13], px, py, pz);
    const float hw = affine3r(vp.proj[3], vp.proj[7], vp.proj[11], vp.proj[15], px, py, pz);
    const float p_w = __frcp_rn(__fadd_rn(hw, 0.0000001f));
    const float ndc_x = __fmul_rn(hx, p_w), ndc_y = __fmul_rn(hy, p_w);

    Ewa e;
    ewa_forward(vp.view, px, py, pz, vp.focal_x, vp.focal_y, vp.tanfovx, vp.tanfovy, cov3D, e);
    const float det = __fmaf_rn(e.a, e.c, -__fmul_rn(e.b, e.b));
    if (det == 0.0f) return false;
    const float det_inv = __frcp_rn(det);
    o.A = __fmul_rn(e.c, det_inv);
    o.B = __fmul_rn(e.b, -det_inv);
    o.C = __fmul_rn(e.a, det_inv);
    const float mid = __fmul_rn(__fadd_rn(e.a, e.c), 0.5f);
    const float s = __fsqrt_rn(fmaxf(__fmaf_rn(mid, mid, -det), 0.1f));
    const float lam = fmaxf(__fadd_rn(mid, s), __fadd_rn(mid, -s));
    const int radius = (int)ceilf(__fmul_rn(__fsqrt_rn(lam), 3.0f));
    o.mx = ndc2pix_r(ndc_x, vp.W);
    o.my = ndc2pix_r(ndc_y, vp.H);
    int x0, y0, x1, y1;
    tile_rect(o.mx, o.my, radius, vp.grid_x, vp.grid_y, x0, y0, x1, y1);
    const int area = (x1 - x0) * (y1 - y0);
    if (area == 0) return false;
    o.radius = radius;
    o.tiles = area;
    o.rect = (uint32_t)x0 | ((uint32_t)y0 << 8) | ((uint32_t)x1 << 16) | ((uint32_t)y1 << 24);
    o.depth = depth;
    return true;
}

// ------------------------------------------------------------------------------------------------
// Operator path
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) preprocess_aos_kernel(PreAosArgs a) {
    __shared__ float s_xyz[256 * 3];
    __shared__ float s_scale[256 * 3];
    __shared__ ViewParams s_vp;
    load_views(&s_vp, a.vp, 1);
    const int base = blockIdx.x * 256;
    const int n = min(256, a.P - base);
    // coalesced staging of the stride-3 arrays
    for (int i = threadIdx.x; i < n * 3; i += 256) {
        s_xyz[i] = __ldg(a.means3D + (size_t)base * 3 + i);
        if (a.scales) s_scale[i] = __ldg(a.scales + (size_t)base * 3 + i);
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= n) return;
    const int idx = base + t;
    const ViewParams& vp = s_vp;
    const float px = s_xyz[3 * t], py = s_xyz[3 * t + 1], pz = s_xyz[3 * t + 2];

    float cov[6];
    if (a.cov3D_precomp) {
#pragma unroll
        for (int k = 0; k < 6; k++) cov[k] = __ldg(a.cov3D_precomp + (size_t)idx * 6 + k);
    } else {
        const float4 q = __ldg(reinterpret_cast<const float4*>(a.rotations) + idx);
        cov3d_from_scale_rot(s_scale[3 * t], s_scale[3 * t + 1], s_scale[3 * t + 2], a.scale_modifier, q.x, q.y, q.z,
                             q.w, cov);
    }
    ProjOut o;
    const bool vis_depth = affine3r(vp.view[2], vp.view[6], vp.view[10], vp.view[14], px, py, pz) > 0.2f;
    // The reference writes cov3D for every Gaussian that passes the frustum test, before the later
    // early-outs (forward.cu:213); keep that so the backward can rely on it.
    if (vis_depth && !a.cov3D_precomp) {
#pragma unroll
        for (int k = 0; k < 6; k++) a.cov3D[(size_t)idx * 6 + k] = cov[k];
    }
    const bool ok = project_geometry(vp, px, py, pz, cov, o);
    a.radii_internal[idx] = ok ? o.radius : 0;
    if (a.radii_out) a.radii_out[idx] = ok ? o.radius : 0;
    a.tiles_touched[idx] = ok ? (uint32_t)o.tiles : 0u;
    if (!ok) return;

    float rgb[3];
    unsigned clampbits = 0;
    if (a.colors_precomp) {
        rgb[0] = __ldg(a.colors_precomp + (size_t)idx * 3);
        rgb[1] = __ldg(a.colors_precomp + (size_t)idx * 3 + 1);
        rgb[2] = __ldg(a.colors_precomp + (size_t)idx * 3 + 2);
    } else {
        const float* sh = a.shs + (size_t)idx * a.M * 3;
        sh_to_rgb(a.D, px - vp.campos[0], py - vp.campos[1], pz - vp.campos[2],
                  [&](int k, int ch) { return __ldg(sh + k * 3 + ch); }, rgb);
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            if (rgb[ch] < 0.0f) clampbits |= 1u << ch;
            rgb[ch] = fmaxf(rgb[ch], 0.0f);
        }
    }
    a.clamped[idx] = (uint8_t)clampbits;
    const float gray = GSEVT_GRAY_R * rgb[0] + GSEVT_GRAY_G * rgb[1] + GSEVT_GRAY_B * rgb[2];
    const float opacity = __ldg(a.opacities + idx);
    a.rec[2 * (size_t)idx] = make_float4(o.mx, o.my, o.A, o.B);
    a.rec[2 * (size_t)idx + 1] = make_float4(o.C, opacity, gray, o.depth);
    a.rgb4[idx] = make_float4(rgb[0], rgb[1], rgb[2], gray);
}

void launch_preprocess_aos(const PreAosArgs& a, cudaStream_t s) {
    if (a.P <= 0) return;
    preprocess_aos_kernel<<<(a.P + 255) / 256, 256, 0, s>>>(a);
}

// ------------------------------------------------------------------------------------------------
// Engine path: packed map, two views per thread.
// ------------------------------------------------------------------------------------------------
// Screen-tile split: can this Gaussian's tile rect reach tile rows [sy0, sy1) in this view?  A CONSERVATIVE test on 20
// bytes (position + the largest eigenvalue of the 3-D covariance, precomputed when the map is packed) and ~40
// instructions, run before the 24 bytes of covariance are even loaded: the rows a rect touches are
// [trunc((my - r) / 16), trunc((my + r + 15) / 16)) with r = ceil(3 sqrt(lambda_max(cov2D))), and
// lambda_max(J W S W^T J^T + 0.3 I) <= lambda_max(J J^T) lambda_max(S) + 0.3 for the rotation W, so
// r <= 3 sqrt(lambda_max(J J^T) smax2 + 0.3) + 1 with the Jacobian of ewa_forward (same clamp of x/z, y/z to 1.3 tan).
// The pixel row comes from the view-space point directly (the reference goes through the projection matrix: the two
// agree to ~1e-3 px) — a pixel of slack and 0.1 % on the eigenvalue cover the rounding.  "false" therefore implies the
// exact path would have produced an empty strip rect; on a rank of an 8-way split 80-90 % of the map takes this exit and
// the replicated part of the projection shrinks from two EWA projections per Gaussian to this test.
__device__ __forceinline__ bool strip_may_touch(const ViewParams& vp, float px, float py, float pz, float smax2, float row0_px,
                                                float row1_px) {
    const float tz = affine3r(vp.view[2], vp.view[6], vp.view[10], vp.view[14], px, py, pz);   // same value as the exact path
    if (!(tz > 0.2f)) return false;
    const float tx = vp.view[0] * px + vp.view[4] * py + vp.view[8] * pz + vp.view[12];
    const float ty = vp.view[1] * px + vp.view[5] * py + vp.view[9] * pz + vp.view[13];
    const float iz = 1.0f / tz;
    const float limx = 1.3f * vp.tanfovx, limy = 1.3f * vp.tanfovy;
    const float cx = fminf(fmaxf(tx * iz, -limx), limx), cy = fminf(fmaxf(ty * iz, -limy), limy);
    const float j00 = vp.focal_x * iz, j11 = vp.focal_y * iz, j02 = -cx * j00, j12 = -cy * j11;
    const float ja = j00 * j00 + j02 * j02, jc = j11 * j11 + j12 * j12, jb = j02 * j12;
    const float mid = 0.5f * (ja + jc), dif = 0.5f * (ja - jc);
    const float lamJ = mid + sqrtf(dif * dif + jb * jb);
    const float r = 3.0f * sqrtf(lamJ * smax2 * 1.001f + 0.3f) + 2.0f;
    const float my = (ty * iz / vp.tanfovy + 1.0f) * (0.5f * (float)vp.H) - 0.5f;
    return !(my + r + 15.0f < row0_px) && !(my - r >= row1_px);
}

// (Measured and dropped, round 2: running the bucket scatter of bucketbin.cu INSIDE this kernel — cursor atomics issued as
// soon as the two rects are known, SH evaluation while they are in flight, keys written at the end.  126 registers, and
// the stage took 0.149 ms against 0.081 + 0.055 ms for the two kernels: the atomics' latency was already hidden by the
// scatter kernel's own occupancy, and the heavy kernel lost more to its third live context than the launch saved.)
//
// One Gaussian, both views: everything from the covariance load to the records.  AOS selects where the SH coefficients
// come from: the planar copy (dense kernel: 48 coalesced lines per warp) or the per-Gaussian copy (split kernel, whose
// survivors are scattered over the map: 192 contiguous bytes per lane instead of 48 lone sectors).
template <int D, bool AOS>
__device__ __forceinline__ uint32_t project_store(const PreMapArgs& a, const ViewParams* s_vp, int idx, const float4 xo) {
    float cov[6];
    {
        const float4 c0 = __ldg(a.cov3D_a + idx);
        const float2 c1 = __ldg(a.cov3D_b + idx);
        cov[0] = c0.x; cov[1] = c0.y; cov[2] = c0.z; cov[3] = c0.w; cov[4] = c1.x; cov[5] = c1.y;
    }
    ProjOut o[2];
    bool ok[2];
#pragma unroll
    for (int v = 0; v < 2; v++) ok[v] = project_geometry(s_vp[v], xo.x, xo.y, xo.z, cov, o[v]);
    if (a.ctl) {
        // screen-tile split: keep only the part of the tile rect inside this engine's strip of tile rows
        // (the whole grid unless split, where this changes nothing)
        const uint32_t sy0 = (uint32_t)a.ctl->strip_y0, sy1 = (uint32_t)a.ctl->strip_y1;
#pragma unroll
        for (int v = 0; v < 2; v++) {
            const uint32_t r = o[v].rect;
            const uint32_t y0 = max(r >> 8 & 255u, sy0), y1 = min(r >> 24, sy1);
            ok[v] = ok[v] && y1 > y0;
            o[v].rect = (r & 0x00FF00FFu) | (y0 << 8) | (y1 << 24);
        }
    }
    // per pair, in index order: the tile rect (0 = not visible here) and the depth bits — the input of the bucket scatter
    // (bucketbin.cu), which reads them with unit stride and skips the zeros: no compaction pass in between
#pragma unroll
    for (int v = 0; v < 2; v++) {
        const size_t j = (size_t)v * a.P + idx;
        a.rect_raw[j] = ok[v] ? o[v].rect : 0u;
        a.depth_raw[j] = __float_as_uint(o[v].depth);
    }
    if (!ok[0] && !ok[1]) return 0u;
    // SH -> RGB for both views from ONE pass over the coefficients.  The degree is a template parameter so that all
    // (D+1)^2 * 3 loads are issued back to back, unconditionally.
    const size_t P = (size_t)a.P;
    constexpr int NB = (D + 1) * (D + 1);
    float coef[NB * 3];
    if constexpr (AOS) {
        const float4* sh4 = reinterpret_cast<const float4*>(a.sh_aos + (size_t)idx * 48);
#pragma unroll
        for (int q = 0; q < (NB * 3 + 3) / 4; q++) {
            const float4 c = __ldg(sh4 + q);
            coef[4 * q] = c.x;
            if (4 * q + 1 < NB * 3) coef[4 * q + 1] = c.y;
            if (4 * q + 2 < NB * 3) coef[4 * q + 2] = c.z;
            if (4 * q + 3 < NB * 3) coef[4 * q + 3] = c.w;
        }
    } else {
        const float* sh = a.sh_planar + idx;
#pragma unroll
        for (int k = 0; k < NB * 3; k++) coef[k] = __ldg(sh + (size_t)k * P);
    }
    float basis[2][16];
#pragma unroll
    for (int v = 0; v < 2; v++)
        sh_basis(D, xo.x - s_vp[v].campos[0], xo.y - s_vp[v].campos[1], xo.z - s_vp[v].campos[2], basis[v]);
    float rgb[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll
    for (int k = 0; k < NB; k++) {
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            rgb[0][ch] += basis[0][k] * coef[k * 3 + ch];
            rgb[1][ch] += basis[1][k] * coef[k * 3 + ch];
        }
    }
#pragma unroll
    for (int v = 0; v < 2; v++) {
        if (!ok[v]) continue;
        unsigned clampbits = 0;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const float c = rgb[v][ch] + 0.5f;
            if (c < 0.0f) clampbits |= 1u << ch;
            rgb[v][ch] = fmaxf(c, 0.0f);
        }
        const float gray = GSEVT_GRAY_R * rgb[v][0] + GSEVT_GRAY_G * rgb[v][1] + GSEVT_GRAY_B * rgb[v][2];
        const size_t j = (size_t)v * P + idx;
        a.clamped[j] = (uint8_t)clampbits;
        a.rec[2 * j] = make_float4(o[v].mx, o[v].my, o[v].A, o[v].B);
        a.rec[2 * j + 1] = make_float4(o[v].C, xo.w, gray, o[v].depth);
        // (the blend-backward accumulators grad8 are all-zero here: geom_bwd clears what it consumes)
    }
    return (ok[0] ? 1u : 0u) | (ok[1] ? 2u : 0u);
}

template <int D, int MINB>
__global__ void __launch_bounds__(256, MINB) preprocess_map_kernel(PreMapArgs a) {
    pdl_prologue();
    if (a.ctl && a.ctl->level_done) return;
    __shared__ ViewParams s_vp[2];
    load_views(s_vp, a.views, 2);
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= a.P) return;
    project_store<D, false>(a, s_vp, idx, __ldg(a.xyz_opacity + idx));
}

// Screen-tile split: the strip pre-test only pays when whole WARPS skip the projection, and in a map without spatial order
// every warp holds a few Gaussians of every strip.  Two kernels:
//   strip_pretest_kernel          a light streaming pass (20 bytes and ~40 instructions per Gaussian and view, 8 CTAs per
//                                 SM): the pre-test over the whole map, the failures' rect words zeroed on the spot, the
//                                 survivors of every 2048 consecutive Gaussians queued in shared memory and appended to a
//                                 global list with ONE reservation per CTA;
//   preprocess_map_list_kernel    the full projection over that list with full warps (grid-stride over the device-side
//                                 count; SH from the per-Gaussian copy, 192 contiguous bytes per lane), which also writes the
//                                 list of visible pairs the bucket scatter walks in split mode.
// On a rank of an 8-way split ~15 % of the map survives: the replicated part of the iteration falls from two EWA
// projections per Gaussian to the pre-test.
constexpr int PRETEST_PER_CTA = 2048;

__global__ void __launch_bounds__(256) strip_pretest_kernel(PreMapArgs a) {
    pdl_prologue();
    if (a.ctl->level_done) return;
    __shared__ ViewParams s_vp[2];
    __shared__ int s_queue[PRETEST_PER_CTA];
    __shared__ int s_n, s_base;
    if (threadIdx.x == 0) s_n = 0;
    load_views(s_vp, a.views, 2);
    __syncthreads();
    const float row0 = (float)(a.ctl->strip_y0 * GSEVT_TILE), row1 = (float)(a.ctl->strip_y1 * GSEVT_TILE);
    const int base = blockIdx.x * PRETEST_PER_CTA + threadIdx.x;
    const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
    for (int c0 = 0; c0 < PRETEST_PER_CTA / 256; c0 += 4) {
        float4 xo[4];
        float sm[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {   // four independent load pairs in flight per thread
            const int idx = base + (c0 + u) * 256;
            xo[u] = idx < a.P ? __ldg(a.xyz_opacity + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            sm[u] = idx < a.P ? __ldg(a.smax2 + idx) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int idx = base + (c0 + u) * 256;
            bool pass = false;
            if (idx < a.P) {
                pass = strip_may_touch(s_vp[0], xo[u].x, xo[u].y, xo[u].z, sm[u], row0, row1) ||
                       strip_may_touch(s_vp[1], xo[u].x, xo[u].y, xo[u].z, sm[u], row0, row1);
                if (!pass) {
                    a.rect_raw[idx] = 0u;
                    a.rect_raw[(size_t)a.P + idx] = 0u;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            if (bal) {
                int pos = 0;
                if ((threadIdx.x & 31) == 0) pos = atomicAdd(&s_n, __popc(bal));
                pos = __shfl_sync(0xffffffffu, pos, 0);
                if (pass) s_queue[pos + __popc(bal & lt)] = idx;
            }
        }
    }
    __syncthreads();
    const int n = s_n;
    if (threadIdx.x == 0) s_base = n ? (int)atomicAdd(a.surv_count, (uint32_t)n) : 0;
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += 256) a.surv_list[s_base + t] = (uint32_t)s_queue[t];
}

template <int D>
__global__ void __launch_bounds__(256, 2) preprocess_map_list_kernel(PreMapArgs a) {
    pdl_prologue();
    if (a.ctl->level_done) return;
    __shared__ ViewParams s_vp[2];
    __shared__ uint32_t s_wn[8];
    __shared__ uint32_t s_base;
    load_views(s_vp, a.views, 2);
    __syncthreads();
    const uint32_t n = *a.surv_count;
    for (uint32_t t0 = blockIdx.x * 256u; t0 < n; t0 += gridDim.x * 256u) {
        const uint32_t t = t0 + threadIdx.x;
        uint32_t vis = 0;
        int idx = 0;
        if (t < n) {
            idx = (int)__ldg(a.surv_list + t);
            vis = project_store<D, true>(a, s_vp, idx, __ldg(a.xyz_opacity + idx));
        }
        // the visible (view, Gaussian) pairs of the round go to the list the bucket scatter walks in split mode (one
        // reservation per round; the list is unordered, like the scatter itself)
        const uint32_t mine = (vis & 1u) + (vis >> 1);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= (unsigned)o) incl += y;
        }
        if ((threadIdx.x & 31) == 31) s_wn[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) { const uint32_t c = s_wn[w]; s_wn[w] = tot; tot += c; }
            s_base = tot ? atomicAdd(a.vis_count, tot) : 0u;
        }
        __syncthreads();
        uint32_t pos = s_base + s_wn[threadIdx.x >> 5] + incl - mine;
        if (vis & 1u) a.vis_list[pos++] = (uint32_t)idx;
        if (vis & 2u) a.vis_list[pos] = (uint32_t)a.P + (uint32_t)idx;
        __syncthreads();   // s_wn / s_base are rewritten by the next round
    }
}

size_t preprocess_map_raw_items(int P) { return (2 * (size_t)P + 1023) / 1024 * 1024; }

void launch_preprocess_map(const PreMapArgs& a, cudaStream_t s) {
    if (a.P <= 0) return;
    if (a.split_pretest) {
        launch_k(strip_pretest_kernel, dim3((a.P + PRETEST_PER_CTA - 1) / PRETEST_PER_CTA), dim3(256), 0, s, a);
        const int blocks = 148 * 4;   // 2 resident CTAs per SM, two waves; grid-stride over the survivors
        switch (a.D) {
            case 0: launch_k(preprocess_map_list_kernel<0>, dim3(blocks), dim3(256), 0, s, a); break;
            case 1: launch_k(preprocess_map_list_kernel<1>, dim3(blocks), dim3(256), 0, s, a); break;
            case 2: launch_k(preprocess_map_list_kernel<2>, dim3(blocks), dim3(256), 0, s, a); break;
            default: launch_k(preprocess_map_list_kernel<3>, dim3(blocks), dim3(256), 0, s, a); break;
        }
        return;
    }
    const int blocks = (a.P + 255) / 256;
    // SH degree 3 needs 105 registers without spills (2 CTAs per SM) or 80 with 56 B of spills (3 CTAs per SM): the
    // former measured 2.3 us faster (0.0985 vs 0.1008 ms for the stage, A/B/A/B on one box); GSEVT_PRE_MINB=3 selects the
    // latter for experiments.
    static const int minb = [] { const char* v = getenv("GSEVT_PRE_MINB"); return v && atoi(v) == 3 ? 3 : 2; }();
    switch (a.D) {
        case 0: launch_k(preprocess_map_kernel<0, 3>, dim3(blocks), dim3(256), 0, s, a); break;
        case 1: launch_k(preprocess_map_kernel<1, 3>, dim3(blocks), dim3(256), 0, s, a); break;
        case 2: launch_k(preprocess_map_kernel<2, 3>, dim3(blocks), dim3(256), 0, s, a); break;
        default:
            if (minb == 2) launch_k(preprocess_map_kernel<3, 2>, dim3(blocks), dim3(256), 0, s, a);
            else launch_k(preprocess_map_kernel<3, 3>, dim3(blocks), dim3(256), 0, s, a);
            break;
    }
}

// checkFrustum (rasterizer_impl.cu:54-66)
__global__ void mark_visible_kernel(int P, const float* __restrict__ means, const float* __restrict__ view,
                                    uint8_t* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float px = means[3 * (size_t)idx], py = means[3 * (size_t)idx + 1], pz = means[3 * (size_t)idx + 2];
    const float depth = affine3r(__ldg(view + 2), __ldg(view + 6), __ldg(view + 10), __ldg(view + 14), px, py, pz);
    present[idx] = depth > 0.2f ? 1 : 0;
}

void launch_mark_visible(int P, const float* means, const float* view, uint8_t* present, cudaStream_t s) {
    if (P <= 0) return;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means, view, present);
}

// Packs the activated AoS map tensors into the engine's resident layout (once per map).
__global__ void pack_map_kernel(int P, int M, const float* __restrict__ xyz, const float* __restrict__ scales,
                                const float* __restrict__ rots, const float* __restrict__ opac,
                                const float* __restrict__ shs, float mod, float4* __restrict__ xyz_opacity,
                                float4* __restrict__ cov_a, float2* __restrict__ cov_b, float* __restrict__ sh_planar,
                                float* __restrict__ sh_aos, float* __restrict__ smax2) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const size_t i = idx;
    xyz_opacity[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], opac[i]);
    {
        // largest eigenvalue of R S^2 R^T = (largest scale x modifier)^2, rounded up: the strip pre-test's radius bound
        const float sm = fmaxf(fmaxf(fabsf(scales[3 * i]), fabsf(scales[3 * i + 1])), fabsf(scales[3 * i + 2])) * fabsf(mod);
        smax2[i] = sm * sm * 1.0001f;
    }
    float cov[6];
    cov3d_from_scale_rot(scales[3 * i], scales[3 * i + 1], scales[3 * i + 2], mod, rots[4 * i], rots[4 * i + 1],
                         rots[4 * i + 2], rots[4 * i + 3], cov);
    cov_a[i] = make_float4(cov[0], cov[1], cov[2], cov[3]);
    cov_b[i] = make_float2(cov[4], cov[5]);
    for (int k = 0; k < 16; k++)
        for (int ch = 0; ch < 3; ch++) {
            const float c = k < M ? shs[(i * M + k) * 3 + ch] : 0.0f;
            sh_planar[(size_t)(k * 3 + ch) * P + i] = c;    // dense access (forward projection)
            sh_aos[i * 48 + k * 3 + ch] = c;                 // sparse access (backward, few Gaussians carry a gradient)
        }
}

void launch_pack_map(int P, int M, const float* xyz, const float* scales, const float* rots, const float* opac,
                     const float* shs, float mod, float4* xyz_opacity, float4* cov_a, float2* cov_b, float* sh_planar,
                     float* sh_aos, float* smax2, cudaStream_t s) {
    if (P <= 0) return;
    pack_map_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, M, xyz, scales, rots, opac, shs, mod, xyz_opacity, cov_a, cov_b,
                                                   sh_planar, sh_aos, smax2);
}

}  // namespace gsevt
