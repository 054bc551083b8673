// Geometry backward fused with the pose / velocity chain and its reduction.
//
// Restates computeCov2DCUDA (dgr/cuda_rasterizer/backward.cu:152-425), the backward preprocessCUDA
// (:497-655), computeColorFromSH backward (:21-147) and computeCov3D backward (:429-492), plus the
// (P,6)->(6,) sums of dgr/diff_gaussian_rasterization/__init__.py:163-169 — in ONE kernel that reads
// each visible Gaussian's blend gradients once and writes 12 floats per CTA.  The reference
// materialises dL_dtau / dL_dvel (P,6) and ten more zero-filled per-Gaussian tensors (352 B/Gaussian
// of memset, rasterize_points.cu:165-176) and reduces them with torch.sum afterwards.
//
// Per Gaussian (Appendix A.5 of SURVEY.md), with g_t = dL/dt through the 2D covariance, dWc_i the
// gradient w.r.t. column i of the view rotation, q = N^T dL/dmean2D, h(x) = R_vel^T x, p' = T_cur p,
// p_k = T_k p, t~ = clamped view-space mean, c_i / c'_i the columns of R_k / R_cur, g_sh the SH view
// direction gradient:
//   dtau[0:3] = h(g_t) + h(q) - g_sh                       (+ depth row)
//   dtau[3:6] = p' x h(g_t) + sum_i c'_i x h(dWc_i) + p' x h(q)
//   dvel[0:3] = dt*g_t + dt*q - g_sh                        (g_sh NOT scaled: reference quirk)
//   dvel[3:6] = dt*(t~ x g_t) + dt*sum_i c_i x dWc_i + dt*(p_k x q)
// HBM-bound; Gaussians whose blend gradients are all zero (culled, occluded) exit after one 32-byte
// read, so the 192-byte SH fetch only happens for Gaussians that actually influenced a pixel.
#include "internal.h"

namespace gsevt {

namespace {

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 cross(const V3& a, const V3& b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// R^T x for a column-major 4x4 (rotation block)
__device__ __forceinline__ V3 rot_t(const float* __restrict__ m, const V3& g) {
    return V3{m[0] * g.x + m[1] * g.y + m[2] * g.z, m[4] * g.x + m[5] * g.y + m[6] * g.z,
              m[8] * g.x + m[9] * g.y + m[10] * g.z};
}

struct SharedCam {
    ViewParams vp;
    float Rp[9];   // R_cur = R_vel_inv * R_k, column-major 3x3 (columns c'_i)
    float tp[3];   // t_cur
};

__device__ __forceinline__ void build_cur(SharedCam& c) {
    // T_CW_prime = T_vel_inv * T_CW (backward.cu:313-314; math.h:344-346)
    const float* vi = c.vp.vel_inv;
    const float* v = c.vp.view;
    for (int col = 0; col < 3; col++)
        for (int r = 0; r < 3; r++)
            c.Rp[col * 3 + r] = vi[0 + r] * v[4 * col + 0] + vi[4 + r] * v[4 * col + 1] + vi[8 + r] * v[4 * col + 2];
    for (int r = 0; r < 3; r++) c.tp[r] = vi[12 + r] + (vi[0 + r] * v[12] + vi[4 + r] * v[13] + vi[8 + r] * v[14]);
}

}  // namespace

// Work distribution.  Only a small fraction of the (view, Gaussian) pairs carries a gradient (a few per cent on
// the benchmark map: most Gaussians are culled, occluded or below the alpha threshold), and the per-pair
// work is heavy and divergent.  Each CTA therefore scans a chunk of GEOM_CHUNK consecutive pairs with fully
// coalesced loads, compacts the active ones into shared memory (deterministic order: ballot + block prefix)
// and then runs the chain densely over the compacted list — full warps instead of one live lane per warp.
#define GEOM_CHUNK 2048

template <bool ENGINE>
__global__ void __launch_bounds__(256) geom_bwd_kernel(GeomBwdArgs a) {
    pdl_prologue();
    if (a.ctl && a.ctl->level_done) return;
    __shared__ SharedCam s_cam[2];
    __shared__ float s_red[8][GSEVT_NPART];
    __shared__ uint32_t s_list[GEOM_CHUNK];
    __shared__ int s_wcount[8];
    __shared__ int s_count;
    for (int v = 0; v < a.nviews; v++) {
        load_views(&s_cam[v].vp, a.views + v, 1);
    }
    if (threadIdx.x < a.nviews) build_cur(s_cam[threadIdx.x]);
    __syncthreads();

    const int P = a.P;
    const long long n = (long long)a.nviews * P;
    const int warp_id = threadIdx.x >> 5, lane_id = threadIdx.x & 31;
    float acc[GSEVT_NPART];
#pragma unroll
    for (int k = 0; k < GSEVT_NPART; k++) acc[k] = 0.0f;

    // ENGINE: the work list was compacted by geom_compact_kernel (only a few per cent of the pairs carry a
    // gradient); a chunk is a slice of that list.  OPERATOR: every visible Gaussian is active (per-Gaussian outputs),
    // compacted per chunk of consecutive ids in shared memory.
    const long long n_work = ENGINE ? (long long)*a.active_count : n;
    constexpr int CH = ENGINE ? 256 : GEOM_CHUNK;   // engine: one dense item per thread, spread over many CTAs
    for (long long chunk = (long long)blockIdx.x * CH; chunk < n_work; chunk += (long long)gridDim.x * CH) {
        if constexpr (ENGINE) {
            const int cnt = (int)min((long long)CH, n_work - chunk);
            for (int i = threadIdx.x; i < cnt; i += 256) s_list[i] = a.active_list[chunk + i];
            if (threadIdx.x == 0) s_count = cnt;
            __syncthreads();
        } else {
        if (threadIdx.x == 0) s_count = 0;
        __syncthreads();
        for (int k = 0; k < GEOM_CHUNK / 256; k++) {
            const long long gid = chunk + k * 256 + threadIdx.x;
            const bool active = gid < n && a.radii[gid] > 0;
            const unsigned bal = __ballot_sync(0xffffffffu, active);
            if (lane_id == 0) s_wcount[warp_id] = __popc(bal);
            __syncthreads();
            int base = s_count;
            for (int w = 0; w < warp_id; w++) base += s_wcount[w];
            if (active) s_list[base + __popc(bal & ((1u << lane_id) - 1u))] = (uint32_t)(gid - chunk);
            __syncthreads();
            if (threadIdx.x == 0) {
                int t = 0;
                for (int w = 0; w < 8; w++) t += s_wcount[w];
                s_count += t;
            }
            __syncthreads();
        }
        }
        const int count = s_count;
    for (int li = threadIdx.x; li < count; li += 256) {
        const long long gid = ENGINE ? (long long)s_list[li] : chunk + s_list[li];
        float4 g0, g1;
        if constexpr (ENGINE) {
            // consume and clear: the engine's accumulators are all-zero between iterations (only the pairs on this
            // list were touched by the blend backward), so the projection never has to zero 32 B per visible pair
            float4* acc8 = const_cast<float4*>(a.grad8) + 2 * gid;
            g0 = acc8[0];
            g1 = acc8[1];
            acc8[0] = make_float4(0.f, 0.f, 0.f, 0.f);
            acc8[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            g0 = __ldg(a.grad8 + 2 * gid);
            g1 = __ldg(a.grad8 + 2 * gid + 1);
        }
        const int view = (int)(gid / P);
        const int idx = (int)(gid - (long long)view * P);
        float dmx = g0.x, dmy = g0.y, dA = g0.z, dB = g0.w, dC = g1.x;
        float dcol[3], ddepth = 0.0f, dop = 0.0f;
        const unsigned clampbits = a.clamped[gid];
        if constexpr (ENGINE) {
            const float dgray = g1.y;
            dcol[0] = GSEVT_GRAY_R * dgray; dcol[1] = GSEVT_GRAY_G * dgray; dcol[2] = GSEVT_GRAY_B * dgray;
        } else {
            const float2 gc = __ldg(a.gradc + idx);
            dop = g1.y; dcol[0] = g1.z; ddepth = g1.w; dcol[1] = gc.x; dcol[2] = gc.y;
        }
        const SharedCam& cam = s_cam[view];
        const ViewParams& vp = cam.vp;
        const float* v = vp.view;

        float mx, my, mz, cov[6];
        if constexpr (ENGINE) {
            const float4 xo = __ldg(a.xyz_opacity + idx);
            mx = xo.x; my = xo.y; mz = xo.z;
            const float4 c0 = __ldg(a.cov3D_a + idx);
            const float2 c1 = __ldg(a.cov3D_b + idx);
            cov[0] = c0.x; cov[1] = c0.y; cov[2] = c0.z; cov[3] = c0.w; cov[4] = c1.x; cov[5] = c1.y;
        } else {
            mx = __ldg(a.means3D + 3 * (size_t)idx); my = __ldg(a.means3D + 3 * (size_t)idx + 1);
            mz = __ldg(a.means3D + 3 * (size_t)idx + 2);
#pragma unroll
            for (int k = 0; k < 6; k++) cov[k] = __ldg(a.cov3D + 6 * (size_t)idx + k);
        }

        // ---- computeCov2DCUDA (backward.cu:179-279) ----
        Ewa e;
        ewa_forward(v, mx, my, mz, vp.focal_x, vp.focal_y, vp.tanfovx, vp.tanfovy, cov, e);
        const float limx = 1.3f * vp.tanfovx, limy = 1.3f * vp.tanfovy;
        const float x_grad_mul = (e.txtz < -limx || e.txtz > limx) ? 0.0f : 1.0f;
        const float y_grad_mul = (e.tytz < -limy || e.tytz > limy) ? 0.0f : 1.0f;
        const float ca = e.a, cb = e.b, cc = e.c;
        const float denom = ca * cc - cb * cb;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
        if (denom2inv != 0.0f) {
            dL_da = denom2inv * (-cc * cc * dA + 2 * cb * cc * dB + (denom - ca * cc) * dC);
            dL_dc = denom2inv * (-ca * ca * dC + 2 * ca * cb * dB + (denom - ca * cc) * dA);
            dL_db = denom2inv * 2 * (cb * cc * dA - (denom + 2 * cb * cb) * dB + ca * cb * dC);
        }
        // U0 = Vrk * T0, U1 = Vrk * T1
        const float U00 = e.T00 * cov[0] + e.T01 * cov[1] + e.T02 * cov[2];
        const float U01 = e.T00 * cov[1] + e.T01 * cov[3] + e.T02 * cov[4];
        const float U02 = e.T00 * cov[2] + e.T01 * cov[4] + e.T02 * cov[5];
        const float U10 = e.T10 * cov[0] + e.T11 * cov[1] + e.T12 * cov[2];
        const float U11 = e.T10 * cov[1] + e.T11 * cov[3] + e.T12 * cov[4];
        const float U12 = e.T10 * cov[2] + e.T11 * cov[4] + e.T12 * cov[5];
        const float dT00 = 2 * U00 * dL_da + U10 * dL_db, dT01 = 2 * U01 * dL_da + U11 * dL_db,
                    dT02 = 2 * U02 * dL_da + U12 * dL_db;
        const float dT10 = 2 * U10 * dL_dc + U00 * dL_db, dT11 = 2 * U11 * dL_dc + U01 * dL_db,
                    dT12 = 2 * U12 * dL_dc + U02 * dL_db;
        const float dJ00 = v[0] * dT00 + v[4] * dT01 + v[8] * dT02;
        const float dJ02 = v[2] * dT00 + v[6] * dT01 + v[10] * dT02;
        const float dJ11 = v[1] * dT10 + v[5] * dT11 + v[9] * dT12;
        const float dJ12 = v[2] * dT10 + v[6] * dT11 + v[10] * dT12;
        const float tz = 1.f / e.tz, tz2 = tz * tz, tz3 = tz2 * tz;
        const float hx = vp.focal_x, hy = vp.focal_y;
        V3 gt;
        gt.x = x_grad_mul * -hx * tz2 * dJ02;
        gt.y = y_grad_mul * -hy * tz2 * dJ12;
        gt.z = -hx * tz2 * dJ00 - hy * tz2 * dJ11 + (2 * hx * e.tx) * tz3 * dJ02 + (2 * hy * e.ty) * tz3 * dJ12;

        // ---- pose chain, part 1 (backward.cu:296-349, 360-423) ----
        const V3 tprime{cam.Rp[0] * mx + cam.Rp[3] * my + cam.Rp[6] * mz + cam.tp[0],
                        cam.Rp[1] * mx + cam.Rp[4] * my + cam.Rp[7] * mz + cam.tp[1],
                        cam.Rp[2] * mx + cam.Rp[5] * my + cam.Rp[8] * mz + cam.tp[2]};
        const V3 tcl{e.tx, e.ty, e.tz};
        const float dt = vp.delta_time;
        float tau[6], vel[6];
        {
            const V3 h = rot_t(vp.vel, gt);
            const V3 ch = cross(tprime, h), cg = cross(tcl, gt);
            tau[0] = h.x; tau[1] = h.y; tau[2] = h.z; tau[3] = ch.x; tau[4] = ch.y; tau[5] = ch.z;
            vel[0] = dt * gt.x; vel[1] = dt * gt.y; vel[2] = dt * gt.z;
            vel[3] = dt * cg.x; vel[4] = dt * cg.y; vel[5] = dt * cg.z;
        }
        {
            const V3 dWc[3] = {{e.J00 * dT00, e.J11 * dT10, e.J02 * dT00 + e.J12 * dT10},
                               {e.J00 * dT01, e.J11 * dT11, e.J02 * dT01 + e.J12 * dT11},
                               {e.J00 * dT02, e.J11 * dT12, e.J02 * dT02 + e.J12 * dT12}};
            V3 sp{0, 0, 0}, s{0, 0, 0};
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const V3 ci{v[4 * i], v[4 * i + 1], v[4 * i + 2]};
                const V3 cpi{cam.Rp[3 * i], cam.Rp[3 * i + 1], cam.Rp[3 * i + 2]};
                const V3 a1 = cross(cpi, rot_t(vp.vel, dWc[i]));
                const V3 a2 = cross(ci, dWc[i]);
                sp.x += a1.x; sp.y += a1.y; sp.z += a1.z;
                s.x += a2.x; s.y += a2.y; s.z += a2.z;
            }
            tau[3] += sp.x; tau[4] += sp.y; tau[5] += sp.z;
            vel[3] += dt * s.x; vel[4] += dt * s.y; vel[5] += dt * s.z;
        }

        // ---- backward preprocessCUDA: projection chain (backward.cu:531-626) ----
        const float* pj = vp.proj;
        const float mhx = pj[0] * mx + pj[4] * my + pj[8] * mz + pj[12];
        const float mhy = pj[1] * mx + pj[5] * my + pj[9] * mz + pj[13];
        const float mhw = pj[3] * mx + pj[7] * my + pj[11] * mz + pj[15];
        const float m_w = 1.0f / (mhw + 0.0000001f);
        const float al = m_w, be = -mhx * m_w * m_w, ga = -mhy * m_w * m_w;
        const V3 q{dmx * al * vp.proj_a, dmy * al * vp.proj_b, (dmx * be + dmy * ga) * vp.proj_e};
        const V3 pC{v[0] * mx + v[4] * my + v[8] * mz + v[12], v[1] * mx + v[5] * my + v[9] * mz + v[13],
                    v[2] * mx + v[6] * my + v[10] * mz + v[14]};
        {
            const V3 hq = rot_t(vp.vel, q);
            const V3 c1 = cross(tprime, hq), c2 = cross(pC, q);
            tau[0] += hq.x; tau[1] += hq.y; tau[2] += hq.z; tau[3] += c1.x; tau[4] += c1.y; tau[5] += c1.z;
            vel[0] += dt * q.x; vel[1] += dt * q.y; vel[2] += dt * q.z;
            vel[3] += dt * c2.x; vel[4] += dt * c2.y; vel[5] += dt * c2.z;
        }
        // depth row (backward.cu:632-644), unscaled and added to both
        if (ddepth != 0.0f) {
            tau[2] += ddepth; tau[3] += ddepth * pC.y; tau[4] += -ddepth * pC.x;
            vel[2] += ddepth; vel[3] += ddepth * pC.y; vel[4] += -ddepth * pC.x;
        }
        // ---- SH view-direction chain (backward.cu:128-146) ----
        float gsh[3] = {0.f, 0.f, 0.f};
        float dRGB[3];
        if (!a.colors_precomp) {
#pragma unroll
            for (int ch = 0; ch < 3; ch++) dRGB[ch] = (clampbits >> ch) & 1u ? 0.0f : dcol[ch];
            const float ox = mx - vp.campos[0], oy = my - vp.campos[1], oz = mz - vp.campos[2];
            if constexpr (ENGINE) {
                // sparse access: the AoS copy costs 6 sectors per Gaussian, the planar one (built for the dense
                // forward pass) would cost 48
                const float4* sh4 = reinterpret_cast<const float4*>(a.sh_aos) + 12 * (size_t)idx;
                float shv[48];
#pragma unroll
                for (int q = 0; q < 12; q++) {
                    const float4 t = __ldg(sh4 + q);
                    shv[4 * q] = t.x; shv[4 * q + 1] = t.y; shv[4 * q + 2] = t.z; shv[4 * q + 3] = t.w;
                }
                sh_dir_grad(a.D, ox, oy, oz, [&](int k, int ch) { return shv[k * 3 + ch]; }, dRGB, gsh);
            } else {
                const float* sh = a.shs + (size_t)idx * a.M * 3;
                sh_dir_grad(a.D, ox, oy, oz, [&](int k, int ch) { return __ldg(sh + k * 3 + ch); }, dRGB, gsh);
            }
            tau[0] -= gsh[0]; tau[1] -= gsh[1]; tau[2] -= gsh[2];
            vel[0] -= gsh[0]; vel[1] -= gsh[1]; vel[2] -= gsh[2];
        }
#pragma unroll
        for (int k = 0; k < 6; k++) {
            acc[k] += tau[k];
            acc[6 + k] += vel[k];
        }

        // ---- optional per-Gaussian outputs (operator only; the frozen map never asks) ----
        if constexpr (!ENGINE) {
            if (a.dL_dtau) for (int k = 0; k < 6; k++) a.dL_dtau[6 * (size_t)idx + k] = tau[k];
            if (a.dL_dvel) for (int k = 0; k < 6; k++) a.dL_dvel[6 * (size_t)idx + k] = vel[k];
            if (a.dL_dmeans2D) {
                a.dL_dmeans2D[3 * (size_t)idx] = dmx; a.dL_dmeans2D[3 * (size_t)idx + 1] = dmy;
            }
            if (a.dL_dopacity) a.dL_dopacity[idx] = dop;
            if (a.dL_dcolors) for (int ch = 0; ch < 3; ch++) a.dL_dcolors[3 * (size_t)idx + ch] = dcol[ch];
            if (a.dL_dmeans3D) {
                // cov part (backward.cu:353-358) + projection part (:539-548) + depth (:632-635) + SH (:139)
                const float mul1 = mhx * m_w * m_w, mul2 = mhy * m_w * m_w;
                float gx = v[0] * gt.x + v[1] * gt.y + v[2] * gt.z;
                float gy = v[4] * gt.x + v[5] * gt.y + v[6] * gt.z;
                float gz = v[8] * gt.x + v[9] * gt.y + v[10] * gt.z;
                gx += (pj[0] * m_w - pj[3] * mul1) * dmx + (pj[1] * m_w - pj[3] * mul2) * dmy;
                gy += (pj[4] * m_w - pj[7] * mul1) * dmx + (pj[5] * m_w - pj[7] * mul2) * dmy;
                gz += (pj[8] * m_w - pj[11] * mul1) * dmx + (pj[9] * m_w - pj[11] * mul2) * dmy;
                gx += ddepth * v[2] + gsh[0]; gy += ddepth * v[6] + gsh[1]; gz += ddepth * v[10] + gsh[2];
                a.dL_dmeans3D[3 * (size_t)idx] = gx; a.dL_dmeans3D[3 * (size_t)idx + 1] = gy;
                a.dL_dmeans3D[3 * (size_t)idx + 2] = gz;
            }
            float dcov[6] = {0, 0, 0, 0, 0, 0};
            if (denom2inv != 0.0f) {  // backward.cu:232-242
                dcov[0] = e.T00 * e.T00 * dL_da + e.T00 * e.T10 * dL_db + e.T10 * e.T10 * dL_dc;
                dcov[3] = e.T01 * e.T01 * dL_da + e.T01 * e.T11 * dL_db + e.T11 * e.T11 * dL_dc;
                dcov[5] = e.T02 * e.T02 * dL_da + e.T02 * e.T12 * dL_db + e.T12 * e.T12 * dL_dc;
                dcov[1] = 2 * e.T00 * e.T01 * dL_da + (e.T00 * e.T11 + e.T01 * e.T10) * dL_db + 2 * e.T10 * e.T11 * dL_dc;
                dcov[2] = 2 * e.T00 * e.T02 * dL_da + (e.T00 * e.T12 + e.T02 * e.T10) * dL_db + 2 * e.T10 * e.T12 * dL_dc;
                dcov[4] = 2 * e.T02 * e.T01 * dL_da + (e.T01 * e.T12 + e.T02 * e.T11) * dL_db + 2 * e.T11 * e.T12 * dL_dc;
            }
            if (a.dL_dcov3D) for (int k = 0; k < 6; k++) a.dL_dcov3D[6 * (size_t)idx + k] = dcov[k];
            if (a.dL_dsh && !a.colors_precomp) {
                // backward.cu:48-98
                const float ox = mx - vp.campos[0], oy = my - vp.campos[1], oz = mz - vp.campos[2];
                const float len = sqrtf(ox * ox + oy * oy + oz * oz);
                const float x = ox / len, y = oy / len, z = oz / len;
                float basis[16];
                basis[0] = kSH_C0;
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                basis[1] = -kSH_C1 * y; basis[2] = kSH_C1 * z; basis[3] = -kSH_C1 * x;
                basis[4] = kSH_C2[0] * xy; basis[5] = kSH_C2[1] * yz; basis[6] = kSH_C2[2] * (2.f * zz - xx - yy);
                basis[7] = kSH_C2[3] * xz; basis[8] = kSH_C2[4] * (xx - yy);
                basis[9] = kSH_C3[0] * y * (3.f * xx - yy); basis[10] = kSH_C3[1] * xy * z;
                basis[11] = kSH_C3[2] * y * (4.f * zz - xx - yy); basis[12] = kSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                basis[13] = kSH_C3[4] * x * (4.f * zz - xx - yy); basis[14] = kSH_C3[5] * z * (xx - yy);
                basis[15] = kSH_C3[6] * x * (xx - 3.f * yy);
                const int nb = (a.D + 1) * (a.D + 1);
                float* out = a.dL_dsh + (size_t)idx * a.M * 3;
                for (int k = 0; k < nb && k < a.M; k++)
                    for (int ch = 0; ch < 3; ch++) out[k * 3 + ch] = basis[k] * dRGB[ch];
            }
            if (a.scales && (a.dL_dscales || a.dL_drotations)) {
                // computeCov3D backward (backward.cu:429-492), q used as given
                const float4 qq = __ldg(reinterpret_cast<const float4*>(a.rotations) + idx);
                const float r = qq.x, x = qq.y, y = qq.z, z = qq.w;
                const float sx = a.scale_modifier * __ldg(a.scales + 3 * (size_t)idx);
                const float sy = a.scale_modifier * __ldg(a.scales + 3 * (size_t)idx + 1);
                const float sz = a.scale_modifier * __ldg(a.scales + 3 * (size_t)idx + 2);
                // R columns as in forward; Rt rows
                const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                                       {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                                       {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
                // M = S * R (GLM): M[c][r] = s_r * R[c][r]
                const float s3[3] = {sx, sy, sz};
                float Mm[3][3];
                for (int c = 0; c < 3; c++) for (int rr = 0; rr < 3; rr++) Mm[c][rr] = s3[rr] * R[c][rr];
                // dL_dSigma (GLM column-major symmetric)
                const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                                        {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                        {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
                // dL_dM = 2 * M * dL_dSigma : (A*B)[c][r] = sum_k A[k][r] * B[c][k]
                float dM[3][3];
                for (int c = 0; c < 3; c++)
                    for (int rr = 0; rr < 3; rr++)
                        dM[c][rr] = 2.0f * (Mm[0][rr] * dS[c][0] + Mm[1][rr] * dS[c][1] + Mm[2][rr] * dS[c][2]);
                // Rt = transpose(R): Rt[c][r] = R[r][c]; dL_dMt[c][r] = dM[r][c]
                float dscale[3];
                for (int i = 0; i < 3; i++) dscale[i] = R[0][i] * dM[0][i] + R[1][i] * dM[1][i] + R[2][i] * dM[2][i];
                if (a.dL_dscales) for (int i = 0; i < 3; i++) a.dL_dscales[3 * (size_t)idx + i] = dscale[i];
                float dMt[3][3];
                for (int c = 0; c < 3; c++) for (int rr = 0; rr < 3; rr++) dMt[c][rr] = dM[rr][c] * s3[c];
                if (a.dL_drotations) {
                    float4 dq;
                    dq.x = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
                    dq.y = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) -
                           4 * x * (dMt[2][2] + dMt[1][1]);
                    dq.z = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) -
                           4 * y * (dMt[2][2] + dMt[0][0]);
                    dq.w = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) -
                           4 * z * (dMt[1][1] + dMt[0][0]);
                    reinterpret_cast<float4*>(a.dL_drotations)[idx] = dq;
                }
            }
        }
    }
        __syncthreads();
    }

    // block reduction of the 12 pose components -> partials[blockIdx.x][12] (fixed order, no atomics)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < GSEVT_NPART; k++) {
        const float s = warp_sum(acc[k]);
        if (lane == 0) s_red[warp][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < GSEVT_NPART) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; w++) s += s_red[w][threadIdx.x];
        a.partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;   // [12][nblocks]
    }
}

int geom_bwd_blocks(int P, int nviews) {
    const long long n = (long long)P * nviews;
    long long b = (n + GEOM_CHUNK - 1) / GEOM_CHUNK;
    const long long cap = 148 * 8;  // one wave of 8 resident CTAs per SM
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

void launch_geom_bwd_aos(const GeomBwdArgs& a, cudaStream_t s) {
    geom_bwd_kernel<false><<<geom_bwd_blocks(a.P, a.nviews), 256, 0, s>>>(a);
}
void launch_geom_bwd_map(const GeomBwdArgs& a, cudaStream_t s) {
    launch_k(geom_bwd_kernel<true>, dim3(geom_bwd_blocks(a.P, a.nviews)), dim3(256), 0, s, a);
}

// Streaming compaction: list of (view, Gaussian) pairs with a non-zero blend gradient.  Every CTA owns a contiguous
// slice of the pairs, collects its hits in shared memory and reserves its part of the list with ONE global atomic
// (a warp-aggregated atomic per hit warp serialises tens of thousands of adds on a single L2 address).
constexpr int GC_PER_CTA = 2048;
__global__ void __launch_bounds__(256) geom_compact_kernel(int n, int per_cta, const uint32_t* __restrict__ rect_raw,
                                                           const float4* __restrict__ grad8, uint32_t* __restrict__ list,
                                                           uint32_t* __restrict__ count, const EngineCtl* __restrict__ ctl) {
    pdl_prologue();
    if (ctl && ctl->level_done) return;
    __shared__ uint32_t s_list[GC_PER_CTA];
    __shared__ uint32_t s_n, s_base;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int first = blockIdx.x * per_cta;
    for (int k0 = 0; k0 < per_cta; k0 += 4 * 256) {
        int gid[4];
        uint32_t rad[4];   // tile rect of the pair, 0 = not visible
#pragma unroll
        for (int u = 0; u < 4; u++) {
            gid[u] = first + k0 + u * 256 + threadIdx.x;
            rad[u] = (k0 + u * 256 < per_cta && gid[u] < n) ? __ldg(rect_raw + gid[u]) : 0u;
        }
        float4 g0[4];
        float2 g1[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (rad[u] != 0u) {
                g0[u] = __ldg(grad8 + 2 * (size_t)gid[u]);
                g1[u] = __ldg(reinterpret_cast<const float2*>(grad8 + 2 * (size_t)gid[u] + 1));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const bool active = rad[u] != 0u && (g0[u].x != 0.f || g0[u].y != 0.f || g0[u].z != 0.f || g0[u].w != 0.f ||
                                               g1[u].x != 0.f || g1[u].y != 0.f);
            const unsigned bal = __ballot_sync(0xffffffffu, active);
            if (bal) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&s_n, (uint32_t)__popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (active) s_list[base + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)gid[u];
            }
        }
    }
    __syncthreads();
    const uint32_t cnt = s_n;
    if (cnt == 0) return;
    if (threadIdx.x == 0) s_base = atomicAdd(count, cnt);
    __syncthreads();
    const uint32_t base = s_base;
    for (uint32_t i = threadIdx.x; i < cnt; i += 256) list[base + i] = s_list[i];
}
void launch_geom_compact(int n_pairs, const uint32_t* rect_raw, const float4* grad8, uint32_t* list, uint32_t* count, const EngineCtl* ctl,
                         cudaStream_t s) {
    if (n_pairs <= 0) return;
    // slices of <= GC_PER_CTA pairs, at least ~4 CTAs per SM
    int per = (n_pairs + 148 * 4 - 1) / (148 * 4);
    per = (per + 255) / 256 * 256;
    if (per > GC_PER_CTA) per = GC_PER_CTA;
    const int blocks = (n_pairs + per - 1) / per;
    launch_k(geom_compact_kernel, dim3(blocks), dim3(256), 0, s, n_pairs, per, rect_raw, grad8, list, count, ctl);
}

// partials[12][nblocks] -> out12, fixed summation order, double accumulation.
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int nblocks, float* __restrict__ out12) {
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k >= GSEVT_NPART) return;
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += (double)partials[(size_t)k * nblocks + b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out12[k] = (float)s;
}
void launch_reduce_partials(const float* partials, int nblocks, float* out12, cudaStream_t s) {
    reduce_partials_kernel<<<1, 32 * GSEVT_NPART, 0, s>>>(partials, nblocks, out12);
}

}  // namespace gsevt
