// Shared definitions for the libgsevt kernels (sm_100a only).
//
// Arithmetic note.  The reference (dgr/cuda_rasterizer/forward.cu) is compiled with nvcc's default
// --fmad=true, so which multiplies are fused into FFMAs is decided by the compiler.  The (tile, depth)
// sort keys and the tile ranges must be BIT-exact against it, so every operation that feeds
// depth / pixel centre / radius / conic is written here with explicit round-to-nearest intrinsics
// (__fmul_rn / __fadd_rn / __fmaf_rn are never contracted or re-associated) in the exact order the
// reference's sm_100a SASS uses (decoded with tools/sass_ssa.py, see DESIGN.md "Rounding contract"):
//     a*x + b*y + c*z + d   ->  add(fma(c,z, fma(a,x, rn(b*y))), d)
//     dot3(a, b)            ->  fma(a2,b2, fma(a0,b0, rn(a1*b1)))
// The CPU oracle mirrors the same sequence with fmaf().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GSEVT_TILE 16
#define GSEVT_TILE_PIX 256

namespace gsevt {

__device__ __constant__ const float kSH_C0 = 0.28209479177387814f;
__device__ __constant__ const float kSH_C1 = 0.4886025119029199f;
__device__ __constant__ const float kSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                                 -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ const float kSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                                 0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                                 -0.5900435899266435f};

// Grayscale weights of RenderFrame.get_intensity_frame (utils/render_camera/frame.py:39-43).
#define GSEVT_GRAY_R 0.2989f
#define GSEVT_GRAY_G 0.5870f
#define GSEVT_GRAY_B 0.1140f

// Per-view camera block, built on the host (operator) or by the pose kernel (engine).
struct ViewParams {
    float view[16];      // world->camera, column-major
    float proj[16];      // full projection (P * T_k), column-major
    float campos[3];
    float tanfovx, tanfovy;
    float focal_x, focal_y;
    int W, H;
    int grid_x, grid_y;
    // backward only
    float proj_a, proj_b, proj_e;   // projmatrix_raw[0], [5], [11]
    float vel[16];                  // vel_transofrm, column-major
    float vel_inv[16];              // vel_transofrm_inv
    float delta_time;
    float bg[3];
    float pad_;
};

// dot3 in the reference's contraction order.
__device__ __forceinline__ float dot3r(float a0, float a1, float a2, float b0, float b1, float b2) {
    return __fmaf_rn(a2, b2, __fmaf_rn(a0, b0, __fmul_rn(a1, b1)));
}
// m0*x + m1*y + m2*z + m3 in the reference's contraction order (auxiliary.h:58-77).
__device__ __forceinline__ float affine3r(float m0, float m1, float m2, float m3, float x, float y, float z) {
    return __fadd_rn(__fmaf_rn(z, m2, __fmaf_rn(x, m0, __fmul_rn(y, m1))), m3);
}

// computeCov3D (forward.cu:120-154) in the reference's rounding order.  q = (r, x, y, z) as given.
__device__ __forceinline__ void cov3d_from_scale_rot(float s0, float s1, float s2, float mod, float r, float x,
                                                     float y, float z, float* __restrict__ cov) {
    const float sx = __fmul_rn(s0, mod), sy = __fmul_rn(s1, mod), sz = __fmul_rn(s2, mod);
    const float xz = __fmul_rn(x, z), rx = __fmul_rn(r, x), rz = __fmul_rn(r, z);
    const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    const float h2 = __fmaf_rn(r, y, xz);     // x*z + r*y
    const float h6 = __fmaf_rn(-r, y, xz);    // x*z - r*y
    const float h5 = __fmaf_rn(y, z, -rx);    // y*z - r*x
    const float h7 = __fmaf_rn(y, z, rx);     // y*z + r*x
    const float h1 = __fmaf_rn(x, y, -rz);    // x*y - r*z
    const float h3 = __fmaf_rn(x, y, rz);     // x*y + r*z
    const float qxy = __fmaf_rn(x, x, yy), qyz = __fadd_rn(yy, zz), qxz = __fmaf_rn(x, x, zz);
    // R columns (GLM constructor order): R0 = (A0,A1,A2), R1 = (A3,A4,A5), R2 = (A6,A7,A8)
    const float A0 = __fadd_rn(-__fadd_rn(qyz, qyz), 1.0f), A1 = __fadd_rn(h1, h1), A2 = __fadd_rn(h2, h2);
    const float A3 = __fadd_rn(h3, h3), A4 = __fadd_rn(-__fadd_rn(qxz, qxz), 1.0f), A5 = __fadd_rn(h5, h5);
    const float A6 = __fadd_rn(h6, h6), A7 = __fadd_rn(h7, h7), A8 = __fadd_rn(-__fadd_rn(qxy, qxy), 1.0f);
    // M = S * R : M[c][r] = s_r * R[c][r]
    const float M00 = __fmul_rn(sx, A0), M01 = __fmul_rn(sy, A1), M02 = __fmul_rn(sz, A2);
    const float M10 = __fmul_rn(sx, A3), M11 = __fmul_rn(sy, A4), M12 = __fmul_rn(sz, A5);
    const float M20 = __fmul_rn(sx, A6), M21 = __fmul_rn(sy, A7), M22 = __fmul_rn(sz, A8);
    // Sigma = M^T M
    cov[0] = dot3r(M00, M01, M02, M00, M01, M02);
    cov[1] = dot3r(M00, M01, M02, M10, M11, M12);
    cov[2] = dot3r(M00, M01, M02, M20, M21, M22);
    cov[3] = dot3r(M10, M11, M12, M10, M11, M12);
    cov[4] = dot3r(M10, M11, M12, M20, M21, M22);
    cov[5] = dot3r(M20, M21, M22, M20, M21, M22);
}

// EWA intermediates shared by forward and backward (forward.cu:76-115, backward.cu:179-214).
struct Ewa {
    float tx, ty, tz;        // clamped view-space mean
    float txtz, tytz;        // unclamped ratios
    float J00, J02, J11, J12;
    float T00, T01, T02, T10, T11, T12;
    float a, b, c;           // 2D covariance (+0.3 on the diagonal)
};

__device__ __forceinline__ void ewa_forward(const float* __restrict__ v, float px, float py, float pz,
                                            float focal_x, float focal_y, float tanfovx, float tanfovy,
                                            const float* __restrict__ c3, Ewa& e) {
    const float tz = affine3r(v[2], v[6], v[10], v[14], px, py, pz);
    const float tx0 = affine3r(v[0], v[4], v[8], v[12], px, py, pz);
    const float ty0 = affine3r(v[1], v[5], v[9], v[13], px, py, pz);
    const float limx = __fmul_rn(tanfovx, 1.3f), limy = __fmul_rn(tanfovy, 1.3f);
    e.txtz = __fdiv_rn(tx0, tz);
    e.tytz = __fdiv_rn(ty0, tz);
    const float cx = fminf(fmaxf(e.txtz, -limx), limx);
    const float cy = fminf(fmaxf(e.tytz, -limy), limy);
    e.tx = __fmul_rn(cx, tz);
    e.ty = __fmul_rn(cy, tz);
    e.tz = tz;
    const float tz2 = __fmul_rn(tz, tz);
    e.J00 = __fdiv_rn(focal_x, tz);
    e.J02 = __fdiv_rn(__fmul_rn(-e.tx, focal_x), tz2);
    e.J11 = __fdiv_rn(focal_y, tz);
    e.J12 = __fdiv_rn(__fmul_rn(-e.ty, focal_y), tz2);
    // T = W * J (GLM): T0r = fma(W2r, J02, rn(W0r*J00)), T1r = fma(W2r, J12, rn(W1r*J11))
    e.T00 = __fmaf_rn(v[2], e.J02, __fmul_rn(v[0], e.J00));
    e.T01 = __fmaf_rn(v[6], e.J02, __fmul_rn(v[4], e.J00));
    e.T02 = __fmaf_rn(v[10], e.J02, __fmul_rn(v[8], e.J00));
    e.T10 = __fmaf_rn(v[2], e.J12, __fmul_rn(v[1], e.J11));
    e.T11 = __fmaf_rn(v[6], e.J12, __fmul_rn(v[5], e.J11));
    e.T12 = __fmaf_rn(v[10], e.J12, __fmul_rn(v[9], e.J11));
    // X = T^T Vrk^T ; cov = X T
    const float X00 = dot3r(e.T00, e.T01, e.T02, c3[0], c3[1], c3[2]);
    const float X10 = dot3r(e.T00, e.T01, e.T02, c3[1], c3[3], c3[4]);
    const float X20 = dot3r(e.T00, e.T01, e.T02, c3[2], c3[4], c3[5]);
    const float X01 = dot3r(e.T10, e.T11, e.T12, c3[0], c3[1], c3[2]);
    const float X11 = dot3r(e.T10, e.T11, e.T12, c3[1], c3[3], c3[4]);
    const float X21 = dot3r(e.T10, e.T11, e.T12, c3[2], c3[4], c3[5]);
    e.a = __fadd_rn(dot3r(e.T00, e.T01, e.T02, X00, X10, X20), 0.3f);
    e.b = dot3r(e.T00, e.T01, e.T02, X01, X11, X21);
    e.c = __fadd_rn(dot3r(e.T10, e.T11, e.T12, X01, X11, X21), 0.3f);
}

// ndc2Pix (auxiliary.h:41-44): evaluated in double, stored as float.
__device__ __forceinline__ float ndc2pix_r(float v, int S) {
    return (float)(fma((double)v + 1.0, (double)S, -1.0) * 0.5);
}

// getRect (auxiliary.h:46-56) for an integer radius.
__device__ __forceinline__ void tile_rect(float px, float py, int radius, int gx, int gy, int& x0, int& y0, int& x1,
                                          int& y1) {
    const float r = (float)radius;
    x0 = min(gx, max(0, (int)__fmul_rn(__fadd_rn(px, -r), 0.0625f)));
    y0 = min(gy, max(0, (int)__fmul_rn(__fadd_rn(py, -r), 0.0625f)));
    x1 = min(gx, max(0, (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(px, r), 16.0f), -1.0f), 0.0625f)));
    y1 = min(gy, max(0, (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(py, r), 16.0f), -1.0f), 0.0625f)));
}

// SH -> RGB (forward.cu:22-73), natural contraction (colours are not a bit-exact gate).
// sh is accessed through a functor so AoS and planar layouts share the code: sh(k, ch).
template <typename ShFn>
__device__ __forceinline__ void sh_to_rgb(int deg, float dx, float dy, float dz, ShFn sh, float rgb[3]) {
    const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
    const float x = dx * inv, y = dy * inv, z = dz * inv;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        float r = kSH_C0 * sh(0, ch);
        if (deg > 0) {
            r = r - kSH_C1 * y * sh(1, ch) + kSH_C1 * z * sh(2, ch) - kSH_C1 * x * sh(3, ch);
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + kSH_C2[0] * xy * sh(4, ch) + kSH_C2[1] * yz * sh(5, ch) +
                    kSH_C2[2] * (2.0f * zz - xx - yy) * sh(6, ch) + kSH_C2[3] * xz * sh(7, ch) +
                    kSH_C2[4] * (xx - yy) * sh(8, ch);
                if (deg > 2) {
                    r = r + kSH_C3[0] * y * (3.0f * xx - yy) * sh(9, ch) + kSH_C3[1] * xy * z * sh(10, ch) +
                        kSH_C3[2] * y * (4.0f * zz - xx - yy) * sh(11, ch) +
                        kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh(12, ch) +
                        kSH_C3[4] * x * (4.0f * zz - xx - yy) * sh(13, ch) + kSH_C3[5] * z * (xx - yy) * sh(14, ch) +
                        kSH_C3[6] * x * (xx - 3.0f * yy) * sh(15, ch);
                }
            }
        }
        rgb[ch] = r + 0.5f;
    }
}

// Real SH basis (with the reference's signs and constants) for direction (dx,dy,dz)/|.|: colour =
// sum_k basis[k] * sh[k] + 0.5, same term order as forward.cu:32-61.
__device__ __forceinline__ void sh_basis(int deg, float dx, float dy, float dz, float b[16]) {
    const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
    const float x = dx * inv, y = dy * inv, z = dz * inv;
    b[0] = kSH_C0;
#pragma unroll
    for (int k = 1; k < 16; k++) b[k] = 0.0f;
    if (deg > 0) {
        b[1] = -kSH_C1 * y; b[2] = kSH_C1 * z; b[3] = -kSH_C1 * x;
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = kSH_C2[0] * xy; b[5] = kSH_C2[1] * yz; b[6] = kSH_C2[2] * (2.0f * zz - xx - yy);
            b[7] = kSH_C2[3] * xz; b[8] = kSH_C2[4] * (xx - yy);
            if (deg > 2) {
                b[9] = kSH_C3[0] * y * (3.0f * xx - yy); b[10] = kSH_C3[1] * xy * z;
                b[11] = kSH_C3[2] * y * (4.0f * zz - xx - yy);
                b[12] = kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                b[13] = kSH_C3[4] * x * (4.0f * zz - xx - yy); b[14] = kSH_C3[5] * z * (xx - yy);
                b[15] = kSH_C3[6] * x * (xx - 3.0f * yy);
            }
        }
    }
}

// Gradient of the SH colour w.r.t. the (unnormalised) view direction, contracted with dL_dRGB and
// pushed through the normalisation: returns dL/dmean (backward.cu:37-139).
template <typename ShFn>
__device__ __forceinline__ void sh_dir_grad(int deg, float ox, float oy, float oz, ShFn sh, const float dRGB[3],
                                            float g[3]) {
    g[0] = g[1] = g[2] = 0.0f;
    if (deg < 1) return;
    const float len = sqrtf(ox * ox + oy * oy + oz * oz);
    const float x = ox / len, y = oy / len, z = oz / len;
    float ddx = 0.f, ddy = 0.f, ddz = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        float gx = -kSH_C1 * sh(3, ch), gy = -kSH_C1 * sh(1, ch), gz = kSH_C1 * sh(2, ch);
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            gx += kSH_C2[0] * y * sh(4, ch) + kSH_C2[2] * 2.f * -x * sh(6, ch) + kSH_C2[3] * z * sh(7, ch) +
                  kSH_C2[4] * 2.f * x * sh(8, ch);
            gy += kSH_C2[0] * x * sh(4, ch) + kSH_C2[1] * z * sh(5, ch) + kSH_C2[2] * 2.f * -y * sh(6, ch) +
                  kSH_C2[4] * 2.f * -y * sh(8, ch);
            gz += kSH_C2[1] * y * sh(5, ch) + kSH_C2[2] * 2.f * 2.f * z * sh(6, ch) + kSH_C2[3] * x * sh(7, ch);
            if (deg > 2) {
                gx += kSH_C3[0] * sh(9, ch) * 3.f * 2.f * xy + kSH_C3[1] * sh(10, ch) * yz +
                      kSH_C3[2] * sh(11, ch) * -2.f * xy + kSH_C3[3] * sh(12, ch) * -3.f * 2.f * xz +
                      kSH_C3[4] * sh(13, ch) * (-3.f * xx + 4.f * zz - yy) + kSH_C3[5] * sh(14, ch) * 2.f * xz +
                      kSH_C3[6] * sh(15, ch) * 3.f * (xx - yy);
                gy += kSH_C3[0] * sh(9, ch) * 3.f * (xx - yy) + kSH_C3[1] * sh(10, ch) * xz +
                      kSH_C3[2] * sh(11, ch) * (-3.f * yy + 4.f * zz - xx) + kSH_C3[3] * sh(12, ch) * -3.f * 2.f * yz +
                      kSH_C3[4] * sh(13, ch) * -2.f * xy + kSH_C3[5] * sh(14, ch) * -2.f * yz +
                      kSH_C3[6] * sh(15, ch) * -3.f * 2.f * xy;
                gz += kSH_C3[1] * sh(10, ch) * xy + kSH_C3[2] * sh(11, ch) * 4.f * 2.f * yz +
                      kSH_C3[3] * sh(12, ch) * 3.f * (2.f * zz - xx - yy) + kSH_C3[4] * sh(13, ch) * 4.f * 2.f * xz +
                      kSH_C3[5] * sh(14, ch) * (xx - yy);
            }
        }
        ddx += gx * dRGB[ch];
        ddy += gy * dRGB[ch];
        ddz += gz * dRGB[ch];
    }
    // dnormvdv (auxiliary.h:107-117)
    const float sum2 = ox * ox + oy * oy + oz * oz;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    g[0] = ((+sum2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * invsum32;
    g[1] = (-ox * oy * ddx + (sum2 - oy * oy) * ddy - oz * oy * ddz) * invsum32;
    g[2] = (-ox * oz * ddx - oy * oz * ddy + (sum2 - oz * oz) * ddz) * invsum32;
}

// Cooperative copy of n ViewParams from global to shared memory (ends with a block barrier).
__device__ __forceinline__ void load_views(ViewParams* dst, const ViewParams* src, int n) {
    const int words = n * (int)(sizeof(ViewParams) / 4);
    for (int i = threadIdx.x; i < words; i += blockDim.x)
        reinterpret_cast<uint32_t*>(dst)[i] = __ldg(reinterpret_cast<const uint32_t*>(src) + i);
    __syncthreads();
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace gsevt
