// Event-frame construction and the normalised intensity-change loss.
//
// Events (replaces the host-side numpy + OpenCV code of utils/event_camera/event.py:116-128 and the
// pyramid of utils/tracker.py:78-91):
//   E0 polarity scatter-add (int32 atomics: order-independent, bit-exact)
//   E1 cv2.undistort  == fixed-point (1/32 px) bilinear remap, constant-0 border
//   E2 cv2.GaussianBlur(9x9, sigma 0) == separable [4,13,30,51,60,51,30,13,4]/256, replicate border,
//      row pass  s = k0*x[-4]; s = fma(x[i], k[i], s)            (left to right)
//      col pass  t = k4*x[0];  t = fma(x[+j] + x[-j], k[4+j], t) (j = 1..4)
//   E3 cv2.normalize(L2) == x * float(1 / sqrt(sum_fp64 x^2))
//   E4 abs, and the INTER_NEAREST pyramid (source index floor(dst * src_size / dst_size): frame[::2^l, ::2^l] for
//      sizes divisible by 2^l)
// Loss (frame.py:86-92 + tracker.py:93-103): d = gray_next - gray_last, u = d/||d||,
//   L = ||u - E||  (signed)  or  || |u| - |E| ||  (unsigned).  One pass computes
//   Sd2 = sum d^2, S2 = sum d*E (or |d||E|), SE2 = sum E^2 in double; then
//   L^2 = 1 - 2*S2/n + SE2,  dL/dd = alpha*d - beta*E_eff,  alpha = S2/(n^3 L),  beta = 1/(L n)
// (closed form of autograd through norm / div / abs / sub / norm; derivation in DESIGN.md).
#include "internal.h"
#include "split_comm.cuh"
#include "loss_finish.cuh"

namespace gsevt {

// ---- E0 ------------------------------------------------------------------------------------------
__global__ void event_accumulate_kernel(const int16_t* __restrict__ x, const int16_t* __restrict__ y,
                                        const uint8_t* __restrict__ p, int n, int W, int H, int* __restrict__ counts,
                                        int* __restrict__ oob) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // numpy indexing (event.py:118-120: frame[y, x] += ...): a negative index in [-size, -1] counts from the end, anything
    // else raises IndexError — here: flagged and dropped
    int xi = x[i], yi = y[i];
    if (xi < 0) xi += W;
    if (yi < 0) yi += H;
    if (xi < 0 || xi >= W || yi < 0 || yi >= H) {
        if (oob) *oob = 1;
        return;
    }
    atomicAdd(counts + yi * W + xi, p[i] ? 1 : -1);
}
void launch_event_accumulate(const int16_t* x, const int16_t* y, const uint8_t* p, int n, int W, int H, int* counts,
                             int* oob, cudaStream_t s) {
    if (n <= 0) return;
    event_accumulate_kernel<<<(n + 255) / 256, 256, 0, s>>>(x, y, p, n, W, H, counts, oob);
}

// ---- E1 ------------------------------------------------------------------------------------------
__global__ void event_undistort_kernel(const int* __restrict__ counts, const int* __restrict__ map_ix,
                                       const int* __restrict__ map_iy, int W, int H, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    const int ix = map_ix[i], iy = map_iy[i];
    const int sx = ix >> 5, sy = iy >> 5;
    const float fx = (float)(ix & 31) * 0.03125f, fy = (float)(iy & 31) * 0.03125f;
    auto tap = [&](int yy, int xx) -> float {
        return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (float)counts[yy * W + xx] : 0.0f;
    };
    const float w00 = __fmul_rn(1.0f - fy, 1.0f - fx), w01 = __fmul_rn(1.0f - fy, fx);
    const float w10 = __fmul_rn(fy, 1.0f - fx), w11 = __fmul_rn(fy, fx);
    // small integers times multiples of 1/1024: every product and partial sum is exact in fp32
    float acc = __fmul_rn(tap(sy, sx), w00);
    acc = __fadd_rn(acc, __fmul_rn(tap(sy, sx + 1), w01));
    acc = __fadd_rn(acc, __fmul_rn(tap(sy + 1, sx), w10));
    acc = __fadd_rn(acc, __fmul_rn(tap(sy + 1, sx + 1), w11));
    out[i] = acc;
}

// ---- E2 ------------------------------------------------------------------------------------------
// cv2.GaussianBlur(frame, (k, k), 0, BORDER_REPLICATE) for odd k <= 31: OpenCV's own coefficients (gauss_taps.inc), the row
// pass accumulated left to right with FMA, the column pass in the symmetric form (centre, then the pairs outwards) —
// bit-exact against cv2 on event frames for every odd k (tests/test_event_oracle.py).
#include "gauss_taps.inc"
struct GaussTaps { float c[16]; int r; };   // c[j]: coefficient at distance j from the centre, r = k / 2

__global__ void event_blur_row_kernel(const float* __restrict__ in, int W, int H, float* __restrict__ out, GaussTaps g) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    const int y = i / W, x = i - y * W;
    const float* row = in + (size_t)y * W;
    const int r = g.r;
    float s = __fmul_rn(row[max(x - r, 0)], g.c[r]);
    for (int k = 1; k <= 2 * r; k++) s = __fmaf_rn(row[min(max(x - r + k, 0), W - 1)], g.c[abs(k - r)], s);
    out[i] = s;
}
__global__ void event_blur_col_kernel(const float* __restrict__ in, int W, int H, float* __restrict__ out, GaussTaps g) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    const int y = i / W, x = i - y * W;
    float t = __fmul_rn(in[i], g.c[0]);
    for (int j = 1; j <= g.r; j++) {
        const float a = in[(size_t)min(y + j, H - 1) * W + x], b = in[(size_t)max(y - j, 0) * W + x];
        t = __fmaf_rn(__fadd_rn(a, b), g.c[j], t);
    }
    out[i] = t;
}

// ---- E3 ------------------------------------------------------------------------------------------
#define EV_NB 64
__global__ void event_sumsq_kernel(const float* __restrict__ in, int n, double* __restrict__ partials) {
    __shared__ double s_w[8];
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double v = (double)in[i];
        s += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += s_w[w];
        partials[blockIdx.x] = t;
    }
}
// E3 + E4 + pyramid
__global__ void event_finish_kernel(const float* __restrict__ in, const double* __restrict__ partials, int W, int H,
                                    int levels, float* __restrict__ sign_out, float* __restrict__ unsign_out) {
    __shared__ float s_scale;
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int b = 0; b < EV_NB; b++) t += partials[b];
        const double nrm = sqrt(t);
        s_scale = nrm > 2.220446049250313e-16 ? (float)(1.0 / nrm) : 0.0f;
    }
    __syncthreads();
    const float scale = s_scale;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    const float v = __fmul_rn(in[i], scale);
    sign_out[i] = v;
    unsign_out[i] = fabsf(v);
    // pyramid levels: cv2.resize(frame, (int(W s), int(H s)), INTER_NEAREST) (tracker.py:87-88) samples source column
    // min(floor(dx * (1 / (Wl / W))), W - 1) — evaluated in double like OpenCV's resizeNN; this is dx << l whenever the size
    // is a multiple of 2^l, and drifts from it otherwise (e.g. 346 x 260 at level 2)
    size_t off = (size_t)W * H;
    for (int l = 1; l < levels; l++) {
        const int Wl = W >> l, Hl = H >> l;
        if (i < Wl * Hl) {
            const int dy = i / Wl, dx = i - dy * Wl;
            const double ifx = 1.0 / ((double)Wl / (double)W), ify = 1.0 / ((double)Hl / (double)H);
            const int sx = min((int)floor((double)dx * ifx), W - 1), sy = min((int)floor((double)dy * ify), H - 1);
            const float u = __fmul_rn(in[(size_t)sy * W + sx], scale);
            sign_out[off + i] = u;
            unsign_out[off + i] = fabsf(u);
        }
        off += (size_t)Wl * Hl;
    }
}

void launch_event_frame(const int* counts, const int* map_ix, const int* map_iy, int W, int H, int levels, int ksize,
                        float* sign_out, float* unsign_out, float* scratch, double* dscratch, cudaStream_t s) {
    const int n = W * H, nb = (n + 255) / 256;
    float* a = scratch;
    float* b = scratch + n;
    GaussTaps g;
    g.r = ksize / 2;
    for (int j = 0; j < 16; j++) g.c[j] = kGaussTaps[g.r][j];
    event_undistort_kernel<<<nb, 256, 0, s>>>(counts, map_ix, map_iy, W, H, a);
    event_blur_row_kernel<<<nb, 256, 0, s>>>(a, W, H, b, g);
    event_blur_col_kernel<<<nb, 256, 0, s>>>(b, W, H, a, g);
    event_sumsq_kernel<<<EV_NB, 256, 0, s>>>(a, n, dscratch);
    event_finish_kernel<<<nb, 256, 0, s>>>(a, dscratch, W, H, levels, sign_out, unsign_out);
}

// ---- loss ----------------------------------------------------------------------------------------
int loss_blocks(int HW) {
    int b = (HW + 1023) / 1024;
    if (b > 296) b = 296;   // two CTAs per SM
    if (b < 1) b = 1;
    return b;
}

// partials: [nblocks][3] doubles followed by one unsigned ticket counter (zero on entry, reset on exit).
__global__ void __launch_bounds__(256) loss_stats_kernel(const float* __restrict__ gray, const float* __restrict__ ev,
                                                         int HW, int pix0, int npix, EngineCtl* __restrict__ ctl,
                                                         double* __restrict__ partials, int nblocks,
                                                         uint32_t* __restrict__ zero_me, SplitComm* comm, int* host_flag) {
    pdl_prologue();
    if (ctl->level_done) return;
    __shared__ double s_w[8][3];
    __shared__ double s_x[8];
    __shared__ bool s_last;
    const bool sgn = ctl->loss_signed != 0;
    double sd2 = 0.0, s2 = 0.0, se2 = 0.0;
    for (int i = pix0 + blockIdx.x * 256 + threadIdx.x; i < pix0 + npix; i += gridDim.x * 256) {
        const float d = gray[(size_t)HW + i] - gray[i];
        const float E = ev[i];
        sd2 += (double)d * (double)d;
        s2 += sgn ? (double)d * (double)E : (double)fabsf(d) * (double)fabsf(E);
        se2 += (double)E * (double)E;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sd2 += __shfl_xor_sync(0xffffffffu, sd2, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        se2 += __shfl_xor_sync(0xffffffffu, se2, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_w[warp][0] = sd2; s_w[warp][1] = s2; s_w[warp][2] = se2; }
    __syncthreads();
    unsigned int* ticket = reinterpret_cast<unsigned int*>(partials + (size_t)nblocks * 3);
    if (threadIdx.x == 0) {
        double t0 = 0, t1 = 0, t2 = 0;
        for (int w = 0; w < 8; w++) { t0 += s_w[w][0]; t1 += s_w[w][1]; t2 += s_w[w][2]; }
        partials[blockIdx.x * 3 + 0] = t0;
        partials[blockIdx.x * 3 + 1] = t1;
        partials[blockIdx.x * 3 + 2] = t2;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == (unsigned)(gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) return;
    __threadfence();
    // last CTA: warp 0 sums the block partials in a fixed order (lane-strided, then a butterfly): deterministic
    double t0 = 0, t1 = 0, t2 = 0;
    for (int b = threadIdx.x; b < nblocks; b += 32) {
        t0 += partials[b * 3 + 0]; t1 += partials[b * 3 + 1]; t2 += partials[b * 3 + 2];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t0 += __shfl_xor_sync(0xffffffffu, t0, o);
        t1 += __shfl_xor_sync(0xffffffffu, t1, o);
        t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
    if (threadIdx.x == 0) *ticket = 0u;
    loss_finish(t0, t1, t2, ctl, comm, host_flag, zero_me, s_x);
}

void launch_loss_stats(const float* gray, const float* event_frame, int HW, int pix0, int npix, EngineCtl* ctl,
                       double* partials, int nblocks, uint32_t* zero_me, SplitComm* comm, int* host_flag, cudaStream_t s) {
    launch_k(loss_stats_kernel, dim3(nblocks), dim3(256), 0, s, gray, event_frame, HW, pix0, npix, ctl, partials, nblocks, zero_me, comm, host_flag);
}

// ---- workload counters (bench / profiling only) ---------------------------------------------------
// out[0], out[1]: Gaussians with radius > 0 per view; out[2], out[3]: sum of n_contrib per view (the
// number of (pixel, instance) pairs a pixel walks before it terminates); out[4]: (view, Gaussian) pairs
// whose blend gradient is non-zero.
__global__ void workload_counters_kernel(int P, const uint32_t* __restrict__ rect_raw, const float4* __restrict__ grad8, int HW,
                                         const uint32_t* __restrict__ n_contrib, unsigned long long* __restrict__ out) {
    unsigned long long c[5] = {0, 0, 0, 0, 0};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 2LL * P; i += stride) {
        if (rect_raw[i] != 0u) {
            c[i >= P ? 1 : 0]++;
            const float4 a = grad8[2 * i], b = grad8[2 * i + 1];
            if (a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f || b.x != 0.f || b.y != 0.f) c[4]++;
        }
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 2LL * HW; i += stride)
        c[2 + (i >= HW ? 1 : 0)] += n_contrib[i];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        unsigned long long v = c[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(out + k, v);
    }
}
void launch_workload_counters(int P, const uint32_t* rect_raw, const float4* grad8, int HW, const uint32_t* n_contrib,
                              unsigned long long* out, cudaStream_t s) {
    workload_counters_kernel<<<296, 256, 0, s>>>(P, rect_raw, grad8, HW, n_contrib, out);
}

}  // namespace gsevt
