/*
 * TEST INFRASTRUCTURE — not part of the product.
 *
 * gsevt_oracle.c: a plain-C, single-threaded CPU restatement of the reference's rasteriser path
 * (ChillTerry/GS-EVT, dgr/ = submodules/diff-gaussian-rasterization/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it; the product
 * (libgsevt.so and the Python host package) never does.
 *
 * Pinning: the reference holds no golden vectors or known-answer tests for this path,
 * so this restatement is pinned against the reference ITSELF: tests/test_parity_reference.py runs
 * the unmodified reference extension (oracle/_ref, built by oracle/build_ref.sh) on the GPU box on the
 * same seeded inputs and compares every stage; tests/golden/ holds vectors generated that way
 * (tests/golden/make_golden.py).  Until those vectors exist the status is "parity unpinned".
 *
 * Rounding: the reference is compiled by nvcc with FMA contraction on.  Where results must be
 * bit-exact (depth bits, pixel centres, radii, tile rects -> sort keys, ranges) this file uses
 * explicit fmaf() in the order of the reference's sm_100a SASS (decoded with tools/sass_ssa.py);
 * compile with -ffp-contract=off so the C compiler adds no contractions of its own.
 *
 * Each function cites the reference lines it restates.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                               0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f,  -0.5900435899266435f};

typedef struct OrcScene {
    int32_t P, D, M, W, H;
    float tanfovx, tanfovy, scale_modifier, delta_time;
    const float* bg;             /* [3] */
    const float* means3D;        /* [P,3] */
    const float* shs;            /* [P,M,3] or NULL */
    const float* colors_precomp; /* [P,3] or NULL */
    const float* opacities;      /* [P] */
    const float* scales;         /* [P,3] or NULL */
    const float* rotations;      /* [P,4] or NULL */
    const float* cov3D_precomp;  /* [P,6] or NULL */
    const float* viewmatrix;     /* [16] column-major */
    const float* projmatrix;     /* [16] */
    const float* projmatrix_raw; /* [16] */
    const float* campos;         /* [3] */
    const float* vel;            /* [16] vel_transofrm */
    const float* vel_inv;        /* [16] */
} OrcScene;

/* a*x + b*y + c*z + d as the reference's SASS evaluates it (auxiliary.h:58-77) */
static float affine3(float m0, float m1, float m2, float m3, float x, float y, float z) {
    return fmaf(z, m2, fmaf(x, m0, y * m1)) + m3;
}
static float dot3r(float a0, float a1, float a2, float b0, float b1, float b2) {
    return fmaf(a2, b2, fmaf(a0, b0, a1 * b1));
}

/* computeCov3D, forward.cu:120-154 */
static void cov3d(const float* s, float mod, const float* q, float* cov) {
    const float sx = s[0] * mod, sy = s[1] * mod, sz = s[2] * mod;
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    const float xz = x * z, rx = r * x, rz = r * z, yy = y * y, zz = z * z;
    const float h2 = fmaf(r, y, xz), h6 = fmaf(-r, y, xz), h5 = fmaf(y, z, -rx), h7 = fmaf(y, z, rx);
    const float h1 = fmaf(x, y, -rz), h3 = fmaf(x, y, rz);
    const float qxy = fmaf(x, x, yy), qyz = yy + zz, qxz = fmaf(x, x, zz);
    const float A0 = -(qyz + qyz) + 1.0f, A1 = h1 + h1, A2 = h2 + h2;
    const float A3 = h3 + h3, A4 = -(qxz + qxz) + 1.0f, A5 = h5 + h5;
    const float A6 = h6 + h6, A7 = h7 + h7, A8 = -(qxy + qxy) + 1.0f;
    const float M00 = sx * A0, M01 = sy * A1, M02 = sz * A2;
    const float M10 = sx * A3, M11 = sy * A4, M12 = sz * A5;
    const float M20 = sx * A6, M21 = sy * A7, M22 = sz * A8;
    cov[0] = dot3r(M00, M01, M02, M00, M01, M02);
    cov[1] = dot3r(M00, M01, M02, M10, M11, M12);
    cov[2] = dot3r(M00, M01, M02, M20, M21, M22);
    cov[3] = dot3r(M10, M11, M12, M10, M11, M12);
    cov[4] = dot3r(M10, M11, M12, M20, M21, M22);
    cov[5] = dot3r(M20, M21, M22, M20, M21, M22);
}

typedef struct Ewa {
    float tx, ty, tz, txtz, tytz;
    float J00, J02, J11, J12;
    float T00, T01, T02, T10, T11, T12;
    float a, b, c;
} Ewa;

/* computeCov2D, forward.cu:76-115 (shared with backward.cu:179-214) */
static void ewa(const float* v, const float* p, float fx, float fy, float tanx, float tany, const float* c3, Ewa* e) {
    const float tz = affine3(v[2], v[6], v[10], v[14], p[0], p[1], p[2]);
    const float tx0 = affine3(v[0], v[4], v[8], v[12], p[0], p[1], p[2]);
    const float ty0 = affine3(v[1], v[5], v[9], v[13], p[0], p[1], p[2]);
    const float limx = tanx * 1.3f, limy = tany * 1.3f;
    e->txtz = tx0 / tz;
    e->tytz = ty0 / tz;
    const float cx = fminf(fmaxf(e->txtz, -limx), limx), cy = fminf(fmaxf(e->tytz, -limy), limy);
    e->tx = cx * tz;
    e->ty = cy * tz;
    e->tz = tz;
    const float tz2 = tz * tz;
    e->J00 = fx / tz;
    e->J02 = (-e->tx * fx) / tz2;
    e->J11 = fy / tz;
    e->J12 = (-e->ty * fy) / tz2;
    e->T00 = fmaf(v[2], e->J02, v[0] * e->J00);
    e->T01 = fmaf(v[6], e->J02, v[4] * e->J00);
    e->T02 = fmaf(v[10], e->J02, v[8] * e->J00);
    e->T10 = fmaf(v[2], e->J12, v[1] * e->J11);
    e->T11 = fmaf(v[6], e->J12, v[5] * e->J11);
    e->T12 = fmaf(v[10], e->J12, v[9] * e->J11);
    const float X00 = dot3r(e->T00, e->T01, e->T02, c3[0], c3[1], c3[2]);
    const float X10 = dot3r(e->T00, e->T01, e->T02, c3[1], c3[3], c3[4]);
    const float X20 = dot3r(e->T00, e->T01, e->T02, c3[2], c3[4], c3[5]);
    const float X01 = dot3r(e->T10, e->T11, e->T12, c3[0], c3[1], c3[2]);
    const float X11 = dot3r(e->T10, e->T11, e->T12, c3[1], c3[3], c3[4]);
    const float X21 = dot3r(e->T10, e->T11, e->T12, c3[2], c3[4], c3[5]);
    e->a = dot3r(e->T00, e->T01, e->T02, X00, X10, X20) + 0.3f;
    e->b = dot3r(e->T00, e->T01, e->T02, X01, X11, X21);
    e->c = dot3r(e->T10, e->T11, e->T12, X01, X11, X21) + 0.3f;
}

/* ndc2Pix, auxiliary.h:41-44 (double) */
static float ndc2pix(float v, int S) { return (float)(fma((double)v + 1.0, (double)S, -1.0) * 0.5); }

static int f2i_trunc(float x) {  /* CUDA (int) cast semantics for the values that occur */
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int)x;
}
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* getRect, auxiliary.h:46-56 */
static void get_rect(float px, float py, int radius, int gx, int gy, int* x0, int* y0, int* x1, int* y1) {
    const float r = (float)radius;
    *x0 = imin(gx, imax(0, f2i_trunc((px - r) * 0.0625f)));
    *y0 = imin(gy, imax(0, f2i_trunc((py - r) * 0.0625f)));
    *x1 = imin(gx, imax(0, f2i_trunc((((px + r) + 16.0f) - 1.0f) * 0.0625f)));
    *y1 = imin(gy, imax(0, f2i_trunc((((py + r) + 16.0f) - 1.0f) * 0.0625f)));
}

/* computeColorFromSH, forward.cu:22-73 */
static void sh_color(int deg, int M, const float* sh, const float* pos, const float* campos, float* rgb, uint8_t* clamped) {
    float d[3] = {pos[0] - campos[0], pos[1] - campos[1], pos[2] - campos[2]};
    const float len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const float x = d[0] / len, y = d[1] / len, z = d[2] / len;
    (void)M;
    for (int ch = 0; ch < 3; ch++) {
#define S(k) sh[(k) * 3 + ch]
        float r = SH_C0 * S(0);
        if (deg > 0) {
            r = r - SH_C1 * y * S(1) + SH_C1 * z * S(2) - SH_C1 * x * S(3);
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + SH_C2[0] * xy * S(4) + SH_C2[1] * yz * S(5) + SH_C2[2] * (2.0f * zz - xx - yy) * S(6) +
                    SH_C2[3] * xz * S(7) + SH_C2[4] * (xx - yy) * S(8);
                if (deg > 2) {
                    r = r + SH_C3[0] * y * (3.0f * xx - yy) * S(9) + SH_C3[1] * xy * z * S(10) +
                        SH_C3[2] * y * (4.0f * zz - xx - yy) * S(11) +
                        SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * S(12) +
                        SH_C3[4] * x * (4.0f * zz - xx - yy) * S(13) + SH_C3[5] * z * (xx - yy) * S(14) +
                        SH_C3[6] * x * (xx - 3.0f * yy) * S(15);
                }
            }
        }
#undef S
        r += 0.5f;
        clamped[ch] = r < 0;
        rgb[ch] = r > 0.0f ? r : 0.0f;
    }
}

/* preprocessCUDA, forward.cu:157-258.  Outputs are zero-filled for culled Gaussians. */
int orc_preprocess(const OrcScene* s, int32_t* radii, float* means2D, float* depths, float* cov3D_out,
                   float* conic_opacity, float* rgb, uint8_t* clamped, uint32_t* tiles_touched) {
    const int P = s->P;
    const int gx = (s->W + TILE - 1) / TILE, gy = (s->H + TILE - 1) / TILE;
    const float focal_y = s->H / (2.0f * s->tanfovy), focal_x = s->W / (2.0f * s->tanfovx); /* rasterizer_impl.cu:225-226 */
    const float* v = s->viewmatrix;
    const float* pm = s->projmatrix;
    for (int i = 0; i < P; i++) {
        radii[i] = 0;
        tiles_touched[i] = 0;
        const float* p = s->means3D + 3 * (size_t)i;
        const float depth = affine3(v[2], v[6], v[10], v[14], p[0], p[1], p[2]);
        if (!(depth > 0.2f)) continue; /* in_frustum, auxiliary.h:154 */
        const float hx = affine3(pm[0], pm[4], pm[8], pm[12], p[0], p[1], p[2]);
        const float hy = affine3(pm[1], pm[5], pm[9], pm[13], p[0], p[1], p[2]);
        const float hw = affine3(pm[3], pm[7], pm[11], pm[15], p[0], p[1], p[2]);
        const float p_w = 1.0f / (hw + 0.0000001f);
        const float ndc_x = hx * p_w, ndc_y = hy * p_w;
        float cov[6];
        if (s->cov3D_precomp) {
            memcpy(cov, s->cov3D_precomp + 6 * (size_t)i, sizeof(cov));
        } else {
            cov3d(s->scales + 3 * (size_t)i, s->scale_modifier, s->rotations + 4 * (size_t)i, cov);
            memcpy(cov3D_out + 6 * (size_t)i, cov, sizeof(cov));
        }
        Ewa e;
        ewa(v, p, focal_x, focal_y, s->tanfovx, s->tanfovy, cov, &e);
        const float det = fmaf(e.a, e.c, -(e.b * e.b));
        if (det == 0.0f) continue;
        const float det_inv = 1.0f / det;
        const float cA = e.c * det_inv, cB = e.b * -det_inv, cC = e.a * det_inv;
        const float mid = (e.a + e.c) * 0.5f;
        const float sq = sqrtf(fmaxf(fmaf(mid, mid, -det), 0.1f));
        const float lam = fmaxf(mid + sq, mid - sq);
        const float rf = ceilf(sqrtf(lam) * 3.0f);
        const int radius = f2i_trunc(rf);
        const float mx = ndc2pix(ndc_x, s->W), my = ndc2pix(ndc_y, s->H);
        int x0, y0, x1, y1;
        get_rect(mx, my, radius, gx, gy, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;
        if (s->colors_precomp == NULL) {
            sh_color(s->D, s->M, s->shs + (size_t)i * s->M * 3, p, s->campos, rgb + 3 * (size_t)i, clamped + 3 * (size_t)i);
        } else {
            memcpy(rgb + 3 * (size_t)i, s->colors_precomp + 3 * (size_t)i, 12);
        }
        depths[i] = depth;
        radii[i] = radius;
        means2D[2 * (size_t)i] = mx;
        means2D[2 * (size_t)i + 1] = my;
        conic_opacity[4 * (size_t)i] = cA;
        conic_opacity[4 * (size_t)i + 1] = cB;
        conic_opacity[4 * (size_t)i + 2] = cC;
        conic_opacity[4 * (size_t)i + 3] = s->opacities[i];
        tiles_touched[i] = (uint32_t)((y1 - y0) * (x1 - x0));
    }
    return 0;
}

typedef struct KV { uint64_t key; uint32_t val; uint32_t seq; } KV;
static int kv_cmp(const void* a, const void* b) {
    const KV* x = (const KV*)a; const KV* y = (const KV*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}

/* duplicateWithKeys + stable sort + identifyTileRanges, rasterizer_impl.cu:70-138,280-321.
 * keys/list must hold sum(tiles_touched) entries; ranges holds 2*tiles uint32 (zero = untouched). */
int64_t orc_bin(int32_t P, int32_t W, int32_t H, const int32_t* radii, const float* means2D, const float* depths,
                const uint32_t* tiles_touched, uint64_t* keys, uint32_t* list, uint32_t* ranges) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    size_t N = 0;
    for (int i = 0; i < P; i++) N += tiles_touched[i];
    KV* kv = (KV*)malloc((N ? N : 1) * sizeof(KV));
    if (!kv) return -1;
    size_t off = 0;
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        int x0, y0, x1, y1;
        get_rect(means2D[2 * (size_t)i], means2D[2 * (size_t)i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
        uint32_t bits;
        memcpy(&bits, depths + i, 4);
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) {
                kv[off].key = ((uint64_t)(uint32_t)(y * gx + x) << 32) | bits;
                kv[off].val = (uint32_t)i;
                kv[off].seq = (uint32_t)off;
                off++;
            }
    }
    if (off != N) { free(kv); return -2; }
    qsort(kv, N, sizeof(KV), kv_cmp);
    memset(ranges, 0, (size_t)gx * gy * 8);
    for (size_t i = 0; i < N; i++) {
        keys[i] = kv[i].key;
        list[i] = kv[i].val;
        const uint32_t cur = (uint32_t)(kv[i].key >> 32);
        if (i == 0) ranges[2 * cur] = 0;
        else {
            const uint32_t prev = (uint32_t)(kv[i - 1].key >> 32);
            if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * cur] = (uint32_t)i; }
        }
        if (i == N - 1) ranges[2 * cur + 1] = (uint32_t)N;
    }
    free(kv);
    return (int64_t)N;
}

/* renderCUDA forward, forward.cu:263-392 (per pixel; rounding order of the sm_100a build) */
void orc_render_fwd(int32_t W, int32_t H, const uint32_t* ranges, const uint32_t* list, const float* means2D,
                    const float* colors, const float* conic_opacity, const float* depths, const float* bg,
                    float* out_color, float* out_depth, float* out_opacity, float* final_T, uint32_t* n_contrib,
                    int32_t* n_touched) {
    const int gx = (W + TILE - 1) / TILE;
    const size_t HW = (size_t)W * H;
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            const int tile = (py / TILE) * gx + (px / TILE);
            const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
            float T = 1.0f, C[3] = {0, 0, 0}, D = 0.0f;
            uint32_t contributor = 0, last = 0;
            for (uint32_t k = r0; k < r1; k++) {
                contributor++;
                const uint32_t id = list[k];
                const float dx = means2D[2 * (size_t)id] - (float)px, dy = means2D[2 * (size_t)id + 1] - (float)py;
                const float* co = conic_opacity + 4 * (size_t)id;
                const float q = fmaf(dx, dx * co[0], dy * (dy * co[2]));
                const float power = fmaf(q, -0.5f, -(dy * (dx * co[1])));
                if (power > 0.0f) continue;
                const float alpha = fminf(0.99f, co[3] * expf(power));
                if (alpha < 1.0f / 255.0f) continue;
                const float test_T = T * (1.0f - alpha);
                if (test_T < 0.0001f) break; /* done = true */
                for (int ch = 0; ch < 3; ch++) C[ch] = fmaf(T, alpha * colors[3 * (size_t)id + ch], C[ch]);
                D = fmaf(T, alpha * depths[id], D);
                if (n_touched && test_T > 0.5f) n_touched[id]++;
                T = test_T;
                last = contributor;
            }
            const size_t pix = (size_t)py * W + px;
            final_T[pix] = T;
            n_contrib[pix] = last;
            for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix] = fmaf(bg[ch], T, C[ch]);
            out_depth[pix] = D;
            out_opacity[pix] = 1.0f - T;
        }
}

/* renderCUDA backward, backward.cu:679-903.  Per-Gaussian sums are accumulated in double (the reference
 * uses float atomics in a non-deterministic order) and stored as float:
 * dL_dmean2D [P,2], dL_dconic [P,3] = (xx, xy, yy slots .x .y .w), dL_dopacity [P], dL_dcolors [P,3],
 * dL_ddepths [P]. */
void orc_render_bwd(int32_t P, int32_t W, int32_t H, const uint32_t* ranges, const uint32_t* list, const float* means2D,
                    const float* colors, const float* conic_opacity, const float* depths, const float* bg,
                    const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, const float* dL_dpix_depth,
                    float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolors, float* dL_ddepths) {
    const int gx = (W + TILE - 1) / TILE;
    const size_t HW = (size_t)W * H;
    double* acc = (double*)calloc((size_t)(P > 0 ? P : 1) * 10, sizeof(double));
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            const size_t pix = (size_t)py * W + px;
            const int tile = (py / TILE) * gx + (px / TILE);
            const uint32_t r0 = ranges[2 * tile];
            const float T_final = final_T[pix];
            float T = T_final;
            const uint32_t last_contributor = n_contrib[pix];
            float accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, dpix[3];
            float accum_rec_depth = 0, last_depth = 0, last_alpha = 0;
            for (int ch = 0; ch < 3; ch++) dpix[ch] = dL_dpix[ch * HW + pix];
            const float dpix_depth = dL_dpix_depth ? dL_dpix_depth[pix] : 0.0f;
            float bg_dot = 0.0f;
            for (int ch = 0; ch < 3; ch++) bg_dot += bg[ch] * dpix[ch];
            for (int pos = (int)last_contributor - 1; pos >= 0; pos--) { /* contributor < last_contributor */
                const uint32_t id = list[r0 + (uint32_t)pos];
                const float dx = means2D[2 * (size_t)id] - (float)px, dy = means2D[2 * (size_t)id + 1] - (float)py;
                const float* co = conic_opacity + 4 * (size_t)id;
                const float q = fmaf(dx, dx * co[0], dy * (dy * co[2]));
                const float power = fmaf(q, -0.5f, -(dy * (dx * co[1])));
                if (power > 0.0f) continue;
                const float G = expf(power);
                const float alpha = fminf(0.99f, co[3] * G);
                if (alpha < 1.0f / 255.0f) continue;
                T = T / (1.f - alpha);
                const float dchannel_dcolor = alpha * T;
                float dL_dalpha = 0.0f;
                double* a = acc + (size_t)id * 10;
                for (int ch = 0; ch < 3; ch++) {
                    const float c = colors[3 * (size_t)id + ch];
                    accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                    last_color[ch] = c;
                    dL_dalpha += (c - accum_rec[ch]) * dpix[ch];
                    a[6 + ch] += (double)(dchannel_dcolor * dpix[ch]);
                }
                const float depth = depths[id];
                accum_rec_depth = last_alpha * last_depth + (1.f - last_alpha) * accum_rec_depth;
                last_depth = depth;
                dL_dalpha += (depth - accum_rec_depth) * dpix_depth;
                a[9] += (double)(dchannel_dcolor * dpix_depth);
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                const float dL_dG = co[3] * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                const float dG_ddely = -gdy * co[2] - gdx * co[1];
                a[0] += (double)(dL_dG * dG_ddelx * ddelx_dx);
                a[1] += (double)(dL_dG * dG_ddely * ddely_dy);
                a[2] += (double)(-0.5f * gdx * dx * dL_dG);
                a[3] += (double)(-0.5f * gdx * dy * dL_dG);
                a[4] += (double)(-0.5f * gdy * dy * dL_dG);
                a[5] += (double)(G * dL_dalpha);
            }
        }
    for (int i = 0; i < P; i++) {
        const double* a = acc + (size_t)i * 10;
        dL_dmean2D[2 * (size_t)i] = (float)a[0];
        dL_dmean2D[2 * (size_t)i + 1] = (float)a[1];
        dL_dconic[3 * (size_t)i] = (float)a[2];
        dL_dconic[3 * (size_t)i + 1] = (float)a[3];
        dL_dconic[3 * (size_t)i + 2] = (float)a[4];
        dL_dopacity[i] = (float)a[5];
        for (int ch = 0; ch < 3; ch++) dL_dcolors[3 * (size_t)i + ch] = (float)a[6 + ch];
        dL_ddepths[i] = (float)a[9];
    }
    free(acc);
}

/* ---- small SE3 helpers mirroring dgr/cuda_rasterizer/math.h ---- */
static void rot_t_mul(const float* m /*col-major 4x4*/, const float* g, float* out) { /* R^T g */
    out[0] = m[0] * g[0] + m[1] * g[1] + m[2] * g[2];
    out[1] = m[4] * g[0] + m[5] * g[1] + m[6] * g[2];
    out[2] = m[8] * g[0] + m[9] * g[1] + m[10] * g[2];
}
static void cross3(const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

/* computeCov2DCUDA (backward.cu:152-425) + backward preprocessCUDA (:497-655) + computeColorFromSH
 * backward (:21-147, view-direction part).  Writes dL_dtau [P,6], dL_dvel [P,6], dL_dmeans3D [P,3],
 * dL_dcov3D [P,6] (all zero for invisible Gaussians). */
void orc_geom_bwd(const OrcScene* s, const int32_t* radii, const float* cov3D, const uint8_t* clamped,
                  const float* dL_dmean2D, const float* dL_dconic, const float* dL_dcolors, const float* dL_ddepths,
                  float* dL_dtau, float* dL_dvel, float* dL_dmeans3D, float* dL_dcov3D) {
    const int P = s->P;
    const float hy = s->H / (2.0f * s->tanfovy), hx = s->W / (2.0f * s->tanfovx);
    const float* v = s->viewmatrix;
    const float* pj = s->projmatrix;
    const float* vel = s->vel;
    const float* vi = s->vel_inv;
    /* T_CW_prime = T_vel_inv * T_CW (backward.cu:313-314, math.h:344-346) */
    float Rp[9], tp[3];
    for (int col = 0; col < 3; col++)
        for (int r = 0; r < 3; r++)
            Rp[col * 3 + r] = vi[r] * v[4 * col] + vi[4 + r] * v[4 * col + 1] + vi[8 + r] * v[4 * col + 2];
    for (int r = 0; r < 3; r++) tp[r] = vi[12 + r] + (vi[r] * v[12] + vi[4 + r] * v[13] + vi[8 + r] * v[14]);
    const float pa = s->projmatrix_raw[0], pb = s->projmatrix_raw[5], pe = s->projmatrix_raw[11];
    const float dt = s->delta_time;
    memset(dL_dtau, 0, (size_t)P * 24);
    memset(dL_dvel, 0, (size_t)P * 24);
    memset(dL_dmeans3D, 0, (size_t)P * 12);
    memset(dL_dcov3D, 0, (size_t)P * 24);
    for (int i = 0; i < P; i++) {
        if (!(radii[i] > 0)) continue;
        const float* m = s->means3D + 3 * (size_t)i;
        const float* c3 = cov3D + 6 * (size_t)i;
        const float dA = dL_dconic[3 * (size_t)i], dB = dL_dconic[3 * (size_t)i + 1], dC = dL_dconic[3 * (size_t)i + 2];
        Ewa e;
        ewa(v, m, hx, hy, s->tanfovx, s->tanfovy, c3, &e);
        const float limx = 1.3f * s->tanfovx, limy = 1.3f * s->tanfovy;
        const float xg = (e.txtz < -limx || e.txtz > limx) ? 0.f : 1.f, yg = (e.tytz < -limy || e.tytz > limy) ? 0.f : 1.f;
        const float a = e.a, b = e.b, c = e.c;
        const float denom = a * c - b * b;
        const float d2 = 1.0f / ((denom * denom) + 0.0000001f);
        float da = 0, db = 0, dc = 0;
        float* dcov = dL_dcov3D + 6 * (size_t)i;
        if (d2 != 0) {
            da = d2 * (-c * c * dA + 2 * b * c * dB + (denom - a * c) * dC);
            dc = d2 * (-a * a * dC + 2 * a * b * dB + (denom - a * c) * dA);
            db = d2 * 2 * (b * c * dA - (denom + 2 * b * b) * dB + a * b * dC);
            dcov[0] = e.T00 * e.T00 * da + e.T00 * e.T10 * db + e.T10 * e.T10 * dc;
            dcov[3] = e.T01 * e.T01 * da + e.T01 * e.T11 * db + e.T11 * e.T11 * dc;
            dcov[5] = e.T02 * e.T02 * da + e.T02 * e.T12 * db + e.T12 * e.T12 * dc;
            dcov[1] = 2 * e.T00 * e.T01 * da + (e.T00 * e.T11 + e.T01 * e.T10) * db + 2 * e.T10 * e.T11 * dc;
            dcov[2] = 2 * e.T00 * e.T02 * da + (e.T00 * e.T12 + e.T02 * e.T10) * db + 2 * e.T10 * e.T12 * dc;
            dcov[4] = 2 * e.T02 * e.T01 * da + (e.T01 * e.T12 + e.T02 * e.T11) * db + 2 * e.T11 * e.T12 * dc;
        }
        const float U00 = e.T00 * c3[0] + e.T01 * c3[1] + e.T02 * c3[2], U01 = e.T00 * c3[1] + e.T01 * c3[3] + e.T02 * c3[4],
                    U02 = e.T00 * c3[2] + e.T01 * c3[4] + e.T02 * c3[5];
        const float U10 = e.T10 * c3[0] + e.T11 * c3[1] + e.T12 * c3[2], U11 = e.T10 * c3[1] + e.T11 * c3[3] + e.T12 * c3[4],
                    U12 = e.T10 * c3[2] + e.T11 * c3[4] + e.T12 * c3[5];
        const float dT00 = 2 * U00 * da + U10 * db, dT01 = 2 * U01 * da + U11 * db, dT02 = 2 * U02 * da + U12 * db;
        const float dT10 = 2 * U10 * dc + U00 * db, dT11 = 2 * U11 * dc + U01 * db, dT12 = 2 * U12 * dc + U02 * db;
        const float dJ00 = v[0] * dT00 + v[4] * dT01 + v[8] * dT02, dJ02 = v[2] * dT00 + v[6] * dT01 + v[10] * dT02;
        const float dJ11 = v[1] * dT10 + v[5] * dT11 + v[9] * dT12, dJ12 = v[2] * dT10 + v[6] * dT11 + v[10] * dT12;
        const float tz = 1.f / e.tz, tz2 = tz * tz, tz3 = tz2 * tz;
        float gt[3];
        gt[0] = xg * -hx * tz2 * dJ02;
        gt[1] = yg * -hy * tz2 * dJ12;
        gt[2] = -hx * tz2 * dJ00 - hy * tz2 * dJ11 + (2 * hx * e.tx) * tz3 * dJ02 + (2 * hy * e.ty) * tz3 * dJ12;
        float tau[6], vl[6];
        const float tprime[3] = {Rp[0] * m[0] + Rp[3] * m[1] + Rp[6] * m[2] + tp[0], Rp[1] * m[0] + Rp[4] * m[1] + Rp[7] * m[2] + tp[1],
                                 Rp[2] * m[0] + Rp[5] * m[1] + Rp[8] * m[2] + tp[2]};
        const float tcl[3] = {e.tx, e.ty, e.tz};
        float h[3], ch[3], cg[3];
        rot_t_mul(vel, gt, h);
        cross3(tprime, h, ch);
        cross3(tcl, gt, cg);
        for (int k = 0; k < 3; k++) { tau[k] = h[k]; tau[3 + k] = ch[k]; vl[k] = dt * gt[k]; vl[3 + k] = dt * cg[k]; }
        /* rotation block (backward.cu:360-423) */
        const float dWc[3][3] = {{e.J00 * dT00, e.J11 * dT10, e.J02 * dT00 + e.J12 * dT10},
                                 {e.J00 * dT01, e.J11 * dT11, e.J02 * dT01 + e.J12 * dT11},
                                 {e.J00 * dT02, e.J11 * dT12, e.J02 * dT02 + e.J12 * dT12}};
        for (int k = 0; k < 3; k++) {
            const float ci[3] = {v[4 * k], v[4 * k + 1], v[4 * k + 2]};
            const float cpi[3] = {Rp[3 * k], Rp[3 * k + 1], Rp[3 * k + 2]};
            float hw_[3], a1[3], a2[3];
            rot_t_mul(vel, dWc[k], hw_);
            cross3(cpi, hw_, a1);
            cross3(ci, dWc[k], a2);
            for (int j = 0; j < 3; j++) { tau[3 + j] += a1[j]; vl[3 + j] += dt * a2[j]; }
        }
        /* projection chain (backward.cu:531-626) */
        const float g2x = dL_dmean2D[2 * (size_t)i], g2y = dL_dmean2D[2 * (size_t)i + 1];
        const float mhx = pj[0] * m[0] + pj[4] * m[1] + pj[8] * m[2] + pj[12];
        const float mhy = pj[1] * m[0] + pj[5] * m[1] + pj[9] * m[2] + pj[13];
        const float mhw = pj[3] * m[0] + pj[7] * m[1] + pj[11] * m[2] + pj[15];
        const float m_w = 1.0f / (mhw + 0.0000001f);
        const float al = m_w, be = -mhx * m_w * m_w, ga = -mhy * m_w * m_w;
        const float q[3] = {g2x * al * pa, g2y * al * pb, (g2x * be + g2y * ga) * pe};
        const float pC[3] = {v[0] * m[0] + v[4] * m[1] + v[8] * m[2] + v[12], v[1] * m[0] + v[5] * m[1] + v[9] * m[2] + v[13],
                             v[2] * m[0] + v[6] * m[1] + v[10] * m[2] + v[14]};
        float hq[3], c1[3], c2[3];
        rot_t_mul(vel, q, hq);
        cross3(tprime, hq, c1);
        cross3(pC, q, c2);
        for (int k = 0; k < 3; k++) { tau[k] += hq[k]; tau[3 + k] += c1[k]; vl[k] += dt * q[k]; vl[3 + k] += dt * c2[k]; }
        /* depth row (backward.cu:632-644) */
        const float dz = dL_ddepths[i];
        tau[2] += dz; tau[3] += dz * pC[1]; tau[4] += -dz * pC[0];
        vl[2] += dz; vl[3] += dz * pC[1]; vl[4] += -dz * pC[0];
        /* mean gradient: cov part + projection part + depth part */
        float* gm = dL_dmeans3D + 3 * (size_t)i;
        const float mul1 = mhx * m_w * m_w, mul2 = mhy * m_w * m_w;
        gm[0] = v[0] * gt[0] + v[1] * gt[1] + v[2] * gt[2] + (pj[0] * m_w - pj[3] * mul1) * g2x + (pj[1] * m_w - pj[3] * mul2) * g2y + dz * v[2];
        gm[1] = v[4] * gt[0] + v[5] * gt[1] + v[6] * gt[2] + (pj[4] * m_w - pj[7] * mul1) * g2x + (pj[5] * m_w - pj[7] * mul2) * g2y + dz * v[6];
        gm[2] = v[8] * gt[0] + v[9] * gt[1] + v[10] * gt[2] + (pj[8] * m_w - pj[11] * mul1) * g2x + (pj[9] * m_w - pj[11] * mul2) * g2y + dz * v[10];
        /* SH view-direction chain (backward.cu:100-146) */
        if (s->shs && s->D > 0) {
            const float* sh = s->shs + (size_t)i * s->M * 3;
            float dRGB[3];
            for (int c_ = 0; c_ < 3; c_++) dRGB[c_] = clamped[3 * (size_t)i + c_] ? 0.0f : dL_dcolors[3 * (size_t)i + c_];
            const float o[3] = {m[0] - s->campos[0], m[1] - s->campos[1], m[2] - s->campos[2]};
            const float len = sqrtf(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
            const float x = o[0] / len, y = o[1] / len, z = o[2] / len;
            float dd[3] = {0, 0, 0};
            for (int c_ = 0; c_ < 3; c_++) {
#define S(k) sh[(k) * 3 + c_]
                float gx_ = -SH_C1 * S(3), gy_ = -SH_C1 * S(1), gz_ = SH_C1 * S(2);
                if (s->D > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    gx_ += SH_C2[0] * y * S(4) + SH_C2[2] * 2.f * -x * S(6) + SH_C2[3] * z * S(7) + SH_C2[4] * 2.f * x * S(8);
                    gy_ += SH_C2[0] * x * S(4) + SH_C2[1] * z * S(5) + SH_C2[2] * 2.f * -y * S(6) + SH_C2[4] * 2.f * -y * S(8);
                    gz_ += SH_C2[1] * y * S(5) + SH_C2[2] * 2.f * 2.f * z * S(6) + SH_C2[3] * x * S(7);
                    if (s->D > 2) {
                        gx_ += SH_C3[0] * S(9) * 3.f * 2.f * xy + SH_C3[1] * S(10) * yz + SH_C3[2] * S(11) * -2.f * xy +
                               SH_C3[3] * S(12) * -3.f * 2.f * xz + SH_C3[4] * S(13) * (-3.f * xx + 4.f * zz - yy) +
                               SH_C3[5] * S(14) * 2.f * xz + SH_C3[6] * S(15) * 3.f * (xx - yy);
                        gy_ += SH_C3[0] * S(9) * 3.f * (xx - yy) + SH_C3[1] * S(10) * xz + SH_C3[2] * S(11) * (-3.f * yy + 4.f * zz - xx) +
                               SH_C3[3] * S(12) * -3.f * 2.f * yz + SH_C3[4] * S(13) * -2.f * xy + SH_C3[5] * S(14) * -2.f * yz +
                               SH_C3[6] * S(15) * -3.f * 2.f * xy;
                        gz_ += SH_C3[1] * S(10) * xy + SH_C3[2] * S(11) * 4.f * 2.f * yz + SH_C3[3] * S(12) * 3.f * (2.f * zz - xx - yy) +
                               SH_C3[4] * S(13) * 4.f * 2.f * xz + SH_C3[5] * S(14) * (xx - yy);
                    }
                }
#undef S
                dd[0] += gx_ * dRGB[c_]; dd[1] += gy_ * dRGB[c_]; dd[2] += gz_ * dRGB[c_];
            }
            const float sum2 = o[0] * o[0] + o[1] * o[1] + o[2] * o[2];
            const float inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            const float gsh[3] = {((+sum2 - o[0] * o[0]) * dd[0] - o[1] * o[0] * dd[1] - o[2] * o[0] * dd[2]) * inv32,
                                  (-o[0] * o[1] * dd[0] + (sum2 - o[1] * o[1]) * dd[1] - o[2] * o[1] * dd[2]) * inv32,
                                  (-o[0] * o[2] * dd[0] - o[1] * o[2] * dd[1] + (sum2 - o[2] * o[2]) * dd[2]) * inv32};
            for (int k = 0; k < 3; k++) { gm[k] += gsh[k]; tau[k] -= gsh[k]; vl[k] -= gsh[k]; }
        }
        memcpy(dL_dtau + 6 * (size_t)i, tau, 24);
        memcpy(dL_dvel + 6 * (size_t)i, vl, 24);
    }
}

/* markVisible / checkFrustum, rasterizer_impl.cu:54-66 */
void orc_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, uint8_t* present) {
    const float* v = viewmatrix;
    for (int i = 0; i < P; i++) {
        const float* p = means3D + 3 * (size_t)i;
        present[i] = affine3(v[2], v[6], v[10], v[14], p[0], p[1], p[2]) > 0.2f;
    }
}
