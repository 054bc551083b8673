// extern "C" entry points of libgsevt.so (declared in include/gsevt.h) and the host-side runtime of
// the tracking engine: buffer carving, launch sequencing, CUDA-graph capture of one iteration.
#include "../../include/gsevt.h"
#include "internal.h"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <string>
#include <vector>

namespace gsevt {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ---- operator work-buffer layouts ----------------------------------------------------------------
struct GeomLayout {
    size_t vp, rec, rgb4, cov3D, radii, clamped, tiles_touched, point_offsets, scan_temp, scan_bytes, total;
};
static GeomLayout geom_layout(int P) {
    GeomLayout L;
    size_t o = 0;
    const size_t p = (size_t)(P > 0 ? P : 1);
    L.vp = o; o = align_up(o + sizeof(ViewParams));
    L.rec = o; o = align_up(o + p * 32);
    L.rgb4 = o; o = align_up(o + p * 16);
    L.cov3D = o; o = align_up(o + p * 24);
    L.radii = o; o = align_up(o + p * 4);
    L.clamped = o; o = align_up(o + p);
    L.tiles_touched = o; o = align_up(o + p * 4);
    L.point_offsets = o; o = align_up(o + p * 4);
    L.scan_bytes = scan_temp_bytes(P);
    L.scan_temp = o; o = align_up(o + L.scan_bytes);
    L.total = o + 256;  // slack for base alignment
    return L;
}
struct ImgLayout { size_t accum_alpha, n_contrib, ranges, total; };
static ImgLayout img_layout(int W, int H) {
    ImgLayout L;
    size_t o = 0;
    const size_t hw = (size_t)W * H;
    const size_t tiles = (size_t)((W + 15) / 16) * ((H + 15) / 16);
    L.accum_alpha = o; o = align_up(o + hw * 4);
    L.n_contrib = o; o = align_up(o + hw * 4);
    L.ranges = o; o = align_up(o + tiles * 8);
    L.total = o + 256;
    return L;
}
struct BinLayout { size_t keys_unsorted, keys, list_unsorted, list, sort_temp, sort_bytes, total; };
static BinLayout bin_layout(int R) {
    BinLayout L;
    size_t o = 0;
    const size_t r = (size_t)(R > 0 ? R : 1);
    L.keys_unsorted = o; o = align_up(o + r * 8);
    L.keys = o; o = align_up(o + r * 8);
    L.list_unsorted = o; o = align_up(o + r * 4);
    L.list = o; o = align_up(o + r * 4);
    L.sort_bytes = sort_temp_bytes(R);
    L.sort_temp = o; o = align_up(o + L.sort_bytes);
    L.total = o + 256;
    return L;
}
struct BwdLayout { size_t vp, grad8, gradc, partials, total; };
static BwdLayout bwd_layout(int P) {
    BwdLayout L;
    size_t o = 0;
    const size_t p = (size_t)(P > 0 ? P : 1);
    L.vp = o; o = align_up(o + sizeof(ViewParams));
    L.grad8 = o; o = align_up(o + p * 32);
    L.gradc = o; o = align_up(o + p * 8);
    L.partials = o; o = align_up(o + (size_t)geom_bwd_blocks(P, 1) * GSEVT_NPART * 4);
    L.total = o + 256;
    return L;
}
static inline char* base_aligned(void* p) { return (char*)align_up((size_t)p); }

static int debug_check(int debug, cudaStream_t s, const char* what) {
    if (!debug) {
        cudaError_t e = cudaPeekAtLastError();
        if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return GSEVT_ECUDA; }
        return 0;
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("[debug] %s: %s", what, cudaGetErrorString(e)); return GSEVT_ECUDA; }
    return 0;
}
#define DBG(what) do { int rc__ = debug_check(a->debug, s, what); if (rc__) return rc__; } while (0)

static int check_common(const GsevtRasterArgs* a) {
    if (!a) { set_error("null args"); return GSEVT_EINVAL; }
    if (a->P < 0 || a->width <= 0 || a->height <= 0) { set_error("bad sizes P=%d W=%d H=%d", a->P, a->width, a->height); return GSEVT_EINVAL; }
    if (!a->means3D && a->P > 0) { set_error("means3D is null"); return GSEVT_EINVAL; }
    if ((a->shs == nullptr) == (a->colors_precomp == nullptr)) {
        set_error("Please provide excatly one of either SHs or precomputed colors!");
        return GSEVT_EINVAL;
    }
    const bool sr = a->scales != nullptr && a->rotations != nullptr;
    if (sr == (a->cov3D_precomp != nullptr) || ((a->scales != nullptr) != (a->rotations != nullptr))) {
        set_error("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
        return GSEVT_EINVAL;
    }
    if (a->shs && (a->sh_degree < 0 || a->sh_degree > 3 || (a->sh_degree + 1) * (a->sh_degree + 1) > a->sh_coeffs)) {
        set_error("sh_degree %d needs %d coefficients, shs has %d", a->sh_degree, (a->sh_degree + 1) * (a->sh_degree + 1), a->sh_coeffs);
        return GSEVT_EINVAL;
    }
    return 0;
}

}  // namespace gsevt

using namespace gsevt;

extern "C" {

GSEVT_API const char* gsevt_last_error(void) { return g_err; }
GSEVT_API int gsevt_abi_version(void) { return GSEVT_ABI_VERSION; }
GSEVT_API int gsevt_device_arch(void) {
    int dev = 0;
    GSEVT_CUDA_OK(cudaGetDevice(&dev));
    int major = 0, minor = 0;
    GSEVT_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    GSEVT_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    return major * 10 + minor;
}

GSEVT_API int gsevt_raster_sizes(int32_t P, int32_t width, int32_t height, size_t* geom_bytes, size_t* img_bytes) {
    if (P < 0 || width <= 0 || height <= 0) { set_error("bad sizes"); return GSEVT_EINVAL; }
    if (geom_bytes) *geom_bytes = geom_layout(P).total;
    if (img_bytes) *img_bytes = img_layout(width, height).total;
    return 0;
}
GSEVT_API size_t gsevt_raster_binning_size(int32_t num_rendered) { return bin_layout(num_rendered).total; }
GSEVT_API size_t gsevt_raster_backward_workspace_size(int32_t P) { return bwd_layout(P).total; }

GSEVT_API int64_t gsevt_raster_geom_offset(const char* name, int32_t P) {
    const GeomLayout L = geom_layout(P);
    const std::string n(name ? name : "");
    if (n == "view_params") return (int64_t)L.vp;
    if (n == "rec") return (int64_t)L.rec;
    if (n == "rgb4") return (int64_t)L.rgb4;
    if (n == "cov3D") return (int64_t)L.cov3D;
    if (n == "radii") return (int64_t)L.radii;
    if (n == "clamped") return (int64_t)L.clamped;
    if (n == "tiles_touched") return (int64_t)L.tiles_touched;
    if (n == "point_offsets") return (int64_t)L.point_offsets;
    return -1;
}
GSEVT_API int64_t gsevt_raster_binning_offset(const char* name, int32_t R) {
    const BinLayout L = bin_layout(R);
    const std::string n(name ? name : "");
    if (n == "point_list_keys_unsorted") return (int64_t)L.keys_unsorted;
    if (n == "point_list_keys") return (int64_t)L.keys;
    if (n == "point_list_unsorted") return (int64_t)L.list_unsorted;
    if (n == "point_list") return (int64_t)L.list;
    return -1;
}
GSEVT_API int64_t gsevt_raster_img_offset(const char* name, int32_t W, int32_t H) {
    const ImgLayout L = img_layout(W, H);
    const std::string n(name ? name : "");
    if (n == "accum_alpha") return (int64_t)L.accum_alpha;
    if (n == "n_contrib") return (int64_t)L.n_contrib;
    if (n == "ranges") return (int64_t)L.ranges;
    return -1;
}

GSEVT_API int gsevt_raster_forward_geometry(const GsevtRasterArgs* a, void* stream) {
    int rc = check_common(a);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (a->P == 0) return 0;
    const GeomLayout L = geom_layout(a->P);
    if (!a->geom_buffer || a->geom_bytes < L.total) { set_error("geom_buffer too small (%zu < %zu)", a->geom_bytes, L.total); return GSEVT_ENOMEM; }
    if (!a->viewmatrix || !a->projmatrix || !a->campos || !a->opacities) { set_error("null camera / opacity input"); return GSEVT_EINVAL; }
    char* g = base_aligned(a->geom_buffer);
    ViewParams* vp = (ViewParams*)(g + L.vp);
    launch_build_view_params(vp, a->viewmatrix, a->projmatrix, a->projmatrix_raw, a->campos, a->vel_transform,
                             a->vel_transform_inv, a->background, a->tanfovx, a->tanfovy, a->width, a->height,
                             a->delta_time, s);
    PreAosArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.P = a->P; pa.D = a->sh_degree; pa.M = a->sh_coeffs; pa.vp = vp;
    pa.means3D = a->means3D; pa.scales = a->scales; pa.rotations = a->rotations; pa.opacities = a->opacities;
    pa.shs = a->shs; pa.colors_precomp = a->colors_precomp; pa.cov3D_precomp = a->cov3D_precomp;
    pa.scale_modifier = a->scale_modifier;
    pa.radii_internal = (int*)(g + L.radii); pa.radii_out = a->radii;
    pa.tiles_touched = (uint32_t*)(g + L.tiles_touched);
    pa.cov3D = (float*)(g + L.cov3D); pa.clamped = (uint8_t*)(g + L.clamped);
    pa.rec = (float4*)(g + L.rec); pa.rgb4 = (float4*)(g + L.rgb4);
    launch_preprocess_aos(pa, s);
    DBG("preprocess");
    launch_scan(g + L.scan_temp, L.scan_bytes, pa.tiles_touched, (uint32_t*)(g + L.point_offsets), a->P, s);
    DBG("scan");
    uint32_t n = 0;
    GSEVT_CUDA_OK(cudaMemcpyAsync(&n, (uint32_t*)(g + L.point_offsets) + (a->P - 1), 4, cudaMemcpyDeviceToHost, s));
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));
    if (n > 0x7fffffffu) { set_error("num_rendered overflow"); return GSEVT_EOVERFLOW; }
    return (int)n;
}

GSEVT_API int gsevt_raster_forward_render(const GsevtRasterArgs* a, void* stream) {
    int rc = check_common(a);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (!a->out_color || !a->out_depth || !a->out_opacity) { set_error("null output image"); return GSEVT_EINVAL; }
    if (a->P == 0) return 0;
    const GeomLayout L = geom_layout(a->P);
    const ImgLayout I = img_layout(a->width, a->height);
    const BinLayout B = bin_layout(a->num_rendered);
    if (!a->geom_buffer || a->geom_bytes < L.total) { set_error("geom_buffer too small"); return GSEVT_ENOMEM; }
    if (!a->img_buffer || a->img_bytes < I.total) { set_error("img_buffer too small (%zu < %zu)", a->img_bytes, I.total); return GSEVT_ENOMEM; }
    if (!a->binning_buffer || a->binning_bytes < B.total) { set_error("binning_buffer too small (%zu < %zu)", a->binning_bytes, B.total); return GSEVT_ENOMEM; }
    char* g = base_aligned(a->geom_buffer);
    char* im = base_aligned(a->img_buffer);
    char* bn = base_aligned(a->binning_buffer);
    const ViewParams* vp = (const ViewParams*)(g + L.vp);
    const int gx = (a->width + 15) / 16, gy = (a->height + 15) / 16;
    const int R = a->num_rendered;
    uint64_t* keys_u = (uint64_t*)(bn + B.keys_unsorted);
    uint64_t* keys = (uint64_t*)(bn + B.keys);
    uint32_t* list_u = (uint32_t*)(bn + B.list_unsorted);
    uint32_t* list = (uint32_t*)(bn + B.list);
    if (R > 0) {
        launch_emit_keys(a->P, 1, vp, (const float4*)(g + L.rec), (const int*)(g + L.radii),
                         (const uint32_t*)(g + L.point_offsets), keys_u, list_u, 0, nullptr, nullptr, s);
        DBG("emit_keys");
        const int bit = (int)higher_msb((uint32_t)(gx * gy));
        launch_sort_pairs(bn + B.sort_temp, B.sort_bytes, keys_u, keys, list_u, list, R, 32 + bit, s);
        DBG("sort");
    }
    launch_identify_ranges(keys, (uint2*)(im + I.ranges), gx * gy, R, nullptr, 0, s);
    DBG("ranges");
    BlendFwdArgs f;
    memset(&f, 0, sizeof(f));
    f.W = a->width; f.H = a->height; f.grid_x = gx; f.grid_y = gy; f.nviews = 1; f.tile_y0 = 0; f.tile_rows = gy;
    f.ranges = (const uint2*)(im + I.ranges); f.point_list = list;
    f.rec = (const float4*)(g + L.rec); f.rgb4 = (const float4*)(g + L.rgb4);
    f.view_stride_gauss = (size_t)a->P; f.bg = a->background; f.views = vp;
    f.final_T = (float*)(im + I.accum_alpha); f.n_contrib = (uint32_t*)(im + I.n_contrib);
    f.out_color = a->out_color; f.out_depth = a->out_depth; f.out_opacity = a->out_opacity;
    f.n_touched = a->want_n_touched ? a->n_touched : nullptr;
    if (!f.bg) { set_error("background is null"); return GSEVT_EINVAL; }
    launch_blend_fwd_rgb(f, s);
    DBG("blend_fwd");
    return 0;
}

GSEVT_API int gsevt_raster_backward(const GsevtRasterArgs* a, void* stream) {
    int rc = check_common(a);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (!a->pose_grads) { set_error("pose_grads is null"); return GSEVT_EINVAL; }
    if (a->P == 0) { GSEVT_CUDA_OK(cudaMemsetAsync(a->pose_grads, 0, 48, s)); return 0; }
    if (!a->dL_dout_color) { set_error("dL_dout_color is null"); return GSEVT_EINVAL; }
    if (!a->projmatrix_raw || !a->vel_transform || !a->vel_transform_inv) { set_error("backward needs projmatrix_raw and the velocity transforms"); return GSEVT_EINVAL; }
    const GeomLayout L = geom_layout(a->P);
    const ImgLayout I = img_layout(a->width, a->height);
    const BinLayout B = bin_layout(a->num_rendered);
    const BwdLayout W = bwd_layout(a->P);
    if (!a->geom_buffer || a->geom_bytes < L.total || !a->img_buffer || a->img_bytes < I.total ||
        !a->binning_buffer || a->binning_bytes < B.total) { set_error("forward buffers missing or too small"); return GSEVT_ENOMEM; }
    if (!a->bwd_workspace || a->bwd_workspace_bytes < W.total) { set_error("bwd_workspace too small (%zu < %zu)", a->bwd_workspace_bytes, W.total); return GSEVT_ENOMEM; }
    char* g = base_aligned(a->geom_buffer);
    char* im = base_aligned(a->img_buffer);
    char* bn = base_aligned(a->binning_buffer);
    char* w = base_aligned(a->bwd_workspace);
    ViewParams* vp = (ViewParams*)(w + W.vp);
    launch_build_view_params(vp, a->viewmatrix, a->projmatrix, a->projmatrix_raw, a->campos, a->vel_transform,
                             a->vel_transform_inv, a->background, a->tanfovx, a->tanfovy, a->width, a->height,
                             a->delta_time, s);
    GSEVT_CUDA_OK(cudaMemsetAsync(w + W.grad8, 0, (size_t)a->P * 32, s));
    GSEVT_CUDA_OK(cudaMemsetAsync(w + W.gradc, 0, (size_t)a->P * 8, s));
    const int gx = (a->width + 15) / 16, gy = (a->height + 15) / 16;
    BlendBwdArgs b;
    memset(&b, 0, sizeof(b));
    b.W = a->width; b.H = a->height; b.grid_x = gx; b.grid_y = gy; b.nviews = 1; b.tile_y0 = 0; b.tile_rows = gy;
    b.ranges = (const uint2*)(im + I.ranges); b.point_list = (const uint32_t*)(bn + B.list);
    b.rec = (const float4*)(g + L.rec); b.rgb4 = (const float4*)(g + L.rgb4);
    b.view_stride_gauss = (size_t)a->P; b.bg = a->background; b.views = vp;
    b.final_T = (const float*)(im + I.accum_alpha); b.n_contrib = (const uint32_t*)(im + I.n_contrib);
    b.dL_dpix = a->dL_dout_color; b.dL_dpix_depth = a->dL_dout_depth;
    b.grad8 = (float4*)(w + W.grad8); b.gradc = (float2*)(w + W.gradc);
    if (a->num_rendered > 0) launch_blend_bwd_rgb(b, s);
    DBG("blend_bwd");
    GeomBwdArgs q;
    memset(&q, 0, sizeof(q));
    q.P = a->P; q.D = a->sh_degree; q.M = a->sh_coeffs; q.nviews = 1; q.views = vp;
    q.radii = (const int*)(g + L.radii); q.clamped = (const uint8_t*)(g + L.clamped);
    q.grad8 = b.grad8; q.gradc = b.gradc;
    q.means3D = a->means3D; q.shs = a->shs;
    q.cov3D = a->cov3D_precomp ? a->cov3D_precomp : (const float*)(g + L.cov3D);
    q.scales = a->scales; q.rotations = a->rotations; q.scale_modifier = a->scale_modifier;
    q.colors_precomp = a->colors_precomp != nullptr;
    q.partials = (float*)(w + W.partials);
    q.dL_dmeans2D = a->dL_dmeans2D; q.dL_dmeans3D = a->dL_dmeans3D; q.dL_dopacity = a->dL_dopacity;
    q.dL_dcolors = a->dL_dcolors; q.dL_dcov3D = a->dL_dcov3D; q.dL_dsh = a->dL_dsh;
    q.dL_dscales = a->dL_dscales; q.dL_drotations = a->dL_drotations; q.dL_dtau = a->dL_dtau; q.dL_dvel = a->dL_dvel;
    launch_geom_bwd_aos(q, s);
    DBG("geom_bwd");
    launch_reduce_partials(q.partials, geom_bwd_blocks(a->P, 1), a->pose_grads, s);
    DBG("reduce");
    return 0;
}

GSEVT_API int gsevt_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                       uint8_t* present, void* stream) {
    (void)projmatrix;
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) { set_error("bad arguments"); return GSEVT_EINVAL; }
    launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
    GSEVT_CUDA_OK(cudaPeekAtLastError());
    return 0;
}

// ---- events ------------------------------------------------------------------------------------
GSEVT_API int gsevt_event_accumulate(const int16_t* x, const int16_t* y, const uint8_t* p, int32_t n, int32_t width,
                           int32_t height, int32_t* counts, int32_t zero_first, int32_t* oob, void* stream) {
    if (n < 0 || width <= 0 || height <= 0 || !counts || (n > 0 && (!x || !y || !p))) { set_error("bad arguments"); return GSEVT_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    if (zero_first) GSEVT_CUDA_OK(cudaMemsetAsync(counts, 0, (size_t)width * height * 4, s));
    launch_event_accumulate(x, y, p, n, width, height, counts, oob, s);
    GSEVT_CUDA_OK(cudaPeekAtLastError());
    return 0;
}

GSEVT_API int gsevt_event_undistort_map(const double* K, const double* D, int32_t W, int32_t H, int32_t* map_ix, int32_t* map_iy) {
    // cv2.undistort == initUndistortRectifyMap(K, D, I, K) + remap; map evaluated in double and
    // quantised to 1/32 px (INTER_BITS = 5).  SURVEY.md 8(a) a2.
    if (!K || !D || !map_ix || !map_iy || W <= 0 || H <= 0) { set_error("bad arguments"); return GSEVT_EINVAL; }
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3], k3 = D[4];
    for (int v = 0; v < H; v++) {
        for (int u = 0; u < W; u++) {
            const double x = (u - cx) / fx, y = (v - cy) / fy;
            const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
            const double kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2;
            const double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2);
            const double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy;
            map_ix[(size_t)v * W + u] = (int32_t)nearbyint((xd * fx + cx) * 32.0);
            map_iy[(size_t)v * W + u] = (int32_t)nearbyint((yd * fy + cy) * 32.0);
        }
    }
    return 0;
}

GSEVT_API size_t gsevt_event_frame_scratch_size(int32_t W, int32_t H) { return align_up((size_t)2 * W * H * 4) + 64 * 8 + 256; }

GSEVT_API int gsevt_event_frame_k(const int32_t* counts, const int32_t* map_ix, const int32_t* map_iy, int32_t W, int32_t H,
                        int32_t levels, int32_t ksize, float* sign_out, float* unsign_out, void* scratch, size_t scratch_bytes, void* stream) {
    if (!counts || !map_ix || !map_iy || !sign_out || !unsign_out || !scratch || W <= 0 || H <= 0 || levels < 1 || levels > GSEVT_MAX_LEVELS) {
        set_error("bad arguments"); return GSEVT_EINVAL;
    }
    // OpenCV's kernels for sigma = 0 are multiples of 1/256 up to 9 taps: every product and partial sum of the blur is then
    // exact in fp32 and the result is cv2's whatever its SIMD width.  From 11 taps on the coefficients are arbitrary floats
    // and cv2's own result depends on how its build splits a row into vector body (FMA) and scalar tail (mul + add), so
    // there is no bit-exact answer to reproduce: rejected rather than approximated.
    if (ksize < 1 || ksize > 9 || (ksize & 1) == 0) { set_error("gaussian kernel size must be 1, 3, 5, 7 or 9 (got %d)", ksize); return GSEVT_EINVAL; }
    if (scratch_bytes < gsevt_event_frame_scratch_size(W, H)) { set_error("scratch too small"); return GSEVT_ENOMEM; }
    char* b = base_aligned(scratch);
    launch_event_frame(counts, map_ix, map_iy, W, H, levels, ksize, sign_out, unsign_out, (float*)b,
                       (double*)(b + align_up((size_t)2 * W * H * 4)), (cudaStream_t)stream);
    GSEVT_CUDA_OK(cudaPeekAtLastError());
    return 0;
}

GSEVT_API int gsevt_event_frame(const int32_t* counts, const int32_t* map_ix, const int32_t* map_iy, int32_t W, int32_t H,
                      int32_t levels, float* sign_out, float* unsign_out, void* scratch, size_t scratch_bytes, void* stream) {
    // the 9-tap kernel of every configs/VECTOR yaml (gaussian_kernel_size: 9)
    return gsevt_event_frame_k(counts, map_ix, map_iy, W, H, levels, 9, sign_out, unsign_out, scratch, scratch_bytes, stream);
}

}  // extern "C"

// =================================================================================================
// Tracking engine
// =================================================================================================
struct GsevtMap {
    int P = 0, D = 0;
    float4* xyz_opacity = nullptr;
    float4* cov_a = nullptr;
    float2* cov_b = nullptr;
    float* sh_planar = nullptr;
    float* sh_aos = nullptr;
    float* smax2 = nullptr;     // [P] largest eigenvalue of the 3-D covariance
    size_t bytes = 0;
};

struct LevelInfo {
    int W, H, gx, gy;
    float tanfovx, tanfovy, focal_x, focal_y;
    float proj_raw[16];
    size_t ev_offset;  // into the event pyramids
};

struct GsevtEngine {
    const GsevtMap* map = nullptr;
    GsevtEngineConfig cfg;
    int nlevels = 0;
    LevelInfo lv[GSEVT_MAX_LEVELS];
    int cur_level = 0;
    // device memory (library-owned)
    EngineCtl* ctl = nullptr;
    ViewParams* views = nullptr;
    float* bg3 = nullptr;
    float4* rec = nullptr;
    float4* grad8 = nullptr;
    uint32_t* rect_raw = nullptr;        // [padded 2P] tile rect per (view, Gaussian), 0 = not visible
    uint32_t* depth_raw = nullptr;       // [padded 2P] depth bits per (view, Gaussian)
    uint8_t* clamped = nullptr;
    uint32_t* active_list = nullptr;   // [2P] compacted pairs with a gradient
    uint32_t* active_count = nullptr;
    uint32_t* vis_list = nullptr;        // screen-tile split: visible pairs of the iteration (split projection kernel -> scatter)
    uint32_t* vis_count = nullptr;       // vis_count[0] = visible pairs, vis_count[1] = survivors of the strip pre-test
    uint32_t* surv_list = nullptr;
    // binning (bucketbin.cu): bucket grid of the current level / strip, per-bucket segment table, key segments, tile lists
    int bin_mode = 0;                    // gsevt_engine_set_binning: 0 automatic, 1 buckets of one tile, 2 buckets of 2 x 2 tiles
    int bk_shift = 0, bk_nbx = 0, bk_nby = 0, bk_by_origin = 0, bk_nb = 0;
    int bk_max_buckets = 0;              // entries of the three tables below (2 views x tiles of level 0)
    uint32_t* bk_cursor = nullptr;       // [bk_max_buckets][GSEVT_BK_CURSOR_STRIDE]
    uint32_t* bk_start = nullptr; uint32_t* bk_cap = nullptr; uint32_t* bk_counts = nullptr;
    uint64_t *bk_keys = nullptr, *bk_keys2 = nullptr;
    long long bk_total = 0;              // key slots of the current level (sum of the bucket capacities)
    long long bk_alloc = 0, vals_alloc = 0;   // allocated key slots (each buffer) / list slots
    int bk_smem_elems = 8, bk_smem_bins = 2048; size_t bk_smem_bytes = 0;   // shared memory of a sort CTA (bucket_sort_smem)
    uint32_t* bk_order = nullptr;        // [bk_max_buckets] buckets by decreasing key count: the order the sort CTAs take them
    uint32_t* tile_order = nullptr;      // [bk_max_buckets] tiles of the strip in the same order: what the blend CTAs take
    uint32_t* vals = nullptr;            // per-tile lists: Gaussian index per slot
    uint32_t* hit_base = nullptr;        // [2 tiles]
    uint2* ranges = nullptr;
    uint32_t* hitmask = nullptr; size_t hitmask_stride = 0;   // forward -> backward: what each warp blended
    int fuse_loss = 0;                   // GSEVT_FUSE_LOSS=1: loss sums in the forward's epilogue instead of a kernel of their own; measured
                                         // no faster (0.6654 vs 0.6630 ms): the epilogue costs the forward what the launch cost the graph
    double* tile_loss = nullptr;         // [tiles of level 0][3] per-tile loss sums (fused loss)
    uint32_t* tile_arrive = nullptr;     // [tiles of level 0] arrival counters + 1 ticket word
    int pdl = 0;                         // GSEVT_PDL=1: programmatic dependent launches along the iteration's kernel chain (internal.h);
                                         // measured 2.6 % SLOWER (0.6775 vs 0.6605 ms): the early-resident dependents take slots from
                                         // the predecessor's tail, and the graph's kernel-to-kernel gaps were only ~1.5 us to begin with
    int blend_bulk = 0;                  // GSEVT_BLEND_BULK=1: id lists staged with cp.async.bulk + mbarrier (blend.cu); measured 1.3 %
                                         // slower per iteration than per-thread loads on the B200 (profiles/README.md), hence off
    float* gray = nullptr; float* final_T = nullptr; uint32_t* n_contrib = nullptr;
    double* loss_partials = nullptr;
    float* geom_partials = nullptr;
    float* lastRT = nullptr;
    int* overflow = nullptr;
    int* host_flag = nullptr;   // pinned, mapped
    int* host_flag_dev = nullptr;
    const float* ev_sign = nullptr;
    int geom_blocks = 0, loss_nb = 0;
    cudaGraphExec_t graph = nullptr;
    int graph_level = -1, graph_y0 = -1, graph_y1 = -1, graph_shift = -1, graph_smem = -1;
    cudaStream_t graph_stream = nullptr;
    // screen-tile split (one engine per rank; strips of tile rows per pyramid level)
    int split_rank = 0, split_n = 1;
    int strip_y0 = 0, strip_y1 = 0;      // current level
    MailBox* mailbox = nullptr;          // this rank's box (its own 2 MiB allocation: exportable through CUDA IPC)
    SplitComm* comm = nullptr;           // device copy of the peer table, NULL unless split
    uint32_t* row_hist = nullptr;        // [256]
    std::vector<void*> allocs;
};
#define GSEVT_MAILBOX_BYTES ((size_t)2 << 20)

namespace gsevt {

thread_local int g_pdl = 0;

template <typename T>
static int dev_alloc(GsevtEngine* e, T** p, size_t n) {
    void* q = nullptr;
    cudaError_t err = cudaMalloc(&q, n * sizeof(T) > 0 ? n * sizeof(T) : 256);
    if (err != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", n * sizeof(T), cudaGetErrorString(err)); return GSEVT_ECUDA; }
    e->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}

static void dev_free(GsevtEngine* e, void* p) {
    if (!p) return;
    for (size_t i = 0; i < e->allocs.size(); i++)
        if (e->allocs[i] == p) { e->allocs.erase(e->allocs.begin() + i); break; }
    cudaFree(p);
}

static void quiesce(GsevtEngine* e, cudaStream_t s) {
    // only this engine's own work uses the buffers (a device-wide sync would also wait for the kernels of OTHER
    // engines, which in a one-process tile-split group may be waiting for this rank)
    cudaStreamSynchronize(s);
    if (e->graph_stream && e->graph_stream != s) cudaStreamSynchronize(e->graph_stream);
    if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
}

// Grows the key segments, the tile lists and the hit-mask table (never inside a captured graph).
static int ensure_capacity(GsevtEngine* e, long long keys, int shift, cudaStream_t s) {
    const long long lists = keys << (2 * shift);
    if (lists > 0x7fffffffLL) { set_error("instance count %lld exceeds the supported maximum", lists); return GSEVT_EOVERFLOW; }
    int rc = 0;
    if (keys > e->bk_alloc) {
        quiesce(e, s);
        dev_free(e, e->bk_keys); dev_free(e, e->bk_keys2);
        e->bk_keys = e->bk_keys2 = nullptr;
        e->bk_alloc = keys + keys / 4;
        rc |= dev_alloc(e, &e->bk_keys, (size_t)e->bk_alloc);
        rc |= dev_alloc(e, &e->bk_keys2, (size_t)e->bk_alloc);
    }
    if (lists > e->vals_alloc) {
        quiesce(e, s);
        dev_free(e, e->vals); dev_free(e, e->hitmask);
        e->vals = nullptr; e->hitmask = nullptr;
        e->vals_alloc = lists + lists / 4;
        // words per warp row of the hit-mask table: (list slot >> 5) + one spare word per tile, see blend.cu / bucketbin.cu
        e->hitmask_stride = (size_t)(e->vals_alloc / 32 + 4 * (long long)e->bk_max_buckets + 64);
        rc |= dev_alloc(e, &e->vals, (size_t)e->vals_alloc);
        rc |= dev_alloc(e, &e->hitmask, 8 * e->hitmask_stride);
    }
    return rc ? GSEVT_ECUDA : 0;
}

// getProjectionMatrix (graphics_utils.py:49-69) evaluated in double like the reference's python floats,
// stored column-major as float32.
static void projection_colmajor(double znear, double zfar, double fovX, double fovY, float* out) {
    const double tanY = tan(fovY / 2), tanX = tan(fovX / 2);
    const double top = tanY * znear, bottom = -top, right = tanX * znear, left = -right;
    float P[4][4];
    memset(P, 0, sizeof(P));
    P[0][0] = (float)(2.0 * znear / (right - left));
    P[1][1] = (float)(2.0 * znear / (top - bottom));
    P[0][2] = (float)((right + left) / (right - left));
    P[1][2] = (float)((top + bottom) / (top - bottom));
    P[3][2] = 1.0f;
    P[2][2] = (float)(zfar / (zfar - znear));
    P[2][3] = (float)(-(zfar * znear) / (zfar - znear));
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) out[4 * c + r] = P[r][c];
}

// Screen-tile split with a strip that is a proper part of the tile grid: the projection runs the strip pre-test + survivor
// queue kernel and hands the scatter a list of visible pairs (preprocess.cu, bucketbin.cu).
static bool split_kernels(const GsevtEngine* e) {
    return e->split_n > 1 && e->vis_list && e->map->smax2 && (e->strip_y0 > 0 || e->strip_y1 < e->lv[e->cur_level].gy);
}

static PreMapArgs premap_args(GsevtEngine* e) {
    const GsevtMap* m = e->map;
    PreMapArgs pa;
    pa.P = m->P; pa.D = m->D; pa.views = e->views; pa.ctl = e->ctl;
    pa.xyz_opacity = m->xyz_opacity; pa.cov3D_a = m->cov_a; pa.cov3D_b = m->cov_b; pa.sh_planar = m->sh_planar; pa.sh_aos = m->sh_aos; pa.smax2 = m->smax2;
    pa.split_pretest = split_kernels(e) ? 1 : 0;
    pa.vis_list = e->vis_list; pa.vis_count = e->vis_count;
    pa.surv_list = e->surv_list; pa.surv_count = e->vis_count ? e->vis_count + 1 : nullptr;
    pa.rect_raw = e->rect_raw; pa.depth_raw = e->depth_raw; pa.clamped = e->clamped;
    pa.rec = e->rec; pa.grad8 = e->grad8;
    return pa;
}

static BucketArgs bucket_args(GsevtEngine* e) {
    const LevelInfo& L = e->lv[e->cur_level];
    BucketArgs b;
    b.P = e->map->P; b.s = e->bk_shift; b.nbx = e->bk_nbx; b.nby = e->bk_nby; b.by_origin = e->bk_by_origin; b.nb = e->bk_nb;
    b.gx = L.gx; b.gy = L.gy; b.tiles_global = L.gx * L.gy;
    b.rect_raw = e->rect_raw; b.depth_raw = e->depth_raw;
    b.cursor = e->bk_cursor; b.bk_start = e->bk_start; b.bk_cap = e->bk_cap;
    b.keys = e->bk_keys; b.keys2 = e->bk_keys2; b.vals = e->vals; b.ranges = e->ranges; b.hit_base = e->hit_base;
    b.bk_order = e->bk_order; b.smem_elems = e->bk_smem_elems; b.smem_bins = e->bk_smem_bins; b.smem_bytes = e->bk_smem_bytes;
    b.overflow = e->overflow; b.ctl = e->ctl;
    b.sparse = split_kernels(e) ? 1 : 0;
    b.vis_list = e->vis_list; b.vis_count = b.sparse ? e->vis_count : nullptr;
    b.surv_count = b.sparse ? e->vis_count + 1 : nullptr;
    return b;
}

// Enqueue one full optimisation iteration (or one evaluation) on stream s.  When `ev` is non-null an
// event is recorded before every stage and after the last one (GSEVT_NSTAGES + 1 events): used by
// gsevt_engine_profile for per-stage device times, never inside a captured graph.
#define GSEVT_NSTAGES 8
static const char* const kStageNames[GSEVT_NSTAGES] = {
    "preprocess_map", "bucket_scatter", "bucket_sort", "blend_fwd_gray", "loss_stats", "blend_bwd_gray", "geom_bwd_pose", "engine_update"};

static void enqueue_iteration(GsevtEngine* e, cudaStream_t s, cudaEvent_t* ev = nullptr) {
    const GsevtMap* m = e->map;
    const LevelInfo& L = e->lv[e->cur_level];
    const int P = m->P;
    int stage = 0;
    auto mark = [&]() { if (ev) cudaEventRecord(ev[stage], s); stage++; };
    struct PdlScope { PdlScope(int on) { g_pdl = on; } ~PdlScope() { g_pdl = 0; } } pdl_scope(e->pdl);
    // the two ViewParams blocks are current on entry: written by the previous iteration's update kernel, or by
    // the stand-alone pose kernel after any host-side change of state / level (probe_buckets, set_state, ...)
    mark();
    launch_preprocess_map(premap_args(e), s);
    mark();
    const BucketArgs ba = bucket_args(e);
    launch_bucket_scatter(ba, false, s);
    mark();
    launch_bucket_sort(ba, s);
    // (an engine whose strip is empty — more ranks than tile rows — launches neither binning kernel: nobody consumes the
    // split projection's two lists, so clear their counters here)
    if (ba.nb <= 0 && e->vis_count && split_kernels(e)) cudaMemsetAsync(e->vis_count, 0, 8, s);
    mark();
    BlendFwdArgs f;
    memset(&f, 0, sizeof(f));
    f.W = L.W; f.H = L.H; f.grid_x = L.gx; f.grid_y = L.gy; f.nviews = 2;
    f.tile_y0 = e->strip_y0; f.tile_rows = e->strip_y1 - e->strip_y0;
    f.ranges = e->ranges; f.point_list = e->vals; f.rec = e->rec; f.view_stride_gauss = (size_t)P;
    f.views = e->views; f.final_T = e->final_T; f.n_contrib = e->n_contrib; f.out_color = e->gray; f.ctl = e->ctl;
    f.hitmask = e->hitmask; f.hitmask_stride = e->hitmask_stride; f.hit_base = e->hit_base; f.tile_order = e->tile_order;
    f.bulk_ids = e->blend_bulk;
    const float* evf = e->ev_sign + L.ev_offset;
    // the loss is evaluated in the forward's epilogue; an engine with an empty strip launches no forward and still has to
    // take part in the exchange of the sums: it keeps the stand-alone kernel
    const bool fuse_loss = e->fuse_loss && f.tile_rows > 0;
    if (fuse_loss) {
        f.event_frame = evf; f.loss_partials = e->tile_loss; f.tile_arrive = e->tile_arrive; f.loss_ticket = e->tile_arrive + e->bk_max_buckets / 2;
        f.ctl_rw = e->ctl; f.comm = e->comm; f.host_flag = e->host_flag_dev; f.zero_me = e->active_count;
    }
    launch_blend_fwd_gray(f, s);
    mark();
    if (!fuse_loss) {
        // pixel rows of this engine's strip (the whole image unless split): one contiguous slice of the image
        const int py0 = e->strip_y0 * GSEVT_TILE < L.H ? e->strip_y0 * GSEVT_TILE : L.H;
        const int py1 = e->strip_y1 * GSEVT_TILE < L.H ? e->strip_y1 * GSEVT_TILE : L.H;
        launch_loss_stats(e->gray, evf, L.W * L.H, py0 * L.W, (py1 - py0) * L.W, e->ctl, e->loss_partials, e->loss_nb,
                          e->active_count, e->comm, e->host_flag_dev, s);
    }
    mark();
    BlendBwdArgs b;
    memset(&b, 0, sizeof(b));
    b.W = L.W; b.H = L.H; b.grid_x = L.gx; b.grid_y = L.gy; b.nviews = 2;
    b.tile_y0 = e->strip_y0; b.tile_rows = e->strip_y1 - e->strip_y0;
    b.ranges = e->ranges; b.point_list = e->vals; b.rec = e->rec; b.view_stride_gauss = (size_t)P; b.views = e->views;
    b.final_T = e->final_T; b.n_contrib = e->n_contrib; b.gray = e->gray; b.event_frame = evf; b.ctl = e->ctl;
    b.grad8 = e->grad8; b.hitmask = e->hitmask; b.hitmask_stride = e->hitmask_stride; b.hit_base = e->hit_base; b.tile_order = e->tile_order;
    b.bulk_ids = e->blend_bulk;
    launch_blend_bwd_gray(b, s);
    mark();
    GeomBwdArgs q;
    memset(&q, 0, sizeof(q));
    q.P = P; q.D = m->D; q.M = 16; q.nviews = 2; q.views = e->views; q.radii = nullptr; q.clamped = e->clamped;
    q.grad8 = e->grad8; q.active_list = e->active_list; q.active_count = e->active_count; q.xyz_opacity = m->xyz_opacity; q.cov3D_a = m->cov_a; q.cov3D_b = m->cov_b;
    q.sh_planar = m->sh_planar; q.sh_aos = m->sh_aos; q.ctl = e->ctl; q.partials = e->geom_partials;
    launch_geom_compact(2 * P, e->rect_raw, e->grad8, e->active_list, e->active_count, e->ctl, s);
    launch_geom_bwd_map(q, s);
    mark();
    launch_engine_update(e->ctl, e->geom_partials, e->geom_blocks, e->host_flag_dev, e->overflow, e->views, e->bg3, e->comm, s);
    mark();
}

static int upload_level(GsevtEngine* e, int level, cudaStream_t s) {
    const LevelInfo& L = e->lv[level];
    struct { int level, W, H, gx, gy; float tx, ty, fx, fy; float proj[16]; } h;
    h.level = level; h.W = L.W; h.H = L.H; h.gx = L.gx; h.gy = L.gy;
    h.tx = L.tanfovx; h.ty = L.tanfovy; h.fx = L.focal_x; h.fy = L.focal_y;
    memcpy(h.proj, L.proj_raw, sizeof(h.proj));
    static_assert(offsetof(EngineCtl, proj_raw) - offsetof(EngineCtl, level) == 9 * 4, "EngineCtl level block layout");
    GSEVT_CUDA_OK(cudaMemcpyAsync((char*)e->ctl + offsetof(EngineCtl, level), &h, sizeof(h), cudaMemcpyHostToDevice, s));
    return 0;
}

static int upload_strip(GsevtEngine* e, int y0, int y1, cudaStream_t s) {
    e->strip_y0 = y0; e->strip_y1 = y1;
    struct { int y0, y1; } h = {y0, y1};
    static_assert(offsetof(EngineCtl, strip_y1) - offsetof(EngineCtl, strip_y0) == 4, "EngineCtl strip layout");
    GSEVT_CUDA_OK(cudaMemcpyAsync((char*)e->ctl + offsetof(EngineCtl, strip_y0), &h, sizeof(h), cudaMemcpyHostToDevice, s));
    // tiles outside the strip are never binned or blended by this engine: their ranges read as empty
    GSEVT_CUDA_OK(cudaMemsetAsync(e->ranges, 0, 2 * (size_t)e->lv[0].gx * e->lv[0].gy * sizeof(uint2), s));
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));   // h is a stack buffer
    return 0;
}

template <typename T>
static int set_field(GsevtEngine* e, size_t off, const T& v, cudaStream_t s) {
    GSEVT_CUDA_OK(cudaMemcpyAsync((char*)e->ctl + off, &v, sizeof(T), cudaMemcpyHostToDevice, s));
    return 0;
}
#define SETF(field, value) do { auto v__ = (value); int rc__ = set_field(e, offsetof(EngineCtl, field), v__, s); if (rc__) return rc__; } while (0)

}  // namespace gsevt

extern "C" {

GSEVT_API int gsevt_map_create(int32_t P, int32_t sh_degree, const float* xyz, const float* scales, const float* rotations,
                     const float* opacities, const float* shs, float scale_modifier, void* stream, GsevtMap** out) {
    if (P <= 0 || !xyz || !scales || !rotations || !opacities || !shs || !out || sh_degree < 0 || sh_degree > 3) {
        set_error("gsevt_map_create: bad arguments"); return GSEVT_EINVAL;
    }
    GsevtMap* m = new GsevtMap();
    m->P = P; m->D = sh_degree;
    const size_t p = (size_t)P;
    if (cudaMalloc(&m->xyz_opacity, p * 16) != cudaSuccess || cudaMalloc(&m->cov_a, p * 16) != cudaSuccess ||
        cudaMalloc(&m->cov_b, p * 8) != cudaSuccess || cudaMalloc(&m->sh_planar, p * 48 * 4) != cudaSuccess ||
        cudaMalloc(&m->sh_aos, p * 48 * 4) != cudaSuccess || cudaMalloc(&m->smax2, p * 4) != cudaSuccess) {
        set_error("gsevt_map_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
        gsevt_map_destroy(m);
        return GSEVT_ECUDA;
    }
    m->bytes = p * (16 + 16 + 8 + 192 + 192 + 4);
    launch_pack_map(P, 16, xyz, scales, rotations, opacities, shs, scale_modifier, m->xyz_opacity, m->cov_a, m->cov_b,
                    m->sh_planar, m->sh_aos, m->smax2, (cudaStream_t)stream);
    GSEVT_CUDA_OK(cudaPeekAtLastError());
    *out = m;
    return 0;
}
GSEVT_API void gsevt_map_destroy(GsevtMap* m) {
    if (!m) return;
    cudaFree(m->xyz_opacity); cudaFree(m->cov_a); cudaFree(m->cov_b); cudaFree(m->sh_planar); cudaFree(m->sh_aos); cudaFree(m->smax2);
    delete m;
}
GSEVT_API int32_t gsevt_map_size(const GsevtMap* m) { return m ? m->P : 0; }
GSEVT_API size_t gsevt_map_bytes(const GsevtMap* m) { return m ? m->bytes : 0; }

GSEVT_API int gsevt_engine_create(const GsevtMap* map, const GsevtEngineConfig* cfg, GsevtEngine** out) {
    if (!map || !cfg || !out || cfg->width <= 0 || cfg->height <= 0 || cfg->levels < 1 || cfg->levels > GSEVT_MAX_LEVELS) {
        set_error("gsevt_engine_create: bad arguments"); return GSEVT_EINVAL;
    }
    GsevtEngine* e = new GsevtEngine();
    e->map = map; e->cfg = *cfg; e->nlevels = cfg->levels;
    if (const char* v = getenv("GSEVT_BLEND_BULK")) e->blend_bulk = atoi(v) != 0;
    if (const char* v = getenv("GSEVT_PDL")) e->pdl = atoi(v) != 0;
    if (const char* v = getenv("GSEVT_FUSE_LOSS")) e->fuse_loss = atoi(v) != 0;
    size_t ev_off = 0;
    for (int l = 0; l < cfg->levels; l++) {
        LevelInfo& L = e->lv[l];
        // frame.py:64-66,75-82: int(W * 0.5**l); FoV from focal2fov(fx*s, W*s) (graphics_utils.py:100-101)
        const double sc = pow(0.5, l);
        L.W = (int)(cfg->width * sc); L.H = (int)(cfg->height * sc);
        if (L.W <= 0 || L.H <= 0) { set_error("pyramid level %d is empty", l); delete e; return GSEVT_EINVAL; }
        if (L.W > 255 * 16 || L.H > 255 * 16) { set_error("image larger than 4080 px is not supported by the packed tile rects"); delete e; return GSEVT_EINVAL; }
        L.gx = (L.W + 15) / 16; L.gy = (L.H + 15) / 16;
        const double fovx = 2 * atan(L.W / (2 * ((double)cfg->fx * sc)));
        const double fovy = 2 * atan(L.H / (2 * ((double)cfg->fy * sc)));
        L.tanfovx = (float)tan(fovx * 0.5); L.tanfovy = (float)tan(fovy * 0.5);
        L.focal_x = L.W / (2.0f * L.tanfovx); L.focal_y = L.H / (2.0f * L.tanfovy);
        projection_colmajor(cfg->znear, cfg->zfar, fovx, fovy, L.proj_raw);
        L.ev_offset = ev_off;
        ev_off += (size_t)(cfg->width >> l) * (cfg->height >> l);
    }
    const int P = map->P;
    const size_t p2 = 2 * (size_t)P;
    if ((long long)P >= (1LL << 31) / 2) { set_error("map too large: pair ids are 32-bit"); delete e; return GSEVT_EINVAL; }
    e->geom_blocks = geom_bwd_blocks(P, 2);
    const LevelInfo& L0 = e->lv[0];
    const size_t hw = (size_t)L0.W * L0.H;
    e->loss_nb = loss_blocks((int)hw);
    int rc = 0;
    rc |= dev_alloc(e, &e->ctl, 1);
    rc |= dev_alloc(e, &e->views, 2);
    rc |= dev_alloc(e, &e->bg3, 4);
    rc |= dev_alloc(e, &e->rec, 2 * p2);
    rc |= dev_alloc(e, &e->grad8, 2 * p2);
    rc |= dev_alloc(e, &e->rect_raw, preprocess_map_raw_items(P));
    rc |= dev_alloc(e, &e->depth_raw, preprocess_map_raw_items(P));
    rc |= dev_alloc(e, &e->clamped, p2);
    rc |= dev_alloc(e, &e->active_list, p2);
    rc |= dev_alloc(e, &e->active_count, 1);
    e->bk_max_buckets = 2 * L0.gx * L0.gy;
    rc |= dev_alloc(e, &e->bk_cursor, (size_t)e->bk_max_buckets * GSEVT_BK_SUB * GSEVT_BK_CURSOR_STRIDE);
    rc |= dev_alloc(e, &e->bk_start, (size_t)e->bk_max_buckets);
    rc |= dev_alloc(e, &e->bk_cap, (size_t)e->bk_max_buckets);
    rc |= dev_alloc(e, &e->bk_counts, (size_t)e->bk_max_buckets * GSEVT_BK_SUB);
    rc |= dev_alloc(e, &e->bk_order, (size_t)e->bk_max_buckets);
    rc |= dev_alloc(e, &e->tile_order, (size_t)e->bk_max_buckets);
    rc |= dev_alloc(e, &e->ranges, (size_t)e->bk_max_buckets);
    rc |= dev_alloc(e, &e->hit_base, (size_t)e->bk_max_buckets);
    rc |= dev_alloc(e, &e->tile_loss, (size_t)(e->bk_max_buckets / 2) * 3);
    rc |= dev_alloc(e, &e->tile_arrive, (size_t)(e->bk_max_buckets / 2) + 1);
    if (bucket_sort_configure()) { set_error("bucket_sort: cannot reserve %d bytes of shared memory", GSEVT_BK_SMEM_MAX_BYTES); gsevt_engine_destroy(e); return GSEVT_ECUDA; }
    if (P >= (1 << 28)) { set_error("map too large: the tile filter packs a Gaussian index into 28 bits"); gsevt_engine_destroy(e); return GSEVT_EINVAL; }
    bucket_sort_smem(8, &e->bk_smem_elems, &e->bk_smem_bins, &e->bk_smem_bytes);
    {
        // first guess of the list memory (grows on demand at begin_level): 16 instances per Gaussian or the caller's hint
        long long keys = cfg->instance_capacity > 0 ? (long long)cfg->instance_capacity * 2 : (long long)P * 4;
        if (keys < (1 << 18)) keys = 1 << 18;
        if (keys > 0x0fffffff) keys = 0x0fffffff;
        if (!rc) rc |= ensure_capacity(e, keys, 1, nullptr);
    }
    rc |= dev_alloc(e, &e->gray, 2 * hw);
    rc |= dev_alloc(e, &e->final_T, 2 * hw);
    rc |= dev_alloc(e, &e->n_contrib, 2 * hw);
    rc |= dev_alloc(e, &e->loss_partials, (size_t)e->loss_nb * 3 + 2);
    rc |= dev_alloc(e, &e->geom_partials, (size_t)e->geom_blocks * GSEVT_NPART);
    rc |= dev_alloc(e, &e->lastRT, 12);
    rc |= dev_alloc(e, &e->overflow, 1);
    rc |= dev_alloc(e, &e->row_hist, 256);
    if (rc) { gsevt_engine_destroy(e); return GSEVT_ECUDA; }
    e->strip_y0 = 0; e->strip_y1 = e->lv[0].gy;
    if (cudaHostAlloc((void**)&e->host_flag, 4, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&e->host_flag_dev, e->host_flag, 0) != cudaSuccess) {
        set_error("pinned flag allocation failed"); gsevt_engine_destroy(e); return GSEVT_ECUDA;
    }
    *e->host_flag = 0;
    EngineCtl h;
    memset(&h, 0, sizeof(h));
    h.R[0] = h.R[4] = h.R[8] = 1.0f;
    h.lr_base[0] = cfg->lr_rot; h.lr_base[1] = cfg->lr_trans; h.lr_base[2] = cfg->lr_w; h.lr_base[3] = cfg->lr_v;
    h.max_optim_iter = cfg->max_optim_iter; h.converged_threshold = cfg->converged_threshold;
    h.level_done = 1;
    h.strip_y0 = 0; h.strip_y1 = e->lv[0].gy;
    cudaMemcpy(e->ctl, &h, sizeof(h), cudaMemcpyHostToDevice);
    float bg[4] = {cfg->background[0], cfg->background[1], cfg->background[2], 0.f};
    cudaMemcpy(e->bg3, bg, sizeof(bg), cudaMemcpyHostToDevice);
    cudaMemset(e->overflow, 0, 4);
    cudaMemset(e->bk_cursor, 0, (size_t)e->bk_max_buckets * GSEVT_BK_SUB * GSEVT_BK_CURSOR_STRIDE * 4);
    cudaMemset(e->ranges, 0, (size_t)e->bk_max_buckets * sizeof(uint2));
    cudaMemset(e->hit_base, 0, (size_t)e->bk_max_buckets * 4);
    cudaMemset(e->tile_arrive, 0, ((size_t)(e->bk_max_buckets / 2) + 1) * 4);
    cudaMemset(e->tile_loss, 0, (size_t)(e->bk_max_buckets / 2) * 3 * 8);
    cudaMemset(e->loss_partials, 0, ((size_t)e->loss_nb * 3 + 2) * 8);
    cudaMemset(e->grad8, 0, 2 * p2 * 16);
    cudaMemset(e->rect_raw, 0, preprocess_map_raw_items(P) * 4);
    cudaMemset(e->depth_raw, 0, preprocess_map_raw_items(P) * 4);
    cudaDeviceSynchronize();
    if (cudaGetLastError() != cudaSuccess) { set_error("engine init failed"); gsevt_engine_destroy(e); return GSEVT_ECUDA; }
    *out = e;
    return 0;
}

GSEVT_API void gsevt_engine_destroy(GsevtEngine* e) {
    if (!e) return;
    if (e->graph) cudaGraphExecDestroy(e->graph);
    for (void* p : e->allocs) cudaFree(p);
    if (e->host_flag) cudaFreeHost(e->host_flag);
    delete e;
}

GSEVT_API int gsevt_engine_set_state(GsevtEngine* e, const float* R, const float* T, const float* w, const float* v, void* stream) {
    if (!e || !R || !T || !w || !v) { set_error("bad arguments"); return GSEVT_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    float h[18];
    memcpy(h, R, 36); memcpy(h + 9, T, 12); memcpy(h + 12, w, 12); memcpy(h + 15, v, 12);
    GSEVT_CUDA_OK(cudaMemcpyAsync(e->ctl, h, sizeof(h), cudaMemcpyHostToDevice, s));
    launch_pose_setup(e->ctl, e->views, e->bg3, e->cfg.znear, e->cfg.zfar, s);
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));  // h is a stack buffer
    return 0;
}
GSEVT_API int gsevt_engine_get_state(GsevtEngine* e, float* R, float* T, float* w, float* v, void* stream) {
    if (!e) { set_error("bad arguments"); return GSEVT_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    float h[18];
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));   // see gsevt_engine_status
    GSEVT_CUDA_OK(cudaMemcpyAsync(h, e->ctl, sizeof(h), cudaMemcpyDeviceToHost, s));
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));
    if (R) memcpy(R, h, 36);
    if (T) memcpy(T, h + 9, 12);
    if (w) memcpy(w, h + 12, 12);
    if (v) memcpy(v, h + 15, 12);
    return 0;
}

GSEVT_API int gsevt_engine_begin_frame(GsevtEngine* e, double delta_tau, const float* sign_pyr, const float* unsign_pyr, void* stream) {
    (void)unsign_pyr;  // |sign| is recomputed on the fly; kept in the ABI for symmetry with the reference
    if (!e || !sign_pyr || !(delta_tau != 0.0)) { set_error("begin_frame: delta_tau must be non-zero and the event pyramid non-null"); return GSEVT_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    e->ev_sign = sign_pyr;
    struct { float dt, half; } h = {(float)delta_tau, (float)(delta_tau / 2)};
    GSEVT_CUDA_OK(cudaMemcpyAsync((char*)e->ctl + offsetof(EngineCtl, delta_tau), &h, sizeof(h), cudaMemcpyHostToDevice, s));
    // fresh Adam per frame (tracker.py:117-129)
    GSEVT_CUDA_OK(cudaMemsetAsync((char*)e->ctl + offsetof(EngineCtl, adam_m), 0,
                                  offsetof(EngineCtl, lr_base) - offsetof(EngineCtl, adam_m), s));
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

// Bucket grid of the current level / strip for bucket edge 1 << shift.
static void set_bucket_grid(GsevtEngine* e, int shift) {
    const LevelInfo& L = e->lv[e->cur_level];
    e->bk_shift = shift;
    e->bk_nbx = ((L.gx - 1) >> shift) + 1;
    e->bk_by_origin = e->strip_y0 >> shift;
    e->bk_nby = e->strip_y1 > e->strip_y0 ? ((e->strip_y1 - 1) >> shift) - e->bk_by_origin + 1 : 0;
    e->bk_nb = e->bk_nbx * e->bk_nby;
}

// pose_setup + projection + a count-only run of the bucket scatter at the current state: keys per (bucket, sub-segment)
// of the current bucket grid -> counts[2 * bk_nb * GSEVT_BK_SUB] (host).  Leaves the cursors at zero.
static int probe_buckets(GsevtEngine* e, cudaStream_t s, std::vector<uint32_t>& counts) {
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));   // pending iterations first (see gsevt_engine_status)
    SETF(level_done, (int)0);
    launch_pose_setup(e->ctl, e->views, e->bg3, e->cfg.znear, e->cfg.zfar, s);
    launch_preprocess_map(premap_args(e), s);
    const int nc = 2 * e->bk_nb * GSEVT_BK_SUB;
    counts.assign((size_t)(nc > 0 ? nc : 1), 0u);
    if (nc <= 0) {
        if (e->vis_count) GSEVT_CUDA_OK(cudaMemsetAsync(e->vis_count, 0, 8, s));
        return 0;
    }
    launch_bucket_scatter(bucket_args(e), true, s);
    launch_bucket_counts(nc, e->bk_cursor, e->bk_counts, s);
    // (split mode) the probe's projection listed its visible pairs, and no bucket_sort follows to consume the list
    if (e->vis_count) GSEVT_CUDA_OK(cudaMemsetAsync(e->vis_count, 0, 8, s));
    GSEVT_CUDA_OK(cudaMemsetAsync(e->bk_cursor, 0, (size_t)nc * GSEVT_BK_CURSOR_STRIDE * 4, s));
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));   // counts is pageable host memory: wait first (see gsevt_engine_status)
    GSEVT_CUDA_OK(cudaMemcpyAsync(counts.data(), e->bk_counts, (size_t)nc * 4, cudaMemcpyDeviceToHost, s));
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

}  // extern "C"

namespace gsevt {
// Contiguous partition of `rows` tile rows into n strips of roughly equal cost (cost = tile instances of the row,
// + 1 so that empty rows still count).  bounds[k] .. bounds[k+1] is strip k.  Every strip gets at least one row
// while rows >= n.  Pure host code: identical inputs give identical strips on every rank.
void balance_rows(const uint32_t* cost, int rows, int n, int* bounds) {
    std::vector<unsigned long long> prefix((size_t)rows + 1, 0ull);
    for (int y = 0; y < rows; y++) prefix[y + 1] = prefix[y] + cost[y] + 1ull;
    const unsigned long long total = prefix[rows];
    bounds[0] = 0;
    for (int k = 1; k < n; k++) {
        const unsigned long long target = total * (unsigned long long)k / (unsigned long long)n;
        int y = bounds[k - 1];
        while (y < rows && prefix[y] < target) y++;
        if (y > 0 && y <= rows && prefix[y] - target > target - prefix[y - 1]) y--;   // nearer boundary
        int lo = bounds[k - 1] + 1, hi = rows - (n - k);
        if (hi < lo) { lo = hi = (bounds[k - 1] < rows ? bounds[k - 1] : rows); }    // fewer rows than ranks: empty strips at the end
        if (rows < n) { lo = bounds[k - 1] < rows ? bounds[k - 1] + 1 : rows; hi = lo; }
        if (y < lo) y = lo;
        if (y > hi) y = hi;
        bounds[k] = y;
    }
    bounds[n] = rows;
}

// Sizes the binning of the current level at the current pose: bucket shape, the segment of every bucket (its key count
// now + slack: the count moves while the pose is optimised; a bucket that outgrows its segment voids the iteration on the
// device and the host comes back here through gsevt_engine_resume), list memory, shared memory of the sort kernel.
// Split mode: (re)balances the strips first.
static int size_level(GsevtEngine* e, cudaStream_t s, bool rebalance, int slack_div) {
    const LevelInfo& L = e->lv[e->cur_level];
    int rc = 0;
    std::vector<uint32_t> counts;
    if (e->split_n > 1 && rebalance) {
        if ((rc = upload_strip(e, 0, L.gy, s))) return rc;
        set_bucket_grid(e, 0);
        if ((rc = probe_buckets(e, s, counts))) return rc;   // (only the projection is needed here)
        launch_row_histogram(2 * e->map->P, e->rect_raw, e->row_hist, s);
        uint32_t h[256];
        GSEVT_CUDA_OK(cudaStreamSynchronize(s));
        GSEVT_CUDA_OK(cudaMemcpyAsync(h, e->row_hist, sizeof(h), cudaMemcpyDeviceToHost, s));
        GSEVT_CUDA_OK(cudaStreamSynchronize(s));
        int bounds[GSEVT_SPLIT_MAX + 1];
        balance_rows(h, L.gy, e->split_n, bounds);
        if ((rc = upload_strip(e, bounds[e->split_rank], bounds[e->split_rank + 1], s))) return rc;
    } else if (e->split_n <= 1) {
        if ((rc = upload_strip(e, 0, L.gy, s))) return rc;
    }
    // Bucket edge: 2 x 2 tiles halve the keys to scatter and sort (a rect of 2 x 2 tiles meets 2.5 buckets on average
    // instead of 4.7 tiles) as long as a bucket still sorts in shared memory; coarse pyramid levels (few tiles, long lists)
    // and very dense maps fall back to one tile per bucket.
    const int strip_tiles = (e->strip_y1 - e->strip_y0) * L.gx;
    int shift = e->bin_mode == 1 ? 0 : (e->bin_mode == 2 ? 1 : (strip_tiles >= 256 ? 1 : 0));
    uint32_t maxc = 0;
    std::vector<uint32_t> keys_of;   // keys per bucket (all sub-segments)
    for (;;) {
        set_bucket_grid(e, shift);
        if ((rc = probe_buckets(e, s, counts))) return rc;   // with this engine's strip in force
        maxc = 0;
        keys_of.assign((size_t)(2 * e->bk_nb > 0 ? 2 * e->bk_nb : 1), 0u);
        for (int b = 0; b < 2 * e->bk_nb; b++) {
            for (int j = 0; j < GSEVT_BK_SUB; j++) keys_of[b] += counts[(size_t)b * GSEVT_BK_SUB + j];
            maxc = keys_of[b] > maxc ? keys_of[b] : maxc;
        }
        if (shift == 1 && e->bin_mode == 0 && maxc + maxc / 4 + 512 > GSEVT_BK_SMEM_MAX_ELEMS) { shift = 0; continue; }
        break;
    }
    const int nbt = 2 * e->bk_nb;
    std::vector<uint32_t> start((size_t)(nbt > 0 ? nbt : 1), 0u), cap((size_t)(nbt > 0 ? nbt : 1), 0u);
    long long total = 0;
    uint32_t maxcap = 8;
    for (int b = 0; b < nbt; b++) {
        // every sub-segment gets the largest sub-count of its bucket + slack (the sub-segments fill evenly: they are picked
        // by the scatter CTA's index)
        uint32_t mj = 0;
        for (int j = 0; j < GSEVT_BK_SUB; j++) mj = counts[(size_t)b * GSEVT_BK_SUB + j] > mj ? counts[(size_t)b * GSEVT_BK_SUB + j] : mj;
        long long c = (long long)mj + mj / 16 + 24;
        if (slack_div > 0) c += mj / slack_div;
        c = (c + 7) / 8 * 8 * GSEVT_BK_SUB;
        start[b] = (uint32_t)total; cap[b] = (uint32_t)c;
        total += c;
        if ((uint32_t)c > maxcap) maxcap = (uint32_t)c;
    }
    if ((rc = ensure_capacity(e, total, shift, s))) return rc;
    e->bk_total = total;
    bucket_sort_smem((int)maxcap, &e->bk_smem_elems, &e->bk_smem_bins, &e->bk_smem_bytes);
    // sort CTAs take the buckets largest first (longest-processing-time order: the tail of the launch is made of small buckets)
    std::vector<uint32_t> order((size_t)(nbt > 0 ? nbt : 1), 0u);
    for (int b = 0; b < nbt; b++) order[b] = (uint32_t)b;
    std::stable_sort(order.begin(), order.begin() + (nbt > 0 ? nbt : 0), [&](uint32_t x, uint32_t y) { return keys_of[x] > keys_of[y]; });
    if (nbt > 0) {
        GSEVT_CUDA_OK(cudaMemcpyAsync(e->bk_start, start.data(), (size_t)nbt * 4, cudaMemcpyHostToDevice, s));
        GSEVT_CUDA_OK(cudaMemcpyAsync(e->bk_cap, cap.data(), (size_t)nbt * 4, cudaMemcpyHostToDevice, s));
        GSEVT_CUDA_OK(cudaMemcpyAsync(e->bk_order, order.data(), (size_t)nbt * 4, cudaMemcpyHostToDevice, s));
    }
    // the blend CTAs take the strip's tiles in the same order (a tile's list is at most its bucket's keys)
    std::vector<uint32_t> torder;
    {
        const int tiles = L.gx * L.gy, edge = 1 << shift;
        torder.reserve((size_t)2 * strip_tiles + 1);
        for (int i = 0; i < nbt; i++) {
            const int b = (int)order[i], view = b >= e->bk_nb ? 1 : 0, bl = b - view * e->bk_nb;
            const int by = bl / e->bk_nbx, bx = bl - by * e->bk_nbx;
            for (int ky = 0; ky < edge; ky++)
                for (int kx = 0; kx < edge; kx++) {
                    const int tx = (bx << shift) + kx, ty = ((by + e->bk_by_origin) << shift) + ky;
                    if (tx < L.gx && ty >= e->strip_y0 && ty < e->strip_y1) torder.push_back((uint32_t)(view * tiles + ty * L.gx + tx));
                }
        }
        if ((int)torder.size() != 2 * strip_tiles) { set_error("internal: tile order covers %zu of %d tiles", torder.size(), 2 * strip_tiles); return GSEVT_ESTATE; }
        if (!torder.empty()) GSEVT_CUDA_OK(cudaMemcpyAsync(e->tile_order, torder.data(), torder.size() * 4, cudaMemcpyHostToDevice, s));
    }
    GSEVT_CUDA_OK(cudaMemsetAsync(e->overflow, 0, 4, s));
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));   // start / cap are stack-frame vectors
    return 0;
}
}  // namespace gsevt

extern "C" {

GSEVT_API int gsevt_engine_begin_level(GsevtEngine* e, int32_t level, int32_t opt_vel, void* stream) {
    if (!e || level < 0 || level >= e->nlevels) { set_error("begin_level: bad level"); return GSEVT_EINVAL; }
    if (!e->ev_sign) { set_error("begin_level before begin_frame"); return GSEVT_ESTATE; }
    cudaStream_t s = (cudaStream_t)stream;
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));   // see gsevt_engine_status
    e->cur_level = level;
    int rc = upload_level(e, level, s);
    if (rc) return rc;
    struct { int opt_vel, optim_iter, start_vel, level_done, iters; } h = {opt_vel ? 1 : 0, 0, 0, 0, 0};
    GSEVT_CUDA_OK(cudaMemcpyAsync((char*)e->ctl + offsetof(EngineCtl, opt_vel), &h, sizeof(h), cudaMemcpyHostToDevice, s));
    SETF(n_losses, (int)0);
    SETF(eval_only, (int)0);
    rc = size_level(e, s, true, 0);
    if (rc) return rc;
    *e->host_flag = 0;
    return 0;
}

GSEVT_API int gsevt_engine_resume(GsevtEngine* e, void* stream) {
    // After the device paused a level because the instance list outgrew the sorted slots (poll_done() == 2):
    // re-count at the current pose, grow, clear the pause.  The voided iteration left no trace in the state.
    if (!e) { set_error("bad arguments"); return GSEVT_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));
    int rc = size_level(e, s, false, 8);   // clears level_done; the strips stay as they are
    if (rc) return rc;
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));
    *e->host_flag = 0;
    return 0;
}

GSEVT_API int gsevt_engine_iterate(GsevtEngine* e, int32_t n, void* stream) {
    if (!e || n < 0) { set_error("bad arguments"); return GSEVT_EINVAL; }
    if (!e->ev_sign) { set_error("iterate before begin_frame"); return GSEVT_ESTATE; }
    cudaStream_t s = (cudaStream_t)stream;
    const bool can_graph = s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread;
    // the captured launches carry the bucket grid, the sort kernel's shared-memory size and every buffer pointer by value
    // (buffers that grow destroy the graph: quiesce())
    if (can_graph && (e->graph == nullptr || e->graph_level != e->cur_level || e->graph_stream != s ||
                      e->graph_y0 != e->strip_y0 || e->graph_y1 != e->strip_y1 ||
                      e->graph_shift != e->bk_shift || e->graph_smem != (int)e->bk_smem_bytes)) {
        if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
        cudaGraph_t g = nullptr;
        GSEVT_CUDA_OK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
        enqueue_iteration(e, s);
        cudaError_t err = cudaStreamEndCapture(s, &g);
        if (err != cudaSuccess || !g) { set_error("graph capture failed: %s", cudaGetErrorString(err)); return GSEVT_ECUDA; }
        err = cudaGraphInstantiate(&e->graph, g, 0);
        cudaGraphDestroy(g);
        if (err != cudaSuccess) { e->graph = nullptr; set_error("graph instantiate failed: %s", cudaGetErrorString(err)); return GSEVT_ECUDA; }
        e->graph_level = e->cur_level; e->graph_stream = s;
        e->graph_y0 = e->strip_y0; e->graph_y1 = e->strip_y1; e->graph_shift = e->bk_shift; e->graph_smem = (int)e->bk_smem_bytes;
    }
    for (int i = 0; i < n; i++) {
        if (can_graph) GSEVT_CUDA_OK(cudaGraphLaunch(e->graph, s));
        else enqueue_iteration(e, s);
    }
    GSEVT_CUDA_OK(cudaPeekAtLastError());
    return 0;
}

GSEVT_API int gsevt_engine_poll_done(GsevtEngine* e) { return e && e->host_flag ? *(volatile int*)e->host_flag : 0; }

GSEVT_API int gsevt_engine_status(GsevtEngine* e, GsevtEngineStatus* out, void* stream) {
    if (!e || !out) { set_error("bad arguments"); return GSEVT_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    static thread_local EngineCtl h;
    uint32_t offs[2] = {0, 0};
    int ov = 0;
    // Wait for the stream BEFORE the device->host copies: a copy into pageable memory waits for the stream inside
    // the driver, and while it does, other host threads of this process cannot submit work — fatal when the work
    // being waited for is a tile-split exchange whose peer is driven by one of those threads.
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));
    GSEVT_CUDA_OK(cudaMemcpyAsync(&h, e->ctl, offsetof(EngineCtl, losses), cudaMemcpyDeviceToHost, s));
    GSEVT_CUDA_OK(cudaMemcpyAsync(&ov, e->overflow, 4, cudaMemcpyDeviceToHost, s));
    {
        // tile instances per view = the lengths of the view's tile lists
        const LevelInfo& L = e->lv[e->cur_level];
        const int tiles = L.gx * L.gy;
        std::vector<uint2> r((size_t)2 * tiles);
        GSEVT_CUDA_OK(cudaMemcpyAsync(r.data(), e->ranges, r.size() * sizeof(uint2), cudaMemcpyDeviceToHost, s));
        GSEVT_CUDA_OK(cudaStreamSynchronize(s));
        for (int t = 0; t < tiles; t++) { offs[0] += r[t].y - r[t].x; offs[1] += r[tiles + t].y - r[tiles + t].x; }
    }
    memset(out, 0, sizeof(*out));
    out->level_done = h.level_done; out->optim_iter = h.optim_iter; out->start_vel_opt_iter = h.start_vel_opt_iter;
    out->opt_vel = h.opt_vel; out->iters_executed = h.iters_executed; out->overflow = ov;
    out->num_rendered[0] = (int)offs[0]; out->num_rendered[1] = (int)offs[1];
    out->last_loss = h.last_loss;
    memcpy(out->pose_grads, h.grads, sizeof(h.grads));
    return 0;
}

GSEVT_API int gsevt_engine_losses(GsevtEngine* e, float* out, int32_t capacity, void* stream) {
    if (!e || !out || capacity < 0) { set_error("bad arguments"); return GSEVT_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    int n = 0;
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));   // see gsevt_engine_status
    GSEVT_CUDA_OK(cudaMemcpyAsync(&n, (char*)e->ctl + offsetof(EngineCtl, n_losses), 4, cudaMemcpyDeviceToHost, s));
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));
    // the device keeps a ring of the last GSEVT_MAX_LOSSES losses: return the most recent min(n, capacity, ring) in order
    int have = n < GSEVT_MAX_LOSSES ? n : GSEVT_MAX_LOSSES;
    int c = have < capacity ? have : capacity;
    if (c > 0) {
        static thread_local float ring[GSEVT_MAX_LOSSES];
        GSEVT_CUDA_OK(cudaMemcpyAsync(ring, (char*)e->ctl + offsetof(EngineCtl, losses), sizeof(ring), cudaMemcpyDeviceToHost, s));
        GSEVT_CUDA_OK(cudaStreamSynchronize(s));
        for (int i = 0; i < c; i++) out[i] = ring[(n - c + i) % GSEVT_MAX_LOSSES];
    }
    return c;
}

GSEVT_API int gsevt_engine_const_vel_model(GsevtEngine* e, double tau, void* stream) {
    if (!e) { set_error("bad arguments"); return GSEVT_EINVAL; }
    launch_const_vel(e->ctl, (float)tau, (cudaStream_t)stream);
    launch_pose_setup(e->ctl, e->views, e->bg3, e->cfg.znear, e->cfg.zfar, (cudaStream_t)stream);
    GSEVT_CUDA_OK(cudaPeekAtLastError());
    return 0;
}
GSEVT_API int gsevt_engine_weighted_velocity(GsevtEngine* e, const float* last_R, const float* last_T, double delta_tau,
                                   double weight, void* stream) {
    if (!e || !last_R || !last_T || delta_tau == 0.0) { set_error("bad arguments"); return GSEVT_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    float h[12];
    memcpy(h, last_R, 36); memcpy(h + 9, last_T, 12);
    GSEVT_CUDA_OK(cudaMemcpyAsync(e->lastRT, h, sizeof(h), cudaMemcpyHostToDevice, s));
    launch_weighted_velocity(e->ctl, e->lastRT, (float)delta_tau, (float)weight, s);
    launch_pose_setup(e->ctl, e->views, e->bg3, e->cfg.znear, e->cfg.zfar, s);
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

GSEVT_API int gsevt_engine_eval(GsevtEngine* e, int32_t level, int32_t signed_loss, float* loss_out, float* grads_out12, void* stream) {
    if (!e || level < 0 || level >= e->nlevels) { set_error("bad arguments"); return GSEVT_EINVAL; }
    if (!e->ev_sign) { set_error("eval before begin_frame"); return GSEVT_ESTATE; }
    cudaStream_t s = (cudaStream_t)stream;
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));   // see gsevt_engine_status
    e->cur_level = level;
    int rc = upload_level(e, level, s);
    if (rc) return rc;
    SETF(eval_only, (int)1);
    SETF(loss_signed, (int)(signed_loss ? 1 : 0));
    rc = size_level(e, s, true, 0);
    if (rc) return rc;
    enqueue_iteration(e, s);
    GsevtEngineStatus st;
    rc = gsevt_engine_status(e, &st, stream);
    if (rc) return rc;
    if (loss_out) *loss_out = st.last_loss;
    if (grads_out12) memcpy(grads_out12, st.pose_grads, 48);
    SETF(eval_only, (int)0);
    SETF(level_done, (int)1);
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));
    if (st.overflow) { set_error("instance capacity exceeded during eval"); return GSEVT_EOVERFLOW; }
    return 0;
}

GSEVT_API int gsevt_engine_render_delta(GsevtEngine* e, int32_t level, float* delta_out, float* gray_last, float* gray_next, void* stream) {
    // Uses the images of the most recent evaluation / iteration at `level`.
    if (!e || level != e->cur_level) { set_error("render_delta: run gsevt_engine_eval at this level first"); return GSEVT_ESTATE; }
    cudaStream_t s = (cudaStream_t)stream;
    const LevelInfo& L = e->lv[level];
    const size_t hw = (size_t)L.W * L.H;
    if (gray_last) GSEVT_CUDA_OK(cudaMemcpyAsync(gray_last, e->gray, hw * 4, cudaMemcpyDeviceToDevice, s));
    if (gray_next) GSEVT_CUDA_OK(cudaMemcpyAsync(gray_next, e->gray + hw, hw * 4, cudaMemcpyDeviceToDevice, s));
    (void)delta_out;
    return 0;
}

GSEVT_API int gsevt_engine_image_state(GsevtEngine* e, int32_t level, float* final_T, uint32_t* n_contrib, void* stream) {
    // Per-pixel blending state of the most recent evaluation / iteration at `level`, both views: [2][H*W] each.
    if (!e || level != e->cur_level) { set_error("image_state: run gsevt_engine_eval at this level first"); return GSEVT_ESTATE; }
    cudaStream_t s = (cudaStream_t)stream;
    const LevelInfo& L = e->lv[level];
    const size_t hw = (size_t)L.W * L.H;
    if (final_T) GSEVT_CUDA_OK(cudaMemcpyAsync(final_T, e->final_T, 2 * hw * 4, cudaMemcpyDeviceToDevice, s));
    if (n_contrib) GSEVT_CUDA_OK(cudaMemcpyAsync(n_contrib, e->n_contrib, 2 * hw * 4, cudaMemcpyDeviceToDevice, s));
    return 0;
}

GSEVT_API int gsevt_engine_view_params(GsevtEngine* e, int32_t view, float* out73, void* stream) {
    // The camera block the kernels of the most recent evaluation / iteration used for `view` (0 last, 1 next), computed on
    // the device by the pose kernel: viewmatrix[16], projmatrix[16] (column-major), campos[3], tanfovx, tanfovy,
    // projmatrix_raw[0], [5], [11], vel_transform[16], vel_transform_inv[16], delta_time.  Host output.
    if (!e || view < 0 || view > 1 || !out73) { set_error("bad arguments"); return GSEVT_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    ViewParams h;
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));
    GSEVT_CUDA_OK(cudaMemcpyAsync(&h, e->views + view, sizeof(h), cudaMemcpyDeviceToHost, s));
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));
    float* o = out73;
    memcpy(o, h.view, 64); o += 16;
    memcpy(o, h.proj, 64); o += 16;
    memcpy(o, h.campos, 12); o += 3;
    *o++ = h.tanfovx; *o++ = h.tanfovy; *o++ = h.proj_a; *o++ = h.proj_b; *o++ = h.proj_e;
    memcpy(o, h.vel, 64); o += 16;
    memcpy(o, h.vel_inv, 64); o += 16;
    *o++ = h.delta_time;
    return 0;
}

GSEVT_API int gsevt_engine_binning(GsevtEngine* e, int32_t view, uint64_t* keys_out, uint32_t* list_out, uint32_t* ranges_out,
                         int32_t capacity, void* stream) {
    if (!e || view < 0 || view > 1 || !keys_out || !list_out || !ranges_out) { set_error("bad arguments"); return GSEVT_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    const LevelInfo& L = e->lv[e->cur_level];
    const int tiles = L.gx * L.gy;
    std::vector<uint2> r((size_t)2 * tiles);
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));   // see gsevt_engine_status
    GSEVT_CUDA_OK(cudaMemcpyAsync(r.data(), e->ranges, r.size() * sizeof(uint2), cudaMemcpyDeviceToHost, s));
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));
    // the engine's tile lists are not packed back to back (bucketbin.cu); the reference's representation is: lists in
    // tile order, ranges into that packed list, (0, 0) for untouched tiles
    std::vector<uint32_t> packed((size_t)tiles, 0u), rr((size_t)2 * tiles, 0u);
    uint32_t count = 0;
    for (int t = 0; t < tiles; t++) {
        const uint2 q = r[(size_t)view * tiles + t];
        packed[t] = count;
        if (q.y > q.x) { rr[2 * t] = count; rr[2 * t + 1] = count + (q.y - q.x); }
        count += q.y - q.x;
    }
    if ((int64_t)count > (int64_t)capacity) { set_error("capacity %d < %u instances", capacity, count); return GSEVT_ENOMEM; }
    uint32_t* packed_dev = nullptr;
    GSEVT_CUDA_OK(cudaMalloc(&packed_dev, (size_t)tiles * 4));
    GSEVT_CUDA_OK(cudaMemcpyAsync(packed_dev, packed.data(), (size_t)tiles * 4, cudaMemcpyHostToDevice, s));
    launch_export_lists(tiles, e->ranges + (size_t)view * tiles, e->vals, e->rec + 2 * (size_t)view * e->map->P, packed_dev, keys_out, list_out, s);
    GSEVT_CUDA_OK(cudaMemcpyAsync(ranges_out, rr.data(), rr.size() * 4, cudaMemcpyHostToDevice, s));
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));
    cudaFree(packed_dev);
    return (int)count;
}

GSEVT_API int gsevt_engine_stage_count(void) { return GSEVT_NSTAGES; }
GSEVT_API const char* gsevt_engine_stage_name(int32_t i) { return i >= 0 && i < GSEVT_NSTAGES ? kStageNames[i] : ""; }

GSEVT_API int gsevt_engine_profile(GsevtEngine* e, int32_t n_iters, float* stage_ms, void* stream) {
    if (!e || n_iters <= 0 || !stage_ms) { set_error("bad arguments"); return GSEVT_EINVAL; }
    if (!e->ev_sign) { set_error("profile before begin_frame"); return GSEVT_ESTATE; }
    cudaStream_t s = (cudaStream_t)stream;
    cudaEvent_t ev[GSEVT_NSTAGES + 1];
    for (int i = 0; i <= GSEVT_NSTAGES; i++) GSEVT_CUDA_OK(cudaEventCreate(&ev[i]));
    double acc[GSEVT_NSTAGES];
    for (int i = 0; i < GSEVT_NSTAGES; i++) acc[i] = 0.0;
    int rc = 0;
    for (int it = 0; it < n_iters && !rc; it++) {
        enqueue_iteration(e, s, ev);
        if (cudaStreamSynchronize(s) != cudaSuccess) { set_error("profile: %s", cudaGetErrorString(cudaGetLastError())); rc = GSEVT_ECUDA; break; }
        for (int i = 0; i < GSEVT_NSTAGES; i++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            acc[i] += ms;
        }
    }
    for (int i = 0; i <= GSEVT_NSTAGES; i++) cudaEventDestroy(ev[i]);
    for (int i = 0; i < GSEVT_NSTAGES; i++) stage_ms[i] = (float)(acc[i] / n_iters);
    return rc;
}

GSEVT_API int gsevt_engine_workload(GsevtEngine* e, int64_t* out8, void* stream) {
    if (!e || !out8) { set_error("bad arguments"); return GSEVT_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    const LevelInfo& L = e->lv[e->cur_level];
    unsigned long long* d = nullptr;
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));   // see gsevt_engine_status
    GSEVT_CUDA_OK(cudaMalloc(&d, 8 * sizeof(unsigned long long)));
    GSEVT_CUDA_OK(cudaMemsetAsync(d, 0, 8 * sizeof(unsigned long long), s));
    launch_workload_counters(e->map->P, e->rect_raw, e->grad8, L.W * L.H, e->n_contrib, d, s);
    unsigned long long h[8];
    uint32_t offs[2] = {0, 0};
    GSEVT_CUDA_OK(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, s));
    uint32_t n_active = 0;   // length of the last iteration's active list (the accumulators themselves are cleared by geom_bwd)
    GSEVT_CUDA_OK(cudaMemcpyAsync(&n_active, e->active_count, 4, cudaMemcpyDeviceToHost, s));
    {
        const int tiles = L.gx * L.gy;
        std::vector<uint2> r((size_t)2 * tiles);
        GSEVT_CUDA_OK(cudaMemcpyAsync(r.data(), e->ranges, r.size() * sizeof(uint2), cudaMemcpyDeviceToHost, s));
        GSEVT_CUDA_OK(cudaStreamSynchronize(s));
        for (int t = 0; t < tiles; t++) { offs[0] += r[t].y - r[t].x; offs[1] += r[tiles + t].y - r[tiles + t].x; }
    }
    cudaFree(d);
    out8[0] = (int64_t)h[0]; out8[1] = (int64_t)h[1];              // visible Gaussians per view
    out8[2] = (int64_t)offs[0]; out8[3] = (int64_t)offs[1];              // tile instances per view
    out8[4] = (int64_t)h[2]; out8[5] = (int64_t)h[3];              // sum of n_contrib per view (pairs walked)
    out8[6] = (int64_t)n_active;                                   // (view, Gaussian) pairs with a non-zero blend gradient
    out8[7] = (int64_t)e->bk_total;                                // key slots of the bucket segments (bucket keys + slack)
    return 0;
}

// ---- screen-tile split ----------------------------------------------------------------------------
GSEVT_API size_t gsevt_split_mailbox_bytes(void) { return GSEVT_MAILBOX_BYTES; }

GSEVT_API int gsevt_split_balance_rows(const uint32_t* row_cost, int32_t rows, int32_t n, int32_t* bounds) {
    if (!row_cost || !bounds || rows < 0 || rows > 255 || n < 1 || n > GSEVT_SPLIT_MAX) { set_error("split_balance_rows: bad arguments"); return GSEVT_EINVAL; }
    int b[GSEVT_SPLIT_MAX + 1];
    balance_rows(row_cost, rows, n, b);
    for (int k = 0; k <= n; k++) bounds[k] = b[k];
    return 0;
}

GSEVT_API int gsevt_engine_split_mailbox(GsevtEngine* e, void** box) {
    if (!e || !box) { set_error("bad arguments"); return GSEVT_EINVAL; }
    if (!e->mailbox) {
        static_assert(sizeof(MailBox) <= GSEVT_MAILBOX_BYTES, "mailbox allocation too small");
        static_assert(sizeof(MailSlot) == 128, "MailSlot layout");
        void* q = nullptr;
        GSEVT_CUDA_OK(cudaMalloc(&q, GSEVT_MAILBOX_BYTES));
        e->allocs.push_back(q);
        e->mailbox = (MailBox*)q;
        GSEVT_CUDA_OK(cudaMemset(q, 0, GSEVT_MAILBOX_BYTES));
        GSEVT_CUDA_OK(cudaDeviceSynchronize());
    }
    *box = e->mailbox;
    return 0;
}

GSEVT_API int gsevt_ipc_export(const void* device_ptr, uint8_t handle64[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!device_ptr || !handle64) { set_error("bad arguments"); return GSEVT_EINVAL; }
    cudaIpcMemHandle_t h;
    GSEVT_CUDA_OK(cudaIpcGetMemHandle(&h, const_cast<void*>(device_ptr)));
    memcpy(handle64, &h, 64);
    return 0;
}
GSEVT_API int gsevt_ipc_open(const uint8_t handle64[64], void** device_ptr) {
    if (!device_ptr || !handle64) { set_error("bad arguments"); return GSEVT_EINVAL; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    GSEVT_CUDA_OK(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
GSEVT_API int gsevt_ipc_close(void* device_ptr) {
    if (!device_ptr) return 0;
    GSEVT_CUDA_OK(cudaIpcCloseMemHandle(device_ptr));
    return 0;
}

GSEVT_API int gsevt_engine_split_attach(GsevtEngine* e, int32_t rank, int32_t n, void* const* boxes, double timeout_s) {
    if (!e || n < 1 || n > GSEVT_SPLIT_MAX || rank < 0 || rank >= n || (n > 1 && !boxes)) { set_error("split_attach: bad arguments"); return GSEVT_EINVAL; }
    GSEVT_CUDA_OK(cudaDeviceSynchronize());
    if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
    if (n == 1) {   // detach
        e->split_rank = 0; e->split_n = 1;
        if (e->comm) { dev_free(e, e->comm); e->comm = nullptr; }
        return 0;
    }
    if (!e->mailbox) { set_error("split_attach: call gsevt_engine_split_mailbox first"); return GSEVT_ESTATE; }
    if (boxes[rank] != (void*)e->mailbox) { set_error("split_attach: boxes[rank] must be this engine's own mailbox"); return GSEVT_EINVAL; }
    SplitComm h;
    memset(&h, 0, sizeof(h));
    h.rank = rank; h.n = n;
    h.timeout_ns = (unsigned long long)((timeout_s > 0.0 ? timeout_s : 5.0) * 1e9);
    for (int r = 0; r < n; r++) {
        if (!boxes[r]) { set_error("split_attach: box %d is NULL", r); return GSEVT_EINVAL; }
        h.box[r] = (MailBox*)boxes[r];
    }
    if (!e->comm) { int rc = dev_alloc(e, &e->comm, 1); if (rc) return rc; }
    if (!e->vis_list) {
        int rc = dev_alloc(e, &e->vis_list, 2 * (size_t)e->map->P) | dev_alloc(e, &e->vis_count, 2) |
                 dev_alloc(e, &e->surv_list, (size_t)e->map->P);
        if (rc) return rc;
        GSEVT_CUDA_OK(cudaMemset(e->vis_count, 0, 8));
    }
    GSEVT_CUDA_OK(cudaMemcpy(e->comm, &h, sizeof(h), cudaMemcpyHostToDevice));
    // a fresh group starts at sequence 0 with clean slots (every rank attaches before any rank iterates:
    // the caller puts a barrier between attach and the first collective call)
    GSEVT_CUDA_OK(cudaMemset(e->mailbox, 0, sizeof(MailBox)));
    GSEVT_CUDA_OK(cudaMemset((char*)e->ctl + offsetof(EngineCtl, comm_error), 0, 4));
    GSEVT_CUDA_OK(cudaDeviceSynchronize());
    e->split_rank = rank; e->split_n = n;
    return 0;
}

GSEVT_API int gsevt_engine_split_info(GsevtEngine* e, int32_t out6[6], void* stream) {
    if (!e || !out6) { set_error("bad arguments"); return GSEVT_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    int err = 0;
    unsigned long long seq[2] = {0, 0};
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));   // see gsevt_engine_status
    GSEVT_CUDA_OK(cudaMemcpyAsync(&err, (char*)e->ctl + offsetof(EngineCtl, comm_error), 4, cudaMemcpyDeviceToHost, s));
    if (e->comm) GSEVT_CUDA_OK(cudaMemcpyAsync(seq, (char*)e->comm + offsetof(SplitComm, seq), 16, cudaMemcpyDeviceToHost, s));
    GSEVT_CUDA_OK(cudaStreamSynchronize(s));
    out6[0] = e->split_rank; out6[1] = e->split_n; out6[2] = e->strip_y0; out6[3] = e->strip_y1; out6[4] = err;
    out6[5] = (int32_t)(seq[0] + seq[1]);
    return 0;
}

GSEVT_API int gsevt_engine_set_binning(GsevtEngine* e, int32_t mode) {
    if (!e || mode < 0 || mode > 2) { set_error("set_binning: bad arguments"); return GSEVT_EINVAL; }
    e->bin_mode = mode;   // takes effect at the next begin_level / eval / resume (the graph is keyed on the bucket shape)
    return 0;
}

GSEVT_API int gsevt_engine_launches_per_iteration(const GsevtEngine* e) {
    // preprocess_map, bucket_scatter, bucket_sort, blend_fwd (+ loss), blend_bwd, geom_compact, geom_bwd, engine_update:
    // all of them this library's own kernels; the screen-tile split runs the projection as two kernels (strip pre-test +
    // projection of the survivors)
    if (!e) return 8;
    const int loss_kernel = e->fuse_loss && e->strip_y1 > e->strip_y0 ? 0 : 1;   // the loss sums ride in the forward's epilogue
    return 8 + loss_kernel + (split_kernels(e) ? 1 : 0);
}

}  // extern "C"
