/*
 * gsevt.h — C ABI of libgsevt.so, the B200-native (sm_100a) implementation of the GS-EVT tracking
 * hot path.  Plain pointers and sizes only: no C++ types, no torch types.  Every device pointer is
 * caller-owned unless stated otherwise; `stream` is a cudaStream_t passed as void* (pass
 * torch.cuda.current_stream().cuda_stream).  All functions return 0 (or a non-negative count) on
 * success and a negative GSEVT_E* code on failure; gsevt_last_error() returns a thread-local
 * message.  Nothing in here falls back to the CPU.
 *
 * Citations are relative to the reference repository root (ChillTerry/GS-EVT), with
 * dgr/ = submodules/diff-gaussian-rasterization/.
 *
 * Matrix convention (same as the reference): a 4x4 matrix is 16 floats read COLUMN-major
 * (dgr/cuda_rasterizer/auxiliary.h:58-77); Python passes M.transpose(0,1) of its row-major tensor.
 */
#ifndef GSEVT_H_
#define GSEVT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSEVT_OK 0
#define GSEVT_EINVAL (-1)   /* bad argument combination */
#define GSEVT_ECUDA (-2)    /* a CUDA runtime call failed; see gsevt_last_error() */
#define GSEVT_ENOMEM (-3)   /* caller-provided buffer too small */
#define GSEVT_EOVERFLOW (-4)/* instance capacity of an engine exceeded */
#define GSEVT_ESTATE (-5)   /* call sequence violated */

#define GSEVT_ABI_VERSION 1

#if defined(__GNUC__)
#define GSEVT_API __attribute__((visibility("default")))
#else
#define GSEVT_API
#endif

GSEVT_API const char* gsevt_last_error(void);
GSEVT_API int gsevt_abi_version(void);
/* Compute capability major*10+minor of the current device, or negative error.  The library only
 * carries sm_100a code; callers use this to fail loudly elsewhere. */
GSEVT_API int gsevt_device_arch(void);

/* ------------------------------------------------------------------------------------------------
 * 1. The rasteriser operator.
 *    Replaces  _C.rasterize_gaussians / _C.rasterize_gaussians_backward / _C.mark_visible
 *    (dgr/ext.cpp:15-19, dgr/rasterize_points.h:18-75) and the C++ statics
 *    CudaRasterizer::Rasterizer::{forward,backward,markVisible} (dgr/cuda_rasterizer/rasterizer.h:20-97).
 *    The reference grows its work buffers through std::function resize callbacks
 *    (dgr/rasterize_points.cu:27-33); here the caller asks for sizes first and the forward is split
 *    at the one point where a size depends on data (num_rendered, dgr/cuda_rasterizer/rasterizer_impl.cu:284).
 * ---------------------------------------------------------------------------------------------- */

typedef struct GsevtRasterArgs {
    /* sizes */
    int32_t P;              /* number of Gaussians */
    int32_t sh_degree;      /* D: active SH degree 0..3 */
    int32_t sh_coeffs;      /* M: coefficients per channel stored in shs (0 if shs == NULL) */
    int32_t width, height;
    /* scalars (GaussianRasterizationSettings, dgr/diff_gaussian_rasterization/__init__.py:189-207) */
    float tanfovx, tanfovy;
    float scale_modifier;
    float delta_time;       /* backward only: +-dtau/2 */
    int32_t prefiltered;
    int32_t debug;          /* !=0: synchronise and check after every launch (auxiliary.h:166-173) */
    int32_t want_n_touched; /* !=0: maintain the n_touched atomics of forward.cu:368-371 */
    int32_t reserved0;
    /* inputs, device pointers (NULL selects the alternative, as the reference's empty tensors do) */
    const float* background;      /* [3] */
    const float* means3D;         /* [P,3] */
    const float* shs;             /* [P,M,3] or NULL */
    const float* colors_precomp;  /* [P,3]  or NULL */
    const float* opacities;       /* [P] */
    const float* scales;          /* [P,3] or NULL */
    const float* rotations;       /* [P,4] or NULL */
    const float* cov3D_precomp;   /* [P,6] or NULL */
    const float* viewmatrix;      /* [16] */
    const float* projmatrix;      /* [16] */
    const float* projmatrix_raw;  /* [16] backward only */
    const float* campos;          /* [3] */
    const float* vel_transform;     /* [16] backward only ("vel_transofrm") */
    const float* vel_transform_inv; /* [16] backward only */
    /* work buffers (sizes from gsevt_raster_sizes / gsevt_raster_binning_size) */
    void* geom_buffer;  size_t geom_bytes;
    void* img_buffer;   size_t img_bytes;
    void* binning_buffer; size_t binning_bytes;
    /* forward outputs */
    float* out_color;    /* [3,H,W] */
    float* out_depth;    /* [1,H,W] */
    float* out_opacity;  /* [1,H,W] */
    int32_t* radii;      /* [P] */
    int32_t* n_touched;  /* [P] (zero-initialised by the caller) or NULL */
    /* backward inputs */
    const float* dL_dout_color;  /* [3,H,W] */
    const float* dL_dout_depth;  /* [1,H,W] or NULL (== 0) */
    int32_t num_rendered;        /* R, as returned by forward */
    int32_t reserved1;
    /* backward outputs.  pose_grads is always written: 12 floats
     *   [0:3] d/drho  [3:6] d/dtheta  [6:9] d/dv  [9:12] d/dw
     * i.e. the column sums of the reference's dL_dtau / dL_dvel (P,6) tensors
     * (dgr/diff_gaussian_rasterization/__init__.py:163-169).  The per-Gaussian outputs are optional
     * (NULL = not wanted; the frozen map of GS-EVT never wants them). */
    float* pose_grads;       /* [12] */
    void* bwd_workspace;  size_t bwd_workspace_bytes; /* gsevt_raster_backward_workspace_size */
    float* dL_dmeans2D;      /* [P,3] or NULL */
    float* dL_dmeans3D;      /* [P,3] or NULL */
    float* dL_dopacity;      /* [P]   or NULL */
    float* dL_dcolors;       /* [P,3] or NULL (colors_precomp gradient) */
    float* dL_dcov3D;        /* [P,6] or NULL */
    float* dL_dsh;           /* [P,M,3] or NULL */
    float* dL_dscales;       /* [P,3] or NULL */
    float* dL_drotations;    /* [P,4] or NULL */
    float* dL_dtau;          /* [P,6] or NULL (un-reduced, reference layout) */
    float* dL_dvel;          /* [P,6] or NULL */
} GsevtRasterArgs;

/* geometry + image buffer sizes for (P, W, H).  Replaces required<GeometryState>/<ImageState>
 * (dgr/cuda_rasterizer/rasterizer_impl.h:63-72). */
GSEVT_API int gsevt_raster_sizes(int32_t P, int32_t width, int32_t height, size_t* geom_bytes, size_t* img_bytes);
/* binning buffer size for num_rendered instances (required<BinningState>). */
GSEVT_API size_t gsevt_raster_binning_size(int32_t num_rendered);
GSEVT_API size_t gsevt_raster_backward_workspace_size(int32_t P);

/* Forward, phase A: projection (preprocessCUDA, dgr/cuda_rasterizer/forward.cu:157-258), prefix sum
 * and the num_rendered read-back (rasterizer_impl.cu:280-284).  Synchronises `stream` once.
 * Returns num_rendered >= 0. */
GSEVT_API int gsevt_raster_forward_geometry(const GsevtRasterArgs* a, void* stream);
/* Forward, phase B: key emission, (tile, depth) sort, tile ranges and blending
 * (rasterizer_impl.cu:292-341, forward.cu:263-392).  a->binning_buffer must hold
 * gsevt_raster_binning_size(a->num_rendered) bytes.  Fully asynchronous. */
GSEVT_API int gsevt_raster_forward_render(const GsevtRasterArgs* a, void* stream);
/* Backward (rasterizer_impl.cu:348-467, backward.cu): blend backward, EWA/projection backward and the
 * pose / velocity chain reduced to 12 floats inside the kernels.  Fully asynchronous. */
GSEVT_API int gsevt_raster_backward(const GsevtRasterArgs* a, void* stream);
/* checkFrustum (rasterizer_impl.cu:54-66): present[i] = depth_i > 0.2 */
GSEVT_API int gsevt_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                       uint8_t* present, void* stream);

/* Introspection of the private work-buffer layouts, for the parity tests: byte offset of a named
 * array inside the buffer, measured from the buffer base rounded up to 256 bytes (negative if
 * unknown).  geom: "radii" i32[P],
 * "rec" f32[8P] = {x, y, conic A, conic B | conic C, opacity, gray, depth} per Gaussian,
 * "rgb4" f32[4P] = {r, g, b, gray}, "cov3D" f32[6P], "clamped" u8[P] (bit c = channel c clamped),
 * "tiles_touched" u32[P], "point_offsets" u32[P].  binning: "point_list_keys" u64[R],
 * "point_list" u32[R], "point_list_keys_unsorted", "point_list_unsorted".
 * img: "accum_alpha" f32[HW], "n_contrib" u32[HW], "ranges" u32[2*tiles]. */
GSEVT_API int64_t gsevt_raster_geom_offset(const char* name, int32_t P);
GSEVT_API int64_t gsevt_raster_binning_offset(const char* name, int32_t num_rendered);
GSEVT_API int64_t gsevt_raster_img_offset(const char* name, int32_t width, int32_t height);

/* ------------------------------------------------------------------------------------------------
 * 2. Event frames.  Replaces EventFrame.integrate_events (utils/event_camera/event.py:116-128) and
 *    Tracker.image_pyramid (utils/tracker.py:78-91), which run on the host with numpy + OpenCV.
 * ---------------------------------------------------------------------------------------------- */

/* Text ingestion.  Replaces the parsing half of load_events_from_txt (utils/event_camera/event.py:11-39:
 * readlines() + split() + int() + one Python object per event, 0.31 M events/s): converts the whitespace-separated
 * decimal integers of `text[0, len)` into out[0, n), in order, with `threads` host threads (<= 0: all cores).
 * out == NULL: returns n without converting (size query).  Returns n, GSEVT_EINVAL on a token that is not an integer
 * (the reference raises ValueError), GSEVT_ENOMEM when n > capacity.  Host pointers; no device involved. */
GSEVT_API int64_t gsevt_parse_int_table(const char* text, size_t len, int64_t* out, size_t capacity, int32_t threads);
/* E0: counts[y*W+x] += p ? +1 : -1  (integer scatter-add, bit-exact).  x,y int16, p uint8, device
 * pointers; counts int32[H*W] must be zeroed by the caller (or pass zero_first != 0).  Coordinates follow
 * numpy's indexing in the reference's loop (frame[y, x] += ..., event.py:118-120): [-size, -1] counts from the
 * end; anything outside [-size, size) is an IndexError in the reference; here it sets the returned device
 * flag *oob (may be NULL) and the event is dropped. */
GSEVT_API int gsevt_event_accumulate(const int16_t* x, const int16_t* y, const uint8_t* p, int32_t n,
                           int32_t width, int32_t height, int32_t* counts, int32_t zero_first,
                           int32_t* oob, void* stream);
/* Fixed-point inverse map of cv2.undistort(K, D) (1/32 px): host helper, fills map_ix/map_iy
 * int32[H*W] (host pointers).  K is 9 doubles row-major, D is 5 doubles. */
GSEVT_API int gsevt_event_undistort_map(const double* K, const double* D, int32_t width, int32_t height,
                              int32_t* map_ix, int32_t* map_iy);
/* E1-E4 + pyramid: counts -> undistort -> 9x9 fixed Gaussian -> L2 normalise -> signed / unsigned
 * frames at `levels` pyramid levels.  map_ix/map_iy are device copies of the map above.
 * sign_out / unsign_out: float32, level l stored at offset sum_{k<l} (W>>k)*(H>>k).
 * scratch: float32[2*H*W] + 64 doubles. */
GSEVT_API int gsevt_event_frame(const int32_t* counts, const int32_t* map_ix, const int32_t* map_iy,
                      int32_t width, int32_t height, int32_t levels,
                      float* sign_out, float* unsign_out, void* scratch, size_t scratch_bytes, void* stream);
/* The same with cv2.GaussianBlur's kernel size as a parameter (the yaml's Event.gaussian_kernel_size, reference
 * utils/event_camera/event.py:122-123): 1, 3, 5, 7 or 9 — the sizes whose OpenCV coefficients (sigma = 0) are multiples of
 * 1/256, so that the blur is exact in fp32 and bit-identical to cv2 on any machine. */
GSEVT_API int gsevt_event_frame_k(const int32_t* counts, const int32_t* map_ix, const int32_t* map_iy,
                        int32_t width, int32_t height, int32_t levels, int32_t ksize,
                        float* sign_out, float* unsign_out, void* scratch, size_t scratch_bytes, void* stream);
GSEVT_API size_t gsevt_event_frame_scratch_size(int32_t width, int32_t height);

/* ------------------------------------------------------------------------------------------------
 * 3. The fused tracking engine.  Replaces the body of Tracker.tracking's innermost loop
 *    (utils/tracker.py:176-240): RenderFrame.get_delta_Ir (utils/render_camera/frame.py:61-94),
 *    render2 / build_rasterizer (gaussian_splatting/gaussian_renderer/__init__.py:234-339), the
 *    Camera / SE3 algebra (utils/render_camera/camera.py:100-155, utils/pose.py:26-91),
 *    Tracker.tracking_loss (utils/tracker.py:93-103), loss.backward(), torch.optim.Adam.step,
 *    check_convergence (utils/tracker.py:65-76) and update_pose / update_vwRT.
 *    One engine = one camera hypothesis on one device; the map is shared and read-only.
 * ---------------------------------------------------------------------------------------------- */

typedef struct GsevtMap GsevtMap;        /* packed, activated, frozen map resident in HBM */
typedef struct GsevtEngine GsevtEngine;  /* per-hypothesis tracking state + workspaces */

/* Pack a map.  Inputs are the ACTIVATED tensors the reference feeds the rasteriser each iteration
 * (gaussian_renderer/__init__.py:277-297): xyz [P,3], scales = exp(_scaling) [P,3], rotations =
 * normalize(_rotation) [P,4], opacities = sigmoid(_opacity) [P], shs = cat(dc, rest) [P,16,3].
 * Device pointers; copied into the library's own SoA layout (library-owned memory). */
GSEVT_API int gsevt_map_create(int32_t P, int32_t sh_degree, const float* xyz, const float* scales, const float* rotations,
                     const float* opacities, const float* shs, float scale_modifier, void* stream,
                     GsevtMap** out);
GSEVT_API void gsevt_map_destroy(GsevtMap* m);
GSEVT_API int32_t gsevt_map_size(const GsevtMap* m);
GSEVT_API size_t gsevt_map_bytes(const GsevtMap* m);

typedef struct GsevtEngineConfig {
    int32_t width, height;       /* level-0 image size */
    int32_t levels;              /* pyramid levels (reference: 3, utils/tracker.py:58) */
    float fx, fy;                /* Gaussian.calib_params */
    float znear, zfar;           /* camera.py:58-59: 0.01 / 100 */
    float background[3];
    float lr_rot, lr_trans, lr_w, lr_v;   /* Optimizer.* */
    float converged_threshold;   /* Optimizer.converged_threshold */
    int32_t max_optim_iter;      /* Optimizer.max_optim_iter */
    int32_t instance_capacity;   /* max tile instances per view (0 = choose from P) */
    int32_t reserved[6];
} GsevtEngineConfig;

GSEVT_API int gsevt_engine_create(const GsevtMap* map, const GsevtEngineConfig* cfg, GsevtEngine** out);
GSEVT_API void gsevt_engine_destroy(GsevtEngine* e);

/* Pose/velocity state: R row-major 3x3 world->camera, T[3], angular_vel[3], linear_vel[3] (host). */
GSEVT_API int gsevt_engine_set_state(GsevtEngine* e, const float* R, const float* T, const float* angular_vel,
                           const float* linear_vel, void* stream);
GSEVT_API int gsevt_engine_get_state(GsevtEngine* e, float* R, float* T, float* angular_vel, float* linear_vel,
                           void* stream);
/* Start a new event frame: sets delta_tau, the event-frame pyramids (device pointers in the layout
 * written by gsevt_event_frame) and resets Adam (fresh optimiser per frame, utils/tracker.py:117-129). */
GSEVT_API int gsevt_engine_begin_frame(GsevtEngine* e, double delta_tau, const float* sign_pyr, const float* unsign_pyr,
                             void* stream);
/* Start optimising a pyramid level; opt_vel = 0 starts in the coarse (pose-only, unsigned) stage,
 * 1 in the fine stage (utils/tracker.py:149-174). */
GSEVT_API int gsevt_engine_begin_level(GsevtEngine* e, int32_t level, int32_t opt_vel, void* stream);
/* Enqueue n optimisation iterations (each: 2 renders, loss, backward, Adam, pose update, convergence
 * logic).  Iterations after the level has finished are no-ops on the device.  No host sync. */
GSEVT_API int gsevt_engine_iterate(GsevtEngine* e, int32_t n, void* stream);

typedef struct GsevtEngineStatus {
    int32_t level_done;       /* 1 once the fine stage converged or hit the cap */
    int32_t optim_iter;       /* reference's optim_iter at loop exit / so far */
    int32_t start_vel_opt_iter;
    int32_t opt_vel;          /* current stage */
    int32_t iters_executed;   /* iterations actually executed in this level */
    int32_t overflow;         /* instance capacity exceeded (results invalid) */
    int32_t num_rendered[2];  /* last / next view, most recent iteration */
    float last_loss;
    float pose_grads[12];     /* most recent: rho, theta, v, w (summed over both views) */
    float reserved[4];
} GsevtEngineStatus;
/* Copies the status block to the host (synchronises `stream`). */
GSEVT_API int gsevt_engine_status(GsevtEngine* e, GsevtEngineStatus* out, void* stream);
/* Per-iteration loss history of the current level (device->host, synchronises): the most recent
 * min(iterations, capacity, 1024) losses in order (the device keeps a ring of 1024; the stopping rule of
 * utils/tracker.py:65-76 only needs the last 11).  Returns the count. */
GSEVT_API int gsevt_engine_losses(GsevtEngine* e, float* out, int32_t capacity, void* stream);
/* End-of-frame velocity blend (Camera.cal_weighted_velocity, camera.py:157-181) and the constant
 * velocity prediction (Camera.const_vel_model, camera.py:183-201), on the device. */
GSEVT_API int gsevt_engine_const_vel_model(GsevtEngine* e, double tau, void* stream);
GSEVT_API int gsevt_engine_weighted_velocity(GsevtEngine* e, const float* last_R, const float* last_T, double delta_tau,
                                   double weight, void* stream);
/* Debug / parity access: renders the normalised signed delta frame of the current state at `level`
 * into out[(H>>level)*(W>>level)] (device pointer) and the two grayscale views if non-NULL. */
GSEVT_API int gsevt_engine_render_delta(GsevtEngine* e, int32_t level, float* delta_out, float* gray_last, float* gray_next,
                              void* stream);
/* Parity access: per-pixel blending state of the most recent evaluation at `level`, both views — final transmittance
 * and last-contributor index (the reference's accum_alpha / n_contrib, rasterizer_impl.h:47-58): device pointers,
 * [2][H*W] each, either may be NULL. */
GSEVT_API int gsevt_engine_image_state(GsevtEngine* e, int32_t level, float* final_T, uint32_t* n_contrib, void* stream);
/* Parity access: the camera block the device-side pose algebra produced for `view` (0 = last, 1 = next) in the most
 * recent evaluation — what render2 / build_rasterizer hand to the rasteriser (gaussian_renderer/__init__.py:318-337):
 * out73 (host) = viewmatrix[16], projmatrix[16] (column-major), campos[3], tanfovx, tanfovy, projmatrix_raw[0], [5], [11],
 * vel_transofrm[16], vel_transofrm_inv[16], delta_time.  Synchronises. */
GSEVT_API int gsevt_engine_view_params(GsevtEngine* e, int32_t view, float* out73, void* stream);
/* One gradient evaluation without optimiser step (parity tests): loss and the 12 pose gradients. */
GSEVT_API int gsevt_engine_eval(GsevtEngine* e, int32_t level, int32_t signed_loss, float* loss_out, float* grads_out12,
                      void* stream);
/* Parity access to the engine's binning of the most recent evaluation / iteration: the sorted instance list
 * of `view` in the reference's representation — 64-bit keys (tile << 32 | depth bits), Gaussian ids, and the
 * tile ranges into that list (rasterizer_impl.cu:70-138).  Device output pointers: keys_out[capacity],
 * list_out[capacity], ranges_out[2 * tiles].  Returns the number of instances.  Synchronises. */
GSEVT_API int gsevt_engine_binning(GsevtEngine* e, int32_t view, uint64_t* keys_out, uint32_t* list_out, uint32_t* ranges_out,
                         int32_t capacity, void* stream);
/* Profiling (bench.py's roofline leg): runs n_iters iterations OUTSIDE the CUDA graph with a CUDA
 * event between stages on `stream` and returns the mean device time of each stage in ms
 * (gsevt_engine_stage_count() entries, names from gsevt_engine_stage_name).  The iterations are real
 * (state advances).  Synchronises. */
GSEVT_API int gsevt_engine_stage_count(void);
GSEVT_API const char* gsevt_engine_stage_name(int32_t i);
GSEVT_API int gsevt_engine_profile(GsevtEngine* e, int32_t n_iters, float* stage_ms, void* stream);
/* Data-dependent work of the most recent iteration: out8 = {visible Gaussians view 0, view 1, tile
 * instances view 0, view 1, sum of n_contrib view 0, view 1, (view, Gaussian) pairs with a non-zero blend
 * gradient, key slots of the bucket segments (bucket keys + slack)}.  Synchronises. */
GSEVT_API int gsevt_engine_workload(GsevtEngine* e, int64_t* out8, void* stream);
/* Number of kernels the engine launches per executed iteration (for bench.py's gpu_launches). */
GSEVT_API int gsevt_engine_launches_per_iteration(const GsevtEngine* e);
/* Bucket shape of the engine's binning from the next begin_level / eval / resume on: 0 = automatic (buckets of 2 x 2
 * tiles while a bucket sorts in shared memory, one tile per bucket on coarse pyramid levels / very dense maps), 1 = one
 * tile per bucket, 2 = 2 x 2 tiles per bucket.  All produce the reference's per-tile lists (rasterizer_impl.cu:70-138) bit
 * for bit; the knob exists so that tests can compare them. */
GSEVT_API int gsevt_engine_set_binning(GsevtEngine* e, int32_t mode);
/* After gsevt_engine_poll_done() returned 2: a bucket of the binning outgrew its key segment
 * (the pose moved a lot inside one level).  The device voided that iteration (state untouched) and paused the
 * level; this call re-counts at the current pose, grows the buffers and clears the pause.  Synchronises. */
GSEVT_API int gsevt_engine_resume(GsevtEngine* e, void* stream);
/* Non-blocking: 1 once the device has flagged the current level as finished, 2 when it paused for
 * gsevt_engine_resume (read from mapped pinned memory, no stream synchronisation). */
GSEVT_API int gsevt_engine_poll_done(GsevtEngine* e);

/* ------------------------------------------------------------------------------------------------
 * 4. Screen-tile split of ONE hypothesis over the GPUs of a box (BASELINE.json configs[4]).
 *    The reference has no multi-GPU path; the only reference arithmetic this touches is the whole-image
 *    L2 normalisation of the rendered difference (utils/render_camera/frame.py:90-91) and the sum over
 *    Gaussians of dL/dtau, dL/dvel (dgr/diff_gaussian_rasterization/__init__.py:163-167).
 *    One engine per rank, identical state and event frame everywhere.  Rank r bins, blends and scores
 *    a contiguous strip of tile rows (balanced on tile instances per row at every begin_level); the
 *    three loss sums and the 12 gradient sums (+ the overflow flag) are exchanged INSIDE the loss and
 *    update kernels by stores into the peers' mailboxes over NVLink and added in rank order, so all
 *    ranks apply bit-identical optimiser steps and an iteration is still one CUDA-graph launch.
 *    Every engine call after attach is collective: all ranks make the same calls in the same order.
 * ---------------------------------------------------------------------------------------------- */
/* This engine's mailbox: library-owned device memory of gsevt_split_mailbox_bytes() bytes in its own
 * cudaMalloc allocation (so that it can be exported).  Created on first call. */
GSEVT_API int gsevt_engine_split_mailbox(GsevtEngine* e, void** box);
GSEVT_API size_t gsevt_split_mailbox_bytes(void);
/* CUDA IPC plumbing for one-process-per-GPU groups: export a 64-byte handle of a mailbox, open a
 * peer's handle (enables peer access lazily), close it again before the owner is destroyed. */
GSEVT_API int gsevt_ipc_export(const void* device_ptr, uint8_t handle64[64]);
GSEVT_API int gsevt_ipc_open(const uint8_t handle64[64], void** device_ptr);
GSEVT_API int gsevt_ipc_close(void* device_ptr);
/* boxes[n]: device pointers, valid on this engine's device, of all ranks' mailboxes in rank order
 * (boxes[rank] = this engine's own).  Resets the exchange sequence: every rank must attach before any
 * rank iterates (put a barrier in between).  n = 1 detaches.  timeout_s <= 0 selects 5 s: a peer that
 * does not answer in time sets level_done = 3 / poll_done() = 3 instead of hanging the GPU. */
GSEVT_API int gsevt_engine_split_attach(GsevtEngine* e, int32_t rank, int32_t n, void* const* boxes, double timeout_s);
/* out6 = {rank, n, first tile row, end tile row of the current level's strip, comm error flag,
 * exchanges completed}.  Synchronises. */
GSEVT_API int gsevt_engine_split_info(GsevtEngine* e, int32_t out6[6], void* stream);
/* The strip partition itself (host code): bounds[k] .. bounds[k+1] = tile rows of rank k, balanced on
 * row_cost (tile instances per tile row).  rows <= 255, n <= 8. */
GSEVT_API int gsevt_split_balance_rows(const uint32_t* row_cost, int32_t rows, int32_t n, int32_t* bounds);

#ifdef __cplusplus
}
#endif
#endif /* GSEVT_H_ */
