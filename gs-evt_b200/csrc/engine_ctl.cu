// Single-thread control kernels of the tracking engine: they keep the whole optimisation loop of
// Tracker.tracking (utils/tracker.py:176-240) on the device, so an iteration needs no host round
// trip (the reference performs ~570 tiny launches, 18 4x4 inversions and ~60 host syncs here).
//
//  pose_setup   : Camera.last/next_vel_transform (camera.py:100-123), RenderFrame pose composition
//                 (frame.py:68-82), world_view / full_proj / camera_center (camera.py:66-92,
//                 graphics_utils.py:33-46) and the GaussianRasterizationSettings of
//                 gaussian_renderer/__init__.py:266-275,313-339 -> two ViewParams blocks.
//  update       : pose-gradient reduction, torch.optim.Adam.step (defaults), check_convergence
//                 (tracker.py:65-76), update_pose / update_vwRT (camera.py:129-155), the fine-stage LR
//                 cross-fade (tracker.py:188-202) and the stage / iteration-cap logic (:224-240).
//  const_vel / weighted_velocity : camera.py:157-201.
// All arithmetic is fp32 in torch's operation order unless noted (python scalars are doubles that
// torch rounds to fp32 when they meet an fp32 tensor).
#include "internal.h"
#include "split_comm.cuh"

namespace gsevt {

namespace {

struct M3 { float m[3][3]; };  // row-major
struct SE3f { M3 R; float t[3]; };

__device__ M3 mat_mul(const M3& a, const M3& b) {
    M3 c;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) c.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return c;
}
__device__ M3 skew(const float* x) {  // pose.py:13-23
    M3 s;
    s.m[0][0] = 0; s.m[0][1] = -x[2]; s.m[0][2] = x[1];
    s.m[1][0] = x[2]; s.m[1][1] = 0; s.m[1][2] = -x[0];
    s.m[2][0] = -x[1]; s.m[2][1] = x[0]; s.m[2][2] = 0;
    return s;
}
// SE3_exp (pose.py:79-91) with SO3_exp (:26-41) and V (:61-76); xi = [rho; theta]
__device__ SE3f se3_exp(const float* xi) {
    const float* rho = xi;
    const float* th = xi + 3;
    const M3 W = skew(th);
    const M3 W2 = mat_mul(W, W);
    const float angle = sqrtf(th[0] * th[0] + th[1] * th[1] + th[2] * th[2]);
    float a1, a2, b1, b2;
    if (angle < 1e-5f) {
        a1 = 1.0f; a2 = 0.5f; b1 = 0.5f; b2 = 1.0f / 6.0f;
    } else {
        const float sn = sinf(angle), cs = cosf(angle);
        a1 = sn / angle;
        a2 = (1.0f - cs) / (angle * angle);
        b1 = (1.0f - cs) / (angle * angle);
        b2 = (angle - sn) / (angle * angle * angle);
    }
    SE3f T;
    M3 V;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const float I = i == j ? 1.0f : 0.0f;
            T.R.m[i][j] = (I + a1 * W.m[i][j]) + a2 * W2.m[i][j];
            V.m[i][j] = (I + W.m[i][j] * b1) + W2.m[i][j] * b2;
        }
    for (int i = 0; i < 3; i++) T.t[i] = V.m[i][0] * rho[0] + V.m[i][1] * rho[1] + V.m[i][2] * rho[2];
    return T;
}
// A @ B for rigid transforms
__device__ SE3f se3_mul(const SE3f& a, const SE3f& b) {
    SE3f c;
    c.R = mat_mul(a.R, b.R);
    for (int i = 0; i < 3; i++) c.t[i] = a.R.m[i][0] * b.t[0] + a.R.m[i][1] * b.t[1] + a.R.m[i][2] * b.t[2] + a.t[i];
    return c;
}
__device__ SE3f se3_inv(const SE3f& a) {
    SE3f c;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) c.R.m[i][j] = a.R.m[j][i];
    for (int i = 0; i < 3; i++) c.t[i] = -(c.R.m[i][0] * a.t[0] + c.R.m[i][1] * a.t[1] + c.R.m[i][2] * a.t[2]);
    return c;
}
__device__ void to_colmajor(const SE3f& T, float* out) {
    for (int c = 0; c < 3; c++) {
        for (int r = 0; r < 3; r++) out[4 * c + r] = T.R.m[r][c];
        out[4 * c + 3] = 0.0f;
    }
    out[12] = T.t[0]; out[13] = T.t[1]; out[14] = T.t[2]; out[15] = 1.0f;
}
__device__ SE3f load_pose(const EngineCtl* c) {
    SE3f T;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) T.R.m[i][j] = c->R[3 * i + j];
        T.t[i] = c->T[i];
    }
    return T;
}
__device__ void store_pose(EngineCtl* c, const SE3f& T) {
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) c->R[3 * i + j] = T.R.m[i][j];
        c->T[i] = T.t[i];
    }
}

}  // namespace

// Builds both ViewParams blocks from the current state (single thread).
__device__ void pose_setup_device(EngineCtl* ctl, ViewParams* views, const float* bg3) {
    if (!ctl->eval_only) ctl->loss_signed = ctl->opt_vel;
    const float s = ctl->half_dtau;
    float rot[3], tr[3];
    for (int i = 0; i < 3; i++) {
        rot[i] = ctl->ang_vel[i] * s;   // camera.py:125-127
        tr[i] = ctl->lin_vel[i] * s;
    }
    const SE3f cur = load_pose(ctl);
    for (int v = 0; v < 2; v++) {
        const float sign = v == 0 ? -1.0f : 1.0f;
        float xi[6] = {sign * tr[0], sign * tr[1], sign * tr[2], sign * rot[0], sign * rot[1], sign * rot[2]};
        const SE3f Tvel = se3_exp(xi);
        const SE3f Tinv = se3_inv(Tvel);
        const SE3f Tk = se3_mul(Tvel, cur);   // frame.py:68-69
        ViewParams& vp = views[v];
        to_colmajor(Tk, vp.view);
        to_colmajor(Tvel, vp.vel);
        to_colmajor(Tinv, vp.vel_inv);
        // full_proj = P * T_k  (camera.py:81-87); proj_raw holds P column-major
        const float* P = ctl->proj_raw;
        for (int c = 0; c < 4; c++)
            for (int r = 0; r < 4; r++) {
                float acc = 0.0f;
                for (int j = 0; j < 4; j++) acc += P[4 * j + r] * vp.view[4 * c + j];
                vp.proj[4 * c + r] = acc;
            }
        // camera centre = -R^T t (camera.py:89-91)
        for (int i = 0; i < 3; i++)
            vp.campos[i] = -(Tk.R.m[0][i] * Tk.t[0] + Tk.R.m[1][i] * Tk.t[1] + Tk.R.m[2][i] * Tk.t[2]);
        vp.tanfovx = ctl->tanfovx; vp.tanfovy = ctl->tanfovy;
        vp.focal_x = ctl->focal_x; vp.focal_y = ctl->focal_y;
        vp.W = ctl->W; vp.H = ctl->H; vp.grid_x = ctl->grid_x; vp.grid_y = ctl->grid_y;
        vp.proj_a = P[0]; vp.proj_b = P[5]; vp.proj_e = P[11];
        vp.delta_time = sign * s;
        vp.bg[0] = bg3[0]; vp.bg[1] = bg3[1]; vp.bg[2] = bg3[2];
        vp.pad_ = 0.0f;
    }
}

// Stand-alone form: run whenever the state or the level changed from the host (set_state, begin_level, eval,
// resume, const_vel_model, weighted_velocity).  Inside the iteration loop the update kernel refreshes the views
// itself, so an iteration has no separate pose launch.
__global__ void pose_setup_kernel(EngineCtl* ctl, ViewParams* views, const float* bg3) {
    if (threadIdx.x != 0) return;
    pose_setup_device(ctl, views, bg3);
}
void launch_pose_setup(EngineCtl* ctl, ViewParams* views, const float* bg3, float, float, cudaStream_t s) {
    pose_setup_kernel<<<1, 32, 0, s>>>(ctl, views, bg3);
}

// ---- Adam --------------------------------------------------------------------------------------
__device__ void adam_group(EngineCtl* c, int group, const float* grad, double lr, float* delta_out) {
    // torch/optim/adam.py::_single_tensor_adam with beta=(0.9,0.999), eps=1e-8, weight_decay=0, amsgrad off
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    c->adam_step[group] += 1;
    const int step = c->adam_step[group];
    const double bc1 = 1.0 - pow(b1, (double)step);
    const double bc2 = 1.0 - pow(b2, (double)step);
    const double step_size = lr / bc1;
    const float bc2_sqrt = (float)sqrt(bc2);
    const float w1 = (float)(1.0 - b1), w2 = (float)(1.0 - b2), fb2 = (float)b2;
    for (int i = 0; i < 3; i++) {
        float& m = c->adam_m[3 * group + i];
        float& v = c->adam_v[3 * group + i];
        const float g = grad[i];
        m = m + w1 * (g - m);                    // exp_avg.lerp_(grad, 1 - beta1)
        v = v * fb2 + w2 * (g * g);              // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
        const float denom = sqrtf(v) / bc2_sqrt + (float)eps;
        delta_out[i] = 0.0f + (float)(-step_size) * (m / denom);   // param.addcdiv_(exp_avg, denom, value=-step_size)
    }
}

// Optimiser step, convergence test, pose / velocity update and stage logic of one iteration (single thread).  `c` points
// at a copy of the control block WITHOUT the loss history (shared memory); the history stays in global memory.
__device__ void engine_control(EngineCtl* c, float* losses, const float* s_g, int* host_flag, ViewParams* views,
                               const float* bg3) {
    for (int k = 0; k < GSEVT_NPART; k++) c->grads[k] = s_g[k];
    if (c->eval_only) {
        c->level_done = 1;   // one-shot
        return;
    }
    c->iters_executed += 1;
    // the history is a ring of the last GSEVT_MAX_LOSSES losses: the stopping rule only ever looks at the last 11
    // (tracker.py:65-76), so levels of any length keep honouring it
    losses[c->n_losses % GSEVT_MAX_LOSSES] = c->last_loss;
    c->n_losses += 1;

    // learning rates (tracker.py:161-170,188-202)
    double lr[4] = {c->lr_base[0], c->lr_base[1], c->lr_base[2], c->lr_base[3]};
    if (c->opt_vel) {
        const int k = c->optim_iter - c->start_vel_opt_iter;
        const double fraction_num = c->max_optim_iter / 2.0;
        const double fraction = (k >= 1 && (double)k <= fraction_num) ? (double)k / fraction_num : 1.0;
        lr[0] *= fraction; lr[1] *= fraction; lr[2] *= (1.0 - fraction); lr[3] *= (1.0 - fraction);
    }
    // gradient slots: grads[0:3] rho -> cam_trans_delta, [3:6] theta -> cam_rot_delta,
    //                 [6:9] v -> cam_v_delta, [9:12] w -> cam_w_delta
    float d_rot[3] = {0, 0, 0}, d_trans[3] = {0, 0, 0}, d_w[3] = {0, 0, 0}, d_v[3] = {0, 0, 0};
    adam_group(c, 0, c->grads + 3, lr[0], d_rot);
    adam_group(c, 1, c->grads + 0, lr[1], d_trans);
    if (c->opt_vel) {
        adam_group(c, 2, c->grads + 9, lr[2], d_w);
        adam_group(c, 3, c->grads + 6, lr[3], d_v);
    }
    // check_convergence (tracker.py:65-76): mean |diff| of the last 11 losses, in double
    bool converged = false;
    if (c->n_losses > 10) {
        double acc = 0.0;
        for (int i = c->n_losses - 10; i < c->n_losses; i++)
            acc += fabs((double)losses[i % GSEVT_MAX_LOSSES] - (double)losses[(i - 1) % GSEVT_MAX_LOSSES]);
        converged = (acc / 10.0) < (double)c->converged_threshold;
    }
    // update_vwRT / update_pose (camera.py:129-155)
    if (c->opt_vel) {
        for (int i = 0; i < 3; i++) {
            c->ang_vel[i] += d_w[i];
            c->lin_vel[i] += d_v[i];
        }
    }
    {
        const float xi[6] = {d_trans[0], d_trans[1], d_trans[2], d_rot[0], d_rot[1], d_rot[2]};
        const SE3f nw = se3_mul(se3_exp(xi), load_pose(c));
        store_pose(c, nw);
    }
    // stage logic (tracker.py:224-240)
    bool done = false;
    if (converged) {
        if (!c->opt_vel) {
            c->opt_vel = 1;
            c->start_vel_opt_iter = c->optim_iter;
        } else {
            done = true;
        }
    }
    if (!done) {
        if (!c->opt_vel) {
            if (c->optim_iter >= c->max_optim_iter) done = true;
        } else if (c->optim_iter >= c->start_vel_opt_iter + c->max_optim_iter) {
            done = true;
        }
    }
    if (done) {
        c->level_done = 1;
        if (host_flag) *host_flag = 1;
        __threadfence_system();
    } else {
        c->optim_iter += 1;
        pose_setup_device(c, views, bg3);   // the next iteration's two views
    }
}

__global__ void __launch_bounds__(32 * GSEVT_NPART) engine_update_kernel(EngineCtl* ctl,
                                                                          const float* __restrict__ partials,
                                                                          int nblocks, int* host_flag,
                                                                          const int* __restrict__ overflow,
                                                                          ViewParams* views, const float* bg3,
                                                                          SplitComm* comm) {
    pdl_prologue();
    if (ctl->level_done) return;
    __shared__ double s_d[16];
    __shared__ double s_t[16];
    __shared__ float s_g[GSEVT_NPART];
    __shared__ int s_overflow;
    {
        const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
        double s = 0.0;
        for (int b = lane; b < nblocks; b += 32) s += (double)partials[(size_t)k * nblocks + b];   // [12][nblocks]: coalesced
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) { s_d[k] = s; s_g[k] = (float)s; }
    }
    if (threadIdx.x == 0) s_overflow = overflow ? *overflow : 0;
    __syncthreads();
    if (comm) {
        // screen-tile split: s_d holds the gradient of this rank's strip.  All ranks exchange {12 sums, overflow flag}
        // over peer memory and add them in rank order, so the replicated Adam / pose state stays bit-identical; an
        // overflow on ANY rank voids the iteration on ALL of them.
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) s_d[GSEVT_NPART] = (double)s_overflow;
            __syncwarp();
            const bool ok = split_exchange(comm, 1, GSEVT_NPART + 1, s_d, s_t);
            __syncwarp();
            if (threadIdx.x < GSEVT_NPART) s_g[threadIdx.x] = (float)s_t[threadIdx.x];
            if (threadIdx.x == 0) {
                s_overflow = s_t[GSEVT_NPART] != 0.0 ? 1 : 0;
                if (!ok) {
                    ctl->comm_error = 1;
                    ctl->level_done = 3;
                    if (host_flag) *host_flag = 3;
                    __threadfence_system();
                    s_overflow = -1;
                }
            }
        }
        __syncthreads();
        if (s_overflow < 0) return;
    }
    if (s_overflow) {
        // The instance list outgrew the slots sorted this iteration: void the iteration (state untouched), pause
        // the level (2) and tell the host, which re-sizes and resumes (gsevt_engine_resume).
        if (threadIdx.x == 0) {
            ctl->level_done = 2;
            ctl->overflow = 1;
            if (host_flag) *host_flag = 2;
            __threadfence_system();
        }
        return;
    }
    // The control block (everything before the loss history) is staged in shared memory: the serial control code
    // below makes ~200 dependent accesses to it, each an L2 round trip when made in place.  All threads copy it in,
    // thread 0 runs the control logic on the copy, all threads copy it back.
    constexpr int HOT_WORDS = (int)(offsetof(EngineCtl, losses) / 4);
    __shared__ __align__(16) uint32_t s_ctl[HOT_WORDS];
    for (int i = threadIdx.x; i < HOT_WORDS; i += blockDim.x) s_ctl[i] = reinterpret_cast<const uint32_t*>(ctl)[i];
    // The two camera blocks of the next iteration are built in shared memory as well (the projection product reads back
    // the view matrix it has just written: in place, 2 x 64 dependent L2 round trips) and copied out by all threads.
    constexpr int VIEW_WORDS = (int)(2 * sizeof(ViewParams) / 4);
    __shared__ __align__(16) uint32_t s_views[VIEW_WORDS];
    __shared__ float s_bg[3];
    if (threadIdx.x < 3) s_bg[threadIdx.x] = bg3[threadIdx.x];
    for (int i = threadIdx.x; i < VIEW_WORDS; i += blockDim.x) s_views[i] = reinterpret_cast<const uint32_t*>(views)[i];
    __syncthreads();
    if (threadIdx.x == 0)
        engine_control(reinterpret_cast<EngineCtl*>(s_ctl), ctl->losses, s_g, host_flag, reinterpret_cast<ViewParams*>(s_views), s_bg);
    __syncthreads();
    for (int i = threadIdx.x; i < HOT_WORDS; i += blockDim.x) reinterpret_cast<uint32_t*>(ctl)[i] = s_ctl[i];
    for (int i = threadIdx.x; i < VIEW_WORDS; i += blockDim.x) reinterpret_cast<uint32_t*>(views)[i] = s_views[i];
}

void launch_engine_update(EngineCtl* ctl, const float* partials, int nblocks, int* host_flag, const int* overflow,
                          ViewParams* views, const float* bg3, SplitComm* comm, cudaStream_t s) {
    launch_k(engine_update_kernel, dim3(1), dim3(32 * GSEVT_NPART), 0, s, ctl, partials, nblocks, host_flag, overflow, views, bg3, comm);
}

// ---- per-frame helpers ----------------------------------------------------------------------------
__global__ void const_vel_kernel(EngineCtl* c, float tau) {  // camera.py:183-201
    if (threadIdx.x != 0) return;
    const float xi[6] = {c->lin_vel[0] * tau, c->lin_vel[1] * tau, c->lin_vel[2] * tau,
                         c->ang_vel[0] * tau, c->ang_vel[1] * tau, c->ang_vel[2] * tau};
    store_pose(c, se3_mul(se3_exp(xi), load_pose(c)));
}
void launch_const_vel(EngineCtl* ctl, float tau, cudaStream_t s) { const_vel_kernel<<<1, 32, 0, s>>>(ctl, tau); }

__global__ void weighted_velocity_kernel(EngineCtl* c, const float* lastRT, float delta_tau, float weight) {
    // camera.py:157-181; lastRT = R(9 row-major) + T(3) on the device
    if (threadIdx.x != 0) return;
    SE3f last;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) last.R.m[i][j] = lastRT[3 * i + j];
        last.t[i] = lastRT[9 + i];
    }
    const SE3f d = se3_mul(load_pose(c), se3_inv(last));
    // SO3_log (pose.py:44-58)
    float th = (d.R.m[0][0] + d.R.m[1][1] + d.R.m[2][2] - 1.0f) / 2.0f;
    th = fminf(fmaxf(th, -1.0f), 1.0f);
    const float theta = acosf(th);
    float rv[3] = {0, 0, 0};
    if (!(fabsf(theta) < 1e-5f)) {
        const float k = 2.0f * sinf(theta);
        rv[0] = theta * ((d.R.m[2][1] - d.R.m[1][2]) / k);
        rv[1] = theta * ((d.R.m[0][2] - d.R.m[2][0]) / k);
        rv[2] = theta * ((d.R.m[1][0] - d.R.m[0][1]) / k);
    }
    for (int i = 0; i < 3; i++) {
        const float lin = d.t[i] / delta_tau, ang = rv[i] / delta_tau;
        c->lin_vel[i] = weight * lin + (1.0f - weight) * c->lin_vel[i];
        c->ang_vel[i] = weight * ang + (1.0f - weight) * c->ang_vel[i];
    }
}
void launch_weighted_velocity(EngineCtl* ctl, const float* lastRT, float delta_tau, float weight, cudaStream_t s) {
    weighted_velocity_kernel<<<1, 32, 0, s>>>(ctl, lastRT, delta_tau, weight);
}

}  // namespace gsevt
