// The end of the loss evaluation, shared by loss_stats_kernel (loss_events.cu) and the epilogue of the engine's blend
// forward (blend.cu): the screen-tile split's exchange of the three sums and the closed-form coefficients of
// dL/d(delta) = alpha * d - beta * E_eff (frame.py:86-92, tracker.py:93-103).
#pragma once
#include "internal.h"
#include "split_comm.cuh"

namespace gsevt {

// Called by ONE full warp with {sum d^2, sum d*E (or |d||E|), sum E^2} over this engine's pixels on every lane.
// s_x: 8 doubles of shared memory.
__device__ __forceinline__ void loss_finish(double t0, double t1, double t2, EngineCtl* ctl, SplitComm* comm, int* host_flag,
                                            uint32_t* zero_me, double* s_x) {
    const int lane = threadIdx.x & 31;
    if (comm) {
        // screen-tile split: these are the sums over this rank's strip; exchange them over peer memory
        if (lane == 0) { s_x[0] = t0; s_x[1] = t1; s_x[2] = t2; }
        __syncwarp();
        const bool ok = split_exchange(comm, 0, 3, s_x, s_x + 4);
        __syncwarp();
        t0 = s_x[4]; t1 = s_x[5]; t2 = s_x[6];
        if (!ok && lane == 0) {
            ctl->comm_error = 1;
            ctl->level_done = 3;
            if (host_flag) *host_flag = 3;
            __threadfence_system();
        }
    }
    if (lane != 0) return;
    if (zero_me) *zero_me = 0u;   // per-iteration counter of the backward's work list (consumed two kernels later)
    const double n = sqrt(t0);
    double L2 = 1.0 - 2.0 * t1 / n + t2;
    if (L2 < 0.0) L2 = 0.0;
    const double L = sqrt(L2);
    if (n > 0.0 && L > 0.0) {
        ctl->loss_alpha = (float)(t1 / (n * n * n * L));
        ctl->loss_beta = (float)(1.0 / (L * n));
        ctl->last_loss = (float)L;
    } else {
        ctl->loss_alpha = 0.0f;
        ctl->loss_beta = 0.0f;
        ctl->last_loss = n > 0.0 ? (float)L : (float)sqrt(t2);
    }
}

}  // namespace gsevt
