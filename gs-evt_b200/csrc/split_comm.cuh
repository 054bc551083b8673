// Screen-tile split (BASELINE configs[4]): the two tiny cross-GPU reductions of an iteration
// — {sum d^2, sum d*E, sum E^2} after the forward (the loss normalises over the WHOLE image,
// frame.py:90-91) and the 12-float pose/velocity gradient (+ the overflow flag) after the backward — done
// INSIDE the kernels that produce them, by stores into the peers' HBM over NVLink, so the iteration stays
// one CUDA-graph launch with no host round trip and no separate collective launch.  Messages are <= 112 B:
// pure latency, which is why an all-gather-by-remote-store + local ordered sum beats a ring / NCCL launch.
#pragma once
#include "internal.h"

namespace gsevt {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(double* p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Called by ONE full warp.  `mine[0..nvals)` must be readable by every lane (shared or global memory).
// On return total[k] (k < nvals, written by lane k into `total`, shared or global) holds the sum over ranks
// in rank order — bit-identical on every rank.  Returns false on every lane if a peer timed out.
__device__ __forceinline__ bool split_exchange(SplitComm* c, int channel, int nvals, const double* mine, double* total) {
    const int lane = threadIdx.x & 31;
    const unsigned long long seq = c->seq[channel] + 1ull;
    const int parity = (int)(seq & 1ull);
    if (lane < c->n) {
        MailSlot* dst = &c->box[lane]->slot[channel][parity][c->rank];
        for (int k = 0; k < nvals; k++) st_relaxed_sys(&dst->v[k], mine[k]);
        __threadfence_system();
        st_release_sys(&dst->seq, seq);
    }
    bool ok = true;
    const MailBox* my = c->box[c->rank];
    if (lane < c->n) {
        const unsigned long long* flag = &my->slot[channel][parity][lane].seq;
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(flag) != seq) {
            if (global_timer_ns() - t0 > c->timeout_ns) { ok = false; break; }
            __nanosleep(20);
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    __syncwarp();
    if (lane < nvals) {
        double t = 0.0;
        for (int r = 0; r < c->n; r++) t += ld_relaxed_sys(&my->slot[channel][parity][r].v[lane]);
        total[lane] = t;
    }
    __syncwarp();
    if (lane == 0) c->seq[channel] = seq;
    return ok;
}

}  // namespace gsevt
