// Tile blending, forward and backward.
//
// Forward restates renderCUDA (dgr/cuda_rasterizer/forward.cu:263-392); backward restates renderCUDA
// (dgr/cuda_rasterizer/backward.cu:679-903).  One CTA = one 16x16 tile, one thread = one pixel, the
// tile's depth-sorted instance list is staged through shared memory in batches of 256 (double-buffered:
// one block barrier per batch).
//
// What is different from the reference (results identical):
//  * each warp owns a compact 8x4 pixel block.  32 staged instances at a time, lane l tests instance j0+l
//    against the block — first the bounding box of the ellipse {alpha >= 1/255}, then (engine forward) the
//    exact minimum of the quadratic form over the block's rectangle — and the ballot is the ordered hit
//    list: only hits are evaluated per pixel, list order is preserved;
//  * one 32-byte record per Gaussian {x, y, A, B | C, opacity, gray, depth} = one sector per gather
//    instead of three separate arrays (RGB rides in a second 16-byte record on the operator path);
//  * engine: the forward records, per warp and per group of 32 list positions, which instances some pixel
//    of the warp actually blended; the backward walks exactly those bits (an instance contributes to the
//    backward of a warp iff it contributed to its forward), so it runs no culling test at all;
//  * backward starts at the deepest list position any pixel of the tile blended (max n_contrib), reduces
//    each instance's per-lane terms with a reduce-scatter over the warp (9 shuffles for 8 values) and
//    issues ONE red.add instruction per warp and instance straight into the Gaussian's 32-byte
//    accumulator record — no shared-memory parking, no barrier inside the per-instance loop (the
//    reference has ~12 barriers and a 256-thread tree reduction per Gaussian-tile instance);
//  * the engine variant renders one grayscale channel for BOTH views in one launch (blockIdx.z = view)
//    and its backward derives dL/dpixel on the fly from the normalised event loss (frame.py:86-92,
//    tracker.py:93-103) instead of reading upstream gradient images.
// The per-pixel arithmetic (power, expf, alpha, transmittance test, accumulation) follows the
// reference's rounding order so n_contrib / final_T are bit-identical given identical inputs.
#include "internal.h"
#include "loss_finish.cuh"

namespace gsevt {

namespace {

constexpr float kAlphaMin = 1.0f / 255.0f;

// Conservative bound on the quadratic form: alpha >= 1/255  =>  A dx^2 + 2B dx dy + C dy^2 <= tau.
__device__ __forceinline__ float cull_tau(float o) { return 2.0f * __logf(255.0f * o) * 1.01f + 0.05f; }

// Half extents (plus the warp block's own half size) of the bounding box of {Q(d) <= tau}; +inf when the
// conic is degenerate, -1 when the instance can never reach alpha >= 1/255.
__device__ __forceinline__ float2 cull_extent(float A, float B, float C, float tau) {
    const float det = A * C - B * B;
    if (!(tau > 0.0f)) return make_float2(-1.0f, -1.0f);
    if (!(det > 0.0f) || !(A > 0.0f) || !(C > 0.0f)) return make_float2(3.0e38f, 3.0e38f);
    const float inv = tau / det;
    return make_float2(sqrtf(inv * C) * 1.001f + 3.5f + 0.05f, sqrtf(inv * A) * 1.001f + 1.5f + 0.05f);
}

// Exact test: does the ellipse {Q(d) <= tau} centred at (x, y) meet the pixel rectangle [bx, bx+7] x [by, by+3]?
// Q is convex, so its minimum over the rectangle is 0 when the centre is inside, else it lies on an edge.
__device__ __forceinline__ bool ellipse_hits_block(float x, float y, float A, float B, float C, float tau, float bx, float by) {
    if (!(A > 0.0f) || !(C > 0.0f)) return true;   // degenerate conic: keep (the bounding-box stage already said "hit")
    const float x0 = bx - x, x1 = bx + 7.0f - x, y0 = by - y, y1 = by + 3.0f - y;
    if (x0 <= 0.0f && x1 >= 0.0f && y0 <= 0.0f && y1 >= 0.0f) return true;
    const float rA = 1.0f / A, rC = 1.0f / C;
    float qmin = 3.0e38f;
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const float xe = e ? x1 : x0;                                   // vertical edge dx = xe
        const float ys = fminf(fmaxf(-B * xe * rC, y0), y1);
        qmin = fminf(qmin, A * xe * xe + (2.0f * B * xe + C * ys) * ys);
        const float ye = e ? y1 : y0;                                   // horizontal edge dy = ye
        const float xs = fminf(fmaxf(-B * ye * rA, x0), x1);
        qmin = fminf(qmin, C * ye * ye + (2.0f * B * ye + A * xs) * xs);
    }
    return qmin <= tau;
}

// ---- bulk-async staging of a tile's id list (engine) -------------------------------------------------
// A tile's list is one contiguous run of 32-bit Gaussian indices (bucketbin.cu lays every run out on a 32-byte boundary),
// so a batch of 256 positions is ONE 1-D bulk copy (cp.async.bulk, the TMA unit's non-tensor mode): a single thread arms
// an mbarrier with the byte count and issues the copy, the TMA unit writes the ids into a ring of shared-memory slots and
// completes the barrier's transaction count.  The ids of batch i + 1 are in shared memory long before the CTA stages the
// batch, so the record gather that follows starts from a shared-memory read instead of a dependent global load, and no
// thread spends registers or issue slots on the list itself.  (The 32-byte records stay a gather: they are indexed by
// Gaussian, not by list position — sorted record copies would cost 2 x 32 B x 6.4 M instances of extra HBM traffic per
// iteration, see DESIGN.md "TMA".)
constexpr int kIdSlots = 4;          // ring depth: batches i + 1 .. i + 2 in flight while batch i is blended
constexpr int kIdAhead = 2;          // prefetch distance in batches

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Ring of id batches.  Batch i lives in slot i % kIdSlots and completes phase (i / kIdSlots) & 1 of the slot's barrier.
// Only thread 0 touches the barriers: it arms, issues and waits; the CTA barrier that every staging round ends with
// anyway then publishes the ids to the other 255 threads (thread 0 observed the completed phase before it arrived at
// that barrier, and bar.sync is cumulative), so the consumers pay nothing — the first version, in which every thread
// polled the mbarrier, cost each staging round a 90-cycle try_wait per warp and was 2.4 % slower than plain loads.
struct IdRing {
    uint32_t (*ids)[256];
    uint64_t* bar;
    const uint32_t* list;   // first id of the tile's run (32-byte aligned)
    int issued, waited;     // batches issued / waited for so far (meaningful on thread 0)
    __device__ __forceinline__ void init(uint32_t (*ids_)[256], uint64_t* bar_, const uint32_t* list_) {
        ids = ids_; bar = bar_; list = list_; issued = 0; waited = 0;
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < kIdSlots; k++) mbar_init(bar + k, 1);
            mbar_init_fence();
        }
    }
    // thread 0: positions [first, first + count) of the list -> slot of batch `issued` (count > 0; first is a multiple of 8)
    __device__ __forceinline__ void issue(int first, int count) {
        const uint32_t bytes = (uint32_t)((count + 3) & ~3) * 4u;   // whole 16-byte units: a run's slack covers the overshoot
        uint64_t* b = bar + (issued & (kIdSlots - 1));
        mbar_expect_tx(b, bytes);
        bulk_copy_g2s(ids[issued & (kIdSlots - 1)], list + first, bytes, b);
        issued++;
    }
    // thread 0: waits for the next batch in order
    __device__ __forceinline__ void wait_next() {
        const int i = waited++;
        mbar_wait(bar + (i & (kIdSlots - 1)), (uint32_t)(i / kIdSlots) & 1u);
    }
    __device__ __forceinline__ const uint32_t* slot(int batch) const { return ids[batch & (kIdSlots - 1)]; }
    // thread 0: a CTA must not exit with copies into its shared memory still in flight
    __device__ __forceinline__ void drain() {
        while (waited < issued) wait_next();
    }
};

// Single-instruction approximations for the engine backward (gated at 1e-3, not bit for bit): MUFU.EX2 / MUFU.RCP without the
// range fix-ups of __expf / __fdividef (the arguments here are power <= 0 and 1 - alpha in [0.01, 1]).
__device__ __forceinline__ float fast_exp(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// alpha and its ingredients in the reference's rounding order (forward.cu:342-353).
__device__ __forceinline__ float eval_power(float dx, float dy, float A, float B, float C) {
    const float q = __fmaf_rn(dx, __fmul_rn(dx, A), __fmul_rn(dy, __fmul_rn(dy, C)));
    return __fmaf_rn(q, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, B)));
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Forward
// ------------------------------------------------------------------------------------------------
// Loss evaluation fused into the engine's forward (what loss_stats_kernel does as a launch of its own): a tile is
// rendered by two CTAs, one per view; whichever finishes SECOND (a per-tile arrival counter) has both renders of the tile
// in L2 and sums {d^2, d*E, E^2} over its 256 pixels (d = gray_next - gray_last); the CTA that delivers the LAST tile of the
// launch adds the per-tile sums in tile order (deterministic: no float atomics), runs the tile split's exchange and writes
// the coefficients the backward needs.  Every sum order is fixed by the tile / pixel layout, so repeated evaluations
// and the ranks of a split agree bit for bit.
__device__ __forceinline__ void blend_fwd_loss_epilogue(const BlendFwdArgs& a, int tile, int pixx, int pixy, bool inside) {
    __shared__ double s_w[8][3];
    __shared__ double s_x[8];
    __shared__ int s_role;   // 0: first view of the tile to finish, 1: second, 2: second AND last tile of the launch
    const int HW = a.W * a.H;
    __threadfence();                                   // this CTA's gray pixels are visible device-wide ...
    __syncthreads();
    if (threadIdx.x == 0) s_role = atomicAdd(a.tile_arrive + tile, 1u) == 1u ? 1 : 0;   // ... before its arrival is
    __syncthreads();
    if (s_role == 0) return;
    __threadfence();
    double sd2 = 0.0, s2 = 0.0, se2 = 0.0;
    if (inside) {
        const size_t pix = (size_t)pixy * a.W + pixx;
        // the other view's pixels were written by another SM: read through L2
        const float d = __ldcg(a.out_color + (size_t)HW + pix) - __ldcg(a.out_color + pix);
        const float E = __ldg(a.event_frame + pix);
        sd2 = (double)d * (double)d;
        s2 = a.ctl->loss_signed ? (double)d * (double)E : (double)fabsf(d) * (double)fabsf(E);
        se2 = (double)E * (double)E;
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
    if (threadIdx.x == 0) {
        double t0 = 0, t1 = 0, t2 = 0;
        for (int w = 0; w < 8; w++) { t0 += s_w[w][0]; t1 += s_w[w][1]; t2 += s_w[w][2]; }
        a.loss_partials[3 * tile + 0] = t0;
        a.loss_partials[3 * tile + 1] = t1;
        a.loss_partials[3 * tile + 2] = t2;
        a.tile_arrive[tile] = 0u;                      // ready for the next iteration
        __threadfence();
        if (atomicAdd(a.loss_ticket, 1u) == (unsigned)(a.tile_rows * a.grid_x - 1)) s_role = 2;
    }
    __syncthreads();
    if (s_role != 2 || threadIdx.x >= 32) return;
    __threadfence();
    // last tile delivered: warp 0 adds the strip's tiles in tile order (lane-strided, then a butterfly)
    const int t_first = a.tile_y0 * a.grid_x, t_end = (a.tile_y0 + a.tile_rows) * a.grid_x;
    double t0 = 0, t1 = 0, t2 = 0;
    for (int t = t_first + lane; t < t_end; t += 32) {
        t0 += __ldcg(a.loss_partials + 3 * t + 0);
        t1 += __ldcg(a.loss_partials + 3 * t + 1);
        t2 += __ldcg(a.loss_partials + 3 * t + 2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t0 += __shfl_xor_sync(0xffffffffu, t0, o);
        t1 += __shfl_xor_sync(0xffffffffu, t1, o);
        t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
    if (lane == 0) *a.loss_ticket = 0u;
    loss_finish(t0, t1, t2, a.ctl_rw, a.comm, a.host_flag, a.zero_me, s_x);
}

template <int C, bool OPERATOR, bool BULK = false>
__global__ void __launch_bounds__(256) blend_fwd_kernel(BlendFwdArgs a) {
    static_assert(!(OPERATOR && BULK), "the operator's packed lists are not aligned for bulk copies");
    pdl_prologue();
    if (a.ctl && a.ctl->level_done) return;
    __shared__ __align__(16) uint32_t s_ids[BULK ? kIdSlots : 1][256];
    __shared__ __align__(8) uint64_t s_bar[BULK ? kIdSlots : 1];
    __shared__ float4 s_r0[2][256];
    __shared__ float4 s_r1[2][256];
    __shared__ float4 s_cull[2][256];   // {x, y, half extent x, half extent y} of the alpha >= 1/255 ellipse's box
    __shared__ float s_tau[OPERATOR ? 1 : 2][OPERATOR ? 1 : 256];
    __shared__ float4 s_rgb[OPERATOR ? 2 : 1][OPERATOR ? 256 : 1];
    __shared__ int s_id[OPERATOR ? 2 : 1][OPERATOR ? 256 : 1];

    const int tiles = a.grid_x * a.grid_y;
    const int HW = a.W * a.H;
    // engine: CTA i takes tile tile_order[i] — the tiles of the largest buckets first, so that the tail of the launch is
    // made of short lists; operator: the 3-D grid is the tile grid
    int view, tile_x, tile_y;
    if (a.tile_order) {
        const uint32_t t = a.tile_order[blockIdx.x];
        view = (int)(t / (uint32_t)tiles);
        const int tt = (int)(t - (uint32_t)view * (uint32_t)tiles);
        tile_y = tt / a.grid_x;
        tile_x = tt - tile_y * a.grid_x;
    } else {
        view = blockIdx.z; tile_x = blockIdx.x; tile_y = blockIdx.y + a.tile_y0;
    }
    const int tile_lin = view * tiles + tile_y * a.grid_x + tile_x;
    const uint2 range = a.ranges[tile_lin];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bx = tile_x * GSEVT_TILE + (warp & 1) * 8, by = tile_y * GSEVT_TILE + (warp >> 1) * 4;
    const int pixx = bx + (lane & 7), pixy = by + (lane >> 3);
    const float cxw = (float)bx + 3.5f, cyw = (float)by + 1.5f;
    const float pxf = (float)pixx, pyf = (float)pixy;
    const bool inside = pixx < a.W && pixy < a.H;
    bool done = !inside;

    const float4* __restrict__ rec = a.rec + 2 * (size_t)view * a.view_stride_gauss;
    const int todo = (int)(range.y - range.x);
    const int rounds = (todo + 255) / 256;
    // engine: this warp's row of the hit-mask table; word hit_base[tile] + g holds list positions [32 g, 32 g + 32) of this
    // tile.  hit_base = (range.x >> 5) + (rank of the tile in memory order), written with the ranges by bucket_sort: rows
    // of consecutive tiles cannot overlap (floor(len/32) + 1 >= ceil(len/32))
    uint32_t* __restrict__ hitrow =
        OPERATOR || !a.hitmask ? nullptr : a.hitmask + (size_t)warp * a.hitmask_stride + a.hit_base[tile_lin];

    float T = 1.0f;
    uint32_t last_contributor = 0;
    float acc[C];
#pragma unroll
    for (int ch = 0; ch < C; ch++) acc[ch] = 0.0f;
    float D = 0.0f;

    IdRing ring;
    auto issue_ids = [&](int i) {   // thread 0
        if (i < rounds) ring.issue(i * 256, min(256, todo - i * 256));
    };
    if constexpr (BULK) {
        if (rounds > 0) {
            ring.init(s_ids, s_bar, a.point_list + range.x);
            if (threadIdx.x == 0) {
#pragma unroll
                for (int i = 0; i <= kIdAhead; i++) issue_ids(i);
                ring.wait_next();                    // batch 0
            }
            __syncthreads();                         // barriers initialised, batch 0 visible to everyone
        }
    }
    // double-buffered staging: ONE block barrier per round (it also counts the finished pixels)
    auto stage = [&](int i, int buf) {
        const int prog = i * 256 + threadIdx.x;
        if (prog < todo) {
            const uint32_t id = BULK ? ring.slot(i)[threadIdx.x] : __ldg(a.point_list + range.x + prog);
            const float4 r0 = __ldg(rec + 2 * (size_t)id);
            const float4 r1 = __ldg(rec + 2 * (size_t)id + 1);
            s_r0[buf][threadIdx.x] = r0;
            s_r1[buf][threadIdx.x] = r1;
            const float tau = cull_tau(r1.y);
            const float2 ext = cull_extent(r0.z, r0.w, r1.x, tau);
            s_cull[buf][threadIdx.x] = make_float4(r0.x, r0.y, ext.x, ext.y);
            if constexpr (OPERATOR) {
                s_rgb[buf][threadIdx.x] = __ldg(a.rgb4 + id);
                s_id[buf][threadIdx.x] = (int)id;
            } else {
                s_tau[buf][threadIdx.x] = tau;
            }
        }
    };
    if (rounds > 0) stage(0, 0);
    for (int i = 0; i < rounds; i++) {
        const int buf = i & 1;
        if constexpr (BULK) {
            if (threadIdx.x == 0 && i + 1 < rounds) ring.wait_next();   // batch i + 1, issued two rounds ago; published by the barrier below
        }
        if (__syncthreads_count(done) == 256) break;
        if (i + 1 < rounds) stage(i + 1, buf ^ 1);
        if constexpr (BULK) {
            if (threadIdx.x == 0) issue_ids(i + 1 + kIdAhead);   // its slot held batch i - 1, last read before the barrier above
        }
        const int base = i * 256;
        const int cnt = min(256, todo - base);
        for (int j0 = 0; j0 < cnt; j0 += 32) {
            if (__all_sync(0xffffffffu, done)) break;
            bool hit = false;
            if (j0 + lane < cnt) {
                const float4 c = s_cull[buf][j0 + lane];
                hit = fabsf(c.x - cxw) <= c.z && fabsf(c.y - cyw) <= c.w;
                if constexpr (!OPERATOR) {
                    if (hit) {
                        const float4 q0 = s_r0[buf][j0 + lane];
                        hit = ellipse_hits_block(c.x, c.y, q0.z, q0.w, s_r1[buf][j0 + lane].x, s_tau[buf][j0 + lane], (float)bx, (float)by);
                    }
                }
            }
            unsigned hits = __ballot_sync(0xffffffffu, hit);
            uint32_t mine = 0;      // engine: bit b set <=> THIS pixel blended instance j0 + b (or-reduced over the warp below)
            int last_b = -1;        // last instance of this group this pixel blended
            while (hits) {
                const int b = __ffs(hits) - 1;
                const int j = j0 + b;
                hits &= hits - 1;
                if (!done) {
                    const float4 r0 = s_r0[buf][j];
                    const float dx = __fadd_rn(r0.x, -pxf), dy = __fadd_rn(r0.y, -pyf);
                    const float4 r1 = s_r1[buf][j];
                    const float power = eval_power(dx, dy, r0.z, r0.w, r1.x);
                    if (!(power > 0.0f)) {
                        const float alpha = fminf(0.99f, __fmul_rn(r1.y, expf(power)));
                        if (!(alpha < kAlphaMin)) {
                            const float test_T = __fmul_rn(T, __fadd_rn(1.0f, -alpha));
                            if (test_T < 0.0001f) {
                                done = true;
                            } else {
                                if constexpr (OPERATOR) {
                                    const float4 col = s_rgb[buf][j];
                                    acc[0] = __fmaf_rn(T, __fmul_rn(alpha, col.x), acc[0]);
                                    acc[1] = __fmaf_rn(T, __fmul_rn(alpha, col.y), acc[1]);
                                    acc[2] = __fmaf_rn(T, __fmul_rn(alpha, col.z), acc[2]);
                                    D = __fmaf_rn(T, __fmul_rn(alpha, r1.w), D);
                                    if (a.n_touched && test_T > 0.5f) atomicAdd(a.n_touched + s_id[buf][j], 1);
                                } else {
                                    acc[0] = __fmaf_rn(T, __fmul_rn(alpha, r1.z), acc[0]);
                                }
                                T = test_T;
                                last_b = b;
                                if constexpr (!OPERATOR) mine |= 1u << b;
                            }
                        }
                    }
                }
            }
            if (last_b >= 0) last_contributor = (uint32_t)(base + j0 + last_b + 1);
            if constexpr (!OPERATOR) {
                // which instances of the group some pixel of this warp blended: ONE warp reduction per group of 32
                // positions (a vote per hit cost 3 instructions x 26 hits per group)
                const uint32_t blended = __reduce_or_sync(0xffffffffu, mine);
                if (hitrow && lane == 0) hitrow[(base + j0) >> 5] = blended;
            }
        }
    }

    if constexpr (BULK) {
        if (rounds > 0 && threadIdx.x == 0) ring.drain();
    }
    if (inside) {
        const size_t pix = (size_t)pixy * a.W + pixx;
        a.final_T[(size_t)view * HW + pix] = T;
        a.n_contrib[(size_t)view * HW + pix] = last_contributor;
        if constexpr (OPERATOR) {
#pragma unroll
            for (int ch = 0; ch < C; ch++) a.out_color[(size_t)ch * HW + pix] = __fmaf_rn(__ldg(a.bg + ch), T, acc[ch]);
            a.out_depth[pix] = D;
            a.out_opacity[pix] = __fadd_rn(1.0f, -T);
        } else {
            const ViewParams& vp = a.views[view];
            const float bgg = GSEVT_GRAY_R * vp.bg[0] + GSEVT_GRAY_G * vp.bg[1] + GSEVT_GRAY_B * vp.bg[2];
            a.out_color[(size_t)view * HW + pix] = __fmaf_rn(bgg, T, acc[0]);
        }
    }
    if constexpr (!OPERATOR) {
        if (a.loss_partials) blend_fwd_loss_epilogue(a, tile_y * a.grid_x + tile_x, pixx, pixy, inside);
    }
}

void launch_blend_fwd_rgb(const BlendFwdArgs& a, cudaStream_t s) {
    dim3 grid(a.grid_x, a.tile_rows, a.nviews);
    if (a.tile_rows <= 0) return;
    blend_fwd_kernel<3, true><<<grid, 256, 0, s>>>(a);
}
void launch_blend_fwd_gray(const BlendFwdArgs& a, cudaStream_t s) {
    dim3 grid(a.grid_x, a.tile_rows, a.nviews);
    if (a.tile_rows <= 0) return;
    if (a.tile_order) grid = dim3((unsigned)(a.grid_x * a.tile_rows * a.nviews), 1, 1);
    if (a.bulk_ids) launch_k(blend_fwd_kernel<1, false, true>, grid, dim3(256), 0, s, a);
    else launch_k(blend_fwd_kernel<1, false, false>, grid, dim3(256), 0, s, a);
}

// ------------------------------------------------------------------------------------------------
// Backward
// ------------------------------------------------------------------------------------------------
// Values reduced per instance.  Engine: dmx dmy dA dB dC dgray (6, padded to 8 for the reduction).
// Operator: dmx dmy dA dB dC dop dc0 ddepth | dc1 dc2 (10, padded to 16).
template <bool OPERATOR> struct BwdCfg;
template <> struct BwdCfg<false> { static constexpr int K = 6, KR = 8; };
template <> struct BwdCfg<true> { static constexpr int K = 10, KR = 16; };

// Sum of each of N (power of two <= 32) per-lane values over the warp.  Returns, on every lane, the total of
// component lane / (32 / N).  Halving exchange: at each step a lane keeps one half of its values and trades
// the other half with its partner, so N values cost N - 1 + log2(32 / N) shuffles instead of 5 N.
template <int N>
__device__ __forceinline__ float warp_reduce_scatter(float (&v)[N], int lane) {
    int off = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1) {
        const int half = n >> 1;
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; i++) {
            const float send = hi ? v[i] : v[half + i];
            const float keep = hi ? v[half + i] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
        off >>= 1;
    }
    float r = v[0];
#pragma unroll
    for (; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
    return r;
}

template <int C, bool OPERATOR, bool BULK = false>
__global__ void __launch_bounds__(256) blend_bwd_kernel(BlendBwdArgs a) {
    static_assert(!(OPERATOR && BULK), "the operator's packed lists are not aligned for bulk copies");
    pdl_prologue();
    if (a.ctl && a.ctl->level_done) return;
    constexpr int K = BwdCfg<OPERATOR>::K, KR = BwdCfg<OPERATOR>::KR;
    __shared__ __align__(16) uint32_t s_ids[BULK ? kIdSlots : 1][256];
    __shared__ __align__(8) uint64_t s_bar[BULK ? kIdSlots : 1];
    __shared__ float4 s_r0[2][256];
    __shared__ float4 s_r1[2][256];
    __shared__ float4 s_cull[OPERATOR ? 2 : 1][OPERATOR ? 256 : 1];
    __shared__ float4 s_rgb[OPERATOR ? 2 : 1][OPERATOR ? 256 : 1];
    __shared__ uint32_t s_id[2][256];
    __shared__ uint32_t s_max[8];

    const int tiles = a.grid_x * a.grid_y;
    const int HW = a.W * a.H;
    // engine: CTA i takes tile tile_order[i] — the tiles of the largest buckets first, so that the tail of the launch is
    // made of short lists; operator: the 3-D grid is the tile grid
    int view, tile_x, tile_y;
    if (a.tile_order) {
        const uint32_t t = a.tile_order[blockIdx.x];
        view = (int)(t / (uint32_t)tiles);
        const int tt = (int)(t - (uint32_t)view * (uint32_t)tiles);
        tile_y = tt / a.grid_x;
        tile_x = tt - tile_y * a.grid_x;
    } else {
        view = blockIdx.z; tile_x = blockIdx.x; tile_y = blockIdx.y + a.tile_y0;
    }
    const int tile_lin = view * tiles + tile_y * a.grid_x + tile_x;
    const uint2 range = a.ranges[tile_lin];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bx = tile_x * GSEVT_TILE + (warp & 1) * 8, by = tile_y * GSEVT_TILE + (warp >> 1) * 4;
    const int pixx = bx + (lane & 7), pixy = by + (lane >> 3);
    const float cxw = (float)bx + 3.5f, cyw = (float)by + 1.5f;
    const float pxf = (float)pixx, pyf = (float)pixy;
    const bool inside = pixx < a.W && pixy < a.H;
    const size_t pix = (size_t)pixy * a.W + pixx;

    const float T_final = inside ? a.final_T[(size_t)view * HW + pix] : 0.0f;
    const uint32_t last_contributor = inside ? a.n_contrib[(size_t)view * HW + pix] : 0u;

    float dpix[C];
    float dpix_depth = 0.0f;
    float bg_dot = 0.0f;
#pragma unroll
    for (int ch = 0; ch < C; ch++) dpix[ch] = 0.0f;
    if (inside) {
        if constexpr (OPERATOR) {
#pragma unroll
            for (int ch = 0; ch < C; ch++) {
                dpix[ch] = a.dL_dpix[(size_t)ch * HW + pix];
                bg_dot += __ldg(a.bg + ch) * dpix[ch];
            }
            if (a.dL_dpix_depth) dpix_depth = a.dL_dpix_depth[pix];
        } else {
            // d = gray_next - gray_last; dL/dd = alpha*d - beta*E_eff (see DESIGN.md "Loss")
            const float d = a.gray[(size_t)HW + pix] - a.gray[pix];
            const float E = a.event_frame[pix];
            const float Eeff = a.ctl->loss_signed ? E : (d > 0.0f ? fabsf(E) : (d < 0.0f ? -fabsf(E) : 0.0f));
            const float g = a.ctl->loss_alpha * d - a.ctl->loss_beta * Eeff;
            dpix[0] = view == 1 ? g : -g;
            const ViewParams& vp = a.views[view];
            bg_dot = (GSEVT_GRAY_R * vp.bg[0] + GSEVT_GRAY_G * vp.bg[1] + GSEVT_GRAY_B * vp.bg[2]) * dpix[0];
        }
    }

    // Deepest list position any pixel of this warp / this tile blended.
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, last_contributor);
    if (lane == 0) s_max[warp] = wmax;
    __syncthreads();
    uint32_t max_lc = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) max_lc = max(max_lc, s_max[w]);
    if (max_lc == 0) return;

    const float4* __restrict__ rec = a.rec + 2 * (size_t)view * a.view_stride_gauss;
    float* __restrict__ grad_f = reinterpret_cast<float*>(a.grad8 + 2 * (size_t)view * a.view_stride_gauss);
    const float ddelx_dx = 0.5f * a.W, ddely_dy = 0.5f * a.H;
    const uint32_t* __restrict__ hitrow =
        OPERATOR ? nullptr : a.hitmask + (size_t)warp * a.hitmask_stride + a.hit_base[tile_lin];

    float T = T_final;
    float accum_rec[C], last_color[C];
#pragma unroll
    for (int ch = 0; ch < C; ch++) accum_rec[ch] = last_color[ch] = 0.0f;
    float accum_rec_depth = 0.0f, last_depth = 0.0f, last_alpha = 0.0f;

    // which component this lane owns after the reduce-scatter, and whether it adds it to memory
    constexpr int LPC = 32 / KR;                       // lanes per component
    const int comp = lane / LPC;
    const bool owner = (lane % LPC) == 0 && comp < K;

    // Batches are aligned to 256 list positions (so the forward's 32-position mask words never straddle one):
    // batch b holds positions [top - 256 b - 255, top - 256 b]; staged entry t <-> position hi - t.
    const int top = (int)((max_lc - 1) | 255u);
    const int len = (int)(range.y - range.x);
    const int nbatches = (top + 1) / 256;
    IdRing ring;
    auto issue_ids = [&](int b) {   // thread 0; batch b = positions [lo, lo + 256) ∩ [0, len), lo = top - 256 b - 255 < max_lc <= len
        if (b < nbatches) {
            const int lo = top - b * 256 - 255;
            ring.issue(lo, min(256, len - lo));
        }
    };
    if constexpr (BULK) {
        ring.init(s_ids, s_bar, a.point_list + range.x);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int b = 0; b <= kIdAhead; b++) issue_ids(b);
            ring.wait_next();                        // batch 0
        }
        __syncthreads();
    }
    auto stage = [&](int b, int buf) {
        const int pos = top - b * 256 - (int)threadIdx.x;
        if (pos < len) {   // pos >= 0 always: top - 256 b - 255 >= 0 for b < nbatches
            const uint32_t id = BULK ? ring.slot(b)[255 - (int)threadIdx.x] : __ldg(a.point_list + range.x + (uint32_t)pos);
            const float4 r0 = __ldg(rec + 2 * (size_t)id);
            const float4 r1 = __ldg(rec + 2 * (size_t)id + 1);
            s_r0[buf][threadIdx.x] = r0;
            s_r1[buf][threadIdx.x] = r1;
            s_id[buf][threadIdx.x] = id;
            if constexpr (OPERATOR) {
                const float2 ext = cull_extent(r0.z, r0.w, r1.x, cull_tau(r1.y));
                s_cull[buf][threadIdx.x] = make_float4(r0.x, r0.y, ext.x, ext.y);
                s_rgb[buf][threadIdx.x] = __ldg(a.rgb4 + id);
            }
        }
    };

    // one instance, all pixels of the warp: recompute alpha, rebuild T and the colour behind, reduce, add
    auto process = [&](int buf, int j, uint32_t pos) {
        const float4 r0 = s_r0[buf][j];
        const float4 r1 = s_r1[buf][j];
        const float dx = __fadd_rn(r0.x, -pxf), dy = __fadd_rn(r0.y, -pyf);
        const float power = eval_power(dx, dy, r0.z, r0.w, r1.x);
        // Operator: the reference's arithmetic (expf, IEEE reciprocal).  Engine: the backward is gated at 1e-3, not bit for
        // bit, so exp is ex2.approx (2 instructions instead of 11) and 1 / (1 - alpha) is rcp.approx (instead of 14) — except
        // within 1e-6 of the alpha >= 1/255 test, which must fall as it did in the forward pass: a pair blended there and
        // skipped here would leave T off by (1 - alpha) for everything in front of it.
        float G = OPERATOR ? expf(power) : fast_exp(power);
        float alpha = fminf(0.99f, __fmul_rn(r1.y, G));
        if constexpr (!OPERATOR) {
            if (fabsf(alpha - kAlphaMin) < 1e-6f) {
                G = expf(power);
                alpha = fminf(0.99f, __fmul_rn(r1.y, G));
            }
        }
        const bool skip = pos >= last_contributor || power > 0.0f || alpha < kAlphaMin;   // !inside => last_contributor == 0
        if constexpr (OPERATOR) {
            if (__all_sync(0xffffffffu, skip)) return;
        }
        float v[KR];
#pragma unroll
        for (int k = 0; k < KR; k++) v[k] = 0.0f;
        if (!skip) {
            const float inv = OPERATOR ? __frcp_rn(1.0f - alpha) : fast_rcp(1.0f - alpha);
            T = T * inv;
            const float w = alpha * T;
            float dL_dalpha = 0.0f;
            if constexpr (OPERATOR) {
                const float4 col = s_rgb[buf][j];
                const float c3[3] = {col.x, col.y, col.z};
#pragma unroll
                for (int ch = 0; ch < C; ch++) {
                    accum_rec[ch] = last_alpha * last_color[ch] + (1.0f - last_alpha) * accum_rec[ch];
                    last_color[ch] = c3[ch];
                    dL_dalpha += (c3[ch] - accum_rec[ch]) * dpix[ch];
                }
                v[6] = w * dpix[0];
                v[8] = w * dpix[1];
                v[9] = w * dpix[2];
                accum_rec_depth = last_alpha * last_depth + (1.0f - last_alpha) * accum_rec_depth;
                last_depth = r1.w;
                dL_dalpha += (r1.w - accum_rec_depth) * dpix_depth;
                v[7] = w * dpix_depth;
            } else {
                accum_rec[0] = last_alpha * last_color[0] + (1.0f - last_alpha) * accum_rec[0];
                last_color[0] = r1.z;
                dL_dalpha += (r1.z - accum_rec[0]) * dpix[0];
                v[5] = w * dpix[0];
            }
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final * inv) * bg_dot;
            const float dL_dG = r1.y * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float dG_ddelx = -gdx * r0.z - gdy * r0.w;
            const float dG_ddely = -gdy * r1.x - gdx * r0.w;
            v[0] = dL_dG * dG_ddelx * ddelx_dx;
            v[1] = dL_dG * dG_ddely * ddely_dy;
            v[2] = -0.5f * gdx * dx * dL_dG;
            v[3] = -0.5f * gdx * dy * dL_dG;
            v[4] = -0.5f * gdy * dy * dL_dG;
            if constexpr (OPERATOR) v[5] = G * dL_dalpha;
        }
        const float mine = warp_reduce_scatter<KR>(v, lane);
        if (owner) {
            const uint32_t id = s_id[buf][j];
            if constexpr (OPERATOR) {
                float* dst = comp < 8 ? grad_f + 8 * (size_t)id + comp : reinterpret_cast<float*>(a.gradc + id) + (comp - 8);
                atomicAdd(dst, mine);
            } else {
                atomicAdd(grad_f + 8 * (size_t)id + comp, mine);
            }
        }
    };

    stage(0, 0);
    for (int b = 0; b < nbatches; b++) {
        const int buf = b & 1;
        if constexpr (BULK) {
            if (threadIdx.x == 0 && b + 1 < nbatches) ring.wait_next();   // ids of batch b + 1; published by the barrier below
        }
        __syncthreads();                                   // batch b staged; everyone is done with buffer buf^1
        if (b + 1 < nbatches) stage(b + 1, buf ^ 1);
        if constexpr (BULK) {
            if (threadIdx.x == 0) issue_ids(b + 1 + kIdAhead);   // its slot held batch b - 1, last read before the barrier above
        }
        const int hi = top - b * 256;                      // list position of staged entry 0
        if ((uint32_t)(hi - 255) >= wmax) continue;        // nothing of this batch can touch this warp's pixels
        if constexpr (OPERATOR) {
            const int jskip = max(0, hi + 1 - (int)wmax);  // first staged entry with position < wmax
            for (int j0 = jskip & ~31; j0 < 256; j0 += 32) {
                bool hit = false;
                const int jl = j0 + lane;
                if (jl >= jskip && hi - jl < len) {
                    const float4 c = s_cull[buf][jl];
                    hit = fabsf(c.x - cxw) <= c.z && fabsf(c.y - cyw) <= c.w;
                }
                unsigned hits = __ballot_sync(0xffffffffu, hit);
                while (hits) {
                    const int j = j0 + __ffs(hits) - 1;
                    hits &= hits - 1;
                    process(buf, j, (uint32_t)(hi - j));
                }
            }
        } else {
            // the forward's record of what this warp blended: 8 words cover this batch's 256 positions
            const uint32_t g0 = (uint32_t)(hi - 255) >> 5;
            const uint32_t gmax = (wmax - 1) >> 5;          // words above were never written for this warp
            uint32_t mw = 0;
            if (lane < 8 && g0 + lane <= gmax) mw = __ldg(hitrow + g0 + lane);
            for (int k = 7; k >= 0; k--) {
                uint32_t m = __shfl_sync(0xffffffffu, mw, k);
                while (m) {
                    const int bit = 31 - __clz(m);
                    m &= ~(1u << bit);
                    const uint32_t pos = ((g0 + k) << 5) + bit;
                    process(buf, hi - (int)pos, pos);
                }
            }
        }
    }
}

void launch_blend_bwd_rgb(const BlendBwdArgs& a, cudaStream_t s) {
    dim3 grid(a.grid_x, a.tile_rows, a.nviews);
    if (a.tile_rows <= 0) return;
    blend_bwd_kernel<3, true><<<grid, 256, 0, s>>>(a);
}
void launch_blend_bwd_gray(const BlendBwdArgs& a, cudaStream_t s) {
    dim3 grid(a.grid_x, a.tile_rows, a.nviews);
    if (a.tile_rows <= 0) return;
    if (a.tile_order) grid = dim3((unsigned)(a.grid_x * a.tile_rows * a.nviews), 1, 1);
    if (a.bulk_ids) launch_k(blend_bwd_kernel<1, false, true>, grid, dim3(256), 0, s, a);
    else launch_k(blend_bwd_kernel<1, false, false>, grid, dim3(256), 0, s, a);
}

}  // namespace gsevt
