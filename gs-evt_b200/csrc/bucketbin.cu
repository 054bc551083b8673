// Binning of the engine path: from the projected (view, Gaussian) pairs straight to the per-tile, depth-ordered lists.
//
// What it replaces.  The reference writes one (tile << 32 | depth, id) record per tile instance and radix-sorts all of
// them with six 8-bit passes (rasterizer_impl.cu:70-111, 303-311), then finds the tile boundaries in the sorted keys
// (:116-138).  Inside a tile the order is (depth bits, Gaussian index) — the second by the stability of the sort over
// the emission order.  Because (depth bits, index) is a TOTAL order on the Gaussians of one view, nothing has to be
// stable here: any procedure that ends with every tile's covering Gaussians in ascending (depth bits << 32 | index)
// reproduces the reference's lists bit for bit.  Two kernels:
//
//   bucket_scatter   one thread per (view, Gaussian) pair, read in index order (rect_raw / depth_raw of the
//                    projection: no compaction pass).  A bucket is a (1 << s) x (1 << s) block of tiles (s = 1 at the
//                    fine levels, 0 at the coarse ones); a visible pair takes one slot in every bucket its tile rect
//                    touches — 2.5 buckets on average at s = 1 against 4.7 tiles — with ONE global atomic and writes its
//                    64-bit key there: depth bits << 32 | Gaussian index << 4 | which of the bucket's 2 x 2 tiles the rect
//                    covers.  A bucket's segment is cut into GSEVT_BK_SUB sub-segments with a cursor each (256 B apart:
//                    one L2 atomic unit each), picked by the CTA index: the L2 serialises atomics on one address, and
//                    5 400 of them per cursor were 55 of this kernel's 64 us [measured r2].  Slot order is whatever the
//                    atomics hand out.  Segments have a fixed capacity, sized when the level begins (count-only run of
//                    this kernel + slack); a sub-segment that outgrows its share sets the overflow flag, the iteration
//                    is voided on the device and the host re-sizes.  (bucket_scatter_list: the same scatter over the
//                    list of visible pairs the split projection kernels write — screen-tile split, where 80-90 % of a
//                    rank's pairs are empty.  The per-pair code lives in bucket_scatter.cuh.)
//   bucket_sort      one CTA per bucket: counting sort of the bucket's keys on the depth in shared memory with one
//                    shared-memory atomic per key + ranking inside the depth bins (below), then the sorted bucket is
//                    FILTERED into its tiles — a tile's list is the subsequence of the bucket whose rect covers the
//                    tile, so the filter is a ballot + prefix per warp, in order — and the tile ranges are written.
//                    Buckets with an adversarial depth distribution or larger than the shared-memory budget take the
//                    merge sort of sortcore.cuh (8-key network per thread, pairwise merge-path rounds, in shared or
//                    global memory).
//
// Tile lists are not packed back to back: tile k of bucket b owns vals[(start[b] << 2s) + k * cap[b], + cap[b]), so no
// count pass over the tiles is needed before the lists are written; ranges[] holds (begin, begin + count) into vals.
// gsevt_engine_binning() re-bases them to the reference's packed representation for the parity tests.
#include "internal.h"
#include "sortcore.cuh"
#include "bucket_scatter.cuh"

namespace gsevt {

using namespace sortcore;

template <bool COUNT_ONLY, int S>
__global__ void __launch_bounds__(256) bucket_scatter_kernel(BucketArgs a) {
    pdl_prologue();
    if (a.ctl && a.ctl->level_done) return;
    const uint32_t j = blockIdx.x * 256u + threadIdx.x;            // pair id = view * P + Gaussian
    const uint32_t n2 = 2u * (uint32_t)a.P;
    const uint32_t rect = j < n2 ? __ldg(a.rect_raw + j) : 0u;
    const uint32_t depth = !COUNT_ONLY && j < n2 ? __ldg(a.depth_raw + j) : 0u;   // independent of the rect: both loads in flight together
    const uint32_t view = j >= (uint32_t)a.P ? 1u : 0u;
    const uint32_t id = j - view * (uint32_t)a.P;
    const uint32_t sub = blockIdx.x & (GSEVT_BK_SUB - 1);          // this CTA's sub-segment of every bucket
    uint32_t* const cur0 = a.cursor + (size_t)sub * GSEVT_BK_CURSOR_STRIDE;   // cursor of (bucket, sub) = cur0 + bucket * CSTEP
    bscatter::Pair p;
    bscatter::prepare<S>(a, view, rect, p);
    // the common case (a rect of 2 x 2 tiles meets 1..4 buckets): all atomics in flight before the first store
    bscatter::issue_small(a, cur0, p);
    if constexpr (!COUNT_ONLY) bscatter::finish_small<S>(a, sub, p, depth, id);
    bscatter::big_rects<COUNT_ONLY, S>(a, sub, p, depth, id);
}

// Screen-tile split: 80-90 % of a rank's pairs are empty, so the scatter walks the list of visible pairs the split
// projection kernel wrote (preprocess.cu) instead of the dense rect array.  Grid-stride over the device-side count.
template <int S>
__global__ void __launch_bounds__(256) bucket_scatter_list_kernel(BucketArgs a) {
    pdl_prologue();
    if (a.ctl && a.ctl->level_done) return;
    const uint32_t n = *a.vis_count;
    for (uint32_t base = blockIdx.x * 256u; base < n; base += gridDim.x * 256u) {   // uniform per CTA: whole warps stay for big_rects
        const uint32_t t = base + threadIdx.x;
        const uint32_t j = t < n ? __ldg(a.vis_list + t) : 0u;
        // the sub-segment the dense kernel would have used for this pair (its CTA index there): the segments were sized
        // from a dense count-only run, so every (bucket, sub-segment) receives exactly the pairs it was sized for
        const uint32_t sub = (j >> 8) & (GSEVT_BK_SUB - 1);
        uint32_t* const cur0 = a.cursor + (size_t)sub * GSEVT_BK_CURSOR_STRIDE;
        const uint32_t rect = t < n ? __ldg(a.rect_raw + j) : 0u;
        const uint32_t depth = t < n ? __ldg(a.depth_raw + j) : 0u;
        const uint32_t view = j >= (uint32_t)a.P ? 1u : 0u;
        const uint32_t id = j - view * (uint32_t)a.P;
        bscatter::Pair p;
        bscatter::prepare<S>(a, view, rect, p);
        bscatter::issue_small(a, cur0, p);
        bscatter::finish_small<S>(a, sub, p, depth, id);
        bscatter::big_rects<false, S>(a, sub, p, depth, id);
    }
}

// ---- per-bucket sort + tile lists -------------------------------------------------------------------
namespace {

// Sorts np = roundup(n, VT) keys.  SMEM: the keys are loaded from `in` into A first; otherwise A == in (sorted in
// place).  A and B are two buffers of np keys; returns the one that holds the result.  All threads of the CTA call it.
template <bool SMEM>
__device__ __forceinline__ uint64_t* cta_merge_sort(const uint64_t* in, uint64_t* A, uint64_t* B, int n, int np) {
    const int T = (int)blockDim.x;
    for (int o = (int)threadIdx.x * VT; o < np; o += T * VT) {
        uint64_t k[VT];
        if (o + VT <= n) {
            const ulonglong2* p = reinterpret_cast<const ulonglong2*>(in + o);   // segments start on 64-byte boundaries
#pragma unroll
            for (int i = 0; i < VT / 2; i++) {
                const ulonglong2 q = p[i];   // (plain loads: the fallback path's input was written by this CTA)
                k[2 * i] = q.x;
                k[2 * i + 1] = q.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < VT; i++) k[i] = o + i < n ? in[o + i] : PAD;
        }
        sort8(k);
#pragma unroll
        for (int i = 0; i < VT; i++) A[o + i] = k[i];
    }
    __syncthreads();
    uint64_t *src = A, *dst = B;
    const int nseg = np / VT;
    for (int run = VT; run < np; run <<= 1) {
        for (int seg = (int)threadIdx.x; seg < nseg; seg += T) merge_segment(src, dst, np, run, seg);
        __syncthreads();
        uint64_t* t = src; src = dst; dst = t;
    }
    return src;
}

}  // namespace

// S = log2 of the bucket edge in tiles (0: bucket == tile, 1: 2 x 2 tiles).
//
// Fast path — a counting sort on the depth with ONE shared-memory atomic per key:
//   1. min / max of the bucket's depths; bin = monotone linear map of the depth VALUE onto NB bins (NB >= keys; the float
//      bits would give bins whose population doubles with every octave);
//   2. r = atomicAdd(count[bin], 1): the histogram AND the key's arrival index inside its bin;
//   3. exclusive scan of the counts; key -> sK[start[bin] + r]: the bucket is now ordered by bin;
//   4. every key ranks itself among the keys of its bin (1.x on average) by counting the smaller 64-bit keys -> its
//      final position.  Keys are unique, so this is a total order whatever the arrival order was.
// The keys are re-read from the bucket's global segment (coalesced, four loads in flight per thread) in each step instead
// of being held in registers, so that two CTAs per SM overlap each other's barriers.
// The merge sort (sortcore.cuh: no assumption about the depth distribution) takes over when a bin holds more than
// BIN_HEAVY keys — ranking by counting is quadratic in the bin — or the bucket exceeds the shared-memory budget.
constexpr int BIN_HEAVY = 64;
constexpr int SORT_THREADS = 512;

template <int S>
__global__ void __launch_bounds__(SORT_THREADS, 2) bucket_sort_kernel(BucketArgs a) {
    pdl_prologue();
    if (a.ctl && a.ctl->level_done) return;
    extern __shared__ __align__(16) uint64_t s_buf[];
    constexpr int NT = 1 << (2 * S);                    // tiles per bucket
    constexpr int T = SORT_THREADS, NW = T / 32, R = GSEVT_BK_SUB;
    __shared__ uint32_t s_nj[R], s_off[R + 1], s_dmin, s_heavy;
    __shared__ float s_scale;
    __shared__ uint32_t s_red[2][NW];
    __shared__ uint32_t s_wc[NW][NT];
    __shared__ uint32_t s_tot[NT];
    if (a.vis_count && blockIdx.x == 0 && threadIdx.x == 0) {   // split mode: both lists were consumed earlier in this iteration
        *a.vis_count = 0u;
        *a.surv_count = 0u;
    }
    const uint32_t b = a.bk_order ? a.bk_order[blockIdx.x] : blockIdx.x;   // largest buckets first
    const uint32_t start = a.bk_start[b], cap = a.bk_cap[b], subcap = cap / R;
    if (threadIdx.x < R) {
        uint32_t* cur = a.cursor + ((size_t)b * R + threadIdx.x) * GSEVT_BK_CURSOR_STRIDE;
        const uint32_t raw = *cur;
        *cur = 0u;                                      // the next iteration's scatter starts from zero
        s_nj[threadIdx.x] = raw < subcap ? raw : subcap;   // raw > subcap: overflow, flagged by the scatter; the iteration is void
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int j = 0; j < R; j++) { s_off[j] = run; run += s_nj[j]; }
        s_off[R] = run;
    }
    __syncthreads();
    const int n = (int)s_off[R];
    const uint32_t view = b >= (uint32_t)a.nb ? 1u : 0u;
    const uint32_t bl = b - view * (uint32_t)a.nb;
    const int by = (int)(bl / (uint32_t)a.nbx), bx = (int)(bl - (uint32_t)by * (uint32_t)a.nbx);
    const int tx0 = bx << S, ty0 = (by + a.by_origin) << S;
    uint32_t* const vbase = a.vals + ((size_t)start << (2 * S));   // this bucket's tiles own vbase[k * cap, + cap)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t* seg = a.keys + start;

    // f(slot, key) for every key of the bucket.  Slot = sub * subcap + arrival index in the sub-segment; thread t owns slots
    // t, t + T, t + 2 T, ... (at most 32: cap <= 32 T on this path) and keeps their validity in one register, so a pass
    // over the keys is coalesced 8-byte loads, four in flight per thread before the first key is used.
    uint32_t vmask = 0;
    {
        const float rcp_sub = 1.0f / (float)subcap;
#pragma unroll 4
        for (int u = 0; u < 32; u++) {
            const uint32_t sl = threadIdx.x + (uint32_t)u * T;
            if (sl < cap) {
                const uint32_t jj = (uint32_t)(((float)sl + 0.5f) * rcp_sub);   // exact: sl < 2^23
                if (sl - jj * subcap < s_nj[jj]) vmask |= 1u << u;
            }
        }
    }
    auto for_keys = [&](auto f) {
        for (uint32_t m = vmask, u0 = 0; m; m >>= 4, u0 += 4) {
            uint64_t k[4];
#pragma unroll
            for (int u = 0; u < 4; u++) k[u] = (m >> u & 1u) ? __ldg(seg + threadIdx.x + (u0 + u) * T) : 0ull;
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (m >> u & 1u) f(threadIdx.x + (u0 + u) * T, k[u]);
        }
    };

    // Stable filter of the sorted bucket into its tiles (S > 0): warp w owns the contiguous chunk [w * cw, w * cw + cw) of
    // the sorted sequence; pass 1 counts per (warp, tile) — one warp reduction of the covered-tile bits spread over 16-bit
    // fields per two tiles — pass 2 writes at the warp's base + ballot prefix.
    // get(p) -> Gaussian index | cover mask << 28 of the p-th key in sorted order.
    auto emit = [&](auto get) {
        const int cw = ((n + NW - 1) / NW + 31) & ~31;
        const int p0 = warp * cw, p1 = min(n, p0 + cw);
        uint32_t lo = 0, hi = 0;
        for (int q = p0; q < p1; q += 32) {
            const int p = q + lane;
            const uint32_t m = p < p1 ? get(p) >> 28 : 0u;
            lo += __reduce_add_sync(0xffffffffu, (m & 1u) | ((m & 2u) << 15));
            if constexpr (NT > 2) hi += __reduce_add_sync(0xffffffffu, ((m >> 2) & 1u) | ((m & 8u) << 13));
        }
        if (lane == 0) {
            s_wc[warp][0] = lo & 0xFFFFu;
            if constexpr (NT > 1) s_wc[warp][1] = lo >> 16;
            if constexpr (NT > 2) { s_wc[warp][2] = hi & 0xFFFFu; s_wc[warp][3] = hi >> 16; }
        }
        __syncthreads();
        if (threadIdx.x < NT) {
            uint32_t run = 0;
            for (int w = 0; w < NW; w++) {
                const uint32_t t = s_wc[w][threadIdx.x];
                s_wc[w][threadIdx.x] = run;
                run += t;
            }
            s_tot[threadIdx.x] = run;
        }
        __syncthreads();
        uint32_t c[NT];
#pragma unroll
        for (int k = 0; k < NT; k++) c[k] = (uint32_t)k * cap + s_wc[warp][k];
        const uint32_t lt = (1u << lane) - 1u;
        for (int q = p0; q < p1; q += 32) {
            const int p = q + lane;
            const uint32_t v = p < p1 ? get(p) : 0u;
            const uint32_t m = v >> 28, id = v & 0x0FFFFFFFu;
#pragma unroll
            for (int k = 0; k < NT; k++) {
                const unsigned bal = __ballot_sync(0xffffffffu, (m >> k) & 1u);
                if ((m >> k) & 1u) vbase[c[k] + (uint32_t)__popc(bal & lt)] = id;
                c[k] += (uint32_t)__popc(bal);
            }
        }
    };

    uint32_t my_count = 0;                              // thread k < NT: instances of tile k
    if (n > 0) {
        const int E = a.smem_elems, NB = a.smem_bins;
        uint64_t* sK = s_buf;                           // [E] keys in bin order
        uint32_t* cnt = reinterpret_cast<uint32_t*>(s_buf + E);   // [NB] counts, then bin starts, then (S > 0) id | mask in sorted order
        uint16_t* sR = reinterpret_cast<uint16_t*>(cnt + NB);     // [E] per slot: arrival index in the bin, then final position
        bool sorted_fast = false;
        if ((int)cap <= E) {
            for (int i = threadIdx.x; i < NB; i += T) cnt[i] = 0u;
            uint32_t dmin = 0xFFFFFFFFu, dmax = 0u;
            for_keys([&](uint32_t, uint64_t key) {
                const uint32_t d = (uint32_t)(key >> 32);
                dmin = min(dmin, d);
                dmax = max(dmax, d);
            });
            dmin = __reduce_min_sync(0xffffffffu, dmin);
            dmax = __reduce_max_sync(0xffffffffu, dmax);
            if (lane == 0) { s_red[0][warp] = dmin; s_red[1][warp] = dmax; }
            __syncthreads();
            if (warp == 0) {
                dmin = __reduce_min_sync(0xffffffffu, lane < NW ? s_red[0][lane] : 0xFFFFFFFFu);
                dmax = __reduce_max_sync(0xffffffffu, lane < NW ? s_red[1][lane] : 0u);
                if (lane == 0) {
                    // depths are positive floats: the order of the bits is the order of the values
                    const float zmin = __uint_as_float(dmin), zmax = __uint_as_float(dmax);
                    s_dmin = dmin;
                    s_scale = zmax > zmin ? (float)NB / (zmax - zmin) : 0.0f;
                    s_heavy = 0u;
                }
            }
            __syncthreads();
            const float z0 = __uint_as_float(s_dmin), scale = s_scale;
            const uint32_t last_bin = (uint32_t)NB - 1u;
            // monotone in the depth: subtraction of a constant, multiplication by a positive constant and truncation all are
            auto bin_of = [&](uint64_t key) -> uint32_t {
                return min((uint32_t)((__uint_as_float((uint32_t)(key >> 32)) - z0) * scale), last_bin);
            };
            for_keys([&](uint32_t sl, uint64_t key) { sR[sl] = (uint16_t)atomicAdd(&cnt[bin_of(key)], 1u); });
            __syncthreads();
            {
                // exclusive scan of the NB counts, CPT consecutive counts per thread; largest count -> s_heavy
                const int CPT = NB / T;                 // NB is a multiple of 4 T
                uint32_t* mine = cnt + threadIdx.x * CPT;
                uint32_t sum = 0, mx = 0;
                for (int k = 0; k < CPT; k += 4) {
                    const uint4 q = *reinterpret_cast<const uint4*>(mine + k);
                    sum += q.x + q.y + q.z + q.w;
                    mx = max(max(mx, max(q.x, q.y)), max(q.z, q.w));
                }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += y;
                }
                mx = __reduce_max_sync(0xffffffffu, mx);
                if (lane == 31) s_red[0][warp] = incl;
                if (lane == 0 && mx > (uint32_t)BIN_HEAVY) s_heavy = 1u;
                __syncthreads();
                uint32_t run = incl - sum;
                for (int w = 0; w < warp; w++) run += s_red[0][w];
                for (int k = 0; k < CPT; k += 4) {
                    uint4 q = *reinterpret_cast<const uint4*>(mine + k);
                    const uint32_t c0 = q.x, c1 = q.y, c2 = q.z;
                    q.x = run; q.y = run + c0; q.z = q.y + c1; run = q.z + c2 + q.w; q.w = q.z + c2;
                    *reinterpret_cast<uint4*>(mine + k) = q;
                }
            }
            __syncthreads();
            if (!s_heavy) {
                sorted_fast = true;
                for_keys([&](uint32_t sl, uint64_t key) { sK[cnt[bin_of(key)] + sR[sl]] = key; });
                __syncthreads();
                for_keys([&](uint32_t sl, uint64_t key) {
                    const uint32_t bn = bin_of(key);
                    const uint32_t s0 = cnt[bn], e0 = bn < last_bin ? cnt[bn + 1u] : (uint32_t)n;
                    uint32_t rank = s0;
#pragma unroll 1   // 1.x keys per bin: the unrolled loop's prologue cost more than it saved
                    for (uint32_t j = s0; j < e0; j++) rank += sK[j] < key ? 1u : 0u;
                    if constexpr (S == 0) vbase[rank] = (uint32_t)key >> 4;
                    else sR[sl] = (uint16_t)rank;
                });
                if constexpr (S == 0) {
                    my_count = (uint32_t)n;
                } else {
                    __syncthreads();                    // the bin starts are dead: their words take id | mask in sorted order
                    for_keys([&](uint32_t sl, uint64_t key) { cnt[sR[sl]] = __funnelshift_r((uint32_t)key, (uint32_t)key, 4); });
                    __syncthreads();
                    emit([&](int p) { return cnt[p]; });
                    if (threadIdx.x < NT) my_count = s_tot[threadIdx.x];
                }
            }
        }
        if (!sorted_fast) {
            // pack the sub-segments back to back in the second key buffer, then the merge sort: in shared memory while two
            // buffers of the bucket fit, else in place over the two global buffers
            __syncthreads();
            uint64_t* packed = a.keys2 + start;
            for (uint32_t sl = threadIdx.x; sl < cap; sl += T) {
                const uint32_t jj = sl / subcap, l = sl - jj * subcap;
                if (l < s_nj[jj]) packed[s_off[jj] + l] = seg[sl];
            }
            __syncthreads();
            const int np = (n + VT - 1) / VT * VT;      // <= cap (a multiple of VT)
            const int half = (int)(a.smem_bytes / 16);
            const uint64_t* sorted = np <= half ? cta_merge_sort<true>(packed, s_buf, s_buf + half, n, np)
                                                : cta_merge_sort<false>(packed, packed, a.keys + start, n, np);
            if constexpr (S == 0) {
                for (int p = threadIdx.x; p < n; p += T) vbase[p] = (uint32_t)sorted[p] >> 4;
                my_count = (uint32_t)n;
            } else {
                emit([&](int p) { const uint32_t w = (uint32_t)sorted[p]; return __funnelshift_r(w, w, 4); });
                if (threadIdx.x < NT) my_count = s_tot[threadIdx.x];
            }
        }
    }
    // ranges of this bucket's tiles (tiles outside the grid exist in edge buckets only on paper)
    if (threadIdx.x < NT) {
        const int k = threadIdx.x;
        const int tx = tx0 + (k & ((1 << S) - 1)), ty = ty0 + (k >> S);
        if (tx < a.gx && ty < a.gy) {
            const uint32_t tile = view * (uint32_t)a.tiles_global + (uint32_t)ty * (uint32_t)a.gx + (uint32_t)tx;
            const uint32_t beg = (start << (2 * S)) + (uint32_t)k * cap;
            a.ranges[tile] = make_uint2(beg, beg + my_count);
            // word of the forward -> backward hit-mask rows where this tile's list starts: tiles are laid out in
            // (bucket, k) order, one spare word per tile keeps the rows of consecutive tiles apart (blend.cu)
            a.hit_base[tile] = (beg >> 5) + (b << (2 * S)) + (uint32_t)k;
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------
void launch_bucket_scatter(const BucketArgs& a, bool count_only, cudaStream_t s) {
    if (a.P <= 0 || a.nb <= 0) return;
    const unsigned blocks = (unsigned)((2 * (size_t)a.P + 255) / 256);
    if (count_only) {
        if (a.s == 0) bucket_scatter_kernel<true, 0><<<blocks, 256, 0, s>>>(a);
        else bucket_scatter_kernel<true, 1><<<blocks, 256, 0, s>>>(a);
    } else if (a.sparse) {
        const unsigned lb = blocks < 148u * 8u ? blocks : 148u * 8u;   // 8 CTAs per SM, grid-stride over the visible pairs
        if (a.s == 0) launch_k(bucket_scatter_list_kernel<0>, dim3(lb), dim3(256), 0, s, a);
        else launch_k(bucket_scatter_list_kernel<1>, dim3(lb), dim3(256), 0, s, a);
    } else {
        if (a.s == 0) launch_k(bucket_scatter_kernel<false, 0>, dim3(blocks), dim3(256), 0, s, a);
        else launch_k(bucket_scatter_kernel<false, 1>, dim3(blocks), dim3(256), 0, s, a);
    }
}

int bucket_sort_configure() {
    cudaError_t e = cudaFuncSetAttribute(bucket_sort_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GSEVT_BK_SMEM_MAX_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bucket_sort_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GSEVT_BK_SMEM_MAX_BYTES);
    return e == cudaSuccess ? 0 : -1;
}

// Shared memory of one sort CTA for buckets of up to `elems` keys: keys in bin order (8 B), one count per bin (4 B,
// bins = elems rounded up to a multiple of 4 x 512 so that every thread scans whole 16-byte groups), arrival index (2 B).
void bucket_sort_smem(int max_keys, int* elems, int* bins, size_t* bytes) {
    int e = max_keys < 8 ? 8 : (max_keys + 7) / 8 * 8;
    if (e > GSEVT_BK_SMEM_MAX_ELEMS) e = GSEVT_BK_SMEM_MAX_ELEMS;
    const int nb = (e + 4 * SORT_THREADS - 1) / (4 * SORT_THREADS) * (4 * SORT_THREADS);
    *elems = e; *bins = nb; *bytes = (size_t)e * 10 + (size_t)nb * 4;
}

void launch_bucket_sort(const BucketArgs& a, cudaStream_t s) {
    if (a.nb <= 0) return;
    if (a.s == 0) launch_k(bucket_sort_kernel<0>, dim3(2 * a.nb), dim3(SORT_THREADS), a.smem_bytes, s, a);
    else launch_k(bucket_sort_kernel<1>, dim3(2 * a.nb), dim3(SORT_THREADS), a.smem_bytes, s, a);
}

// padded cursor array -> packed counts, one per (bucket, sub-segment) (probe)
__global__ void bucket_counts_kernel(int n, const uint32_t* __restrict__ cursor, uint32_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = cursor[(size_t)i * GSEVT_BK_CURSOR_STRIDE];
}
void launch_bucket_counts(int n, const uint32_t* cursor, uint32_t* out, cudaStream_t s) {
    if (n > 0) bucket_counts_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, cursor, out);
}

// Screen-tile split: tile instances per tile row (both views), the cost model the strips are balanced on.
__global__ void __launch_bounds__(256) row_histogram_kernel(int n, const uint32_t* __restrict__ rect_raw, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const uint32_t r = __ldg(rect_raw + i);
        const uint32_t w = (r >> 16 & 255u) - (r & 255u);
        if (w == 0) continue;
        for (uint32_t y = r >> 8 & 255u; y < r >> 24; y++) atomicAdd(&s_h[y], w);
    }
    __syncthreads();
    if (s_h[threadIdx.x]) atomicAdd(hist + threadIdx.x, s_h[threadIdx.x]);
}
void launch_row_histogram(int n_pairs, const uint32_t* rect_raw, uint32_t* hist256, cudaStream_t s) {
    cudaMemsetAsync(hist256, 0, 256 * sizeof(uint32_t), s);
    if (n_pairs <= 0) return;
    int blocks = (n_pairs + 255) / 256;
    if (blocks > 1184) blocks = 1184;   // 8 CTAs per SM
    row_histogram_kernel<<<blocks, 256, 0, s>>>(n_pairs, rect_raw, hist256);
}

// Parity-test helper: the per-tile lists of one view packed back to back in tile order with the reference's 64-bit
// keys rebuilt (tile << 32 | depth bits), one CTA per tile.  packed_start[t] = first slot of tile t in the output.
__global__ void export_lists_kernel(int tiles, const uint2* __restrict__ ranges_view, const uint32_t* __restrict__ vals,
                                    const float4* __restrict__ rec_view, const uint32_t* __restrict__ packed_start,
                                    uint64_t* __restrict__ keys_out, uint32_t* __restrict__ list_out) {
    const int t = blockIdx.x;
    if (t >= tiles) return;
    const uint2 r = ranges_view[t];
    const uint32_t o = packed_start[t];
    for (uint32_t i = threadIdx.x; i < r.y - r.x; i += blockDim.x) {
        const uint32_t id = vals[r.x + i];
        keys_out[o + i] = ((uint64_t)(uint32_t)t << 32) | __float_as_uint(rec_view[2 * (size_t)id + 1].w);
        list_out[o + i] = id;
    }
}
void launch_export_lists(int tiles, const uint2* ranges_view, const uint32_t* vals, const float4* rec_view, const uint32_t* packed_start,
                         uint64_t* keys_out, uint32_t* list_out, cudaStream_t s) {
    if (tiles > 0) export_lists_kernel<<<tiles, 128, 0, s>>>(tiles, ranges_view, vals, rec_view, packed_start, keys_out, list_out);
}

}  // namespace gsevt
