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
//                    touches — 2.5 buckets on average at s = 1 against 4.7 tiles — with ONE global atomic on the
//                    bucket's cursor (cursors 256 B apart: one L2 atomic unit each) and writes its 64-bit key there.
//                    Slot order is whatever the atomics hand out.  Bucket segments have a fixed capacity, sized when
//                    the level begins (count-only run of this kernel + slack); a bucket that outgrows its segment
//                    sets the overflow flag, the iteration is voided on the device and the host re-sizes.
//   bucket_sort      one CTA per bucket: merge sort of the bucket's keys in shared memory (sortcore.cuh: 8-key
//                    network per thread, then pairwise merge-path rounds; buckets larger than the shared-memory
//                    budget run the same rounds over their global segments), then the sorted bucket is FILTERED into
//                    its tiles — a tile's list is the subsequence of the bucket whose rect covers the tile, so the
//                    filter is a ballot + prefix per warp, in order — and the tile ranges are written.  No atomics:
//                    a shared-memory atomic on 32 different addresses costs 2 cycles per lane on this GPU, which is
//                    what bounded the previous counting kernels (tile_count / tile_scatter, r1).
//
// Tile lists are not packed back to back: tile k of bucket b owns vals[(start[b] << 2s) + k * cap[b], + cap[b]), so no
// count pass over the tiles is needed before the lists are written; ranges[] holds (begin, begin + count) into vals.
// gsevt_engine_binning() re-bases them to the reference's packed representation for the parity tests.
#include "internal.h"
#include "sortcore.cuh"

namespace gsevt {

using namespace sortcore;

namespace {

template <bool COUNT_ONLY>
__device__ __forceinline__ void take_slot(const BucketArgs& a, uint32_t b, uint64_t key) {
    const uint32_t slot = atomicAdd(a.cursor + (size_t)b * GSEVT_BK_CURSOR_STRIDE, 1u);
    if constexpr (!COUNT_ONLY) {
        if (slot < __ldg(a.bk_cap + b)) a.keys[(size_t)__ldg(a.bk_start + b) + slot] = key;
        else *a.overflow = 1;
    }
}

}  // namespace

template <bool COUNT_ONLY>
__global__ void __launch_bounds__(256) bucket_scatter_kernel(BucketArgs a) {
    if (a.ctl && a.ctl->level_done) return;
    const uint32_t j = blockIdx.x * 256u + threadIdx.x;            // pair id = view * P + Gaussian
    const uint32_t n2 = 2u * (uint32_t)a.P;
    const uint32_t rect = j < n2 ? __ldg(a.rect_raw + j) : 0u;
    const uint32_t view = j >= (uint32_t)a.P ? 1u : 0u;
    const uint32_t id = j - view * (uint32_t)a.P;
    int bx0 = 0, bx1 = 0, by0 = 0, by1 = 0;
    if (rect) bucket_rect(rect, a.s, a.by_origin, bx0, bx1, by0, by1);
    const uint32_t w = (uint32_t)(bx1 - bx0), cnt = w * (uint32_t)(by1 - by0);
    uint64_t key = 0;
    if (!COUNT_ONLY && rect) key = ((uint64_t)__ldg(a.depth_raw + j) << 32) | id;
    const uint32_t bbase = view * (uint32_t)a.nb + (uint32_t)by0 * (uint32_t)a.nbx + (uint32_t)bx0;
    constexpr uint32_t SMALL = 4;
    if (cnt && cnt <= SMALL) {
        // the common case (a rect of 2 x 2 tiles meets 1..4 buckets): all atomics in flight before the first store
        uint32_t bb[SMALL], slot[SMALL];
#pragma unroll
        for (uint32_t t = 0; t < SMALL; t++) {
            const uint32_t ty = (t >= w ? 1u : 0u) + (t >= 2u * w ? 1u : 0u) + (t >= 3u * w ? 1u : 0u);
            bb[t] = bbase + ty * (uint32_t)a.nbx + (t - ty * w);
            slot[t] = t < cnt ? atomicAdd(a.cursor + (size_t)bb[t] * GSEVT_BK_CURSOR_STRIDE, 1u) : 0u;
        }
        if constexpr (!COUNT_ONLY) {
#pragma unroll
            for (uint32_t t = 0; t < SMALL; t++) {
                if (t < cnt) {
                    if (slot[t] < __ldg(a.bk_cap + bb[t])) a.keys[(size_t)__ldg(a.bk_start + bb[t]) + slot[t]] = key;
                    else *a.overflow = 1;
                }
            }
        }
    }
    // large rects: the whole warp walks one pair's buckets, 32 per step, so that no lane loops over a screen-filling
    // Gaussian alone
    unsigned bigs = __ballot_sync(0xffffffffu, cnt > SMALL);
    const uint32_t lane = threadIdx.x & 31u;
    while (bigs) {
        const int src = __ffs(bigs) - 1;
        bigs &= bigs - 1;
        const uint32_t b_cnt = __shfl_sync(0xffffffffu, cnt, src), b_w = __shfl_sync(0xffffffffu, w, src);
        const uint32_t b_base = __shfl_sync(0xffffffffu, bbase, src);
        const uint32_t k_lo = __shfl_sync(0xffffffffu, (uint32_t)key, src), k_hi = __shfl_sync(0xffffffffu, (uint32_t)(key >> 32), src);
        const uint64_t b_key = ((uint64_t)k_hi << 32) | k_lo;
        for (uint32_t t = lane; t < b_cnt; t += 32u) {
            const uint32_t ty = t / b_w;
            take_slot<COUNT_ONLY>(a, b_base + ty * (uint32_t)a.nbx + (t - ty * b_w), b_key);
        }
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
                const ulonglong2 q = SMEM ? __ldg(p + i) : p[i];
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
template <int S>
__global__ void __launch_bounds__(1024, 1) bucket_sort_kernel(BucketArgs a) {
    if (a.ctl && a.ctl->level_done) return;
    extern __shared__ __align__(16) uint64_t s_buf[];   // [2][smem_elems]
    constexpr int NT = 1 << (2 * S);                    // tiles per bucket
    __shared__ uint32_t s_n;
    __shared__ uint32_t s_wc[32][NT];
    __shared__ uint32_t s_tot[NT];
    const uint32_t b = blockIdx.x;
    if (threadIdx.x == 0) {
        uint32_t* cur = a.cursor + (size_t)b * GSEVT_BK_CURSOR_STRIDE;
        const uint32_t raw = *cur, cap = a.bk_cap[b];
        *cur = 0u;                                      // the next iteration's scatter starts from zero
        s_n = raw < cap ? raw : cap;                    // raw > cap: overflow, flagged by the scatter; the iteration is void
    }
    __syncthreads();
    const int n = (int)s_n;
    const uint32_t start = a.bk_start[b], cap = a.bk_cap[b];
    const uint32_t view = b >= (uint32_t)a.nb ? 1u : 0u;
    const uint32_t bl = b - view * (uint32_t)a.nb;
    const int by = (int)(bl / (uint32_t)a.nbx), bx = (int)(bl - (uint32_t)by * (uint32_t)a.nbx);
    const int tx0 = bx << S, ty0 = (by + a.by_origin) << S;
    const size_t voff = ((size_t)start << (2 * S));     // this bucket's tiles own vals[voff + k * cap, + cap)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (int)blockDim.x >> 5;

    uint32_t my_count = 0;                              // thread k < NT: instances of tile k
    if (n > 0) {
        const int np = (n + VT - 1) / VT * VT;          // <= cap (a multiple of VT)
        uint64_t* seg = a.keys + start;
        const bool in_smem = np <= a.smem_elems;
        const uint64_t* sorted;
        uint32_t* rects = nullptr;                      // S > 0, shared-memory path: tile rect of every sorted key
        if (in_smem) {
            uint64_t* r = cta_merge_sort<true>(seg, s_buf, s_buf + a.smem_elems, n, np);
            sorted = r;
            if constexpr (S > 0) {
                rects = reinterpret_cast<uint32_t*>(r == s_buf ? s_buf + a.smem_elems : s_buf);
                const uint32_t* rr = a.rect_raw + (size_t)view * a.P;
                for (int p = threadIdx.x; p < n; p += blockDim.x) rects[p] = __ldg(rr + (uint32_t)r[p]);
                __syncthreads();
            }
        } else {
            sorted = cta_merge_sort<false>(seg, seg, a.keys2 + start, n, np);
        }
        if constexpr (S == 0) {
            for (int p = threadIdx.x; p < n; p += blockDim.x) a.vals[voff + p] = (uint32_t)sorted[p];
            my_count = (uint32_t)n;
        } else {
            // Stable filter into the bucket's tiles: warp w owns the contiguous chunk [w * cw, w * cw + cw) of the
            // sorted keys; pass 1 counts per (warp, tile), pass 2 writes at the warp's base + ballot prefix.
            const uint32_t* rr = a.rect_raw + (size_t)view * a.P;
            const int cw = ((n + nwarps - 1) / nwarps + 31) & ~31;
            const int p0 = warp * cw, p1 = min(n, p0 + cw);
            uint32_t c[NT];
#pragma unroll
            for (int k = 0; k < NT; k++) c[k] = 0;
            for (int q = p0; q < p1; q += 32) {
                const int p = q + lane;
                uint32_t m = 0;
                if (p < p1) m = cover_mask4(rects ? rects[p] : __ldg(rr + (uint32_t)sorted[p]), tx0, ty0);
#pragma unroll
                for (int k = 0; k < NT; k++) c[k] += (uint32_t)__popc(__ballot_sync(0xffffffffu, (m >> k) & 1u));
            }
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < NT; k++) s_wc[warp][k] = c[k];
            }
            __syncthreads();
            if (threadIdx.x < NT) {
                uint32_t run = 0;
                for (int w = 0; w < nwarps; w++) {
                    const uint32_t t = s_wc[w][threadIdx.x];
                    s_wc[w][threadIdx.x] = run;
                    run += t;
                }
                s_tot[threadIdx.x] = run;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < NT; k++) c[k] = s_wc[warp][k];
            const uint32_t lt = (1u << lane) - 1u;
            for (int q = p0; q < p1; q += 32) {
                const int p = q + lane;
                uint32_t m = 0, id = 0;
                if (p < p1) {
                    id = (uint32_t)sorted[p];
                    m = cover_mask4(rects ? rects[p] : __ldg(rr + id), tx0, ty0);
                }
#pragma unroll
                for (int k = 0; k < NT; k++) {
                    const unsigned bal = __ballot_sync(0xffffffffu, (m >> k) & 1u);
                    if ((m >> k) & 1u) a.vals[voff + (size_t)k * cap + c[k] + (uint32_t)__popc(bal & lt)] = id;
                    c[k] += (uint32_t)__popc(bal);
                }
            }
            if (threadIdx.x < NT) my_count = s_tot[threadIdx.x];
        }
    }
    // ranges of this bucket's tiles (tiles outside the grid exist in edge buckets only on paper)
    if (threadIdx.x < NT) {
        const int k = threadIdx.x;
        const int tx = tx0 + (k & ((1 << S) - 1)), ty = ty0 + (k >> S);
        if (tx < a.gx && ty < a.gy) {
            const uint32_t tile = view * (uint32_t)a.tiles_global + (uint32_t)ty * (uint32_t)a.gx + (uint32_t)tx;
            const uint32_t beg = (uint32_t)(voff + (size_t)k * cap);
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
    if (count_only) bucket_scatter_kernel<true><<<blocks, 256, 0, s>>>(a);
    else bucket_scatter_kernel<false><<<blocks, 256, 0, s>>>(a);
}

int bucket_sort_configure() {
    cudaError_t e = cudaFuncSetAttribute(bucket_sort_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GSEVT_BK_SMEM_MAX_ELEMS * 16);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bucket_sort_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GSEVT_BK_SMEM_MAX_ELEMS * 16);
    return e == cudaSuccess ? 0 : -1;
}

void launch_bucket_sort(const BucketArgs& a, cudaStream_t s) {
    if (a.nb <= 0) return;
    const size_t smem = (size_t)a.smem_elems * 16;
    // two CTAs of 512 threads per SM while two buffers of the largest bucket fit twice, else one CTA of 1024
    const int threads = smem > 110 * 1024 ? 1024 : 512;
    if (a.s == 0) bucket_sort_kernel<0><<<2 * a.nb, threads, smem, s>>>(a);
    else bucket_sort_kernel<1><<<2 * a.nb, threads, smem, s>>>(a);
}

// cursors[i] of the padded cursor array -> packed counts (probe)
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
