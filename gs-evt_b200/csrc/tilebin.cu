// Tile binning of the engine path: from the depth-sorted (view, Gaussian) pairs straight to the per-tile lists.
//
// What it replaces.  The reference writes one (tile << 32 | depth, id) record per tile instance and radix-sorts
// all of them (rasterizer_impl.cu:70-111, 303-311), then finds the tile boundaries in the sorted keys (:116-138).
// The order inside a tile is (depth bits, Gaussian index).  The engine already holds the pairs in exactly that
// order (stable depth sort of the index-ordered compact list; view 0 in front of view 1 because the sort key carries
// the view in bit 31), so the per-tile lists are a STABLE PARTITION of the instance sequence by tile id — a one-pass
// counting sort over the tiles of one view — instead of a sort of 12-byte records:
//
//   tile_count    chunk = 4096 consecutive instances of ONE view, generated from the pairs' tile rects; counts per
//                 (chunk, tile) in shared memory -> hist[chunk][tile] (u16); (u16 tile, u32 id) per instance staged
//                 in shared memory and written out coalesced
//   tile_scan     per tile: exclusive prefix of the counts over the view's chunks -> base[chunk][tile]; totals per
//                 tile; the LAST CTA to finish prefixes the totals over the tiles -> ranges[tile] (untouched tiles
//                 stay (0,0) as in the reference)
//   tile_scatter  chunk again (coalesced read of its instances): every instance gets slot ranges[tile].x +
//                 base[chunk][tile] + (its rank among the chunk's earlier instances of the same tile) and stores
//                 its Gaussian index there.
//
// Stability inside a chunk: the chunk is cut into 8 slices in sequence order, one per warp; a first pass counts per
// (slice, tile), the prefix over the slices gives every warp its own cursor per tile, and each warp then walks its
// slice in order, 32 instances per step (shared-memory atomics of one warp execute in program order); equal tiles
// inside a step are ranked by lane with match.any, which a stamp test skips when the 32 tiles are all different.
// Traffic: the pairs once (12 B each), hist/base once each way, per instance 6 B out + 6 B in + 4 B out — against
// (6 B + 2 x 12 B + 2 B) per instance for emit + two radix passes + range scan (the fallback above 2048 tiles per view).
#include <cstdlib>
#include "internal.h"

namespace gsevt {

namespace {

constexpr int TB_THREADS = 256;
constexpr int TB_IPT = 16;                        // instances per thread
constexpr int TB_CH = TB_THREADS * TB_IPT;        // instances per chunk (4096)
constexpr int TB_SLICE = TB_CH / 8;               // instances per warp slice

__device__ __forceinline__ uint32_t rect_area_d(uint32_t r) {
    return ((r >> 16 & 255u) - (r & 255u)) * ((r >> 24) - (r >> 8 & 255u));
}

// Chunk b of the instance sequence.  The depth-sorted pairs hold view 0 first (the sort key carries the view in bit 31),
// so instances [0, N0) belong to view 0 and [N0, total) to view 1; chunks never straddle the boundary: rows [0, R0) of
// hist / base cut view 0 into pieces of TB_CH instances, rows [R0, R0 + R1) view 1.  A chunk therefore bins one view's
// tiles only — half the shared memory and half the per-chunk bookkeeping of binning both views at once.
struct ChunkRange { uint32_t begin, end, view, rows0, rows1; bool active, overflow; };
__device__ __forceinline__ ChunkRange chunk_range(const uint32_t* __restrict__ off, const uint32_t* __restrict__ n_vis, int n_pairs,
                                                  int cap, uint32_t b) {
    ChunkRange c;
    const uint32_t total = __ldg(off + n_pairs - 1);
    const uint32_t nv0 = __ldg(n_vis + 1);
    const uint32_t N0 = nv0 ? __ldg(off + nv0 - 1) : 0u;
    c.overflow = total > (uint32_t)cap;
    c.rows0 = (N0 + TB_CH - 1) / TB_CH;
    c.rows1 = (total - N0 + TB_CH - 1) / TB_CH;
    if (b < c.rows0) {
        c.view = 0; c.begin = b * (uint32_t)TB_CH; c.end = min(c.begin + (uint32_t)TB_CH, N0);
        c.active = true;
    } else {
        c.view = 1; c.begin = N0 + (b - c.rows0) * (uint32_t)TB_CH; c.end = min(c.begin + (uint32_t)TB_CH, total);
        c.active = c.begin < total;
    }
    if (c.overflow) c.active = false;
    return c;
}

// First index p in [0, n) with off[p] > o, for a non-decreasing off[] with off[n-1] > o.  All 32 lanes call it;
// every round probes 32 positions, so the depth is log32(n) dependent loads instead of log2(n).
__device__ __forceinline__ uint32_t upper_bound_warp(const uint32_t* __restrict__ off, uint32_t n, uint32_t o) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t lo = 0, len = n;
    while (len > 1) {
        const uint32_t step = (len + 31u) / 32u;
        const uint32_t last = lo + len - 1u;
        uint32_t idx = lo + (lane + 1u) * step - 1u;
        if (idx > last) idx = last;
        const unsigned b = __ballot_sync(0xffffffffu, __ldg(off + idx) > o);   // monotone in the lane; lane 31 is true
        const uint32_t L = (uint32_t)__ffs(b) - 1u;
        const uint32_t nlo = lo + L * step;
        uint32_t nhi = nlo + step - 1u;
        if (nhi > last) nhi = last;
        lo = nlo;
        len = nhi - nlo + 1u;
    }
    return lo;
}

// Pair range [p_lo, p_hi] whose instances intersect [begin, end), into shared memory (warps 0 and 1 search).
__device__ __forceinline__ void chunk_pair_range(const uint32_t* __restrict__ off, uint32_t n, uint32_t begin, uint32_t end,
                                                 uint32_t* s_p) {
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        const uint32_t p = upper_bound_warp(off, n, begin);
        if ((threadIdx.x & 31) == 0) s_p[0] = p;
    } else if (warp == 1) {
        const uint32_t p = upper_bound_warp(off, n, end - 1u);
        if ((threadIdx.x & 31) == 0) s_p[1] = p;
    }
}

// One pair's share of a chunk: instances t in [t0, t1) of its rect (row-major), the first at local index excl - begin.
struct PairWork { uint32_t excl, t0, t1, tbase, wd, id; };

__device__ __forceinline__ PairWork load_pair(const uint64_t* __restrict__ pairs, const uint32_t* __restrict__ off, uint32_t p,
                                              uint32_t p_hi, uint32_t begin, uint32_t end, int P, int gx, int tiles_per_view, int row0) {
    PairWork w = {0u, 0u, 0u, 0u, 1u, 0u};
    if (p <= p_hi) {
        const uint64_t pr = __ldg(pairs + p);
        const uint32_t rect = (uint32_t)(pr >> 32);
        const uint32_t incl = __ldg(off + p);
        w.id = (uint32_t)pr >= (uint32_t)P ? (uint32_t)pr - (uint32_t)P : (uint32_t)pr;
        w.excl = incl - rect_area_d(rect);
        w.t0 = (w.excl > begin ? w.excl : begin) - w.excl;
        w.t1 = (incl < end ? incl : end) - w.excl;
        if (w.t1 < w.t0) w.t1 = w.t0;
        w.wd = (rect >> 16 & 255u) - (rect & 255u);
        w.tbase = ((rect >> 8 & 255u) - (uint32_t)row0) * (uint32_t)gx + (rect & 255u);   // bin = tile of the strip (one view per chunk)
    }
    return w;
}

// Visits every instance of the warp's 32 pairs: f(local_instance, tile, Gaussian).  A lane walks its own rect when it
// has at most TB_SMALL instances in the chunk; larger ones are walked by the whole warp, 32 instances per step, so no
// lane ever loops over a screen-filling Gaussian alone.
__constant__ uint32_t TB_SMALL = 32;   // set once from GSEVT_TB_SMALL (tuning knob)
template <typename F>
__device__ __forceinline__ void visit_pair(const PairWork& w, uint32_t begin, int gx, F f) {
    const bool big = w.t1 - w.t0 > TB_SMALL;
    if (!big && w.t1 > w.t0) {
        uint32_t ty = w.t0 / w.wd, tx = w.t0 - ty * w.wd;
        for (uint32_t t = w.t0; t < w.t1; t++) {
            f(w.excl + t - begin, w.tbase + ty * (uint32_t)gx + tx, w.id);
            if (++tx == w.wd) { tx = 0; ty++; }
        }
    }
    unsigned bigs = __ballot_sync(0xffffffffu, big);
    const uint32_t lane = threadIdx.x & 31u;
    while (bigs) {
        const int src = __ffs(bigs) - 1;
        bigs &= bigs - 1;
        const uint32_t b_t0 = __shfl_sync(0xffffffffu, w.t0, src), b_t1 = __shfl_sync(0xffffffffu, w.t1, src);
        const uint32_t b_excl = __shfl_sync(0xffffffffu, w.excl, src), b_wd = __shfl_sync(0xffffffffu, w.wd, src);
        const uint32_t b_tbase = __shfl_sync(0xffffffffu, w.tbase, src), b_id = __shfl_sync(0xffffffffu, w.id, src);
        const float rcpw = 1.0f / (float)b_wd;
        for (uint32_t t = b_t0 + lane; t < b_t1; t += 32u) {
            const uint32_t ty = (uint32_t)(((float)t + 0.5f) * rcpw);   // exact: t < 65536, wd <= 255
            const uint32_t tx = t - ty * b_wd;
            f(b_excl + t - begin, b_tbase + ty * (uint32_t)gx + tx, b_id);
        }
    }
}

// All pairs [p_lo, p_hi] of a chunk, 256 per round.  (Keeping several rounds' loads in flight costs more in registers and
// code size than the latency it hides: measured 25 -> 47 us for the count kernel.)
template <typename F>
__device__ __forceinline__ void visit_chunk(const uint64_t* __restrict__ pairs, const uint32_t* __restrict__ off, uint32_t p_lo,
                                            uint32_t p_hi, uint32_t begin, uint32_t end, int P, int gx, int tiles_per_view, int row0, F f) {
    for (uint32_t p0 = p_lo; p0 <= p_hi; p0 += TB_THREADS) {
        const PairWork w = load_pair(pairs, off, p0 + threadIdx.x, p_hi, begin, end, P, gx, tiles_per_view, row0);
        visit_pair(w, begin, gx, f);
    }
}

}  // namespace

// ---- 1. counts per (chunk, tile) ------------------------------------------------------------------
__global__ void __launch_bounds__(TB_THREADS) tile_count_kernel(int P, int n_pairs, int gx, int tiles_per_view, int row0,
                                                                 const uint64_t* __restrict__ pairs,
                                                                 const uint32_t* __restrict__ off, const uint32_t* __restrict__ n_vis,
                                                                 uint16_t* __restrict__ hist, uint16_t* __restrict__ inst_tile, uint32_t* __restrict__ inst_id,
                                                                 int cap, int* __restrict__ overflow,
                                                                 const EngineCtl* __restrict__ ctl) {
    if (ctl && ctl->level_done) return;
    extern __shared__ uint32_t s_hist[];   // [tiles_per_view]
    __shared__ uint32_t s_p[2];
    __shared__ uint16_t s_tile[TB_CH];     // the chunk's instances in sequence order, staged so that they leave coalesced
    __shared__ uint32_t s_id[TB_CH];
    const ChunkRange cr = chunk_range(off, n_vis, n_pairs, cap, blockIdx.x);
    if (cr.overflow && blockIdx.x == 0 && threadIdx.x == 0) *overflow = 1;
    if (!cr.active) return;
    const uint32_t begin = cr.begin, end = cr.end;
    const int nt = tiles_per_view;
    for (int t = threadIdx.x; t < nt; t += TB_THREADS) s_hist[t] = 0;
    chunk_pair_range(off, (uint32_t)n_pairs, begin, end, s_p);
    __syncthreads();
    const uint32_t p_lo = s_p[0], p_hi = s_p[1];
    visit_chunk(pairs, off, p_lo, p_hi, begin, end, P, gx, tiles_per_view, row0, [&](uint32_t li, uint32_t tile, uint32_t id) {
        atomicAdd(&s_hist[tile], 1u);
        s_tile[li] = (uint16_t)tile;
        s_id[li] = id;
    });
    __syncthreads();
    uint16_t* row = hist + (size_t)blockIdx.x * nt;
    for (int t = threadIdx.x; t < nt; t += TB_THREADS) row[t] = (uint16_t)s_hist[t];
    // (tile, id) of every instance, for the scatter pass: 6 B per instance, written and read once, both coalesced —
    // cheaper than walking the rects a second time (measured: the walk was 44 % of the scatter kernel's instructions)
    const uint32_t n_local = end - begin;
    for (uint32_t li = threadIdx.x; li < n_local; li += TB_THREADS) {
        inst_tile[begin + li] = s_tile[li];
        inst_id[begin + li] = s_id[li];
    }
}

// ---- 3. prefix over the tiles: the tile ranges ----------------------------------------------------
// Run by the 1024 threads of the LAST tile_scan CTA to finish (ticket below): no launch of its own.
__device__ void tile_starts(int nt, int tiles_local, int tiles_global, int tile_origin, const uint32_t* tile_total,
                            uint2* __restrict__ ranges) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int t0 = 0; t0 < nt; t0 += 1024) {
        const int t = t0 + threadIdx.x;
        const uint32_t v = t < nt ? __ldcg(tile_total + t) : 0u;   // written by other CTAs of this launch: L2, not L1
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane], xs = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, xs, o);
                if (lane >= o) xs += y;
            }
            s_warp[lane] = xs - w;   // exclusive
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t start = carry + s_warp[warp] + x - v;
        if (t < nt) {
            const int view = t >= tiles_local ? 1 : 0;   // bins are strip-local; the ranges are indexed by the global tile id
            ranges[view * tiles_global + tile_origin + (t - view * tiles_local)] = v ? make_uint2(start, start + v) : make_uint2(0u, 0u);
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = start + v;
        __syncthreads();
    }
}

// ---- 2. prefix over the chunks, per tile ----------------------------------------------------------
// CTA = 16 tile columns (one 32-byte sector per hist row) x 64 row groups: thread (col, g) loads its rows into shared
// memory (all loads independent, in flight together) and sums them, the groups are scanned through shared memory, then
// every thread re-walks its rows from shared memory writing the running prefix.  150 CTAs for 2 x 1200 tiles.
constexpr int TS_COLS = 16, TS_GROUPS = 64;
__global__ void __launch_bounds__(TS_COLS * TS_GROUPS) tile_scan_kernel(int nt, int n_pairs, const uint32_t* __restrict__ off,
                                                                         const uint32_t* __restrict__ n_vis, int cap,
                                                                         const uint16_t* __restrict__ hist,
                                                                         uint32_t* __restrict__ base,
                                                                         uint32_t* tile_total, uint32_t* ticket, int rows_cached,
                                                                         int tiles_global, int tile_origin, uint2* __restrict__ ranges,
                                                                         const EngineCtl* __restrict__ ctl) {
    if (ctl && ctl->level_done) return;
    extern __shared__ uint16_t s_rows[];                 // [rows_cached][TS_COLS]
    __shared__ bool s_last;
    __shared__ uint32_t s_part[TS_GROUPS][TS_COLS + 1];
    const int view = blockIdx.y;                         // nt = tiles per view; this view's chunks are rows [row_lo, row_lo + nact)
    const ChunkRange cr = chunk_range(off, n_vis, n_pairs, cap, 0u);
    const int col = threadIdx.x & (TS_COLS - 1), g = threadIdx.x / TS_COLS;
    const int tile = blockIdx.x * TS_COLS + col;
    const int row_lo = view ? (int)cr.rows0 : 0;
    const int nact = cr.overflow ? 0 : (int)(view ? cr.rows1 : cr.rows0);
    const int R = (nact + TS_GROUPS - 1) / TS_GROUPS;
    const int r0 = g * R, r1 = min(r0 + R, nact);
    uint32_t sum = 0;
    if (tile < nt) {
#pragma unroll 8
        for (int r = r0; r < r1; r++) {
            const uint16_t v = __ldg(hist + (size_t)(row_lo + r) * nt + tile);
            if (r < rows_cached) s_rows[r * TS_COLS + col] = v;
            sum += v;
        }
    }
    s_part[g][col] = sum;
    __syncthreads();
    if (g == 0) {
        uint32_t run = 0;
#pragma unroll 8
        for (int k = 0; k < TS_GROUPS; k++) {
            const uint32_t v = s_part[k][col];
            s_part[k][col] = run;
            run += v;
        }
        if (tile < nt) tile_total[view * nt + tile] = run;
    }
    __syncthreads();
    if (tile < nt) {
        uint32_t run = s_part[g][col];
        for (int r = r0; r < r1; r++) {
            const size_t i = (size_t)(row_lo + r) * nt + tile;
            base[i] = run;
            run += r < rows_cached ? s_rows[r * TS_COLS + col] : __ldg(hist + i);
        }
    }
    // the last CTA to get here turns the per-tile totals into the tile ranges
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) *ticket = 0u;
    tile_starts(2 * nt, nt, tiles_global, tile_origin, tile_total, ranges);
}

// ---- 4. stable scatter ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TB_THREADS, 4) tile_scatter_kernel(int n_pairs, int gx, int tiles_per_view, int row0, int tiles_global,
                                                                   const uint32_t* __restrict__ off, const uint32_t* __restrict__ n_vis,
                                                                   const uint16_t* __restrict__ inst_tile,
                                                                   const uint32_t* __restrict__ inst_id,
                                                                   const uint32_t* __restrict__ base,
                                                                   const uint2* __restrict__ ranges, uint32_t* __restrict__ values,
                                                                   int cap, const EngineCtl* __restrict__ ctl) {
    if (ctl && ctl->level_done) return;
    extern __shared__ uint32_t s_mem[];
    const int nt = tiles_per_view;
    uint32_t* s_gbase = s_mem;                 // [nt]    first slot of (this chunk, tile)
    uint32_t* s_cnt = s_mem + nt;              // [4][nt] per-slice counts, then cursors: slice w in word w & 3, half w >> 2
    uint32_t* s_lstart = s_mem + 5 * nt;       // [nt]    first position of the tile in the chunk's tile-ordered staging
    uint16_t* s_stamp = reinterpret_cast<uint16_t*>(s_mem + 6 * nt);   // [nt] scratch of the distinct-tiles test
    __shared__ uint32_t s_out_id[TB_CH];       // the chunk's instances ordered by (tile, rank): what leaves for tile t is one
    __shared__ uint16_t s_out_tile[TB_CH];     // contiguous run in shared AND in global memory -> few store sectors per warp
    __shared__ uint32_t s_wsum[8];
    const ChunkRange cr = chunk_range(off, n_vis, n_pairs, cap, blockIdx.x);
    if (!cr.active) return;                    // (an overflow was flagged by tile_count_kernel: the iteration is void)
    const uint32_t begin = cr.begin, end = cr.end;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // this lane's instances: slice `warp` of the chunk, step st -> local index warp * TB_SLICE + st * 32 + lane
    constexpr int STEPS = TB_SLICE / 32;
    const uint32_t first = begin + (uint32_t)warp * TB_SLICE + (uint32_t)lane;
    uint32_t tile[STEPS], id[STEPS];
#pragma unroll
    for (int st = 0; st < STEPS; st++) {
        const uint32_t o = first + (uint32_t)st * 32u;
        tile[st] = o < end ? (uint32_t)__ldg(inst_tile + o) : 0xFFFFFFFFu;
        id[st] = o < end ? __ldg(inst_id + o) : 0u;
    }
    const uint32_t* brow = base + (size_t)blockIdx.x * nt;
    for (int t = threadIdx.x; t < nt; t += TB_THREADS) {
        s_gbase[t] = __ldg(&ranges[(int)cr.view * tiles_global + row0 * gx + t].x) + __ldg(brow + t);
        s_cnt[t] = 0; s_cnt[nt + t] = 0; s_cnt[2 * nt + t] = 0; s_cnt[3 * nt + t] = 0;
    }
    __syncthreads();
    const uint32_t shift = 16u * (uint32_t)(warp >> 2);
    uint32_t* my_cnt = s_cnt + (warp & 3) * nt;
#pragma unroll
    for (int st = 0; st < STEPS; st++)
        if (tile[st] != 0xFFFFFFFFu) atomicAdd(&my_cnt[tile[st]], 1u << shift);
    __syncthreads();
    // counts -> exclusive prefix over the 8 slices, in place
    for (int t = threadIdx.x; t < nt; t += TB_THREADS) {
        uint32_t w[4] = {s_cnt[t], s_cnt[nt + t], s_cnt[2 * nt + t], s_cnt[3 * nt + t]};
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { const uint32_t c = w[k] & 0xFFFFu; w[k] = (w[k] & 0xFFFF0000u) | run; run += c; }
#pragma unroll
        for (int k = 0; k < 4; k++) { const uint32_t c = w[k] >> 16; w[k] = (w[k] & 0xFFFFu) | (run << 16); run += c; }
        s_cnt[t] = w[0]; s_cnt[nt + t] = w[1]; s_cnt[2 * nt + t] = w[2]; s_cnt[3 * nt + t] = w[3];
        s_lstart[t] = run;   // the chunk's count for this tile; prefixed over the tiles below
    }
    __syncthreads();
    {
        // exclusive prefix of the per-tile counts over the tiles: where each tile's run starts in the staging arrays
        const int per = (nt + TB_THREADS - 1) / TB_THREADS, t0 = threadIdx.x * per;
        uint32_t sum = 0;
        for (int k = 0; k < per; k++) sum += t0 + k < nt ? s_lstart[t0 + k] : 0u;
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t run = incl - sum;
        for (int w = 0; w < warp; w++) run += s_wsum[w];
        for (int k = 0; k < per; k++) {
            if (t0 + k < nt) { const uint32_t c = s_lstart[t0 + k]; s_lstart[t0 + k] = run; run += c; }
        }
    }
    __syncthreads();
    // every warp walks its slice in order: rank among the chunk's instances of the same tile -> staging position.
    // Two lanes of a step need ranking against each other only when they hit the same tile.  match.any answers that but
    // runs ~100 cycles per instruction on a shared unit (it bounded this kernel: 64 % unit utilisation at 26 % issue
    // slots), so a step first tests "all tiles distinct" with a stamp: every lane writes its own id to s_stamp[tile]
    // and reads it back — if two lanes share a tile at most one reads its own id back.  A foreign warp's stamp can only
    // produce a false alarm.  Steps with distinct tiles (about two in three at 1200 tiles) take one atomic per lane;
    // only the others fall back to match.any.  (compute-sanitizer racecheck reports exactly these two lines — the race
    // between warps on s_stamp is the mechanism, see above; nothing else in the library is flagged:
    // profiles/r1_sanitizer_v38.log.)  Shared-memory atomics of one warp execute in program order, which is
    // what keeps the ranks stable across steps.  Software-pipelined in groups of 8 steps.
    const uint32_t lt = (1u << lane) - 1u;
    const uint16_t my_stamp = (uint16_t)threadIdx.x;
    constexpr int GROUP = 8;
#pragma unroll
    for (int g0 = 0; g0 < STEPS; g0 += GROUP) {
        unsigned clash = 0;   // bit k: step g0 + k has two lanes on one tile (or a false alarm)
        uint32_t old[GROUP];
#pragma unroll
        for (int k = 0; k < GROUP; k++)
            if (tile[g0 + k] != 0xFFFFFFFFu) s_stamp[tile[g0 + k]] = (uint16_t)(my_stamp ^ (k << 8));
        __syncwarp();
#pragma unroll
        for (int k = 0; k < GROUP; k++) {
            const bool lost = tile[g0 + k] != 0xFFFFFFFFu && s_stamp[tile[g0 + k]] != (uint16_t)(my_stamp ^ (k << 8));
            if (__any_sync(0xffffffffu, lost)) clash |= 1u << k;
        }
#pragma unroll
        for (int k = 0; k < GROUP; k++) {
            const uint32_t t = tile[g0 + k];
            uint32_t o = 0, rank = 0;
            if (clash >> k & 1u) {
                const unsigned m = __match_any_sync(0xffffffffu, t);
                const int leader = __ffs(m) - 1;
                if (t != 0xFFFFFFFFu && lane == leader) o = atomicAdd(&my_cnt[t], (uint32_t)__popc(m) << shift);
                o = __shfl_sync(0xffffffffu, o, leader);
                rank = (uint32_t)__popc(m & lt);
            } else if (t != 0xFFFFFFFFu) {
                o = atomicAdd(&my_cnt[t], 1u << shift);
            }
            old[k] = ((o >> shift) & 0xFFFFu) + rank;
        }
#pragma unroll
        for (int k = 0; k < GROUP; k++) {
            if (tile[g0 + k] != 0xFFFFFFFFu) {
                const uint32_t pos = s_lstart[tile[g0 + k]] + old[k];
                s_out_id[pos] = id[g0 + k];
                s_out_tile[pos] = (uint16_t)tile[g0 + k];
            }
        }
    }
    __syncthreads();
    const uint32_t n_local = end - begin;
    for (uint32_t i = threadIdx.x; i < n_local; i += TB_THREADS) {
        const uint32_t t = s_out_tile[i];
        values[s_gbase[t] + (i - s_lstart[t])] = s_out_id[i];
    }
}

// ---- host side -----------------------------------------------------------------------------------
int tilebin_chunk() { return TB_CH; }
size_t tilebin_chunks(int cap) { return (size_t)((cap + TB_CH - 1) / TB_CH); }

int tilebin_configure(int max_tiles_per_view) {
    if (const char* v = getenv("GSEVT_TB_SMALL")) {
        const uint32_t k = (uint32_t)atoi(v);
        if (k >= 1 && k <= 65536) cudaMemcpyToSymbol(TB_SMALL, &k, sizeof(k));
    }
    const size_t smem = (size_t)max_tiles_per_view * 26;
    if (smem + TB_CH * 6 + 128 > 227 * 1024) return -1;
    cudaFuncSetAttribute(tile_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(tile_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)max_tiles_per_view * 4));
    cudaFuncSetAttribute(tile_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * TS_COLS * 2);
    return 0;
}

void launch_tile_count(const TileBinArgs& a, cudaStream_t s) {
    const int chunks = (a.cap + TB_CH - 1) / TB_CH + 1;   // + 1: the view boundary may cut one chunk in two
    if (a.n_pairs <= 0) return;
    tile_count_kernel<<<chunks, TB_THREADS, (size_t)a.tiles_per_view * 4, s>>>(
        a.P, a.n_pairs, a.grid_x, a.tiles_per_view, a.row0, a.pairs, a.offsets, a.n_vis, a.hist, a.inst_tile, a.inst_id, a.cap,
        a.overflow, a.ctl);
}
void launch_tile_scan(const TileBinArgs& a, cudaStream_t s) {
    const int nt = a.tiles_per_view;
    if (a.n_pairs <= 0) return;
    const int chunks = (a.cap + TB_CH - 1) / TB_CH + 1;
    const int rows_cached = chunks < 4096 ? chunks : 4096;   // 32 B per cached row: <= 128 KB
    tile_scan_kernel<<<dim3((nt + TS_COLS - 1) / TS_COLS, 2), TS_COLS * TS_GROUPS, (size_t)rows_cached * TS_COLS * 2, s>>>(
        nt, a.n_pairs, a.offsets, a.n_vis, a.cap, a.hist, a.base, a.tile_total, a.ticket, rows_cached, a.tiles_global, a.row0 * a.grid_x, a.ranges,
        a.ctl);
}
void launch_tile_scatter(const TileBinArgs& a, cudaStream_t s) {
    const int chunks = (a.cap + TB_CH - 1) / TB_CH + 1;
    if (a.n_pairs <= 0) return;
    tile_scatter_kernel<<<chunks, TB_THREADS, (size_t)a.tiles_per_view * 26, s>>>(
        a.n_pairs, a.grid_x, a.tiles_per_view, a.row0, a.tiles_global, a.offsets, a.n_vis, a.inst_tile, a.inst_id, a.base, a.ranges,
        a.values, a.cap, a.ctl);
}

// Parity-test helper: tile id of every slot of the per-tile lists (what the sorted keys of the reference hold in
// their high word), from the ranges.
__global__ void keys_from_ranges_kernel(int nt, const uint2* __restrict__ ranges, uint16_t* __restrict__ keys) {
    const int t = blockIdx.x;
    if (t >= nt) return;
    const uint2 r = ranges[t];
    for (uint32_t i = r.x + threadIdx.x; i < r.y; i += blockDim.x) keys[i] = (uint16_t)t;
}
void launch_keys_from_ranges(int nt, const uint2* ranges, uint16_t* keys, cudaStream_t s) {
    if (nt <= 0) return;
    keys_from_ranges_kernel<<<nt, 128, 0, s>>>(nt, ranges, keys);
}

}  // namespace gsevt
