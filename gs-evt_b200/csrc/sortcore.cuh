// Building blocks of the per-bucket merge sort (bucketbin.cu), written as host/device code so that the CPU test
// tests/test_sortcore.py can drive the very same functions thread by thread (tests/sortcore_host.cpp).
//
// Keys are 64-bit and UNIQUE inside a bucket: (depth bits << 32) | Gaussian index.  Their ascending order is the
// reference's order inside a tile list — depth first, ties by Gaussian index, which the reference gets from the
// stability of its radix sort over the emission order (dgr/cuda_rasterizer/rasterizer_impl.cu:85-109, 306-311).
// Padding keys are ~0 (depth bits of a positive float never reach 0xFFFFFFFF) and sort behind everything.
//
// Scheme: every thread sorts VT = 8 consecutive keys with a 19-comparator network, then runs of 8, 16, 32, ...
// keys are merged pairwise; in each round the output is cut into segments of VT keys, a thread finds its segment's
// two input cursors with a merge-path search and merges VT keys serially.  No atomics, no ranking by counters: on
// this GPU a shared-memory atomic with 32 different addresses costs 2 cycles per lane (B300_MICROARCH.md), which
// is what bounded the previous counting kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GSEVT_HD __host__ __device__ __forceinline__
#else
#define GSEVT_HD inline
#endif

namespace gsevt {
namespace sortcore {

constexpr int VT = 8;
constexpr uint64_t PAD = ~0ull;

GSEVT_HD void cswap(uint64_t& a, uint64_t& b) {
    const bool sw = b < a;
    const uint64_t lo = sw ? b : a, hi = sw ? a : b;
    a = lo;
    b = hi;
}

// Optimal 19-comparator, 7-layer network for 8 keys (Knuth, TAOCP 3, 5.3.4).
GSEVT_HD void sort8(uint64_t (&k)[VT]) {
    cswap(k[0], k[1]); cswap(k[2], k[3]); cswap(k[4], k[5]); cswap(k[6], k[7]);
    cswap(k[0], k[2]); cswap(k[1], k[3]); cswap(k[4], k[6]); cswap(k[5], k[7]);
    cswap(k[1], k[2]); cswap(k[5], k[6]); cswap(k[0], k[4]); cswap(k[3], k[7]);
    cswap(k[1], k[5]); cswap(k[2], k[6]);
    cswap(k[1], k[4]); cswap(k[3], k[6]);
    cswap(k[2], k[4]); cswap(k[3], k[5]);
    cswap(k[3], k[4]);
}

// Number of keys of A among the first `diag` keys of merge(A, B), A first on ties.
GSEVT_HD int merge_path(const uint64_t* A, int lenA, const uint64_t* B, int lenB, int diag) {
    int lo = diag > lenB ? diag - lenB : 0;
    int hi = diag < lenA ? diag : lenA;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (A[mid] <= B[diag - 1 - mid]) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One segment of one merge round: keys [seg * VT, seg * VT + VT) of dst, where src holds sorted runs of `run` keys
// (run = VT * 2^r) over np keys (np a multiple of VT) and dst receives the runs merged pairwise.
GSEVT_HD void merge_segment(const uint64_t* src, uint64_t* dst, int np, int run, int seg) {
    const int o = seg * VT;
    const int pb = o & ~(2 * run - 1);
    const int a0 = pb;
    const int a1 = pb + run < np ? pb + run : np;
    const int b1 = pb + 2 * run < np ? pb + 2 * run : np;
    const int diag = o - pb;
    const int i = merge_path(src + a0, a1 - a0, src + a1, b1 - a1, diag);
    int ai = a0 + i, bi = a1 + (diag - i);
    uint64_t ka = ai < a1 ? src[ai] : PAD, kb = bi < b1 ? src[bi] : PAD;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int t = 0; t < VT; t++) {
        const bool take_a = bi >= b1 || (ai < a1 && ka <= kb);
        dst[o + t] = take_a ? ka : kb;
        if (take_a) {
            ++ai;
            ka = ai < a1 ? src[ai] : PAD;
        } else {
            ++bi;
            kb = bi < b1 ? src[bi] : PAD;
        }
    }
}

// Tile rect packing shared by the projection, the bucket scatter and the list emission:
// x0 | y0 << 8 | x1 << 16 | y1 << 24 in tile units, [x0, x1) x [y0, y1); 0 = not visible.
// Buckets are (1 << s) x (1 << s) blocks of tiles aligned to the tile grid; rows are counted from by_origin = (first
// tile row of the engine's strip) >> s.
GSEVT_HD void bucket_rect(uint32_t rect, int s, int by_origin, int& bx0, int& bx1, int& by0, int& by1) {
    const int x0 = (int)(rect & 255u), y0 = (int)(rect >> 8 & 255u), x1 = (int)(rect >> 16 & 255u), y1 = (int)(rect >> 24);
    bx0 = x0 >> s;
    bx1 = ((x1 - 1) >> s) + 1;
    by0 = (y0 >> s) - by_origin;
    by1 = ((y1 - 1) >> s) + 1 - by_origin;
}

// Which of the bucket's tiles (tx0 + kx, ty0 + ky), kx, ky < 2, does the tile rect cover?  bit ky * 2 + kx.
GSEVT_HD uint32_t cover_mask4(uint32_t rect, int tx0, int ty0) {
    const int x0 = (int)(rect & 255u), y0 = (int)(rect >> 8 & 255u), x1 = (int)(rect >> 16 & 255u), y1 = (int)(rect >> 24);
    const uint32_t cx0 = (x0 <= tx0 && tx0 < x1) ? 1u : 0u, cx1 = (x0 <= tx0 + 1 && tx0 + 1 < x1) ? 1u : 0u;
    const uint32_t cy0 = (y0 <= ty0 && ty0 < y1) ? 1u : 0u, cy1 = (y0 <= ty0 + 1 && ty0 + 1 < y1) ? 1u : 0u;
    return (cy0 & cx0) | ((cy0 & cx1) << 1) | ((cy1 & cx0) << 2) | ((cy1 & cx1) << 3);
}

}  // namespace sortcore
}  // namespace gsevt
