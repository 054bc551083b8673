// TEST INFRASTRUCTURE (CPU): drives gs-evt_b200/csrc/sortcore.cuh — the very functions the bucket binning kernels
// run per thread — sequentially, phase by phase, the way the CTA runs them between its barriers.  Built by
// tests/test_sortcore.py with g++; nothing in the product links this file.
#include <algorithm>
#include <cstring>
#include <vector>
#include "../gs-evt_b200/csrc/sortcore.cuh"

using namespace gsevt::sortcore;

extern "C" {

// 0/1 principle: returns the number of 0/1 inputs (of 256) the 8-key network fails to sort.
int sc_network_failures() {
    int bad = 0;
    for (int m = 0; m < 256; m++) {
        uint64_t k[VT];
        for (int i = 0; i < VT; i++) k[i] = (m >> i) & 1;
        sort8(k);
        for (int i = 1; i < VT; i++) bad += k[i - 1] > k[i];
    }
    return bad;
}

// The CTA's sort of one bucket: n keys (in place), `threads` simulated threads; returns the number of rounds.
int sc_sort(uint64_t* keys, int n, int threads) {
    const int np = (n + VT - 1) / VT * VT;
    std::vector<uint64_t> a(np, PAD), b(np, PAD);
    // phase 1: thread t of pass c loads keys [(c * threads + t) * VT, +VT), pads past n, sorts them
    for (int base = 0; base < np; base += threads * VT)
        for (int t = 0; t < threads; t++) {
            const int o = base + t * VT;
            if (o >= np) continue;
            uint64_t k[VT];
            for (int i = 0; i < VT; i++) k[i] = o + i < n ? keys[o + i] : PAD;
            sort8(k);
            for (int i = 0; i < VT; i++) a[o + i] = k[i];
        }
    uint64_t *src = a.data(), *dst = b.data();
    int rounds = 0;
    for (int run = VT; run < np; run *= 2, rounds++) {
        for (int t = 0; t < threads; t++)
            for (int seg = t; seg < np / VT; seg += threads) merge_segment(src, dst, np, run, seg);
        std::swap(src, dst);
    }
    memcpy(keys, src, sizeof(uint64_t) * (size_t)n);
    return rounds;
}

// The whole binning of one view on the host, in the kernels' data flow:
//   rect[P] (0 = not visible), depth_bits[P]  ->  per-tile lists (ids) + ranges, bucket shift s, strip rows [y0, y1).
// `order` permutes the pairs before the scatter (the GPU's atomics hand out bucket slots in arbitrary order).
// Returns the total number of tile instances; lists are written tile after tile (row-major) into list_out, ranges_out
// holds (begin, end) per tile of the whole grid, (0, 0) for untouched tiles — the reference's representation.
long long sc_bin_view(int P, const uint32_t* rect, const uint32_t* depth_bits, const int* order, int gx, int gy, int s, int y0, int y1,
                      int threads, uint32_t* list_out, uint32_t* ranges_out) {
    const int by_origin = y0 >> s;
    const int nbx = ((gx - 1) >> s) + 1, nby = y1 > y0 ? ((y1 - 1) >> s) - by_origin + 1 : 0;
    std::vector<std::vector<uint64_t>> bucket((size_t)nbx * nby);
    for (int q = 0; q < P; q++) {
        const int i = order ? order[q] : q;
        if (!rect[i]) continue;
        int bx0, bx1, by0, by1;
        bucket_rect(rect[i], s, by_origin, bx0, bx1, by0, by1);
        for (int by = by0; by < by1; by++)
            for (int bx = bx0; bx < bx1; bx++) bucket[(size_t)by * nbx + bx].push_back(((uint64_t)depth_bits[i] << 32) | (uint32_t)i);
    }
    std::vector<std::vector<uint32_t>> tile((size_t)gx * gy);
    for (int by = 0; by < nby; by++)
        for (int bx = 0; bx < nbx; bx++) {
            std::vector<uint64_t>& k = bucket[(size_t)by * nbx + bx];
            if (k.empty()) continue;
            sc_sort(k.data(), (int)k.size(), threads);
            const int tx0 = bx << s, ty0 = (by + by_origin) << s;
            for (uint64_t key : k) {
                const uint32_t id = (uint32_t)key;
                if (s == 0) {
                    tile[(size_t)ty0 * gx + tx0].push_back(id);
                } else {
                    const uint32_t m = cover_mask4(rect[id], tx0, ty0);
                    for (int kk = 0; kk < 4; kk++)
                        if (m >> kk & 1u) tile[(size_t)(ty0 + (kk >> 1)) * gx + tx0 + (kk & 1)].push_back(id);
                }
            }
        }
    long long total = 0;
    for (int t = 0; t < gx * gy; t++) {
        const std::vector<uint32_t>& l = tile[t];
        if (l.empty()) {
            ranges_out[2 * t] = ranges_out[2 * t + 1] = 0;
            continue;
        }
        ranges_out[2 * t] = (uint32_t)total;
        for (uint32_t id : l) list_out[total++] = id;
        ranges_out[2 * t + 1] = (uint32_t)total;
    }
    return total;
}

}  // extern "C"
