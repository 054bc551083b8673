// Host side of the event ingestion: the text parser behind utils/event_camera/event.py:load_events_from_txt.
//
// The reference reads "ts x y p" lines with readlines() + split() + int() and builds one Python object per event
// (event.py:11-39): 0.31 M events/s measured, ~100 s for a 1000-frame sequence.  This is a two-pass, multi-threaded
// parser over the file bytes: pass 1 counts the whitespace-separated tokens of every thread's slice (slices start on
// token boundaries), pass 2 converts them in place into the caller's int64 table.  Token grammar = what bytes.split() +
// int() accept for the files GS-EVT reads: optional sign, decimal digits; anything else is an error, as in the
// reference (ValueError).  No CUDA in this file: it is host code living in the same library as the kernels.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <climits>
#include <cstring>
#include <thread>
#include <vector>
#include "internal.h"
#include "../../include/gsevt.h"

namespace {

inline bool is_space(unsigned char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// Tokens whose first byte lies in [lo, hi): count them, or convert them into out[].  Returns the number of tokens, or
// -1 - (byte offset of the offending token) on a malformed token.
int64_t scan_slice(const unsigned char* s, size_t len, size_t lo, size_t hi, int64_t* out, size_t room = ~(size_t)0) {
    int64_t n = 0;
    size_t i = lo;
    while (i < hi) {
        while (i < hi && is_space(s[i])) i++;
        if (i >= hi) break;
        const size_t start = i;
        bool neg = false;
        if (s[i] == '-' || s[i] == '+') { neg = s[i] == '-'; i++; }
        if (i >= len || s[i] < '0' || s[i] > '9') return -1 - (int64_t)start;
        uint64_t v = 0;
        int digits = 0;
        while (i < len && s[i] >= '0' && s[i] <= '9') { v = v * 10u + (uint64_t)(s[i] - '0'); i++; digits++; }   // a token may run past hi
        if (digits > 18) return -1 - (int64_t)start;                      // would not fit an int64 for sure
        if (i < len && !is_space(s[i])) return -1 - (int64_t)start;       // "12a", "1.5", "1_000": not an integer here
        if (out) {
            if ((size_t)n >= room) return INT64_MIN;                       // caller's table is full
            out[n] = neg ? -(int64_t)v : (int64_t)v;
        }
        n++;
    }
    return n;
}

}  // namespace

extern "C" {

GSEVT_API int64_t gsevt_parse_int_table(const char* text, size_t len, int64_t* out, size_t capacity, int32_t threads) {
    if (!text && len) { gsevt::set_error("parse_int_table: null text"); return GSEVT_EINVAL; }
    const unsigned char* s = reinterpret_cast<const unsigned char*>(text);
    int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    if (len < (size_t)1 << 20) T = 1;
    // slice boundaries on token starts: move each cut forward to the first byte after a whitespace
    std::vector<size_t> cut(T + 1);
    cut[0] = 0; cut[T] = len;
    for (int t = 1; t < T; t++) {
        size_t p = len / T * t;
        if (p < cut[t - 1]) p = cut[t - 1];
        while (p < len && p > 0 && !is_space(s[p - 1])) p++;   // inside a token: skip to its end
        cut[t] = p;
    }
    if (T == 1 && out) {
        // one pass: convert straight into the caller's table
        const int64_t n = scan_slice(s, len, 0, len, out, capacity);
        if (n == INT64_MIN) { gsevt::set_error("parse_int_table: more than %zu integers", capacity); return GSEVT_ENOMEM; }
        if (n < 0) { gsevt::set_error("parse_int_table: not an integer at byte %lld", (long long)(-1 - n)); return GSEVT_EINVAL; }
        return n;
    }
    std::vector<int64_t> count(T, 0);
    auto run = [&](bool convert, const std::vector<int64_t>& offset) {
        std::vector<std::thread> pool;
        for (int t = 0; t < T; t++)
            pool.emplace_back([&, t] { count[t] = scan_slice(s, len, cut[t], cut[t + 1], convert ? out + offset[t] : nullptr); });
        for (auto& th : pool) th.join();
    };
    std::vector<int64_t> offset(T, 0);
    run(false, offset);
    int64_t total = 0;
    for (int t = 0; t < T; t++) {
        if (count[t] < 0) { gsevt::set_error("parse_int_table: not an integer at byte %lld", (long long)(-1 - count[t])); return GSEVT_EINVAL; }
        offset[t] = total;
        total += count[t];
    }
    if (!out) return total;                                     // size query
    if ((size_t)total > capacity) { gsevt::set_error("parse_int_table: %lld integers, room for %zu", (long long)total, capacity); return GSEVT_ENOMEM; }
    run(true, offset);
    return total;
}

}  // extern "C"
