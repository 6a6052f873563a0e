// Host harness around beta9_b200/csrc/ring_place.h: the same bookkeeping b9gpu.cu keeps per push
// (deque of segments, write position = end of the newest segment, reset when empty), driven by a
// script of pushes and pops, with an independent overlap check over byte intervals.
//   g++ -O2 -std=c++17 -shared -fPIC -o libringplace.so ring_place_shim.cpp
#include <cstdint>
#include <deque>
#include "../../beta9_b200/csrc/ring_place.h"

namespace {
struct Seg { uint64_t start, bytes; };
}

extern "C" {

// ops[i] > 0: push of ops[i] - 1 bytes; ops[i] == 0: pop the oldest segment.
// out[i]: push -> start offset, or -1 when refused; pop -> 0 (or -2 if nothing to pop).
// Returns -1 - i at the first step whose placement overlaps a live segment or leaves the ring, else the number of steps run.
long b9_ring_place_script(uint64_t ring_bytes, const int64_t* ops, long n, int64_t* out) {
    std::deque<Seg> segs;
    uint64_t wp = 0;
    for (long i = 0; i < n; ++i) {
        if (ops[i] == 0) {
            if (segs.empty()) { out[i] = -2; continue; }
            segs.pop_front();
            if (segs.empty()) wp = 0;
            out[i] = 0;
            continue;
        }
        const uint64_t bytes = (uint64_t)(ops[i] - 1);
        uint64_t live = 0;
        for (const Seg& s : segs) live += s.bytes;
        uint64_t start = 0;
        if (!b9_ring_place(ring_bytes, !segs.empty(), segs.empty() ? 0 : segs.front().start, wp, live, bytes, &start)) { out[i] = -1; continue; }
        const uint64_t span = b9_seg_span(bytes);
        if (start % B9_SEG_ALIGN || start + span > ring_bytes) return -1 - i;
        for (const Seg& s : segs) {
            const uint64_t a0 = s.start, a1 = s.start + b9_seg_span(s.bytes);
            if (s.bytes && bytes && start < a1 && a0 < start + span) return -1 - i;
        }
        segs.push_back(Seg{start, bytes});
        wp = start + span;
        out[i] = (int64_t)start;
    }
    return n;
}

}  // extern "C"
