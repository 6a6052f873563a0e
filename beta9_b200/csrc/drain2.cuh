// drain kernel v2 — the production drain.
//
// Persistent CTAs, ticket work stealing and ballot compaction as in v1 (drain_kernel.cuh), with the
// data path rebuilt around the B200 memory system and the inter-CTA ordering chain removed:
//
//   * result RECORDS (id, status, has, offset, length) are FIFO-dense: record j belongs to the j-th
//     ready task. When no pending task is cancelled (the host knows) j is plain arithmetic on the
//     ticket; otherwise the per-tile ready counts — known from the slot words alone, long before
//     the tile is processed — go through a decoupled look-back;
//   * result BYTES are placed by ONE atomicAdd per tile on a byte cursor: dense, but in completion
//     order. (v1 chained the byte prefix through an in-order look-back; ncu showed the whole grid
//     falling into lockstep behind it, 32 tiles resolved per L2 round trip — profiles/r1_v2a_*.)
//
//   * a tile's payload bytes are ONE contiguous range of the ring in the common case; warp 0 pulls
//     it into shared memory with a single bulk async copy (cp.async.bulk / TMA 1-D, completion on
//     an mbarrier), double buffered so the next tile streams in while this one is processed;
//   * phase A and B are thread-per-task over the staged bytes (16-byte shared loads, SWAR
//     classification of the string body, 16-byte global stores): small uniform tasks keep all 32
//     lanes busy, which a warp-per-task layout cannot (a 284-byte task fills 18 of 32 lanes);
//   * strings with escapes / non-ASCII (the 1 % "adversarial" share, and most of configs[2]) are
//     NOT walked by one lane: the warp splits the body into 32 chunks, every lane finds its first
//     code-unit boundary by a bounded look-behind and transcodes its chunk (esc_* below);
//   * anything that is not the SDK's canonical frame goes through the sequential validating parser
//     (json_device.cuh) — correctness first, it is rare.
//
// Reference behaviour realised: see drain_kernel.cuh's header (pop, decode, loads, call, result).
#pragma once
#include <stdint.h>
#include "drain_kernel.cuh"

namespace b9 {

constexpr int D2_TASKS   = 64;               // tasks per tile
constexpr int D2_THREADS = D2_TASKS;         // (host code sizes tiles with this name)
constexpr int D2_STAGES  = 1;
// G threads share one task. Shared memory caps the TASKS resident on an SM (~900 for 284-byte
// payloads, one stage); thread-per-task would stop at ~16-25 warps/SM, too few to cover the
// shared-memory, atomic and barrier latencies (profiles/r1_v2c_*). G = 2 doubles the warps for the
// same bytes; G = 4 with two stages was slower (8-warp CTAs idling at block barriers, r1_v2d_*).
template <int HANDLER> struct D2Cfg { static constexpr int G = (HANDLER == 0) ? 2 : 1; static constexpr int THREADS = D2_TASKS * G; };

// ------------------------------------------------------------------ PTX: mbarrier + bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
// global -> shared::cta bulk async copy; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------ per-stage tile metadata
struct D2Meta {
    uint64_t goff[D2_TASKS];     // physical ring offset of each task's payload
    uint32_t soff[D2_TASKS];     // offset inside the stage buffer (when staged)
    uint32_t len[D2_TASKS];
    uint8_t  ready[D2_TASKS];
    unsigned long long tile;
    uint32_t nt;
    uint32_t staged;               // 1: payload bytes are (arriving) in shared memory; 0: read from global
    uint32_t base_cnt;             // records of all earlier tiles (index of this tile's first record)
    uint32_t ready_cnt;            // ready tasks in this tile
};

// ------------------------------------------------------------------ SWAR classification
// any byte of the four words outside printable ASCII, or equal to '"' or '\\'?
// Exactness: a false positive can only occur in a group that also holds a byte >= 0x80, which is
// "special" anyway (carries out of a byte need a byte >= 0x80 / 0xA0 below them).
__device__ __forceinline__ bool swar_special16(uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    const uint32_t K1 = 0x01010101u, K60 = 0x60606060u, K7F = 0x7F7F7F7Fu, Q = 0x22222222u, S = 0x5C5C5C5Cu, H = 0x80808080u;
    uint32_t hi = (x | y | z | w) | ((x + K1) | (y + K1) | (z + K1) | (w + K1));          // >= 0x7F
    uint32_t lo = (x + K60) & (y + K60) & (z + K60) & (w + K60);                          // bit7 clear: < 0x20
    uint32_t eq = ((x ^ Q) + K7F) & ((y ^ Q) + K7F) & ((z ^ Q) + K7F) & ((w ^ Q) + K7F)   // bit7 clear: == '"'
                & ((x ^ S) + K7F) & ((y ^ S) + K7F) & ((z ^ S) + K7F) & ((w ^ S) + K7F);  //            == '\\'
    return ((hi | ~lo | ~eq) & H) != 0;
}
__device__ __forceinline__ bool byte_special(uint32_t c) { return c < 0x20u || c >= 0x7Fu || c == '"' || c == '\\'; }

// 4 bytes at an arbitrary address: two aligned words + funnel shift
__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t* __restrict__ p) {
    const uint32_t a = (uint32_t)((uintptr_t)p & 3u);
    const uint32_t* w = (const uint32_t*)(p - a);
    return __funnelshift_r(w[0], w[1], a * 8u);
}
// per-byte 0xFF where the byte's index (0..15 across the four words) is >= lo and < hi
__device__ __forceinline__ uint32_t byte_range_mask(uint32_t word, uint32_t lo, uint32_t hi) {
    const uint32_t idx = 0x03020100u + 0x04040404u * word;
    return __vcmpgeu4(idx, lo * 0x01010101u) & __vcmpltu4(idx, hi * 0x01010101u);
}

// G threads per task (sub = 0..G-1, adjacent lanes): is the payload exactly
//   {"args": ["<body>"], "kwargs": {}}
// with a body of printable ASCII free of '"' and '\\'?  Returns bit0 = frame matches, bit1 = body
// clean, identical on all G lanes. Must be called by all 32 lanes of the warp (`active` masks the
// loads of lanes whose task does not exist).
template <int G>
__device__ __forceinline__ uint32_t quick_clean_framed(const uint8_t* __restrict__ p, uint32_t len, int sub, bool active) {
    uint32_t bad_frame = 0, special = 0;
    if (active && len >= FRAME_PRE_LEN + FRAME_SUF_LEN) {
        // frame: 11 + 17 bytes compared as 32-bit words ("{\"ar" "gs\":" " [\""  /  "\"], " "\"kwa" "rgs\"" ": {}" "}")
        const uint8_t* q = p + len - FRAME_SUF_LEN;
        if (0 % G == sub) bad_frame |= ld_u32_unaligned(p) != 0x7261227Bu;
        if (1 % G == sub) bad_frame |= ld_u32_unaligned(p + 4) != 0x3A227367u;
        if (2 % G == sub) bad_frame |= (ld_u32_unaligned(p + 8) & 0x00FFFFFFu) != 0x00225B20u;
        if (3 % G == sub) bad_frame |= ld_u32_unaligned(q) != 0x202C5D22u;
        if (4 % G == sub) bad_frame |= ld_u32_unaligned(q + 4) != 0x61776B22u;
        if (5 % G == sub) bad_frame |= ld_u32_unaligned(q + 8) != 0x22736772u;
        if (6 % G == sub) bad_frame |= ld_u32_unaligned(q + 12) != 0x7D7B203Au;
        if (7 % G == sub) bad_frame |= q[16] != '}';
        // body = [b0, b1): aligned 16-byte groups, group i handled by lane i % G; bytes of the first /
        // last group outside the body are replaced by 'a' before the SWAR test. (A garbage frame makes
        // the answer irrelevant, but the loads stay inside the payload: len >= 28.)
        const uint8_t* b0 = p + FRAME_PRE_LEN;
        const uint8_t* b1 = q;
        // (pointer - integer keeps the address space the compiler inferred; an integer round trip loses it)
        const uint8_t* g0 = b0 - ((uintptr_t)b0 & 15u);
        const uint8_t* g1 = b1 + ((16u - ((uintptr_t)b1 & 15u)) & 15u);
        const uint32_t lead = (uint32_t)(b0 - g0);                  // bytes to ignore at the front of the first group
        const uint32_t tail_keep = 16u - (uint32_t)(g1 - b1);       // bytes to keep in the last group
        const uint32_t A = 0x61616161u;
        for (const uint8_t* g = g0 + 16 * sub; g < g1; g += 16 * G) {
            uint4 v = *(const uint4*)g;
            const bool first = g == g0, last = g + 16 == g1;
            if (first | last) {
                const uint32_t lo = first ? lead : 0u, hi = last ? tail_keep : 16u;
                uint32_t m;
                m = byte_range_mask(0, lo, hi); v.x = (v.x & m) | (A & ~m);
                m = byte_range_mask(1, lo, hi); v.y = (v.y & m) | (A & ~m);
                m = byte_range_mask(2, lo, hi); v.z = (v.z & m) | (A & ~m);
                m = byte_range_mask(3, lo, hi); v.w = (v.w & m) | (A & ~m);
            }
            special |= swar_special16(v.x, v.y, v.z, v.w) ? 1u : 0u;
        }
    } else bad_frame = 1;
    uint32_t bits = bad_frame | (special << 1);
    #pragma unroll
    for (int d = 1; d < G; d <<= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, d);
    return (bits & 1u ? 0u : 1u) | (bits & 2u ? 0u : 2u);
}

// up to 15 bytes, destination alignment known to allow the 1/2/4/8-byte ladder used by the callers
__device__ __forceinline__ void copy_small_up(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n) {
    // dst + n is 16-byte aligned (head of a copy): ascending sizes keep every store naturally aligned
    uint32_t i = 0;
    if (n & 1u) { dst[0] = src[0]; i = 1; }
    if (n & 2u) { *(uint16_t*)(dst + i) = (uint16_t)(src[i] | (src[i + 1] << 8)); i += 2; }
    if (n & 4u) { *(uint32_t*)(dst + i) = ld_u32_unaligned(src + i); i += 4; }
    if (n & 8u) { *(uint2*)(dst + i) = make_uint2(ld_u32_unaligned(src + i), ld_u32_unaligned(src + i + 4)); }
}
__device__ __forceinline__ void copy_small_down(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n) {
    // dst is 16-byte aligned (tail of a copy): descending sizes
    uint32_t i = 0;
    if (n & 8u) { *(uint2*)(dst) = make_uint2(ld_u32_unaligned(src), ld_u32_unaligned(src + 4)); i = 8; }
    if (n & 4u) { *(uint32_t*)(dst + i) = ld_u32_unaligned(src + i); i += 4; }
    if (n & 2u) { *(uint16_t*)(dst + i) = (uint16_t)(src[i] | (src[i + 1] << 8)); i += 2; }
    if (n & 1u) dst[i] = src[i];
}

// Copy of n bytes to global memory by the G threads of a task: 16-byte stores on the destination
// (vector v by lane v % G), 4-byte loads + funnel shift on the (arbitrarily aligned) source.
template <int G>
__device__ __forceinline__ void group_copy(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n, int sub) {
    uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);
    if (head > n) {                                              // tiny copy that never reaches an aligned vector
        if (sub == 0) for (uint32_t i = 0; i < n; ++i) dst[i] = src[i];
        return;
    }
    if (sub == 0) copy_small_up(dst, src, head);
    dst += head; src += head; n -= head;
    const uint32_t nvec = n >> 4;
    const uint32_t sh = (uint32_t)((uintptr_t)src & 3u), bits = sh * 8;
    const uint32_t* sw = (const uint32_t*)(src - sh);
    uint4* dv = (uint4*)dst;
    for (uint32_t v = sub; v < nvec; v += G) {
        const uint32_t* s4 = sw + 4 * v;
        const uint32_t w0 = s4[0], w1 = s4[1], w2 = s4[2], w3 = s4[3], w4 = s4[4];   // w4: <= 3 bytes of over-read, inside the buffers' slack
        uint4 o;
        o.x = __funnelshift_r(w0, w1, bits); o.y = __funnelshift_r(w1, w2, bits);
        o.z = __funnelshift_r(w2, w3, bits); o.w = __funnelshift_r(w3, w4, bits);
        dv[v] = o;
    }
    if (sub == G - 1) { const uint32_t done = nvec << 4; copy_small_down(dst + done, src + done, n - done); }
}

// ------------------------------------------------------------------ chunk-parallel string transcoding
// body = the bytes between the frame's quotes. A "unit" is what Go's unquote consumes at once: a
// plain byte, an escape, a \uXXXX (with its low surrogate partner), a UTF-8 sequence.
__device__ __forceinline__ int hex4(const uint8_t* __restrict__ b) {
    int h0 = hexval(b[0]), h1 = hexval(b[1]), h2 = hexval(b[2]), h3 = hexval(b[3]);
    if ((h0 | h1 | h2 | h3) < 0) return -1;
    return (h0 << 12) | (h1 << 8) | (h2 << 4) | h3;
}
__device__ __forceinline__ uint32_t bs_run_before(const uint8_t* __restrict__ b, uint32_t q) {   // backslashes ending at q-1
    uint32_t r = 0;
    while (r < q && b[q - 1 - r] == '\\') ++r;
    return r;
}
// is there a "\uDC00..\uDFFF" escape text at position q?
__device__ __forceinline__ bool low_surrogate_at(const uint8_t* __restrict__ b, uint32_t q, uint32_t n) {
    if (q + 6 > n || b[q] != '\\' || b[q + 1] != 'u') return false;
    int r = hex4(b + q + 2);
    return r >= 0xDC00 && r <= 0xDFFF;
}
__device__ __forceinline__ bool high_surrogate_escape_at(const uint8_t* __restrict__ b, uint32_t q, uint32_t n) {
    if (q + 6 > n || b[q] != '\\' || b[q + 1] != 'u') return false;
    int r = hex4(b + q + 2);
    return r >= 0xD800 && r <= 0xDBFF;
}
__device__ __forceinline__ uint32_t utf8_valid_len(const uint8_t* __restrict__ b, uint32_t q, uint32_t n) {
    // length of the valid UTF-8 sequence starting at q (Go utf8.DecodeRune), or 0
    uint8_t c = b[q]; uint32_t rem = n - q;
    if (c >= 0xC2 && c <= 0xDF) return (rem >= 2 && (b[q + 1] & 0xC0) == 0x80) ? 2u : 0u;
    if (c >= 0xE0 && c <= 0xEF) {
        uint8_t lo = (c == 0xE0) ? 0xA0 : 0x80, hi = (c == 0xED) ? 0x9F : 0xBF;
        return (rem >= 3 && b[q + 1] >= lo && b[q + 1] <= hi && (b[q + 2] & 0xC0) == 0x80) ? 3u : 0u;
    }
    if (c >= 0xF0 && c <= 0xF4) {
        uint8_t lo = (c == 0xF0) ? 0x90 : 0x80, hi = (c == 0xF4) ? 0x8F : 0xBF;
        return (rem >= 4 && b[q + 1] >= lo && b[q + 1] <= hi && (b[q + 2] & 0xC0) == 0x80 && (b[q + 3] & 0xC0) == 0x80) ? 4u : 0u;
    }
    return 0u;
}

// First unit boundary at or after `lo`, decided from a bounded neighbourhood of lo. Exact for a
// body whose escapes are all well-formed; for a malformed body some lane reports !ok and the
// answer is discarded.
__device__ __noinline__ uint32_t first_unit_start(const uint8_t* __restrict__ b, uint32_t n, uint32_t lo) {
    if (lo == 0 || lo >= n) return lo;
    // (1) lo is the character after an escape's backslash
    if (bs_run_before(b, lo) & 1u) {
        uint32_t next = lo + 1;
        if (b[lo] == 'u') {
            next = lo + 5;
            if (high_surrogate_escape_at(b, lo - 1, n) && low_surrogate_at(b, next, n)) next += 6;
        }
        return next;
    }
    // (2) lo is one of the hex digits of a \uXXXX that began 2..5 bytes earlier
    #pragma unroll
    for (uint32_t k = 2; k <= 5; ++k) {
        if (lo >= k) {
            uint32_t q = lo - k;
            if (b[q] == '\\' && b[q + 1] == 'u' && !(bs_run_before(b, q) & 1u)) {
                uint32_t next = q + 6;
                if (high_surrogate_escape_at(b, q, n) && low_surrogate_at(b, next, n)) next += 6;
                return next;
            }
        }
    }
    // (3) lo starts the low-surrogate escape that the high surrogate 6 bytes earlier consumes
    if (lo >= 6 && low_surrogate_at(b, lo, n) && high_surrogate_escape_at(b, lo - 6, n) && !(bs_run_before(b, lo - 6) & 1u))
        return lo + 6;
    // (4) lo is a continuation byte of a valid UTF-8 sequence that began 1..3 bytes earlier
    if ((b[lo] & 0xC0) == 0x80) {
        for (uint32_t k = 1; k <= 3 && k <= lo; ++k) {
            uint8_t c = b[lo - k];
            if ((c & 0xC0) == 0x80) continue;            // another continuation byte: keep looking back
            if (c >= 0xC2) { uint32_t L = utf8_valid_len(b, lo - k, n); if (L > k) return lo - k + L; }
            break;
        }
    }
    return lo;
}

// One unit at b[i] (i < n): advances i, returns the code point Go's unquote yields; *ok = false on
// anything that is not a well-formed JSON string body byte (bad escape, raw control byte, raw '"').
__device__ __noinline__ uint32_t next_unit(const uint8_t* __restrict__ b, uint32_t& i, uint32_t n, bool* ok) {
    uint8_t c = b[i];
    if (c == '\\') {
        if (i + 1 >= n) { *ok = false; ++i; return 0; }
        uint8_t e = b[i + 1];
        if (e == 'u') {
            int r = (i + 6 <= n) ? hex4(b + i + 2) : -1;
            if (r < 0) { *ok = false; i += 2; return 0; }
            i += 6;
            if (r >= 0xD800 && r <= 0xDFFF) {
                if (r <= 0xDBFF && i + 6 <= n && b[i] == '\\' && b[i + 1] == 'u') {
                    int r1 = hex4(b + i + 2);
                    if (r1 >= 0xDC00 && r1 <= 0xDFFF) { i += 6; return 0x10000u + (((uint32_t)r - 0xD800u) << 10) + ((uint32_t)r1 - 0xDC00u); }
                }
                return 0xFFFDu;
            }
            return (uint32_t)r;
        }
        i += 2;
        switch (e) {
        case '"': case '\\': case '/': return e;
        case 'b': return 8; case 'f': return 12; case 'n': return 10; case 'r': return 13; case 't': return 9;
        default: *ok = false; return 0;
        }
    }
    if (c < 0x20 || c == '"') { *ok = false; ++i; return 0; }
    if (c < 0x80) { ++i; return c; }
    uint32_t L = utf8_valid_len(b, i, n);
    if (!L) { ++i; return 0xFFFDu; }
    uint32_t cp;
    if (L == 2) cp = ((c & 0x1Fu) << 6) | (b[i + 1] & 0x3Fu);
    else if (L == 3) cp = ((c & 0x0Fu) << 12) | ((b[i + 1] & 0x3Fu) << 6) | (b[i + 2] & 0x3Fu);
    else cp = ((c & 0x07u) << 18) | ((b[i + 1] & 0x3Fu) << 12) | ((b[i + 2] & 0x3Fu) << 6) | (b[i + 3] & 0x3Fu);
    i += L;
    return cp;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane) {
    uint32_t inc = v;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    return inc - v;
}

__device__ __forceinline__ bool plain_byte(uint32_t c) { return c >= 0x20u && c < 0x7Fu && c != '"' && c != '\\'; }

// Warp-cooperative: json.dumps length of the framed string body, or false if the body is not a
// well-formed JSON string ending exactly at the frame's closing quote. Every lane walks its chunk
// as  (run of plain bytes)* (one escape / UTF-8 unit)  so that the branchy unit decoder runs once
// per round for the whole warp. `lane_info` (optional, shared memory, 32 words) keeps each lane's
// first-unit offset and output length for esc_emit.
__device__ __forceinline__ bool esc_measure(const uint8_t* __restrict__ body, uint32_t n, int lane, uint32_t* out_len, uint32_t* lane_info) {
    const uint32_t S = (n + 31u) / 32u;
    const uint32_t lo = min(n, (uint32_t)lane * S), hi = min(n, lo + S);
    bool ok = true;
    uint32_t mine = 0, start = lo;
    if (lo < hi) {
        start = first_unit_start(body, n, lo);
        uint32_t i = start;
        while (i < hi) {
            while (i < hi && plain_byte(body[i])) { ++i; ++mine; }
            if (i >= hi) break;
            mine += py_escaped_len(next_unit(body, i, n, &ok));
            if (!ok) break;
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    *out_len = 2u + warp_sum(mine);
    if (lane_info) lane_info[lane] = mine | ((start - lo) << 24);     // chunk output < 2^24 bytes for any accepted payload
    return ok;
}

// Warp-cooperative: write json.dumps(body) to dst (global). Only called after esc_measure said ok.
__device__ __forceinline__ void esc_emit(const uint8_t* __restrict__ body, uint32_t n, int lane, uint8_t* __restrict__ dst, const uint32_t* lane_info) {
    const uint32_t S = (n + 31u) / 32u;
    const uint32_t lo = min(n, (uint32_t)lane * S), hi = min(n, lo + S);
    bool ok = true;
    uint32_t mine = 0, start = lo;
    if (lane_info) { const uint32_t w = lane_info[lane]; mine = w & 0xFFFFFFu; start = lo + (w >> 24); }
    else if (lo < hi) {
        start = first_unit_start(body, n, lo);
        uint32_t i = start;
        while (i < hi) {
            while (i < hi && plain_byte(body[i])) { ++i; ++mine; }
            if (i >= hi) break;
            mine += py_escaped_len(next_unit(body, i, n, &ok));
        }
    }
    const uint32_t at = 1u + warp_excl_scan(mine, lane);
    if (lane == 0) dst[0] = '"';
    if (lane == 31) dst[at + mine] = '"';
    if (lo < hi) {
        uint8_t* o = dst + at;
        uint32_t i = start;
        while (i < hi) {
            while (i < hi) { const uint32_t c = body[i]; if (!plain_byte(c)) break; *o++ = (uint8_t)c; ++i; }
            if (i >= hi) break;
            o += py_emit(next_unit(body, i, n, &ok), o);
        }
    }
}

// ------------------------------------------------------------------ the kernel
constexpr int D2_ESC_SLOTS = 8;          // escaped strings per tile whose per-lane sizes are kept for phase B
template <int HANDLER>
struct D2Shared {
    D2Meta meta[D2_STAGES];
    TaskRec rec[D2_TASKS];
    uint32_t excl_bytes[D2_TASKS];       // exclusive prefix of out_len inside the tile
    uint32_t excl_cnt[D2_TASKS];         // exclusive prefix of ready inside the tile
    uint32_t slow_list[D2_TASKS]; uint32_t n_slow;
    uint32_t esc_info[D2_ESC_SLOTS][32];
    unsigned long long base;
    uint32_t crc_table[HANDLER == 1 ? 256 : 1];
    alignas(8) uint64_t mbar[D2_STAGES];
};

// warp 0, step 1: start the loads of a tile's slot words (two tasks per lane); nothing is consumed
// here, so the latency overlaps whatever the warp does until d2_stage_tile.
struct D2MetaRegs { uint64_t off[2], hdr[2]; };
__device__ __forceinline__ void d2_load_meta(const DrainArgs& a, unsigned long long tile, int lane, D2MetaRegs& r) {
    const uint32_t t0 = (uint32_t)tile * D2_TASKS;
    const uint32_t nt = min((uint32_t)D2_TASKS, a.n_tasks - t0);
    #pragma unroll
    for (int q = 0; q < 2; ++q) {
        const uint32_t k = lane + 32 * q;
        r.off[q] = 0; r.hdr[q] = 0;
        if (k < nt) {
            const uint32_t slot = (uint32_t)((a.first_task + t0 + k) & a.slot_mask);
            r.hdr[q] = __ldg(a.hdr + slot);
            r.off[q] = __ldg(a.off + slot);
        }
    }
}

// warp 0, step 2: decide how to stage the tile and fire the bulk copies
__device__ inline void d2_stage_tile(const DrainArgs& a, unsigned long long tile, const D2MetaRegs& r, D2Meta& m, uint8_t* buf,
                                     uint32_t in_cap, uint64_t* bar, int lane) {
    const uint32_t t0 = (uint32_t)tile * D2_TASKS;
    const uint32_t nt = min((uint32_t)D2_TASKS, a.n_tasks - t0);
    uint64_t off[2], end[2]; uint32_t len[2]; bool valid[2];
    #pragma unroll
    for (int q = 0; q < 2; ++q) {
        const uint32_t k = lane + 32 * q;
        valid[q] = k < nt;
        off[q] = r.off[q]; len[q] = valid[q] ? hdr_len(r.hdr[q]) : 0u;
        if (valid[q]) { m.goff[k] = off[q]; m.len[k] = len[q]; m.ready[k] = !(hdr_flags(r.hdr[q]) & 1u); }
        end[q] = off[q] + len[q];
    }
    // contiguous?  task k starts where task k-1 ended
    uint64_t prev0 = __shfl_up_sync(0xffffffffu, end[0], 1);
    uint64_t prev1 = __shfl_up_sync(0xffffffffu, end[1], 1);
    const uint64_t end0_last = __shfl_sync(0xffffffffu, end[0], 31);
    if (lane == 0) { prev0 = off[0]; prev1 = end0_last; }
    bool contig = (!valid[0] || off[0] == prev0) && (!valid[1] || off[1] == prev1);
    contig = __all_sync(0xffffffffu, contig);
    const uint64_t gs = __shfl_sync(0xffffffffu, off[0], 0);
    const uint32_t last = nt - 1;
    const uint64_t ge0 = __shfl_sync(0xffffffffu, end[0], last & 31), ge1 = __shfl_sync(0xffffffffu, end[1], last & 31);
    const uint64_t ge = (last < 32) ? ge0 : ge1;
    uint32_t staged = 0;
    if (contig) {
        const uint64_t as = gs & ~15ull;
        const uint64_t bytes = ((ge + 15ull) & ~15ull) - as;
        if (bytes <= in_cap) {
            staged = 1;
            #pragma unroll
            for (int q = 0; q < 2; ++q) if (valid[q]) m.soff[lane + 32 * q] = (uint32_t)(off[q] - as);
            if (lane == 0) {
                mbar_expect_tx(bar, (uint32_t)bytes);
                if (bytes) bulk_g2s(buf, a.payload + as, (uint32_t)bytes, bar);
            }
        }
    } else {
        // scattered tile (it spans pushes): one copy per task, each widened to 16-byte boundaries
        uint32_t asz[2];
        #pragma unroll
        for (int q = 0; q < 2; ++q) asz[q] = (valid[q] && len[q]) ? (uint32_t)(((end[q] + 15ull) & ~15ull) - (off[q] & ~15ull)) : 0u;
        const uint32_t ex0 = warp_excl_scan(asz[0], lane);
        const uint32_t tot0 = __shfl_sync(0xffffffffu, ex0 + asz[0], 31);
        const uint32_t ex1 = tot0 + warp_excl_scan(asz[1], lane);
        const uint32_t total = __shfl_sync(0xffffffffu, ex1 + asz[1], 31);
        if (total <= in_cap) {
            staged = 1;
            if (lane == 0) mbar_expect_tx(bar, total);
            __syncwarp();
            if (valid[0]) { m.soff[lane] = ex0 + (uint32_t)(off[0] & 15ull); if (asz[0]) bulk_g2s(buf + ex0, a.payload + (off[0] & ~15ull), asz[0], bar); }
            if (valid[1]) { m.soff[lane + 32] = ex1 + (uint32_t)(off[1] & 15ull); if (asz[1]) bulk_g2s(buf + ex1, a.payload + (off[1] & ~15ull), asz[1], bar); }
        }
    }
    // ---- record indices: ready counts are known from the slot words alone
    const uint32_t rc = __popc(__ballot_sync(0xffffffffu, valid[0] && !(hdr_flags(r.hdr[0]) & 1u)))
                      + __popc(__ballot_sync(0xffffffffu, valid[1] && !(hdr_flags(r.hdr[1]) & 1u)));
    uint32_t base_cnt = t0;                                   // nothing cancelled anywhere: pure arithmetic
    if (a.count_mode) {                                       // some pending task is cancelled: chain the counts
        uint64_t excl = 0;
        if (tile == 0) { if (lane == 0) st_volatile_u64(a.tile_state + 0, LB_INC | rc); }
        else {
            if (lane == 0) st_volatile_u64(a.tile_state + tile, LB_AGG | rc);
            long long look = (long long)tile - 1;
            for (;;) {
                const long long idx = look - lane;
                uint64_t w = (idx >= 0) ? ld_volatile_u64(a.tile_state + idx) : LB_INC;
                while (__any_sync(0xffffffffu, (w & LB_STATUS) == 0)) { if ((w & LB_STATUS) == 0) w = ld_volatile_u64(a.tile_state + idx); }
                const uint32_t inc_mask = __ballot_sync(0xffffffffu, (w & LB_STATUS) == LB_INC);
                uint64_t v = lb_value(w);
                if (inc_mask) { const int first = __ffs(inc_mask) - 1; if (lane > first) v = 0; }
                #pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                excl += v;
                if (inc_mask) break;
                look -= 32;
            }
            if (lane == 0) st_volatile_u64(a.tile_state + tile, LB_INC | (excl + rc));
        }
        base_cnt = (uint32_t)excl;
    }
    if (lane == 0) {
        m.tile = tile; m.nt = nt; m.staged = staged; m.base_cnt = base_cnt; m.ready_cnt = rc;
        if (tile == a.n_tiles - 1) a.ctl->total_cnt = base_cnt + rc;
    }
}

// ---- per-task bodies, inlined once for shared-memory payloads (LDS) and once for global ones ----
template <int HANDLER>
__device__ __forceinline__ void d2_phase_b_task(const uint8_t* __restrict__ p, const TaskRec& rec, uint8_t* __restrict__ o) {
    if (rec.mode == OM_VADD) vadd_write(p, rec.src_off, rec.src_len, o);
    else if (rec.mode == OM_U32_DEC || rec.mode == OM_I64_DEC) {
        long long v = rec.value; uint32_t l = rec.out_len;
        if (v < 0) { *o++ = '-'; --l; v = -v; }
        write_dec(o, (unsigned long long)v, l);
    } else if (rec.mode == OM_STR_ESC) {                          // string the sequential parser sized (non-canonical frame)
        uint32_t i = rec.src_off + 1, end = rec.src_off + rec.src_len - 1;
        *o++ = '"';
        while (i < end) o += py_emit(next_cp(p, i, end), o);
        *o = '"';
    }
}

// the sequential validating parser + handler sizing, out of line: rare for identity, and it keeps the
// hot loops' registers and instruction-cache footprint small
template <int HANDLER>
__device__ __noinline__ void d2_parse_and_size(const uint8_t* p, uint32_t len, TaskRec& rec, const uint32_t* crc_table) {
    Parsed pr = parse_payload(p, len);
    handler_phase_a(HANDLER, p, pr, rec, crc_table);
}

template <int HANDLER>
__global__ void __launch_bounds__(D2Cfg<HANDLER>::THREADS, 8) drain2_kernel(DrainArgs a, uint32_t in_cap) {
    constexpr int G = D2Cfg<HANDLER>::G, THREADS = D2Cfg<HANDLER>::THREADS, WARPS = THREADS / 32;
    extern __shared__ __align__(128) uint8_t d2_smem[];
    using Sh = D2Shared<HANDLER>;
    Sh& S = *reinterpret_cast<Sh*>(d2_smem);
    uint8_t* const bufs = d2_smem + ((sizeof(Sh) + 127u) & ~127u);
    const uint32_t buf_stride = (in_cap + 64u + 127u) & ~127u;             // 64 bytes of readable slack behind each stage
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = tid / G, sub = tid % G;                                  // my task inside the tile, my share of it

    if (tid == 0) { mbar_init(&S.mbar[0], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); S.n_slow = 0; }
    if (HANDLER == 1) for (int i = tid; i < 256; i += THREADS) S.crc_table[i] = crc_table_entry(i);

    // warp 0 keeps two tickets ahead of the tile being processed: the slot words of the next tile are
    // already in registers when its turn comes (no global latency between the end of one tile and the
    // bulk copy of the next), and the ticket after that is in flight. Holding tickets is harmless: no
    // CTA waits on another CTA's unprocessed tile (the byte cursor is an atomic, not a chain).
    unsigned long long t_cur = ~0ull;      // tile to process next; its slot words are in `mregs`
    unsigned long long t_raw = ~0ull;      // lane 0: ticket claimed for the tile after t_cur
    D2MetaRegs mregs; mregs.off[0] = mregs.off[1] = mregs.hdr[0] = mregs.hdr[1] = 0;
    if (warp == 0) {
        if (lane == 0) { t_cur = atomicAdd(&a.ctl->ticket, 1ull); t_raw = atomicAdd(&a.ctl->ticket, 1ull); }
        t_cur = __shfl_sync(0xffffffffu, t_cur, 0);
        if (t_cur < a.n_tiles) d2_load_meta(a, t_cur, lane, mregs);
    }
    __syncthreads();
    uint32_t parity = 0;
    const uint32_t stage = 0;
    D2Meta& M = S.meta[0];

    for (;;) {
        if (warp == 0) {
            if (t_cur < a.n_tiles) d2_stage_tile(a, t_cur, mregs, M, bufs, in_cap, &S.mbar[0], lane);
            else if (lane == 0) M.tile = t_cur;
            t_cur = __shfl_sync(0xffffffffu, t_raw, 0);
            if (t_cur < a.n_tiles) d2_load_meta(a, t_cur, lane, mregs);
            if (lane == 0) t_raw = atomicAdd(&a.ctl->ticket, 1ull);
        }
        __syncthreads();                                                   // [0] tile metadata visible
        const unsigned long long tile = M.tile;
        if (tile >= a.n_tiles) break;
        const uint32_t nt = M.nt;
        const uint32_t t0 = (uint32_t)tile * D2_TASKS;
        const bool staged = M.staged != 0;
        const uint8_t* const sbuf = bufs + (size_t)stage * buf_stride;
        if (staged) { mbar_wait(&S.mbar[0], parity); parity ^= 1u; }

        // ---------------- phase A: G threads per task ------------------------------------------------
        const bool mine = k < (int)nt && M.ready[k];
        const uint32_t my_len = mine ? M.len[k] : 0u;
        TaskRec rec; rec.ready = mine; rec.status = 0; rec.has = 0; rec.mode = OM_NONE; rec.out_len = 0; rec.src_off = 0; rec.src_len = 0; rec.value = 0;
        if (HANDLER == 0) {
            uint32_t q;
            if (staged) q = quick_clean_framed<G>(sbuf + (mine ? M.soff[k] : 0u), my_len, sub, mine);
            else        q = quick_clean_framed<G>(a.payload + (mine ? M.goff[k] : 0ull), my_len, sub, mine);
            if (mine) {
                if (q == 3u) {
                    const uint32_t tok = my_len - FRAME_PRE_LEN - FRAME_SUF_LEN + 2;
                    if (tok > 2) { rec.has = 1; rec.mode = OM_COPY; rec.src_off = FRAME_PRE_LEN - 1; rec.src_len = tok; rec.out_len = tok; }
                } else {
                    rec.mode = (q & 1u) ? OM_STR_PAR : OM_NONE;          // settled in the cooperative pass
                    if (sub == 0) S.slow_list[atomicAdd(&S.n_slow, 1u)] = (uint32_t)k;
                }
            }
        } else if (mine) {
            const uint8_t* p = staged ? (const uint8_t*)(sbuf + M.soff[k]) : a.payload + M.goff[k];
            d2_parse_and_size<HANDLER>(p, my_len, rec, S.crc_table);
        }
        if (sub == 0 && k < D2_TASKS) S.rec[k] = rec;
        __syncthreads();                                                   // [1] records + slow list complete
        // ---------------- phase A, cooperative pass over the tasks the quick look could not settle --
        const uint32_t ns = (HANDLER == 0) ? S.n_slow : 0u;
        if (ns) {
            for (uint32_t s = warp; s < ns; s += WARPS) {
                const uint32_t ks = S.slow_list[s];
                const uint32_t len = M.len[ks];
                TaskRec r2 = S.rec[ks];
                bool done = false;
                if (r2.mode == OM_STR_PAR) {                               // canonical frame, body needs transcoding
                    uint32_t ol;
                    const uint32_t n = len - FRAME_PRE_LEN - FRAME_SUF_LEN;
                    uint32_t* info = s < D2_ESC_SLOTS ? S.esc_info[s] : nullptr;
                    const bool ok = staged ? esc_measure(sbuf + M.soff[ks] + FRAME_PRE_LEN, n, lane, &ol, info)
                                           : esc_measure(a.payload + M.goff[ks] + FRAME_PRE_LEN, n, lane, &ol, info);
                    if (ok) { r2.has = 1; r2.src_off = FRAME_PRE_LEN; r2.src_len = n; r2.out_len = ol; done = true; }
                }
                if (!done && lane == 0) {
                    const uint8_t* p = staged ? (const uint8_t*)(sbuf + M.soff[ks]) : a.payload + M.goff[ks];
                    r2.mode = OM_NONE; r2.has = 0; r2.out_len = 0;
                    d2_parse_and_size<0>(p, len, r2, nullptr);
                    r2.ready = 1;
                }
                if (lane == 0) S.rec[ks] = r2;
            }
            __syncthreads();                                               // [2] slow tasks sized
        }

        // ---------------- compaction (ballot) + sizes (scan) + ONE cursor add, all in warp 0 ---------
        if (warp == 0) {
            const uint32_t l0 = S.rec[lane].out_len, l1 = S.rec[lane + 32].out_len;
            const uint32_t m0 = __ballot_sync(0xffffffffu, S.rec[lane].ready), m1 = __ballot_sync(0xffffffffu, S.rec[lane + 32].ready);
            const uint32_t e0 = warp_excl_scan(l0, lane);
            const uint32_t tot0 = __shfl_sync(0xffffffffu, e0 + l0, 31);
            const uint32_t e1 = tot0 + warp_excl_scan(l1, lane);
            const uint32_t tb = __shfl_sync(0xffffffffu, e1 + l1, 31);
            S.excl_bytes[lane] = e0; S.excl_bytes[lane + 32] = e1;
            const uint32_t below = (1u << lane) - 1u;
            S.excl_cnt[lane] = __popc(m0 & below); S.excl_cnt[lane + 32] = __popc(m0) + __popc(m1 & below);
            if (lane == 0) {
                const unsigned long long base = tb ? atomicAdd(&a.ctl->bytes, (unsigned long long)tb) : 0ull;
                S.base = base;
                if (base + tb > a.out_cap) { a.ctl->overflow = 1u; S.base = ~0ull; }
                S.n_slow = 0;                                              // for the next tile (touched again only after [4])
            }
        }
        __syncthreads();                                                   // [3] offsets known
        const unsigned long long base_bytes = S.base;
        const bool fits = base_bytes != ~0ull;
        const uint32_t base_cnt = M.base_cnt;

        // ---------------- phase B: G threads per task ------------------------------------------------
        if (mine) {
            if (ns) rec = S.rec[k];                                        // the cooperative pass may have rewritten it
            const uint64_t ob = base_bytes + S.excl_bytes[k];
            if (sub == 0) {
                const uint32_t slot = (uint32_t)((a.first_task + t0 + k) & a.slot_mask);
                const uint32_t j = base_cnt + S.excl_cnt[k];
                a.out_off[j] = fits ? ob : 0; a.out_len[j] = rec.out_len; a.out_ids[j] = __ldg(a.ids + slot); a.out_status[j] = rec.status; a.out_has[j] = rec.has;
            }
            if (rec.has && fits) {
                if (rec.mode == OM_COPY) {
                    if (staged) group_copy<G>(a.out_payload + ob, sbuf + M.soff[k] + rec.src_off, rec.src_len, sub);
                    else        group_copy<G>(a.out_payload + ob, a.payload + M.goff[k] + rec.src_off, rec.src_len, sub);
                } else if (sub == 0 && rec.mode != OM_STR_PAR) {
                    const uint8_t* p = staged ? (const uint8_t*)(sbuf + M.soff[k]) : a.payload + M.goff[k];
                    d2_phase_b_task<HANDLER>(p, rec, a.out_payload + ob);
                }
            }
        }
        // ---------------- phase B, cooperative: transcode the escaped strings -----------------------
        if (ns && fits) {
            for (uint32_t s = warp; s < ns; s += WARPS) {
                const uint32_t ks = S.slow_list[s];
                const TaskRec r2 = S.rec[ks];
                if (r2.mode != OM_STR_PAR || !r2.has) continue;
                const uint32_t* info = s < D2_ESC_SLOTS ? S.esc_info[s] : nullptr;
                uint8_t* o = a.out_payload + base_bytes + S.excl_bytes[ks];
                if (staged) esc_emit(sbuf + M.soff[ks] + r2.src_off, r2.src_len, lane, o, info);
                else        esc_emit(a.payload + M.goff[ks] + r2.src_off, r2.src_len, lane, o, info);
            }
        }
        __syncthreads();                                                   // [4] stage buffer, records, metadata free again
    }
}

}  // namespace b9
