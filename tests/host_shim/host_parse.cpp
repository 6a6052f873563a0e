// TEST INFRASTRUCTURE: compiles the DEVICE's sequential path — the JSON parser (json_device.cuh), the
// handlers (handlers_device.cuh) and what a parsed payload means to each handler (handler_seq.cuh), all plain
// C++ behind CUDA qualifiers — for the host, so that its decisions and result bytes can be fuzzed against the
// oracle on a CPU (tests/test_device_parser_on_host.py). Nothing in the product links or loads this; the
// warp-cooperative fast paths (drain2.cuh) are GPU-only and covered by tests/test_gpu_parity.py.
#include <stdint.h>
#include <string.h>
#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__ static const
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __fadd_rn(float a, float b) { volatile float z = a + b; return z; }     // IEEE binary32, round to nearest even
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh & 31u)); }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {         // PRMT, default mode: result byte i = byte (nibble i of s) of {y:x}
    const uint64_t v = (((uint64_t)y) << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
}
#include "../../beta9_b200/csrc/handler_seq.cuh"
#include "../../beta9_b200/csrc/vadd_fast.cuh"
#define B9_WIRE_ENCODE_HELPERS_ONLY 1
#include "../../beta9_b200/csrc/wire_encode.cuh"

extern "C" int b9_host_parse(const uint8_t* p, uint32_t n, int http, uint32_t* out /* [12] */) {
    const b9::Parsed r = b9::parse_payload(p, n, http != 0);
    out[0] = r.status; out[1] = r.nargs; out[2] = r.kwargs_nonempty; out[3] = r.a0_kind;
    out[4] = r.a0_off; out[5] = r.a0_len; out[6] = r.a0_flags; out[7] = r.kw_merged;
    out[8] = r.args_off; out[9] = r.args_len; out[10] = r.kw_off; out[11] = r.kw_len;
    return 0;
}

// the whole sequential path for one task: parse -> handler (phase A) -> result bytes (phase B).
// Returns the result length (0 with *has == 0: no result bytes), or -1 if `cap` is too small.
extern "C" long b9_host_run(const uint8_t* p, uint32_t n, int http, int handler, uint8_t* status, uint8_t* has, uint8_t* out, uint32_t cap) {
    static uint32_t table[256];
    static bool have = false;
    if (!have) { for (uint32_t i = 0; i < 256; ++i) table[i] = b9::crc_table_entry(i); have = true; }
    const b9::Parsed pr = b9::parse_payload(p, n, http != 0);
    b9::TaskRec rec; memset(&rec, 0, sizeof rec); rec.ready = 1;
    b9::handler_phase_a(handler, p, pr, rec, table);
    *status = rec.status; *has = rec.has;
    if (!rec.has) return 0;
    if (rec.out_len > cap) return -1;
    if (rec.mode == b9::OM_COPY) memcpy(out, p + rec.src_off, rec.src_len);
    else b9::seq_emit(p, rec, out);
    return (long)rec.out_len;
}

// Go re-encoding of one validated JSON value (the args list / kwargs object of a TaskMessage): bytes written, or -1
// when the device encoder declines (unsorted or duplicate keys, non-integer numbers), -2 if `cap` is too small.
extern "C" long b9_host_go_transcode(const uint8_t* p, uint32_t n, uint8_t* out, uint32_t cap) {
    const int64_t need = b9::go_transcode(p, 0, n, nullptr);
    if (need < 0) return -1;
    if ((uint64_t)need > cap) return -2;
    return (long)b9::go_transcode(p, 0, n, out);
}
extern "C" long b9_host_rfc3339nano(long long unix_ns, uint8_t* out) { return (long)b9::rfc3339nano(unix_ns, out); }

// vadd_f32's in-place fast path on a private copy of the payload. Returns the result length when it decided
// the task (status COMPLETE, bytes in out), -1 when it left the task to the sequential path (return code 0 or 2).
extern "C" long b9_host_vadd_fast(const uint8_t* p, uint32_t n, uint8_t* out, uint32_t cap) {
    static uint8_t tab[320];
    static bool have = false;
    if (!have) { for (int i = 0; i < 320; ++i) tab[i] = i < 256 ? (uint8_t)b9::b64_val((uint8_t)i) : b9::b64_chr((uint32_t)i - 256u); have = true; }
    // the device buffer is 4-byte aligned storage with slack around the task: mimic it (and vary the task's alignment)
    static uint32_t store[(1u << 16) / 4];
    if (n + 64u > sizeof store) return -1;
    for (uint32_t al = 0; al < 4; ++al) {
        uint8_t* buf = (uint8_t*)store + 16 + al;
        memset(store, 0x5A, sizeof store);
        memcpy(buf, p, n);
        b9::TaskRec rec; memset(&rec, 0, sizeof rec);
        const int fr = b9::vadd_fast(buf, n, tab, rec);
        if (fr != 1) { if (al == 0) return -1; return -3; }                 // the decision must not depend on the alignment
        if (rec.out_len > cap || rec.mode != b9::OM_COPY || !rec.has) return -2;
        if (al == 0) memcpy(out, buf + rec.src_off, rec.src_len);
        else if (memcmp(out, buf + rec.src_off, rec.src_len) != 0) return -4;   // nor the bytes
        if (al == 3) return (long)rec.out_len;
    }
    return -5;
}
