// vadd_f32's thread-per-task fast path (used by drain3_kernel<vadd_f32>, drain2.cuh). Plain C++ plus three
// byte-shuffle intrinsics: tests/host_shim/host_parse.cpp builds it for the host (with the intrinsics written
// out) and tests/test_device_parser_on_host.py runs it against the oracle on a CPU.
#pragma once
#include <stdint.h>
#include "handler_seq.cuh"

namespace b9 {

// frame constants of the SDK's put payload (also in drain_kernel.cuh as byte arrays): 11 bytes before / 17 after the string body
constexpr uint32_t VADD_FRAME_PRE_LEN = 11, VADD_FRAME_SUF_LEN = 17;

// 4 bytes at an arbitrary address: two aligned words + funnel shift
__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t* __restrict__ p) {
    const uint32_t a = (uint32_t)((uintptr_t)p & 3u);
    const uint32_t* w = (const uint32_t*)(p - a);
    return __funnelshift_r(w[0], w[1], a * 8u);
}

// ------------------------------------------------------------------ vadd_f32, thread per task, in place in the stage buffer
// The common case (canonical frame, plain base64 body, two equal fp32 vectors) is computed without a
// data-dependent branch: characters go through a 256-entry table in shared memory (0..63, 0xFF = not
// in the alphabet), three floats (16 characters of a, 16..20 of b) per step, and the result text is
// written over the bytes of `a` already consumed, 4-byte aligned, so that phase B is a plain copy.
// tab[0..255] = decode, tab[256..319] = encode.
__device__ __forceinline__ uint32_t b64_dec4(uint32_t x, const uint8_t* tab, uint32_t& bad) {
    const uint32_t v0 = tab[x & 0xFFu], v1 = tab[(x >> 8) & 0xFFu], v2 = tab[(x >> 16) & 0xFFu], v3 = tab[x >> 24];
    bad |= v0 | v1 | v2 | v3;
    return (v0 << 18) | (v1 << 12) | (v2 << 6) | v3;                   // stream bytes: bits 23..16, 15..8, 7..0
}
__device__ __forceinline__ uint32_t b64_enc4(uint32_t w, const uint8_t* tab) {
    return (uint32_t)tab[256 + ((w >> 18) & 63u)] | ((uint32_t)tab[256 + ((w >> 12) & 63u)] << 8) |
           ((uint32_t)tab[256 + ((w >> 6) & 63u)] << 16) | ((uint32_t)tab[256 + (w & 63u)] << 24);
}
// group g of the body; the padding characters of the last group read as 'A'
__device__ __forceinline__ uint32_t b64_group_padded(const uint8_t* body, uint32_t g, uint32_t last, uint32_t pad) {
    uint32_t x = ld_u32_unaligned(body + 4u * g);
    if (g == last && pad) x = pad == 1u ? ((x & 0x00FFFFFFu) | 0x41000000u) : ((x & 0x0000FFFFu) | 0x41410000u);
    return x;
}
// little-endian u32 at stream byte s (s..s+3) of the decoded body; used for the <= 2 floats after the last full step
__device__ __forceinline__ uint32_t b64_u32_at(const uint8_t* body, uint32_t s, uint32_t last, uint32_t pad, const uint8_t* tab, uint32_t& bad) {
    const uint32_t g = s / 3u, q = s - 3u * g;
    const uint32_t u0 = b64_dec4(b64_group_padded(body, g, last, pad), tab, bad);
    const uint32_t u1 = b64_dec4(b64_group_padded(body, g + 1u, last, pad), tab, bad);
    const uint32_t s0 = __byte_perm(u0, u1, 0x6012), s1 = __byte_perm(u1, 0u, 0x4401);      // stream bytes 0..3, 4..5
    return __funnelshift_r(s0, s1, 8u * q);
}
// returns 0 = not the common case, nothing touched; 1 = done (rec filled); 2 = stage bytes overwritten and
// a non-alphabet character found: the caller re-reads the task from the ring and takes the general path
__device__ __forceinline__ int vadd_fast(uint8_t* p, uint32_t len, const uint8_t* tab, TaskRec& rec) {
    if (len < VADD_FRAME_PRE_LEN + VADD_FRAME_SUF_LEN + 4u) return 0;
    const uint8_t* q = p + len - VADD_FRAME_SUF_LEN;
    bool bad_frame = ld_u32_unaligned(p) != 0x7261227Bu;
    bad_frame |= ld_u32_unaligned(p + 4) != 0x3A227367u;
    bad_frame |= (ld_u32_unaligned(p + 8) & 0x00FFFFFFu) != 0x00225B20u;
    bad_frame |= ld_u32_unaligned(q) != 0x202C5D22u;
    bad_frame |= ld_u32_unaligned(q + 4) != 0x61776B22u;
    bad_frame |= ld_u32_unaligned(q + 8) != 0x22736772u;
    bad_frame |= ld_u32_unaligned(q + 12) != 0x7D7B203Au;
    bad_frame |= q[16] != '}';
    const uint32_t L = len - VADD_FRAME_PRE_LEN - VADD_FRAME_SUF_LEN;
    if (bad_frame || (L & 3u)) return 0;
    uint8_t* const body = p + VADD_FRAME_PRE_LEN;
    const uint32_t pad = body[L - 1] == '=' ? (body[L - 2] == '=' ? 2u : 1u) : 0u;
    const uint32_t G = L >> 2, nbytes = G * 3u - pad;
    if (nbytes & 7u) return 0;
    const uint32_t n = nbytes >> 3;                                     // floats per vector (>= 1 here)
    const uint32_t nblk = n / 3u, rem = n - 3u * nblk, ph = rem;        // (4n) % 3 == n % 3
    const uint32_t gb0 = (4u * n) / 3u;
    uint8_t* const out = body - ((uintptr_t)body & 3u);                 // result characters start here (<= body)
    uint32_t bad = 0;
    uint32_t unext = nblk ? b64_dec4(ld_u32_unaligned(body + 4u * gb0), tab, bad) : 0u;
    for (uint32_t j = 0; j < nblk; ++j) {
        const uint8_t* ca = body + 16u * j;
        const uint8_t* cb = body + 4u * (gb0 + 4u * j);
        const uint32_t a0 = b64_dec4(ld_u32_unaligned(ca), tab, bad), a1 = b64_dec4(ld_u32_unaligned(ca + 4), tab, bad);
        const uint32_t a2 = b64_dec4(ld_u32_unaligned(ca + 8), tab, bad), a3 = b64_dec4(ld_u32_unaligned(ca + 12), tab, bad);
        const uint32_t u0 = unext;
        const uint32_t u1 = b64_dec4(ld_u32_unaligned(cb + 4), tab, bad), u2 = b64_dec4(ld_u32_unaligned(cb + 8), tab, bad);
        const uint32_t u3 = b64_dec4(ld_u32_unaligned(cb + 12), tab, bad);
        if (ph || j + 1u < nblk) unext = b64_dec4(ld_u32_unaligned(cb + 16), tab, bad);
        const uint32_t s0 = __byte_perm(u0, u1, 0x6012), s1 = __byte_perm(u1, u2, 0x5601), s2 = __byte_perm(u2, u3, 0x4560);
        const uint32_t s3 = __byte_perm(unext, 0u, 0x4012);
        const uint32_t z0 = vadd_bits(__byte_perm(a0, a1, 0x6012), __funnelshift_r(s0, s1, 8u * ph));
        const uint32_t z1 = vadd_bits(__byte_perm(a1, a2, 0x5601), __funnelshift_r(s1, s2, 8u * ph));
        const uint32_t z2 = vadd_bits(__byte_perm(a2, a3, 0x4560), __funnelshift_r(s2, s3, 8u * ph));
        uint32_t* o = (uint32_t*)(out + 16u * j);
        o[0] = b64_enc4(__byte_perm(z0, 0u, 0x4012), tab);              // bytes 0..2
        o[1] = b64_enc4(__byte_perm(z0, z1, 0x4345), tab);              // bytes 3..5: z0.b3, z1.b0, z1.b1
        o[2] = b64_enc4(__byte_perm(z1, z2, 0x4234), tab);              // bytes 6..8: z1.b2, z1.b3, z2.b0
        o[3] = b64_enc4(__byte_perm(z2, 0u, 0x4123), tab);              // bytes 9..11: z2.b1, z2.b2, z2.b3
    }
    uint32_t nchars = 16u * nblk;
    if (rem) {
        const uint32_t last = G - 1u;
        uint32_t z[2] = {0u, 0u};
        #pragma unroll
        for (uint32_t t = 0; t < 2; ++t) {
            if (t < rem) {
                const uint32_t i = 3u * nblk + t;
                z[t] = vadd_bits(b64_u32_at(body, 4u * i, last, pad, tab, bad), b64_u32_at(body, 4u * (n + i), last, pad, tab, bad));
            }
        }
        uint32_t* o = (uint32_t*)(out + nchars);
        o[0] = b64_enc4(__byte_perm(z[0], 0u, 0x4012), tab);
        if (rem == 1u) {
            o[1] = (b64_enc4(__byte_perm(z[0], 0u, 0x4344), tab) & 0x0000FFFFu) | 0x3D3D0000u;          // z0.b3 + "=="
            nchars += 8u;
        } else {
            o[1] = b64_enc4(__byte_perm(z[0], z[1], 0x4345), tab);
            o[2] = (b64_enc4(__byte_perm(z[1], 0u, 0x4234), tab) & 0x00FFFFFFu) | 0x3D000000u;           // z1.b2, z1.b3 + "="
            nchars += 12u;
        }
    }
    if (bad & 0x80u) return 2;
    out[-1] = '"'; out[nchars] = '"';
    rec.src_off = (uint32_t)(out - 1 - p); rec.src_len = nchars + 2u; rec.out_len = nchars + 2u; rec.mode = OM_COPY; rec.has = 1;
    return 1;
}

}  // namespace b9
