// CRC-32 (IEEE 802.3, the one in every BGZF block trailer) by carry-less multiplication -- host code of the native ingest.
//
// htslib verifies the CRC of every BGZF block it inflates, and so does the ingest (ingest.hpp).  zlib 1.3's crc32() runs at
// 2-3 GB/s, a fifth of the time the library's own inflate needs for the same block; folding 64 bytes per step with
// PCLMULQDQ (Gopal et al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ Instruction", Intel 2009: fold by
// x^(512+-32) mod P, then 128 -> 64 -> 32 bits with a Barrett reduction) runs at memory speed.  Used where the CPU has the
// instruction AND the function reproduces zlib's result on a test pattern when it is first called; zlib's crc32() otherwise.
#pragma once
#include <zlib.h>
#include <cstdint>
#include <cstddef>
#include <initializer_list>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace lvc_crc {

#if defined(__x86_64__)
// internal (pre-inverted) state in, internal state out; len >= 64 and a multiple of 16
__attribute__((target("pclmul,sse4.1"))) static uint32_t fold_clmul(const uint8_t* buf, size_t len, uint32_t state) {
    alignas(16) static const uint64_t k1k2[2] = {0x0154442bd4ull, 0x01c6e41596ull};
    alignas(16) static const uint64_t k3k4[2] = {0x01751997d0ull, 0x00ccaa009eull};
    alignas(16) static const uint64_t k5k0[2] = {0x0163cd6124ull, 0x0000000000ull};
    alignas(16) static const uint64_t poly[2] = {0x01db710641ull, 0x01f7011641ull};
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
    x2 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
    x3 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
    x4 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)state));
    x0 = _mm_load_si128((const __m128i*)k1k2);
    buf += 64; len -= 64;
    while (len >= 64) {
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00); x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11); x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128((const __m128i*)(buf + 0x00)); y6 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
        y7 = _mm_loadu_si128((const __m128i*)(buf + 0x20)); y8 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5); x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7); x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64; len -= 64;
    }
    // four lanes -> one
    x0 = _mm_load_si128((const __m128i*)k3k4);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (len >= 16) {
        x2 = _mm_loadu_si128((const __m128i*)buf);
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16; len -= 16;
    }
    // 128 -> 64 bits
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_loadl_epi64((const __m128i*)k5k0);
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, x3);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    // Barrett reduction to 32 bits
    x0 = _mm_load_si128((const __m128i*)poly);
    x2 = _mm_and_si128(x1, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
    x2 = _mm_and_si128(x2, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}

static uint32_t crc32_with_clmul(const uint8_t* buf, size_t len) {
    uint32_t crc = 0;
    const size_t body = len >= 64 ? (len & ~(size_t)15) : 0;
    if (body) crc = ~fold_clmul(buf, body, ~crc);
    if (len > body) crc = (uint32_t)::crc32(crc, buf + body, (uInt)(len - body));
    return crc;
}

// the CPU has the instruction and the folding reproduces zlib's CRC on a pattern that exercises every stage
static bool clmul_usable() {
    static const bool ok = [] {
        if (!__builtin_cpu_supports("pclmul") || !__builtin_cpu_supports("sse4.1")) return false;
        uint8_t pat[1024 + 37];
        uint32_t s = 0x12345678u;
        for (size_t i = 0; i < sizeof pat; ++i) { s = s * 1664525u + 1013904223u; pat[i] = (uint8_t)(s >> 24); }
        for (size_t n : {(size_t)64, (size_t)80, (size_t)127, (size_t)128, (size_t)333, sizeof pat})
            if (crc32_with_clmul(pat, n) != (uint32_t)::crc32(::crc32(0L, Z_NULL, 0), pat, (uInt)n)) return false;
        return true;
    }();
    return ok;
}
#endif

// CRC-32 of buf[0, len) (len < 2^32: a BGZF block holds at most 64 KiB)
static inline uint32_t crc32_block(const uint8_t* buf, size_t len) {
#if defined(__x86_64__)
    if (clmul_usable()) return crc32_with_clmul(buf, len);
#endif
    return (uint32_t)::crc32(::crc32(0L, Z_NULL, 0), buf, (uInt)len);
}

}  // namespace lvc_crc
