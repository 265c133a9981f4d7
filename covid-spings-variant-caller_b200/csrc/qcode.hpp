// 2-bit quality codes (lvc_batch::qual_bits == 2): the integer transforms the generation-5 tiled kernel applies while it
// stages a quality-code batch, as host + device functions so that tests/test_qcode_cpu.py can check them on the CPU
// against a base-by-base restatement, and the host-side packer behind lvc_pack_quality_codes.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LVC_HD __host__ __device__ __forceinline__
#else
#define LVC_HD inline
#endif

namespace lvc {

// w = 16 codes, base i of the group in bits 2i.  Result: bit 2i set iff code i == c.
LVC_HD uint32_t qc_eq_flags(uint32_t w, uint32_t c) {
    const uint32_t x = w ^ (c * 0x55555555u);
    return ~(x | (x >> 1)) & 0x55555555u;
}
// v = 8 flags in the even bits of the low 16 bits (higher bits ignored).  Result: nibble i = 0xF iff flag i is set.
LVC_HD uint32_t qc_spread8(uint32_t v) {
    v &= 0xFFFFu;
    v = (v | (v << 8)) & 0x00FF00FFu;
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x11111111u;
    return v * 15u;
}
// the 16 staged keys of one group: (s0, s1) = base nibbles of bases 0..7 / 8..15 in little-endian nibble order,
// w = their codes.  key = the base's nibble if its code is the primary one, else 0.
LVC_HD void qc_keys16(uint32_t s0, uint32_t s1, uint32_t w, uint32_t pcode, uint32_t& k0, uint32_t& k1) {
    const uint32_t eq = qc_eq_flags(w, pcode);
    k0 = s0 & qc_spread8(eq);
    k1 = s1 & qc_spread8(eq >> 16);
}
// 2-bit base codes (lvc_batch seq form 2): w = 16 codes (A,C,G,T = 0..3), base i of the group in bits 2i.  Result: the
// bases' one-hot BAM nibbles (A,C,G,T = 1,2,4,8) in little-endian nibble order, bases 0..7 in s0 and 8..15 in s1 --
// what the staging loop of the tiled kernel makes of the 4-bit form with two byte swaps.
LVC_HD uint32_t b2_onehot8(uint32_t v) {
    v &= 0xFFFFu;
    v = (v | (v << 8)) & 0x00FF00FFu;
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u;                 // nibble i = code i
    const uint32_t lo = v & 0x11111111u, hi = (v >> 1) & 0x11111111u;
    const uint32_t x = 0x11111111u + lo;              // 1 << (code & 1)
    const uint32_t hm = hi * 15u;                     // nibbles whose code has bit 1 set: shift by two more
    return (x & ~hm) | ((x << 2) & hm);
}
LVC_HD void b2_onehot16(uint32_t w, uint32_t& s0, uint32_t& s1) {
    s0 = b2_onehot8(w);
    s1 = b2_onehot8(w >> 16);
}
// flags (bit 2i) of the bases of a group whose code is in the set `cold` (bit c = code c)
LVC_HD uint32_t qc_cold_flags(uint32_t w, uint32_t cold) {
    uint32_t f = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t c = 0; c < 4; ++c)
        if ((cold >> c) & 1u) f |= qc_eq_flags(w, c);
    return f;
}

}  // namespace lvc

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>
namespace lvc {
// Byte qualities -> 2-bit codes (lvc_pack_quality_codes in include/lvc.h).  Returns the number of distinct values among the
// qualities of the l_qseq bases (sum of the query-consuming CIGAR ops) of the reads with keep bit0 set -- every byte when
// the per-read arrays are null --, 0 if there are more than four.  Every byte is coded (a value outside the dictionary,
// possible only in a dropped read or in the pad byte of an odd-length read, becomes code 0).
inline int pack_quality_codes(const uint8_t* qual, uint64_t nq, uint32_t n_reads, const uint8_t* keep,
                              const uint64_t* seq_off, const uint32_t* cigar_off, const uint32_t* cigar, int n_threads,
                              uint8_t dict_out[4], uint8_t* codes_out) {
    const bool per_read = keep && seq_off && cigar_off && cigar && n_reads;
    n_threads = std::max(1, std::min(n_threads, 64));
    if (nq < (1u << 20)) n_threads = 1;
    std::vector<std::vector<uint8_t>> seen((size_t)n_threads, std::vector<uint8_t>(256, 0));
    auto run = [&](auto&& fn) {
        std::vector<std::thread> th;
        for (int t = 1; t < n_threads; ++t) th.emplace_back(fn, t);
        fn(0);
        for (auto& x : th) x.join();
    };
    run([&](int t) {
        uint8_t* sn = seen[(size_t)t].data();
        if (per_read) {
            const uint64_t r0 = (uint64_t)n_reads * t / n_threads, r1 = (uint64_t)n_reads * (t + 1) / n_threads;
            for (uint64_t i = r0; i < r1; ++i)
                if (keep[i] & 1u) {
                    uint64_t lq = 0;
                    for (uint32_t k = cigar_off[i]; k < cigar_off[i + 1]; ++k) {
                        const uint32_t op = cigar[k] & 15u;
                        if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) lq += cigar[k] >> 4;
                    }
                    const uint64_t x1 = std::min<uint64_t>(seq_off[i] + lq, seq_off[i + 1]);
                    for (uint64_t x = seq_off[i]; x < x1; ++x) sn[qual[x]] = 1;
                }
        } else {
            const uint64_t a = nq * t / n_threads, b = nq * (t + 1) / n_threads;
            for (uint64_t x = a; x < b; ++x) sn[qual[x]] = 1;
        }
    });
    uint8_t dict[4] = {0, 0, 0, 0};
    int nd = 0;
    for (int v = 0; v < 256; ++v) {
        bool any = false;
        for (int t = 0; t < n_threads; ++t) any |= seen[(size_t)t][(size_t)v] != 0;
        if (any) { if (nd == 4) return 0; dict[nd++] = (uint8_t)v; }
    }
    if (nd == 0) nd = 1;                                   // no base at all: one entry, phred 0
    // two qualities -> four code bits per lookup
    std::vector<uint8_t> lut2(65536);
    {
        uint8_t lut[256];
        memset(lut, 0, sizeof lut);
        for (int c = 0; c < nd; ++c) lut[dict[c]] = (uint8_t)c;
        for (int hi = 0; hi < 256; ++hi)
            for (int lo = 0; lo < 256; ++lo) lut2[(size_t)((hi << 8) | lo)] = (uint8_t)(lut[lo] | (lut[hi] << 2));
    }
    const uint64_t n_out = (nq + 3) / 4, n_full = nq / 4;
    run([&](int t) {
        const uint64_t a = n_full * t / n_threads, b = n_full * (t + 1) / n_threads;
        for (uint64_t o = a; o < b; ++o) {
            uint16_t p0, p1;
            memcpy(&p0, qual + 4 * o, 2);                 // little endian: the first quality is the low byte
            memcpy(&p1, qual + 4 * o + 2, 2);
            codes_out[o] = (uint8_t)(lut2[p0] | (lut2[p1] << 4));
        }
    });
    if (n_out > n_full) {
        uint8_t v = 0;
        for (uint64_t x = 4 * n_full; x < nq; ++x) v |= (uint8_t)((lut2[qual[x]] & 3u) << (2 * (x & 3)));
        codes_out[n_full] = v;
    }
    memcpy(dict_out, dict, 4);
    return nd;
}

// 4-bit BAM base codes -> 2-bit codes (lvc_pack_base_codes in include/lvc.h).  Returns 1 after writing
// codes_out[(nq + 3) / 4] if every base that can reach the tables is A, C, G or T: a base of a read the admission dropped,
// a base whose quality is below `min_bq` (the pileup never shows it: live_variant_caller.py:58) and the pad nibble of an
// odd-length read may be anything and get code 0.  Returns 0 (codes_out unspecified) otherwise: the batch keeps its nibbles.
// `qual`: one phred byte per base (the byte form; the quality CODES of the batch are made from the same bytes).
inline int pack_base_codes(const uint8_t* seq4, const uint8_t* qual, uint64_t nq, uint32_t n_reads, const uint8_t* keep,
                           const uint64_t* seq_off, const uint32_t* cigar_off, const uint32_t* cigar, int min_bq, int n_threads,
                           uint8_t* codes_out) {
    if (!(seq4 && qual && keep && seq_off && cigar_off && cigar && codes_out)) return 0;
    (void)nq;
    n_threads = std::max(1, std::min(n_threads, 64));
    if (n_reads < (1u << 14)) n_threads = 1;
    // byte of two nibbles (first base in the HIGH nibble) -> two codes (first base in the LOW bits); 0x80: not both A/C/G/T
    uint8_t pair[256];
    auto is_acgt = [](int nib) { return nib == 1 || nib == 2 || nib == 4 || nib == 8; };
    for (int v = 0; v < 256; ++v) {
        auto code = [](int nib) { return nib == 2 ? 1 : nib == 4 ? 2 : nib == 8 ? 3 : 0; };
        pair[v] = (uint8_t)((is_acgt(v >> 4) && is_acgt(v & 15) ? 0 : 0x80) | code(v >> 4) | (code(v & 15) << 2));
    }
    // four bases (two bytes) per lookup: low byte = their codes, bit 8 = one of them is not A/C/G/T
    static const std::vector<uint16_t> quad = [&] {
        std::vector<uint16_t> q(65536);
        for (int v = 0; v < 65536; ++v)
            q[(size_t)v] = (uint16_t)((pair[v & 255] & 15u) | ((pair[v >> 8] & 15u) << 4) | (((pair[v & 255] | pair[v >> 8]) & 0x80u) << 1));
        return q;
    }();
    // one read: its bases x in [seq_off[i], seq_off[i + 1]) (even bounds).  An output byte holds the bases 4k .. 4k + 3:
    // its low half is written with `=`, its high half OR-ed in afterwards (by the same thread, see the cuts)
    auto do_read = [&](uint64_t i) -> bool {
        const uint64_t x0 = seq_off[i], x1 = seq_off[i + 1];
        uint64_t lq = 0;
        for (uint32_t k = cigar_off[i]; k < cigar_off[i + 1]; ++k) {
            const uint32_t op = cigar[k] & 15u;
            if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) lq += cigar[k] >> 4;
        }
        const bool live = (keep[i] & 1u) != 0;
        const uint64_t xe = std::min<uint64_t>(x0 + lq, x1);       // the pad nibble of an odd-length read is not a base
        auto two = [&](uint64_t x) -> bool {                       // the bases x, x + 1 (one byte of nibbles)
            const uint8_t by = seq4[x >> 1], p = pair[by];
            if ((p & 0x80) && live) {
                if ((!is_acgt(by >> 4) && x < xe && (int)qual[x] >= min_bq) ||
                    (!is_acgt(by & 15) && x + 1 < xe && (int)qual[x + 1] >= min_bq)) return false;
            }
            const uint8_t v = (uint8_t)((p & 15u) << (2 * (x & 3)));
            if (x & 2) codes_out[x >> 2] |= v; else codes_out[x >> 2] = v;
            return true;
        };
        uint64_t x = x0;
        if ((x & 2) && x < x1) { if (!two(x)) return false; x += 2; }
        for (; x + 4 <= x1; x += 4) {
            uint16_t w;
            memcpy(&w, seq4 + (x >> 1), 2);                        // little endian: bases x, x + 1 in the low byte
            const uint16_t q = quad[w];
            if ((q & 0x100u) && live) { if (!two(x) || !two(x + 2)) return false; }
            else codes_out[x >> 2] = (uint8_t)q;
        }
        if (x < x1 && !two(x)) return false;
        return true;
    };
    // threads take runs of reads that start where seq_off is a multiple of four: no output byte is shared between them
    std::vector<uint64_t> cut((size_t)n_threads + 1);
    for (int t = 0; t <= n_threads; ++t) {
        uint64_t r = (uint64_t)n_reads * t / n_threads;
        while (t > 0 && r < n_reads && (seq_off[r] & 3u)) ++r;
        cut[(size_t)t] = t == n_threads ? n_reads : r;
    }
    std::vector<int> bad((size_t)n_threads, 0);
    auto work = [&](int t) {
        for (uint64_t i = cut[(size_t)t]; i < cut[(size_t)t + 1]; ++i)
            if (!do_read(i)) { bad[(size_t)t] = 1; return; }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    for (int t = 0; t < n_threads; ++t) if (bad[(size_t)t]) return 0;
    return 1;
}
}  // namespace lvc
