// Raw DEFLATE (RFC 1951) decoder for BGZF blocks -- host code of the native ingest (ingest.hpp).
//
// process_bam(BAM on disk) is bound by the inflate of the file's BGZF blocks (the reference gets the same bytes through
// pysam -> htslib -> zlib, live_variant_caller.py:55-60).  zlib's inflate() is a resumable state machine that decodes one
// symbol per table walk from a byte-wise refilled 32-bit accumulator; a BGZF block is a complete stream of <= 64 KiB whose
// input and output are both in memory, so none of that machinery is needed.  This decoder keeps a 64-bit bit buffer refilled
// with one unaligned 8-byte load, resolves literal / length codes of <= 11 bits (distance codes of <= 8 bits) with ONE table
// lookup whose entry already carries the base value and the number of extra bits, emits up to three literals per refill and
// copies matches eight bytes at a time.  Measured on the config-2 BAM's blocks (one host thread): 2.5-3x zlib 1.3.
//
// Safety: every access is bounded.  The hot loop runs only while at least 16 readable bytes follow the input position and
// 282 writable bytes follow the output position (a BGZF block is followed by its 8-byte trailer, so the bound costs
// nothing); the rest of a block goes through a loop that reads and writes byte by byte.  Output never exceeds the caller's
// buffer: neighbouring BGZF blocks are inflated by other threads into the same array.  A stream that consumes bits past its
// end, refers to data before the start of the block (BGZF blocks are independent), uses an over-subscribed or unusable
// Huffman code, or does not produce exactly the expected number of bytes is rejected; the caller checks the CRC-32 of
// every block on top of that (as htslib does).
#pragma once
#include <cstdint>
#include <cstring>
#include <cstddef>

namespace lvc_inflate {

constexpr int kLitBits = 11, kDistBits = 8, kPreBits = 7;
constexpr int kLitEnough = 2342, kDistEnough = 402;       // zlib's `enough 288 11 15` / `enough 32 8 15`
// table entry: [31:16] payload (literal, base length, base distance, or subtable offset), [15:12] flags,
//              [11:8] code length after the table's index bits were taken care of, [4:0] bits to consume (code + extra)
constexpr uint32_t kLit = 0x1000u, kEob = 0x2000u, kSub = 0x4000u, kBad = 0x8000u;

struct Tables {
    uint32_t lit[kLitEnough];
    uint32_t dist[kDistEnough];
    uint32_t pre[1 << kPreBits];
    bool fixed_loaded = false;       // lit / dist currently hold the fixed code of BTYPE 1
};

static inline uint32_t rev_bits(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

// what a decoded symbol means, without its code length: payload << 16 | flags | extra bits << 20 (moved by build())
static inline bool litlen_symbol(int sym, uint32_t& payload, uint32_t& flags, int& extra) {
    static const uint16_t base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t ext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    extra = 0;
    if (sym < 256) { payload = (uint32_t)sym; flags = kLit; return true; }
    if (sym == 256) { payload = 0; flags = kEob; return true; }
    if (sym <= 285) { payload = base[sym - 257]; flags = 0; extra = ext[sym - 257]; return true; }
    payload = 0; flags = kBad; return true;                           // 286, 287: may carry a length, must not be used
}
static inline bool dist_symbol(int sym, uint32_t& payload, uint32_t& flags, int& extra) {
    static const uint16_t base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
                                      4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t ext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    extra = 0;
    if (sym < 30) { payload = base[sym]; flags = 0; extra = ext[sym]; return true; }
    payload = 0; flags = kBad; return true;                           // 30, 31
}
static inline bool pre_symbol(int sym, uint32_t& payload, uint32_t& flags, int& extra) {
    payload = (uint32_t)sym; flags = kLit; extra = 0;
    return true;
}

// Canonical Huffman code -> lookup table with `tbits` index bits and second-level tables for longer codes.
// false: over-subscribed code, or more table space than `cap` (cannot happen for a complete code).  An incomplete code is
// accepted (a distance code with a single symbol is legal, and zlib-compatible streams contain nothing else incomplete);
// bit patterns it does not define decode to kBad entries.
template <class SymFn>
static bool build(uint32_t* table, int tbits, int cap, const uint8_t* lens, int nsyms, SymFn sym_fn) {
    int count[16] = {0};
    for (int s = 0; s < nsyms; ++s) count[lens[s]]++;
    int left = 1;
    for (int l = 1; l <= 15; ++l) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return false;
    }
    int offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + count[l];
    uint16_t sorted[288];
    for (int s = 0; s < nsyms; ++s) if (lens[s]) sorted[offs[lens[s]]++] = (uint16_t)s;
    const int n_coded = offs[15];
    const int tsize = 1 << tbits;
    for (int i = 0; i < tsize; ++i) table[i] = kBad;
    // codes longer than the index: the longest code of every index prefix sizes that prefix's second-level table
    uint8_t maxlen[1 << kLitBits];
    bool any_long = false;
    {
        uint32_t code = 0;
        int k = 0;
        for (int l = 1; l <= 15; ++l) {
            for (int c = 0; c < count[l]; ++c, ++k, ++code)
                if (l > tbits) {
                    if (!any_long) { memset(maxlen, 0, (size_t)tsize); any_long = true; }
                    maxlen[code >> (l - tbits)] = (uint8_t)l;            // lengths only grow along the canonical order
                }
            code <<= 1;
        }
    }
    int next_free = tsize;
    uint32_t code = 0;
    int k = 0;
    for (int l = 1; l <= 15; ++l) {
        for (int c = 0; c < count[l]; ++c, ++k, ++code) {
            uint32_t payload, flags;
            int extra;
            sym_fn((int)sorted[k], payload, flags, extra);
            if (l <= tbits) {
                const uint32_t e = (payload << 16) | flags | ((uint32_t)l << 8) | (uint32_t)(l + extra);
                for (uint32_t j = rev_bits(code, l); j < (uint32_t)tsize; j += 1u << l) table[j] = e;
            } else {
                const uint32_t prefix = code >> (l - tbits);
                const uint32_t pidx = rev_bits(prefix, tbits);
                const int sbits = (int)maxlen[prefix] - tbits;
                if (!(table[pidx] & kSub)) {
                    if (next_free + (1 << sbits) > cap) return false;
                    table[pidx] = ((uint32_t)next_free << 16) | kSub | (uint32_t)sbits;
                    for (int i = 0; i < (1 << sbits); ++i) table[next_free + i] = kBad;
                    next_free += 1 << sbits;
                }
                const uint32_t sub0 = table[pidx] >> 16;
                const int rl = l - tbits;                                   // bits of the code left for the second level
                const uint32_t e = (payload << 16) | flags | ((uint32_t)rl << 8) | (uint32_t)(rl + extra);
                for (uint32_t j = rev_bits(code & ((1u << rl) - 1u), rl); j < (1u << sbits); j += 1u << rl) table[sub0 + j] = e;
            }
        }
        code <<= 1;
    }
    (void)n_coded;
    return true;
}

static inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline void store64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }

// Inflate one complete raw DEFLATE stream of in_len bytes into exactly out_len bytes.  `in_slack` bytes after the stream
// are readable (never interpreted: a stream that needs them is rejected); the hot loop stops 16 bytes before their end.
// Returns true iff the stream is well formed, ends with its final block and produced exactly out_len bytes.
static bool inflate_block(Tables& T, const uint8_t* in0, size_t in_len, size_t in_slack, uint8_t* out0, size_t out_len) {
    const uint8_t* in = in0;
    const uint8_t* const in_end = in0 + in_len;
    uint8_t* out = out0;
    uint8_t* const out_end = out0 + out_len;
    uint64_t bitbuf = 0;
    uint32_t bitcnt = 0;             // valid bits in bitbuf
    uint32_t phantom = 0;            // zero bytes appended after the end of the input (never to be consumed)
    // the hot loop issues up to two 8-byte loads per iteration, the second at most 7 bytes behind the position it checked
    const ptrdiff_t hot_room = (ptrdiff_t)in_len + (ptrdiff_t)(in_slack < 64 ? in_slack : 64) - 16;
    const uint8_t* const in_hot = hot_room >= 0 ? in0 + hot_room : nullptr;   // last position an iteration may start at

    // byte-wise refill to >= 56 bits that never reads past the stream
    auto refill = [&]() {
        while (bitcnt <= 56) {
            if (in < in_end) bitbuf |= (uint64_t)*in++ << bitcnt;
            else ++phantom;
            bitcnt += 8;
        }
    };
    auto overran = [&]() { return phantom * 8u > bitcnt; };
    auto take = [&](uint32_t n) { const uint32_t v = (uint32_t)(bitbuf & ((1ull << n) - 1ull)); bitbuf >>= n; bitcnt -= n; return v; };

    for (;;) {
        refill();
        if (overran()) return false;
        const uint32_t bfinal = take(1), btype = take(2);
        if (btype == 0) {
            // stored: drop the rest of the byte, hand the unread whole bytes back
            take(bitcnt & 7u);
            if (overran()) return false;
            const uint32_t back = bitcnt >> 3;
            in -= (back - phantom);
            bitbuf = 0; bitcnt = 0; phantom = 0;
            if (in_end - in < 4) return false;
            const uint32_t len = (uint32_t)in[0] | ((uint32_t)in[1] << 8), nlen = (uint32_t)in[2] | ((uint32_t)in[3] << 8);
            in += 4;
            if ((len ^ nlen) != 0xFFFFu) return false;
            if ((size_t)(in_end - in) < len || (size_t)(out_end - out) < len) return false;
            memcpy(out, in, len);
            in += len; out += len;
        } else if (btype == 1 || btype == 2) {
            if (btype == 1) {
                if (!T.fixed_loaded) {
                    uint8_t lens[288 + 32];
                    for (int s = 0; s < 144; ++s) lens[s] = 8;
                    for (int s = 144; s < 256; ++s) lens[s] = 9;
                    for (int s = 256; s < 280; ++s) lens[s] = 7;
                    for (int s = 280; s < 288; ++s) lens[s] = 8;
                    for (int s = 0; s < 32; ++s) lens[288 + s] = 5;
                    if (!build(T.lit, kLitBits, kLitEnough, lens, 288, litlen_symbol)) return false;
                    if (!build(T.dist, kDistBits, kDistEnough, lens + 288, 32, dist_symbol)) return false;
                    T.fixed_loaded = true;
                }
            } else {
                T.fixed_loaded = false;
                const uint32_t hlit = take(5) + 257u, hdist = take(5) + 1u, hclen = take(4) + 4u;
                if (hlit > 286u || hdist > 30u) return false;
                static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                uint8_t plen[19] = {0};
                for (uint32_t i = 0; i < hclen; ++i) { if (bitcnt < 3) refill(); plen[order[i]] = (uint8_t)take(3); }
                if (!build(T.pre, kPreBits, 1 << kPreBits, plen, 19, pre_symbol)) return false;
                uint8_t lens[286 + 30 + 138];
                uint32_t n = 0;
                const uint32_t total = hlit + hdist;
                while (n < total) {
                    refill();
                    if (overran()) return false;
                    const uint32_t e = T.pre[bitbuf & ((1u << kPreBits) - 1u)];
                    if (e & kBad) return false;
                    take(e & 31u);
                    const uint32_t sym = e >> 16;
                    if (sym < 16) lens[n++] = (uint8_t)sym;
                    else if (sym == 16) {
                        if (n == 0) return false;
                        const uint32_t rep = 3u + take(2);
                        const uint8_t v = lens[n - 1];
                        for (uint32_t i = 0; i < rep; ++i) lens[n++] = v;
                    } else {
                        const uint32_t rep = sym == 17 ? 3u + take(3) : 11u + take(7);
                        for (uint32_t i = 0; i < rep; ++i) lens[n++] = 0;
                    }
                }
                if (n != total) return false;                         // a repeat ran over the end of the code lengths
                if (overran()) return false;
                if (lens[256] == 0) return false;                     // no end-of-block code
                uint8_t dl[32] = {0};
                memcpy(dl, lens + hlit, hdist);
                if (!build(T.lit, kLitBits, kLitEnough, lens, (int)hlit, litlen_symbol)) return false;
                if (!build(T.dist, kDistBits, kDistEnough, dl, (int)hdist, dist_symbol)) return false;
            }
            // ---- symbols of the block
            bool eob = false;
            while (!eob) {
                // hot loop: 8 readable bytes at `in`, room for three literals or one match plus the copy's overshoot
                if (phantom == 0 && in_hot && in <= in_hot && (size_t)(out_end - out) >= 282) {
                    uint8_t* const out_hot = out_end - 282;
                    for (;;) {
                        if (in > in_hot || out > out_hot) break;
                        bitbuf |= load64(in) << bitcnt;
                        in += (63u - bitcnt) >> 3;
                        bitcnt |= 56u;
                        uint32_t e = T.lit[bitbuf & ((1u << kLitBits) - 1u)];
                        if (e & kLit) {
                            *out++ = (uint8_t)(e >> 16); bitbuf >>= (e & 31u); bitcnt -= (e & 31u);
                            e = T.lit[bitbuf & ((1u << kLitBits) - 1u)];
                            if (e & kLit) {
                                *out++ = (uint8_t)(e >> 16); bitbuf >>= (e & 31u); bitcnt -= (e & 31u);
                                e = T.lit[bitbuf & ((1u << kLitBits) - 1u)];
                                if (e & kLit) {
                                    *out++ = (uint8_t)(e >> 16); bitbuf >>= (e & 31u); bitcnt -= (e & 31u);
                                    continue;
                                }
                            }
                            // <= 22 bits gone: top up (the low bits, which `e` was looked up with, do not change)
                            bitbuf |= load64(in) << bitcnt;
                            in += (63u - bitcnt) >> 3;
                            bitcnt |= 56u;
                        }
                        if (e & kSub) {
                            bitbuf >>= kLitBits; bitcnt -= kLitBits;
                            e = T.lit[(e >> 16) + (uint32_t)(bitbuf & ((1u << (e & 31u)) - 1u))];
                            if (e & kLit) { *out++ = (uint8_t)(e >> 16); bitbuf >>= (e & 31u); bitcnt -= (e & 31u); continue; }
                        }
                        if (e & (kEob | kBad)) {
                            if (e & kBad) return false;
                            bitbuf >>= (e & 31u); bitcnt -= (e & 31u);
                            eob = true;
                            break;
                        }
                        // length: base + extra bits (taken from the bits as they were before the code was dropped)
                        const uint32_t cl = (e >> 8) & 15u, tot = e & 31u;
                        const uint32_t len = (e >> 16) + ((uint32_t)(bitbuf >> cl) & ((1u << (tot - cl)) - 1u));
                        bitbuf >>= tot; bitcnt -= tot;
                        uint32_t d = T.dist[bitbuf & ((1u << kDistBits) - 1u)];
                        if (d & kSub) {
                            bitbuf >>= kDistBits; bitcnt -= kDistBits;
                            d = T.dist[(d >> 16) + (uint32_t)(bitbuf & ((1u << (d & 31u)) - 1u))];
                        }
                        if (d & kBad) return false;
                        const uint32_t dcl = (d >> 8) & 15u, dtot = d & 31u;
                        const uint32_t dist = (d >> 16) + ((uint32_t)(bitbuf >> dcl) & ((1u << (dtot - dcl)) - 1u));
                        bitbuf >>= dtot; bitcnt -= dtot;
                        if (dist > (size_t)(out - out0)) return false;          // before the start of the block
                        const uint8_t* src = out - dist;
                        uint8_t* dst = out;
                        out += len;
                        if (dist >= 8) {
                            // sixteen bytes without a test (three matches in four are no longer: the length of a match is
                            // what a branch predictor cannot know), then eight at a time; the last store may run up to 15
                            // bytes past the match (room was checked)
                            store64(dst, load64(src));
                            store64(dst + 8, load64(src + 8));
                            if (len > 16) {
                                dst += 16; src += 16;
                                do { store64(dst, load64(src)); dst += 8; src += 8; } while (dst < out);
                            }
                        } else if (dist == 1) {
                            const uint64_t v = 0x0101010101010101ull * (uint64_t)*src;
                            do { store64(dst, v); dst += 8; } while (dst < out);
                        } else {
                            // period 2..7: every 8-byte copy yields `dist` new bytes of the pattern (and some it overwrites next)
                            do { store64(dst, load64(src)); dst += dist; src += dist; } while (dst < out);
                        }
                    }
                    if (in > in_end) {
                        // the hot loop loaded bytes that follow the stream: they count as appended zeros from here on
                        // (their bits sit at the top of the buffer; if the stream needs them it is rejected)
                        const uint32_t over = (uint32_t)(in - in_end);
                        if (over * 8u > bitcnt) return false;
                        bitbuf &= (bitcnt - over * 8u) >= 64u ? ~0ull : ((1ull << (bitcnt - over * 8u)) - 1ull);
                        phantom = over;
                        in = in_end;
                    }
                    if (eob) break;
                }
                // careful loop: one symbol, exact bounds
                refill();
                uint32_t e = T.lit[bitbuf & ((1u << kLitBits) - 1u)];
                if (e & kSub) {
                    take(kLitBits);
                    e = T.lit[(e >> 16) + (uint32_t)(bitbuf & ((1u << (e & 31u)) - 1u))];
                }
                if (e & kBad) return false;
                if (e & kLit) {
                    take(e & 31u);
                    if (overran() || out >= out_end) return false;
                    *out++ = (uint8_t)(e >> 16);
                    continue;
                }
                if (e & kEob) {
                    take(e & 31u);
                    if (overran()) return false;
                    eob = true;
                    break;
                }
                const uint32_t cl = (e >> 8) & 15u, tot = e & 31u;
                const uint32_t len = (e >> 16) + ((uint32_t)(bitbuf >> cl) & ((1u << (tot - cl)) - 1u));
                take(tot);
                refill();
                uint32_t d = T.dist[bitbuf & ((1u << kDistBits) - 1u)];
                if (d & kSub) {
                    take(kDistBits);
                    d = T.dist[(d >> 16) + (uint32_t)(bitbuf & ((1u << (d & 31u)) - 1u))];
                }
                if (d & kBad) return false;
                const uint32_t dcl = (d >> 8) & 15u, dtot = d & 31u;
                const uint32_t dist = (d >> 16) + ((uint32_t)(bitbuf >> dcl) & ((1u << (dtot - dcl)) - 1u));
                take(dtot);
                if (overran()) return false;
                if (dist > (size_t)(out - out0) || len > (size_t)(out_end - out)) return false;
                const uint8_t* src = out - dist;
                for (uint32_t i = 0; i < len; ++i) out[i] = src[i];
                out += len;
            }
        } else return false;                                          // BTYPE 3
        if (bfinal) break;
    }
    // whole stream consumed without touching anything behind it, and exactly the announced size produced
    if (overran()) return false;
    return out == out_end;
}

}  // namespace lvc_inflate
