// Host-side admission with htslib's mate-overlap handling (pysam pileup default ignore_overlaps=True).
//
// The reference calls pysam.AlignmentFile.pileup() without ignore_overlaps (live_variant_caller.py:56-60), so
// htslib's pileup engine runs with an overlap hash (bam_plp_init_overlaps): when the second read of a proper pair
// is pushed while the first one is still buffered, tweak_overlap_quality() REWRITES the base qualities of both
// mates at the reference positions they share, and everything downstream (pysam's base-quality filter, the phreds
// the reference appends, live_variant_caller.py:97-103) sees the rewritten values.  That rewrite is a pure function
// of the read stream, so it is done here, on the host, on the packed quality array, inside the same sequential pass
// that computes the max_depth admission mask (the hash is only consulted for reads that are admitted, and entries
// leave it when a read is swept from the buffer or dropped by max_depth -- SURVEY B4/B5).
//
// htslib's behaviour differs between releases and pysam is unpinned in the reference (requirements.txt:1), so two
// models are offered (restated from htslib sam.c; [EXT], parity unpinned -- no pysam in this image to pin them):
//   LVC_OVERLAP_HTSLIB_1_10  (htslib <= 1.10): every read of a proper pair with |tlen| < 2*l_qseq enters the hash;
//       matching bases: first mate gets min(qa+qb, 200), second 0; mismatch: the higher quality keeps 0.8*q (the first
//       mate wins ties), the other 0; only positions where both mates have a match op.
//   LVC_OVERLAP_HTSLIB_1_13  (htslib >= 1.13, default): only reads whose mate is still to arrive enter the hash;
//       which mate keeps the combined quality is decided by a hash of the read name; a deletion in one mate zeroes
//       (or scales by 0.8) the other mate's bases across it; ties on a mismatch follow the name hash.
#pragma once
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/lvc.h"

namespace lvc_overlap {

// ---- khash.h string hash + Wang mix, as htslib uses them to pick the mate that keeps its qualities
static inline uint32_t x31_hash(const char* s, size_t n) {
    if (n == 0) return 0;
    uint32_t h = (uint32_t)(int32_t)(signed char)s[0];
    for (size_t k = 1; k < n; ++k) h = (h << 5) - h + (uint32_t)(int32_t)(signed char)s[k];
    return h;
}
static inline uint32_t wang_hash(uint32_t key) {
    key += ~(key << 15);
    key ^= (key >> 10);
    key += (key << 3);
    key ^= (key >> 6);
    key += ~(key << 11);
    key ^= (key >> 16);
    return key;
}

// ---- htslib cigar_iref2iseq_set / cigar_iref2iseq_next over one read's BAM-encoded CIGAR
struct Cursor {
    const uint32_t* cig;       // current op
    const uint32_t* end;
    const uint32_t* begin;
    int64_t icig = 0, iseq = 0, iref = 0;
};
// returns 0 (on a match base) or -1 (no more cigar / position not covered)
static inline int cursor_set(Cursor& c, int64_t want_iref) {
    int64_t pos = want_iref;
    if (pos < 0) return -1;
    c.icig = 0; c.iseq = 0; c.iref = 0;
    while (c.cig < c.end) {
        const uint32_t op = *c.cig & 15u;
        const int64_t n = *c.cig >> 4;
        if (op == 4) { ++c.cig; c.iseq += n; c.icig = 0; continue; }
        if (op == 5 || op == 6) { ++c.cig; c.icig = 0; continue; }
        if (op == 0 || op == 7 || op == 8) {
            pos -= n;
            if (pos < 0) { c.icig = n + pos; c.iseq += c.icig; c.iref += c.icig; return 0; }
            ++c.cig; c.iseq += n; c.icig = 0; c.iref += n;
            continue;
        }
        if (op == 1) { ++c.cig; c.iseq += n; c.icig = 0; continue; }
        if (op == 2 || op == 3) {
            pos -= n;
            if (pos < 0) pos = 0;
            ++c.cig; c.icig = 0; c.iref += n;
            continue;
        }
        return -2;
    }
    c.iseq = -1;
    return -1;
}
static inline int cursor_next(Cursor& c) {
    while (c.cig < c.end) {
        const uint32_t op = *c.cig & 15u;
        const int64_t n = *c.cig >> 4;
        if (op == 0 || op == 7 || op == 8) {
            if (c.icig >= n - 1) { c.icig = -1; ++c.cig; continue; }
            ++c.iseq; ++c.icig; ++c.iref;
            return 0;
        }
        if (op == 2 || op == 3) { ++c.cig; c.iref += n; c.icig = -1; continue; }
        if (op == 1 || op == 4) { ++c.cig; c.iseq += n; c.icig = -1; continue; }
        if (op == 5 || op == 6) { ++c.cig; c.icig = -1; continue; }
        return -2;
    }
    c.iseq = -1;
    c.iref = -1;
    return -1;
}

struct ReadRef {
    int64_t pos;
    const uint32_t* cig; uint32_t n_cig;
    const uint8_t* seq4;         // BAM nibbles, high nibble first
    uint8_t* qual;
    int64_t l_qseq;
};
static inline uint32_t seqi(const uint8_t* s, int64_t i) { return (s[i >> 1] >> ((~i & 1) << 2)) & 15u; }
static inline uint8_t scale08(uint8_t q) { return (uint8_t)(0.8 * q); }

// the rule for one reference position both mates cover with a base (htslib tweak_overlap_quality)
static inline void rewrite_base(uint8_t& qa, uint8_t& qb, bool same_base, bool legacy, uint8_t amul, uint8_t bmul) {
    if (same_base) {
        const int q = (int)qa + (int)qb;
        const uint8_t qq = (uint8_t)(q > 200 ? 200 : q);
        if (legacy) { qa = qq; qb = 0; }
        else { qa = (uint8_t)(amul * qq); qb = (uint8_t)(bmul * qq); }
    } else if (legacy) {
        if (qa >= qb) { qa = scale08(qa); qb = 0; }
        else { qb = scale08(qb); qa = 0; }
    } else {
        if (qa > qb) { qa = scale08(qa); qb = 0; }
        else if (qa < qb) { qb = scale08(qb); qa = 0; }
        else { qa = (uint8_t)((amul * 0.8) * qa); qb = (uint8_t)((bmul * 0.8) * qb); }
    }
}

// a read whose CIGAR is [clips] ONE match op [clips] (the usual short read): first query index and length of the match
struct SimpleMatch { bool ok; int64_t q0, len; };
static inline SimpleMatch simple_match(const ReadRef& r) {
    SimpleMatch m{false, 0, 0};
    uint32_t k = 0;
    while (k < r.n_cig && ((r.cig[k] & 15u) == 4 || (r.cig[k] & 15u) == 5)) { if ((r.cig[k] & 15u) == 4) m.q0 += r.cig[k] >> 4; ++k; }
    if (k >= r.n_cig) return m;
    const uint32_t op = r.cig[k] & 15u;
    if (!(op == 0 || op == 7 || op == 8)) return m;
    m.len = r.cig[k] >> 4;
    ++k;
    while (k < r.n_cig && ((r.cig[k] & 15u) == 4 || (r.cig[k] & 15u) == 5)) ++k;
    m.ok = k == r.n_cig && m.len > 0 && m.q0 + m.len <= r.l_qseq;
    return m;
}

// htslib tweak_overlap_quality(a = the buffered first read, b = the read being pushed).  `a_keeps`: model 1.13 picks the
// mate that keeps the combined quality by the name hash; model 1.10 always keeps a.  Returns the number of rewritten
// positions.
static inline uint64_t tweak(const ReadRef& a, const ReadRef& b, int model, bool a_keeps) {
    {
        // Two reads of one match op each (clips aside): the walk below visits exactly the reference positions
        // [b.pos, min(end of a, end of b)) with both cursors on a base, so the rule can be applied in a plain loop.
        const SimpleMatch ma = simple_match(a), mb = simple_match(b);
        if (ma.ok && mb.ok && b.pos >= a.pos) {
            const bool legacy = model == LVC_OVERLAP_HTSLIB_1_10;
            const uint8_t amul = legacy ? 1 : (a_keeps ? 1 : 0), bmul = legacy ? 0 : (a_keeps ? 0 : 1);
            const int64_t hi = a.pos + ma.len < b.pos + mb.len ? a.pos + ma.len : b.pos + mb.len;
            uint64_t touched = 0;
            for (int64_t r = b.pos; r < hi; ++r, ++touched) {
                const int64_t ia = ma.q0 + (r - a.pos), ib = mb.q0 + (r - b.pos);
                rewrite_base(a.qual[ia], b.qual[ib], seqi(a.seq4, ia) == seqi(b.seq4, ib), legacy, amul, bmul);
            }
            return touched;
        }
    }
    Cursor ca{a.cig, a.cig + a.n_cig, a.cig}, cb{b.cig, b.cig + b.n_cig, b.cig};
    int64_t iref = b.pos;
    int a_ret = cursor_set(ca, iref - a.pos);
    if (a_ret < 0) return 0;
    int b_ret = cursor_set(cb, iref - b.pos);
    if (b_ret < 0) return 0;
    const bool legacy = model == LVC_OVERLAP_HTSLIB_1_10;
    const uint8_t amul = legacy ? 1 : (a_keeps ? 1 : 0), bmul = legacy ? 0 : (a_keeps ? 0 : 1);
    uint64_t touched = 0;
    for (;;) {
        while (a_ret >= 0 && ca.iref >= 0 && ca.iref < iref - a.pos) a_ret = cursor_next(ca);
        if (a_ret < 0) break;
        if (iref < ca.iref + a.pos) iref = ca.iref + a.pos;
        while (b_ret >= 0 && cb.iref >= 0 && cb.iref < iref - b.pos) b_ret = cursor_next(cb);
        if (b_ret < 0) break;
        if (iref < cb.iref + b.pos) iref = cb.iref + b.pos;
        ++iref;
        if (ca.iref + a.pos != cb.iref + b.pos) {
            if (legacy) continue;                      // only positions matched in both mates
            if (ca.iref + a.pos < cb.iref + b.pos && cb.cig > cb.begin && (*(cb.cig - 1) & 15u) == 2) {
                // deletion in b: a catches up, its bases across the deletion lose their quality
                bool done = false;
                do {
                    if (ca.iseq >= 0 && ca.iseq < a.l_qseq) { a.qual[ca.iseq] = amul ? scale08(a.qual[ca.iseq]) : 0; ++touched; }
                    a_ret = cursor_next(ca);
                    if (a_ret < 0) { done = true; break; }
                } while (ca.iref + a.pos < cb.iref + b.pos);
                if (done) return touched;
            } else if (ca.cig > ca.begin && (*(ca.cig - 1) & 15u) == 2) {
                bool done = false;
                do {
                    if (cb.iseq >= 0 && cb.iseq < b.l_qseq) { b.qual[cb.iseq] = bmul ? scale08(b.qual[cb.iseq]) : 0; ++touched; }
                    b_ret = cursor_next(cb);
                    if (b_ret < 0) { done = true; break; }
                } while (cb.iref + b.pos < ca.iref + a.pos);
                if (done) return touched;
            } else {
                continue;                              // reference skips and the like: untouched
            }
        }
        if (ca.iseq < 0 || cb.iseq < 0 || ca.iseq >= a.l_qseq || cb.iseq >= b.l_qseq) return touched;   // bad CIGAR
        ++touched;
        rewrite_base(a.qual[ca.iseq], b.qual[cb.iseq], seqi(a.seq4, ca.iseq) == seqi(b.seq4, cb.iseq), legacy, amul, bmul);
    }
    return touched;
}

struct NameKey {
    const char* p; uint32_t n;
    bool operator==(const NameKey& o) const { return n == o.n && memcmp(p, o.p, n) == 0; }
};
struct NameHash {
    size_t operator()(const NameKey& k) const {
        uint64_t h = 1469598103934665603ull;
        for (uint32_t i = 0; i < k.n; ++i) { h ^= (uint8_t)k.p[i]; h *= 1099511628211ull; }
        return (size_t)h;
    }
};

// 64-bit hash of a read name (what the name table below is keyed by; the ingest computes it for every read on all threads
// before the sequential admission pass)
static inline uint64_t name_hash64(const char* p, uint32_t n) {
    uint64_t h = 1469598103934665603ull;
    for (uint32_t i = 0; i < n; ++i) { h ^= (uint8_t)p[i]; h *= 1099511628211ull; }
    h ^= h >> 32; h *= 0x9E3779B97F4A7C15ull; h ^= h >> 29;
    return h;
}

// The overlap hash of htslib's pileup iterator (name -> buffered read), as an open-addressing table of (hash, read):
// linear probing, deletion by backward shift, names compared only where the 64-bit hashes agree.  Real paired data
// sends every read through it up to three times (lookup / insert, removal when the mate arrives, removal when the read
// leaves the pileup); a node-based map spends ~200 ns per read there, which made the admission pass the longest phase
// of the ingest on files whose mates overlap.
struct NameTable {
    struct Slot { uint64_t h; uint32_t idx1; };          // idx1 = read index + 1, 0 = free
    std::vector<Slot> t;
    size_t mask = 0, count = 0;
    bool empty() const { return count == 0; }
    void grow() {
        std::vector<Slot> old;
        old.swap(t);
        const size_t cap = old.empty() ? 1024 : old.size() * 2;
        t.assign(cap, Slot{0, 0});
        mask = cap - 1;
        for (const Slot& s : old)
            if (s.idx1) { size_t k = (size_t)s.h & mask; while (t[k].idx1) k = (k + 1) & mask; t[k] = s; }
    }
    // slot of the entry whose name equals that of the probe (same(read) compares the names), or -1
    template <class SameFn>
    long find(uint64_t h, SameFn same) const {
        if (!count) return -1;
        for (size_t k = (size_t)h & mask;; k = (k + 1) & mask) {
            if (!t[k].idx1) return -1;
            if (t[k].h == h && same(t[k].idx1 - 1)) return (long)k;
        }
    }
    void insert(uint64_t h, uint32_t idx) {               // the name is known to be absent
        if ((count + 1) * 2 > t.size()) grow();
        size_t k = (size_t)h & mask;
        while (t[k].idx1) k = (k + 1) & mask;
        t[k] = Slot{h, idx + 1};
        ++count;
    }
    void erase(long at) {
        size_t k = (size_t)at;
        for (;;) {                                        // close the gap: move back every entry the gap separates from its home
            size_t j = k;
            for (;;) {
                j = (j + 1) & mask;
                if (!t[j].idx1) { t[k].idx1 = 0; --count; return; }
                const size_t home = (size_t)t[j].h & mask;
                // the entry at j may move to k iff its home is NOT cyclically in (k, j]
                const bool stays = k <= j ? (home > k && home <= j) : (home > k || home <= j);
                if (!stays) break;
            }
            t[k] = t[j];
            k = j;
        }
    }
};

// a pair of overlapping mates found by the admission pass whose quality rewrite is still to be done (apply_pending)
struct PendingTweak {
    uint32_t a, b;       // read indices: `a` was buffered first
    bool a_keeps;
};

// Admission (SURVEY B2 + B4, htslib bam_plp_push / bam_plp_next) with the overlap hash (B5).
// name(i) -> NameKey of read i; mate_* may be null when overlap_model == LVC_OVERLAP_OFF.
// `pending` != nullptr: the quality rewrites are not done here but listed (the admission itself never looks at a base or
// a quality, so it can run while the payload is still being packed); seq4 / qual are not touched then.
template <class NameFn>
static int admit_core(uint32_t n, const int32_t* pos, const uint16_t* flag, const uint8_t* mapq, const uint32_t* cigar_off,
                      const uint32_t* cigar, const uint64_t* seq_off, const uint8_t* seq4, uint8_t* qual, NameFn name,
                      const int32_t* mate_pos, const int8_t* mate_ref, const int32_t* tlen, int min_mq, int max_depth,
                      int overlap_model, uint8_t* keep, uint64_t* n_pairs, uint64_t* n_bases,
                      std::vector<PendingTweak>* pending = nullptr, const uint64_t* name_hash = nullptr) {
    constexpr uint32_t kFilter = 0x4u | 0x100u | 0x200u | 0x400u;
    const bool ov = overlap_model != LVC_OVERLAP_OFF;
    std::vector<uint32_t> ring;            // buffered reads per end position
    std::vector<uint32_t> ring_head;       // overlap handling: list of the reads that end there (index + 1)
    std::vector<uint32_t> next_in_slot;    // linked through this
    if (ov) next_in_slot.assign(n, 0);
    NameTable olap;
    // name_hash (optional): name_hash64 of every read's name, made beforehand; else hashed here
    auto hash_of = [&](uint32_t i) { if (name_hash) return name_hash[i]; const NameKey k = name(i); return name_hash64(k.p, k.n); };
    auto find_name = [&](uint32_t i, uint64_t h) {
        const NameKey key = name(i);
        return olap.find(h, [&](uint32_t j) { return j == i || name(j) == key; });
    };
    auto erase_name = [&](uint32_t i) { const long at = find_name(i, hash_of(i)); if (at >= 0) olap.erase(at); };
    int64_t ring_base = 0;
    auto ring_add = [&](int64_t e, int64_t p, uint32_t i) {
        if (ring.empty()) { ring.assign(4096, 0); if (ov) ring_head.assign(4096, 0); ring_base = p; }
        if (e < ring_base) return;
        size_t off = (size_t)(e - ring_base);
        if (off >= ring.size()) {
            const size_t want = std::max(ring.size() * 2, off + 1);
            ring.resize(want, 0);
            if (ov) ring_head.resize(want, 0);
        }
        ring[off]++;
        if (ov) { next_in_slot[i] = ring_head[off]; ring_head[off] = i + 1; }
    };
    auto read_ref = [&](uint32_t i) {
        ReadRef r;
        r.pos = pos[i]; r.cig = cigar + cigar_off[i]; r.n_cig = cigar_off[i + 1] - cigar_off[i];
        r.seq4 = seq4 + (seq_off[i] >> 1); r.qual = qual + seq_off[i];
        int64_t lq = 0;
        for (uint32_t k = 0; k < r.n_cig; ++k) { const uint32_t op = r.cig[k] & 15u; if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) lq += r.cig[k] >> 4; }
        r.l_qseq = lq;
        return r;
    };
    uint64_t pairs = 0, bases = 0;
    int64_t iter_pos = 0, max_pos = -1, nbuf = 0;
    for (uint32_t i = 0; i < n; ++i) {
        keep[i] = 0;
        const uint32_t f = flag[i];
        if (f & kFilter) continue;                     // filtered by the stepper: never pushed
        if ((int)mapq[i] < min_mq) continue;
        if ((f & 0x1u) && !(f & 0x2u)) continue;
        int64_t rlen = 0;
        for (uint32_t k = cigar_off[i]; k < cigar_off[i + 1]; ++k) {
            const uint32_t op = cigar[k] & 15u;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += cigar[k] >> 4;
        }
        if (rlen == 0) continue;                       // malformed: no reference-consuming op (htslib asserts)
        const int64_t p = pos[i], e = p + rlen;
        if (p < max_pos) return LVC_EUNSORTED;
        if (p == iter_pos && nbuf + 1 > (int64_t)max_depth) {       // bam_plp_push: cnt > maxcnt -> overlap_remove, drop
            if (ov && !olap.empty()) erase_name(i);
            continue;
        }
        max_pos = p;
        keep[i] = 1;
        nbuf++;
        ring_add(e, p, i);
        if (ov && (f & 0x2u) && !(f & 0x8u)) {
            // overlap_push
            const int64_t lq = (int64_t)(seq_off[i + 1] - seq_off[i]);      // l_qseq (+1 pad byte for odd lengths)
            int64_t l_qseq = 0;
            for (uint32_t k = cigar_off[i]; k < cigar_off[i + 1]; ++k) { const uint32_t op = cigar[k] & 15u; if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) l_qseq += cigar[k] >> 4; }
            (void)lq;
            const int64_t isz = tlen ? (int64_t)tlen[i] : 0;
            const int64_t aisz = isz < 0 ? -isz : isz;
            const int64_t mp = mate_pos ? (int64_t)mate_pos[i] : -1;
            const int mref = mate_ref ? (int)mate_ref[i] : -1;
            bool possible;
            if (overlap_model == LVC_OVERLAP_HTSLIB_1_10) possible = !(aisz >= 2 * l_qseq);
            else possible = !((mref == 0) || (aisz >= 2 * l_qseq && mp >= e));
            if (possible) {
                // (nothing is buffered: the lookup cannot hit, and a read that is not added either never has its name
                // touched -- the name bytes live in the inflated file, one cache miss per read)
                const bool add = overlap_model == LVC_OVERLAP_HTSLIB_1_10 ? true : (mp >= p || ((f & 0x1u) && mp == -1));
                const uint64_t hi = hash_of(i);
                if (olap.empty()) {
                    if (add) olap.insert(hi, i);
                    goto pushed;
                }
                const long at = find_name(i, hi);
                if (at < 0) {
                    if (add) olap.insert(hi, i);
                } else {
                    const uint32_t a = olap.t[(size_t)at].idx1 - 1;
                    olap.erase(at);
                    const bool a_keeps = (wang_hash(x31_hash(name(a).p, name(a).n)) & 1u) != 0;
                    if (pending) pending->push_back(PendingTweak{a, i, a_keeps});
                    else {
                        const uint64_t t = tweak(read_ref(a), read_ref(i), overlap_model, a_keeps);
                        if (t) { ++pairs; bases += t; }
                    }
                }
            }
        }
    pushed:
        // bam_plp_next: emit columns while max_pos > iter_pos; each built column frees ended reads
        while (max_pos > iter_pos) {
            const int64_t c = iter_pos;
            if (!ring.empty() && c >= ring_base && (size_t)(c - ring_base) < ring.size()) {
                const size_t off = (size_t)(c - ring_base);
                nbuf -= ring[off];
                ring[off] = 0;
                if (ov) {
                    if (!olap.empty())
                        for (uint32_t r = ring_head[off]; r; r = next_in_slot[r - 1]) { erase_name(r - 1); if (olap.empty()) break; }
                    ring_head[off] = 0;
                }
            }
            if (nbuf - 1 == 0) iter_pos = max_pos;     // only the new read is buffered: jump to it
            else iter_pos = c + 1;
        }
        // slide the ring so it does not grow with the genome
        if (!ring.empty() && iter_pos - ring_base > (int64_t)ring.size() / 2) {
            const size_t shift = (size_t)(iter_pos - ring_base);
            if (shift >= ring.size()) {
                std::fill(ring.begin(), ring.end(), 0);
                if (ov) std::fill(ring_head.begin(), ring_head.end(), 0);
            } else {
                std::move(ring.begin() + shift, ring.end(), ring.begin());
                std::fill(ring.end() - shift, ring.end(), 0);
                if (ov) {
                    std::move(ring_head.begin() + shift, ring_head.end(), ring_head.begin());
                    std::fill(ring_head.end() - shift, ring_head.end(), 0);
                }
            }
            ring_base = iter_pos;
        }
    }
    if (n_pairs) *n_pairs = pairs;
    if (n_bases) *n_bases = bases;
    return LVC_OK;
}

// The rewrites admit_core listed.  A read takes part in at most one of them (the buffered mate leaves the hash when its
// partner arrives, and the partner never enters it), so they are independent: entries [k0, k1) can be applied by any thread.
static void apply_pending(const PendingTweak* list, size_t k0, size_t k1, const int32_t* pos, const uint32_t* cigar_off,
                          const uint32_t* cigar, const uint64_t* seq_off, const uint8_t* seq4, uint8_t* qual, int overlap_model,
                          uint64_t* n_pairs, uint64_t* n_bases) {
    auto read_ref = [&](uint32_t i) {
        ReadRef r;
        r.pos = pos[i]; r.cig = cigar + cigar_off[i]; r.n_cig = cigar_off[i + 1] - cigar_off[i];
        r.seq4 = seq4 + (seq_off[i] >> 1); r.qual = qual + seq_off[i];
        int64_t lq = 0;
        for (uint32_t k = 0; k < r.n_cig; ++k) { const uint32_t op = r.cig[k] & 15u; if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) lq += r.cig[k] >> 4; }
        r.l_qseq = lq;
        return r;
    };
    uint64_t pairs = 0, bases = 0;
    for (size_t k = k0; k < k1; ++k) {
        const uint64_t t = tweak(read_ref(list[k].a), read_ref(list[k].b), overlap_model, list[k].a_keeps);
        if (t) { ++pairs; bases += t; }
    }
    *n_pairs = pairs; *n_bases = bases;
}

}  // namespace lvc_overlap
