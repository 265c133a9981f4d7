// General deposit kernel: one thread per read, one global reduction (RED) per deposited base.
//
// Replaces, for every read the pileup engine would admit, the per-(column, read) Python loop of
// live_variant_caller.py:69-70,89-103 and the htslib CIGAR walk behind it (resolve_cigar2).
// It takes ANY well-formed record (every CIGAR op, every nibble code, any read length), so it is both
// the correctness anchor for the tiled kernel and the path for reads the tiled kernel defers.
#pragma once
#include "lvc_common.cuh"

namespace lvc {

__device__ __forceinline__ void mark_unmapped(const TableView& tv, uint32_t key) {
    atomicOr(&tv.newkeys[key >> 5], 1u << (key & 31));
    atomicAdd(&tv.status[ST_UNMAPPED], 1u);
}

// deposit one passing base (nibble `nib`, quality `q`) of read ordinal `ord` at reference column r.
// PEER = true: peer tables are attached (lvc_peer_attach) and the cell is looked up at the rank that owns the column;
// the PEER = false instantiations contain no routing code.
template <bool PEER = false>
__device__ __forceinline__ void deposit_base(const TableView& tv, const DepositParams& dp, int64_t r, uint32_t nib,
                                             uint32_t q, uint32_t ord) {
    const uint32_t gs = nibble_gs(nib);
    const uint32_t key = ((gs >> 2) << 8) | q;
    if (dp.replay && !((dp.replay_keys[key >> 5] >> (key & 31)) & 1u)) return;
    const uint16_t pl = tv.lut[key];
    if (pl == kNoPlane) {
        if (!dp.replay) mark_unmapped(tv, key);
        return;
    }
    atomicAdd((PEER ? plane_row(tv, pl, r) : tv.planes[pl] + r * 4) + (gs & 3u), 1u);
    // first-seen ordinal.  Every quality plane of a group shares ONE first-seen cell per (column, allele): at depth d
    // an unconditional RED.MIN would put d same-address atomics per batch on it (they serialise in L2), so test first;
    // after the first reads of a column the test fails and nothing is written.  The test may read a STALE line from L1:
    // the cell only ever decreases, so a stale value is >= the current one and can only cause a redundant atomicMin,
    // never a missed one (and the reads of a CTA share a few hundred columns, so the test mostly hits L1).
    uint32_t* f = (PEER ? first_row(tv, gs >> 2, r) : tv.first[gs >> 2] + r * 4) + (gs & 3u);
    if (*f > ord) atomicMin(f, ord);
}

// Walk one read (one thread).
template <bool PEER = false>
__device__ __forceinline__ void deposit_read_general(const BatchView& b, const TableView& tv, const DepositParams& dp,
                                                     uint32_t i) {
    const uint32_t flag = b.flag[i];
    if (!read_passes_filter(flag, b.mapq[i], b.keep[i], dp.min_mq)) return;
    const uint32_t c0 = b.cigar_off[i], c1 = b.cigar_off[i + 1];
    // first pass over the ops: reference length and l_qseq
    int64_t rlen = 0;
    uint32_t lq = 0;
    for (uint32_t k = c0; k < c1; ++k) {
        const uint32_t c = b.cigar[k], op = c & 15u, len = c >> 4;
        if (op_consumes_ref(op)) rlen += len;
        if (op_consumes_query(op)) lq += len;
    }
    if (rlen == 0) return;   // no M/D/N/=/X op: htslib asserts on such records; skipped
    const int64_t pos = b.pos[i];
    if (pos < 0 || pos + rlen > tv.G) {
        atomicAdd(&tv.status[ST_RANGE_ERR], 1u);
        return;
    }
    if (!dp.replay) {
        atomicAdd(PEER ? covdiff_cell(tv, pos) : tv.covdiff + pos, 1);
        atomicAdd(PEER ? covdiff_cell(tv, pos + rlen) : tv.covdiff + pos + rlen, -1);
    }
    const uint64_t qb = b.seq_off[i];
    const uint32_t ord = dp.ord_base + i;
    int64_t r = pos;
    uint32_t qi = 0;
    for (uint32_t k = c0; k < c1; ++k) {
        const uint32_t c = b.cigar[k], op = c & 15u, len = c >> 4;
        if (op_is_match(op)) {
            for (uint32_t j = 0; j < len; ++j, ++qi, ++r) {
                const uint32_t q = batch_qual(b, qb + qi);
                if ((int)q < dp.min_bq) continue;
                deposit_base<PEER>(tv, dp, r, batch_nibble(b, qb + qi), q, ord);
            }
        } else if (op == 2 || op == 3) {
            // deletion / ref-skip entries are kept iff the NEXT query base passes the quality rule
            // (pysam pileup_base_qual_skip on qpos = y; 0 if qpos >= l_qseq) -- SURVEY B3
            const uint32_t q = (qi < lq) ? batch_qual(b, qb + qi) : 0u;
            if (!dp.replay && (int)q >= dp.min_bq) {
                for (uint32_t j = 0; j < len; ++j) atomicAdd(PEER ? dels_cell(tv, r + j) : tv.dels + r + j, 1u);
            }
            r += len;
        } else if (op == 1 || op == 4) {
            qi += len;
        }   // H, P: nothing
    }
}

// The same walk with one WARP per read.  The CIGAR is read cooperatively: lane k holds op k of a group of 32 ops, the
// reference / query offsets of every op come from a warp scan, and the ops that deposit something (match runs,
// deletions, skips) are visited by broadcasting them from their lane -- no dependent global load per op.  Inside a
// match run the lanes stride over the bases (coalesced loads, 32 reductions in flight).
// COMPACT (the warp-per-read kernel): the quality test and the deposit are separated.  Per group of 32 ops the lanes
// test the group's query bases FOUR AT A TIME (aligned 32-bit words of the quality array, byte-parallel compare) and
// append the passing ones, packed as (query index | quality << 24), to a 256-entry ring in shared memory; 32 entries
// at a time are then resolved to their op with a shuffle binary search over the group's query offsets (entries in
// soft clips and insertions are dropped there) and deposited with every lane busy.  With a threshold that most
// bases fail (ONT at minBQ 30: 85 %) the ~70-instruction resolve-and-deposit sequence runs once per 32 PASSING
// bases, and the test costs ~1/4 instruction per base.
constexpr uint32_t kRingEntries = 256;
struct WarpRing {
    uint32_t e[kRingEntries];
};

// 0x80 in every byte of w that is >= min_bq (any min_bq)
__device__ __forceinline__ uint32_t ge_flags4(uint32_t w, int min_bq) {
    if (min_bq <= 0) return 0x80808080u;
    if (min_bq <= 128) return (((w & 0x7F7F7F7Fu) + (uint32_t)(0x80 - min_bq) * 0x01010101u) | w) & 0x80808080u;
    uint32_t f = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) f |= (int)((w >> (8 * k)) & 255u) >= min_bq ? (0x80u << (8 * k)) : 0u;
    return f;
}

template <bool COMPACT, bool PEER = false>
__device__ __forceinline__ void deposit_read_warp_impl(const BatchView& b, const TableView& tv, const DepositParams& dp,
                                                       uint32_t i, uint32_t lane, WarpRing* ring = nullptr) {
    if (!read_passes_filter(b.flag[i], b.mapq[i], b.keep[i], dp.min_mq)) return;
    const uint32_t c0 = b.cigar_off[i], c1 = b.cigar_off[i + 1];
    const int64_t pos = b.pos[i];
    const uint64_t qb = b.seq_off[i];
    // group 0 of the ops (all of them for reads with <= 32 ops) stays in registers
    const uint32_t cg_first = c0 + lane < c1 ? b.cigar[c0 + lane] : 0u;       // padding: a match of length 0
    // scan of group 0 (inclusive prefix of the reference / query lengths), needed below anyway
    uint32_t r_in0, q_in0;
    {
        const uint32_t op = cg_first & 15u, len = cg_first >> 4;
        r_in0 = op_consumes_ref(op) ? len : 0u;
        q_in0 = op_consumes_query(op) ? len : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t ur = __shfl_up_sync(0xFFFFFFFFu, r_in0, d);
            const uint32_t uq = __shfl_up_sync(0xFFFFFFFFu, q_in0, d);
            if ((int)lane >= d) { r_in0 += ur; q_in0 += uq; }
        }
    }
    // totals: reference length and l_qseq.  One group of ops whose lengths are all < 2^26 (any real read): the scan
    // has them.  Otherwise sum group by group (per-op lengths are < 2^28, so 32 of them cannot wrap 64 bits).  A read
    // whose reference span does not fit 31 bits cannot lie inside any contig and is reported as out of range.
    uint64_t rlen = 0, lq64 = 0;
    if (c1 - c0 <= 32u && !__any_sync(0xFFFFFFFFu, (cg_first >> 4) >= (1u << 26))) {
        rlen = __shfl_sync(0xFFFFFFFFu, r_in0, 31);
        lq64 = __shfl_sync(0xFFFFFFFFu, q_in0, 31);
    } else {
        for (uint32_t g0 = c0; g0 < c1; g0 += 32) {
            const uint32_t c = g0 == c0 ? cg_first : (g0 + lane < c1 ? b.cigar[g0 + lane] : 0u);
            const uint32_t op = c & 15u, len = c >> 4;
            const uint32_t rl = op_consumes_ref(op) ? len : 0u, ql = op_consumes_query(op) ? len : 0u;
            rlen += (uint64_t)__reduce_add_sync(0xFFFFFFFFu, rl >> 8) * 256u + __reduce_add_sync(0xFFFFFFFFu, rl & 255u);
            lq64 += (uint64_t)__reduce_add_sync(0xFFFFFFFFu, ql >> 8) * 256u + __reduce_add_sync(0xFFFFFFFFu, ql & 255u);
        }
    }
    if (rlen == 0) return;   // no M/D/N/=/X op: htslib asserts on such records; skipped
    if (pos < 0 || rlen > 0x7FFFFFFFull || lq64 > 0xFFFFFFFFull || pos + (int64_t)rlen > tv.G) {
        if (lane == 0) atomicAdd(&tv.status[ST_RANGE_ERR], 1u);
        return;
    }
    const uint32_t lq = (uint32_t)lq64;
    if (!dp.replay && lane == 0) {
        atomicAdd(PEER ? covdiff_cell(tv, pos) : tv.covdiff + pos, 1);
        atomicAdd(PEER ? covdiff_cell(tv, pos + (int64_t)rlen) : tv.covdiff + pos + (int64_t)rlen, -1);
    }
    // (byte form; a quality-code batch keeps its codes at a quarter of the offset and is read through batch_qual)
    const uint8_t* qual = b.qual + (b.qbits == 2u ? (qb >> 2) : qb);
    // (bases: 4-bit codes at half the offset, or 2-bit codes at a quarter of it: read through batch_nibble)
    const uint8_t* seq = b.seq4 + (b.sbits == 2u ? (qb >> 2) : (qb >> 1));
    // request the read's whole payload now (one line per lane): the per-run loads below then hit L1 instead of
    // paying a DRAM latency per run
    for (uint32_t off = lane * 128u; off < lq; off += 32u * 128u) {
        if (b.qbits != 2u || off < (lq + 3) / 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(qual + off));
        if (off < (b.sbits == 2u ? (lq + 3) / 4 : (lq + 1) / 2)) asm volatile("prefetch.global.L1 [%0];" ::"l"(seq + off));
    }
    const uint32_t ord = dp.ord_base + i;
    uint32_t ring_head = 0, ring_n = 0;
    const bool wide = COMPACT && lq < (1u << 24);           // ring entries carry 24-bit query indices
    uint32_t r_base = 0, q_base = 0;                         // offsets from pos / from the first query base
    for (uint32_t g0 = c0; g0 < c1; g0 += 32) {
        const uint32_t c = g0 == c0 ? cg_first : (g0 + lane < c1 ? b.cigar[g0 + lane] : 0u);
        const uint32_t op = c & 15u, len = c >> 4;
        const uint32_t rl = op_consumes_ref(op) ? len : 0u, ql = op_consumes_query(op) ? len : 0u;
        // inclusive prefix of the reference / query lengths inside the group
        uint32_t r_in = r_in0, q_in = q_in0;
        if (g0 != c0) {
            r_in = rl; q_in = ql;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t ur = __shfl_up_sync(0xFFFFFFFFu, r_in, d);
                const uint32_t uq = __shfl_up_sync(0xFFFFFFFFu, q_in, d);
                if ((int)lane >= d) { r_in += ur; q_in += uq; }
            }
        }
        const uint32_t r_off = r_base + r_in - rl, q_off = q_base + q_in - ql;
        if (wide && !dp.replay && len != 0 && (op == 2 || op == 3)) {
            // deletion / ref-skip entries, every such op of the group in its own lane: kept iff the NEXT query base
            // passes the quality rule (pysam pileup_base_qual_skip on qpos = y; 0 if qpos >= l_qseq) -- SURVEY B3
            const uint32_t q = (q_off < lq) ? batch_qual(b, qb + q_off) : 0u;
            if ((int)q >= dp.min_bq)
                for (uint32_t j = 0; j < len; ++j) atomicAdd(PEER ? dels_cell(tv, pos + r_off + j) : tv.dels + pos + r_off + j, 1u);
        }
        // without the word-parallel path: the ops with something to deposit, one after the other
        uint32_t work = __ballot_sync(0xFFFFFFFFu, !wide && len != 0 && (op_is_match(op) || op == 2 || op == 3));
        while (work) {
            const int k = __ffs(work) - 1;
            work &= work - 1;
            const uint32_t ck = __shfl_sync(0xFFFFFFFFu, c, k);
            const int64_t r = pos + (int64_t)__shfl_sync(0xFFFFFFFFu, r_off, k);
            const uint32_t qi = __shfl_sync(0xFFFFFFFFu, q_off, k);
            const uint32_t opk = ck & 15u, lenk = ck >> 4;
            if (op_is_match(opk)) {
                for (uint32_t j = lane; j < lenk; j += 32) {
                    const uint32_t q = batch_qual(b, qb + qi + j);
                    if ((int)q < dp.min_bq) continue;
                    deposit_base<PEER>(tv, dp, r + j, batch_nibble(b, qb + qi + j), q, ord);
                }
            } else if (!dp.replay) {
                // deletion / ref-skip entries are kept iff the NEXT query base passes the quality rule
                // (pysam pileup_base_qual_skip on qpos = y; 0 if qpos >= l_qseq) -- SURVEY B3
                const uint32_t q = (qi < lq) ? batch_qual(b, qb + qi) : 0u;
                if ((int)q >= dp.min_bq)
                    for (uint32_t j = lane; j < lenk; j += 32) atomicAdd(PEER ? dels_cell(tv, r + j) : tv.dels + r + j, 1u);
            }
        }
        const uint32_t r_tot = __shfl_sync(0xFFFFFFFFu, r_in, 31), q_tot = __shfl_sync(0xFFFFFFFFu, q_in, 31);
        const uint32_t match_ops = __ballot_sync(0xFFFFFFFFu, wide && len != 0 && op_is_match(op));
        if (match_ops) {
            // only the query bases between the first and the last match run of the group are tested (soft clips at the
            // ends of the read would be dropped at the op lookup anyway)
            const uint32_t t_lo = __shfl_sync(0xFFFFFFFFu, q_off, __ffs(match_ops) - 1);
            const uint32_t t_hi = __shfl_sync(0xFFFFFFFFu, q_off + ql, 31 - __clz(match_ops));
            // query offsets of the group's ops, padded with +inf, for the op lookup
            const uint32_t q_key = g0 + lane < c1 ? q_off : 0xFFFFFFFFu;
            // resolve and deposit the first `cnt` (<= 32) entries of the ring
            auto flush = [&](uint32_t cnt) {
                uint32_t x = 0, q = 0;
                if (lane < cnt) { const uint32_t e = ring->e[(ring_head + lane) & (kRingEntries - 1)]; x = e & 0xFFFFFFu; q = e >> 24; }
                else x = t_lo;                                // idle lanes take part in the shuffles
                uint32_t k = 0;                               // number of ops with q_off <= x  (>= 1)
#pragma unroll
                for (int st = 16; st >= 1; st >>= 1)
                    if (__shfl_sync(0xFFFFFFFFu, q_key, (int)(k + st - 1)) <= x) k += st;
                if (__shfl_sync(0xFFFFFFFFu, q_key, (int)k) <= x) ++k;
                const int ko = (int)k - 1;
                const uint32_t co = __shfl_sync(0xFFFFFFFFu, c, ko);
                const uint32_t qo = __shfl_sync(0xFFFFFFFFu, q_off, ko), ro = __shfl_sync(0xFFFFFFFFu, r_off, ko);
                if (lane < cnt && op_is_match(co & 15u)) {
                    deposit_base<PEER>(tv, dp, pos + (int64_t)ro + (x - qo), batch_nibble(b, qb + x), q, ord);
                }
                ring_head = (ring_head + cnt) & (kRingEntries - 1);
                ring_n -= cnt;
            };
            // the query bases [t_lo, t_hi) as aligned words of the quality array
            const uint32_t mis = (uint32_t)(qb & 3u);                       // the read starts `mis` bytes into a word
            const uint32_t* qw = reinterpret_cast<const uint32_t*>(b.qual + (qb - mis));
            const uint32_t w_end = (mis + t_hi + 3u) >> 2;
            const uint32_t lt = (1u << lane) - 1u;
            for (uint32_t w0 = (mis + t_lo) >> 2; w0 < w_end; w0 += 32) {
                const uint32_t w = w0 + lane;
                uint32_t f = 0, word = 0;
                const int32_t x0 = (int32_t)(w << 2) - (int32_t)mis;          // query index of byte 0 (may be < t_lo)
                if (w < w_end) {
                    word = qw[w];
                    const int32_t lo = max((int32_t)t_lo - x0, 0), hi = min((int32_t)t_hi - x0, 4);
                    if (hi > lo) f = ge_flags4(word, dp.min_bq) & (0xFFFFFFFFu << (8 * lo)) & (0xFFFFFFFFu >> (8 * (4 - hi)));
                }
                // exclusive prefix of the per-lane counts (0..4) from three ballots
                const uint32_t cnt = __popc(f);
                const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, cnt & 1u), b1 = __ballot_sync(0xFFFFFFFFu, cnt & 2u),
                               b2 = __ballot_sync(0xFFFFFFFFu, cnt & 4u);
                const uint32_t total = __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
                if (total == 0) continue;
                uint32_t slot = ring_head + ring_n + __popc(b0 & lt) + 2u * __popc(b1 & lt) + 4u * __popc(b2 & lt);
                while (f) {
                    const int bb = (__ffs(f) - 1) >> 3;
                    f &= f - 1;
                    ring->e[slot & (kRingEntries - 1)] = (uint32_t)(x0 + bb) | (((word >> (8 * bb)) & 255u) << 24);
                    ++slot;
                }
                ring_n += total;
                __syncwarp();
                while (ring_n >= 32) flush(32);
                __syncwarp();
            }
            // entries are resolved against THIS group's ops: nothing may stay in the ring
            __syncwarp();
            if (ring_n) flush(ring_n);
            __syncwarp();
        }
        r_base += r_tot;
        q_base += q_tot;
    }
}

// out of line: what the tiled kernels call for the few reads they hand over
__device__ __noinline__ void deposit_read_warp(const BatchView& b, const TableView& tv, const DepositParams& dp,
                                               uint32_t i, uint32_t lane) {
    deposit_read_warp_impl<false, false>(b, tv, dp, i, lane);
}
// the same with peer tables attached (columns of other ranks are reduced into their tables)
__device__ __noinline__ void deposit_read_warp_peer(const BatchView& b, const TableView& tv, const DepositParams& dp,
                                                    uint32_t i, uint32_t lane) {
    deposit_read_warp_impl<false, true>(b, tv, dp, i, lane);
}

// one thread per read of the batch
template <bool PEER>
__global__ void __launch_bounds__(128) k_deposit_general(const __grid_constant__ BatchView b,
                                                         const __grid_constant__ TableView tv,
                                                         const __grid_constant__ DepositParams dp, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) deposit_read_general<PEER>(b, tv, dp, i);
}

// one WARP per read of the batch: long reads with many CIGAR ops and a wide quality alphabet (ONT), where neither a
// primary quality nor a short run table exists.  Every read of the batch is in flight at once.
constexpr int kWarpKernelThreads = 256;
template <bool PEER>
__global__ void __launch_bounds__(kWarpKernelThreads) k_deposit_warp(const __grid_constant__ BatchView b,
                                                                     const __grid_constant__ TableView tv,
                                                                     const __grid_constant__ DepositParams dp, uint32_t n) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");     // the tables may still be read by the previous kernel
    __shared__ WarpRing rings[kWarpKernelThreads / 32];
    const uint32_t i = blockIdx.x * (kWarpKernelThreads / 32) + (threadIdx.x >> 5);
    if (i < n) deposit_read_warp_impl<true, PEER>(b, tv, dp, i, threadIdx.x & 31u, &rings[threadIdx.x >> 5]);
}

}  // namespace lvc
