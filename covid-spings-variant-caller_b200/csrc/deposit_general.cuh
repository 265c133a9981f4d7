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

// deposit one passing base (nibble `nib`, quality `q`) of read ordinal `ord` at reference column r
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
    const int64_t cell = r * 4 + (gs & 3u);
    atomicAdd(&tv.planes[pl][cell], 1u);
    // first-seen ordinal: an unconditional reduction (fire and forget) instead of a load the warp would wait for
    atomicMin(&tv.first[gs >> 2][cell], ord);
}

// Walk one read (one thread).
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
        atomicAdd(&tv.covdiff[pos], 1);
        atomicAdd(&tv.covdiff[pos + rlen], -1);
    }
    const uint64_t qb = b.seq_off[i];
    const uint8_t* qual = b.qual + qb;
    const uint8_t* seq = b.seq4 + (qb >> 1);
    const uint32_t ord = dp.ord_base + i;
    int64_t r = pos;
    uint32_t qi = 0;
    for (uint32_t k = c0; k < c1; ++k) {
        const uint32_t c = b.cigar[k], op = c & 15u, len = c >> 4;
        if (op_is_match(op)) {
            for (uint32_t j = 0; j < len; ++j, ++qi, ++r) {
                const uint32_t q = qual[qi];
                if ((int)q < dp.min_bq) continue;
                const uint32_t byte = seq[qi >> 1];
                const uint32_t nib = (qi & 1u) ? (byte & 15u) : (byte >> 4);
                deposit_base(tv, dp, r, nib, q, ord);
            }
        } else if (op == 2 || op == 3) {
            // deletion / ref-skip entries are kept iff the NEXT query base passes the quality rule
            // (pysam pileup_base_qual_skip on qpos = y; 0 if qpos >= l_qseq) -- SURVEY B3
            const uint32_t q = (qi < lq) ? (uint32_t)qual[qi] : 0u;
            if (!dp.replay && (int)q >= dp.min_bq) {
                for (uint32_t j = 0; j < len; ++j) atomicAdd(&tv.dels[r + j], 1u);
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
// COMPACT: the quality test and the deposit are separated.  Lanes test one base each and append the passing ones to a
// 64-entry ring in shared memory (`ring`, 64 x {column offset u32, query index u32, quality u8} per warp); whenever 32
// entries are waiting they are deposited with every lane busy.  With a threshold that most bases fail (ONT at minBQ
// 30: 85 %) the ~40-instruction deposit sequence runs once per 32 PASSING bases instead of once per 32 bases.
struct WarpRing {
    uint32_t rr[64];
    uint32_t qx[64];
    uint8_t q[64];
};

template <bool COMPACT>
__device__ __forceinline__ void deposit_read_warp_impl(const BatchView& b, const TableView& tv, const DepositParams& dp,
                                                       uint32_t i, uint32_t lane, WarpRing* ring = nullptr) {
    if (!read_passes_filter(b.flag[i], b.mapq[i], b.keep[i], dp.min_mq)) return;
    const uint32_t c0 = b.cigar_off[i], c1 = b.cigar_off[i + 1];
    const int64_t pos = b.pos[i];
    const uint64_t qb = b.seq_off[i];
    // group 0 of the ops (all of them for reads with <= 32 ops) stays in registers
    const uint32_t cg_first = c0 + lane < c1 ? b.cigar[c0 + lane] : 0u;       // padding: a match of length 0
    // totals: reference length and l_qseq.  Per-op lengths are < 2^28, so 32 of them cannot wrap 64 bits; a read
    // whose reference span does not fit 31 bits cannot lie inside any contig and is reported as out of range.
    uint64_t rlen = 0, lq64 = 0;
    for (uint32_t g0 = c0; g0 < c1; g0 += 32) {
        const uint32_t c = g0 == c0 ? cg_first : (g0 + lane < c1 ? b.cigar[g0 + lane] : 0u);
        const uint32_t op = c & 15u, len = c >> 4;
        const uint32_t rl = op_consumes_ref(op) ? len : 0u, ql = op_consumes_query(op) ? len : 0u;
        rlen += (uint64_t)__reduce_add_sync(0xFFFFFFFFu, rl >> 8) * 256u + __reduce_add_sync(0xFFFFFFFFu, rl & 255u);
        lq64 += (uint64_t)__reduce_add_sync(0xFFFFFFFFu, ql >> 8) * 256u + __reduce_add_sync(0xFFFFFFFFu, ql & 255u);
    }
    if (rlen == 0) return;   // no M/D/N/=/X op: htslib asserts on such records; skipped
    if (pos < 0 || rlen > 0x7FFFFFFFull || lq64 > 0xFFFFFFFFull || pos + (int64_t)rlen > tv.G) {
        if (lane == 0) atomicAdd(&tv.status[ST_RANGE_ERR], 1u);
        return;
    }
    const uint32_t lq = (uint32_t)lq64;
    if (!dp.replay && lane == 0) {
        atomicAdd(&tv.covdiff[pos], 1);
        atomicAdd(&tv.covdiff[pos + (int64_t)rlen], -1);
    }
    const uint8_t* qual = b.qual + qb;
    const uint8_t* seq = b.seq4 + (qb >> 1);
    // request the read's whole payload now (one line per lane): the per-run loads below then hit L1 instead of
    // paying a DRAM latency per run
    for (uint32_t off = lane * 128u; off < lq; off += 32u * 128u) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(qual + off));
        if (off < (lq + 1) / 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(seq + off));
    }
    const uint32_t ord = dp.ord_base + i;
    uint32_t ring_head = 0, ring_n = 0;
    uint32_t r_base = 0, q_base = 0;                         // offsets from pos / from the first query base
    for (uint32_t g0 = c0; g0 < c1; g0 += 32) {
        const uint32_t c = g0 == c0 ? cg_first : (g0 + lane < c1 ? b.cigar[g0 + lane] : 0u);
        const uint32_t op = c & 15u, len = c >> 4;
        const uint32_t rl = op_consumes_ref(op) ? len : 0u, ql = op_consumes_query(op) ? len : 0u;
        // inclusive prefix of the reference / query lengths inside the group
        uint32_t r_in = rl, q_in = ql;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t ur = __shfl_up_sync(0xFFFFFFFFu, r_in, d);
            const uint32_t uq = __shfl_up_sync(0xFFFFFFFFu, q_in, d);
            if ((int)lane >= d) { r_in += ur; q_in += uq; }
        }
        const uint32_t r_off = r_base + r_in - rl, q_off = q_base + q_in - ql;
        // ops with something to deposit
        uint32_t work = __ballot_sync(0xFFFFFFFFu, len != 0 && (op_is_match(op) || op == 2 || op == 3));
        while (work) {
            const int k = __ffs(work) - 1;
            work &= work - 1;
            const uint32_t ck = __shfl_sync(0xFFFFFFFFu, c, k);
            const int64_t r = pos + (int64_t)__shfl_sync(0xFFFFFFFFu, r_off, k);
            const uint32_t qi = __shfl_sync(0xFFFFFFFFu, q_off, k);
            const uint32_t opk = ck & 15u, lenk = ck >> 4;
            if (op_is_match(opk)) {
                if (COMPACT) {
                    const uint32_t r_rel = (uint32_t)(r - pos);
                    for (uint32_t j0 = 0; j0 < lenk; j0 += 32) {
                        const uint32_t j = j0 + lane;
                        uint32_t q = 0;
                        if (j < lenk) q = qual[qi + j];
                        const bool pass = j < lenk && (int)q >= dp.min_bq;
                        const uint32_t m = __ballot_sync(0xFFFFFFFFu, pass);
                        if (m == 0) continue;
                        if (pass) {
                            const uint32_t slot = (ring_head + ring_n + __popc(m & ((1u << lane) - 1u))) & 63u;
                            ring->rr[slot] = r_rel + j; ring->qx[slot] = qi + j; ring->q[slot] = (uint8_t)q;
                        }
                        ring_n += __popc(m);
                        if (ring_n >= 32) {
                            __syncwarp();
                            const uint32_t e = (ring_head + lane) & 63u;
                            const uint32_t qx = ring->qx[e];
                            const uint32_t byte = seq[qx >> 1];
                            deposit_base(tv, dp, pos + ring->rr[e], (qx & 1u) ? (byte & 15u) : (byte >> 4), ring->q[e], ord);
                            ring_head = (ring_head + 32) & 63u;
                            ring_n -= 32;
                            __syncwarp();
                        }
                    }
                } else {
                    for (uint32_t j = lane; j < lenk; j += 32) {
                        const uint32_t q = qual[qi + j];
                        if ((int)q < dp.min_bq) continue;
                        const uint32_t byte = seq[(qi + j) >> 1];
                        const uint32_t nib = ((qi + j) & 1u) ? (byte & 15u) : (byte >> 4);
                        deposit_base(tv, dp, r + j, nib, q, ord);
                    }
                }
            } else if (!dp.replay) {
                // deletion / ref-skip entries are kept iff the NEXT query base passes the quality rule
                // (pysam pileup_base_qual_skip on qpos = y; 0 if qpos >= l_qseq) -- SURVEY B3
                const uint32_t q = (qi < lq) ? (uint32_t)qual[qi] : 0u;
                if ((int)q >= dp.min_bq)
                    for (uint32_t j = lane; j < lenk; j += 32) atomicAdd(&tv.dels[r + j], 1u);
            }
        }
        r_base += __shfl_sync(0xFFFFFFFFu, r_in, 31);
        q_base += __shfl_sync(0xFFFFFFFFu, q_in, 31);
    }
    if (COMPACT) {                                           // what is left in the ring (< 32 entries)
        __syncwarp();
        if (lane < ring_n) {
            const uint32_t e = (ring_head + lane) & 63u;
            const uint32_t qx = ring->qx[e];
            const uint32_t byte = seq[qx >> 1];
            deposit_base(tv, dp, pos + ring->rr[e], (qx & 1u) ? (byte & 15u) : (byte >> 4), ring->q[e], ord);
        }
        __syncwarp();
    }
}

// out of line: what the tiled kernels call for the few reads they hand over
__device__ __noinline__ void deposit_read_warp(const BatchView& b, const TableView& tv, const DepositParams& dp,
                                               uint32_t i, uint32_t lane) {
    deposit_read_warp_impl<false>(b, tv, dp, i, lane);
}

// one thread per read of the batch
__global__ void __launch_bounds__(128) k_deposit_general(const __grid_constant__ BatchView b,
                                                         const __grid_constant__ TableView tv,
                                                         const __grid_constant__ DepositParams dp, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) deposit_read_general(b, tv, dp, i);
}

// one WARP per read of the batch: long reads with many CIGAR ops and a wide quality alphabet (ONT), where neither a
// primary quality nor a short run table exists.  Every read of the batch is in flight at once.
constexpr int kWarpKernelThreads = 256;
__global__ void __launch_bounds__(kWarpKernelThreads) k_deposit_warp(const __grid_constant__ BatchView b,
                                                                     const __grid_constant__ TableView tv,
                                                                     const __grid_constant__ DepositParams dp, uint32_t n) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");     // the tables may still be read by the previous kernel
    __shared__ WarpRing rings[kWarpKernelThreads / 32];
    const uint32_t i = blockIdx.x * (kWarpKernelThreads / 32) + (threadIdx.x >> 5);
    if (i < n) deposit_read_warp_impl<true>(b, tv, dp, i, threadIdx.x & 31u, &rings[threadIdx.x >> 5]);
}

}  // namespace lvc
