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
    uint32_t* f = tv.first[gs >> 2];
    if (f[cell] > ord) atomicMin(&f[cell], ord);
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

// The same walk with one WARP per read: the lanes stride over the bases of each match op (coalesced loads,
// 32 reductions in flight) while the CIGAR walk itself is warp-uniform.  Used by the tiled kernel for the
// reads it cannot take.
__device__ __noinline__ void deposit_read_warp(const BatchView& b, const TableView& tv, const DepositParams& dp,
                                               uint32_t i, uint32_t lane) {
    if (!read_passes_filter(b.flag[i], b.mapq[i], b.keep[i], dp.min_mq)) return;
    const uint32_t c0 = b.cigar_off[i], c1 = b.cigar_off[i + 1];
    int64_t rlen = 0;
    uint32_t lq = 0;
    for (uint32_t k = c0; k < c1; ++k) {
        const uint32_t c = b.cigar[k], op = c & 15u, len = c >> 4;
        if (op_consumes_ref(op)) rlen += len;
        if (op_consumes_query(op)) lq += len;
    }
    if (rlen == 0) return;
    const int64_t pos = b.pos[i];
    if (pos < 0 || pos + rlen > tv.G) {
        if (lane == 0) atomicAdd(&tv.status[ST_RANGE_ERR], 1u);
        return;
    }
    if (!dp.replay && lane == 0) {
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
            for (uint32_t j = lane; j < len; j += 32) {
                const uint32_t q = qual[qi + j];
                if ((int)q < dp.min_bq) continue;
                const uint32_t byte = seq[(qi + j) >> 1];
                const uint32_t nib = ((qi + j) & 1u) ? (byte & 15u) : (byte >> 4);
                deposit_base(tv, dp, r + j, nib, q, ord);
            }
            qi += len; r += len;
        } else if (op == 2 || op == 3) {
            const uint32_t q = (qi < lq) ? (uint32_t)qual[qi] : 0u;
            if (!dp.replay && (int)q >= dp.min_bq)
                for (uint32_t j = lane; j < len; j += 32) atomicAdd(&tv.dels[r + j], 1u);
            r += len;
        } else if (op == 1 || op == 4) {
            qi += len;
        }
    }
}

// one thread per read of the batch
__global__ void __launch_bounds__(128) k_deposit_general(BatchView b, TableView tv, DepositParams dp, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) deposit_read_general(b, tv, dp, i);
}

}  // namespace lvc
