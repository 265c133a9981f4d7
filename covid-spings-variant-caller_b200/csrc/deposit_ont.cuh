// Long-read deposit kernel (ONT-like batches: tens of CIGAR ops per read, a wide quality alphabet, most bases below
// the base-quality threshold).
//
// Replaces, for such batches, the per-(column, read) Python loop of live_variant_caller.py:69-70,89-103 and the
// htslib CIGAR walk behind it.  The warp-per-read kernel (deposit_general.cuh: k_deposit_warp) spends ~1,100 warp
// instructions per 520-base read: a CIGAR scan with 21 of 32 lanes busy, a quality test over the whole query
// (soft clips included), a ring of passing bases and a shuffle binary search per passing base to find its op.
// Here nothing is searched.  A CTA (128 threads) owns up to 16 consecutive (coordinate-sorted) reads and works in two
// phases ("k_deposit_ont"):
//   (1) CIGAR phase, EIGHT THREADS PER READ, four ops each, the whole CTA in one pass: per-thread sums of the
//       reference / query lengths, a 3-level scan over the read's 8 lanes for the offsets, a second one for the unit
//       counts.  Every MATCH op becomes a run record (payload offset, length, first column) in a per-read slot table
//       in shared memory and appends the run's UNITS -- the aligned 16-byte pieces of the quality array it touches --
//       to a CTA-wide unit list; deletion / ref-skip ops go to a list of their own.  Nothing is written to the tables,
//       so the phase runs before griddepcontrol.wait and overlaps the previous kernel of the stream;
//   (2) unit phase, every warp on its own, 32 units per step: one aligned 16-byte load of qualities + one 8-byte load
//       of bases per lane (the next step's loads in flight), byte-parallel threshold test, bytes outside the run masked
//       off; the step's PASSING bases (~16 % at minBQ 30) are compacted into the warp's entry list and deposited with
//       every lane busy: one RED into the plane of the base's quality -- the unit knows its run, so the column is an
//       addition.  Soft clips and insertions are never looked at.  First-seen ordinals are tested only for alleles
//       the genotype pass has not yet marked as seen (TableView::seen).
// The (column, allele, quality) histogram of a CTA is sparse (16 reads put ~6 passing bases on a column, spread over
// 61 quality planes x 4 alleles), so a shared-memory count tile would merge almost nothing: the reductions go to L2.
// Counts stay exact for any input: reads with more than 32 ops, other base codes than A/C/G/T, or more units / columns
// than the tables hold take the warp-per-read path inside the same CTA; unknown (group, quality) keys are recorded for
// the replay exactly as in the other kernels.
#pragma once
#include "lvc_common.cuh"
#include "deposit_general.cuh"
#include <limits.h>

namespace lvc {

#ifndef LVC_ONT_THREADS
#define LVC_ONT_THREADS 128
#endif
constexpr int kOntThreads = LVC_ONT_THREADS;
constexpr int kOntWarps = kOntThreads / 32;
constexpr int kOntMaxReads = kOntThreads / 8;   // reads per CTA, at most (the launch picks fewer for longer reads)
constexpr int kOntSlots = 32;                   // one slot per CIGAR op of a read (reads with more ops: warp path)
constexpr int kOntSlotStride = 33;              // slots of consecutive reads are 33 words apart: the 32 threads of the CIGAR
                                                // phase write the same op index of 32 reads without bank conflicts
constexpr uint32_t kOntMaxUnits = kOntThreads * 10;   // units per CTA (more: the remaining reads take the warp path)
constexpr uint32_t kOntMaxWindow = 1u << 30;    // payload window of a CTA addressed with 32-bit offsets
constexpr uint32_t kOntMaxSpan = 1u << 15;      // columns of a CTA addressed with 15 bits in an entry (more: warp path)
constexpr uint32_t kOntNoQual = 0xFFFFFFFFu;    // deletion at the very end of the query: its quality is 0
constexpr uint32_t kOntMaxEntries = kOntThreads * 16;    // passing bases of one step (one unit per thread)

struct OntSmem {
    uint32_t run_q[kOntMaxReads * kOntSlotStride];   // match run: first byte, offset in the CTA's payload window; deletion: the NEXT query byte
    uint32_t run_len[kOntMaxReads * kOntSlotStride];
    int32_t run_ref[kOntMaxReads * kOntSlotStride];  // first reference column of the op
    uint32_t unit[kOntMaxUnits];                // match runs: slot | unit index inside the run << 16
    uint16_t del_list[kOntMaxReads * kOntSlots];  // deletion / ref-skip ops: slot (one entry per op: cannot overflow)
    // per warp: the 32 units of a step as they were loaded, and the step's passing bases as (lane << 4 | byte)
    uint4 st_q[kOntThreads];                    // 16 qualities
    uint2 st_s[kOntThreads];                    // 16 bases (4 bit each, as packed in the batch)
    uint2 st_n[kOntThreads];                    // "seen before" nibbles of the unit's 16 columns
    uint32_t st_c[kOntThreads];                 // column of byte 0 - col_min + 16 | read << 16
    uint16_t entry[kOntMaxEntries];
    uint32_t* plane[128];                       // group 0 (A,C,G,T) plane of quality q < 128, or nullptr
    uint32_t hdr_c0[kOntMaxReads], hdr_nc[kOntMaxReads], hdr_so[kOntMaxReads], hdr_rlen[kOntMaxReads];
    int32_t hdr_pos[kOntMaxReads];
    uint32_t deferred[kOntMaxReads];
    uint32_t n_deferred, n_units, n_valid, n_dels;
    int32_t col_min;
};

#ifndef LVC_ONT_CTAS
#define LVC_ONT_CTAS 8
#endif
__global__ void __launch_bounds__(kOntThreads, LVC_ONT_CTAS)
k_deposit_ont(const __grid_constant__ BatchView b, const __grid_constant__ TableView tv,
              const __grid_constant__ DepositParams dp, uint32_t n, uint32_t reads_per_cta) {
    __shared__ OntSmem sm;
    asm volatile("griddepcontrol.launch_dependents;");
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t r0 = blockIdx.x * reads_per_cta;
    const uint32_t nr_cta = min(reads_per_cta, n - r0);
    // payload window of the CTA: the reads' qualities are contiguous, [seq_off[r0], seq_off[r0 + nr_cta])
    const uint64_t win0 = b.seq_off[r0] & ~15ull;                        // 16-byte aligned
    const uint64_t win1 = b.seq_off[r0 + nr_cta];
    const bool window_ok = win1 - win0 < kOntMaxWindow;
    if (tid == 0) {
        sm.n_deferred = 0; sm.n_units = 0; sm.n_valid = kOntMaxUnits; sm.n_dels = 0; sm.col_min = INT32_MAX;
    }
    __syncthreads();
    for (uint32_t q = tid; q < 128u; q += kOntThreads) {
        const uint16_t pl = tv.lut[q];
        sm.plane[q] = pl == kNoPlane ? nullptr : tv.planes[pl];
    }
    if (tid < nr_cta) {
        // read headers, one read per thread: one round of loads for the whole CTA
        const uint32_t rl = tid, i = r0 + rl;
        const uint32_t keep = b.keep[i];
        const bool ok = read_passes_filter(b.flag[i], b.mapq[i], keep, dp.min_mq);
        const uint32_t c0 = b.cigar_off[i];
        const int32_t pos = b.pos[i];
        sm.hdr_c0[rl] = c0;
        // 0 ops: nothing to do for this read.  Reads without the "every base is A/C/G/T" hint (keep bit 1) take the
        // warp path: the unit phase then never meets another base code (marked by 33+ ops here)
        sm.hdr_nc[rl] = ok ? ((keep & 2u) ? b.cigar_off[i + 1] - c0 : max(b.cigar_off[i + 1] - c0, 33u)) : 0u;
        sm.hdr_so[rl] = (uint32_t)(b.seq_off[i] - win0);
        sm.hdr_pos[rl] = pos;
        sm.hdr_rlen[rl] = 0;                                              // > 0: coverage to deposit after the CIGAR phase
        if (ok) atomicMin(&sm.col_min, pos);
    }
    __syncthreads();
    const int32_t col_min = sm.col_min;

    // ---- (1) CIGAR phase: EIGHT THREADS PER READ, four ops each: the CTA's 32 reads are walked in one pass by all
    //      256 threads (~45 warp instructions per read where a warp per read with lane = op spent 200, and no warp
    //      idles as with a thread per read).  Each thread sums the reference / query lengths of its four ops, a 3-level
    //      scan over the read's 8 lanes gives the offsets, a second one the unit counts; then every thread writes the
    //      run records and the unit-list entries of its own ops.  Nothing is written to the tables here, so the phase
    //      overlaps the previous kernel of the stream.
    {
        static_assert(kOntThreads == 8 * kOntMaxReads && kOntSlots == 32, "8 threads x 4 ops per read");
        const uint32_t rl = tid >> 3, part = tid & 7u;
        const uint32_t nc = rl < nr_cta ? sm.hdr_nc[rl] : 0u;
        const bool fits = nc != 0 && nc <= (uint32_t)kOntSlots && window_ok && !dp.replay;   // the same for the read's 8 lanes
        const uint32_t c0 = sm.hdr_c0[rl < nr_cta ? rl : 0], so = sm.hdr_so[rl < nr_cta ? rl : 0];
        const int64_t pos = sm.hdr_pos[rl < nr_cta ? rl : 0];
        uint32_t cg[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t k = part * 4u + j;
            cg[j] = (fits && k < nc) ? b.cigar[c0 + k] : 0u;                  // padding: a match of length 0
        }
        uint32_t r4 = 0, q4 = 0;
        bool huge = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t op = cg[j] & 15u, len = cg[j] >> 4;
            r4 += op_consumes_ref(op) ? len : 0u;
            q4 += op_consumes_query(op) ? len : 0u;
            huge |= len >= (1u << 19);
        }
        uint32_t r_in = r4, q_in = q4;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            const uint32_t ur = __shfl_up_sync(0xFFFFFFFFu, r_in, d, 8);
            const uint32_t uq = __shfl_up_sync(0xFFFFFFFFu, q_in, d, 8);
            if ((int)part >= d) { r_in += ur; q_in += uq; }
        }
        const uint32_t rlen = __shfl_sync(0xFFFFFFFFu, r_in, 7, 8), lq = __shfl_sync(0xFFFFFFFFu, q_in, 7, 8);
        // units of this thread's ops: a match run touches some aligned 16-byte groups of the window; a deletion is one
        uint32_t u4 = 0;
        {
            uint32_t qo = q_in - q4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t op = cg[j] & 15u, len = cg[j] >> 4;
                if (len != 0) {
                    if (op_is_match(op)) u4 += (((so + qo) & 15u) + len + 15u) >> 4;
                    else if (op == 2 || op == 3) u4 += 1u << 16;                 // deletions counted in the high half
                }
                qo += op_consumes_query(op) ? len : 0u;
            }
        }
        uint32_t u_in = u4;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            const uint32_t uu = __shfl_up_sync(0xFFFFFFFFu, u_in, d, 8);
            if ((int)part >= d) u_in += uu;
        }
        const uint32_t ud_tot = __shfl_sync(0xFFFFFFFFu, u_in, 7, 8);
        const uint32_t u_tot = ud_tot & 0xFFFFu, d_tot = ud_tot >> 16;
        const bool any_huge = ((__ballot_sync(0xFFFFFFFFu, huge) >> (lane & 24u)) & 255u) != 0u;
        // anything the tables cannot hold goes to the warp-per-read path (which does its own coverage / deletions)
        const bool big = nc != 0 && (!fits || any_huge || u_tot > kOntMaxUnits || (uint64_t)(pos - col_min) + rlen >= kOntMaxSpan);
        // rlen == 0: no M/D/N/=/X op: htslib asserts on such records; skipped
        const bool ok = nc != 0 && !big && rlen != 0;
        const bool range_err = ok && (pos < 0 || pos + (int64_t)rlen > tv.G);
        uint32_t u_base = 0, d_base = 0;
        if (ok && !range_err && part == 0) { u_base = atomicAdd(&sm.n_units, u_tot); d_base = atomicAdd(&sm.n_dels, d_tot); }
        u_base = __shfl_sync(0xFFFFFFFFu, u_base, 0, 8);
        d_base = __shfl_sync(0xFFFFFFFFu, d_base, 0, 8);
        const bool full = ok && !range_err && u_base + u_tot > kOntMaxUnits;   // list full: this read and every later one
        if (part == 0) {
            if (full) atomicMin(&sm.n_valid, u_base);
            if (big || full) sm.deferred[atomicAdd(&sm.n_deferred, 1u)] = rl;
            else if (range_err) sm.hdr_rlen[rl] = 0xFFFFFFFFu;                 // out of range: reported after the phase
            else if (ok) sm.hdr_rlen[rl] = rlen;
        }
        if (ok && !range_err) {
            uint32_t* up = sm.unit + u_base + ((u_in - u4) & 0xFFFFu);
            uint16_t* dl = sm.del_list + d_base + ((u_in - u4) >> 16);
            uint32_t r_off = r_in - r4, q_off = q_in - q4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t op = cg[j] & 15u, len = cg[j] >> 4;
                const uint32_t slot = rl * kOntSlotStride + part * 4u + j;
                if (len != 0 && op_is_match(op) && !full) {
                    const uint32_t rq = so + q_off;
                    sm.run_q[slot] = rq;
                    sm.run_len[slot] = len;
                    sm.run_ref[slot] = (int32_t)(pos + r_off);
                    const uint32_t n_u = ((rq & 15u) + len + 15u) >> 4;
                    for (uint32_t k = 0; k < n_u; ++k) *up++ = slot | (k << 16);
                } else if (len != 0 && (op == 2 || op == 3)) {
                    // deletion / ref-skip entries are kept iff the NEXT query base passes the quality rule (pysam
                    // pileup_base_qual_skip on qpos = y; quality 0 if qpos >= l_qseq) -- SURVEY B3; tested after the unit
                    // phase.  (A read the unit list could not take leaves its reserved entries empty.)
                    sm.run_q[slot] = q_off < lq ? so + q_off : kOntNoQual;
                    sm.run_len[slot] = len;
                    sm.run_ref[slot] = (int32_t)(pos + r_off);
                    *dl++ = full ? (uint16_t)0xFFFFu : (uint16_t)slot;
                }
                r_off += op_consumes_ref(op) ? len : 0u;
                q_off += op_consumes_query(op) ? len : 0u;
            }
        }
    }
    __syncthreads();
    // the tables may still be read by the previous kernel of the stream (programmatic stream serialization)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tid < nr_cta) {
        // coverage difference array of the reads this CTA deposits itself
        const uint32_t rlen = sm.hdr_rlen[tid];
        if (rlen == 0xFFFFFFFFu) atomicAdd(&tv.status[ST_RANGE_ERR], 1u);
        else if (rlen) {
            const int64_t pos = sm.hdr_pos[tid];
            atomicAdd(&tv.covdiff[pos], 1);
            atomicAdd(&tv.covdiff[pos + (int64_t)rlen], -1);
        }
    }

    // ---- (2) unit phase.  Each warp works on its own: 32 units per step (one per lane): test, COMPACT the step's
    //      passing bases into the warp's entry list (a short loop over the set bits of a 16-bit mask: two integer
    //      instructions and a store per base), then deposit the list with every lane busy -- the quality, the base and
    //      the column of an entry are read back from the warp's staging slots.  No CTA barrier; the next step's loads
    //      are in flight while this one is processed.
    {
        const uint32_t n_units = min(sm.n_units, sm.n_valid);
        const uint8_t* qbase = b.qual + win0;
        const uint8_t* sbase = b.seq4 + (win0 >> 1);
        uint32_t* first0 = tv.first[0];
        const uint32_t* seen = tv.seen;
        const int mbq = dp.min_bq;
        struct Unit {
            uint4 q4;            // 16 qualities
            uint2 sraw;          // 16 bases, 4 bit each
            uint32_t w0, w1, w2; // "allele seen in an earlier batch" nibbles of the columns around the unit
            uint32_t slot, len;
            int32_t lo, hi;      // bytes [lo, hi) of the unit belong to the run
            int32_t col0;        // column of the unit's byte 0
        };
        auto fetch = [&](uint32_t u, Unit& x) {
            const uint32_t e = sm.unit[u];
            x.slot = e & 0xFFFFu;
            const uint32_t rq = sm.run_q[x.slot];
            x.len = sm.run_len[x.slot];
            const int32_t ref0 = sm.run_ref[x.slot];
            const uint32_t A = (rq & ~15u) + ((e >> 16) << 4);                  // window offset of the unit's byte 0
            x.q4 = __ldcs(reinterpret_cast<const uint4*>(qbase + A));
            x.sraw = __ldcs(reinterpret_cast<const uint2*>(sbase + (A >> 1)));
            x.lo = (int32_t)rq - (int32_t)A;                                    // > 0 only in the run's first unit
            x.hi = (int32_t)(rq + x.len) - (int32_t)A;                          // < 16 only in its last unit
            x.col0 = ref0 - x.lo;
            x.w0 = x.w1 = x.w2 = 0;
            if (seen) {
                const uint32_t* sw = seen + (x.col0 >> 3);                      // two words of padding in front: col0 >= -15
                x.w0 = sw[0]; x.w1 = sw[1]; x.w2 = sw[2];
            }
        };
        const uint32_t wbase = warp * 32u;                                      // the warp's staging slots
        uint16_t* ent = sm.entry + warp * 512u;                                 // and its entry list (32 units x 16 bases)
        const uint8_t* stq = reinterpret_cast<const uint8_t*>(sm.st_q + wbase);
        const uint8_t* sts = reinterpret_cast<const uint8_t*>(sm.st_s + wbase);
        const uint32_t* stn = reinterpret_cast<const uint32_t*>(sm.st_n + wbase);
        // deletion / ref-skip entries: the one quality that decides each of them is requested now and used after the unit
        // loop (its latency is hidden; the first 256 entries cover a CTA of 520-base ONT reads)
        const uint32_t n_dels = sm.n_dels;
        uint32_t d_slot = 0xFFFFu, d_q = 0;
        if (tid < n_dels) {
            d_slot = sm.del_list[tid];
            if (d_slot != 0xFFFFu) { const uint32_t rq = sm.run_q[d_slot]; d_q = rq == kOntNoQual ? 0u : (uint32_t)qbase[rq]; }
        }
        Unit cur, nxt;
        if (tid < n_units) fetch(tid, cur);
        for (uint32_t ub = 0; ub + wbase < n_units; ub += kOntThreads) {
            const bool have = ub + tid < n_units;
            const bool have_n = ub + kOntThreads + tid < n_units;
            if (have_n) fetch(ub + kOntThreads + tid, nxt);
            uint32_t m16 = 0;
            if (have) {
                {
                    const uint32_t f0 = ge_flags4(cur.q4.x, mbq), f1 = ge_flags4(cur.q4.y, mbq), f2 = ge_flags4(cur.q4.z, mbq),
                                   f3 = ge_flags4(cur.q4.w, mbq);
                    // 0x80 per passing byte -> one bit per base
                    m16 = ((((f0 >> 7) * 0x00204081u) >> 21) & 15u) | (((((f1 >> 7) * 0x00204081u) >> 21) & 15u) << 4) |
                          (((((f2 >> 7) * 0x00204081u) >> 21) & 15u) << 8) | (((((f3 >> 7) * 0x00204081u) >> 21) & 15u) << 12);
                    if (cur.lo > 0) m16 &= 0xFFFFu << cur.lo;
                    if (cur.hi < 16) m16 &= 0xFFFFu >> (16 - cur.hi);
                    if ((cur.q4.x | cur.q4.y | cur.q4.z | cur.q4.w) & 0x80808080u) {
                        // a quality >= 128 (no real file): this unit's bases one by one, nothing to compact
                        const uint32_t ord = dp.ord_base + r0 + cur.slot / kOntSlotStride;
                        while (m16) {
                            const uint32_t j = (uint32_t)__ffs(m16) - 1u;
                            m16 &= m16 - 1u;
                            const uint32_t qw = j < 8u ? (j < 4u ? cur.q4.x : cur.q4.y) : (j < 12u ? cur.q4.z : cur.q4.w);
                            const uint32_t sw = j < 8u ? cur.sraw.x : cur.sraw.y;
                            const uint32_t byte = (sw >> ((j & 6u) * 4u)) & 255u;
                            deposit_base(tv, dp, (int64_t)cur.col0 + j, (j & 1u) ? (byte & 15u) : (byte >> 4),
                                         (qw >> ((j & 3u) * 8u)) & 255u, ord);
                        }
                    }
                }
            }
            if (m16) {
                // stage what the deposit needs: qualities, bases, seen nibbles (column j at bits 4j of y:x), column, read
                sm.st_q[tid] = cur.q4;
                sm.st_s[tid] = cur.sraw;
                const uint32_t sh = ((uint32_t)cur.col0 & 7u) * 4u;
                sm.st_n[tid] = make_uint2(__funnelshift_r(cur.w0, cur.w1, sh), __funnelshift_r(cur.w1, cur.w2, sh));
                sm.st_c[tid] = (uint32_t)(cur.col0 - col_min + 16) | ((cur.slot / kOntSlotStride) << 16);
            }
            // where this lane's entries go: warp scan of the counts
            const uint32_t cnt = __popc(m16);
            uint32_t c_in = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t uu = __shfl_up_sync(0xFFFFFFFFu, c_in, d);
                if ((int)lane >= d) c_in += uu;
            }
            const uint32_t n_ent = __shfl_sync(0xFFFFFFFFu, c_in, 31);
            {
                uint16_t* ep = ent + (c_in - cnt);
                const uint32_t tag = lane << 4;
                while (m16) {
                    *ep++ = (uint16_t)(tag | ((uint32_t)__ffs(m16) - 1u));
                    m16 &= m16 - 1u;
                }
            }
            __syncwarp();
            for (uint32_t i = lane; i < n_ent; i += 32) {
                const uint32_t e = ent[i];
                const uint32_t t = e >> 4, j = e & 15u;
                const uint32_t q = stq[t * 16u + j];                                      // < 128 (tested with the unit)
                const uint32_t byte = sts[t * 8u + (j >> 1)];
                const uint32_t nib = (j & 1u) ? (byte & 15u) : (byte >> 4);               // A,C,G,T = 1,2,4,8 (read hint)
                const uint32_t sn = (stn[t * 2u + (j >> 3)] >> ((j & 7u) * 4u)) & nib;
                const uint32_t uc = sm.st_c[wbase + t];
                const uint32_t col = (uint32_t)col_min + (uc & 0xFFFFu) - 16u + j;
                uint32_t* pl = sm.plane[q];
                if (pl) {
                    const uint32_t cell = col * 4u + ((nib >> 1) - (nib >> 3));           // slot 0..3
                    atomicAdd(pl + cell, 1u);
                    // First-seen ordinal.  An allele the genotype pass found present after an earlier batch cannot get a
                    // smaller ordinal from this one: its bit in `seen` is set and nothing is read.  Otherwise test before
                    // reducing (all quality planes of a group share the cell; a stale L1 line is safe: it only decreases).
                    if (!sn) {
                        const uint32_t ord = dp.ord_base + r0 + (uc >> 16);
                        uint32_t* f = first0 + cell;
                        if (*f > ord) atomicMin(f, ord);
                    }
                } else {
                    deposit_base(tv, dp, (int64_t)col, nib, q, dp.ord_base + r0 + (uc >> 16));   // a quality without a plane yet
                }
            }
            __syncwarp();
            cur = nxt;
        }
        for (uint32_t i = tid; i < n_dels; i += kOntThreads) {
            if (i != tid) {
                d_slot = sm.del_list[i];
                if (d_slot != 0xFFFFu) { const uint32_t rq = sm.run_q[d_slot]; d_q = rq == kOntNoQual ? 0u : (uint32_t)qbase[rq]; }
            }
            // kept iff the NEXT query base passes the quality rule
            if (d_slot != 0xFFFFu && (int)d_q >= mbq) {
                uint32_t* d = tv.dels + sm.run_ref[d_slot];
                const uint32_t len = sm.run_len[d_slot];
                for (uint32_t j = 0; j < len; ++j) atomicAdd(d + j, 1u);
            }
        }
    }
    // ---- reads the tables could not hold: the warp-per-read path, one warp each (the list is complete since the barrier
    //      after the CIGAR phase: no barrier here, a warp that is done with its units leaves)
    const uint32_t n_def = sm.n_deferred;
    for (uint32_t d = warp; d < n_def; d += kOntWarps) deposit_read_warp(b, tv, dp, r0 + sm.deferred[d], lane);
}

}  // namespace lvc
