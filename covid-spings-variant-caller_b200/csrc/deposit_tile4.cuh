// Tiled deposit kernel, generation 4: 4-bit KEY staging.
//
// Same decomposition as deposit_tile.cuh (chunk of 256 coordinate-sorted reads per CTA, match-run table,
// (32-column slab, group of kTaskRuns runs) tasks, register SWAR counters, one RED per (column, allele) per chunk),
// but the staged payload is no longer the raw 1.5 bytes per base: while it is copied from global memory (16-byte
// coalesced loads) each base is reduced to a 4-bit KEY = its one-hot nibble if its quality equals the batch's
// primary quality, else 0.  Bases with another passing quality are deposited right there (exact; their run is found
// by binary search).  Consequences:
//   * shared memory per CTA drops from 73 KB to 36 KB and the kernel is held to 48 registers: 5 CTAs per SM overlap
//     the per-chunk latency phases (header loads -> CIGAR loads -> payload -> flush);
//   * the inner loop reads one packed run record and 3 shared key words per 16 bases and needs no quality arithmetic:
//     counts for allele c are  (keys >> c) & 0x11111111  -- 8 bases per instruction.
// Launched with programmatic stream serialization: everything before `griddepcontrol.wait` only reads the batch.
// What was tried and measured around this kernel is listed in DESIGN.md sections 3.4 and 4.
#pragma once
#include "deposit_tile.cuh"

namespace lvc {

#ifndef LVC_CTAS_PER_SM
#define LVC_CTAS_PER_SM 5
#endif
constexpr int kTile4CtasPerSM = LVC_CTAS_PER_SM;
#ifndef LVC_TASK_RUNS
#define LVC_TASK_RUNS 128
#endif
constexpr int kTaskRuns = LVC_TASK_RUNS;         // runs per task (passes of 16 runs; the 4-bit counters hold <= 15 passes)
static_assert(kTaskRuns % 32 == 0 && kTaskRuns <= 224, "4-bit counters: at most 15 units per lane");
constexpr uint32_t kKeyBasesPerWin = kTileReads * 160u;              // bases per staged window (stride)
constexpr uint32_t kKeyCapBases = kKeyBasesPerWin + kMaxReadBytes;   // + one longest tileable read

struct Tile4Smem {
    static constexpr uint32_t key_off = 0;                                     // 4-bit keys, little-endian nibble order
    static constexpr uint32_t key_bytes = kSlack + kKeyCapBases / 2 + 16 + kSlack;
    static constexpr uint32_t tab_off = key_off + key_bytes;                   // u32 [kTabCols*4]
    static constexpr uint32_t tab_bytes = kTabCols * 4 * 4;
    static constexpr uint32_t pos_off = tab_off + tab_bytes;                   // i32 [kMaxRuns]
    static constexpr uint32_t qo_off = pos_off + kMaxRuns * 4;                 // u32 [kMaxRuns]
    static constexpr uint32_t len_off = qo_off + kMaxRuns * 4;                 // u16 [kMaxRuns]
    static constexpr uint32_t rd_off = len_off + kMaxRuns * 2;                 // u16 [kMaxRuns]
    static constexpr uint32_t rix_off = rd_off + kMaxRuns * 2;                 // u16 [kMaxRuns]
    static constexpr uint32_t items_off = rix_off + kMaxRuns * 2;              // u16 [kTabCols*4]
    static constexpr uint32_t dlist_off = items_off + kTabCols * 4 * 2;        // u16 [kTileReads]
    static constexpr uint32_t slab_a_off = dlist_off + kTileReads * 2;         // u32 [kMaxSlabs]
    static constexpr uint32_t slab_pre_off = slab_a_off + kMaxSlabs * 4;       // u32 [kMaxSlabs+1]
    static constexpr uint32_t slab_n_off = slab_pre_off + (kMaxSlabs + 1) * 4; // u32 [kMaxSlabs]
    static constexpr uint32_t misc_off = (slab_n_off + kMaxSlabs * 4 + 15) & ~15u;
    static constexpr uint32_t pk_off = misc_off + 256;                         // uint2 [kMaxRuns]: what the pass loop reads
    static constexpr uint32_t total = pk_off + kMaxRuns * 8;
};
constexpr size_t kTile4SmemBytes = Tile4Smem::total;
static_assert((kTile4SmemBytes + 1024) * kTile4CtasPerSM <= 227 * 1024, "the intended CTAs per SM must fit");

// one (run, 8 columns) unit: 8 keys (4 bit each) from two aligned shared words
struct PassUnit4 {
    uint32_t k, k1;          // 16 keys (4 bit each)
    int32_t j, len;
};
__device__ __forceinline__ PassUnit4 pass_load4(uint32_t k_smem, int32_t ka, int32_t j, int32_t len) {
    PassUnit4 u;
    u.j = j; u.len = len;
    const uint32_t a = k_smem + (uint32_t)((ka >> 3) << 2);
    const uint32_t x0 = lds32(a), x1 = lds32(a + 4), x2 = lds32(a + 8);
    const uint32_t sh = (uint32_t)(ka & 7) * 4u;
    u.k = __funnelshift_r(x0, x1, sh);
    u.k1 = __funnelshift_r(x1, x2, sh);
    return u;
}
// PRMT in its default mode: selector nibbles with bit 3 set replicate the SIGN of the selected byte over the
// result byte (0xFF / 0x00).  (__byte_perm masks that bit off, hence the inline PTX.)
__device__ __forceinline__ uint32_t prmt_sign(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// (a & c) | (b & ~c) in one LOP3
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// byte flags (0x80 per byte) of 4 bases -> nibble mask (0xF per passing base) in bits 0..15
__device__ __forceinline__ uint32_t flags_to_nibbles(uint32_t f80) {
    const uint32_t m = f80 >> 7;                                   // bits 0, 8, 16, 24
    const uint32_t x = (m | (m >> 4)) & 0x00110011u;               // bits 0, 4, 16, 20
    return ((x | (x >> 8)) & 0x1111u) * 15u;                       // nibbles 0..3
}

// kernel parameters stay in the constant bank even where their address is taken (the warp-per-read helper takes
// the views by reference): without this every thread copies them to local memory first
#define LVC_GC const __grid_constant__
template <bool GE_ALL>
__global__ void __launch_bounds__(kTileThreads, kTile4CtasPerSM)
k_deposit_tile4(LVC_GC BatchView b, LVC_GC TableView tv, LVC_GC DepositParams dp, LVC_GC TileParams tp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem + Tile4Smem::tab_off);
    int32_t* s_pos = reinterpret_cast<int32_t*>(smem + Tile4Smem::pos_off);
    uint32_t* s_qo = reinterpret_cast<uint32_t*>(smem + Tile4Smem::qo_off);
    uint16_t* s_len = reinterpret_cast<uint16_t*>(smem + Tile4Smem::len_off);
    uint16_t* s_rd = reinterpret_cast<uint16_t*>(smem + Tile4Smem::rd_off);
    uint16_t* s_rix = reinterpret_cast<uint16_t*>(smem + Tile4Smem::rix_off);
    uint16_t* s_items = reinterpret_cast<uint16_t*>(smem + Tile4Smem::items_off);
    uint16_t* s_dlist = reinterpret_cast<uint16_t*>(smem + Tile4Smem::dlist_off);
    uint32_t* s_slab_a = reinterpret_cast<uint32_t*>(smem + Tile4Smem::slab_a_off);
    uint32_t* s_slab_pre = reinterpret_cast<uint32_t*>(smem + Tile4Smem::slab_pre_off);
    uint32_t* s_slab_n = reinterpret_cast<uint32_t*>(smem + Tile4Smem::slab_n_off);
    uint32_t* s_misc = reinterpret_cast<uint32_t*>(smem + Tile4Smem::misc_off);
    uint2* s_pk = reinterpret_cast<uint2*>(smem + Tile4Smem::pk_off);
    // s_misc: [0,1] mbarrier  [2] task counter  [3] n_items  [4] window max end column (long reads only)  [6] deferred reads
    //         [5] run table end (overflow only)
    //         [8..11] min read byte, max read end byte, max reference span, max end column   [32..39] runs per warp
    const uint32_t k_smem = sbase + Tile4Smem::key_off + kSlack;      // staged keys start here

    // let a kernel launched with programmatic stream serialization (the genotype pass) become resident while this
    // grid drains; it still waits for this grid's completion before reading
    asm volatile("griddepcontrol.launch_dependents;");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t qprim4 = tp.qprim * 0x01010101u;
    const int mbq = dp.min_bq < 1 ? 1 : (dp.min_bq > 128 ? 128 : dp.min_bq);
    const uint32_t ge_add4 = (uint32_t)(0x80 - mbq) * 0x01010101u;

    // ---- (1) this chunk's read headers: issue the global loads first, then set up shared memory
    const uint32_t cur = blockIdx.x;
    ReadHdr hd;
    uint64_t so0;
    hdr_load1(b, cur, tid, hd, so0);
    uint32_t* sc = s_misc + 8;                                       // per-chunk scalars
    uint32_t* wc = s_misc + 32;                                      // runs per warp
    if (tid == 0) {
        s_misc[3] = 0; s_misc[6] = 0;
        sc[0] = 0xFFFFFFFFu; sc[1] = 0; sc[2] = 0; sc[3] = 0;
    }
    for (int k = tid; k < kTabCols * 4; k += kTileThreads) s_tab[k] = 0;
    hdr_load2(b, hd, dp.min_mq);       // CIGAR ops, only for reads that pass the read-level filter
    // a chunk in which no read passes the read-level filter (everything dropped by the depth cap) ends here
    if (!__syncthreads_or(read_passes_filter(hd.flag, hd.mapq, hd.keep, dp.min_mq))) return;
    // Up to here only the batch was read.  The tables may still be in use by the previous kernel of the stream (this
    // kernel is launched with programmatic stream serialization): wait for it before the first table access.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // byte extent of the reads that pass the read-level filter (a superset of what will be deposited):
    // known before the CIGARs arrive, so the bulk copy overlaps classification
    const uint32_t so_rel = hd.so - (uint32_t)so0, so1_rel = hd.so1 - (uint32_t)so0;
    {
        const bool pass = read_passes_filter(hd.flag, hd.mapq, hd.keep, dp.min_mq) && (hd.keep & 2u) &&
                          (so1_rel - so_rel) <= kMaxReadBytes && so1_rel > so_rel;
        uint32_t lo = pass ? so_rel : 0xFFFFFFFFu, hi = pass ? so1_rel : 0u;
        lo = __reduce_min_sync(0xFFFFFFFFu, lo);
        hi = __reduce_max_sync(0xFFFFFFFFu, hi);
        if (lane == 0 && hi) { atomicMin(&sc[0], lo); atomicMax(&sc[1], hi); }
    }
    __syncthreads();
    const uint32_t min_rel = sc[0], max_rel = sc[1];
    if (max_rel == 0) {
        // no read of this chunk can take the tiled path: hand over what must be deposited and leave
        if (read_passes_filter(hd.flag, hd.mapq, hd.keep, dp.min_mq) && hd.nc) {
            bool any_ref = false;
            for (uint32_t k = 0; k < hd.nc; ++k) any_ref |= op_consumes_ref(b.cigar[hd.c0 + k] & 15u);
            if (any_ref) s_dlist[atomicAdd(&s_misc[6], 1u)] = (uint16_t)tid;
        }
        __syncthreads();
        const uint32_t n_def = s_misc[6];
        for (uint32_t d = warp; d < n_def; d += kTileWarps) deposit_read_warp(b, tv, dp, cur * kTileReads + s_dlist[d], lane);
        return;
    }
    // staging base: 16-byte aligned start of the first such read
    const uint64_t base_abs = (so0 + min_rel) & ~15ull;
    const uint32_t base_rel = (uint32_t)(base_abs - so0);            // may wrap below zero: used mod 2^32
    {
        const uint32_t chunk0 = cur * kTileReads;
        // ---- (2) classify this thread's read: filter, match runs, deletion entries
        int32_t run_pos[kMaxRunsPerRead] = {0, 0, 0};
        uint32_t run_len[kMaxRunsPerRead] = {0, 0, 0}, run_q[kMaxRunsPerRead] = {0, 0, 0};
        int32_t del_pos[kMaxDelsPerRead] = {0, 0};
        uint32_t del_len[kMaxDelsPerRead] = {0, 0}, del_q[kMaxDelsPerRead] = {0, 0};
        uint32_t nr = 0, nd = 0, rspan = 0;
        bool defer = false;
        const bool rpass = read_passes_filter(hd.flag, hd.mapq, hd.keep, dp.min_mq);
        // warp-uniform shortcut: every read of the warp that passes the filter is one match op (150M): no CIGAR walk
        const uint32_t l0 = hd.cg0 >> 4;
        const bool one_m = hd.nc == 1 && op_is_match(hd.cg0 & 15u) && (hd.keep & 2u) && (hd.so1 - hd.so) <= kMaxReadBytes &&
                           l0 >= 1u && l0 <= 65535u;
        if (__all_sync(0xFFFFFFFFu, !rpass || one_m)) {
            if (rpass) {
                if (hd.pos < 0 || (int64_t)hd.pos + l0 > tv.G) atomicAdd(&tv.status[ST_RANGE_ERR], 1u);
                else { nr = 1; run_pos[0] = hd.pos; run_len[0] = l0; run_q[0] = 0; rspan = l0; }
            }
        } else if (__all_sync(0xFFFFFFFFu, !rpass || (hd.nc >= 1u && hd.nc <= 3u && (hd.keep & 2u) &&
                                                        (hd.so1 - hd.so) <= kMaxReadBytes))) {
            // second warp-uniform shortcut: at most 3 CIGAR ops per read (one indel or soft clips), all of them
            // already in registers: straight-line code, same results as the general walk below
            if (rpass) {
                const bool p1 = hd.nc > 1u, p2 = hd.nc > 2u;
                const uint32_t o0 = hd.cg0 & 15u, o1 = hd.cg1 & 15u, o2 = hd.cg2 & 15u;
                const uint32_t n0 = hd.cg0 >> 4, n1 = p1 ? hd.cg1 >> 4 : 0u, n2 = p2 ? hd.cg2 >> 4 : 0u;
                const bool m0 = op_is_match(o0), m1 = p1 && op_is_match(o1), m2 = p2 && op_is_match(o2);
                const bool d0 = o0 == 2 || o0 == 3, d1 = p1 && (o1 == 2 || o1 == 3), d2 = p2 && (o2 == 2 || o2 == 3);
                const uint32_t qc0 = op_consumes_query(o0) ? n0 : 0u, qc1 = (p1 && op_consumes_query(o1)) ? n1 : 0u,
                               qc2 = (p2 && op_consumes_query(o2)) ? n2 : 0u;
                const uint32_t rc0 = (m0 || d0) ? n0 : 0u, rc1 = (m1 || d1) ? n1 : 0u, rc2 = (m2 || d2) ? n2 : 0u;
                const uint32_t qo1 = qc0, qo2 = qc0 + qc1, lq = qo2 + qc2;
                const uint32_t ro1 = rc0, ro2 = rc0 + rc1;
                rspan = ro2 + rc2;
                // match runs: a run starts at a match op that does not follow a match op
                const bool s0 = m0, s1 = m1 && !m0, s2 = m2 && !m1;
                const uint32_t len2 = n2, len1 = n1 + (m2 ? n2 : 0u), len0 = n0 + (m1 ? len1 : 0u);
                nr = (uint32_t)s0 + (uint32_t)s1 + (uint32_t)s2;
                if (s0) { run_pos[0] = hd.pos; run_len[0] = len0; run_q[0] = 0; }
                else if (s1) { run_pos[0] = hd.pos + (int32_t)ro1; run_len[0] = len1; run_q[0] = qo1; }
                else if (s2) { run_pos[0] = hd.pos + (int32_t)ro2; run_len[0] = len2; run_q[0] = qo2; }
                if (s0 && s2) { run_pos[1] = hd.pos + (int32_t)ro2; run_len[1] = len2; run_q[1] = qo2; }
                // deletion / ref-skip entries, in op order; their quality (the NEXT query base, 0 past the end) is
                // requested here and tested where the entries are deposited
                const uint32_t ndel = (uint32_t)d0 + (uint32_t)d1 + (uint32_t)d2;
                bool tileable = ndel <= (uint32_t)kMaxDelsPerRead;
                if (tileable && ndel) {
                    const uint8_t* qrd = b.qual + so0 + so_rel;
                    if (d0) { del_pos[0] = hd.pos; del_len[0] = n0; del_q[0] = 0u < lq ? (uint32_t)qrd[0] : 0u; nd = 1; }
                    if (d1) {
                        const uint32_t qv = qo1 < lq ? (uint32_t)qrd[qo1] : 0u;
                        if (nd == 0) { del_pos[0] = hd.pos + (int32_t)ro1; del_len[0] = n1; del_q[0] = qv; }
                        else { del_pos[1] = hd.pos + (int32_t)ro1; del_len[1] = n1; del_q[1] = qv; }
                        ++nd;
                    }
                    if (d2) {
                        const uint32_t qv = qo2 < lq ? (uint32_t)qrd[qo2] : 0u;
                        if (nd == 0) { del_pos[0] = hd.pos + (int32_t)ro2; del_len[0] = n2; del_q[0] = qv; }
                        else { del_pos[1] = hd.pos + (int32_t)ro2; del_len[1] = n2; del_q[1] = qv; }
                        ++nd;
                    }
                }
                bool any_ref = m0 || m1 || m2 || d0 || d1 || d2;
                if (tileable && any_ref && (hd.pos < 0 || (int64_t)hd.pos + rspan > tv.G)) {
                    atomicAdd(&tv.status[ST_RANGE_ERR], 1u);
                    nr = 0; nd = 0; rspan = 0; any_ref = false;
                }
                if ((nr >= 1 && run_len[0] > 65535u) || (nr >= 2 && run_len[1] > 65535u)) tileable = false;
                if (!tileable) { defer = any_ref; nr = 0; nd = 0; rspan = 0; }
                else if (!any_ref) { nr = 0; nd = 0; rspan = 0; }
            }
        } else if (rpass) {
            bool tileable = hd.nc <= (uint32_t)kMaxCigarTile && hd.nc > 0 && (hd.so1 - hd.so) <= kMaxReadBytes &&
                            (hd.keep & 2u);
            uint32_t lq = 0;
            bool any_ref = false;
            if (tileable) {
                // l_qseq first (a deletion at the very end tests a quality that does not exist)
                for (uint32_t k = 0; k < hd.nc; ++k) {
                    const uint32_t c = k == 0 ? hd.cg0 : (k == 1 ? hd.cg1 : (k == 2 ? hd.cg2 : b.cigar[hd.c0 + k]));
                    if (op_consumes_query(c & 15u)) lq += c >> 4;
                }
                uint32_t qi = 0;
                int32_t r = hd.pos;
                bool prev_match = false;
                for (uint32_t k = 0; k < hd.nc && tileable; ++k) {
                    const uint32_t c = k == 0 ? hd.cg0 : (k == 1 ? hd.cg1 : (k == 2 ? hd.cg2 : b.cigar[hd.c0 + k]));
                    const uint32_t op = c & 15u, l = c >> 4;
                    if (op_is_match(op)) {
                        any_ref = true;
                        if (prev_match) {
                            if (nr == 1) run_len[0] += l; else if (nr == 2) run_len[1] += l; else run_len[2] += l;
                        } else if (nr == (uint32_t)kMaxRunsPerRead) tileable = false;
                        else {
                            if (nr == 0) { run_pos[0] = r; run_len[0] = l; run_q[0] = qi; }
                            else if (nr == 1) { run_pos[1] = r; run_len[1] = l; run_q[1] = qi; }
                            else { run_pos[2] = r; run_len[2] = l; run_q[2] = qi; }
                            ++nr;
                        }
                        qi += l; r += (int32_t)l; prev_match = true;
                    } else {
                        prev_match = false;
                        if (op == 2 || op == 3) {
                            any_ref = true;
                            // kept iff the NEXT query base passes the quality rule (0 if past the end).  The quality
                            // is only requested here; it is tested where the entries are deposited, after the barrier,
                            // so its latency does not hold up the chunk
                            if (nd == (uint32_t)kMaxDelsPerRead) tileable = false;
                            else {
                                const uint32_t q = qi < lq ? (uint32_t)b.qual[so0 + so_rel + qi] : 0u;
                                if (nd == 0) { del_pos[0] = r; del_len[0] = l; del_q[0] = q; }
                                else { del_pos[1] = r; del_len[1] = l; del_q[1] = q; }
                                ++nd;
                            }
                            r += (int32_t)l;
                        } else if (op == 1 || op == 4) qi += l;
                    }
                }
                rspan = (uint32_t)(r - hd.pos);
                if (tileable && any_ref && (hd.pos < 0 || (int64_t)hd.pos + rspan > tv.G)) {
                    atomicAdd(&tv.status[ST_RANGE_ERR], 1u);
                    nr = 0; nd = 0; rspan = 0; tileable = true; any_ref = false;
                }
#pragma unroll
                for (int k = 0; k < kMaxRunsPerRead; ++k) if ((uint32_t)k < nr && run_len[k] > 65535u) tileable = false;
            }
            if (!tileable) {
                // a record with no reference-consuming op at all is skipped everywhere
                if (!any_ref)
                    for (uint32_t k = 0; k < hd.nc; ++k) any_ref |= op_consumes_ref(b.cigar[hd.c0 + k] & 15u);
                defer = any_ref;
                nr = 0; nd = 0; rspan = 0;
            } else if (!any_ref) { nr = 0; nd = 0; rspan = 0; }
        }
        // ---- (3) warp-level compaction bookkeeping + chunk extents, then ONE barrier
        const uint32_t b1 = __ballot_sync(0xFFFFFFFFu, nr >= 1), b2 = __ballot_sync(0xFFFFFFFFu, nr >= 2),
                       b3 = __ballot_sync(0xFFFFFFFFu, nr >= 3);
        const uint32_t lt = (1u << lane) - 1u;
        const uint32_t wprefix = __popc(b1 & lt) + __popc(b2 & lt) + __popc(b3 & lt);
        {
            uint32_t sp = nr ? rspan : 0u;
            int32_t ce = nr ? (int32_t)(hd.pos + rspan) : 0;
            sp = __reduce_max_sync(0xFFFFFFFFu, sp);
            ce = __reduce_max_sync(0xFFFFFFFFu, ce);
            if (lane == 0) {
                wc[warp] = __popc(b1) + __popc(b2) + __popc(b3);
                if (sp) { atomicMax(&sc[2], sp); atomicMax(reinterpret_cast<int32_t*>(&sc[3]), ce); }
            }
        }
        __syncthreads();                                                   // barrier A
        uint32_t n_runs = 0, my_base = 0;
#pragma unroll
        for (int w = 0; w < kTileWarps; ++w) {
            const uint32_t c = wc[w];
            if (w < warp) my_base += c;
            n_runs += c;
        }
        const uint32_t maxspan = sc[2];
        const int32_t chunk_cmax = (int32_t)sc[3];
        if (n_runs > (uint32_t)kMaxRuns) {                                // run table full (indel-dense chunk): rare
            // reads whose runs do not fit are handed to the general kernel; the table ends where the first
            // such read would have started (prefix sums are monotone, so everything after it overflows too)
            if (tid == 0) s_misc[5] = 0;
            __syncthreads();
            const uint32_t my_end = my_base + wprefix + nr;
            if (nr && my_end > (uint32_t)kMaxRuns) { defer = true; nr = 0; nd = 0; rspan = 0; }
            uint32_t ok_end = nr ? my_end : 0u;
            ok_end = __reduce_max_sync(0xFFFFFFFFu, ok_end);
            if (lane == 0) atomicMax(&s_misc[5], ok_end);
            __syncthreads();
            n_runs = s_misc[5];
        }
        const bool active = rspan != 0;                                   // deposited by this kernel
        if (defer) s_dlist[atomicAdd(&s_misc[6], 1u)] = (uint16_t)tid;
        {
            // coverage difference array: one atomic per distinct start / end among the warp's reads
            const int32_t ks = active ? hd.pos : (int32_t)(0x80000000u + lane);
            const uint32_t ms = __match_any_sync(0xFFFFFFFFu, ks);
            if (active && lane == __ffs(ms) - 1) atomicAdd(&tv.covdiff[hd.pos], (int32_t)__popc(ms));
            const int32_t ke = active ? (int32_t)(hd.pos + rspan) : (int32_t)(0x80000000u + lane);
            const uint32_t me = __match_any_sync(0xFFFFFFFFu, ke);
            if (active && lane == __ffs(me) - 1) atomicAdd(&tv.covdiff[hd.pos + rspan], -(int32_t)__popc(me));
#pragma unroll
            for (int k = 0; k < kMaxDelsPerRead; ++k)
                if ((uint32_t)k < nd && (int)del_q[k] >= dp.min_bq)
                    for (uint32_t j = 0; j < del_len[k]; ++j) atomicAdd(&tv.dels[del_pos[k] + j], 1u);
        }
        if (n_runs) {
            if (nr) {
                const uint32_t off = so_rel - base_rel;                   // read's first byte relative to the base
                const uint32_t win = off / kWinStride;
                const uint32_t idx0 = my_base + wprefix;
#pragma unroll
                for (int k = 0; k < kMaxRunsPerRead; ++k) {
                    if ((uint32_t)k < nr) {
                        const uint32_t idx = idx0 + k;
                        s_pos[idx] = run_pos[k];
                        s_qo[idx] = off + run_q[k];
                        s_len[idx] = (uint16_t)run_len[k];
                        s_rd[idx] = (uint16_t)(run_pos[k] - hd.pos);
                        s_rix[idx] = (uint16_t)((win << 8) | (uint32_t)tid);
                        // (start column, key offset inside the read's window | length << 16): one 8-byte load per unit
                        s_pk[idx] = make_uint2((uint32_t)run_pos[k], ((off + run_q[k] - win * kWinStride) & 0xFFFFu) | (run_len[k] << 16));
                    }
                }
            }
            __syncthreads();                                               // barrier B: run table visible
        }

        // ---- (5) staged windows of runs (one per chunk unless reads are long)
        if (n_runs) {
            const uint32_t n_win = (uint32_t)(s_rix[n_runs - 1] >> 8) + 1u;
            uint32_t a0 = 0;
            for (uint32_t win = 0; win < n_win; ++win) {
                uint32_t a1 = n_runs;
                if (win + 1 < n_win) {            // first run of a later window
                    uint32_t lo = a0, hi = n_runs;
                    while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if ((uint32_t)(s_rix[m] >> 8) <= win) lo = m + 1; else hi = m; }
                    a1 = lo;
                }
                if (a1 == a0 && win > 0) continue;
                const uint32_t w_rel = win * kWinStride;                   // window start relative to the base
                const uint64_t qbeg = base_abs + w_rel;                    // 16-byte aligned
                // column range of these runs
                const int32_t cmin = s_pos[a0] - (int32_t)s_rd[a0];
                int32_t cmax = chunk_cmax;                                 // chunk-wide (an upper bound for any window)
                if (n_win > 1) {                                           // long reads: exact range of this window's runs
                    __syncthreads();
                    if (tid == 0) s_misc[4] = 0;
                    __syncthreads();
                    int32_t e = 0;
                    for (uint32_t r = a0 + tid; r < a1; r += kTileThreads) e = max(e, s_pos[r] + (int32_t)s_len[r]);
                    e = __reduce_max_sync(0xFFFFFFFFu, e);
                    if (lane == 0 && e) atomicMax(reinterpret_cast<int32_t*>(&s_misc[4]), e);
                    __syncthreads();
                    cmax = (int32_t)s_misc[4];
                }
                // Task table of one column window: per 32-column slab the candidate run range (binary search over the
                // sorted read start of each run) and the exclusive scan of the task counts.  One warp does it.  For the
                // first column window that happens while the warp's first payload loads are in flight.
                auto slab_setup = [&](int32_t wc0) {
                    const int nslab = min(kMaxSlabs, (cmax - wc0 + kSlabCols - 1) / kSlabCols);
                    uint32_t cnt = 0, first_run = 0;
                    if (lane < nslab) {
                        const int32_t s_lo = wc0 + lane * kSlabCols, s_hi = s_lo + kSlabCols;
                        uint32_t lo = a0, hi = a1;               // first run whose read starts at or after s_hi
                        while (lo < hi) {
                            const uint32_t m = (lo + hi) >> 1;
                            if (s_pos[m] - (int32_t)s_rd[m] < s_hi) lo = m + 1; else hi = m;
                        }
                        const uint32_t bnd = lo;
                        const int64_t thr = (int64_t)s_lo - (int64_t)maxspan;   // first run whose read starts after thr
                        lo = a0; hi = bnd;
                        while (lo < hi) {
                            const uint32_t m = (lo + hi) >> 1;
                            if ((int64_t)(s_pos[m] - (int32_t)s_rd[m]) <= thr) lo = m + 1; else hi = m;
                        }
                        first_run = lo;
                        cnt = bnd - lo;
                    }
                    const uint32_t groups = (cnt + kTaskRuns - 1) / kTaskRuns;
                    uint32_t incl = groups;
#pragma unroll
                    for (int d = 1; d < kMaxSlabs; d <<= 1) {
                        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                        if (lane >= d) incl += up;
                    }
                    if (lane < nslab) { s_slab_a[lane] = first_run; s_slab_n[lane] = cnt; s_slab_pre[lane] = incl - groups; }
                    if (lane == nslab - 1) s_slab_pre[nslab] = incl;
                    if (lane == 0) s_misc[2] = 0;
                };
                // ---- stage this window as KEYS: coalesced 16-byte loads, 16 bases per step, written in place once
                if (a1 > a0) {
                    const uint64_t qend_all = so0 + max_rel;
                    const uint64_t qend = qend_all < qbeg + kKeyCapBases ? qend_all : qbeg + kKeyCapBases;
                    const uint32_t n_grp = (uint32_t)(((qend - qbeg) + 15) >> 4);
                    const uint32_t chunk_ord = dp.ord_base + chunk0;
                    const uint4* gq = reinterpret_cast<const uint4*>(b.qual + qbeg);
                    const uint2* gs = reinterpret_cast<const uint2*>(b.seq4 + (qbeg >> 1));
                    // one group = 16 bases: 16 quality bytes + 8 sequence bytes -> 16 keys
                    auto stage_group = [&](uint32_t gg, const uint4& q, const uint2& sraw) {
                        // base nibbles in little-endian nibble order (base k at bits 4k)
                        const uint32_t s0 = bitsel(sraw.x >> 4, sraw.x << 4, 0x0F0F0F0Fu);
                        const uint32_t s1 = bitsel(sraw.y >> 4, sraw.y << 4, 0x0F0F0F0Fu);
                        // Qualities below 128 (every real file): no carries between bytes, so per word
                        //   bit 7 of (q ^ qprim) + 0x7F = "differs from the primary quality"
                        //   bit 7 of  q + (0x80 - minBQ) = "passes the base-quality threshold"
                        const uint32_t t0 = (q.x ^ qprim4) + 0x7F7F7F7Fu, t1 = (q.y ^ qprim4) + 0x7F7F7F7Fu;
                        const uint32_t t2 = (q.z ^ qprim4) + 0x7F7F7F7Fu, t3 = (q.w ^ qprim4) + 0x7F7F7F7Fu;
                        uint32_t cold = q.x | q.y | q.z | q.w;                       // a byte >= 128: exact path
                        if (GE_ALL) cold |= t0 | t1 | t2 | t3;
                        else cold |= (t0 & (q.x + ge_add4)) | (t1 & (q.y + ge_add4)) | (t2 & (q.z + ge_add4)) |
                                     (t3 & (q.w + ge_add4));                          // passing, not primary
                        // "differs" flags of the even / odd bases gathered and widened to bytes (PRMT sign mode)
                        uint32_t k0 = s0 & ~bitsel(prmt_sign(t0, t1, 0xECA8u), prmt_sign(t0, t1, 0xFDB9u), 0x0F0F0F0Fu);
                        uint32_t k1 = s1 & ~bitsel(prmt_sign(t2, t3, 0xECA8u), prmt_sign(t2, t3, 0xFDB9u), 0x0F0F0F0Fu);
                        if (cold & 0x80808080u) {
                            // rare: exact flags; a passing quality other than the primary one is deposited individually
                            k0 = 0; k1 = 0;
#pragma unroll 1
                            for (int w = 0; w < 4; ++w) {
                                const uint32_t qv = w == 0 ? q.x : (w == 1 ? q.y : (w == 2 ? q.z : q.w));
                                const uint32_t e80 = bytes_eq80(qv, qprim4);
                                const uint32_t sx = ((w < 2) ? s0 : s1) >> (16 * (w & 1));     // these 4 bases' nibbles
                                const uint32_t kw = sx & flags_to_nibbles(e80);
                                if (w < 2) k0 |= kw << (16 * (w & 1)); else k1 |= kw << (16 * (w & 1));
                                uint32_t m80 = (GE_ALL ? 0x80808080u : bytes_ge80(qv, ge_add4)) & ~e80;
                                while (m80) {
                                    const int bb = (__ffs(m80) - 1) >> 3;
                                    m80 &= ~(0x80u << (8 * bb));
                                    const uint32_t x_rel = w_rel + 16u * gg + 4u * (uint32_t)w + (uint32_t)bb;
                                    uint32_t lo = a0, hi = a1;             // last run with s_qo <= x_rel
                                    while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (s_qo[m] <= x_rel) lo = m + 1; else hi = m; }
                                    if (lo > a0) {
                                        const uint32_t r = lo - 1, d = x_rel - s_qo[r];
                                        if (d < (uint32_t)s_len[r])
                                            deposit_base(tv, dp, (int64_t)s_pos[r] + d, (sx >> (4 * bb)) & 15u,
                                                         (qv >> (8 * bb)) & 255u, chunk_ord + (s_rix[r] & 255u));
                                    }
                                }
                            }
                        }
                        asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(k_smem + 8u * gg), "r"(k0), "r"(k1) : "memory");
                    };
                    // two groups per iteration, loaded then reduced.  (A software-pipelined version with the next two
                    // groups in flight costs 12 more live registers under the 48-register cap and measured 4 % slower:
                    // with 5 CTAs per SM the load latency is covered by the other CTAs.  One group per iteration is 8 % slower.)
                    if (warp == 0 && cmin < cmax) slab_setup(cmin);
                    for (uint32_t g = tid; g < n_grp; g += 2 * kTileThreads) {
                        const bool hasB = g + kTileThreads < n_grp;
                        // streaming loads: the payload is read exactly once and should not displace the few hot lines
                        // (plane pointers, key LUT) in the 38 KB of L1 left beside the shared memory: -4 %
                        const uint4 qA = __ldcs(gq + g);
                        const uint2 sA = __ldcs(gs + g);
                        uint4 qB = make_uint4(0, 0, 0, 0);
                        uint2 sB = make_uint2(0, 0);
                        if (hasB) { qB = __ldcs(gq + g + kTileThreads); sB = __ldcs(gs + g + kTileThreads); }
                        stage_group(g, qA, sA);
                        if (hasB) stage_group(g + kTileThreads, qB, sB);
                    }
                }

                // ---- column windows of kTabCols (one for amplicon / deep shotgun chunks)
                for (int32_t wc0 = cmin; wc0 < cmax; wc0 += kTabCols) {
                    const int nslab = min(kMaxSlabs, (cmax - wc0 + kSlabCols - 1) / kSlabCols);
                    if (warp == 0 && wc0 != cmin) slab_setup(wc0);
                    __syncthreads();                                       // barrier C
                    const uint32_t n_tasks = s_slab_pre[nslab];

                    // ---- tasks: (slab, group of kTaskRuns runs); lane = 16 columns of one run per unit (2 lanes per run,
                    //      16 runs per pass): the per-task overhead is paid once per 4096 bases
                    const int w2 = lane & 1, sread = lane >> 1;
                    for (;;) {
                        uint32_t t = 0;
                        if (lane == 0) t = atomicAdd(&s_misc[2], 1u);
                        t = __shfl_sync(0xFFFFFFFFu, t, 0);
                        if (t >= n_tasks) break;
                        int k = 0;
                        while (k + 1 < nslab && s_slab_pre[k + 1] <= t) ++k;
                        const uint32_t ra = s_slab_a[k] + (t - s_slab_pre[k]) * (uint32_t)kTaskRuns;
                        const uint32_t rb = min(s_slab_a[k] + s_slab_n[k], ra + (uint32_t)kTaskRuns);
                        const int32_t col_lane = wc0 + k * kSlabCols + 16 * w2;
                        uint32_t acc4[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};    // [key word][allele]: 8 columns x 4-bit counters (<= 4)
                        const uint32_t pk_smem = sbase + Tile4Smem::pk_off;
                        // one unit = 16 columns of one run.  A run that misses this lane's columns (or a slot past the end
                        // of the group) becomes a unit of length 0, which the edge mask below empties: no other branch.
                        auto unit_load = [&](uint32_t r) {
                            uint32_t px = (uint32_t)col_lane, py = 0;
                            if (r < rb) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(px), "=r"(py) : "r"(pk_smem + 8u * r));
                            int32_t j = col_lane - (int32_t)px, len = (int32_t)(py >> 16);
                            const bool on = j > -16 && j < len;
                            j = on ? j : 0; len = on ? len : 0;
                            return pass_load4(k_smem, (int32_t)(py & 0xFFFFu) + j, j, len);
                        };
                        auto consume = [&](const PassUnit4& u) {
                            uint32_t k0 = u.k, k1 = u.k1;
                            if (u.j < 0 || u.j + 16 > u.len) {
                                // partial overlap at a run edge: keep keys with 0 <= j+b < len
                                const int lo = u.j < 0 ? -u.j : 0, hi = (u.len - u.j) < 16 ? (u.len - u.j) : 16;
                                const uint64_t vm = (hi >= 16 ? ~0ull : ((1ull << (4 * hi)) - 1ull)) & ~((1ull << (4 * lo)) - 1ull);
                                k0 &= (uint32_t)vm; k1 &= (uint32_t)(vm >> 32);
                            }
                            acc4[0][0] += k0 & 0x11111111u;        acc4[1][0] += k1 & 0x11111111u;
                            acc4[0][1] += (k0 >> 1) & 0x11111111u; acc4[1][1] += (k1 >> 1) & 0x11111111u;
                            acc4[0][2] += (k0 >> 2) & 0x11111111u; acc4[1][2] += (k1 >> 2) & 0x11111111u;
                            acc4[0][3] += (k0 >> 3) & 0x11111111u; acc4[1][3] += (k1 >> 3) & 0x11111111u;
                        };
                        // one unit per iteration: two in flight cost 8 more live registers under the 48-register cap and
                        // measured 2 % slower
#pragma unroll 1
                        for (uint32_t r = ra + sread; r < rb; r += 16) {
                            const PassUnit4 uA = unit_load(r);
                            consume(uA);
                        }
                        // widen to 8-bit fields: v[word][parity][allele]; parity 0 = even columns of the word
                        uint32_t v16[16];
#pragma unroll
                        for (int wd = 0; wd < 2; ++wd)
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                v16[(wd * 2 + 0) * 4 + c] = acc4[wd][c] & 0x0F0F0F0Fu;
                                v16[(wd * 2 + 1) * 4 + c] = (acc4[wd][c] >> 4) & 0x0F0F0F0Fu;
                            }
                        // ---- reduce-scatter over the 16 runs of a pass (lane bits 1..4): 16 -> 8 -> 4 -> 2 -> 1 registers,
                        //      fields stay <= 64
                        {
                            const bool b4 = lane & 16, b3x = lane & 8, b2x = lane & 4, b1x = lane & 2;
                            uint32_t v8[8], v4[4], v2[2];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const uint32_t keepv = b4 ? v16[8 + i] : v16[i], send = b4 ? v16[i] : v16[8 + i];
                                v8[i] = keepv + __shfl_xor_sync(0xFFFFFFFFu, send, 16);
                            }
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint32_t keepv = b3x ? v8[4 + i] : v8[i], send = b3x ? v8[i] : v8[4 + i];
                                v4[i] = keepv + __shfl_xor_sync(0xFFFFFFFFu, send, 8);
                            }
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                const uint32_t keepv = b2x ? v4[2 + i] : v4[i], send = b2x ? v4[i] : v4[2 + i];
                                v2[i] = keepv + __shfl_xor_sync(0xFFFFFFFFu, send, 4);
                            }
                            const uint32_t keepv = b1x ? v2[1] : v2[0], send = b1x ? v2[0] : v2[1];
                            const uint32_t v = keepv + __shfl_xor_sync(0xFFFFFFFFu, send, 2);
                            // this lane now owns register index  b4*8 + b3*4 + b2*2 + b1  = (word*2 + parity)*4 + allele
                            const int word = b4 ? 1 : 0, parity = b3x ? 1 : 0;
                            const int code = (b2x ? 2 : 0) + (b1x ? 1 : 0);
                            const int colrel = (col_lane - wc0) + 8 * word + parity;
                            if (v) {
#pragma unroll
                                for (int bb = 0; bb < 4; ++bb) {
                                    const uint32_t f = (v >> (8 * bb)) & 255u;
                                    if (f) atomicAdd(&s_tab[(colrel + 2 * bb) * 4 + code], f);
                                }
                            }
                        }
                    }
                    __syncthreads();

                    // ---- flush: one global RED per non-zero (column, allele); collect first-seen work
                    const uint32_t chunk_ord0 = dp.ord_base + chunk0;
                    uint32_t* plane = tv.planes[tp.prim_plane];
                    uint32_t* first0 = tv.first[0];
                    const uint32_t ord_lo = chunk_ord0 + (s_rix[a0] & 255u);
                    const int ncols = min(kTabCols, cmax - wc0);
                    {
                        // 4 table cells per thread; the 4 first-seen loads are issued together (one memory latency, not 4)
                        static_assert(kTabCols * 4 == 4 * kTileThreads, "flush: 4 cells per thread");
                        uint32_t fv[4], ff[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int e = tid + k * kTileThreads;
                            fv[k] = e < ncols * 4 ? s_tab[e] : 0u;
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int e = tid + k * kTileThreads;
                            ff[k] = 0;
                            if (fv[k]) {
                                s_tab[e] = 0;
                                const int64_t cell = (int64_t)wc0 * 4 + e;
                                atomicAdd(&plane[cell], fv[k]);
                                ff[k] = first0[cell];
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (fv[k] && ff[k] > ord_lo) s_items[atomicAdd(&s_misc[3], 1u)] = (uint16_t)(tid + k * kTileThreads);
                    }
                    __syncthreads();                                       // barrier E: table flushed, items known
                    const uint32_t n_items = s_misc[3];
                    if (n_items) {
                        // exact first-seen ordinal for new (column, allele) pairs: scan the runs in read order
                        for (uint32_t it = warp; it < n_items; it += kTileWarps) {
                            const uint32_t e = s_items[it];
                            const int32_t col = wc0 + (int32_t)(e >> 2);
                            const uint32_t want = 1u << (e & 3u);
                            for (uint32_t r0 = a0; r0 < a1; r0 += 32) {
                                const uint32_t r = r0 + lane;
                                bool hit = false;
                                if (r < a1) {
                                    const int32_t j = col - s_pos[r];
                                    if (j >= 0 && j < (int32_t)s_len[r]) {
                                        const uint32_t ka = (s_qo[r] - w_rel) + (uint32_t)j;
                                        const uint32_t key = (lds32(k_smem + ((ka >> 3) << 2)) >> ((ka & 7u) * 4u)) & 15u;
                                        hit = key == want;
                                    }
                                }
                                const uint32_t hb = __ballot_sync(0xFFFFFFFFu, hit);
                                if (hb) {
                                    if (lane == 0)
                                        atomicMin(&first0[(int64_t)col * 4 + (e & 3u)],
                                                  chunk_ord0 + (s_rix[r0 + (__ffs(hb) - 1)] & 255u));
                                    break;
                                }
                            }
                        }
                        __syncthreads();
                        if (tid == 0) s_misc[3] = 0;
                        __syncthreads();
                    }
                }
                if (win + 1 < n_win) __syncthreads();                      // staging buffer is reused by the next window
                a0 = a1;
            }
        }
        // ---- reads the tiled path could not take (many runs, long, exotic base codes): general path, one warp each
        __syncthreads();
        const uint32_t n_def = s_misc[6];
        for (uint32_t d = warp; d < n_def; d += kTileWarps) deposit_read_warp(b, tv, dp, chunk0 + s_dlist[d], lane);
    }
}

}  // namespace lvc
