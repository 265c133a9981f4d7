// Tiled deposit kernel (the fast path): CIGAR expansion + count accumulation without per-base atomics.
//
// Replaces the per-(column, read) loop of live_variant_caller.py:69-70,89-103 and the htslib CIGAR walk
// behind it for every read with at most kMaxRunsPerRead match runs (all Illumina reads, with or
// without an indel).  Reads with more runs / very long reads / base codes beyond A,C,G,T are
// handled by the same CTA on the general path (one warp per read) once its tiles are flushed.
//
// Work decomposition (B200: 148 SMs, 3 persistent CTAs per SM at ~73 KB shared memory each)
//   CTA    = persistent; pulls chunks of kTileReads consecutive (coordinate-sorted) reads from a global
//            counter.  The NEXT chunk's read headers are prefetched into registers while the current
//            chunk is processed, so the dependent global loads (offsets -> CIGAR) never stall a CTA.
//   stage  = after classification, ONLY the byte range of the reads that will be deposited is copied
//            into shared memory with two TMA bulk copies (cp.async.bulk + mbarrier): chunks whose
//            reads were all dropped (max_depth / flags / mapq) touch no payload bytes at all, and
//            every staged byte crosses HBM/L2 exactly once.
//   run    = a maximal match run of a read (start column, length, first quality byte); a read with
//            one indel contributes two runs.  Deletion entries are counted straight into dels[].
//   task   = (32-column slab, group of 32 runs), fetched by warps from a shared-memory queue.
//   lane   = 8 columns of one run per unit (4 lanes per run, 8 runs per pass, 2 units in flight).  A
//            lane keeps its 8 columns for the whole task, so the A/C/G/T counts of those columns live
//            in REGISTERS as SWAR fields (4 columns x 8 bit per register): no atomics and no
//            cross-lane traffic in the inner loop.
//   flush  = butterfly reduce-scatter across the 8 runs of a pass, 4 shared-memory atomics per lane
//            into the CTA's column table, then ONE global RED per non-zero (column, allele) per chunk.
// Only bases whose quality equals the batch's primary quality `qprim` (the most frequent passing
// value; the only passing one for binned Illumina data at minBQ 30) take the register path; any other
// passing quality is deposited individually (exact, slower) -- correctness never depends on qprim.
#pragma once
#include "lvc_common.cuh"
#include "deposit_general.cuh"

namespace lvc {

constexpr int kTileThreads = 256;
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kTileReads = 256;                 // reads per chunk (one header per thread)
constexpr int kMaxRuns = 384;                   // match runs per chunk (reads beyond are deferred)
constexpr int kMaxRunsPerRead = 3;
constexpr int kMaxDelsPerRead = 2;
constexpr int kMaxCigarTile = 8;
constexpr uint32_t kQCap = 41u * 1024u;         // staged quality bytes per window
constexpr uint32_t kMaxReadBytes = 2048;        // longer reads take the general kernel
constexpr uint32_t kWinStride = kQCap - kMaxReadBytes;   // 39,936 >= 256 x 152 + 16: one window per chunk of 150 bp reads
constexpr int kTabCols = 256;                   // columns per shared count table window
constexpr int kSlabCols = 32;
constexpr int kMaxSlabs = kTabCols / kSlabCols; // 8
constexpr uint32_t kSlack = 32;                 // bytes of slack before/after the staged arrays

// dynamic shared memory layout
struct TileSmem {
    static constexpr uint32_t qual_off = 0;
    static constexpr uint32_t qual_bytes = kSlack + kQCap + 16 + kSlack;
    static constexpr uint32_t seq_off = qual_off + qual_bytes;
    static constexpr uint32_t seq_bytes = kSlack + kQCap / 2 + 32 + kSlack;
    static constexpr uint32_t tab_off = seq_off + seq_bytes;                   // u32 [kTabCols*4]
    static constexpr uint32_t tab_bytes = kTabCols * 4 * 4;
    // per match run (compacted, read order); byte offsets relative to the chunk's staging base
    static constexpr uint32_t pos_off = tab_off + tab_bytes;                   // i32 [kMaxRuns] run start column
    static constexpr uint32_t qo_off = pos_off + kMaxRuns * 4;                 // u32 [kMaxRuns] run's first quality byte
    static constexpr uint32_t len_off = qo_off + kMaxRuns * 4;                 // u16 [kMaxRuns] run length
    static constexpr uint32_t rd_off = len_off + kMaxRuns * 2;                 // u16 [kMaxRuns] run start - read start
    static constexpr uint32_t rix_off = rd_off + kMaxRuns * 2;                 // u16 [kMaxRuns] window<<8 | read index in chunk
    static constexpr uint32_t items_off = rix_off + kMaxRuns * 2;              // u16 [kTabCols*4]
    static constexpr uint32_t dlist_off = items_off + kTabCols * 4 * 2;        // u16 [kTileReads] reads left to the general path
    static constexpr uint32_t slab_a_off = dlist_off + kTileReads * 2;         // u32 [kMaxSlabs]
    static constexpr uint32_t slab_pre_off = slab_a_off + kMaxSlabs * 4;       // u32 [kMaxSlabs+1]
    static constexpr uint32_t slab_n_off = slab_pre_off + (kMaxSlabs + 1) * 4; // u32 [kMaxSlabs]
    static constexpr uint32_t misc_off = (slab_n_off + kMaxSlabs * 4 + 15) & ~15u;   // mbarrier + scalars
    static constexpr uint32_t total = misc_off + 128;
};
constexpr size_t kTileSmemBytes = TileSmem::total;
static_assert(kTileSmemBytes * 3 + 3 * 1024 <= 227 * 1024, "three CTAs per SM must fit");

struct TileParams {
    uint32_t grid;
    uint32_t n_chunks;
    uint32_t qprim;        // primary quality (255 = none)
    uint32_t prim_plane;   // plane id of (group 0, qprim)
    uint32_t qc_pcode;     // quality-code batches: the code of qprim (0..3)
    uint32_t qc_cold;      // quality-code batches: bit c = code c passes the base-quality threshold and is not qc_pcode
};

inline TileParams make_tile_params(uint32_t n_reads, int sm_count) {
    TileParams tp;
    tp.n_chunks = (n_reads + kTileReads - 1) / kTileReads;
    tp.grid = tp.n_chunks;          // one CTA per chunk: the hardware block scheduler balances dead and live chunks
    (void)sm_count;
    tp.qprim = 255;
    tp.prim_plane = 0;
    tp.qc_pcode = 0;
    tp.qc_cold = 0;
    return tp;
}

// ---- PTX helpers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// ---- SWAR helpers (4 bytes per register) -----------------------------------------------------------
// 0x80 in every byte of x that equals the corresponding byte of pattern p4
__device__ __forceinline__ uint32_t bytes_eq80(uint32_t x, uint32_t p4) {
    const uint32_t y = x ^ p4;
    const uint32_t t = (y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | y) & 0x80808080u;
}
// 0x80 in every byte of x that is >= m (1 <= m <= 128); add4 = (0x80 - m) replicated
__device__ __forceinline__ uint32_t bytes_ge80(uint32_t x, uint32_t add4) {
    return (((x & 0x7F7F7F7Fu) + add4) | x) & 0x80808080u;
}
// spread the 4 nibbles held in two bytes of V (selected by PRMT selector `sel`) to one nibble per byte
__device__ __forceinline__ uint32_t spread_nibbles(uint32_t V, uint32_t sel) {
    const uint32_t y = __byte_perm(V, 0, sel);
    return ((y >> 4) & 0x000F000Fu) | (y & 0x0F000F00u);
}

// deposit the bytes flagged in `m80` (0x80 per byte) one by one: quality != qprim but passing
__device__ __noinline__ void tile_slow_bytes(const TableView& tv, const DepositParams& dp, uint32_t m80, uint32_t qw,
                                             uint32_t sw, int64_t col0, uint32_t ord) {
    while (m80) {
        const int b = (__ffs(m80) - 1) >> 3;
        m80 &= ~(0x80u << (8 * b));
        deposit_base(tv, dp, col0 + b, (sw >> (8 * b)) & 15u, (qw >> (8 * b)) & 255u, ord);
    }
}

// one (run, 8 columns) unit of a pass: aligned shared loads -> 8 qualities + 8 spread base nibbles
struct PassUnit {
    uint32_t q0, q1, sw0, sw1;
    int32_t j, len;
};

__device__ __forceinline__ PassUnit pass_load(uint32_t q_smem, uint32_t s_smem, int32_t qa, int32_t sn_delta,
                                              int32_t j, int32_t len) {
    PassUnit u;
    u.j = j; u.len = len;
    // 8 qualities, byte aligned from three aligned shared words
    const uint32_t a = q_smem + (uint32_t)(qa & ~3);
    const uint32_t x0 = lds32(a), x1 = lds32(a + 4), x2 = lds32(a + 8);
    const uint32_t sh = (uint32_t)(qa & 3) * 8u;
    u.q0 = __funnelshift_r(x0, x1, sh);
    u.q1 = __funnelshift_r(x1, x2, sh);
    // 8 base nibbles (big-endian within bytes) from two aligned shared words
    const int32_t ni = qa + sn_delta;
    const uint32_t sa = s_smem + (uint32_t)((ni >> 1) & ~3);
    const uint32_t y0 = __byte_perm(lds32(sa), 0, 0x0123);
    const uint32_t y1 = __byte_perm(lds32(sa + 4), 0, 0x0123);
    const uint32_t V = __funnelshift_l(y1, y0, (uint32_t)(ni & 7) * 4u);
    u.sw0 = spread_nibbles(V, 0x2233);
    u.sw1 = spread_nibbles(V, 0x0011);
    return u;
}

// per-read header registers (prefetched one chunk ahead; scalar members only, so they stay in registers)
struct ReadHdr {
    int32_t pos;
    uint32_t flag, mapq, keep;
    uint32_t c0, nc;
    uint32_t so, so1;          // low words of seq_off[i], seq_off[i+1] (differences are < 2^32)
    uint32_t cg0, cg1, cg2;    // first CIGAR ops
};

template <int READS = kTileReads>
__device__ __forceinline__ void hdr_load1(const BatchView& b, uint32_t chunk, int tid, ReadHdr& h, uint64_t& so0) {
    const uint32_t chunk0 = chunk * READS;
    const uint32_t n = min((uint32_t)READS, b.n_reads - chunk0);
    const bool mine = (uint32_t)tid < n;
    const uint32_t i = chunk0 + (mine ? tid : 0);
    h.pos = b.pos[i];
    h.flag = b.flag[i];
    h.mapq = b.mapq[i];
    h.keep = mine ? (uint32_t)b.keep[i] : 0u;
    h.c0 = b.cigar_off[i];
    h.nc = b.cigar_off[i + 1] - h.c0;
    const uint32_t* so32 = reinterpret_cast<const uint32_t*>(b.seq_off);
    h.so = so32[2 * (size_t)i];
    h.so1 = so32[2 * (size_t)i + 2];
    so0 = b.seq_off[chunk0];
}
// second stage: the CIGAR ops, only for reads that survive the read-level filter (dropped reads never
// need them, so chunks of dropped reads cost a single round of loads)
__device__ __forceinline__ void hdr_load2(const BatchView& b, ReadHdr& h, int min_mq) {
    const bool need = read_passes_filter(h.flag, h.mapq, h.keep, min_mq);
    h.cg0 = (need && h.nc > 0) ? b.cigar[h.c0] : 0u;
    h.cg1 = (need && h.nc > 1) ? b.cigar[h.c0 + 1] : 0u;
    h.cg2 = (need && h.nc > 2) ? b.cigar[h.c0 + 2] : 0u;
}

template <bool GE_ALL>
__global__ void __launch_bounds__(kTileThreads, 3)
k_deposit_tile(BatchView b, TableView tv, DepositParams dp, TileParams tp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem + TileSmem::tab_off);
    int32_t* s_pos = reinterpret_cast<int32_t*>(smem + TileSmem::pos_off);
    uint32_t* s_qo = reinterpret_cast<uint32_t*>(smem + TileSmem::qo_off);
    uint16_t* s_len = reinterpret_cast<uint16_t*>(smem + TileSmem::len_off);
    uint16_t* s_rd = reinterpret_cast<uint16_t*>(smem + TileSmem::rd_off);
    uint16_t* s_rix = reinterpret_cast<uint16_t*>(smem + TileSmem::rix_off);
    uint16_t* s_items = reinterpret_cast<uint16_t*>(smem + TileSmem::items_off);
    uint16_t* s_dlist = reinterpret_cast<uint16_t*>(smem + TileSmem::dlist_off);
    uint32_t* s_slab_a = reinterpret_cast<uint32_t*>(smem + TileSmem::slab_a_off);
    uint32_t* s_slab_pre = reinterpret_cast<uint32_t*>(smem + TileSmem::slab_pre_off);
    uint32_t* s_slab_n = reinterpret_cast<uint32_t*>(smem + TileSmem::slab_n_off);
    uint32_t* s_misc = reinterpret_cast<uint32_t*>(smem + TileSmem::misc_off);
    // s_misc: [0,1] mbarrier  [2] task counter  [3] n_items  [4] window max end column (long reads only)  [6] deferred reads
    //         [5] run table end (overflow only)
    //         [8..11] min read byte, max read end byte, max reference span, max end column   [32..39] runs per warp
    const uint32_t bar = sbase + TileSmem::misc_off;
    const uint32_t q_smem = sbase + TileSmem::qual_off + kSlack;     // staged qualities start here
    const uint32_t s_smem = sbase + TileSmem::seq_off + kSlack;      // staged 4-bit bases start here

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
        mbar_init(bar, 1);
        s_misc[3] = 0; s_misc[6] = 0;
        sc[0] = 0xFFFFFFFFu; sc[1] = 0; sc[2] = 0; sc[3] = 0;
    }
    for (int k = tid; k < kTabCols * 4; k += kTileThreads) s_tab[k] = 0;
    hdr_load2(b, hd, dp.min_mq);       // CIGAR ops, only for reads that pass the read-level filter
    __syncthreads();
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
    // staging base: 16-byte aligned start of the first such read; window 0 is staged right away
    const uint64_t base_abs = (so0 + min_rel) & ~15ull;
    const uint32_t base_rel = (uint32_t)(base_abs - so0);            // may wrap below zero: used mod 2^32
    auto stage_window = [&](uint32_t win) {
        const uint64_t qbeg = base_abs + (uint64_t)win * kWinStride;      // 16-byte aligned
        const uint64_t qend_all = so0 + max_rel;
        const uint64_t qend = qend_all < qbeg + kQCap ? qend_all : qbeg + kQCap;
        const uint64_t sbeg16 = (qbeg >> 1) & ~15ull;
        const uint32_t qbytes = (uint32_t)(((qend - qbeg) + 15) & ~15ull);
        const uint32_t sbytes = (uint32_t)((((qend + 1) >> 1) - sbeg16 + 15) & ~15ull);
        mbar_expect_tx(bar, qbytes + sbytes);
        if (qbytes) tma_bulk_g2s(q_smem, b.qual + qbeg, qbytes, bar);
        if (sbytes) tma_bulk_g2s(s_smem, b.seq4 + sbeg16, sbytes, bar);
    };
    if (tid == 0) stage_window(0);
    uint32_t phase = 0;
    {
        const uint32_t chunk0 = cur * kTileReads;
        // ---- (2) classify this thread's read: filter, match runs, deletion entries
        int32_t run_pos[kMaxRunsPerRead] = {0, 0, 0};
        uint32_t run_len[kMaxRunsPerRead] = {0, 0, 0}, run_q[kMaxRunsPerRead] = {0, 0, 0};
        int32_t del_pos[kMaxDelsPerRead] = {0, 0};
        uint32_t del_len[kMaxDelsPerRead] = {0, 0};
        uint32_t nr = 0, nd = 0, rspan = 0;
        bool defer = false;
        const uint32_t i = chunk0 + tid;
        if (read_passes_filter(hd.flag, hd.mapq, hd.keep, dp.min_mq)) {
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
                            // kept iff the NEXT query base passes the quality rule (0 if past the end)
                            const uint32_t q = qi < lq ? (uint32_t)b.qual[b.seq_off[i] + qi] : 0u;
                            if ((int)q >= dp.min_bq) {
                                if (nd == (uint32_t)kMaxDelsPerRead) tileable = false;
                                else {
                                    if (nd == 0) { del_pos[0] = r; del_len[0] = l; } else { del_pos[1] = r; del_len[1] = l; }
                                    ++nd;
                                }
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
                if ((uint32_t)k < nd)
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
                const int32_t sn_delta = (int32_t)(qbeg - 2 * ((qbeg >> 1) & ~15ull));      // 0 or 16
                if (win > 0 && tid == 0) stage_window(win);
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
                bool waited = false;

                // ---- column windows of kTabCols (one for amplicon / deep shotgun chunks)
                for (int32_t wc0 = cmin; wc0 < cmax; wc0 += kTabCols) {
                    const int nslab = min(kMaxSlabs, (cmax - wc0 + kSlabCols - 1) / kSlabCols);
                    // per-slab candidate run range by binary search over the (sorted) read start of each run;
                    // warp 0 does the searches and the exclusive scan of the task counts
                    if (warp == 0) {
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
                        const uint32_t groups = (cnt + 31) >> 5;
                        uint32_t incl = groups;
#pragma unroll
                        for (int d = 1; d < kMaxSlabs; d <<= 1) {
                            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                            if (lane >= d) incl += up;
                        }
                        if (lane < nslab) { s_slab_a[lane] = first_run; s_slab_n[lane] = cnt; s_slab_pre[lane] = incl - groups; }
                        if (lane == nslab - 1) s_slab_pre[nslab] = incl;
                        if (lane == 0) s_misc[2] = 0;
                    }
                    __syncthreads();                                       // barrier C
                    const uint32_t n_tasks = s_slab_pre[nslab];
                    if (!waited) { mbar_wait(bar, phase); phase ^= 1; waited = true; }   // staged bytes have landed

                    // ---- tasks: (slab, group of 32 runs); lane = 8 columns of one run per unit
                    const int w4 = lane & 3, sread = lane >> 2;
                    for (;;) {
                        uint32_t t = 0;
                        if (lane == 0) t = atomicAdd(&s_misc[2], 1u);
                        t = __shfl_sync(0xFFFFFFFFu, t, 0);
                        if (t >= n_tasks) break;
                        int k = 0;
                        while (k + 1 < nslab && s_slab_pre[k + 1] <= t) ++k;
                        const uint32_t ra = s_slab_a[k] + ((t - s_slab_pre[k]) << 5);
                        const uint32_t rb = min(s_slab_a[k] + s_slab_n[k], ra + 32u);
                        const int32_t col_lane = wc0 + k * kSlabCols + 8 * w4;
                        uint32_t acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
                        auto consume = [&](const PassUnit& u, uint32_t r) {
                            uint32_t p0 = bytes_eq80(u.q0, qprim4), p1 = bytes_eq80(u.q1, qprim4);
                            uint32_t g0 = GE_ALL ? 0x80808080u : bytes_ge80(u.q0, ge_add4);
                            uint32_t g1 = GE_ALL ? 0x80808080u : bytes_ge80(u.q1, ge_add4);
                            if (u.j < 0 || u.j + 8 > u.len) {
                                // partial overlap at a run edge: keep bytes with 0 <= j+b < len
                                const int lo = u.j < 0 ? -u.j : 0, hi = (u.len - u.j) < 8 ? (u.len - u.j) : 8;
                                const uint64_t vm = ((hi >= 8 ? ~0ull : ((1ull << (8 * hi)) - 1ull)) & ~((1ull << (8 * lo)) - 1ull));
                                const uint32_t v0 = (uint32_t)vm, v1 = (uint32_t)(vm >> 32);
                                p0 &= v0; p1 &= v1; g0 &= v0; g1 &= v1;
                            }
                            const uint32_t o0 = g0 & ~p0, o1 = g1 & ~p1;
                            const uint32_t m0 = p0 >> 7, m1 = p1 >> 7;
                            acc[0][0] += u.sw0 & m0;        acc[1][0] += u.sw1 & m1;
                            acc[0][1] += (u.sw0 >> 1) & m0; acc[1][1] += (u.sw1 >> 1) & m1;
                            acc[0][2] += (u.sw0 >> 2) & m0; acc[1][2] += (u.sw1 >> 2) & m1;
                            acc[0][3] += (u.sw0 >> 3) & m0; acc[1][3] += (u.sw1 >> 3) & m1;
                            if (o0 | o1) {
                                const uint32_t ord = dp.ord_base + chunk0 + (s_rix[r] & 255u);
                                if (o0) tile_slow_bytes(tv, dp, o0, u.q0, u.sw0, (int64_t)col_lane, ord);
                                if (o1) tile_slow_bytes(tv, dp, o1, u.q1, u.sw1, (int64_t)col_lane + 4, ord);
                            }
                        };
                        // two independent (run, 8 columns) units per iteration: twice the loads in flight
#pragma unroll 1
                        for (uint32_t r = ra + sread; r < rb; r += 16) {
                            const uint32_t r2 = r + 8;
                            const int32_t jA = col_lane - s_pos[r], lenA = (int32_t)s_len[r];
                            const bool onA = jA > -8 && jA < lenA;
                            int32_t jB = 0, lenB = 0;
                            bool onB = false;
                            if (r2 < rb) { jB = col_lane - s_pos[r2]; lenB = (int32_t)s_len[r2]; onB = jB > -8 && jB < lenB; }
                            PassUnit uA, uB;
                            if (onA) uA = pass_load(q_smem, s_smem, (int32_t)(s_qo[r] - w_rel) + jA, sn_delta, jA, lenA);
                            if (onB) uB = pass_load(q_smem, s_smem, (int32_t)(s_qo[r2] - w_rel) + jB, sn_delta, jB, lenB);
                            if (onA) consume(uA, r);
                            if (onB) consume(uB, r2);
                        }
                        // ---- reduce-scatter over the 8 runs of a pass (lane bits 2..4), fields stay <= 32
                        {
                            const bool b4 = lane & 16, b3x = lane & 8, b2x = lane & 4;
                            uint32_t m4[4];
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const uint32_t keepv = b4 ? acc[1][c] : acc[0][c];
                                const uint32_t send = b4 ? acc[0][c] : acc[1][c];
                                m4[c] = keepv + __shfl_xor_sync(0xFFFFFFFFu, send, 16);
                            }
                            uint32_t m2[2];
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                const uint32_t keepv = b3x ? m4[2 + c] : m4[c];
                                const uint32_t send = b3x ? m4[c] : m4[2 + c];
                                m2[c] = keepv + __shfl_xor_sync(0xFFFFFFFFu, send, 8);
                            }
                            const uint32_t keepv = b2x ? m2[1] : m2[0];
                            const uint32_t send = b2x ? m2[0] : m2[1];
                            const uint32_t v = keepv + __shfl_xor_sync(0xFFFFFFFFu, send, 4);
                            // this lane now owns: word h = b4, allele = 2*b3 + b2, columns col_lane + 4h + {0..3}
                            const int code = (b3x ? 2 : 0) + (b2x ? 1 : 0);
                            const int colrel = (col_lane - wc0) + (b4 ? 4 : 0);
                            if (v) {
#pragma unroll
                                for (int bb = 0; bb < 4; ++bb) {
                                    const uint32_t f = (v >> (8 * bb)) & 255u;
                                    if (f) atomicAdd(&s_tab[(colrel + bb) * 4 + code], f);
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
                    for (int e = tid; e < ncols * 4; e += kTileThreads) {
                        const uint32_t v = s_tab[e];
                        if (v) {
                            s_tab[e] = 0;
                            const int64_t cell = (int64_t)wc0 * 4 + e;
                            atomicAdd(&plane[cell], v);
                            if (first0[cell] > ord_lo) s_items[atomicAdd(&s_misc[3], 1u)] = (uint16_t)e;
                        }
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
                                        const uint32_t qa = (s_qo[r] - w_rel) + (uint32_t)j;
                                        const uint32_t q = (lds32(q_smem + (qa & ~3u)) >> ((qa & 3u) * 8u)) & 255u;
                                        const uint32_t ni = qa + (uint32_t)sn_delta;
                                        const uint32_t by = (lds32(s_smem + ((ni >> 1) & ~3u)) >> (((ni >> 1) & 3u) * 8u)) & 255u;
                                        const uint32_t nib = (ni & 1u) ? (by & 15u) : (by >> 4);
                                        hit = (q == tp.qprim) && (nib == want);
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
                if (!waited) { mbar_wait(bar, phase); phase ^= 1; }       // never leave a bulk copy in flight
                if (win + 1 < n_win) __syncthreads();                      // staging buffer is reused by the next window
                a0 = a1;
            }
        } else {
            mbar_wait(bar, phase);                                         // window 0 was staged speculatively
        }
        // ---- reads the tiled path could not take (many runs, long, exotic base codes): general path, one warp each
        __syncthreads();
        const uint32_t n_def = s_misc[6];
        for (uint32_t d = warp; d < n_def; d += kTileWarps) deposit_read_warp(b, tv, dp, chunk0 + s_dlist[d], lane);
    }
}

}  // namespace lvc
