// Tiled deposit kernel (the fast path): CIGAR expansion + count accumulation without per-base atomics.
//
// Replaces the per-(column, read) loop of live_variant_caller.py:69-70,89-103 for "simple" reads
// (one contiguous match run, optional clips: every Illumina read without an indel).  Everything
// else (indels, ref-skips, long reads, exotic base codes) is appended to a deferred list that the
// general kernel (deposit_general.cuh) processes right after.
//
// Work decomposition (B200: 148 SMs, 3 CTAs/SM at ~72 KB smem each)
//   CTA    = chunk of kTileReads consecutive (coordinate-sorted) reads.  Their packed qualities and
//            4-bit bases are CONTIGUOUS in the batch buffers, so one elected thread stages them into
//            shared memory with two TMA bulk copies (cp.async.bulk + mbarrier); every input byte
//            crosses HBM/L2 exactly once and all later accesses are shared-memory loads.
//   task   = (32-column slab, group of 32 reads), fetched by warps from a shared-memory queue.
//   lane   = 8 columns of one read per pass (4 lanes per read, 8 reads per pass).  A lane keeps its 8
//            columns for the whole task, so the A/C/G/T counts of those columns live in REGISTERS as
//            SWAR fields (4 columns x 8 bit per register): no atomics, no cross-lane traffic in the
//            inner loop.  Per pass a lane does 3+2 shared loads, two funnel shifts to byte/nibble
//            align, and ~25 integer ops per 4 bases.
//   flush  = butterfly reduce-scatter across the 8 reads of a pass (28 instr), 4 shared-memory
//            atomics per lane into the CTA's column table, then ONE global RED per non-zero
//            (column, allele) per chunk.
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
constexpr uint32_t kQCap = 39u * 1024u;         // staged quality bytes per sub-chunk (256 x 152)
constexpr int kTabCols = 256;                   // columns per shared count table window
constexpr int kSlabCols = 32;
constexpr int kMaxSlabs = kTabCols / kSlabCols; // 8
constexpr int kMaxCigarSimple = 8;
constexpr uint32_t kSlack = 32;                 // bytes of slack before/after the staged arrays

// dynamic shared memory layout
struct TileSmem {
    static constexpr uint32_t qual_off = 0;                                    // [kSlack + kQCap + 16 + kSlack]
    static constexpr uint32_t qual_bytes = kSlack + kQCap + 16 + kSlack;
    static constexpr uint32_t seq_off = qual_off + qual_bytes;                 // [kSlack + kQCap/2 + 32 + kSlack]
    static constexpr uint32_t seq_bytes = kSlack + kQCap / 2 + 32 + kSlack;
    static constexpr uint32_t tab_off = seq_off + seq_bytes;                   // u32 [kTabCols*4]
    static constexpr uint32_t tab_bytes = kTabCols * 4 * 4;
    // per ACTIVE read (compacted, coordinate order), offsets relative to the chunk's first byte
    static constexpr uint32_t pos_off = tab_off + tab_bytes;                   // i32 [kTileReads] run start column
    static constexpr uint32_t len_off = pos_off + kTileReads * 4;              // u32 [kTileReads] run length
    static constexpr uint32_t qo_off = len_off + kTileReads * 4;               // u32 [kTileReads] run's first quality byte
    static constexpr uint32_t beg_off = qo_off + kTileReads * 4;               // u32 [kTileReads] read's first byte
    static constexpr uint32_t end_off = beg_off + kTileReads * 4;              // u32 [kTileReads] read's end byte
    static constexpr uint32_t rix_off = end_off + kTileReads * 4;              // u16 [kTileReads] index within the chunk
    static constexpr uint32_t items_off = rix_off + kTileReads * 2;            // u16 [kTabCols*4]
    static constexpr uint32_t slab_a_off = items_off + kTabCols * 4 * 2;       // u32 [kMaxSlabs]
    static constexpr uint32_t slab_pre_off = slab_a_off + kMaxSlabs * 4;       // u32 [kMaxSlabs+1]
    static constexpr uint32_t slab_n_off = slab_pre_off + (kMaxSlabs + 1) * 4; // u32 [kMaxSlabs]
    static constexpr uint32_t misc_off = (slab_n_off + kMaxSlabs * 4 + 15) & ~15u;   // mbarrier + scalars
    static constexpr uint32_t total = misc_off + 128;
};
constexpr size_t kTileSmemBytes = TileSmem::total;

struct TileParams {
    uint32_t grid;
    uint32_t n_chunks;
    uint32_t qprim;        // primary quality (255 = none)
    uint32_t prim_plane;   // plane id of (group 0, qprim)
};

inline TileParams make_tile_params(uint32_t n_reads, int sm_count) {
    TileParams tp;
    tp.n_chunks = (n_reads + kTileReads - 1) / kTileReads;
    tp.grid = tp.n_chunks;
    tp.qprim = 255;
    tp.prim_plane = 0;
    (void)sm_count;
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

// one (read, 8 columns) unit of a pass: aligned shared loads -> 8 qualities + 8 spread base nibbles
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

__global__ void __launch_bounds__(kTileThreads, 3)
k_deposit_tile(BatchView b, TableView tv, DepositParams dp, TileParams tp, uint32_t* __restrict__ defer_list) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem + TileSmem::tab_off);
    int32_t* s_pos = reinterpret_cast<int32_t*>(smem + TileSmem::pos_off);
    uint32_t* s_len = reinterpret_cast<uint32_t*>(smem + TileSmem::len_off);
    uint32_t* s_qo = reinterpret_cast<uint32_t*>(smem + TileSmem::qo_off);
    uint32_t* s_beg = reinterpret_cast<uint32_t*>(smem + TileSmem::beg_off);
    uint32_t* s_end = reinterpret_cast<uint32_t*>(smem + TileSmem::end_off);
    uint16_t* s_rix = reinterpret_cast<uint16_t*>(smem + TileSmem::rix_off);
    uint16_t* s_items = reinterpret_cast<uint16_t*>(smem + TileSmem::items_off);
    uint32_t* s_slab_a = reinterpret_cast<uint32_t*>(smem + TileSmem::slab_a_off);
    uint32_t* s_slab_pre = reinterpret_cast<uint32_t*>(smem + TileSmem::slab_pre_off);
    uint32_t* s_slab_n = reinterpret_cast<uint32_t*>(smem + TileSmem::slab_n_off);
    uint32_t* s_misc = reinterpret_cast<uint32_t*>(smem + TileSmem::misc_off);
    // s_misc: [0,1] mbarrier, [2] task counter, [3] n_items, [4] cmin, [5] cmax, [6] maxlen, [8..15] warp counts
    const uint32_t bar = sbase + TileSmem::misc_off;
    const uint32_t q_smem = sbase + TileSmem::qual_off + kSlack;     // staged qualities start here
    const uint32_t s_smem = sbase + TileSmem::seq_off + kSlack;      // staged 4-bit bases start here

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t chunk0 = blockIdx.x * kTileReads;
    const uint32_t n_chunk = min((uint32_t)kTileReads, b.n_reads - chunk0);

    // ---- (1) per-read headers: issue the global loads first, then the shared-memory setup
    const bool mine = (uint32_t)tid < n_chunk;
    const uint32_t i = chunk0 + (mine ? tid : 0);
    const int32_t pos = b.pos[i];
    const uint32_t flag = b.flag[i], mapq = b.mapq[i], keep = b.keep[i];
    const uint32_t c0 = b.cigar_off[i], c1 = b.cigar_off[i + 1];
    const uint64_t so0 = b.seq_off[chunk0];
    const uint64_t so_r = b.seq_off[i], so_r1 = b.seq_off[i + 1];

    if (tid == 0) mbar_init(bar, 1);
    if (tid < 8) s_misc[2 + tid] = (tid == 2) ? 0x7FFFFFFFu : 0u;     // [4] = cmin = INT_MAX, the others 0
    for (int k = tid; k < kTabCols * 4; k += kTileThreads) s_tab[k] = 0;
    __syncthreads();

    const uint32_t qprim4 = tp.qprim * 0x01010101u;
    const int mbq = dp.min_bq < 1 ? 1 : (dp.min_bq > 128 ? 128 : dp.min_bq);
    const uint32_t ge_add4 = (uint32_t)(0x80 - mbq) * 0x01010101u;
    const bool ge_all = dp.min_bq <= 0;           // every quality passes

    // ---- (2) classify: filter, "simple" test (one contiguous match run + clips), defer the rest
    uint32_t my_len = 0, my_qstart = 0;
    {
        bool defer = false;
        if (mine && read_passes_filter(flag, mapq, keep, dp.min_mq)) {
            uint32_t qstart = 0, len = 0, phase_c = 0;
            bool simple = (c1 - c0) <= (uint32_t)kMaxCigarSimple && c1 > c0;
            bool has_ref = false;
            for (uint32_t k = c0; k < c1 && simple; ++k) {
                const uint32_t c = b.cigar[k], op = c & 15u, l = c >> 4;
                if (op_is_match(op)) {
                    if (phase_c == 2) simple = false;
                    phase_c = 1; len += l; has_ref = true;
                } else if (op == 4) {
                    if (phase_c == 0) qstart += l; else phase_c = 2;
                } else if (op == 5) {
                    if (phase_c == 1) phase_c = 2;
                } else simple = false;
            }
            if (!simple) {
                // a record with no reference-consuming op at all is skipped everywhere
                bool any_ref = false;
                for (uint32_t k = c0; k < c1; ++k) any_ref |= op_consumes_ref(b.cigar[k] & 15u);
                defer = any_ref;
            } else if (!has_ref || len == 0) {
                // only clips: skipped like the general kernel does
            } else if ((so_r1 - so_r) > kQCap - 32 || !(keep & 2u)) {
                defer = true;       // larger than the stage, or base codes beyond A/C/G/T possible
            } else if (pos < 0 || (int64_t)pos + len > tv.G) {
                atomicAdd(&tv.status[ST_RANGE_ERR], 1u);
            } else {
                my_len = len;
                my_qstart = qstart;
            }
        }
        if (defer) defer_list[atomicAdd(&tv.status[ST_DEFERRED], 1u)] = i;
    }
    // coverage difference array: one atomic per distinct start / end among the warp's reads
    {
        const int32_t ks = my_len ? pos : (int32_t)(0x80000000u + lane);
        const uint32_t ms = __match_any_sync(0xFFFFFFFFu, ks);
        if (my_len && lane == __ffs(ms) - 1) atomicAdd(&tv.covdiff[pos], (int32_t)__popc(ms));
        const int32_t ke = my_len ? (int32_t)(pos + my_len) : (int32_t)(0x80000000u + lane);
        const uint32_t me = __match_any_sync(0xFFFFFFFFu, ke);
        if (my_len && lane == __ffs(me) - 1) atomicAdd(&tv.covdiff[pos + my_len], -(int32_t)__popc(me));
    }
    // ---- (3) compact the active reads; chunk column range
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, my_len != 0);
    if (lane == 0) s_misc[8 + warp] = __popc(bal);
    {
        int32_t lo = my_len ? pos : 0x7FFFFFFF;
        int32_t hi = my_len ? (int32_t)(pos + my_len) : 0;
        uint32_t ml = my_len;
        lo = __reduce_min_sync(0xFFFFFFFFu, lo);
        hi = __reduce_max_sync(0xFFFFFFFFu, hi);
        ml = __reduce_max_sync(0xFFFFFFFFu, ml);
        if (lane == 0 && ml) {
            atomicMin(reinterpret_cast<int32_t*>(&s_misc[4]), lo);
            atomicMax(reinterpret_cast<int32_t*>(&s_misc[5]), hi);
            atomicMax(&s_misc[6], ml);
        }
    }
    __syncthreads();
    uint32_t n_act = 0, my_base = 0;
#pragma unroll
    for (int w = 0; w < kTileWarps; ++w) {
        const uint32_t c = s_misc[8 + w];
        if (w < warp) my_base += c;
        n_act += c;
    }
    if (n_act == 0) return;                     // nothing to deposit from this chunk: no bytes are staged at all
    if (my_len) {
        const uint32_t idx = my_base + __popc(bal & ((1u << lane) - 1u));
        s_pos[idx] = pos; s_len[idx] = my_len;
        s_beg[idx] = (uint32_t)(so_r - so0);
        s_end[idx] = (uint32_t)(so_r1 - so0);
        s_qo[idx] = (uint32_t)(so_r - so0) + my_qstart;
        s_rix[idx] = (uint16_t)tid;
    }
    __syncthreads();
    const uint32_t maxlen = s_misc[6];
    uint32_t phase = 0;

    // ---- (4) sub-ranges of active reads whose bytes fit the stage (one for 150 bp reads)
    uint32_t a0 = 0;
    while (a0 < n_act) {
        const uint64_t qbeg = (so0 + s_beg[a0]) & ~15ull;          // 16-byte aligned start of the staged range
        const uint32_t qbeg_rel = (uint32_t)(qbeg - so0);          // may wrap below zero: only used mod 2^32
        uint32_t a1 = n_act;
        if ((so0 + s_end[n_act - 1]) - qbeg > kQCap) {
            uint32_t lo = a0 + 1, hi = n_act;                       // largest a1 with end[a1-1] - qbeg <= kQCap
            while (lo < hi) {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if ((so0 + s_end[mid - 1]) - qbeg <= kQCap) lo = mid; else hi = mid - 1;
            }
            a1 = lo;
        }
        const uint64_t qend = so0 + s_end[a1 - 1];
        const uint64_t sbeg16 = (qbeg >> 1) & ~15ull;
        const int32_t sn_delta = (int32_t)(qbeg - 2 * sbeg16);      // 0 or 16
        if (tid == 0) {
            const uint32_t qbytes = (uint32_t)(((qend - qbeg) + 15) & ~15ull);
            const uint32_t sbytes = (uint32_t)((((qend + 1) >> 1) - sbeg16 + 15) & ~15ull);
            mbar_expect_tx(bar, qbytes + sbytes);
            if (qbytes) tma_bulk_g2s(q_smem, b.qual + qbeg, qbytes, bar);
            if (sbytes) tma_bulk_g2s(s_smem, b.seq4 + sbeg16, sbytes, bar);
        }
        // column range of this sub-range
        int32_t cmin = s_pos[a0], cmax = (int32_t)s_misc[5];
        if (!(a0 == 0 && a1 == n_act)) {
            __syncthreads();
            if (tid == 0) s_misc[5] = 0;
            __syncthreads();
            for (uint32_t r = a0 + tid; r < a1; r += kTileThreads)
                atomicMax(reinterpret_cast<int32_t*>(&s_misc[5]), (int32_t)(s_pos[r] + s_len[r]));
            __syncthreads();
            cmax = (int32_t)s_misc[5];
        }
        bool waited = false;

        // ---- column windows of kTabCols (one for amplicon / deep shotgun chunks)
        for (int32_t wc0 = cmin; wc0 < cmax; wc0 += kTabCols) {
            const int nslab = min(kMaxSlabs, (cmax - wc0 + kSlabCols - 1) / kSlabCols);
            // per-slab candidate read range [a, a+n) by binary search over the sorted run starts
            if (tid < nslab) {
                const int32_t s_lo = wc0 + tid * kSlabCols, s_hi = s_lo + kSlabCols;
                uint32_t lo = a0, hi = a1;               // first read with pos >= s_hi
                while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (s_pos[m] < s_hi) lo = m + 1; else hi = m; }
                const uint32_t bnd = lo;
                const int64_t thr = (int64_t)s_lo - (int64_t)maxlen;   // first read with pos > thr
                lo = a0; hi = bnd;
                while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if ((int64_t)s_pos[m] <= thr) lo = m + 1; else hi = m; }
                s_slab_a[tid] = lo;
                s_slab_n[tid] = bnd - lo;
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t acc = 0;
                for (int k = 0; k < nslab; ++k) { s_slab_pre[k] = acc; acc += (s_slab_n[k] + 31) >> 5; }
                s_slab_pre[nslab] = acc;
                s_misc[2] = 0;
            }
            __syncthreads();
            const uint32_t n_tasks = s_slab_pre[nslab];
            if (!waited) { mbar_wait(bar, phase); phase ^= 1; waited = true; }   // staged bytes have landed

            // ---- tasks: (slab, group of 32 reads); lane = 8 columns of one read per pass
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
                    uint32_t g0 = ge_all ? 0x80808080u : bytes_ge80(u.q0, ge_add4);
                    uint32_t g1 = ge_all ? 0x80808080u : bytes_ge80(u.q1, ge_add4);
                    if (u.j < 0 || u.j + 8 > u.len) {
                        // partial overlap at a read edge: keep bytes with 0 <= j+b < len
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
                        const uint32_t ord = dp.ord_base + chunk0 + s_rix[r];
                        if (o0) tile_slow_bytes(tv, dp, o0, u.q0, u.sw0, (int64_t)col_lane, ord);
                        if (o1) tile_slow_bytes(tv, dp, o1, u.q1, u.sw1, (int64_t)col_lane + 4, ord);
                    }
                };
                // two independent (read, 8 columns) units per iteration: twice the loads in flight
#pragma unroll 1
                for (uint32_t r = ra + sread; r < rb; r += 16) {
                    const uint32_t r2 = r + 8;
                    const int32_t jA = col_lane - s_pos[r], lenA = (int32_t)s_len[r];
                    const bool onA = jA > -8 && jA < lenA;
                    int32_t jB = 0, lenB = 0;
                    bool onB = false;
                    if (r2 < rb) { jB = col_lane - s_pos[r2]; lenB = (int32_t)s_len[r2]; onB = jB > -8 && jB < lenB; }
                    PassUnit uA, uB;
                    if (onA) uA = pass_load(q_smem, s_smem, (int32_t)(s_qo[r] - qbeg_rel) + jA, sn_delta, jA, lenA);
                    if (onB) uB = pass_load(q_smem, s_smem, (int32_t)(s_qo[r2] - qbeg_rel) + jB, sn_delta, jB, lenB);
                    if (onA) consume(uA, r);
                    if (onB) consume(uB, r2);
                }
                // ---- reduce-scatter over the 8 reads of a pass (lane bits 2..4), fields stay <= 32
                {
                    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
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
                        const uint32_t keepv = b3 ? m4[2 + c] : m4[c];
                        const uint32_t send = b3 ? m4[c] : m4[2 + c];
                        m2[c] = keepv + __shfl_xor_sync(0xFFFFFFFFu, send, 8);
                    }
                    const uint32_t keepv = b2 ? m2[1] : m2[0];
                    const uint32_t send = b2 ? m2[0] : m2[1];
                    const uint32_t v = keepv + __shfl_xor_sync(0xFFFFFFFFu, send, 4);
                    // this lane now owns: word h = b4, allele = 2*b3 + b2, columns col_lane + 4h + {0..3}
                    const int code = (b3 ? 2 : 0) + (b2 ? 1 : 0);
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

            // ---- flush the window: one global RED per non-zero (column, allele); collect first-seen work
            const uint32_t chunk_ord0 = dp.ord_base + chunk0;
            uint32_t* plane = tv.planes[tp.prim_plane];
            uint32_t* first0 = tv.first[0];
            const int ncols = min(kTabCols, cmax - wc0);
            for (int e = tid; e < ncols * 4; e += kTileThreads) {
                const uint32_t v = s_tab[e];
                if (v) {
                    s_tab[e] = 0;
                    const int64_t cell = (int64_t)wc0 * 4 + e;
                    atomicAdd(&plane[cell], v);
                    if (first0[cell] > chunk_ord0 + s_rix[a0]) s_items[atomicAdd(&s_misc[3], 1u)] = (uint16_t)e;
                }
            }
            __syncthreads();
            const uint32_t n_items = s_misc[3];
            if (n_items) {
                // exact first-seen ordinal for new (column, allele) pairs: scan the active reads in order
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
                                const uint32_t qa = (s_qo[r] - qbeg_rel) + (uint32_t)j;
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
                                atomicMin(&first0[(int64_t)col * 4 + (e & 3u)], chunk_ord0 + s_rix[r0 + (__ffs(hb) - 1)]);
                            break;
                        }
                    }
                }
                __syncthreads();
                if (tid == 0) s_misc[3] = 0;
            }
            __syncthreads();
        }
        if (!waited) { mbar_wait(bar, phase); phase ^= 1; }       // never leave a bulk copy in flight
        __syncthreads();
        a0 = a1;
    }
}

// General path over the deferred list, one WARP per read: the lanes stride over the bases of each match
// op (coalesced loads, 32 reductions in flight) while the CIGAR walk itself is warp-uniform.  The count
// of deferred reads lives on the device (status[ST_DEFERRED]).
__global__ void __launch_bounds__(128) k_deposit_general_deferred(BatchView b, TableView tv, DepositParams dp,
                                                                  const uint32_t* __restrict__ list) {
    const uint32_t n = tv.status[ST_DEFERRED];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n; t += warps) {
        const uint32_t i = list[t];
        if (!read_passes_filter(b.flag[i], b.mapq[i], b.keep[i], dp.min_mq)) continue;
        const uint32_t c0 = b.cigar_off[i], c1 = b.cigar_off[i + 1];
        int64_t rlen = 0;
        uint32_t lq = 0;
        for (uint32_t k = c0; k < c1; ++k) {
            const uint32_t c = b.cigar[k], op = c & 15u, len = c >> 4;
            if (op_consumes_ref(op)) rlen += len;
            if (op_consumes_query(op)) lq += len;
        }
        if (rlen == 0) continue;
        const int64_t pos = b.pos[i];
        if (pos < 0 || pos + rlen > tv.G) {
            if (lane == 0) atomicAdd(&tv.status[ST_RANGE_ERR], 1u);
            continue;
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
}

}  // namespace lvc
