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
    static constexpr uint32_t pos_off = tab_off + tab_bytes;                   // i32 [kTileReads]
    static constexpr uint32_t len_off = pos_off + kTileReads * 4;              // u32 [kTileReads]
    static constexpr uint32_t qo_off = len_off + kTileReads * 4;               // u32 [kTileReads]
    static constexpr uint32_t sn_off = qo_off + kTileReads * 4;                // u32 [kTileReads]
    static constexpr uint32_t so_off = sn_off + kTileReads * 4;                // u64 [kTileReads+1] seq_off copy
    static constexpr uint32_t rix_off = so_off + (kTileReads + 1) * 8 + 8;     // u16 [kTileReads] compacted -> chunk index
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

__global__ void __launch_bounds__(kTileThreads, 3)
k_deposit_tile(BatchView b, TableView tv, DepositParams dp, TileParams tp, uint32_t* __restrict__ defer_list) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem + TileSmem::tab_off);
    int32_t* s_pos = reinterpret_cast<int32_t*>(smem + TileSmem::pos_off);
    uint32_t* s_len = reinterpret_cast<uint32_t*>(smem + TileSmem::len_off);
    uint32_t* s_qo = reinterpret_cast<uint32_t*>(smem + TileSmem::qo_off);
    uint32_t* s_sn = reinterpret_cast<uint32_t*>(smem + TileSmem::sn_off);
    uint64_t* s_so = reinterpret_cast<uint64_t*>(smem + TileSmem::so_off);
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

    if (tid == 0) mbar_init(bar, 1);
    for (uint32_t r = tid; r <= n_chunk; r += kTileThreads) s_so[r] = b.seq_off[chunk0 + r];
    for (int i = tid; i < kTabCols * 4; i += kTileThreads) s_tab[i] = 0;
    __syncthreads();

    const uint32_t qprim4 = tp.qprim * 0x01010101u;
    const int mbq = dp.min_bq < 1 ? 1 : (dp.min_bq > 128 ? 128 : dp.min_bq);
    const uint32_t ge_add4 = (uint32_t)(0x80 - mbq) * 0x01010101u;
    const bool ge_all = dp.min_bq <= 0;           // every quality passes
    uint32_t phase = 0;

    // ---- sub-chunks: maximal runs of reads whose staged bytes fit kQCap (one for 150 bp reads) ----
    uint32_t sub0 = 0;
    while (sub0 < n_chunk) {
        // (1) sub-chunk extent (uniform across the CTA: computed from the shared seq_off copy)
        uint32_t sub1 = sub0;
        const uint64_t qbeg = s_so[sub0] & ~15ull;           // 16-byte aligned start of the staged range
        {
            // largest sub1 with s_so[sub1] - qbeg <= kQCap  (binary search, s_so is monotone)
            uint32_t lo = sub0, hi = n_chunk;
            while (lo < hi) {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if (s_so[mid] - qbeg <= kQCap) lo = mid; else hi = mid - 1;
            }
            sub1 = lo;
        }
        const bool oversize = (sub1 == sub0);               // a single read larger than the stage: defer it
        if (oversize) sub1 = sub0 + 1;
        const uint32_t n_sub = sub1 - sub0;
        const uint64_t qend = s_so[sub1];
        // (2) stage the bytes: two TMA bulk copies issued by one thread
        if (!oversize && tid == 0) {
            const uint32_t qbytes = (uint32_t)(((qend - qbeg) + 15) & ~15ull);
            const uint64_t sbeg = qbeg >> 1;                 // qbeg is a multiple of 16 -> sbeg multiple of 8
            const uint64_t sbeg16 = sbeg & ~15ull;
            const uint32_t sbytes = (uint32_t)((((qend + 1) >> 1) - sbeg16 + 15) & ~15ull);
            mbar_expect_tx(bar, qbytes + sbytes);
            if (qbytes) tma_bulk_g2s(q_smem, b.qual + qbeg, qbytes, bar);
            if (sbytes) tma_bulk_g2s(s_smem, b.seq4 + sbeg16, sbytes, bar);
        }
        const uint64_t sbeg16 = (qbeg >> 1) & ~15ull;

        // (3) per-read headers: filter, classify, coverage, shared header arrays
        if (tid < 8) s_misc[2 + tid] = (tid == 2) ? 0x7FFFFFFFu : 0u;     // [4]=cmin=INT_MAX, others 0
        __syncthreads();
        {
            int32_t my_pos = 0x7FFFFFFF;
            uint32_t my_len = 0, my_qo = 0, my_sn = 0;
            if ((uint32_t)tid < n_sub) {
                const uint32_t r = sub0 + tid, i = chunk0 + r;
                const int32_t pos = b.pos[i];
                my_pos = pos;
                const uint32_t flag = b.flag[i], keep = b.keep[i];
                bool defer = false;
                if (read_passes_filter(flag, b.mapq[i], keep, dp.min_mq)) {
                    const uint32_t c0 = b.cigar_off[i], c1 = b.cigar_off[i + 1];
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
                        // is it a record with no reference-consuming op at all? (skipped everywhere)
                        bool any_ref = false;
                        for (uint32_t k = c0; k < c1; ++k) any_ref |= op_consumes_ref(b.cigar[k] & 15u);
                        defer = any_ref;
                    } else if (!has_ref || len == 0) {
                        // no match op (e.g. only clips): skipped like the general kernel does
                    } else if (oversize || !(keep & 2u)) {
                        defer = true;       // larger than the stage, or base codes beyond A/C/G/T possible
                    } else if (pos < 0 || (int64_t)pos + len > tv.G) {
                        atomicAdd(&tv.status[ST_RANGE_ERR], 1u);
                    } else {
                        my_len = len;
                        atomicAdd(&tv.covdiff[pos], 1);
                        atomicAdd(&tv.covdiff[pos + len], -1);
                        my_qo = (uint32_t)(s_so[r] - qbeg) + qstart;
                        my_sn = (uint32_t)(((s_so[r] >> 1) - sbeg16) * 2) + qstart;
                    }
                }
                if (defer) defer_list[atomicAdd(&tv.status[ST_DEFERRED], 1u)] = i;
            }
            // compact the active (simple, kept) reads: later loops never touch dropped / deferred reads
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, my_len != 0);
            if (lane == 0) s_misc[8 + warp] = __popc(bal);
            // chunk column range over the simple reads
            int32_t lo = my_len ? my_pos : 0x7FFFFFFF;
            int32_t hi = my_len ? (int32_t)(my_pos + my_len) : 0;
            uint32_t ml = my_len;
            lo = __reduce_min_sync(0xFFFFFFFFu, lo);
            hi = __reduce_max_sync(0xFFFFFFFFu, hi);
            ml = __reduce_max_sync(0xFFFFFFFFu, ml);
            if (lane == 0) {
                atomicMin(reinterpret_cast<int32_t*>(&s_misc[4]), lo);
                atomicMax(reinterpret_cast<int32_t*>(&s_misc[5]), hi);
                atomicMax(&s_misc[6], ml);
            }
            __syncthreads();
            uint32_t base = 0;
            for (int w = 0; w < warp; ++w) base += s_misc[8 + w];
            if (my_len) {
                const uint32_t idx = base + __popc(bal & ((1u << lane) - 1u));
                s_pos[idx] = my_pos; s_len[idx] = my_len; s_qo[idx] = my_qo; s_sn[idx] = my_sn;
                s_rix[idx] = (uint16_t)tid;
            }
        }
        __syncthreads();
        uint32_t n_act = 0;
        for (int w = 0; w < kTileWarps; ++w) n_act += s_misc[8 + w];
        const int32_t cmin = (int32_t)s_misc[4], cmax = (int32_t)s_misc[5];
        const uint32_t maxlen = s_misc[6];
        const bool any_simple = cmax > cmin && cmin != 0x7FFFFFFF;
        if (!oversize) mbar_wait(bar, phase);      // staged bytes have landed (every thread observes it)
        if (!oversize) phase ^= 1;

        // ---- column windows of kTabCols (one for amplicon / deep shotgun chunks) ----
        for (int32_t wc0 = cmin; any_simple && wc0 < cmax; wc0 += kTabCols) {
            const int nslab = min(kMaxSlabs, (cmax - wc0 + kSlabCols - 1) / kSlabCols);
            // per-slab candidate read range [a, a+n) by binary search over the sorted positions
            if (tid < nslab) {
                const int32_t s_lo = wc0 + tid * kSlabCols, s_hi = s_lo + kSlabCols;
                uint32_t lo = 0, hi = n_act;             // first read with pos >= s_hi
                while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (s_pos[m] < s_hi) lo = m + 1; else hi = m; }
                const uint32_t bnd = lo;
                const int64_t thr = (int64_t)s_lo - (int64_t)maxlen;   // first read with pos > thr
                lo = 0; hi = bnd;
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

            // ---- tasks: (slab, group of 32 reads) ----
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
#pragma unroll 1
                for (uint32_t r = ra + sread; r < rb; r += 8) {
                    uint32_t other = 0, q0 = 0, q1 = 0, sw0 = 0, sw1 = 0;
                    int32_t j = 0;
                    {
                        const int32_t pos = s_pos[r];
                        const int32_t len = (int32_t)s_len[r];
                        j = col_lane - pos;
                        if (j > -8 && j < len) {
                            // --- 8 qualities, byte aligned from three aligned shared words
                            const int32_t qa = (int32_t)s_qo[r] + j;
                            const uint32_t a = q_smem + (uint32_t)(qa & ~3);
                            const uint32_t x0 = lds32(a), x1 = lds32(a + 4), x2 = lds32(a + 8);
                            const uint32_t sh = (uint32_t)(qa & 3) * 8u;
                            q0 = __funnelshift_r(x0, x1, sh);
                            q1 = __funnelshift_r(x1, x2, sh);
                            // --- 8 base nibbles (big-endian within bytes) from two aligned shared words
                            const int32_t ni = (int32_t)s_sn[r] + j;
                            const uint32_t sa = s_smem + (uint32_t)((ni >> 1) & ~3);
                            const uint32_t y0 = __byte_perm(lds32(sa), 0, 0x0123);
                            const uint32_t y1 = __byte_perm(lds32(sa + 4), 0, 0x0123);
                            const uint32_t V = __funnelshift_l(y1, y0, (uint32_t)(ni & 7) * 4u);
                            sw0 = spread_nibbles(V, 0x2233);
                            sw1 = spread_nibbles(V, 0x0011);
                            // --- quality masks
                            uint32_t p0 = bytes_eq80(q0, qprim4), p1 = bytes_eq80(q1, qprim4);
                            uint32_t g0 = ge_all ? 0x80808080u : bytes_ge80(q0, ge_add4);
                            uint32_t g1 = ge_all ? 0x80808080u : bytes_ge80(q1, ge_add4);
                            if (j < 0 || j + 8 > len) {
                                // partial overlap at a read edge: keep bytes with 0 <= j+b < len
                                const int lo = j < 0 ? -j : 0, hi = (len - j) < 8 ? (len - j) : 8;
                                const uint64_t vm = ((hi >= 8 ? ~0ull : ((1ull << (8 * hi)) - 1ull)) &
                                                     ~((1ull << (8 * lo)) - 1ull));
                                const uint32_t v0 = (uint32_t)vm, v1 = (uint32_t)(vm >> 32);
                                p0 &= v0; p1 &= v1; g0 &= v0; g1 &= v1;
                            }
                            const uint32_t o0 = g0 & ~p0, o1 = g1 & ~p1;
                            other = o0 | o1;
                            const uint32_t m0 = p0 >> 7, m1 = p1 >> 7;
                            acc[0][0] += sw0 & m0;        acc[1][0] += sw1 & m1;
                            acc[0][1] += (sw0 >> 1) & m0; acc[1][1] += (sw1 >> 1) & m1;
                            acc[0][2] += (sw0 >> 2) & m0; acc[1][2] += (sw1 >> 2) & m1;
                            acc[0][3] += (sw0 >> 3) & m0; acc[1][3] += (sw1 >> 3) & m1;
                            if (other) {
                                const uint32_t ord = dp.ord_base + chunk0 + sub0 + s_rix[r];
                                if (o0) tile_slow_bytes(tv, dp, o0, q0, sw0, (int64_t)col_lane, ord);
                                if (o1) tile_slow_bytes(tv, dp, o1, q1, sw1, (int64_t)col_lane + 4, ord);
                            }
                        }
                    }
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
            const uint32_t chunk_ord0 = dp.ord_base + chunk0 + sub0;
            uint32_t* plane = tv.planes[tp.prim_plane];
            uint32_t* first0 = tv.first[0];
            const int ncols = min(kTabCols, cmax - wc0);
            for (int i = tid; i < ncols * 4; i += kTileThreads) {
                const uint32_t v = s_tab[i];
                if (v) {
                    s_tab[i] = 0;
                    const int64_t cell = (int64_t)wc0 * 4 + i;
                    atomicAdd(&plane[cell], v);
                    if (first0[cell] > chunk_ord0) s_items[atomicAdd(&s_misc[3], 1u)] = (uint16_t)i;
                }
            }
            __syncthreads();
            const uint32_t n_items = s_misc[3];
            // exact first-seen ordinal for new (column, allele) pairs: scan the chunk's reads in order
            for (uint32_t it = warp; it < n_items; it += kTileWarps) {
                const uint32_t e = s_items[it];
                const int32_t col = wc0 + (int32_t)(e >> 2);
                const uint32_t want = 1u << (e & 3u);
                for (uint32_t r0 = 0; r0 < n_act; r0 += 32) {
                    const uint32_t r = r0 + lane;
                    bool hit = false;
                    if (r < n_act) {
                        const int32_t j = col - s_pos[r];
                        if (j >= 0 && j < (int32_t)s_len[r]) {
                            const uint32_t qa = s_qo[r] + (uint32_t)j;
                            const uint32_t q = (lds32(q_smem + (qa & ~3u)) >> ((qa & 3u) * 8u)) & 255u;
                            const uint32_t ni = s_sn[r] + (uint32_t)j;
                            const uint32_t by = (lds32(s_smem + ((ni >> 1) & ~3u)) >> (((ni >> 1) & 3u) * 8u)) & 255u;
                            const uint32_t nib = (ni & 1u) ? (by & 15u) : (by >> 4);
                            hit = (q == tp.qprim) && (nib == want);
                        }
                    }
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, hit);
                    if (bal) {
                        if (lane == 0) atomicMin(&first0[(int64_t)col * 4 + (e & 3u)], chunk_ord0 + s_rix[r0 + (__ffs(bal) - 1)]);
                        break;
                    }
                }
            }
            __syncthreads();
            if (tid == 0) s_misc[3] = 0;
            __syncthreads();
        }
        __syncthreads();
        sub0 = sub1;
    }
}

// general kernel over the deferred list; the count lives on the device (status[ST_DEFERRED])
__global__ void __launch_bounds__(128) k_deposit_general_deferred(BatchView b, TableView tv, DepositParams dp,
                                                                  const uint32_t* __restrict__ list) {
    const uint32_t n = tv.status[ST_DEFERRED];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
        deposit_read_general(b, tv, dp, list[t]);
}

}  // namespace lvc
