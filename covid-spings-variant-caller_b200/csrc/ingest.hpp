// Native host ingest: BAM (BGZF, multi-threaded inflate) and SAM text -> packed structure-of-arrays batch.
//
// Replaces, for this path, what pysam/htslib do for the reference before the pileup loop:
// pysam.AlignmentFile(inputBam, 'rb') + region iteration (live_variant_caller.py:55-60) and
// pysam.sort for SAM input (client_server/vc_queue.py:34; samtools order: position, forward strand first,
// input order).  Mate overlaps (pysam's ignore_overlaps=True default) are handled at pack time by the admission pass
// of overlap.hpp.  Included by lvc_api.cu (host code only).
#include <zlib.h>
#include <sys/mman.h>
#include "inflate_fast.hpp"
#include "crc32_clmul.hpp"

#include <atomic>
#include <memory>
#include <functional>
#include <mutex>
#include <chrono>
#include <fstream>
#include <sstream>
#include <thread>

struct lvc_reads {
    std::string contig;
    int64_t contig_len = 0;
    std::vector<std::pair<std::string, int64_t>> contigs;
    uint32_t n = 0;
    uint64_t n_cigar = 0, n_qual = 0;
    // arrays (page-locked when a CUDA device is present, so lvc_push_batch can read the payload in place)
    int32_t* pos = nullptr; uint16_t* flag = nullptr; uint8_t* mapq = nullptr; uint8_t* keep = nullptr;
    uint32_t* cigar_off = nullptr; uint32_t* cigar = nullptr; uint64_t* seq_off = nullptr;
    uint8_t* seq4 = nullptr; uint8_t* qual = nullptr;
    uint8_t* qcode = nullptr;                        // 2-bit quality codes (lvc_batch::qual_bits == 2) when the file qualifies
    uint8_t qdict[4] = {0, 0, 0, 0};
    bool codes_tried = false;                        // the code form is made on the first lvc_reads_batch
    uint8_t* scode = nullptr;                        // 2-bit base codes (lvc_batch::seq_form) made by lvc_reads_batch_for
    int scode_min_bq = -1;                           // the base-quality threshold they were made for (-1: not tried)
    int n_threads = 1;
    bool pinned = false;
    uint64_t overlap_pairs = 0, overlap_bases = 0;   // mate pairs / quality bytes rewritten by the overlap model
    std::vector<std::pair<void*, size_t>> allocs;    // (pointer, capacity)
};

namespace ingest {
// LVC_INGEST_TIMING=1: phase times on stderr
struct PhaseTimer {
    bool on; std::chrono::steady_clock::time_point t;
    PhaseTimer() : on(getenv("LVC_INGEST_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[ingest] %-18s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

// Page-locking half a gigabyte costs ~200 ms, more than inflating and packing the file: page-locked buffers of freed
// batches are kept in a small process-wide pool and handed to the next batch (live batches arrive every few seconds
// and have similar sizes).  At most kPoolMaxBytes stay parked.
struct PinnedPool {
    std::mutex mu;
    std::vector<std::pair<void*, size_t>> free_list;
    size_t parked = 0;
    static constexpr size_t kPoolMaxBytes = size_t(4) << 30;
    void* take(size_t bytes, size_t* cap) {
        std::lock_guard<std::mutex> g(mu);
        size_t best = free_list.size();
        for (size_t k = 0; k < free_list.size(); ++k)
            if (free_list[k].second >= bytes && free_list[k].second <= 2 * bytes + (1 << 20) &&
                (best == free_list.size() || free_list[k].second < free_list[best].second)) best = k;
        if (best == free_list.size()) return nullptr;
        void* p = free_list[best].first;
        *cap = free_list[best].second;
        parked -= *cap;
        free_list.erase(free_list.begin() + (long)best);
        return p;
    }
    void give(void* p, size_t cap) {
        {
            std::lock_guard<std::mutex> g(mu);
            if (parked + cap <= kPoolMaxBytes && free_list.size() < 64) { free_list.push_back({p, cap}); parked += cap; return; }
        }
        cudaFreeHost(p);
    }
};
static PinnedPool g_pinned_pool;

static void* host_alloc(lvc_reads* r, size_t bytes) {
    void* p = nullptr;
    bytes = bytes ? bytes : 1;
    size_t cap = bytes;
    if (r->pinned) {
        p = g_pinned_pool.take(bytes, &cap);
        if (!p) {
            cap = (bytes + 4095) & ~size_t(4095);
            if (cudaHostAlloc(&p, cap, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); p = nullptr; }
        }
    }
    if (!p) {
        if (r->pinned && !r->allocs.empty()) return nullptr;   // do not mix: caller falls back as a whole
        r->pinned = false;
        cap = (bytes + 63) & ~size_t(63);
        p = aligned_alloc(64, cap);
    }
    if (p) r->allocs.push_back({p, cap});
    return p;
}

static void free_all(lvc_reads* r) {
    for (auto& a : r->allocs) { if (r->pinned) g_pinned_pool.give(a.first, a.second); else free(a.first); }
    r->allocs.clear();
}

struct Rec {                 // one alignment before packing
    int32_t pos; uint16_t flag; uint8_t mapq;
    uint32_t n_cig, l_seq;
    const uint8_t* cig;      // BAM-encoded u32 ops (little endian) -- or owned storage for SAM
    const uint8_t* seq;      // BAM 4-bit packed
    const uint8_t* qual;     // raw phred
    int32_t next_pos; int8_t next_ref;      // mate: PNEXT (0-based), RNEXT 1 = this contig, 0 = another, -1 = absent
    int32_t tlen;
    const char* name; uint32_t l_name;       // QNAME (not terminated)
};

static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline int32_t rdi32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
static inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

static std::string fail(const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    return buf;
}

// ---- BGZF: block directory, then parallel raw inflate into one contiguous buffer
// a byte buffer that is NOT zero-filled when it grows (std::vector::resize would touch every page twice).  Buffers of
// several MB are anonymous mappings aligned to 2 MB with MADV_HUGEPAGE: the inflate threads first-touch the array, and with
// 4 KB pages those 125,000 page faults per config-2 BAM cost more host time than the decoder itself (measured: 0.53 s of
// thread time for inflate + walk, of which 0.19 s were faults).  Where transparent huge pages are off the advice is a no-op.
// One such mapping is parked between calls (live batches arrive every few seconds and have similar sizes): the next file
// neither faults its pages in again nor pays for unmapping half a gigabyte at the end of the call.  At most kRawParkMax
// bytes stay parked; LVC_INGEST_PARK=0 disables it.
struct RawPark {
    std::mutex mu;
    void* map = nullptr; size_t map_len = 0;
};
static RawPark g_raw_park;
constexpr size_t kRawParkMax = 2ull << 30;

struct RawBuf {
    uint8_t* p = nullptr; size_t n = 0;
    void* map = nullptr; size_t map_len = 0;
    static constexpr size_t kHuge = 2u << 20;
    ~RawBuf() { release(); }
    static uint8_t* aligned(void* q) { return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(q) + kHuge - 1) & ~(uintptr_t)(kHuge - 1)); }
    void release() {
        if (map) {
            const char* env = getenv("LVC_INGEST_PARK");
            bool parked = false;
            if (map_len <= kRawParkMax && !(env && atoi(env) == 0)) {
                std::lock_guard<std::mutex> g(g_raw_park.mu);
                if (!g_raw_park.map || g_raw_park.map_len < map_len) {      // keep the larger one
                    if (g_raw_park.map) munmap(g_raw_park.map, g_raw_park.map_len);
                    g_raw_park.map = map; g_raw_park.map_len = map_len;
                    parked = true;
                }
            }
            if (!parked) munmap(map, map_len);
        } else free(p);
        p = nullptr; n = 0; map = nullptr; map_len = 0;
    }
    bool resize(size_t m) {
        release();
        if (m >= 4 * kHuge) {
            const size_t len = ((m + kHuge - 1) & ~(kHuge - 1)) + kHuge;
            {
                // a parked mapping that is large enough, and not more than four times too large
                std::lock_guard<std::mutex> g(g_raw_park.mu);
                if (g_raw_park.map && g_raw_park.map_len >= len && g_raw_park.map_len / 4 <= len) {
                    map = g_raw_park.map; map_len = g_raw_park.map_len;
                    g_raw_park.map = nullptr; g_raw_park.map_len = 0;
                    p = aligned(map); n = m;
                    return true;
                }
            }
            void* q = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (q != MAP_FAILED) {
                map = q; map_len = len;
                p = aligned(q);
#ifdef MADV_HUGEPAGE
                madvise(p, len - kHuge, MADV_HUGEPAGE);
#endif
                n = m;
                return true;
            }
        }
        p = (uint8_t*)malloc(m ? m : 1); n = p ? m : 0;
        return p != nullptr;
    }
    size_t size() const { return n; }
    uint8_t* data() { return p; }
    const uint8_t* data() const { return p; }
    uint8_t& operator[](size_t i) { return p[i]; }
    const uint8_t& operator[](size_t i) const { return p[i]; }
};

// the input file: a read-only mapping (no copy: the inflate threads fault the pages in), or a buffer for a pipe
struct Bytes {
    const uint8_t* p = nullptr; size_t n = 0;
    size_t size() const { return n; }
    const uint8_t* data() const { return p; }
    const uint8_t& operator[](size_t i) const { return p[i]; }
};

// `follower(raw, total, need)`, if given, runs on a thread of its own WHILE the blocks are inflated: need(end) blocks until
// every byte below `end` is in place (false: inflation failed, give up).  The BAM record walk -- a chain of dependent hops
// that no second thread can help with -- hides behind the inflate that way.
using NeedFn = std::function<bool(size_t)>;
using FollowerFn = std::function<void(const uint8_t*, size_t, const NeedFn&)>;
static std::string bgzf_inflate(const Bytes& file, RawBuf& out, int n_threads, const FollowerFn& follower = nullptr) {
    struct Blk { size_t coff, clen, uoff; uint32_t ulen, crc; };
    std::vector<Blk> blks;
    size_t i = 0, n = file.size(), total = 0;
    while (i + 18 <= n) {
        if (!(file[i] == 0x1f && file[i + 1] == 0x8b && file[i + 2] == 8 && (file[i + 3] & 4))) return fail("not a BGZF/BAM file");
        const uint16_t xlen = rd16(&file[i + 10]);
        size_t j = i + 12, xend = i + 12 + xlen;
        if (xend > n) return fail("corrupt BGZF block header");
        int bsize = -1;
        while (j + 4 <= xend) {
            const uint16_t slen = rd16(&file[j + 2]);
            if (file[j] == 66 && file[j + 1] == 67) bsize = rd16(&file[j + 4]);
            j += 4 + slen;
        }
        // the block (bsize + 1 bytes) holds the 12 + xlen header bytes, the deflate stream and the 8-byte trailer
        if (bsize < 0 || i + bsize + 1 > n || (size_t)bsize + 1 < (xend - i) + 8) return fail("corrupt BGZF block header");
        const uint32_t isize = rd32(&file[i + bsize + 1 - 4]);
        if (isize > 65536u) return fail("corrupt BGZF block (uncompressed size > 64 KiB)");
        blks.push_back({xend, (size_t)(bsize + 1) - (xend - i) - 8, total, isize, rd32(&file[i + bsize + 1 - 8])});
        total += isize;
        i += bsize + 1;
    }
    if (!out.resize(total)) return fail("out of host memory");
    constexpr size_t kGrab = 16;                                 // blocks (~1 MB) per grab
    const size_t n_grabs = (blks.size() + kGrab - 1) / kGrab;
    std::unique_ptr<std::atomic<uint8_t>[]> done(new std::atomic<uint8_t>[n_grabs ? n_grabs : 1]);
    for (size_t g = 0; g < n_grabs; ++g) done[g].store(0, std::memory_order_relaxed);
    std::atomic<size_t> next{0};
    std::atomic<bool> bad{false};
    // the library's own DEFLATE decoder (inflate_fast.hpp: 64-bit bit buffer, one table lookup per symbol; ~2x zlib's
    // inflate on BAM blocks); LVC_INFLATE=zlib selects zlib's inflate() for an A/B measurement.  The CRC-32 of every block
    // is checked either way (htslib checks it too).
    const char* inf_env = getenv("LVC_INFLATE");
    const bool use_zlib = inf_env && strcmp(inf_env, "zlib") == 0;
    auto work = [&]() {
        z_stream zs; memset(&zs, 0, sizeof zs);                 // one inflate state per thread, reset per block
        if (use_zlib && inflateInit2(&zs, -15) != Z_OK) { bad = true; return; }
        std::unique_ptr<lvc_inflate::Tables> tables(use_zlib ? nullptr : new lvc_inflate::Tables());
        for (;;) {
            const size_t k0 = next.fetch_add(kGrab);
            if (k0 >= blks.size() || bad) break;
            for (size_t k = k0; k < std::min(blks.size(), k0 + kGrab); ++k) {
                const Blk& b = blks[k];
                if (b.ulen == 0) continue;
                if (use_zlib) {
                    inflateReset(&zs);
                    zs.next_in = const_cast<Bytef*>(&file[b.coff]); zs.avail_in = (uInt)b.clen;
                    zs.next_out = &out[b.uoff]; zs.avail_out = b.ulen;
                    const int rc = inflate(&zs, Z_FINISH);
                    if (rc != Z_STREAM_END || zs.avail_out != 0) { bad = true; break; }
                } else if (!lvc_inflate::inflate_block(*tables, &file[b.coff], b.clen, n - (b.coff + b.clen), &out[b.uoff], b.ulen)) {
                    bad = true; break;                          // (the 8-byte trailer follows every block's stream)
                }
                if (lvc_crc::crc32_block(&out[b.uoff], b.ulen) != b.crc) { bad = true; break; }
            }
            if (!bad) done[k0 / kGrab].store(1, std::memory_order_release);
        }
        if (use_zlib) inflateEnd(&zs);
    };
    n_threads = std::max(1, std::min(n_threads, 64));
    std::vector<std::thread> th;
    std::thread follow;
    if (follower && n_threads > 1) {
        follow = std::thread([&]() {
            size_t ready = 0, g = 0;                              // bytes below `ready` are in place; g = next grab to wait for
            const NeedFn need = [&](size_t end) -> bool {
                end = std::min(end, total);
                while (ready < end) {
                    if (g >= n_grabs) { ready = total; break; }
                    while (!done[g].load(std::memory_order_acquire)) {
                        if (bad.load()) return false;
                        std::this_thread::yield();
                    }
                    ++g;
                    ready = g * kGrab < blks.size() ? blks[g * kGrab].uoff : total;
                }
                return true;
            };
            follower(out.data(), total, need);
        });
    }
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    if (follow.joinable()) follow.join();
    if (bad) return fail("BGZF inflate failed (corrupt block or CRC mismatch)");
    if (follower && n_threads <= 1) follower(out.data(), total, [](size_t) { return true; });
    return "";
}

struct NibTable {
    uint8_t t[256];
    NibTable() {
        memset(t, 15, sizeof t);                       // unknown letters -> N (htslib seq_nt16_table)
        const char* letters = "=ACMGRSVTWYHKDBN";
        for (int i = 0; i < 16; ++i) {
            t[(uint8_t)letters[i]] = (uint8_t)i;
            if (letters[i] >= 'A' && letters[i] <= 'Z') t[(uint8_t)(letters[i] + 32)] = (uint8_t)i;
        }
    }
};
static const NibTable kNib;
// 1 for a byte of two packed bases unless both nibbles are A, C, G or T (1, 2, 4, 8)
struct NotAcgtPair {
    uint8_t t[256];
    NotAcgtPair() {
        for (int b = 0; b < 256; ++b) {
            const int hi = b >> 4, lo = b & 15;
            const bool ok = (hi == 1 || hi == 2 || hi == 4 || hi == 8) && (lo == 1 || lo == 2 || lo == 4 || lo == 8);
            t[b] = ok ? 0 : 1;
        }
    }
};
static const NotAcgtPair kNotAcgtPair;
#define kAsciiToNib kNib.t

static void release_alloc(lvc_reads* r, void* p) {
    for (size_t k = 0; k < r->allocs.size(); ++k)
        if (r->allocs[k].first == p) {
            if (r->pinned) g_pinned_pool.give(p, r->allocs[k].second); else free(p);
            r->allocs.erase(r->allocs.begin() + (long)k);
            return;
        }
}

// Instrument-binned qualities (at most four distinct values among the admitted reads, after the mate-overlap rewrite):
// the batch also gets the 2-bit code form, which is what lvc_reads_batch hands out (half the payload bytes over PCIe).
// LVC_QUALITY_CODES=0 keeps the byte form only.
static void make_quality_codes(lvc_reads* r, int n_threads) {
    if (r->qcode) { release_alloc(r, r->qcode); r->qcode = nullptr; }
    const char* qc_env = getenv("LVC_QUALITY_CODES");
    if (!r->n || !r->n_qual || (qc_env && atoi(qc_env) == 0)) return;
    uint8_t* codes = (uint8_t*)host_alloc(r, (size_t)(r->n_qual / 4) + 64);
    if (!codes) return;
    const int nd = lvc::pack_quality_codes(r->qual, r->n_qual, r->n, r->keep, r->seq_off, r->cigar_off, r->cigar, n_threads,
                                           r->qdict, codes);
    if (nd > 0) { memset(codes + (r->n_qual + 3) / 4, 0, 64 - 4); r->qcode = codes; }
    else release_alloc(r, codes);
}

// lvc_reads_compact: leave out the reads the admission dropped (keep bit0 clear).  No kernel reads them, so the tables
// that result are the same; first-seen ordinals then number the admitted reads (same order).  The arrays are re-packed
// into fresh buffers (page-locked like the old ones), the old ones go back to the pool.
static int compact(lvc_reads* r, int n_threads) {
    const size_t n = r->n;
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) m += r->keep[i] & 1u;
    if (m == n) return 0;
    std::vector<uint32_t> src(m);
    std::vector<uint64_t> coff(m + 1, 0), soff(m + 1, 0);
    for (size_t i = 0, j = 0; i < n; ++i)
        if (r->keep[i] & 1u) {
            src[j] = (uint32_t)i;
            coff[j + 1] = coff[j] + (r->cigar_off[i + 1] - r->cigar_off[i]);
            soff[j + 1] = soff[j] + (r->seq_off[i + 1] - r->seq_off[i]);
            ++j;
        }
    int32_t* pos = (int32_t*)host_alloc(r, m * 4); uint16_t* flag = (uint16_t*)host_alloc(r, m * 2);
    uint8_t* mapq = (uint8_t*)host_alloc(r, m); uint8_t* keep = (uint8_t*)host_alloc(r, m);
    uint32_t* cigar_off = (uint32_t*)host_alloc(r, (m + 1) * 4); uint32_t* cigar = (uint32_t*)host_alloc(r, coff[m] * 4 + 4);
    uint64_t* seq_off = (uint64_t*)host_alloc(r, (m + 1) * 8);
    uint8_t* seq4 = (uint8_t*)host_alloc(r, soff[m] / 2 + 64); uint8_t* qual = (uint8_t*)host_alloc(r, soff[m] + 64);
    void* fresh[9] = {pos, flag, mapq, keep, cigar_off, cigar, seq_off, seq4, qual};
    for (void* p : fresh)
        if (!p) {                                            // out of (page-locked) memory: the batch stays as it is
            for (void* q : fresh) if (q) release_alloc(r, q);
            return 0;
        }
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (;;) {
            const size_t b0 = next.fetch_add(4096);
            if (b0 >= m) return;
            const size_t b1 = std::min(m, b0 + 4096);
            for (size_t j = b0; j < b1; ++j) {
                const size_t i = src[j];
                pos[j] = r->pos[i]; flag[j] = r->flag[i]; mapq[j] = r->mapq[i]; keep[j] = r->keep[i];
                cigar_off[j] = (uint32_t)coff[j]; seq_off[j] = soff[j];
                memcpy(cigar + coff[j], r->cigar + r->cigar_off[i], (size_t)(coff[j + 1] - coff[j]) * 4);
                const size_t nq = (size_t)(soff[j + 1] - soff[j]);
                memcpy(qual + soff[j], r->qual + r->seq_off[i], nq);
                memcpy(seq4 + soff[j] / 2, r->seq4 + r->seq_off[i] / 2, nq / 2);
            }
        }
    };
    n_threads = std::max(1, std::min(n_threads, 64));
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    cigar_off[m] = (uint32_t)coff[m];
    seq_off[m] = soff[m];
    memset(seq4 + soff[m] / 2, 0, 64);
    memset(qual + soff[m], 0, 64);
    void* old[9] = {r->pos, r->flag, r->mapq, r->keep, r->cigar_off, r->cigar, r->seq_off, r->seq4, r->qual};
    for (void* p : old) release_alloc(r, p);
    r->pos = pos; r->flag = flag; r->mapq = mapq; r->keep = keep; r->cigar_off = cigar_off; r->cigar = cigar;
    r->seq_off = seq_off; r->seq4 = seq4; r->qual = qual;
    r->n = (uint32_t)m; r->n_cigar = coff[m]; r->n_qual = soff[m];
    if (r->qcode) { release_alloc(r, r->qcode); r->qcode = nullptr; }
    if (r->scode) { release_alloc(r, r->scode); r->scode = nullptr; }
    r->scode_min_bq = -1;
    r->codes_tried = false;                                  // made again, for the new layout, when a batch is asked for
    return 1;
}

// pack `recs` (already in coordinate order) into the SoA arrays of `r`
static std::string pack(lvc_reads* r, const Rec* recs, const size_t n, int min_mapq, int max_depth, int n_threads,
                        int overlap_model, PhaseTimer& timer) {
    n_threads = std::max(1, std::min(n_threads, 64));
    // offsets: a two-level scan over the records (sum per slice, exclusive scan of the slices, prefix inside each slice)
    std::unique_ptr<uint64_t[]> coff(new uint64_t[n + 1]), soff(new uint64_t[n + 1]);
    {
        const int nt = n < (1u << 16) ? 1 : n_threads;
        std::vector<uint64_t> csum((size_t)nt + 1, 0), ssum((size_t)nt + 1, 0);
        auto slice = [&](int t, size_t& a, size_t& b) { a = n * (size_t)t / (size_t)nt; b = n * ((size_t)t + 1) / (size_t)nt; };
        auto run = [&](auto&& fn) {
            std::vector<std::thread> th;
            for (int t = 1; t < nt; ++t) th.emplace_back(fn, t);
            fn(0);
            for (auto& x : th) x.join();
        };
        run([&](int t) {
            size_t a, b; slice(t, a, b);
            uint64_t c = 0, q = 0;
            for (size_t i = a; i < b; ++i) { c += recs[i].n_cig; q += recs[i].l_seq + (recs[i].l_seq & 1u); }
            csum[(size_t)t + 1] = c; ssum[(size_t)t + 1] = q;
        });
        for (int t = 0; t < nt; ++t) { csum[(size_t)t + 1] += csum[(size_t)t]; ssum[(size_t)t + 1] += ssum[(size_t)t]; }
        run([&](int t) {
            size_t a, b; slice(t, a, b);
            uint64_t c = csum[(size_t)t], q = ssum[(size_t)t];
            for (size_t i = a; i < b; ++i) { coff[i] = c; soff[i] = q; c += recs[i].n_cig; q += recs[i].l_seq + (recs[i].l_seq & 1u); }
        });
        coff[n] = csum[(size_t)nt]; soff[n] = ssum[(size_t)nt];
    }
    if (coff[n] > 0xFFFFFFFFull) return fail("too many CIGAR operations in one batch");
    r->n = (uint32_t)n; r->n_cigar = coff[n]; r->n_qual = soff[n];
    int ndev = 0;
    r->pinned = cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0;
    if (!r->pinned) cudaGetLastError();
    for (int attempt = 0; attempt < 2; ++attempt) {
        r->pos = (int32_t*)host_alloc(r, n * 4); r->flag = (uint16_t*)host_alloc(r, n * 2);
        r->mapq = (uint8_t*)host_alloc(r, n); r->keep = (uint8_t*)host_alloc(r, n);
        r->cigar_off = (uint32_t*)host_alloc(r, (n + 1) * 4); r->cigar = (uint32_t*)host_alloc(r, coff[n] * 4 + 4);
        r->seq_off = (uint64_t*)host_alloc(r, (n + 1) * 8);
        r->seq4 = (uint8_t*)host_alloc(r, soff[n] / 2 + 64); r->qual = (uint8_t*)host_alloc(r, soff[n] + 64);
        if (r->pos && r->flag && r->mapq && r->keep && r->cigar_off && r->cigar && r->seq_off && r->seq4 && r->qual) break;
        free_all(r);
        if (attempt == 1) return fail("out of host memory");
        r->pinned = false;
    }
    timer.mark("offsets + alloc");
    memset(r->seq4 + soff[n] / 2, 0, 64);
    memset(r->qual + soff[n], 0, 64);
    // mate fields for the overlap pass, filled by the same threads (unpaired data -- ONT, single-end -- has nothing to
    // pair up and skips the overlap bookkeeping altogether)
    std::unique_ptr<int32_t[]> mpos, tlen;
    std::unique_ptr<int8_t[]> mref;
    std::unique_ptr<uint64_t[]> nhash;         // hash of every read name, for the admission pass's name table
    if (overlap_model != LVC_OVERLAP_OFF && n) { mpos.reset(new int32_t[n]); tlen.reset(new int32_t[n]); mref.reset(new int8_t[n]); nhash.reset(new uint64_t[n]); }
    // Two passes over the records, each on all threads.  Pass A fills what the admission reads (position, flag, mapping
    // quality, CIGAR, offsets, mate fields); pass B copies the payload (1.5 bytes per base) and makes the A/C/G/T hint.
    // The admission (htslib's bam_plp_push rule and the mate-overlap hash: sequential by definition) runs on a thread of
    // its own beside pass B -- it never looks at a base or a quality; the quality rewrites of the overlapping pairs it
    // finds are listed and applied, on all threads, once the payload is in place.
    std::atomic<bool> any_pair_seen{false};
    auto run_all = [&](auto&& fn, int reserve) {
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (;;) {
                const size_t b0 = next.fetch_add(4096);
                if (b0 >= n) return;
                fn(b0, std::min(n, b0 + 4096));
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < std::max(1, n_threads - reserve); ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
    };
    run_all([&](size_t b0, size_t b1) {
        bool pair_here = false;
        if (mpos)
            for (size_t i = b0; i < b1; ++i) {
                mpos[i] = recs[i].next_pos; tlen[i] = recs[i].tlen; mref[i] = recs[i].next_ref;
                // (every read's name may be looked up: a read that leaves the pileup, or is dropped at the depth cap,
                // removes the entry of its NAME, as htslib's overlap_remove does)
                nhash[i] = lvc_overlap::name_hash64(recs[i].name, recs[i].l_name);
                pair_here |= (recs[i].flag & 0x3u) == 0x3u && !(recs[i].flag & 0x8u);
            }
        if (pair_here) any_pair_seen.store(true, std::memory_order_relaxed);
        for (size_t i = b0; i < b1; ++i) {
            const Rec& x = recs[i];
            r->pos[i] = x.pos; r->flag[i] = x.flag; r->mapq[i] = x.mapq;
            r->cigar_off[i] = (uint32_t)coff[i]; r->seq_off[i] = soff[i];
            memcpy(r->cigar + coff[i], x.cig, (size_t)x.n_cig * 4);
        }
    }, 0);
    r->cigar_off[n] = (uint32_t)coff[n];
    r->seq_off[n] = soff[n];
    timer.mark("pack: per-read arrays");
    std::vector<uint8_t> adm(n ? n : 1);
    const bool any_pair = any_pair_seen.load();
    std::vector<lvc_overlap::PendingTweak> pend;
    int rc = 0;
    auto admit = [&]() {
        auto name = [&](uint32_t i) { return lvc_overlap::NameKey{recs[i].name, recs[i].l_name}; };
        rc = lvc_overlap::admit_core((uint32_t)n, r->pos, r->flag, r->mapq, r->cigar_off, r->cigar, r->seq_off, r->seq4,
                                     r->qual, name, any_pair ? mpos.get() : nullptr, any_pair ? mref.get() : nullptr,
                                     any_pair ? tlen.get() : nullptr, min_mapq, max_depth,
                                     any_pair ? overlap_model : LVC_OVERLAP_OFF, adm.data(), nullptr, nullptr, &pend,
                                     any_pair ? nhash.get() : nullptr);
    };
    std::thread admit_thread;
    if (n_threads > 1) admit_thread = std::thread(admit);
    run_all([&](size_t b0, size_t b1) {
        for (size_t i = b0; i < b1; ++i) {
            const Rec& x = recs[i];
            const size_t nb = (x.l_seq + 1) / 2;
            uint8_t* sq = r->seq4 + soff[i] / 2;
            memcpy(sq, x.seq, nb);
            uint8_t* q = r->qual + soff[i];
            memcpy(q, x.qual, x.l_seq);
            // A/C/G/T-only hint: two bases per table lookup (the pad nibble of an odd-length read is not a base)
            uint32_t bad_b = 0;
            const size_t full = x.l_seq / 2;
            for (size_t k = 0; k < full; ++k) bad_b |= kNotAcgtPair.t[sq[k]];
            if (x.l_seq & 1) { bad_b |= kNotAcgtPair.t[(sq[nb - 1] & 0xF0u) | 1u]; q[x.l_seq] = 0; sq[nb - 1] &= 0xF0; }
            const bool acgt = bad_b == 0;
            r->keep[i] = acgt ? 2 : 0;          // bit1: ACGT-only hint; bit0 is OR-ed in after admission
        }
    }, n_threads > 2 ? 1 : 0);
    if (admit_thread.joinable()) admit_thread.join(); else admit();
    timer.mark("pack: payload | admission");
    if (rc == LVC_EUNSORTED) return fail("reads are not coordinate sorted");
    if (rc) return fail("admission failed (%d)", rc);
    r->overlap_pairs = 0; r->overlap_bases = 0;
    if (!pend.empty()) {
        const size_t np = pend.size();
        const int nt = (int)std::min<size_t>((size_t)n_threads, (np + 1023) / 1024);
        std::vector<uint64_t> pr((size_t)nt, 0), bs((size_t)nt, 0);
        std::vector<std::thread> th;
        auto part = [&](int t) {
            lvc_overlap::apply_pending(pend.data(), np * (size_t)t / (size_t)nt, np * ((size_t)t + 1) / (size_t)nt, r->pos, r->cigar_off,
                                       r->cigar, r->seq_off, r->seq4, r->qual, overlap_model, &pr[(size_t)t], &bs[(size_t)t]);
        };
        for (int t = 1; t < nt; ++t) th.emplace_back(part, t);
        part(0);
        for (auto& t : th) t.join();
        for (int t = 0; t < nt; ++t) { r->overlap_pairs += pr[(size_t)t]; r->overlap_bases += bs[(size_t)t]; }
    }
    timer.mark("mate-overlap rewrites");
    for (size_t i = 0; i < n; ++i) r->keep[i] |= adm[i];
    timer.mark("keep bits");
    r->n_threads = n_threads;
    return "";
}

static std::string validate(const Rec& x, const char* what) {
    uint64_t lq = 0, rl = 0;
    for (uint32_t k = 0; k < x.n_cig; ++k) {
        uint32_t c; memcpy(&c, x.cig + 4 * k, 4);
        const uint32_t op = c & 15u, len = c >> 4;
        if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) lq += len;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += len;
    }
    if (rl > 0 && lq != x.l_seq) return fail("%s: read at %d: CIGAR query length %llu != sequence length %u", what, x.pos, (unsigned long long)lq, x.l_seq);
    if (rl > 0 && x.l_seq && x.qual[0] == 0xFF && !(x.flag & 4)) return fail("%s: a read has no base qualities (the reference raises TypeError)", what);
    return "";
}

static std::string read_bam(lvc_reads* r, const Bytes& file, const char* contig, int min_mapq, int max_depth,
                            int n_threads, RawBuf& raw, int overlap_model, PhaseTimer& timer) {
    // header + pass 1 (serial, a hop per record: where the records of this contig start) follow the inflate threads
    int tid = -1;
    std::vector<size_t> offs;
    std::string scan_err;
    const FollowerFn scan = [&](const uint8_t* raw_p, size_t raw_n, const NeedFn& need) {
        auto bail = [&](const std::string& m) { scan_err = m; };
        if (!need(12)) return;
        if (raw_n < 12 || memcmp(raw_p, "BAM\1", 4) != 0) return bail(fail("bad BAM magic"));
        const int32_t l_text = rdi32(raw_p + 4);
        if (l_text < 0 || (size_t)l_text > raw_n - 12) return bail(fail("truncated BAM header"));
        size_t off = 8 + (size_t)l_text;
        if (off + 4 > raw_n) return bail(fail("truncated BAM header"));
        if (!need(off + 4)) return;
        const int32_t n_ref = rdi32(raw_p + off); off += 4;
        if (n_ref < 0) return bail(fail("corrupt BAM header"));
        for (int32_t k = 0; k < n_ref; ++k) {
            if (off + 8 > raw_n) return bail(fail("truncated BAM header"));
            if (!need(off + 4)) return;
            const int32_t l_name = rdi32(raw_p + off);
            if (l_name < 1 || (size_t)l_name > raw_n - off - 8) return bail(fail("corrupt BAM header"));
            if (!need(off + 8 + (size_t)l_name)) return;
            std::string name((const char*)raw_p + off + 4, (size_t)std::max(0, l_name - 1));
            const int32_t l_ref = rdi32(raw_p + off + 4 + l_name);
            r->contigs.push_back({name, l_ref});
            off += 8 + l_name;
        }
        for (size_t k = 0; k < r->contigs.size(); ++k)
            if ((!contig || !*contig) ? k == 0 : r->contigs[k].first == contig) { tid = (int)k; break; }
        if (tid < 0) return bail(fail("invalid contig `%s`", contig ? contig : ""));
        r->contig = r->contigs[tid].first; r->contig_len = r->contigs[tid].second;
        offs.reserve(raw_n / 256);
        // (every hop is a dependent cache miss; records of one run have similar sizes, so the line a few records ahead
        // is requested by extrapolation: right most of the time, harmless when not)
        while (off + 4 <= raw_n) {
            if (!need(off + 8)) return;
            const int32_t bs = rdi32(raw_p + off);
            if (bs < 32 || off + 4 + (size_t)bs > raw_n) return bail(fail("truncated BAM record"));
            const size_t step = 4 + (size_t)bs;
            if (off + 12 * step + 64 < raw_n) {
                __builtin_prefetch(raw_p + off + 6 * step); __builtin_prefetch(raw_p + off + 6 * step + 64);
                __builtin_prefetch(raw_p + off + 12 * step); __builtin_prefetch(raw_p + off + 12 * step + 64);
            }
            if (rdi32(raw_p + off + 4) == tid) offs.push_back(off);
            off += step;
        }
    };
    std::string e = bgzf_inflate(file, raw, n_threads, scan);
    if (!e.empty()) return e;
    if (!scan_err.empty()) return scan_err;
    timer.mark("bgzf inflate + walk");
    // pass 2 (threads): decode and validate; the error of the first bad record wins
    std::unique_ptr<Rec[]> recs(new Rec[offs.size() ? offs.size() : 1]);     // not zero-filled: every slot is written below
    std::atomic<size_t> next{0};
    std::mutex emu;
    size_t err_at = offs.size();
    std::string err_msg;
    auto work = [&]() {
        for (;;) {
            const size_t i0 = next.fetch_add(8192);
            if (i0 >= offs.size()) return;
            for (size_t i = i0; i < std::min(offs.size(), i0 + 8192); ++i) {
                const int32_t bs = rdi32(&raw[offs[i]]);
                const uint8_t* p = &raw[offs[i] + 4];
                Rec x;
                x.pos = rdi32(p + 4);
                const uint32_t l_rn = p[8];
                x.mapq = p[9];
                x.n_cig = rd16(p + 12); x.flag = rd16(p + 14); x.l_seq = (uint32_t)rdi32(p + 16);
                const int32_t nref = rdi32(p + 20);
                x.next_ref = nref < 0 ? -1 : (nref == tid ? 1 : 0); x.next_pos = rdi32(p + 24); x.tlen = rdi32(p + 28);
                x.name = (const char*)(p + 32); x.l_name = l_rn ? l_rn - 1 : 0;
                x.cig = p + 32 + l_rn;
                x.seq = x.cig + 4 * (size_t)x.n_cig;
                x.qual = x.seq + (x.l_seq + 1) / 2;
                std::string v;
                if (x.qual + x.l_seq > p + bs || (int32_t)x.l_seq < 0) v = fail("corrupt BAM record");
                if (v.empty() && x.n_cig == 2) {
                    uint32_t c0; memcpy(&c0, x.cig, 4);
                    uint32_t c1; memcpy(&c1, x.cig + 4, 4);
                    if ((c0 & 15u) == 4 && (c0 >> 4) == x.l_seq && (c1 & 15u) == 3)
                        v = fail("CIGAR stored in the CG tag (> 65535 operations) is not supported");
                }
                if (v.empty()) v = validate(x, "BAM");
                if (!v.empty()) {
                    std::lock_guard<std::mutex> g(emu);
                    if (i < err_at) { err_at = i; err_msg = v; }
                    x.n_cig = 0; x.l_seq = 0;
                } else {
                    uint64_t rl = 0;
                    for (uint32_t k = 0; k < x.n_cig; ++k) { uint32_t c; memcpy(&c, x.cig + 4 * k, 4); const uint32_t op = c & 15u; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += c >> 4; }
                    if (rl == 0) { x.n_cig = 0; x.l_seq = 0; }       // never reaches the pileup: keep the slot only
                }
                recs[i] = x;
            }
        }
    };
    {
        const int nt = std::max(1, std::min(n_threads, 64));
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
    }
    if (err_at < offs.size()) return err_msg;
    timer.mark("record scan");
    return pack(r, recs.get(), offs.size(), min_mapq, max_depth, n_threads, overlap_model, timer);
}

static inline int64_t parse_int(const char* p, size_t n) {
    int64_t v = 0; bool neg = false; size_t k = 0;
    if (k < n && (p[k] == '-' || p[k] == '+')) { neg = p[k] == '-'; ++k; }
    for (; k < n && p[k] >= '0' && p[k] <= '9'; ++k) v = v * 10 + (p[k] - '0');
    return neg ? -v : v;
}

static std::string read_sam(lvc_reads* r, const Bytes& file, const char* contig, int min_mapq, int max_depth,
                            int n_threads, int overlap_model, PhaseTimer& timer) {
    // text -> per-record binary blobs (cigar u32s, packed seq, qual), parsed by all threads on line-aligned pieces
    // of the file, then the samtools-order sort and the common packer
    struct Tmp { Rec rec; size_t cig_o, seq_o, qual_o; uint32_t piece; };
    std::string want = contig ? contig : "";
    const char* p = (const char*)file.data();
    const char* end = p + file.size();
    // header lines (top of the file)
    while (p < end && *p == '@') {
        const char* nl = (const char*)memchr(p, '\n', end - p);
        const char* le = nl ? nl : end;
        if (le > p && le[-1] == '\r') --le;
        if (le - p > 3 && p[1] == 'S' && p[2] == 'Q') {
            std::string line(p, le), name; int64_t ln = 0;
            size_t a = 0;
            while (a < line.size()) {
                size_t b = line.find('\t', a); if (b == std::string::npos) b = line.size();
                if (line.compare(a, 3, "SN:") == 0) name = line.substr(a + 3, b - a - 3);
                if (line.compare(a, 3, "LN:") == 0) ln = atoll(line.c_str() + a + 3);
                a = b + 1;
            }
            r->contigs.push_back({name, ln});
        }
        p = nl ? nl + 1 : end;
    }
    if (want.empty()) {
        if (!r->contigs.empty()) want = r->contigs[0].first;
        else {                                        // headerless SAM: the contig of the first record
            const char* q = p;
            while (q < end && want.empty()) {
                const char* nl = (const char*)memchr(q, '\n', end - q);
                const char* le = nl ? nl : end;
                const char* t1 = (const char*)memchr(q, '\t', le - q);
                const char* t2 = t1 ? (const char*)memchr(t1 + 1, '\t', le - t1 - 1) : nullptr;
                const char* t3 = t2 ? (const char*)memchr(t2 + 1, '\t', le - t2 - 1) : nullptr;
                if (t3 && *q != '@') want.assign(t2 + 1, t3);
                q = nl ? nl + 1 : end;
            }
        }
    }
    r->contig = want;
    for (auto& c : r->contigs) if (c.first == want) r->contig_len = c.second;
    if (!r->contigs.empty() && r->contig_len == 0 && contig && *contig) return fail("invalid contig `%s`", contig);
    // line-aligned pieces of the body
    const int nt = std::max(1, std::min(n_threads, 64));
    const size_t n_piece = (size_t)nt * 4;
    std::vector<const char*> cut(n_piece + 1, end);
    cut[0] = p;
    for (size_t k = 1; k < n_piece; ++k) {
        const char* c = p + (size_t)((end - p) / (double)n_piece * (double)k);
        if (c < cut[k - 1]) c = cut[k - 1];
        const char* nl = c < end ? (const char*)memchr(c, '\n', end - c) : nullptr;
        cut[k] = nl ? nl + 1 : end;
    }
    struct Piece { std::vector<Tmp> tmp; std::vector<uint8_t> store; std::string err; };
    std::vector<Piece> pieces(n_piece);
    std::atomic<size_t> next{0};
    auto parse_piece = [&](size_t pk) {
        Piece& pc = pieces[pk];
        const char* q = cut[pk];
        const char* pend = cut[pk + 1];
        pc.store.reserve((size_t)(pend - q));
        while (q < pend) {
            const char* nl = (const char*)memchr(q, '\n', pend - q);
            const char* le = nl ? nl : pend;
            const char* next_line = nl ? nl + 1 : pend;
            if (le > q && le[-1] == '\r') --le;
            if (q < le && *q != '@') {
                const char* f[11]; size_t fl[11]; int nf = 0;
                const char* a = q;
                while (nf < 11 && a <= le) {
                    const char* b = (const char*)memchr(a, '\t', le - a); if (!b) b = le;
                    f[nf] = a; fl[nf] = (size_t)(b - a); ++nf; a = b + 1;
                }
                if (nf == 11 && fl[2] == want.size() && memcmp(f[2], want.data(), fl[2]) == 0) {
                    Tmp t;
                    t.piece = (uint32_t)pk;
                    t.rec.flag = (uint16_t)parse_int(f[1], fl[1]);
                    t.rec.pos = (int32_t)parse_int(f[3], fl[3]) - 1;
                    t.rec.mapq = (uint8_t)parse_int(f[4], fl[4]);
                    const bool next_same = (fl[6] == 1 && f[6][0] == '=') || (fl[6] == want.size() && memcmp(f[6], want.data(), fl[6]) == 0);
                    t.rec.next_ref = next_same ? 1 : ((fl[6] == 1 && f[6][0] == '*') ? -1 : 0);
                    t.rec.next_pos = (int32_t)parse_int(f[7], fl[7]) - 1;
                    t.rec.tlen = (int32_t)parse_int(f[8], fl[8]);
                    t.rec.name = f[0]; t.rec.l_name = (uint32_t)fl[0];
                    // CIGAR
                    t.cig_o = pc.store.size(); t.rec.n_cig = 0;
                    if (!(fl[5] == 1 && f[5][0] == '*')) {
                        uint32_t num = 0;
                        for (size_t k = 0; k < fl[5]; ++k) {
                            const char c = f[5][k];
                            if (c >= '0' && c <= '9') num = num * 10 + (uint32_t)(c - '0');
                            else {
                                const char* ops = "MIDNSHP=XB"; const char* o = strchr(ops, c);
                                if (!o || !c) { pc.err = fail("SAM: bad CIGAR operation '%c'", c); return; }
                                if (num) { const uint32_t v = (num << 4) | (uint32_t)(o - ops); const uint8_t* vb = (const uint8_t*)&v; pc.store.insert(pc.store.end(), vb, vb + 4); t.rec.n_cig++; }
                                num = 0;
                            }
                        }
                    }
                    const bool has_seq = !(fl[9] == 1 && f[9][0] == '*');
                    t.rec.l_seq = has_seq ? (uint32_t)fl[9] : 0;
                    t.seq_o = pc.store.size();
                    pc.store.resize(t.seq_o + (t.rec.l_seq + 1) / 2 + t.rec.l_seq);
                    uint8_t* sq = pc.store.data() + t.seq_o;
                    for (uint32_t k = 0; k < t.rec.l_seq; k += 2) {
                        const uint8_t hi = kAsciiToNib[(uint8_t)f[9][k]];
                        const uint8_t lo = k + 1 < t.rec.l_seq ? kAsciiToNib[(uint8_t)f[9][k + 1]] : 0;
                        sq[k >> 1] = (uint8_t)(hi << 4 | lo);
                    }
                    t.qual_o = t.seq_o + (t.rec.l_seq + 1) / 2;
                    uint8_t* ql = pc.store.data() + t.qual_o;
                    const bool has_q = !(fl[10] == 1 && f[10][0] == '*');
                    if (has_q && has_seq && fl[10] != fl[9]) { pc.err = fail("SAM: QUAL length %zu != SEQ length %zu", fl[10], fl[9]); return; }
                    for (uint32_t k = 0; k < t.rec.l_seq; ++k) ql[k] = has_q ? (uint8_t)(f[10][k] - 33) : 0xFF;
                    pc.tmp.push_back(t);
                }
            }
            q = next_line;
        }
    };
    auto work = [&]() {
        for (;;) {
            const size_t pk = next.fetch_add(1);
            if (pk >= n_piece) return;
            parse_piece(pk);
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
    }
    size_t total = 0;
    for (auto& pc : pieces) { if (!pc.err.empty()) return pc.err; total += pc.tmp.size(); }
    timer.mark("sam parse");
    // samtools sort: position, forward before reverse strand, input order.  The records stay where the parser put them;
    // what is sorted is one 64-bit key per record (position | strand | input index), and a file that is in order already --
    // the usual case -- is not sorted at all.
    std::vector<const Tmp*> order(total);
    {
        size_t k = 0;
        for (auto& pc : pieces) for (auto& t : pc.tmp) order[k++] = &t;
    }
    bool sorted = true;
    for (size_t i = 1; i < total && sorted; ++i) {
        const Rec &x = order[i - 1]->rec, &y = order[i]->rec;
        sorted = x.pos < y.pos || (x.pos == y.pos && ((x.flag >> 4) & 1) <= ((y.flag >> 4) & 1));
    }
    if (!sorted) {
        if (total >= (1ull << 31)) return fail("too many records in one SAM file");
        std::vector<uint64_t> key(total);
        for (size_t i = 0; i < total; ++i) {
            const Rec& x = order[i]->rec;
            key[i] = ((uint64_t)(uint32_t)((int64_t)x.pos + 2) << 32) | ((uint64_t)((x.flag >> 4) & 1) << 31) | (uint64_t)i;   // pos >= -1
        }
        std::sort(key.begin(), key.end());
        std::vector<const Tmp*> by_key(total);
        for (size_t i = 0; i < total; ++i) by_key[i] = order[(size_t)(key[i] & 0x7FFFFFFFu)];
        order.swap(by_key);
    }
    timer.mark("sort");
    // pointers, validation and the reference length of every record, on all threads; the error of the first bad record wins
    std::vector<Rec> recs(total);
    {
        std::atomic<size_t> next{0};
        std::mutex emu;
        size_t err_at = total;
        std::string err_msg;
        auto work = [&]() {
            for (;;) {
                const size_t i0 = next.fetch_add(2048);
                if (i0 >= total) return;
                for (size_t i = i0; i < std::min(total, i0 + 2048); ++i) {
                    const Tmp& t = *order[i];
                    Rec x = t.rec;
                    const uint8_t* base = pieces[t.piece].store.data();
                    x.cig = base + t.cig_o; x.seq = base + t.seq_o; x.qual = base + t.qual_o;
                    std::string v = validate(x, "SAM");
                    if (!v.empty()) {
                        std::lock_guard<std::mutex> g(emu);
                        if (i < err_at) { err_at = i; err_msg = v; }
                    }
                    uint64_t rl = 0;
                    for (uint32_t k = 0; k < x.n_cig; ++k) { uint32_t c; memcpy(&c, x.cig + 4 * k, 4); const uint32_t op = c & 15u; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += c >> 4; }
                    if (rl == 0) { x.n_cig = 0; x.l_seq = 0; }
                    recs[i] = x;
                }
            }
        };
        const int nt = total < 4096 ? 1 : std::max(1, std::min(n_threads, 64));
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
        if (err_at < total) return err_msg;
    }
    timer.mark("validate");
    return pack(r, recs.data(), recs.size(), min_mapq, max_depth, n_threads, overlap_model, timer);
}

}  // namespace ingest

extern "C" {

int lvc_read_alignments_ex(const char* path, const char* contig, int min_mapq, int max_depth, int n_threads,
                           int overlap_model, lvc_reads** out, char* errbuf, int errlen) {
    NvtxRange nvtx_range("lvc_read_alignments");
    ingest::PhaseTimer timer;                                    // per call: the entry point is re-entrant
    auto seterr = [&](const std::string& s) { if (errbuf && errlen > 0) snprintf(errbuf, (size_t)errlen, "%s", s.c_str()); };
    if (!path || !out || overlap_model < LVC_OVERLAP_OFF || overlap_model > LVC_OVERLAP_HTSLIB_1_13) { seterr("bad arguments"); return LVC_EINVAL; }
    // a regular file is mapped read-only (no copy; the worker threads fault its pages in); anything else (a pipe) is
    // read in blocks
    FILE* fh = fopen(path, "rb");
    if (!fh) { seterr(std::string("cannot open ") + path); return LVC_EIO; }
    std::vector<uint8_t> filebuf;
    ingest::Bytes file;
    void* mapped = nullptr; size_t mapped_len = 0;
    struct Unmap { void*& p; size_t& n; ~Unmap() { if (p) munmap(p, n); } } unmap_guard{mapped, mapped_len};
    if (fseek(fh, 0, SEEK_END) == 0) {
        const long sz = ftell(fh);
        rewind(fh);
        if (sz > 0) {
            void* m = mmap(nullptr, (size_t)sz, PROT_READ, MAP_PRIVATE, fileno(fh), 0);
            if (m != MAP_FAILED) {
                mapped = m; mapped_len = (size_t)sz;
                madvise(m, (size_t)sz, MADV_WILLNEED);
                file.p = (const uint8_t*)m; file.n = (size_t)sz;
            } else {
                filebuf.resize((size_t)sz);
                filebuf.resize(fread(filebuf.data(), 1, (size_t)sz, fh));
                file.p = filebuf.data(); file.n = filebuf.size();
            }
        }
    } else {                                                     // not seekable (a pipe): read in blocks
        uint8_t buf[1 << 16];
        size_t got;
        while ((got = fread(buf, 1, sizeof buf, fh)) > 0) filebuf.insert(filebuf.end(), buf, buf + got);
        file.p = filebuf.data(); file.n = filebuf.size();
    }
    fclose(fh);
    timer.mark("file read");
    lvc_reads* r = new lvc_reads();
    ingest::RawBuf rawbuf;
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::string e;
    if (file.size() >= 2 && file[0] == 0x1f && file[1] == 0x8b) e = ingest::read_bam(r, file, contig, min_mapq, max_depth, n_threads, rawbuf, overlap_model, timer);
    else e = ingest::read_sam(r, file, contig, min_mapq, max_depth, n_threads, overlap_model, timer);
    if (!e.empty()) {
        seterr(e);
        const bool unsorted = e.find("not coordinate sorted") != std::string::npos;
        ingest::free_all(r);
        delete r;
        return unsorted ? LVC_EUNSORTED : LVC_EINVAL;
    }
    timer.mark("total tail");
    *out = r;
    return LVC_OK;
}

int lvc_read_alignments(const char* path, const char* contig, int min_mapq, int max_depth, int n_threads,
                        lvc_reads** out, char* errbuf, int errlen) {
    return lvc_read_alignments_ex(path, contig, min_mapq, max_depth, n_threads, LVC_OVERLAP_DEFAULT, out, errbuf, errlen);
}

int lvc_reads_overlap_stats(const lvc_reads* r, uint64_t* n_pairs, uint64_t* n_bases) {
    if (!r) return LVC_EINVAL;
    if (n_pairs) *n_pairs = r->overlap_pairs;
    if (n_bases) *n_bases = r->overlap_bases;
    return LVC_OK;
}

int lvc_reads_batch_bytes(const lvc_reads* r, lvc_batch* b) {
    if (!r || !b) return LVC_EINVAL;
    b->n_reads = r->n; b->qual_bits = 8; b->seq_form = 0; b->n_cigar_ops = r->n_cigar; b->n_qual_bytes = r->n_qual;
    b->pos = r->pos; b->flag = r->flag; b->mapq = r->mapq; b->keep = r->keep; b->cigar_off = r->cigar_off;
    b->cigar = r->cigar; b->seq_off = r->seq_off; b->seq4 = r->seq4; b->qual = r->qual;
    memset(b->qual_dict, 0, 4);
    return LVC_OK;
}

int lvc_reads_batch(const lvc_reads* r, lvc_batch* b) {
    const int rc = lvc_reads_batch_bytes(r, b);
    if (rc == LVC_OK && !r->codes_tried) {
        lvc_reads* w = const_cast<lvc_reads*>(r);            // a cache inside the object: the batch itself is unchanged
        ingest::PhaseTimer timer;
        ingest::make_quality_codes(w, w->n_threads);
        w->codes_tried = true;
        timer.mark("quality codes");
    }
    if (rc == LVC_OK && r->qcode) { b->qual_bits = 2; b->qual = r->qcode; memcpy(b->qual_dict, r->qdict, 4); }
    return rc;
}

int lvc_reads_batch_for(const lvc_reads* r, int min_base_quality, lvc_batch* b) {
    const int rc = lvc_reads_batch(r, b);
    if (rc != LVC_OK || b->qual_bits != 2 || !r->n || !r->n_qual) return rc;
    const char* env = getenv("LVC_BASE_CODES");
    if (env && atoi(env) == 0) return rc;
    lvc_reads* w = const_cast<lvc_reads*>(r);                // a cache inside the object, as for the quality codes
    const int mbq = std::max(0, std::min(min_base_quality, 255));
    // codes made for a threshold stay valid for every higher one; a failed attempt is only repeated for a higher threshold,
    // under which more non-A/C/G/T bases drop out
    const bool make = w->scode_min_bq < 0 || (w->scode ? w->scode_min_bq > mbq : w->scode_min_bq < mbq);
    if (make) {
        ingest::PhaseTimer timer;
        if (w->scode) { ingest::release_alloc(w, w->scode); w->scode = nullptr; }
        uint8_t* codes = (uint8_t*)ingest::host_alloc(w, (size_t)(w->n_qual / 4) + 64);
        if (codes) {
            if (lvc::pack_base_codes(w->seq4, w->qual, w->n_qual, w->n, w->keep, w->seq_off, w->cigar_off, w->cigar, mbq,
                                     w->n_threads, codes)) {
                memset(codes + (w->n_qual + 3) / 4, 0, 64 - 4);
                w->scode = codes;
            } else ingest::release_alloc(w, codes);
        }
        w->scode_min_bq = mbq;
        timer.mark("base codes");
    }
    if (r->scode && r->scode_min_bq <= mbq) { b->seq4 = r->scode; b->seq_form = 2u | ((uint32_t)r->scode_min_bq << 8); }
    return rc;
}

int lvc_reads_compact(lvc_reads* r, int n_threads) {
    if (!r) return LVC_EINVAL;
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    ingest::PhaseTimer timer;
    const int rc = ingest::compact(r, n_threads);
    timer.mark("compact");
    return rc;
}

int lvc_reads_info(const lvc_reads* r, char* contig_name, int name_cap, int64_t* contig_len, int* n_contigs, int* pinned) {
    if (!r) return LVC_EINVAL;
    if (contig_name && name_cap > 0) snprintf(contig_name, (size_t)name_cap, "%s", r->contig.c_str());
    if (contig_len) *contig_len = r->contig_len;
    if (n_contigs) *n_contigs = (int)r->contigs.size();
    if (pinned) *pinned = r->pinned ? 1 : 0;
    return LVC_OK;
}

void lvc_reads_free(lvc_reads* r) {
    if (!r) return;
    ingest::free_all(r);
    delete r;
}

}  // extern "C"
