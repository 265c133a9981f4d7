/*
 * CPU ORACLE in C (test infrastructure, NOT product code).
 *
 * A plain, single-threaded restatement of the reference hot path for inputs too large for the
 * pure-Python oracle (oracle/pileup_oracle.py).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.
 *
 * It follows (paths relative to /root/reference):
 *   - pysam/htslib read filter + bam_plp_push max_depth admission          ([EXT])
 *   - htslib resolve_cigar2 + pysam pileup_base_qual_skip                  ([EXT])
 *   - variant_caller/live_variant_caller.py:74-103  process_pileup_column / process_svn
 *   - variant_caller/utils.py:9-24                   from_phred_scale / genotype_likelihood
 *   - variant_caller/live_variant_caller.py:131-157  gates of prepare_variants
 *
 * The reference keeps per-allele Python lists and multiplies them left to right with np.prod.  The
 * product is a streaming statistic, so this oracle keeps, per (position, allele), the running
 * products in READ ORDER -- bit-identical to the reference's sequential products, including their
 * underflow to 0.0 -- without materialising the lists.
 *
 * Parity status: pinned against oracle/pileup_oracle.py (which is pinned to the real reference by
 * tests/golden/) in tests/test_oracle_c.py.  The htslib half is "parity unpinned" (see DESIGN.md).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FLAG_FILTER (0x4u | 0x100u | 0x200u | 0x400u)

typedef struct {
    int64_t G;
    uint32_t* depth;   /* [G]      totalDepth                                             */
    uint32_t* cov;     /* [G]      number of admitted reads covering the column (site exists iff > 0) */
    uint32_t* ad;      /* [G*16]   len(snvs[allele]) by BAM nibble code                   */
    uint64_t* qsum;    /* [G*16]   sum of qualities                                       */
    uint64_t* q2sum;   /* [G*16]   sum of squared qualities (checksum)                    */
    uint32_t* first;   /* [G*16]   ordinal of the first read depositing the allele        */
    double* pe;        /* [G*16]   running prod e      (append order)                     */
    double* p1;        /* [G*16]   running prod (1-e)  (append order)                     */
    double* esum;      /* [G*16]   running sum e                                          */
    uint32_t* hist;    /* [G*16*256] or NULL                                              */
    uint64_t ordinal;  /* reads consumed so far                                           */
    double e_lut[256]; /* math.pow(10, q/-10) supplied by the caller (host libm)          */
} orc_state;

static int ref_op(uint32_t op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }
static int qry_op(uint32_t op) { return op == 0 || op == 1 || op == 4 || op == 7 || op == 8; }
static int match_op(uint32_t op) { return op == 0 || op == 7 || op == 8; }

/* SURVEY B2 + B4.  keep[i] = 1 iff the read enters the pileup buffer. Returns -4 if unsorted. */
int orc_admit(uint32_t n, const int32_t* pos, const uint16_t* flag, const uint8_t* mapq, const uint32_t* cigar_off,
              const uint32_t* cigar, int min_mq, int max_depth, int64_t G, uint8_t* keep) {
    /* ends[e] = number of buffered reads ending at e; positions bounded by G + longest read */
    int64_t cap = G + 1;
    for (uint32_t i = 0; i < n; ++i) {
        int64_t rl = 0;
        for (uint32_t k = cigar_off[i]; k < cigar_off[i + 1]; ++k)
            if (ref_op(cigar[k] & 15u)) rl += cigar[k] >> 4;
        if (pos[i] + rl + 1 > cap) cap = pos[i] + rl + 1;
    }
    uint32_t* ends = (uint32_t*)calloc((size_t)cap + 1, sizeof(uint32_t));
    if (!ends) return -3;
    int64_t iter_pos = 0, max_pos = -1, nbuf = 0;
    for (uint32_t i = 0; i < n; ++i) {
        keep[i] = 0;
        uint32_t f = flag[i];
        if (f & FLAG_FILTER) continue;
        if ((int)mapq[i] < min_mq) continue;
        if ((f & 1u) && !(f & 2u)) continue;
        int64_t rl = 0;
        for (uint32_t k = cigar_off[i]; k < cigar_off[i + 1]; ++k)
            if (ref_op(cigar[k] & 15u)) rl += cigar[k] >> 4;
        if (rl == 0) continue;
        int64_t p = pos[i];
        if (p < max_pos) { free(ends); return -4; }
        if (p == iter_pos && nbuf + 1 > max_depth) continue;
        max_pos = p;
        keep[i] = 1;
        nbuf++;
        ends[p + rl]++;
        while (max_pos > iter_pos) {
            nbuf -= ends[iter_pos];
            ends[iter_pos] = 0;
            if (nbuf - 1 == 0) iter_pos = max_pos; else iter_pos++;
        }
    }
    free(ends);
    return 0;
}

static void deposit(orc_state* st, int64_t r, uint32_t nib, uint32_t q, uint32_t ord) {
    int64_t c = r * 16 + nib;
    double e = st->e_lut[q];
    if (st->ad[c] == 0) { st->pe[c] = e; st->p1[c] = 1.0 - e; st->esum[c] = e; st->first[c] = ord; }
    else { st->pe[c] *= e; st->p1[c] *= (1.0 - e); st->esum[c] += e; }
    st->ad[c]++;
    st->qsum[c] += q;
    st->q2sum[c] += (uint64_t)q * q;
    st->depth[r]++;
    if (st->hist) st->hist[c * 256 + q]++;
}

/* one batch == one process_bam call (the max_depth rule applies per call). keep may be NULL (computed). */
int orc_process(orc_state* st, uint32_t n, const int32_t* pos, const uint16_t* flag, const uint8_t* mapq,
                const uint8_t* keep_in, const uint32_t* cigar_off, const uint32_t* cigar, const uint64_t* seq_off,
                const uint8_t* seq4, const uint8_t* qual, int min_bq, int min_mq, int max_depth) {
    uint8_t* keep = (uint8_t*)malloc(n ? n : 1);
    if (!keep) return -3;
    int rc = orc_admit(n, pos, flag, mapq, cigar_off, cigar, min_mq, max_depth, st->G, keep);
    if (rc) { free(keep); return rc; }
    (void)keep_in;
    for (uint32_t i = 0; i < n; ++i) {
        if (!keep[i]) continue;
        int64_t rl = 0;
        uint32_t lq = 0;
        for (uint32_t k = cigar_off[i]; k < cigar_off[i + 1]; ++k) {
            uint32_t op = cigar[k] & 15u, len = cigar[k] >> 4;
            if (ref_op(op)) rl += len;
            if (qry_op(op)) lq += len;
        }
        int64_t r = pos[i];
        if (r < 0 || r + rl > st->G) { free(keep); return -5; }
        for (int64_t c = r; c < r + rl; ++c) st->cov[c]++;
        const uint8_t* q = qual + seq_off[i];
        const uint8_t* s = seq4 + (seq_off[i] >> 1);
        uint32_t qi = 0, ord = (uint32_t)(st->ordinal + i);
        for (uint32_t k = cigar_off[i]; k < cigar_off[i + 1]; ++k) {
            uint32_t op = cigar[k] & 15u, len = cigar[k] >> 4;
            if (match_op(op)) {
                for (uint32_t j = 0; j < len; ++j, ++qi, ++r) {
                    if ((int)q[qi] < min_bq) continue;
                    uint32_t nib = (qi & 1u) ? (s[qi >> 1] & 15u) : (s[qi >> 1] >> 4);
                    deposit(st, r, nib, q[qi], ord);
                }
            } else if (op == 2 || op == 3) {
                uint32_t dq = qi < lq ? q[qi] : 0;
                if ((int)dq >= min_bq)
                    for (uint32_t j = 0; j < len; ++j) st->depth[r + j]++;
                r += len;
            } else if (op == 1 || op == 4) {
                qi += len;
            }
        }
    }
    st->ordinal += n;
    free(keep);
    return 0;
}

/* utils.genotype_likelihood for every allele of every site + the emission gates of prepare_variants.
 * L[G*16], S[G]; emit[G*16] = 1 where a record would be written.  Returns the number of records. */
int64_t orc_genotype(const orc_state* st, const uint8_t* ref, int64_t min_dp, int64_t min_ad, double ratio, double* L,
                     double* S, uint8_t* emit) {
    static const char letters[] = "=ACMGRSVTWYHKDBN";
    int64_t n_emit = 0;
    for (int64_t p = 0; p < st->G; ++p) {
        int order[16], na = 0;
        for (int c = 0; c < 16; ++c) { L[p * 16 + c] = 0.0; emit[p * 16 + c] = 0; if (st->ad[p * 16 + c]) order[na++] = c; }
        S[p] = 1.0;
        if (na == 0) continue;
        /* dict order == first-seen order */
        for (int a = 1; a < na; ++a) {
            int v = order[a], b = a - 1;
            while (b >= 0 && st->first[p * 16 + order[b]] > st->first[p * 16 + v]) { order[b + 1] = order[b]; --b; }
            order[b + 1] = v;
        }
        double sum = 0.0;
        for (int a = 0; a < na; ++a) {
            double non = 1.0;
            for (int b = 0; b < na; ++b)
                if (b != a) non = non * st->pe[p * 16 + order[b]];
            double l = st->p1[p * 16 + order[a]] * non;
            L[p * 16 + order[a]] = l;
            sum = sum + l;
        }
        S[p] = sum != 0.0 ? sum : 1.0;
        if ((int64_t)st->depth[p] < min_dp) continue;
        for (int a = 0; a < na; ++a) {
            int c = order[a];
            uint32_t ad = st->ad[p * 16 + c];
            if ((uint8_t)letters[c] != ref[p] && (int64_t)ad >= min_ad && (double)ad / (double)st->depth[p] >= ratio) {
                emit[p * 16 + c] = 1;
                n_emit++;
            }
        }
    }
    return n_emit;
}

int orc_sizeof_state(void) { return (int)sizeof(orc_state); }
