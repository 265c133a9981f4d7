// Position-parallel fp64 genotype-likelihood kernel.
//
// Replaces prepare_variants (live_variant_caller.py:120-231) + utils.genotype_likelihood
// (utils.py:16-24) up to, but not including, the log10/round/format of the few emitted records
// (done on the host with the host libm).
//
// The reference multiplies per-read error probabilities sequentially in fp64 and lets the product
// underflow to exactly 0.0.  Here each allele's products are evaluated from the integer quality
// histogram as  prod_q e[q]^n  by square-and-multiply in an extended-range representation
// (mantissa in [0.5,1) + 64-bit binary exponent), so nothing underflows early; the final value is
// rounded to fp64 ONCE (ldexp), giving 0.0 exactly where the true product is below 2^-1075 and a
// correctly rounded denormal inside the denormal band (where the reference itself is order
// dependent).  e[q] / 1-e[q] come from the host (math.pow) -- never pow() on the device.
#pragma once
#include "lvc_common.cuh"

namespace lvc {

struct XF {            // value = m * 2^x ; m == 0 means exactly zero
    double m;
    long long x;
};

__device__ __forceinline__ XF xf_one() { return XF{0.5, 1}; }

__device__ __forceinline__ XF xf_from_double(double v) {
    if (v == 0.0) return XF{0.0, 0};
    int e;
    double m = frexp(v, &e);
    return XF{m, (long long)e};
}

__device__ __forceinline__ XF xf_mul(XF a, XF b) {
    XF r;
    r.m = a.m * b.m;            // in [0.25, 1) or 0
    r.x = a.x + b.x;
    if (r.m < 0.5 && r.m != 0.0) { r.m *= 2.0; r.x -= 1; }
    return r;
}

__device__ __forceinline__ XF xf_div(XF a, XF b) {   // b.m != 0
    XF r;
    r.m = a.m / b.m;            // in (0.5, 2)
    r.x = a.x - b.x;
    if (r.m >= 1.0) { r.m *= 0.5; r.x += 1; }
    return r;
}

__device__ __forceinline__ XF xf_pow(XF b, uint32_t n) {
    XF r = xf_one();
    while (n) {
        if (n & 1u) r = xf_mul(r, b);
        n >>= 1;
        if (n) b = xf_mul(b, b);
    }
    return r;
}

__device__ __forceinline__ double xf_to_double(XF a) {
    if (a.m == 0.0) return 0.0;
    if (a.x < -1200) return 0.0;
    if (a.x > 1100) return a.m * 8.98846567431158e307 * 4.0;   // +inf (cannot happen: all factors <= 1)
    return ldexp(a.m, (int)a.x);                                // single rounding, gradual underflow
}

struct GenoParams {
    int64_t G;
    int64_t p0, p1;          // positions [p0, p1) are genotyped (the whole contig by default)
    int64_t min_total_depth;
    int64_t min_allele_depth;
    double min_ratio;
    uint32_t flags;
    int n_planes;            // all planes, ordered by key (group 0 first)
    int grp_begin[5];        // planes of group g are [grp_begin[g], grp_begin[g+1])
    uint32_t cand_cap;
};

constexpr int kGenoThreads = 128;            // 32 positions x 4 allele slots
constexpr int kGenoBatch = 8;                // plane counts requested together
constexpr int kGenoPowBits = 32;

// Power tables, built once per (plane set, phred table) and cached on the device: for every plane the 32
// repeated squarings e^(2^k) and (1-e)^(2^k) in extended range.  Layout: [plane][e | 1-e][32].
__global__ void k_pow_tables(int n_planes, const uint16_t* __restrict__ plane_keys, const double* __restrict__ e_lut,
                             const double* __restrict__ om_lut, XF* __restrict__ pow_tab, double* __restrict__ ed) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n_planes) return;
    const int k = t >> 1;
    const bool is_om = t & 1;
    const uint32_t q = plane_keys[k] & 255u;
    XF v = xf_from_double(is_om ? om_lut[q] : e_lut[q]);
    XF* dst = pow_tab + (size_t)t * kGenoPowBits;
    for (int i = 0; i < kGenoPowBits; ++i) { dst[i] = v; v = xf_mul(v, v); }
    if (!is_om) ed[k] = e_lut[q];
}

struct AlleleStat {
    XF pe;       // prod e
    XF p1;       // prod (1-e)
    double es;   // sum e
    uint32_t ad;
};

__device__ __forceinline__ XF xf_shfl_xor(XF v, int lanemask) {
    XF r;
    r.m = __shfl_xor_sync(0xFFFFFFFFu, v.m, lanemask);
    r.x = __shfl_xor_sync(0xFFFFFFFFu, v.x, lanemask);
    return r;
}

// x^n from the table of x^(2^k): one multiply per set bit of n.  The table entries are normalised
// (mantissa in [0.5, 1)), so a chain of <= 32 products stays far inside the normal fp64 range: the mantissas are
// multiplied as they are and the product is re-normalised ONCE (scaling by a power of two is exact, so the bits
// equal those of a chain normalised after every step).
__device__ __forceinline__ XF xf_pow_tab(const XF* __restrict__ tab, uint32_t n) {
    double m = 1.0;
    long long x = 0;
    while (n) {
        const int k = __ffs(n) - 1;
        n &= n - 1;
        const XF t = tab[k];
        m *= t.m;
        x += t.x;
    }
    int e;
    m = frexp(m, &e);
    return XF{m, x + e};
}

// One thread per (position, allele slot); the 4 slots of a position sit in 4 adjacent lanes and are
// combined with shuffles.  Alleles of the rare groups 1..3 are handled by the same lanes in turn.
__global__ void __launch_bounds__(kGenoThreads) k_genotype(GenoParams gp, const uint32_t* const* __restrict__ plane_ptrs,
                                                           const XF* __restrict__ pow_tab,
                                                           const double* __restrict__ ed_tab,
                                                           const uint32_t* __restrict__ dels,
                                                           const uint8_t* __restrict__ ref,
                                                           const uint32_t* const* __restrict__ first,
                                                           uint32_t* __restrict__ out_depth, uint32_t* __restrict__ out_ad,
                                                           double* __restrict__ out_lik, lvc_candidate* __restrict__ cand,
                                                           uint32_t* __restrict__ cand_count,
                                                           uint32_t* __restrict__ cand_count_next) {
    const int tid = threadIdx.x;
    // everything below reads tables written by the deposit kernel launched before this one
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");        // the next deposit kernel may start reading its batch
    if (blockIdx.x == 0 && tid == 0) *cand_count_next = 0;
    const int slot = tid & 3;
    const int64_t p = gp.p0 + (int64_t)blockIdx.x * (kGenoThreads / 4) + (tid >> 2);
    const bool live = p < gp.p1;
    const int64_t pc = live ? p : gp.p1 - 1;         // clamp: every lane takes part in the shuffles

    AlleleStat st[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) { st[g].pe = xf_one(); st[g].p1 = xf_one(); st[g].es = 0.0; st[g].ad = 0; }

#pragma unroll
    for (int g = 0; g < 4; ++g) {
        // counts of up to kGenoBatch planes are requested together (one memory latency per batch, not per plane)
        for (int k0 = gp.grp_begin[g]; k0 < gp.grp_begin[g + 1]; k0 += kGenoBatch) {
            uint32_t cnt[kGenoBatch];
#pragma unroll
            for (int j = 0; j < kGenoBatch; ++j)
                cnt[j] = k0 + j < gp.grp_begin[g + 1] ? __ldcg(&plane_ptrs[k0 + j][pc * 4 + slot]) : 0u;
#pragma unroll
            for (int j = 0; j < kGenoBatch; ++j) {
                const uint32_t n = cnt[j];
                if (n) {
                    const int k = k0 + j;
                    st[g].ad += n;
                    st[g].es += (double)n * ed_tab[k];
                    st[g].pe = xf_mul(st[g].pe, xf_pow_tab(pow_tab + (size_t)(2 * k) * kGenoPowBits, n));
                    st[g].p1 = xf_mul(st[g].p1, xf_pow_tab(pow_tab + (size_t)(2 * k + 1) * kGenoPowBits, n));
                }
            }
        }
    }
    // ---- combine the (up to 16) alleles of the position
    const bool has_other = gp.grp_begin[4] > gp.grp_begin[1];
    XF own = st[0].ad ? st[0].pe : xf_one();                  // product of e over this lane's alleles
    uint32_t own_ad = st[0].ad;
    if (has_other) {
#pragma unroll
        for (int g = 1; g < 4; ++g) { if (st[g].ad) own = xf_mul(own, st[g].pe); own_ad += st[g].ad; }
    }
    const XF v1 = xf_shfl_xor(own, 1);
    const XF pair = xf_mul(own, v1);
    const XF opp = xf_shfl_xor(pair, 2);
    const XF others_lanes = xf_mul(v1, opp);                  // product over the other three lanes
    uint32_t dsum = own_ad + __shfl_xor_sync(0xFFFFFFFFu, own_ad, 1);
    dsum += __shfl_xor_sync(0xFFFFFFFFu, dsum, 2);
    const uint64_t depth = (uint64_t)dels[pc] + dsum;
    // L(a) = prod(1-e | a) * prod over b != a of prod(e | b)      (utils.py:16-24)
    double L[4];
    double Ssum = 0.0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        L[g] = 0.0;
        if ((g == 0 || has_other) && st[g].ad) {
            XF others = others_lanes;
            if (has_other) {
#pragma unroll
                for (int h = 0; h < 4; ++h)
                    if (h != g && st[h].ad) others = xf_mul(others, st[h].pe);
            }
            L[g] = xf_to_double(xf_mul(st[g].p1, others));
            Ssum += L[g];
        }
    }
    Ssum += __shfl_xor_sync(0xFFFFFFFFu, Ssum, 1);
    Ssum += __shfl_xor_sync(0xFFFFFFFFu, Ssum, 2);
    const double S = Ssum == 0.0 ? 1.0 : Ssum;                                   // live_variant_caller.py:146
    if (!live) return;
    const uint32_t depth32 = (uint32_t)(depth > 0xFFFFFFFFull ? 0xFFFFFFFFull : depth);
    if (slot == 0) out_depth[p] = depth32;
    out_ad[p * 4 + slot] = st[0].ad;
    out_lik[p * 4 + slot] = L[0];

    if ((int64_t)depth < gp.min_total_depth) return;                 // :131
    if (depth == 0) return;
    const uint8_t rb = ref[p];
    const bool all = gp.flags & 1u;
    const double ddepth = (double)depth;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        if (!((g == 0 || has_other) && st[g].ad)) continue;
        const uint32_t gs = (uint32_t)(g * 4 + slot);
        const uint32_t nib = gs_nibble(gs);
        const char letter = "=ACMGRSVTWYHKDBN"[nib];
        const bool ok = all || ((uint8_t)letter != rb && (int64_t)st[g].ad >= gp.min_allele_depth &&
                                ((double)st[g].ad / ddepth) >= gp.min_ratio);     // :151-155
        if (!ok) continue;
        const uint32_t idx = atomicAdd(cand_count, 1u);
        if (idx >= gp.cand_cap) continue;
        lvc_candidate c;
        c.pos = (int32_t)p; c.code = (uint8_t)nib; c.ref = rb; c.pad0 = 0;
        c.ad = st[g].ad; c.dp = depth32;
        const uint32_t* f = first[g];
        c.first = f ? f[p * 4 + slot] : kUnsetOrdinal;
        c.pad1 = 0; c.L = L[g]; c.S = S; c.esum = st[g].es;
        cand[idx] = c;
    }
}

}  // namespace lvc
