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
    int64_t min_total_depth;
    int64_t min_allele_depth;
    double min_ratio;
    uint32_t flags;
    int n_planes;            // all planes, ordered so that the n_g0 group-0 planes come first
    int n_g0;
    uint32_t cand_cap;
};

struct AlleleStat {
    XF pe;       // prod e
    XF p1;       // prod (1-e)
    double es;   // sum e
    uint32_t ad;
};

__device__ __forceinline__ void stat_init(AlleleStat& s) {
    s.pe = xf_one(); s.p1 = xf_one(); s.es = 0.0; s.ad = 0;
}
__device__ __forceinline__ void stat_add(AlleleStat& s, uint32_t n, XF e, XF om, double ed) {
    s.ad += n;
    s.pe = xf_mul(s.pe, xf_pow(e, n));
    s.p1 = xf_mul(s.p1, xf_pow(om, n));
    s.es += (double)n * ed;
}

// plane_ptrs[k], plane_keys[k] for k < n_planes (ordered: group 0 first); e_lut / om_lut [256].
__global__ void __launch_bounds__(128) k_genotype(GenoParams gp, const uint32_t* const* __restrict__ plane_ptrs,
                                                  const uint16_t* __restrict__ plane_keys,
                                                  const double* __restrict__ e_lut, const double* __restrict__ om_lut,
                                                  const uint32_t* __restrict__ dels, const uint8_t* __restrict__ ref,
                                                  const uint32_t* const* __restrict__ first,   // [4] device array of ptrs
                                                  uint32_t* __restrict__ out_depth, uint32_t* __restrict__ out_ad,
                                                  double* __restrict__ out_lik, lvc_candidate* __restrict__ cand,
                                                  uint32_t* __restrict__ cand_count) {
    extern __shared__ unsigned char smem_raw[];
    // per-plane constants staged once per CTA: e, 1-e as XF and e as double
    XF* s_e = reinterpret_cast<XF*>(smem_raw);
    XF* s_om = s_e + gp.n_planes;
    double* s_ed = reinterpret_cast<double*>(s_om + gp.n_planes);
    for (int k = threadIdx.x; k < gp.n_planes; k += blockDim.x) {
        const uint32_t q = plane_keys[k] & 255u;
        s_e[k] = xf_from_double(e_lut[q]);
        s_om[k] = xf_from_double(om_lut[q]);
        s_ed[k] = e_lut[q];
    }
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= gp.G) return;

    AlleleStat st[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) stat_init(st[s]);
    uint64_t depth = dels[p];
    // ---- group 0 (A,C,G,T): register resident
    for (int k = 0; k < gp.n_g0; ++k) {
        const uint4 c = *reinterpret_cast<const uint4*>(plane_ptrs[k] + p * 4);
        if ((c.x | c.y | c.z | c.w) == 0u) continue;
        const XF e = s_e[k], om = s_om[k];
        const double ed = s_ed[k];
        if (c.x) stat_add(st[0], c.x, e, om, ed);
        if (c.y) stat_add(st[1], c.y, e, om, ed);
        if (c.z) stat_add(st[2], c.z, e, om, ed);
        if (c.w) stat_add(st[3], c.w, e, om, ed);
    }
    // ---- groups 1..3 (the other 12 nibble codes): rare, local-memory resident
    AlleleStat ot[12];
    bool any_other = false;
    if (gp.n_planes > gp.n_g0) {
        for (int s = 0; s < 12; ++s) stat_init(ot[s]);
        for (int k = gp.n_g0; k < gp.n_planes; ++k) {
            const uint4 c = *reinterpret_cast<const uint4*>(plane_ptrs[k] + p * 4);
            if ((c.x | c.y | c.z | c.w) == 0u) continue;
            any_other = true;
            const int g = (plane_keys[k] >> 8) - 1;
            const XF e = s_e[k], om = s_om[k];
            const double ed = s_ed[k];
            if (c.x) stat_add(ot[g * 4 + 0], c.x, e, om, ed);
            if (c.y) stat_add(ot[g * 4 + 1], c.y, e, om, ed);
            if (c.z) stat_add(ot[g * 4 + 2], c.z, e, om, ed);
            if (c.w) stat_add(ot[g * 4 + 3], c.w, e, om, ed);
        }
    }
    // total product of e over every allele, total depth
    XF tot = xf_one();
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (st[s].ad) { tot = xf_mul(tot, st[s].pe); depth += st[s].ad; }
    }
    if (any_other) {
        for (int s = 0; s < 12; ++s)
            if (ot[s].ad) { tot = xf_mul(tot, ot[s].pe); depth += ot[s].ad; }
    }
    // L(a) = prod(1-e | a) * prod over b != a of prod(e | b)      (utils.py:16-24)
    double L[4];
    double S = 0.0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        L[s] = 0.0;
        if (st[s].ad) {
            L[s] = xf_to_double(xf_mul(st[s].p1, xf_div(tot, st[s].pe)));
            S += L[s];
        }
    }
    double Lo[12];
    if (any_other) {
        for (int s = 0; s < 12; ++s) {
            Lo[s] = 0.0;
            if (ot[s].ad) {
                Lo[s] = xf_to_double(xf_mul(ot[s].p1, xf_div(tot, ot[s].pe)));
                S += Lo[s];
            }
        }
    }
    if (S == 0.0) S = 1.0;                                           // live_variant_caller.py:146
    const uint32_t depth32 = (uint32_t)(depth > 0xFFFFFFFFull ? 0xFFFFFFFFull : depth);
    out_depth[p] = depth32;
    *reinterpret_cast<uint4*>(out_ad + p * 4) = make_uint4(st[0].ad, st[1].ad, st[2].ad, st[3].ad);
    *reinterpret_cast<double2*>(out_lik + p * 4) = make_double2(L[0], L[1]);
    *reinterpret_cast<double2*>(out_lik + p * 4 + 2) = make_double2(L[2], L[3]);

    if ((int64_t)depth < gp.min_total_depth) return;                 // :131
    if (depth == 0) return;
    const uint8_t rb = ref[p];
    const bool all = gp.flags & 1u;
    const double ddepth = (double)depth;
    auto emit = [&](uint32_t gs, const AlleleStat& a, double Lval) {
        const uint32_t nib = gs_nibble(gs);
        const char letter = "=ACMGRSVTWYHKDBN"[nib];
        bool ok = all || ((uint8_t)letter != rb && (int64_t)a.ad >= gp.min_allele_depth &&
                          ((double)a.ad / ddepth) >= gp.min_ratio);              // :151-155
        if (!ok) return;
        const uint32_t slot = atomicAdd(cand_count, 1u);
        if (slot >= gp.cand_cap) return;
        lvc_candidate c;
        c.pos = (int32_t)p; c.code = (uint8_t)nib; c.ref = rb; c.pad0 = 0;
        c.ad = a.ad; c.dp = depth32;
        const uint32_t* f = first[gs >> 2];
        c.first = f ? f[p * 4 + (gs & 3u)] : kUnsetOrdinal;
        c.pad1 = 0; c.L = Lval; c.S = S; c.esum = a.es;
        cand[slot] = c;
    };
#pragma unroll
    for (int s = 0; s < 4; ++s)
        if (st[s].ad) emit((uint32_t)s, st[s], L[s]);
    if (any_other)
        for (int s = 0; s < 12; ++s)
            if (ot[s].ad) emit((uint32_t)(4 + s), ot[s], Lo[s]);
}

}  // namespace lvc
