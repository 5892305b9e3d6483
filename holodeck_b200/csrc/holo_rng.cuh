// holo_rng.cuh -- counter-based Philox4x32-10 and exact Poisson / normal sampling.
//
// The reference draws with numpy's C `random_poisson` / `random_normal` on an unseeded PCG64
// (cyutils.pyx:29-31, 875, 1315, ...), one sequential stream per call.  A sequential stream cannot
// be partitioned over 10^5 threads, so the B200 path keys a counter-based generator on
// (seed; flat (cell,f) index, global realization, stream id, retry): any partition of cells or
// realizations over threads / launches / GPUs gives the same numbers.  Parity with the reference
// is therefore statistical for realised quantities and bit-exact in supplied-count mode.
//
// Samplers (all exact, fp64):
//   lam <  10      : inversion by sequential search on the survival function, 64-bit uniform so
//                    P(n>=1) = -expm1(-lam) is resolved down to lam ~ 5e-20
//   lam >= 10      : PTRS transformed rejection (Hormann 1993), the algorithm numpy uses
//                    (numpy/random/src/distributions/distributions.c: random_poisson_ptrs)
//   lam >  thresh  : Normal(lam, sqrt(lam)) by Box-Muller, NOT floored -- as cyutils.pyx:890-891,
//                    1329-1330 (gravwaves.poisson_as_needed floors it; that wrapper floors on top)
#pragma once

#include "holo_common.cuh"

namespace holo {

HOLO_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

struct Philox4 {
    uint32_t v[4];
};

// Philox4x32-10 (Salmon et al. 2011, Random123)
HOLO_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n1 = lo1;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        uint32_t n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 r;
    r.v[0] = c0; r.v[1] = c1; r.v[2] = c2; r.v[3] = c3;
    return r;
}

// Identifies one random draw: which (cell,f) element, which realization, which consumer.
struct DrawKey {
    uint32_t k0, k1;     // seed
    uint32_t idx_lo, idx_hi;   // flat element index
    uint32_t real;       // global realization index
    uint32_t stream;     // consumer id (so that e.g. `gwb` and `hc_bg` draw independently)
};

HOLO_HD Philox4 draw_bits(const DrawKey& k, uint32_t trial) {
    return philox4x32_10(k.idx_lo, k.idx_hi, k.real, (k.stream << 20) | trial, k.k0, k.k1);
}

HOLO_HD double u53(uint32_t hi, uint32_t lo) {   // numpy next_double: (x >> 11) * 2^-53
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

// ---- per-element sampler set-up (computed once per staged element, shared by all realizations) ----
enum { CLS_SMALL = 0, CLS_PTRS = 1, CLS_NORMAL = 2 };

struct DrawPrep {
    double lam;
    double a0, a1, a2, a3;
    int cls;
};

HOLO_HD uint64_t prob_to_u64(double p) {   // floor(p * 2^64) for p in [0,1)
    double s = p * 18446744073709551616.0;
    if (s >= 18446744073709549568.0) return 0xFFFFFFFFFFFFF800ull;
#if defined(__CUDA_ARCH__)
    return __double2ull_rd(s);
#else
    return (uint64_t)s;
#endif
}

HOLO_HD DrawPrep prep_draw(double lam, double thresh) {
    DrawPrep p;
    p.lam = lam;
    p.a0 = p.a1 = p.a2 = p.a3 = 0.0;
    if (lam > thresh) {
        p.cls = CLS_NORMAL;
        p.a0 = sqrt(lam);
    } else if (lam >= 10.0) {
        p.cls = CLS_PTRS;
        double slam = sqrt(lam);
        double b = 0.931 + 2.53 * slam;
        p.a0 = b;
        p.a1 = 1.1239 + 1.1328 / (b - 3.4);   // invalpha
        p.a2 = 0.9277 - 3.6224 / (b - 2.0);   // vr
        p.a3 = log(lam);
    } else {
        p.cls = CLS_SMALL;
        double T = -expm1(-lam);              // P(n >= 1)
        uint64_t t = prob_to_u64(T);
#if defined(__CUDA_ARCH__)
        p.a0 = __longlong_as_double((long long)t);
#else
        union { uint64_t u; double d; } cv; cv.u = t; p.a0 = cv.d;
#endif
        p.a1 = T;
        p.a2 = lam * exp(-lam);               // P(n == 1)
    }
    return p;
}

HOLO_HD double draw_small(const DrawPrep& p, const Philox4& b) {
#if defined(__CUDA_ARCH__)
    uint64_t t = (uint64_t)__double_as_longlong(p.a0);
#else
    union { uint64_t u; double d; } cv; cv.d = p.a0; uint64_t t = cv.u;
#endif
    uint64_t u = ((uint64_t)b.v[0] << 32) | b.v[1];
    if (u >= t) return 0.0;
    double v = (double)u * (1.0 / 18446744073709551616.0);
    double n = 1.0;
    double pm = p.a2;            // pmf(n)
    double S = p.a1 - pm;        // P(N >= n+1)
    while (v < S && n < 1000.0) {
        n += 1.0;
        pm *= p.lam / n;
        S -= pm;
    }
    return n;
}

// numpy random_poisson_ptrs, one trial per Philox block (U from words 0,1; V from words 2,3)
HOLO_HD double draw_ptrs(const DrawPrep& p, const DrawKey& key, Philox4 bits) {
    double lam = p.lam, b = p.a0, invalpha = p.a1, vr = p.a2, loglam = p.a3;
    double a = -0.059 + 0.02483 * b;
    uint32_t trial = 0;
    while (true) {
        double U = u53(bits.v[0], bits.v[1]) - 0.5;
        double V = u53(bits.v[2], bits.v[3]);
        double us = 0.5 - fabs(U);
        double k = floor((2.0 * a / us + b) * U + lam + 0.43);
        if ((us >= 0.07) && (V <= vr)) return k;
        bool retry = (k < 0.0) || ((us < 0.013) && (V > us));
        if (!retry) {
            if ((log(V) + log(invalpha) - log(a / (us * us) + b)) <= (-lam + k * loglam - lgamma(k + 1.0)))
                return k;
        }
        ++trial;
        if (trial >= 1000u) return floor(lam);   // unreachable in practice (acceptance ~ 0.9 per trial)
        bits = draw_bits(key, trial);
    }
}

HOLO_HD double draw_normal(const DrawPrep& p, const Philox4& b) {
    double u1 = 1.0 - u53(b.v[0], b.v[1]);   // (0, 1]
    double u2 = u53(b.v[2], b.v[3]);
    double z = sqrt(-2.0 * log(u1)) * cos(2.0 * CY_PI * u2);
    return p.lam + p.a0 * z;
}

HOLO_HD double draw_count(const DrawPrep& p, const DrawKey& key) {
    Philox4 bits = draw_bits(key, 0);
    if (p.cls == CLS_SMALL) return draw_small(p, bits);
    if (p.cls == CLS_PTRS) return draw_ptrs(p, key, bits);
    return draw_normal(p, bits);
}

}  // namespace holo
