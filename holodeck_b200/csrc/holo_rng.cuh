// holo_rng.cuh -- counter-based Philox4x32-10 and exact Poisson / normal sampling.
//
// The reference draws with numpy's C `random_poisson` / `random_normal` on an unseeded PCG64
// (cyutils.pyx:29-31, 875, 1315, ...), one sequential stream per call.  A sequential stream cannot
// be partitioned over 10^5 threads, so the B200 path keys a counter-based generator on
// (seed; grid element, global realization, purpose, retry): any partition of cells or realizations
// over threads / launches / GPUs gives the same numbers.  Parity with the reference is therefore
// statistical for realised quantities and bit-exact in supplied-count mode.
//
// Samplers (all exact in distribution, fp64):
//   lam < 2^-8   TINY  : P(n>=1) = -expm1(-lam) <= 0.4%.  One 32-bit word decides "n = 0" for all
//                        but a fraction ~lam of the draws (four draws share one Philox block); the
//                        rest extend the word to 64 bits and invert the survival function.
//   lam < 32     SMALL : inversion by sequential search on the survival function, 64-bit uniform
//                        (all lanes of a warp share lam, so the walk lengths are similar: lock-step friendly).
//   lam >= 32    PTRS  : transformed rejection (Hormann 1993; the algorithm numpy uses, numpy/random/
//                        src/distributions/distributions.c:random_poisson_ptrs).  The exact
//                        acceptance test is evaluated without lgamma: Stirling's series for ln k!
//                        and lam*[x-(1+x)ln(1+x)], x=(k-lam)/lam, for the log-pmf (one log1p + one log).
//   lam > thresh NORMAL: Normal(lam, sqrt(lam)) by Box-Muller, NOT floored -- as cyutils.pyx:890-891,
//                        1329-1330 (gravwaves.poisson_as_needed floors it; that wrapper floors on top).
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

// What a Philox block is used for: the purpose is folded into the counter so no two uses collide.
enum {
    PURPOSE_GROUP_HI = 1,   // one block per (cell, frequency group, realization): word j -> frequency j
    PURPOSE_GROUP_LO = 2,   // second block of the same group: low halves of the 64-bit uniforms
    PURPOSE_ELEMENT = 3,    // one block per (cell, f, realization, trial): PTRS / normal / TINY refinement
};

struct DrawKey {
    uint32_t k0, k1;     // seed
    uint32_t real;       // global realization index
    uint32_t stream;     // consumer id (so that e.g. `gwb` and `hc_bg` draw independently), < 256
};

// block shared by the (up to 4) frequencies of group `fg` of cell `cell`
HOLO_HD Philox4 group_bits(const DrawKey& k, uint32_t cell, uint32_t fg, int purpose) {
    return philox4x32_10(cell, fg | ((uint32_t)purpose << 28), k.real, k.stream << 24, k.k0, k.k1);
}

// block private to grid element `idx` (= cell*F + f)
HOLO_HD Philox4 element_bits(const DrawKey& k, uint64_t idx, uint32_t trial) {
    return philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32) | ((uint32_t)PURPOSE_ELEMENT << 28), k.real,
                         (k.stream << 24) | trial, k.k0, k.k1);
}

HOLO_HD double u53(uint32_t hi, uint32_t lo) {   // numpy next_double: (x >> 11) * 2^-53
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

HOLO_HD double bits_as_double(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    union { uint64_t u; double d; } cv; cv.u = u; return cv.d;
#endif
}

HOLO_HD uint64_t double_as_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    union { uint64_t u; double d; } cv; cv.d = d; return cv.u;
#endif
}

HOLO_HD uint64_t prob_to_u64(double p) {   // floor(p * 2^64) for p in [0,1)
    double s = p * 18446744073709551616.0;
    if (s >= 18446744073709549568.0) return 0xFFFFFFFFFFFFF800ull;
#if defined(__CUDA_ARCH__)
    return __double2ull_rd(s);
#else
    return (uint64_t)s;
#endif
}

// ---- per-element sampler set-up (computed once per staged element, shared by all realizations) ----
enum { CLS_EMPTY = 0, CLS_TINY = 1, CLS_SMALL = 2, CLS_PTRS = 3, CLS_NORMAL = 4 };

constexpr double TINY_LAM = 0.00390625;   // 2^-8
constexpr double PTRS_MIN_LAM = 32.0;      // below: inversion from 0 (lock-step friendly), above: PTRS

struct FPrep {
    double lam;
    double h;
    double a0, a1, a2;
};
//   TINY  : a0 = bits(floor(T 2^64)),  T = P(n>=1)
//   SMALL : a0 = bits(floor(T 2^64)),  a1 = T,  a2 = P(n==1)
//   PTRS  : a0 = b,  a1 = vr,  a2 = 1/lam
//   NORMAL: a0 = sqrt(lam)

// class of a draw from its expectation value alone (same thresholds as prep_draw)
HOLO_HD int classify_draw(double lam, double thresh) {
    if (!(lam > 0.0)) return CLS_EMPTY;
    if (lam > thresh) return CLS_NORMAL;
    if (lam >= PTRS_MIN_LAM) return CLS_PTRS;
    return lam < TINY_LAM ? CLS_TINY : CLS_SMALL;
}

HOLO_HD int prep_draw(double lam, double thresh, FPrep& p) {
    p.lam = lam;
    p.a0 = p.a1 = p.a2 = 0.0;
    if (!(lam > 0.0)) return CLS_EMPTY;
    if (lam > thresh) {
        p.a0 = sqrt(lam);
        return CLS_NORMAL;
    }
    if (lam >= PTRS_MIN_LAM) {
        double b = 0.931 + 2.53 * sqrt(lam);
        p.a0 = b;
        p.a1 = 0.9277 - 3.6224 / (b - 2.0);
        p.a2 = 1.0 / lam;
        return CLS_PTRS;
    }
    double T = -expm1(-lam);
    p.a0 = bits_as_double(prob_to_u64(T));
    if (lam < TINY_LAM) return CLS_TINY;
    p.a1 = T;
    p.a2 = lam * exp(-lam);
    return CLS_SMALL;
}

// n >= 1 is already known (u < t): find n by walking the survival function.  v = u * 2^-64.
// `rcp` (optional) is a table of 1/n for n < RCP_TABLE so the walk has no division.
constexpr int RCP_TABLE = 96;

HOLO_HD double invert_survival(double lam, double T, double p1, double v, const double* rcp = nullptr) {
    double n = 1.0;
    int ni = 1;
    double pm = p1;           // pmf(n)
    double S = T - pm;        // P(N >= n+1)
    while (v < S && ni < 1000) {
        ni += 1;
        n += 1.0;
        pm = (rcp != nullptr && ni < RCP_TABLE) ? pm * lam * rcp[ni] : pm * lam / n;
        S -= pm;
    }
    return n;
}

HOLO_HD double draw_small(const FPrep& p, uint32_t hi, uint32_t lo, const double* rcp = nullptr) {
    uint64_t t = double_as_bits(p.a0);
    uint64_t u = ((uint64_t)hi << 32) | lo;
    if (u >= t) return 0.0;
    return invert_survival(p.lam, p.a1, p.a2, (double)u * (1.0 / 18446744073709551616.0), rcp);
}

// `hi` is this draw's word of the shared group block; nearly always it alone proves n = 0.
HOLO_HD double draw_tiny(const FPrep& p, uint32_t hi, const DrawKey& key, uint64_t idx) {
    uint64_t t = double_as_bits(p.a0);
    if (hi > (uint32_t)(t >> 32)) return 0.0;
    Philox4 ext = element_bits(key, idx, 0);
    uint64_t u = ((uint64_t)hi << 32) | ext.v[0];
    if (u >= t) return 0.0;
    double T = -expm1(-p.lam);
    return invert_survival(p.lam, T, p.lam * exp(-p.lam), (double)u * (1.0 / 18446744073709551616.0));
}

// ln(k!) for k < 10 (exact doubles of lgamma(k+1))
HOLO_HD double log_factorial_small(int k) {
    const double tab[10] = {0.0, 0.0, 0.6931471805599453, 1.791759469228055, 3.1780538303479458,
                            4.787491742782046, 6.579251212010101, 8.525161361065415, 10.60460290274525,
                            12.801827480081469};
    return tab[k];
}

// Exact PTRS acceptance test:  log(V*invalpha/(a/us^2+b)) <= -lam + k log(lam) - log(k!)
HOLO_NOINLINE_STATIC bool ptrs_accept(double lam, double inv_lam, double a, double b, double us, double V, double k) {
    double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    double den = a / (us * us) + b;
    if (k >= 10.0) {
        // ln k! = (k+1/2) ln k - k + ln(2 pi)/2 + 1/(12k) - 1/(360k^3) + 1/(1260k^5) - 1/(1680k^7)  (|err| < 1e-12)
        // => log pmf(k) + ln(2 pi k)/2 = lam*[x - (1+x) ln(1+x)] - corr(k),   x = (k - lam)/lam
        double x = (k - lam) * inv_lam;
        double g;
        if (fabs(x) < 0.0625) {
            // x - (1+x)ln(1+x) = -x^2 * sum_m (-x)^m / ((m+1)(m+2)),  truncated at m = 9 (|err| < 1e-13 rel.)
            double s = 1.0 / 110.0;
            s = 1.0 / 90.0 - x * s;
            s = 1.0 / 72.0 - x * s;
            s = 1.0 / 56.0 - x * s;
            s = 1.0 / 42.0 - x * s;
            s = 1.0 / 30.0 - x * s;
            s = 1.0 / 20.0 - x * s;
            s = 1.0 / 12.0 - x * s;
            s = 1.0 / 6.0 - x * s;
            s = 0.5 - x * s;
            g = -(x * x) * s;
        } else {
            g = x - (1.0 + x) * log1p(x);
        }
        double ik = 1.0 / k;
        double ik2 = ik * ik;
        double corr = ik * (1.0 / 12.0 - ik2 * (1.0 / 360.0 - ik2 * (1.0 / 1260.0 - ik2 * (1.0 / 1680.0))));
        double rhs = lam * g - corr;
        double lhs = log(V * invalpha * sqrt(6.283185307179586 * k) / den);
        return lhs <= rhs;
    }
    double lhs = log(V * invalpha / den);
    double rhs = -lam + k * log(lam) - log_factorial_small((int)k);
    return lhs <= rhs;
}

// fp32 pre-screen of the same test.  Returns +1 accept, -1 reject, 0 undecided (|lhs-rhs| within the
// fp32 error band; the caller then runs the fp64 test).  The band (1e-3) is >= 10x the worst-case
// fp32 evaluation error of either side (|lam*g| <= ~100 in the region PTRS proposes, eps = 6e-8).
HOLO_HD int ptrs_screen(double lam, double inv_lam, double a, double b, double us, double V, double k) {
    if (k < 10.0) return 0;
    const double x = (k - lam) * inv_lam;
    const float xf = (float)x;
    float lg;   // lam * g(x),  g = x - (1+x) ln(1+x)
    if (fabsf(xf) <= 0.25f) {
        // g = -x^2 * sum_m (-x)^m/((m+1)(m+2)); 14 terms: |x| <= 1/4 -> rel. err < 1e-9.  lam*x^2 is formed
        // in fp64 (two multiplies) so that the large factor carries no fp32 error.
        float s = 1.0f / 210.0f;
        s = 1.0f / 182.0f - xf * s;
        s = 1.0f / 156.0f - xf * s;
        s = 1.0f / 132.0f - xf * s;
        s = 1.0f / 110.0f - xf * s;
        s = 1.0f / 90.0f - xf * s;
        s = 1.0f / 72.0f - xf * s;
        s = 1.0f / 56.0f - xf * s;
        s = 1.0f / 42.0f - xf * s;
        s = 1.0f / 30.0f - xf * s;
        s = 1.0f / 20.0f - xf * s;
        s = 1.0f / 12.0f - xf * s;
        s = 1.0f / 6.0f - xf * s;
        s = 0.5f - xf * s;
        lg = -(float)(lam * x * x) * s;
    } else {
        // |x| > 1/4 only happens for lam <~ 500 (k within ~6 sigma): lam*|g| <~ 100, the cancellation in
        // x - (1+x)log1p(x) costs < 10 ulp -> abs. error < 1e-4, inside the 1e-3 band
        lg = (float)lam * (xf - (1.0f + xf) * log1pf(xf));
    }
    const float kf = (float)k;
    const float rhs = lg - 1.0f / (12.0f * kf);
    const float usf = (float)us, bf = (float)b;
    const float invalpha = 1.1239f + 1.1328f / (bf - 3.4f);
    const float den = (float)a / (usf * usf) + bf;
    const float lhs = logf((float)V * invalpha * sqrtf(6.2831853f * kf) / den);
    const float d = lhs - rhs;
    if (d < -1.0e-3f) return 1;
    if (d > 1.0e-3f) return -1;
    return 0;
}

// One PTRS trial (numpy random_poisson_ptrs body): U from words 0,1; V from words 2,3.
// Returns true and sets k when the proposal is accepted.
HOLO_HD bool ptrs_trial(const FPrep& p, const Philox4& bits, double* kout) {
    const double lam = p.lam, b = p.a0, vr = p.a1, inv_lam = p.a2;
    const double a = -0.059 + 0.02483 * b;
    const double U = u53(bits.v[0], bits.v[1]) - 0.5;
    const double V = u53(bits.v[2], bits.v[3]);
    const double us = 0.5 - fabs(U);
    // 1/us by fp32 seed + two Newton steps (rel. err < 1e-14): an error here can only move a proposal that
    // sits within 1e-14 of an integer boundary, far below the resolution of U itself (2^-53)
    double rus = (double)(1.0f / (float)us);
    rus = rus * (2.0 - us * rus);
    rus = rus * (2.0 - us * rus);
    const double k = floor((2.0 * a * rus + b) * U + lam + 0.43);
    *kout = k;
    if ((us >= 0.07) && (V <= vr)) return true;
    if ((k < 0.0) || ((us < 0.013) && (V > us))) return false;
    const int scr = ptrs_screen(lam, inv_lam, a, b, us, V, k);
    if (scr != 0) return scr > 0;
    return ptrs_accept(lam, inv_lam, a, b, us, V, k);
}

HOLO_HD double draw_ptrs(const FPrep& p, const DrawKey& key, uint64_t idx) {
    for (uint32_t trial = 0; trial < 4096u; ++trial) {
        double k;
        if (ptrs_trial(p, element_bits(key, idx, trial), &k)) return k;
    }
    return floor(p.lam);   // unreachable in practice (acceptance ~0.9 per trial)
}

HOLO_HD double draw_normal(const FPrep& p, const DrawKey& key, uint64_t idx) {
    Philox4 b = element_bits(key, idx, 0);
    double u1 = 1.0 - u53(b.v[0], b.v[1]);   // (0, 1]
    double u2 = u53(b.v[2], b.v[3]);
    double z = sqrt(-2.0 * log(u1)) * cos(2.0 * CY_PI * u2);
    return p.lam + p.a0 * z;
}

// Stand-alone draw of one element (bulk sampling, eccentric kernel, host tests): same classes, but
// the 64-bit uniform comes from the element's own block instead of a shared group block.
HOLO_HD double draw_element(double lam, double thresh, const DrawKey& key, uint64_t idx) {
    FPrep p;
    int cls = prep_draw(lam, thresh, p);
    if (cls == CLS_EMPTY) return 0.0;
    if (cls == CLS_PTRS) return draw_ptrs(p, key, idx);
    if (cls == CLS_NORMAL) return draw_normal(p, key, idx);
    Philox4 bits = element_bits(key, idx, 0x800u);
    if (cls == CLS_TINY) {
        uint64_t t = double_as_bits(p.a0);
        uint64_t u = ((uint64_t)bits.v[0] << 32) | bits.v[1];
        if (u >= t) return 0.0;
        double T = -expm1(-lam);
        return invert_survival(lam, T, lam * exp(-lam), (double)u * (1.0 / 18446744073709551616.0));
    }
    return draw_small(p, bits.v[0], bits.v[1]);
}

}  // namespace holo
