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
//   2^-8 <= lam <= 4000, inside the realization kernel only
//                TABLE : the CDF over the window [lam - 9.5 sd, lam + 6.5 sd] is tabulated ONCE per CTA as 32-bit
//                        thresholds in shared memory and shared by all realizations: a draw is one 32-bit
//                        word + a fixed-depth binary search, free of lane divergence.  Draws whose word
//                        falls within one unit of a threshold (prob. ~3W 2^-32) are resolved exactly with
//                        32 more bits against the fp64 CDF, so the sampler stays exact to 64-bit uniforms.
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
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;   // one IMAD.WIDE.U32 each
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
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
    PURPOSE_PASS_COUNT = 4, // event count of a pass's superposition group (draw_group), keyed on the pass
    PURPOSE_PASS_PICK = 5,  // member picks of the same group: two 53-bit uniforms per block
    PURPOSE_QUAD_HI = 6,    // one block per (element, QUAD of four consecutive realizations): word u -> realization 4q + u
    PURPOSE_QUAD_LO = 7,    // second block of the same quad: low halves of the 64-bit uniforms
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

// block shared by the four consecutive realizations 4q .. 4q+3 of ONE element (cell, frequency slot `fi` of group
// `fg`): the parameter variants of the realization kernel give a thread four realizations of one frequency at a time
HOLO_HD Philox4 quad_bits(uint32_t k0, uint32_t k1, uint32_t stream, uint32_t cell, uint32_t fg, uint32_t fi, uint32_t quad,
                          int purpose) {
    return philox4x32_10(cell, fg | (fi << 24) | ((uint32_t)purpose << 28), quad, stream << 24, k0, k1);
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
enum { CLS_EMPTY = 0, CLS_TINY = 1, CLS_SMALL = 2, CLS_PTRS = 3, CLS_NORMAL = 4, CLS_TABLE = 5 };

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
HOLO_NOINLINE_STATIC double draw_tiny_extend(FPrep p, uint32_t hi, DrawKey key, uint64_t idx);

HOLO_HD double draw_tiny(const FPrep& p, uint32_t hi, const DrawKey& key, uint64_t idx) {
    const uint64_t t = double_as_bits(p.a0);
    if (hi > (uint32_t)(t >> 32)) return 0.0;
    return draw_tiny_extend(p, hi, key, idx);
}

HOLO_NOINLINE_STATIC double draw_tiny_extend(FPrep p, uint32_t hi, DrawKey key, uint64_t idx) {
    const uint64_t t = double_as_bits(p.a0);
    Philox4 ext = element_bits(key, idx, 0);
    uint64_t u = ((uint64_t)hi << 32) | ext.v[0];
    if (u >= t) return 0.0;
    double T = -expm1(-p.lam);
    return invert_survival(p.lam, T, p.lam * exp(-p.lam), (double)u * (1.0 / 18446744073709551616.0));
}

// ln(k!) for k < 32 (correctly rounded).  Kept in constant memory on the device: a function-local array
// indexed at run time would be rebuilt on the stack of every thread.
#define HOLO_LNFACT32_INIT                                                                                         \
    {0.0, 0.0, 0.6931471805599453, 1.791759469228055, 3.1780538303479458, 4.787491742782046, 6.579251212010101,    \
     8.525161361065415, 10.60460290274525, 12.801827480081469, 15.104412573075516, 17.502307845873887,             \
     19.987214495661885, 22.552163853123425, 25.19122118273868, 27.89927138384089, 30.671860106080672,             \
     33.50507345013689, 36.39544520803305, 39.339884187199495, 42.335616460753485, 45.38013889847691,              \
     48.47118135183523, 51.60667556776438, 54.78472939811232, 58.00360522298052, 61.261701761002,                  \
     64.55753862700634, 67.88974313718154, 71.25703896716801, 74.65823634883016, 78.0922235533153}
static const double h_lnfact32[32] = HOLO_LNFACT32_INIT;
#if defined(__CUDACC__)
static __constant__ double d_lnfact32[32] = HOLO_LNFACT32_INIT;
#endif

HOLO_HD double log_factorial_32(int k) {
#if defined(__CUDA_ARCH__)
    return d_lnfact32[k];
#else
    return h_lnfact32[k];
#endif
}

HOLO_HD double log_factorial_small(int k) { return log_factorial_32(k); }   // k < 10

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

// The same trial in two parts, for callers that advance several independent draws in lock-step: the straight-line
// part (proposal k and the two cheap tests, numpy's squeeze steps) and the rarely needed exact acceptance test.
// `ptrs_propose` returns +1 accept, -1 reject, 0 undecided; the outcome of a trial is the same as `ptrs_trial`'s.
HOLO_HD int ptrs_propose(double lam, double b, double vr, const Philox4& bits, double* kout, double* us_out, double* V_out) {
    const double a = -0.059 + 0.02483 * b;
    const double U = u53(bits.v[0], bits.v[1]) - 0.5;
    const double V = u53(bits.v[2], bits.v[3]);
    const double us = 0.5 - fabs(U);
    double rus = (double)(1.0f / (float)us);
    rus = rus * (2.0 - us * rus);
    rus = rus * (2.0 - us * rus);
    const double k = floor((2.0 * a * rus + b) * U + lam + 0.43);
    *kout = k;
    *us_out = us;
    *V_out = V;
    if ((us >= 0.07) && (V <= vr)) return 1;
    if ((k < 0.0) || ((us < 0.013) && (V > us))) return -1;
    return 0;
}
HOLO_NOINLINE_STATIC bool ptrs_decide(double lam, double b, double us, double V, double k) {
    const double a = -0.059 + 0.02483 * b, inv_lam = 1.0 / lam;
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

// ---- TABLE class ----------------------------------------------------------------------------------
// Window of the tabulated CDF: below kmin the Poisson mass is < 2^-64 (lighter than the normal tail at
// -9.5 sd), above kmin+W-1 it is ~1e-10: draws that land there take the exact slow path.
#ifndef HOLO_TABLE_MAX_LAM
#define HOLO_TABLE_MAX_LAM 4000.0
#endif
constexpr double TABLE_MAX_LAM = HOLO_TABLE_MAX_LAM;
#ifndef HOLO_TABLE_WMAX
#define HOLO_TABLE_WMAX 1024
#endif
constexpr int TABLE_WMAX = HOLO_TABLE_WMAX;    // >= table_spec(TABLE_MAX_LAM).W

struct TableSpec {
    int kmin, W;
};

HOLO_HD TableSpec table_spec(double lam) {
    const double sd = sqrt(lam);
    double lo = floor(lam - 9.5 * sd - 2.0);
    if (lo < 0.0) lo = 0.0;
    const double hi = ceil(lam + 6.5 * sd + 8.0);
    TableSpec t;
    t.kmin = (int)lo;
    t.W = (int)(hi - lo) + 1;
    return t;
}

// Poisson pmf at integer k >= 0, relative error ~1e-14: exact ln k! below 32, above it Stirling's series
// with the large terms combined analytically into lam*[x - (1+x)ln(1+x)], x = (k-lam)/lam.
HOLO_HD double poisson_pmf(double k, double lam, double ln_lam, double inv_lam) {
    if (k < 32.0) return exp(k * ln_lam - lam - log_factorial_32((int)k));
    const double x = (k - lam) * inv_lam;
    double g;
    if (fabs(x) < 0.0625) {
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
    const double ik = 1.0 / k, ik2 = ik * ik;
    const double corr = ik * (1.0 / 12.0 - ik2 * (1.0 / 360.0 - ik2 * (1.0 / 1260.0 - ik2 * (1.0 / 1680.0))));
    return exp(lam * g - corr) / sqrt(6.283185307179586 * k);
}

HOLO_HD uint32_t cdf_to_u32(double c) {   // floor(c * 2^32), clamped to [0, 2^32 - 1]
    if (!(c > 0.0)) return 0u;
    const double s = c * 4294967296.0;
    if (s >= 4294967295.0) return 0xFFFFFFFFu;
    return (uint32_t)s;
}

// A table is built by 32 cooperating lanes: lane l owns window entries [j0, j1).  Step 1 evaluates the pmf
// at the TOP of the segment and recurs downwards (pmf(k-1) = pmf(k) k / lam: no division), returning the
// segment's mass; an inclusive scan over the lanes gives the CDF at each segment top; step 2 recurs
// downwards again, writing floor(CDF 2^32).
HOLO_HD double table_segment_mass(double lam, double ln_lam, double inv_lam, int kmin, int j0, int j1, double* ptop) {
    *ptop = 0.0;
    if (j1 <= j0) return 0.0;
    double k = (double)(kmin + j1 - 1);
    double p = poisson_pmf(k, lam, ln_lam, inv_lam);
    *ptop = p;
    double s = 0.0;
    for (int j = j1 - 1; j >= j0; --j) {
        s += p;
        p *= k * inv_lam;
        k -= 1.0;
    }
    return s;
}

HOLO_HD void table_segment_write(uint32_t* t, double cdf_top, double ptop, double inv_lam, int kmin, int j0, int j1) {
    double c = cdf_top, p = ptop, k = (double)(kmin + j1 - 1);
    for (int j = j1 - 1; j >= j0; --j) {
        t[j] = cdf_to_u32(c);
        c -= p;
        p *= k * inv_lam;
        k -= 1.0;
    }
}

// #{j < W : t[j] <= hi}; the trip count depends on W only (lock-step over a warp)
HOLO_HD int table_count(const uint32_t* t, int W, uint32_t hi) {
    int base = 0, len = W;
    while (len > 1) {
        const int half = len >> 1;
        if (t[base + half - 1] <= hi) base += half;
        len -= half;
    }
    return base + ((t[base] <= hi) ? 1 : 0);
}

// Exact resolution of a draw whose 32-bit word is within one unit of a threshold (or beyond the window):
// all entries below the first threshold >= hi-1 are certainly <= u; from there the CDF is summed in
// fp64 (downward recurrence for the starting value, upward walk afterwards) and compared with the
// 64-bit uniform.
HOLO_NOINLINE_STATIC double table_resolve(double lam, const uint32_t* t, int kmin, int W, int nidx, uint32_t hi, uint32_t lo) {
    int js = nidx < W ? nidx : W - 1;
    while (js > 0 && (int64_t)t[js - 1] >= (int64_t)hi - 1) --js;
    const double ln_lam = log(lam), inv_lam = 1.0 / lam;
    const double u = (double)(((uint64_t)hi << 32) | lo) * (1.0 / 18446744073709551616.0);
    double k = (double)(kmin + js);
    const double pk = poisson_pmf(k, lam, ln_lam, inv_lam);
    double S = 0.0, p = pk, kk = k;
    for (int j = js; j >= 0; --j) {      // S = sum_{i=kmin}^{kmin+js} pmf(i)
        S += p;
        p *= kk * inv_lam;
        kk -= 1.0;
    }
    p = pk;
    for (int it = 0; it < 100000 && !(u < S); ++it) {
        k += 1.0;
        p *= lam / k;
        S += p;
        if (p < 1e-300) break;
    }
    return k;
}

// Fast path: returns the count, or -1 when the word alone does not decide it.
HOLO_HD double draw_table_fast(const uint32_t* t, int kmin, int W, uint32_t hi, int* nidx_out) {
    const int n = table_count(t, W, hi);
    *nidx_out = n;
    bool amb = (n >= W) || (n == 0 && hi <= 1u);   // (same rule as the sentinel form of draw_table_ladder)
    if (n > 0 && hi - t[n - 1] <= 1u) amb = true;
    if (n < W && t[n] - hi <= 1u) amb = true;
    return amb ? -1.0 : (double)(kmin + n);
}

// #{j < W : t[j] <= hi} by a fixed sequence of probes (lg = floor(log2 W)): the first probe picks the lower
// or the upper 2^lg entries, the unrolled ladder then halves the step with immediate offsets.  The table is
// addressed as pool[q + j] with a 32-bit index so that the device code is one LDS + compare + predicated add
// per rung; N independent draws from the same table (the realization slots a thread carries) climb the
// ladder together, which shares the rung dispatch and gives the scheduler N independent dependency chains.
// On return q[u] = toff + count (pool index of the first threshold above draw u).
// (thresholds are addressed by BYTE offset from `pool`, so that a rung is LDS [reg + imm], compare, predicated add)
HOLO_HD uint32_t pool_at(const uint32_t* pool, uint32_t byte_off) {
    return *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(pool) + byte_off);
}

template <int N>
HOLO_HD void table_ladder_n(const uint32_t* pool, uint32_t toff, int W, int lg, const uint32_t (&hi)[N], uint32_t (&q)[N]) {
    const uint32_t P = 1u << lg;
    uint32_t qb[N];   // byte offsets
#define HOLO_RUNG(step)                                                         \
    _Pragma("unroll") for (int u = 0; u < N; ++u) {                             \
        if (pool_at(pool, qb[u] + 4u * ((step) - 1u)) <= hi[u]) qb[u] += 4u * (step); \
    }
    {
        const uint32_t top = pool_at(pool, 4u * (toff + P - 1u)), skip = 4u * ((uint32_t)W - P);
#pragma unroll
        for (int u = 0; u < N; ++u) qb[u] = 4u * toff + ((top <= hi[u]) ? skip : 0u);
    }
    switch (lg) {
        case 10: HOLO_RUNG(512u)   // fall through
        case 9: HOLO_RUNG(256u)
        case 8: HOLO_RUNG(128u)
        case 7: HOLO_RUNG(64u)
        case 6: HOLO_RUNG(32u)
        case 5: HOLO_RUNG(16u)
        case 4: HOLO_RUNG(8u)
        case 3: HOLO_RUNG(4u)
        case 2: HOLO_RUNG(2u)
        case 1: HOLO_RUNG(1u)
        default: break;
    }
    HOLO_RUNG(1u)
#undef HOLO_RUNG
#pragma unroll
    for (int u = 0; u < N; ++u) q[u] = qb[u] >> 2;
}

// Is the draw whose ladder ended at pool index q decided by its 32-bit word?  pool[toff - 1] = 0 and
// pool[toff + W] = 2^32 - 1 are sentinels, so the two neighbouring thresholds can be read blindly.
HOLO_HD bool table_ambiguous(const uint32_t* pool, uint32_t toff, int W, uint32_t q, uint32_t hi) {
    const uint32_t below = hi - pool_at(pool, 4u * q - 4u), above = pool_at(pool, 4u * q) - hi;
    return (q >= toff + (uint32_t)W) | (below <= 1u) | (above <= 1u);
}

// single TABLE draw: the count, or -1 when the 32-bit word does not decide it (toff >= 1)
HOLO_HD double draw_table_ladder(const uint32_t* pool, uint32_t toff, int kmin, int W, int lg, uint32_t hi, int* nidx) {
    const uint32_t h1[1] = {hi};
    uint32_t q1[1];
    table_ladder_n<1>(pool, toff, W, lg, h1, q1);
    *nidx = (int)(q1[0] - toff);
    return table_ambiguous(pool, toff, W, q1[0], hi) ? -1.0 : (double)(kmin + *nidx);
}

// The superposition group of a pass: its members (expectation values lam_k, inclusive cumulative sums `gcum`,
// total `lam_tot` = gcum[ngrp-1]) are drawn as ONE Poisson process of rate lam_tot; each of its N events
// belongs to member k with probability lam_k / lam_tot.  By the superposition / thinning theorem the
// member counts are independent Poisson(lam_k) -- exact, and O(1 + N) instead of O(ngrp) per realization.
// pool[toff ...] is the CDF table of Poisson(lam_tot); the Philox counter is keyed on the pass (its first cell).
template <class OnEvent>
HOLO_HD void draw_group(const uint32_t* pool, uint32_t toff, int kmin, int W, int lg, double lam_tot, const double* gcum,
                        int ngrp, uint32_t pass_id, uint32_t fg, const DrawKey& key, OnEvent&& on_event) {
    const Philox4 gb = philox4x32_10(pass_id, fg | ((uint32_t)PURPOSE_PASS_COUNT << 28), key.real, key.stream << 24,
                                     key.k0, key.k1);
    int nidx;
    double nev = draw_table_ladder(pool, toff, kmin, W, lg, gb.v[0], &nidx);
    if (nev < 0.0) nev = table_resolve(lam_tot, pool + toff, kmin, W, nidx, gb.v[0], gb.v[1]);
    const int nevents = (int)nev;
    Philox4 pb;
    pb.v[0] = pb.v[1] = pb.v[2] = pb.v[3] = 0u;
    for (int ev = 0; ev < nevents; ++ev) {
        if ((ev & 1) == 0)
            pb = philox4x32_10(pass_id, fg | ((uint32_t)PURPOSE_PASS_PICK << 28), key.real,
                               (key.stream << 24) | (uint32_t)(ev >> 1), key.k0, key.k1);
        const double v = ((ev & 1) == 0 ? u53(pb.v[0], pb.v[1]) : u53(pb.v[2], pb.v[3])) * lam_tot;
        int base = 0, len = ngrp;      // member index = #{j : gcum[j] <= v}
        while (len > 1) {
            const int half = len >> 1;
            if (gcum[base + half - 1] <= v) base += half;
            len -= half;
        }
        if (gcum[base] <= v && base < ngrp - 1) base += 1;
        on_event(base);
    }
}

HOLO_HD double draw_normal(const FPrep& p, const DrawKey& key, uint64_t idx) {
    Philox4 b = element_bits(key, idx, 0);
    double u1 = 1.0 - u53(b.v[0], b.v[1]);   // (0, 1]
    double u2 = u53(b.v[2], b.v[3]);
    double z = sqrt(-2.0 * log(u1)) * cos(2.0 * CY_PI * u2);
    return p.lam + p.a0 * z;
}

// (the out-of-line helpers take the key BY VALUE: a reference would pin the caller's key to local memory)
HOLO_NOINLINE_STATIC double draw_normal_lam(double lam, DrawKey key, uint64_t idx) {
    FPrep p;
    p.lam = lam;
    p.a0 = sqrt(lam);
    return draw_normal(p, key, idx);
}

// slow path of a TABLE draw inside the realization kernel: fetch the low word of the group's 64-bit uniform
HOLO_NOINLINE_STATIC double table_resolve_keyed(double lam, const uint32_t* t, int kmin, int W, int nidx, uint32_t hi,
                                                DrawKey key, uint32_t cell, uint32_t fg, int fi) {
    const Philox4 lo = group_bits(key, cell, fg, PURPOSE_GROUP_LO);
    return table_resolve(lam, t, kmin, W, nidx, hi, lo.v[fi]);
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
