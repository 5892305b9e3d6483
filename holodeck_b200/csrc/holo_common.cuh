// holo_common.cuh -- shared constants and small device helpers for libholo_b200.
//
// Every function marked HOLO_HD is plain C++ arithmetic that compiles for the host as well, so
// the per-cell math of each kernel can be exercised by `tests/hostemu` on a machine without a GPU.
// That host build is a debugging aid for the *tests*; the product only ever calls the __global__
// kernels.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define HOLO_HD __host__ __device__ __forceinline__
#define HOLO_D __device__ __forceinline__
#define HOLO_NOINLINE_STATIC static __host__ __device__ __noinline__
#else
#define HOLO_HD inline
#define HOLO_D inline
#define HOLO_NOINLINE_STATIC static inline
#endif

namespace holo {

// ---- constant set A: what the reference's Cython kernels hard-code
//      (holodeck/sams/sam_cyutils.pyx:30-42, holodeck/cyutils.pyx:43-48)
constexpr double CY_NWTG = 6.6742999e-08;
constexpr double CY_SPLC = 29979245800.0;
constexpr double CY_MPC = 3.08567758e+24;
constexpr double CY_MSOL = 1.988409870698051e+33;
constexpr double CY_YR = 31557600.0;
constexpr double CY_SCHW = 1.4852320538237328e-28;
constexpr double CY_PI = 3.14159265358979323846;

// ---- constant set B: what the reference's numpy layer gets from astropy
//      (holodeck/constants.py:23-57; CODATA-2018)
constexpr double AP_NWTG = 6.6743e-08;
constexpr double AP_SPLC = 29979245800.0;
constexpr double AP_PC = 3.0856775814913674e+18;
constexpr double AP_MPC = 1.0e6 * AP_PC;

// Derived constants.  The reference evaluates these once at module import with libm `pow`/`sqrt`
// (sam_cyutils.pyx:36-42); the host wrapper computes them the same way and passes them in, so that
// the kernels see bit-identical values (no device `pow` in the constant path).
struct CyConsts {
    double gw_dadt_sep_const;   // -64 G^3 / (5 c^5)              sam_cyutils.pyx:36
    double kepler_const_freq;   // sqrt(G) / (2 pi)               sam_cyutils.pyx:40
    double kepler_const_sepa;   // G^(1/3) / (2 pi)^(2/3)         sam_cyutils.pyx:41
    double four_pi_c_over_mpc;  // 4 pi c / Mpc                   sam_cyutils.pyx:42
};

// ---- reference helper restatements -------------------------------------------------------------

// cyutils.pyx:326-331
HOLO_HD double interp_between_vals(double xnew, double xl, double xr, double yl, double yr) {
    return yl + (yr - yl) * (xnew - xl) / (xr - xl);
}

// cyutils.pyx:338-358
HOLO_HD double interp_at_index(int idx, double xnew, const double* xold, const double* yold) {
    return interp_between_vals(xnew, xold[idx], xold[idx + 1], yold[idx], yold[idx + 1]);
}

// Stateless equivalent of `while_while_increasing` (sam_cyutils.pyx:66-84): the left index of the
// interval of the INCREASING array `edges` that bounds `val`, clamped to [0, size-2] (so values
// beyond either end are linearly extrapolated by `interp_at_index`, exactly as in the reference).
// At an exact tie (val == edges[k]) the reference's answer depends on its carried hint (k-1 or k);
// both give the same interpolated value up to rounding.  We return k.
HOLO_HD int bracket_increasing(int size, double val, const double* edges) {
    int lo = 0, hi = size - 1;   // invariant: answer in [lo, hi)
    // largest idx in [0, size-2] with edges[idx] <= val  (0 if none)
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (edges[mid] <= val) lo = mid; else hi = mid;
    }
    return lo;
}

// Stateless equivalent of `while_while_decreasing` (sam_cyutils.pyx:89-107) started from a hint
// that does not overshoot: smallest idx with edges[idx+1] <= val on a DECREASING array, clamped to
// [0, size-2].
HOLO_HD int bracket_decreasing(int size, double val, const double* edges) {
    int lo = 0, hi = size - 1;   // answer in [lo, hi)
    // smallest idx such that edges[idx+1] <= val  <=> largest idx with edges[idx] > val (or 0)
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (edges[mid] > val) lo = mid; else hi = mid;
    }
    return lo;
}

// sam_cyutils.pyx:45-49
HOLO_HD double hard_gw(const CyConsts& cc, double mtot, double mrat, double sepa) {
    return cc.gw_dadt_sep_const * pow(mtot, 3.0) * mrat / pow(sepa, 3.0) / pow(1.0 + mrat, 2.0);
}

// sam_cyutils.pyx:52-55
HOLO_HD double kepler_freq_from_sepa(const CyConsts& cc, double mtot, double sepa) {
    return cc.kepler_const_freq * sqrt(mtot) / pow(sepa, 1.5);
}

// sam_cyutils.pyx:58-61
HOLO_HD double kepler_sepa_from_freq(const CyConsts& cc, double mtot, double freq) {
    return cc.kepler_const_sepa * pow(mtot, 1.0 / 3.0) / pow(freq, 2.0 / 3.0);
}

// sam_cyutils.pyx:240-253
HOLO_HD double hard_func_2pwl(double norm, double xx, double gamma_inner, double gamma_outer) {
    return -norm * pow(1.0 + xx, -gamma_outer + gamma_inner) / pow(xx, gamma_inner - 1.0);
}

// x^e for the per-(z,f) evaluation of K1b: when 2e is a small integer (the library's gamma_inner = -1,
// gamma_outer = +2.5 give e = -3.5 and e = -2) the power is a square root and a few multiplications (<= 3 ulp
// from libm's pow, far inside the 1e-10 parity tolerance) instead of a ~150-instruction generic pow.
struct HalfPow {
    double e;
    int n2;      // 2e when that is an integer with |2e| <= 32, else HALF_POW_GENERIC
};
constexpr int HALF_POW_GENERIC = 1 << 20;

HOLO_HD HalfPow half_pow_spec(double e) {
    HalfPow s;
    s.e = e;
    const double t = 2.0 * e;
    s.n2 = (t == floor(t) && fabs(t) <= 32.0) ? (int)t : HALF_POW_GENERIC;
    return s;
}

HOLO_HD double half_pow(double x, const HalfPow& s) {
    if (s.n2 == HALF_POW_GENERIC) return pow(x, s.e);
    const int n = s.n2 < 0 ? -s.n2 : s.n2;
    double r = (n & 1) ? sqrt(x) : 1.0;
    double b = x;
    for (int m = n >> 1; m; m >>= 1) {
        if (m & 1) r *= b;
        b *= b;
    }
    return s.n2 < 0 ? 1.0 / r : r;
}

HOLO_HD double hard_func_2pwl_spec(double norm, double xx, const HalfPow& p_outer, const HalfPow& p_inner) {
    return -norm * half_pow(1.0 + xx, p_outer) / half_pow(xx, p_inner);   // sam_cyutils.pyx:240-253
}

HOLO_HD double hard_func_2pwl_gw(const CyConsts& cc, double mtot, double mrat, double sepa, double norm,
                                 double rchar, double gamma_inner, double gamma_outer) {
    double dadt = hard_func_2pwl(norm, sepa / rchar, gamma_inner, gamma_outer);
    dadt += hard_gw(cc, mtot, mrat, sepa);
    return dadt;
}

// ---- flat-LCDM comoving distance (host twin: holodeck_b200/cosmology.py:comoving_distance) -----
constexpr int GL_ORDER = 16;

struct GLTable {
    double x[GL_ORDER];
    double w[GL_ORDER];
};

// d_c(z) [cm] = hubble_distance * int_{s0}^{1} 2 ds / sqrt(Om0 + OL s^6),  s0 = (1+z)^(-1/2)
HOLO_HD double comoving_distance_cm(const GLTable& gl, double hubble_distance, double om0, double zz) {
    double sq = sqrt(1.0 + zz);
    double half = 0.5 * zz / (sq * (sq + 1.0));
    double mid = 1.0 - half;
    double ode0 = 1.0 - om0;
    double tot = 0.0;
    for (int i = 0; i < GL_ORDER; ++i) {
        double ss = mid + half * gl.x[i];
        double s2 = ss * ss;
        tot = tot + gl.w[i] * (2.0 / sqrt(om0 + ode0 * s2 * s2 * s2));
    }
    return hubble_distance * half * tot;
}

// The same distance from a table (optional, built by the host once per cosmology: holodeck_b200/_lib.py:dc_table):
// with w = 1 - (1+z)^(-1/2), d_c = hubble_distance * w * R(w), R(w) = (1/w) int_{1-w}^{1} 2 ds / sqrt(Om0 + OL s^6), a
// smooth O(1) function tabulated at n+1 uniform nodes in w as pairs (R_i, h R'_i) and evaluated by cubic Hermite
// interpolation (relative error 2e-15 at n = 2048, uniformly down to z -> 0 because R, not d_c, is interpolated).
// One sqrt, one division and two 16 B loads instead of 16 sqrt + 16 divisions.  Returns < 0 outside the table.
HOLO_HD double comoving_distance_table(const double* tab, int n, double inv_h, double hubble_distance, double zz) {
    const double sq = sqrt(1.0 + zz);
    const double w = zz / (sq * (sq + 1.0));
    const double u = w * inv_h;
    const int i = (int)u;
    if (!(u >= 0.0) || i >= n) return -1.0;
    const double t = u - (double)i;
    const double r0 = tab[2 * i], d0 = tab[2 * i + 1], r1 = tab[2 * i + 2], d1 = tab[2 * i + 3];
    const double omt = 1.0 - t, t2 = t * t, o2 = omt * omt;
    const double rr = ((1.0 + 2.0 * t) * o2) * r0 + (t * o2) * d0 + (t2 * (3.0 - 2.0 * t)) * r1 - (t2 * omt) * d1;
    return hubble_distance * w * rr;
}

}  // namespace holo
