// holo_eccen_math.cuh -- per-element math of the eccentric GWB kernel (K5), host/device inline so
// that tests/hostemu can exercise it on a CPU.  See holo_eccen.cu for the kernels.
#pragma once

#include "holo_common.cuh"

namespace holo {

constexpr double ECCEN_ZERO_LIMIT = 1.0e-12;   // cyutils.pyx:39

struct EccConsts {
    double gw_dadt_sep_const;   // cyutils.pyx:47
    double gw_src_const;        // cyutils.pyx:48
};

// J_{n-2}(x), J_{n-1}(x) for n >= 1, 0 < x < n, by Miller's backward recurrence (relative accuracy
// ~1e-15 at every order, unlike a forward recurrence or an absolute-error library routine).
// The reference calls scipy.special.jv (cyutils.pyx:180-181).
HOLO_HD void bessel_pair(int nn, double x, double* jm2, double* jm1) {
    // order -1: J_{-1} = -J_1
    const int hi = nn - 1;              // highest order needed (>= 0)
    const int lo = nn - 2;              // may be -1
    const int want_lo = lo < 0 ? 1 : lo;
    int top = (hi > want_lo ? hi : want_lo);
    int m = 2 * ((top + 8 + (int)sqrt(160.0 * (double)(top + 8))) / 2);
    double bjp = 0.0, bj = 1.0, sum = 0.0;
    double ans_hi = 0.0, ans_lo = 0.0;
    const double tox = 2.0 / x;
    for (int j = m; j > 0; --j) {
        double bjm = j * tox * bj - bjp;
        bjp = bj;
        bj = bjm;
        if (fabs(bj) > 1.0e150) {
            bj *= 1.0e-150; bjp *= 1.0e-150; sum *= 1.0e-150; ans_hi *= 1.0e-150; ans_lo *= 1.0e-150;
        }
        // bj now holds the (unnormalised) order j-1
        const int ord = j - 1;
        if ((ord & 1) == 0) sum += (ord == 0) ? bj : 2.0 * bj;   // 1 = J0 + 2 sum_{k>=1} J_{2k}
        if (ord == hi) ans_hi = bj;
        if (ord == want_lo) ans_lo = bj;
    }
    const double inv = 1.0 / sum;
    *jm1 = ans_hi * inv;
    *jm2 = (lo < 0) ? -(ans_lo * inv) : ans_lo * inv;
}

// gw_freq_dist_func__scalar_scalar, cyutils.pyx:145-194
HOLO_HD double gw_freq_dist_func(int nn, double ee) {
    if (ee < ECCEN_ZERO_LIMIT) return (nn == 2) ? 1.0 : 0.0;
    const double ne = nn * ee;
    const double n2 = (double)nn * (double)nn;
    double jn_m2, jn_m1;
    bessel_pair(nn, ne, &jn_m2, &jn_m1);
    // bessel_recursive, cyutils.pyx:54-76
    const double jn = (2 * (nn - 1) / ne) * jn_m1 - jn_m2;
    const double jn_p1 = (2 * nn / ne) * jn - jn_m1;
    const double jn_p2 = (2 * (nn + 1) / ne) * jn_p1 - jn;
    double aa = jn_m2 - 2.0 * ee * jn_m1 + (2.0 / nn) * jn + 2 * ee * jn_p1 - jn_p2;
    aa = aa * aa;
    // NOTE: the reference multiplies J_n by `ee` here (cyutils.pyx:188: `bb = jn_m2 - 2*ee*jn + jn_p2`);
    // Peters & Mathews (1963) have -2 J_n.  Parity means matching the reference.
    double bb = jn_m2 - 2 * ee * jn + jn_p2;
    bb = (1 - ee * ee) * bb * bb;
    const double cc = (4.0 / (3.0 * n2)) * jn * jn;
    return (n2 * n2 / 32) * (aa + bb + cc);
}

// _gw_ecc_func, cyutils.pyx:79-97
HOLO_HD double gw_ecc_func(double eccen) {
    const double e2 = eccen * eccen;
    return (1.0 + (73.0 / 24.0) * e2 + (37.0 / 96.0) * e2 * e2) / pow(1.0 - e2, 7.0 / 2.0);
}

// my_trapz_grid_weight, cyutils.pyx:102-142: rv[0] = inverse weight (1 or 2), rv[1] = width
HOLO_HD void trapz_grid_weight(int index, int size, const double* grid, double* w, double* dx) {
    if (index == 0) { *w = 2.0; *dx = grid[1] - grid[0]; return; }
    if (index == size - 1) { *w = 2.0; *dx = grid[index] - grid[index - 1]; return; }
    *w = 1.0;
    *dx = 0.5 * (grid[index + 1] - grid[index - 1]);
}

struct EccGeom {
    const double* ndens;        // (M,Q,Z)
    const double* mtot_log10;   // (M,)
    const double* mrat;         // (Q,)
    const double* redz;         // (Z,)
    const double* dcom;         // (Z,) [Mpc]
    const double* gwfobs;       // (F,)
    const double* sepa_evo;     // (E,)
    const double* eccen_evo;    // (E,)
    int M, Q, Z, F, E, H;
};

// The (f,n)-dependent factor for one (z, M):  returns false when (f,n) lies outside the evolution track.
//   *afac   = -sa^4/(C F(e) mt) * g(n,e) * 4/n^2 * f_r^(4/3)    [continuous: hterm = hterm_pref/(m1 m2) * afac]
//   *taufac = -sa^4/(C F(e) mt)                                  [discrete: tau = taufac/(m1 m2)]
//   *hfac   = g(n,e) * 4/n^2 * f_r^(4/3)
HOLO_HD bool eccen_factor(const EccGeom& g, const EccConsts& cc, const double* frst_pref, int kk, int ii, int ff,
                          int nh, double* afac, double* taufac, double* hfac) {
    const double zterm = 1.0 + g.redz[kk];
    const double mt = pow(10.0, g.mtot_log10[ii]);                       // pyx:498
    const double mt_sqrt = sqrt(mt);
    const double kep_sa_term = CY_NWTG / pow(2.0 * CY_PI, 2.0);          // pyx:423
    const double kep_sa_mass_term = kep_sa_term * mt;
    const double gwfr = g.gwfobs[ff] * zterm / nh;                       // pyx:542
    const double sa = pow(kep_sa_mass_term / pow(gwfr, 2.0), 1.0 / 3.0);  // pyx:545
    const double sa_fourth = pow(sa, 4.0);
    // smallest idx in [0, E-2] with gwfr <= frst_hi(idx)  (pyx:551-557)
    int lo = 0, hi = g.E - 2;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (gwfr > frst_pref[mid + 1] * mt_sqrt) lo = mid + 1; else hi = mid;
    }
    const int idx = lo;
    const double frst_lo = frst_pref[idx] * mt_sqrt;
    const double frst_hi = frst_pref[idx + 1] * mt_sqrt;
    if (gwfr < frst_lo) return false;                                    // pyx:560
    if (gwfr > frst_hi) return false;                                    // pyx:564
    double ecc = (g.eccen_evo[idx + 1] - g.eccen_evo[idx]) / (frst_hi - frst_lo);   // pyx:568
    ecc = g.eccen_evo[idx] + (gwfr - frst_lo) * ecc;
    const double gne = gw_freq_dist_func(nh, ecc);
    const double fe_ecc = gw_ecc_func(ecc);
    const double four_over_nh_squared = 4.0 / ((double)nh * (double)nh);
    const double tf = -sa_fourth / (cc.gw_dadt_sep_const * fe_ecc * mt);
    const double hf = gne * four_over_nh_squared * pow(gwfr, 4.0 / 3.0);
    *taufac = tf;
    *hfac = hf;
    *afac = tf * hf;
    return true;
}

}  // namespace holo
