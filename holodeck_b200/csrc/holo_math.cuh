// holo_math.cuh -- per-element math of the deterministic kernels (K0, K1, K2, K2b), written as
// HOLO_HD inline functions so the CUDA kernels in holo_deterministic.cu are thin parallel wrappers
// and `tests/hostemu` can run the identical arithmetic on a CPU for debugging.
#pragma once

#include "holo_common.cuh"
#include "../../include/holo_b200.h"

namespace holo {

// =================================================================================================
// K0: static binary density   (reference: holodeck/sams/sam.py:250-365)
// =================================================================================================

// BF_Sigmoid (host_relations.py:198-331): its inverse relations are scipy quadratic interpolants; the host hands
// them over as piecewise polynomials [breaks (n+1) | c0 (n) | c1 (n) | c2 (n)] (first / last piece extrapolate).
HOLO_HD double bf_spline(const double* tab, int n, double x) {
    int lo = 0, hi = n;                       // largest lo in [0, n-1] with breaks[lo] <= x (clamped)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tab[mid] <= x) lo = mid; else hi = mid;
    }
    const double dx = x - tab[lo];
    const double* c = tab + (n + 1);
    return (c[lo] * dx + c[n + lo]) * dx + c[2 * n + lo];
}

// BF_Sigmoid.bulge_frac host_relations.py:286-295
HOLO_HD double bf_sigmoid_frac(const holo_sam_params& p, double mstar) {
    double mm = mstar / p.bf[2];
    if (mm > 1.0) mm = 1.0;
    const double frac = p.bf[0] + (p.bf[1] - p.bf[0]) / (1.0 + pow((1.0 / mm) - 1.0, p.bf[3]));
    return ((mm >= 1.0) || (frac > p.bf[1])) ? p.bf[1] : frac;
}

// MMBulge_Standard.mstar_from_mbh: host_relations.py:768-771 -> mbulge_from_mbh :745-765 ->
// _log10_relation_reverse :1137-1178 ; BF_Constant.mstar_from_mbulge :190-192 / BF_Sigmoid.mstar_from_mbulge :297-308
HOLO_HD double mstar_from_mbh(const holo_sam_params& p, const double* bf_tab, double mbh) {
    double xx = log10(mbh / p.mmb[0]);
    xx = 1.0 / p.mmb[1] * xx;
    double mbulge = p.mmb[2] * pow(10.0, xx);
    if (p.bf_kind == 0) return mbulge / p.mmb[3];
    const double mstar = mbulge / p.bf[1];
    return ((mstar / p.bf[2]) < 1.0) ? bf_spline(bf_tab, p.bf_n, mbulge) : mstar;
}

// _MMBulge_Relation.dmstar_dmbh host_relations.py:483-512 with MMBulge_Standard.dmbulge_dmbh :720-743
HOLO_HD double dmstar_dmbh(const holo_sam_params& p, const double* bf_tab, double mstar) {
    double mbulge, dmstar_dmbulge;
    if (p.bf_kind == 0) {
        mbulge = mstar * p.mmb[3];
        dmstar_dmbulge = 1.0 / p.mmb[3];
    } else {
        mbulge = mstar * bf_sigmoid_frac(p, mstar);                    // _Bulge_Frac.mbulge_from_mstar :113-131
        // BF_Sigmoid.dmstar_dmbulge :310-321 (second table)
        dmstar_dmbulge = ((mbulge / p.bf[2]) < p.bf[1]) ? bf_spline(bf_tab + 4 * p.bf_n + 1, p.bf_n, mbulge) : 1.0 / p.bf[1];
    }
    // mbh_from_mbulge -> _log10_relation (host_relations.py:1102-1134)
    double yy = log10(mbulge / p.mmb[2]) * p.mmb[1];
    double mbh = p.mmb[0] * pow(10.0, yy);
    double dmbulge_dmbh = mbulge / (p.mmb[1] * mbh);
    return dmstar_dmbulge * dmbulge_dmbh;
}

// GSMF_Schechter.__call__ components.py:132-172 ; GSMF_Double_Schechter :315-329 (+ :236-270)
HOLO_HD double gsmf_eval(const holo_sam_params& p, double mstar, double redz) {
    const double LN10 = 2.302585092994046;   // np.log(10.0)
    if (p.gsmf_kind == 0) {
        double phi = pow(10.0, p.gsmf[0] + p.gsmf[1] * redz);
        double mchar = p.gsmf[2] + p.gsmf[3] * redz;
        double alpha = p.gsmf[4] + p.gsmf[5] * redz;
        double xx = mstar / mchar;
        return LN10 * phi * pow(xx, 1.0 + alpha) * exp(-xx);
    }
    double z2 = redz * redz;   // numpy `redz**2`
    double mchar = p.gsmf[11] * pow(10.0, p.gsmf[6] + p.gsmf[7] * redz + p.gsmf[8] * z2);
    double xx = mstar / mchar;
    double phi1 = pow(10.0, p.gsmf[0] + p.gsmf[1] * redz + p.gsmf[2] * z2);
    double phi2 = pow(10.0, p.gsmf[3] + p.gsmf[4] * redz + p.gsmf[5] * z2);
    double v1 = LN10 * phi1 * pow(xx, 1.0 + p.gsmf[9]) * exp(-xx);
    double v2 = LN10 * phi2 * pow(xx, 1.0 + p.gsmf[10]) * exp(-xx);
    return v1 + v2;
}

// GMR_Illustris.__call__ components.py:452-482
HOLO_HD double gmr_eval(const holo_sam_params& p, double mtot, double mrat, double redz) {
    double zp1 = 1.0 + redz;
    double norm = p.gmr[0] * pow(zp1, p.gmr[1]);
    double malpha = p.gmr[2] * pow(zp1, p.gmr[3]);
    double mdelta = p.gmr[4] * pow(zp1, p.gmr[5]);
    double qgamma = p.gmr[6] * pow(zp1, p.gmr[7]);
    qgamma = qgamma + p.gmr[8] * log10(mtot / p.gmr[9]);
    double xx = mtot / p.gmr[9];
    double mt = pow(xx, malpha);
    double yy = mtot / p.gmr[10];
    double mp1t = pow(1.0 + yy, mdelta);
    double qt = pow(mrat, qgamma);
    return norm * mt * mp1t * qt;
}

// closed-form inverse of the flat-LCDM age (host twin: cosmology.py:tage_to_z)
HOLO_HD double tage_to_z(const holo_sam_params& p, double age) {
    double ode0 = 1.0 - p.om0;
    double sq = sqrt(ode0);
    double sh = sinh(1.5 * sq * age / p.hubble_time);
    double zp1 = pow(sqrt(ode0 / p.om0) / sh, 2.0 / 3.0);
    return zp1 - 1.0;
}

struct DensityOut {
    double dens, gmt_time, redz_prime;
};

// One (M,q,z) grid point of sam.py:310-365.
HOLO_HD DensityOut density_point(const holo_sam_params& p, const double* bf_tab, double mtot, double mrat, double redz,
                                 double age_z, double dtdz_z) {
    DensityOut out;
    // mass_stellar(): sam.py:250-278 ; utils.m1m2_from_mtmr utils.py:1620-1642
    double m1 = mtot / (1.0 + mrat);
    double m2 = mtot - m1;
    double mstar_pri = mstar_from_mbh(p, bf_tab, m1);
    double mstar_sec = mstar_from_mbh(p, bf_tab, m2);
    double mstar_rat = mstar_sec / mstar_pri;
    double mstar_tot = mstar_pri + mstar_sec;
    double mass_gsmf = p.gsmf_uses_mtot ? mstar_tot : mstar_pri;

    out.gmt_time = 0.0;
    out.redz_prime = redz;
    if (p.has_gmt) {
        // GMT_Power_Law.__call__ components.py:646-675 ; zprime :620-626 ; redz_after utils.py:1772
        double mass_gmt = p.gmt_uses_mtot ? mstar_tot : mstar_pri;
        double tau0 = p.gmt[0] * pow(mass_gmt / p.gmt[1], p.gmt[2]) * pow(1.0 + redz, p.gmt[3]) *
                      pow(mstar_rat, p.gmt[4]);
        double new_age = age_z + tau0;
        out.gmt_time = tau0;
        out.redz_prime = (new_age < p.age_universe) ? tage_to_z(p, new_age) : -1.0;
    }

    double rate;
    if (!p.use_gmr) {
        // GPF_Power_Law.__call__ components.py:556-583
        double mass_gpf = p.gpf_uses_mtot ? mstar_tot : mstar_pri;
        double rv = p.gpf[0] * pow(mass_gpf / p.gpf[1], p.gpf[2]) * pow(1.0 + redz, p.gpf[3]) *
                    pow(mstar_rat, p.gpf[4]);
        if (rv > p.gpf[5]) rv = p.gpf[5];
        rate = rv / out.gmt_time;
    } else {
        rate = gmr_eval(p, mstar_tot, mstar_rat, redz);
    }

    double dens = gsmf_eval(p, mass_gsmf, redz) * rate * dtdz_z;    // sam.py:347
    double mplaw = p.mmb[1];
    double dqbh_dqgal = mplaw * pow(mstar_rat, mplaw - 1.0);         // sam.py:355
    double dmstar_dmbh_pri = dmstar_dmbh(p, bf_tab, mstar_pri);      // sam.py:357
    double qterm = (1.0 + mstar_rat) / (1.0 + mrat);                 // sam.py:358
    double dms = dmstar_dmbh_pri * qterm;
    dens *= (mtot / mstar_tot) * (dms / dqbh_dqgal);                 // sam.py:365
    out.dens = dens;
    return out;
}

// =================================================================================================
// K1a: scipy/optimize/Zeros/brentq.c restated (scipy is an un-vendored dependency of the reference:
// `from scipy.optimize.cython_optimize cimport brentq`, sam_cyutils.pyx:14; published algorithm:
// Brent 1973, "Algorithms for Minimization without Derivatives", ch. 4, in C. Harris's formulation).
// Called as brentq(f, -20, +20, args, xtol=1e-3, rtol=1e-5, iter=100) at sam_cyutils.pyx:346-349;
// on a sign error scipy's C routine returns 0.0, on non-convergence the last iterate.
// =================================================================================================
template <class Fn>
HOLO_HD double brentq(const Fn& f, double xa, double xb, double xtol, double rtol, int iter) {
    double xpre = xa, xcur = xb;
    double xblk = 0., fpre, fcur, fblk = 0., spre = 0., scur = 0., sbis;
    double delta, stry, dpre, dblk;
    fpre = f(xpre);
    fcur = f(xcur);
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    if (signbit(fpre) == signbit(fcur)) return 0.;
    for (int i = 0; i < iter; i++) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre;
            fblk = fpre;
            spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        delta = (xtol + rtol * fabs(xcur)) / 2;
        sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);                 // secant
            } else {
                dpre = (fpre - fcur) / (xpre - xcur);                          // inverse quadratic
                dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            if (2 * fabs(stry) < fmin(fabs(spre), 3 * fabs(sbis) - delta)) {
                spre = scur; scur = stry;                                      // good short step
            } else {
                spre = sbis; scur = sbis;                                      // bisect
            }
        } else {
            spre = sbis; scur = sbis;
        }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = f(xcur);
    }
    return xcur;
}

// =================================================================================================
// K1b: per-(z,f) solve of _dynamic_binary_number_at_fobs_2pwl (sam_cyutils.pyx:654-768)
// given the per-(M,q) evolution track: sepa[s], dadt[s], frst[s] at the nsteps+1 separation edges,
// tevo[s] = cumulative evolution time at edge s (tevo[0] = 0), dt[s] = duration of step s.
// =================================================================================================

// Per-(M,q) factors of the Kepler / GW-hardening formulas, hoisted out of the per-(z,f) evaluation with the
// reference's association order kept (sam_cyutils.pyx:48, 60); x^3 and f^(2/3) are formed by
// multiplication / cbrt (<= 2 ulp from libm pow).
struct MqConsts {
    double kep_sepa_m;   // KEPLER_CONST_SEPA * pow(mtot, 1/3)
    double gw_num;       // GW_DADT_SEP_CONST * pow(mtot, 3) * mrat
    double opmr2;        // pow(1 + mrat, 2)
};

HOLO_HD MqConsts mq_consts(const CyConsts& cc, double mt, double mr) {
    MqConsts k;
    k.kep_sepa_m = cc.kepler_const_sepa * pow(mt, 1.0 / 3.0);
    k.gw_num = cc.gw_dadt_sep_const * pow(mt, 3.0) * mr;
    k.opmr2 = pow(1.0 + mr, 2.0);
    return k;
}

HOLO_HD double kepler_sepa_fast(const MqConsts& k, double freq) {
    const double cb = cbrt(freq);
    return k.kep_sepa_m / (cb * cb);
}

HOLO_HD double hard_gw_fast(const MqConsts& k, double sepa) {
    return k.gw_num / (sepa * sepa * sepa) / k.opmr2;
}

struct Track2pwl {
    const double* frst;   // (nsteps+1,)
    const double* tevo;   // (nsteps+1,)
    const double* dt;     // (nsteps,)
    int nsteps;
    const double* tage;   // cosmology tables, (n_interp,)
    const double* gz;
    const double* gdc;
    int n_interp;
    double age_universe;
};

// Observed frequency of the RIGHT edge of step `s` for a binary that formed at (gmt + age_z):
// sam_cyutils.pyx:659, 682-690, 697.
HOLO_HD double fobs_right_of_step(const Track2pwl& t, int s, double gmt, double age_z) {
    double time_right = t.tevo[s + 1] + gmt + age_z;
    int ir = bracket_increasing(t.n_interp, time_right, t.tage);
    double redz_right = interp_at_index(ir, time_right, t.tage, t.gz);
    if (redz_right < 0.0) redz_right = 0.0;
    return t.frst[s + 1] / (1.0 + redz_right);
}

// First step whose right edge reaches `ftarget` (fobs_right is non-decreasing in s): bisection.
HOLO_HD int dbn_2pwl_first_step(const Track2pwl& t, double gmt, double age_z, double ftarget) {
    int lo = 0, hi = t.nsteps;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (fobs_right_of_step(t, mid, gmt, age_z) >= ftarget) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// The same quantity for the next (higher) target frequency, walking up from the previous answer.
HOLO_HD int dbn_2pwl_next_step(const Track2pwl& t, double gmt, double age_z, double ftarget, int lo) {
    while (lo < t.nsteps && fobs_right_of_step(t, lo, gmt, age_z) < ftarget) ++lo;
    return lo;
}

// `fobs_right_of_step` for a walk over non-decreasing steps: the time at the right edge grows with the step, so the
// bracket in the age table only ever moves up -- it is advanced from `hint` (<= the true bracket, e.g. the bracket of an
// earlier step) instead of being bisected afresh.  Same index as `bracket_increasing`, hence the same value.
HOLO_HD double fobs_right_of_step_walk(const Track2pwl& t, int s, double gmt, double age_z, int& hint) {
    const double time_right = t.tevo[s + 1] + gmt + age_z;
    int ir = hint;
    const int last = t.n_interp - 2;
    while (ir < last && t.tage[ir + 1] <= time_right) ++ir;
    hint = ir;
    double redz_right = interp_at_index(ir, time_right, t.tage, t.gz);
    if (redz_right < 0.0) redz_right = 0.0;
    return t.frst[s + 1] / (1.0 + redz_right);
}

// (`fr_lo` / `fr` remember the right-edge frequency of the step the previous target stopped at: several target
//  frequencies usually fall into the same step, and the walk would evaluate that step again for each of them)
HOLO_HD int dbn_2pwl_next_step_walk(const Track2pwl& t, double gmt, double age_z, double ftarget, int lo, int& hint,
                                    int& fr_lo, double& fr) {
    while (lo < t.nsteps) {
        if (fr_lo != lo) {
            fr = fobs_right_of_step_walk(t, lo, gmt, age_z, hint);
            fr_lo = lo;
        }
        if (!(fr < ftarget)) break;
        ++lo;
    }
    return lo;
}

// Given `lo` = first step whose right edge reaches the target: returns true (and fills redz/dnum) iff some
// integration step brackets `ftarget`; when several do (exact ties at step boundaries) the LAST one wins,
// as in the reference's step-major loop order.
HOLO_HD bool dbn_2pwl_cell_from(const CyConsts& cc, const MqConsts& mq, const Track2pwl& t, double norm,
                                double rchar, double gamma_inner, double gamma_outer, double nden,
                                double gmt, double age_z, double ftarget, int lo, double* redz_out,
                                double* dnum_out) {
    bool found = false;
    const HalfPow p_outer = half_pow_spec(-gamma_outer + gamma_inner), p_inner = half_pow_spec(gamma_inner - 1.0);
    for (int s = lo; s < t.nsteps; ++s) {
        double time_right = t.tevo[s + 1] + gmt + age_z;               // pyx:659
        double time_left = time_right - t.dt[s];                       // pyx:661
        if (time_left > t.age_universe) break;                         // pyx:667 (monotone in s)
        int il = bracket_increasing(t.n_interp, time_left, t.tage);    // pyx:672
        double redz_left = interp_at_index(il, time_left, t.tage, t.gz);
        if (redz_left < 0.0) continue;                                 // pyx:678
        int ir = bracket_increasing(t.n_interp, time_right, t.tage);   // pyx:682
        double redz_right = interp_at_index(ir, time_right, t.tage, t.gz);
        if (redz_right < 0.0) redz_right = 0.0;                        // pyx:689
        double fobs_left = t.frst[s] / (1.0 + redz_left);              // pyx:696
        double fobs_right = t.frst[s + 1] / (1.0 + redz_right);
        if (ftarget < fobs_left) break;                                // later steps start even higher
        if (fobs_right < ftarget) continue;                            // pyx:709
        double new_time = interp_between_vals(ftarget, fobs_left, fobs_right, time_left, time_right);
        if (new_time > t.age_universe) continue;                       // pyx:723 (break of the f loop)
        int in = bracket_increasing(t.n_interp, new_time, t.tage);     // pyx:728
        double new_redz = interp_at_index(in, new_time, t.tage, t.gz);
        double dcom = interp_at_index(in, new_time, t.tage, t.gdc);
        double target_frst_orb = ftarget * (1.0 + new_redz);           // pyx:754
        double sepa = kepler_sepa_fast(mq, target_frst_orb);
        double dadt = hard_func_2pwl_spec(norm, sepa / rchar, p_outer, p_inner) + hard_gw_fast(mq, sepa);
        double tres = -(2.0 / 3.0) * sepa / dadt;                      // pyx:764
        const double dmpc = dcom / CY_MPC;
        double cosmo_fact = cc.four_pi_c_over_mpc * (1.0 + new_redz) * (dmpc * dmpc);
        *redz_out = new_redz;
        *dnum_out = nden * tres * cosmo_fact;                          // pyx:768
        found = true;
        // The next step can only bracket the target too if its left edge -- the same point of the track as this
        // step's right edge, recomputed: equal to a few ulp -- is <= ftarget, i.e. at an exact tie.  1e-12 is
        // three orders of magnitude above the rounding of that recomputation, so nothing is skipped wrongly.
        if (ftarget < fobs_right * (1.0 - 1.0e-12)) break;
    }
    return found;
}

HOLO_HD bool dbn_2pwl_cell(const CyConsts& cc, const MqConsts& mq, const Track2pwl& t, double norm,
                           double rchar, double gamma_inner, double gamma_outer, double nden,
                           double gmt, double age_z, double ftarget, double* redz_out,
                           double* dnum_out) {
    const int lo = dbn_2pwl_first_step(t, gmt, age_z, ftarget);
    return dbn_2pwl_cell_from(cc, mq, t, norm, rchar, gamma_inner, gamma_outer, nden, gmt, age_z, ftarget, lo,
                              redz_out, dnum_out);
}

// =================================================================================================
// K1c: one (M,q,z,f) element of _dynamic_binary_number_at_fobs_gw (sam_cyutils.pyx:853-897)
// =================================================================================================
HOLO_HD void dbn_gw_cell(const CyConsts& cc, double mt, double mr, double nden, double rzp,
                         const double* fobs, int ff, const double* gz, const double* gdc, int n_interp,
                         double* redz_out, double* dnum_out) {
    *redz_out = -1.0;
    *dnum_out = 0.0;
    if (rzp <= 0.0) return;                                            // pyx:867
    *redz_out = rzp;                                                   // pyx:871, 879-880
    double rad_isco = 3.0 * CY_SCHW * mt;
    double frst_orb_isco = kepler_freq_from_sepa(cc, mt, rad_isco);
    // the reference `break`s at the first frequency above ISCO (pyx:877-882): every later frequency
    // keeps dnum = 0 even if `fobs` were not sorted.
    for (int fp = 0; fp < ff; ++fp) {
        if (fobs[fp] * (1.0 + rzp) > frst_orb_isco) return;
    }
    double target_frst_orb = fobs[ff] * (1.0 + rzp);
    if (target_frst_orb > frst_orb_isco) return;
    int idx = bracket_decreasing(n_interp, rzp, gz);                   // pyx:885
    double dcom = interp_at_index(idx, rzp, gz, gdc);
    const MqConsts mq = mq_consts(cc, mt, mr);
    double sepa = kepler_sepa_fast(mq, target_frst_orb);
    double dadt = hard_gw_fast(mq, sepa);
    double tres = -(2.0 / 3.0) * sepa / dadt;
    const double dmpc = dcom / CY_MPC;
    double cosmo_fact = cc.four_pi_c_over_mpc * (1.0 + rzp) * (dmpc * dmpc);
    *dnum_out = nden * tres * cosmo_fact;
}

// =================================================================================================
// K2: one output bin of _integrate_differential_number_3dx1d (sam_cyutils.pyx:194-214)
// =================================================================================================
HOLO_HD double integrate_bin(const double* dnum, int64_t sM, int64_t sQ, int64_t sZ, int64_t base,
                             double dmdqdz, double dlnf) {
    double temp = 0.0;
    for (int ii = 0; ii < 2; ++ii)
        for (int jj = 0; jj < 2; ++jj)
            for (int kk = 0; kk < 2; ++kk) temp += dnum[base + ii * sM + jj * sQ + kk * sZ];
    return temp * dmdqdz * dlnf / 8.0;
}

// =================================================================================================
// K2b: cell-centre redshift and strain (gravwaves.py:694-725, single_sources.py:112-139)
// =================================================================================================

// three successive midpoint passes over axes 0,1,2 (gravwaves.py:705-708), -1 sentinels included
// `bad` (optional) reports a corner that is negative but not the -1 sentinel: the input check of
// single_sources.py:95-99, done where the values are loaded anyway.
HOLO_HD double corner_mean_redz(const double* rz, int64_t sM, int64_t sQ, int64_t sZ, int64_t base, bool* bad = nullptr) {
    const double c000 = rz[base], c100 = rz[base + sM], c001 = rz[base + sZ], c101 = rz[base + sM + sZ];
    const double c010 = rz[base + sQ], c110 = rz[base + sM + sQ], c011 = rz[base + sQ + sZ], c111 = rz[base + sM + sQ + sZ];
    if (bad) {
        *bad = (c000 < 0.0 && c000 != -1.0) || (c100 < 0.0 && c100 != -1.0) || (c001 < 0.0 && c001 != -1.0) ||
               (c101 < 0.0 && c101 != -1.0) || (c010 < 0.0 && c010 != -1.0) || (c110 < 0.0 && c110 != -1.0) ||
               (c011 < 0.0 && c011 != -1.0) || (c111 < 0.0 && c111 != -1.0);
    }
    // axis 0
    double a00 = 0.5 * (c100 + c000);
    double a01 = 0.5 * (c101 + c001);
    double a10 = 0.5 * (c110 + c010);
    double a11 = 0.5 * (c111 + c011);
    // axis 1
    double b0 = 0.5 * (a10 + a00);
    double b1 = 0.5 * (a11 + a01);
    // axis 2
    return 0.5 * (b1 + b0);
}

// utils.chirp_mass_mtmr utils.py:1978-1997
HOLO_HD double chirp_mass_mtmr(double mt, double mr) {
    return mt * pow(mr, 3.0 / 5.0) / pow(1.0 + mr, 6.0 / 5.0);
}

struct StrainOut {
    double h2fdf, zmid, dcom, sepa, angs;
};

HOLO_HD StrainOut strain_cell(const GLTable& gl, double hubble_distance, double om0,
                              double gw_src_const, double nwtg, double zc, double mc, double mt_mid,
                              double fc, double fc_over_df, bool want_params, const double* dc_tab = nullptr,
                              int dc_n = 0, double dc_inv_h = 0.0) {
    StrainOut o;
    const double inf = 1.0 / 0.0;
    const bool sel = (zc > 0.0);
    double dc = inf;
    if (sel) {
        dc = dc_tab ? comoving_distance_table(dc_tab, dc_n, dc_inv_h, hubble_distance, zc) : -1.0;
        if (dc < 0.0) dc = comoving_distance_cm(gl, hubble_distance, om0, zc);      // no table, or beyond it
    }
    double h2 = 0.0;
    if (sel) {
        const double fr = fc * (1.0 + zc);                               // utils.frst_from_fobs
        // utils.gw_strain_source utils.py:2260-2285; x^(2/3) as cbrt(x)^2 (|diff| to pow <= 2 ulp)
        const double cb = cbrt(2.0 * mc * fr);
        const double hs = gw_src_const * mc * (cb * cb) / dc;
        h2 = (hs * hs) * fc_over_df;                                     // gravwaves.py:723
    }
    o.h2fdf = h2;                                                        // z <= 0: d_c = inf -> hs = 0
    o.zmid = zc; o.dcom = dc; o.sepa = 0.0; o.angs = 0.0;
    if (want_params) {
        // single_sources.py:124-139
        const double rz = sel ? zc : -1.0;
        const double frp = fc * (1.0 + rz);
        const double two_pi_f = 2.0 * CY_PI * frp;
        const double sepa = cbrt(nwtg * mt_mid / (two_pi_f * two_pi_f));   // utils.py:1705-1724 (1/0 -> inf for rz = -1)
        const double dang = dc / (1.0 + rz);                             // utils.py:1897-1917
        o.zmid = rz; o.sepa = sepa; o.angs = sepa / dang;
    }
    return o;
}

}  // namespace holo
