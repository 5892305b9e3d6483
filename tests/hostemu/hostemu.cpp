// hostemu.cpp -- TEST-ONLY host build of the per-element kernel math in holodeck_b200/csrc/holo_math.cuh.
//
// The CUDA kernels are thin parallel wrappers around HOLO_HD inline functions; this file drives
// the same functions with plain loops so that the arithmetic (not the launch machinery) can be
// compared against the reference on a machine without a GPU.  It is never loaded by the product
// (holodeck_b200/_lib.py only ever opens libholo_b200.so) and it is not an oracle: the oracle is
// oracle/_ref (the compiled reference) plus oracle/*.py|c.
//
// build: g++ -O2 -ffp-contract=off -shared -fPIC -o tests/hostemu/libhostemu.so tests/hostemu/hostemu.cpp
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../holodeck_b200/csrc/holo_math.cuh"

using namespace holo;

static CyConsts to_cc(const holo_cy_consts& c) {
    CyConsts cc;
    cc.gw_dadt_sep_const = c.gw_dadt_sep_const;
    cc.kepler_const_freq = c.kepler_const_freq;
    cc.kepler_const_sepa = c.kepler_const_sepa;
    cc.four_pi_c_over_mpc = c.four_pi_c_over_mpc;
    return cc;
}

extern "C" {

void emu_sam_density(const double* mtot, const double* mrat, const double* redz, const double* age_z,
                     const double* dtdz_z, int M, int Q, int Z, const holo_sam_params* par, const double* bf_tables,
                     double* dens, double* gmt_time, double* redz_prime) {
    for (int ii = 0; ii < M; ++ii)
        for (int jj = 0; jj < Q; ++jj)
            for (int kk = 0; kk < Z; ++kk) {
                int64_t i = ((int64_t)ii * Q + jj) * Z + kk;
                DensityOut o = density_point(*par, bf_tables, mtot[ii], mrat[jj], redz[kk], age_z[kk], dtdz_z[kk]);
                dens[i] = o.dens;
                if (gmt_time) gmt_time[i] = o.gmt_time;
                if (redz_prime) redz_prime[i] = o.redz_prime;
            }
}

struct HostLifetime {
    const double *sepa, *p1, *p2, *gw;
    int nsteps;
    double target;
    double operator()(double norm_log10) const {
        double norm = pow(10.0, norm_log10);
        double part[32];
        for (int lane = 0; lane < 32; ++lane) {
            double acc = 0.0;
            for (int s = lane; s < nsteps; s += 32) {
                double dl = (-norm * p1[s]) / p2[s] + gw[s];
                double dr = (-norm * p1[s + 1]) / p2[s + 1] + gw[s + 1];
                acc += 2.0 * (sepa[s + 1] - sepa[s]) / (dl + dr);
            }
            part[lane] = acc;
        }
        for (int off = 16; off > 0; off >>= 1) {
            double nxt[32];
            for (int l = 0; l < 32; ++l) nxt[l] = part[l] + part[l ^ off];
            memcpy(part, nxt, sizeof(part));
        }
        return part[0] - target;
    }
};

void emu_norm_2pwl(holo_cy_consts c, double target_time, const double* mtot, const double* mrat, int N,
                   double sepa_init, double rchar, double gi, double go, int nsteps, int mode,
                   const double* norm_in, double* out) {
    CyConsts cc = to_cc(c);
    int ne = nsteps + 1;
    std::vector<double> sepa(ne), p1(ne), p2(ne), gw(ne);
    double sepa_init_log10 = log10(sepa_init);
    for (int i = 0; i < N; ++i) {
        double mt = mtot[i], mr = mrat[i];
        double dx = (sepa_init_log10 - log10(3.0 * CY_SCHW * mt)) / nsteps;
        double slog = sepa_init_log10;
        for (int k = 0; k < ne; ++k) {
            double sp = pow(10.0, slog);
            double xx = sp / rchar;
            sepa[k] = sp;
            p1[k] = pow(1.0 + xx, -go + gi);
            p2[k] = pow(xx, gi - 1.0);
            gw[k] = hard_gw(cc, mt, mr, sp);
            slog -= dx;
        }
        HostLifetime fn{sepa.data(), p1.data(), p2.data(), gw.data(), nsteps, mode == 0 ? target_time : 0.0};
        out[i] = (mode == 0) ? brentq(fn, -20.0, 20.0, 1e-3, 1e-5, 100) : fn(norm_in[i]);
    }
}

void emu_dbn_2pwl(holo_cy_consts c, const double* fobs, int F, double sepa_init, int nsteps,
                  const double* hard_norm, double rchar, double gi, double go, const double* nden,
                  const double* mtot, const double* mrat, const double* redz, const double* gmt_time,
                  int M, int Q, int Z, const double* grid_z, const double* grid_dcom,
                  const double* grid_age, int n_interp, double* redz_final, double* diff_num) {
    CyConsts cc = to_cc(c);
    int ne = nsteps + 1;
    std::vector<double> sepa(ne), dadt(ne), frst(ne), tevo(ne), dt(ne), zage(Z);
    double sepa_init_log10 = log10(sepa_init);
    for (int k = 0; k < Z; ++k) {
        int idx = bracket_decreasing(n_interp, redz[k], grid_z);
        zage[k] = interp_at_index(idx, redz[k], grid_z, grid_age);
    }
    for (int ii = 0; ii < M; ++ii)
        for (int jj = 0; jj < Q; ++jj) {
            int mq = ii * Q + jj;
            double mt = mtot[ii], mr = mrat[jj], norm = hard_norm[mq];
            double dx = (sepa_init_log10 - log10(3.0 * CY_SCHW * mt)) / nsteps;
            double slog = sepa_init_log10;
            for (int k = 0; k < ne; ++k) {
                double sp = pow(10.0, slog);
                sepa[k] = sp;
                dadt[k] = hard_func_2pwl_gw(cc, mt, mr, sp, norm, rchar, gi, go);
                frst[k] = kepler_freq_from_sepa(cc, mt, sp);
                slog -= dx;
            }
            tevo[0] = 0.0;
            double te = 0.0;
            for (int k = 0; k < nsteps; ++k) {
                dt[k] = 2.0 * (sepa[k + 1] - sepa[k]) / (dadt[k] + dadt[k + 1]);
                te += dt[k];
                tevo[k + 1] = te;
            }
            const MqConsts mqc = mq_consts(cc, mt, mr);
            Track2pwl t;
            t.frst = frst.data(); t.tevo = tevo.data(); t.dt = dt.data(); t.nsteps = nsteps;
            t.tage = grid_age; t.gz = grid_z; t.gdc = grid_dcom; t.n_interp = n_interp;
            t.age_universe = grid_age[n_interp - 1];
            for (int kk = 0; kk < Z; ++kk)
                for (int ff = 0; ff < F; ++ff) {
                    double rz = -1.0, dn = 0.0;
                    int64_t b = (int64_t)mq * Z + kk;
                    dbn_2pwl_cell(cc, mqc, t, norm, rchar, gi, go, nden[b], gmt_time[b], zage[kk],
                                  fobs[ff], &rz, &dn);
                    redz_final[b * F + ff] = rz;
                    diff_num[b * F + ff] = dn;
                }
        }
}

void emu_dbn_gw(holo_cy_consts c, const double* fobs, int F, const double* nden, const double* mtot,
                const double* mrat, const double* redz_prime, int M, int Q, int Z, const double* grid_z,
                const double* grid_dcom, int n_interp, double* redz_final, double* diff_num) {
    CyConsts cc = to_cc(c);
    for (int ii = 0; ii < M; ++ii)
        for (int jj = 0; jj < Q; ++jj)
            for (int kk = 0; kk < Z; ++kk)
                for (int ff = 0; ff < F; ++ff) {
                    int64_t b = ((int64_t)ii * Q + jj) * Z + kk;
                    dbn_gw_cell(cc, mtot[ii], mrat[jj], nden[b], redz_prime[b], fobs, ff, grid_z,
                                grid_dcom, n_interp, &redz_final[b * F + ff], &diff_num[b * F + ff]);
                }
}

void emu_integrate_and_strain(const holo_cosmo_params* cosmo, double gw_src_const, double nwtg,
                              const double* log10_mtot, const double* mrat, const double* redz,
                              const double* dln_freq, const double* dnum, const double* redz_final,
                              const double* rz_mid, const double* mt_mid, const double* mr_mid,
                              const double* fc, const double* fc_over_df, int M, int Q, int Z, int F,
                              double* numb, double* h2fdf, double* zmid, double* dcom, double* sepa,
                              double* angs) {
    GLTable gl;
    for (int i = 0; i < GL_ORDER; ++i) { gl.x[i] = cosmo->gl_x[i]; gl.w[i] = cosmo->gl_w[i]; }
    int Mb = M - 1, Qb = Q - 1, Zb = Z - 1;
    int64_t sZ = F, sQ = (int64_t)Z * F, sM = (int64_t)Q * Z * F;
    bool want_par = zmid || dcom || sepa || angs;
    for (int mm = 0; mm < Mb; ++mm)
        for (int qq = 0; qq < Qb; ++qq) {
            double mc = chirp_mass_mtmr(mt_mid[mm], mr_mid[qq]);
            for (int zz = 0; zz < Zb; ++zz)
                for (int ff = 0; ff < F; ++ff) {
                    int64_t i = (((int64_t)mm * Qb + qq) * Zb + zz) * F + ff;
                    int64_t base = mm * sM + qq * sQ + zz * sZ + ff;
                    if (numb) {
                        double dm = log10_mtot[mm + 1] - log10_mtot[mm];
                        double dmdq = dm * (mrat[qq + 1] - mrat[qq]);
                        double dmdqdz = dmdq * (redz[zz + 1] - redz[zz]);
                        numb[i] = integrate_bin(dnum, sM, sQ, sZ, base, dmdqdz, dln_freq[ff]);
                    }
                    if (h2fdf) {
                        double zc = redz_final ? corner_mean_redz(redz_final, sM, sQ, sZ, base) : rz_mid[zz];
                        StrainOut o = strain_cell(gl, cosmo->hubble_distance, cosmo->om0, gw_src_const,
                                                  nwtg, zc, mc, mt_mid[mm], fc[ff], fc_over_df[ff], want_par);
                        h2fdf[i] = o.h2fdf;
                        if (zmid) zmid[i] = o.zmid;
                        if (dcom) dcom[i] = o.dcom;
                        if (sepa) sepa[i] = o.sepa;
                        if (angs) angs[i] = o.angs;
                    }
                }
        }
}

}  // extern "C"

// ---- samplers (holo_rng.cuh) ---------------------------------------------------------------------
#include "../../holodeck_b200/csrc/holo_rng.cuh"

extern "C" {

// n independent draws of Poisson(lam) (realization index = 0..n-1) through the stand-alone path
void emu_draw_elements(double lam, int64_t n, uint64_t seed, double thresh, uint64_t idx, double* out) {
    DrawKey key;
    key.k0 = (uint32_t)seed; key.k1 = (uint32_t)(seed >> 32); key.stream = 7;
    for (int64_t i = 0; i < n; ++i) {
        key.real = (uint32_t)i;
        out[i] = draw_element(lam, thresh, key, idx);
    }
}

// n draws through the shared-group path used by the realization kernel: 4 frequencies of one cell
// share a HI block (and a LO block for the SMALL class); lam4 holds the 4 expectation values.
void emu_draw_group(const double* lam4, int64_t n, uint64_t seed, double thresh, uint32_t cell, double* out /* (n,4) */) {
    DrawKey key;
    key.k0 = (uint32_t)seed; key.k1 = (uint32_t)(seed >> 32); key.stream = 2;
    FPrep p[4];
    int cls[4];
    for (int j = 0; j < 4; ++j) cls[j] = prep_draw(lam4[j], thresh, p[j]);
    for (int64_t i = 0; i < n; ++i) {
        key.real = (uint32_t)i;
        Philox4 hi = group_bits(key, cell, 3, PURPOSE_GROUP_HI);
        Philox4 lo = group_bits(key, cell, 3, PURPOSE_GROUP_LO);
        for (int j = 0; j < 4; ++j) {
            uint64_t idx = (uint64_t)cell * 40 + 12 + j;
            double v = 0.0;
            if (cls[j] == CLS_TINY) v = draw_tiny(p[j], hi.v[j], key, idx);
            else if (cls[j] == CLS_SMALL) v = draw_small(p[j], hi.v[j], lo.v[j]);
            else if (cls[j] == CLS_PTRS) v = draw_ptrs(p[j], key, idx);
            else if (cls[j] == CLS_NORMAL) v = draw_normal(p[j], key, idx);
            out[i * 4 + j] = v;
        }
    }
}

// lane-by-lane emulation of build_table_sub<SW> (holo_realize.cu; SW = 8 lanes per element table, 32 for the table of
// a superposition group): thresholds at t[0..W), sentinels at t[-1], t[W]
static void emu_build_table(double lam, uint32_t* t, int kmin, int W, int SW = 8) {
    const int seg = (W + SW - 1) / SW;
    const double ln_lam = log(lam), inv_lam = 1.0 / lam;
    double incl[32], ptop[32];
    int j0[32], j1[32];
    for (int l = 0; l < SW; ++l) {
        j0[l] = l * seg > W ? W : l * seg;
        j1[l] = j0[l] + seg > W ? W : j0[l] + seg;
        incl[l] = table_segment_mass(lam, ln_lam, inv_lam, kmin, j0[l], j1[l], &ptop[l]);
    }
    for (int off = 1; off < SW; off <<= 1) {
        double prev[32];
        for (int l = 0; l < SW; ++l) prev[l] = incl[l];
        for (int l = off; l < SW; ++l) incl[l] = prev[l] + prev[l - off];
    }
    for (int l = 0; l < SW; ++l) table_segment_write(t, incl[l], ptop[l], inv_lam, kmin, j0[l], j1[l]);
    t[-1] = 0u;
    t[W] = 0xFFFFFFFFu;
}

// The TABLE class of the realization kernel: the 32-lane table build is emulated lane by lane (same
// segment arithmetic, same Hillis-Steele scan order), then n draws go through the fast path / slow path.
// mode 0: as the kernel does; mode 1: every draw through the exact slow path (table_resolve) -- both must
// be exact Poisson samplers.  `tab_out` (optional, W entries) receives the thresholds; returns W, kmin.
int emu_draw_table(double lam, int64_t n, uint64_t seed, int mode, double* out, uint32_t* tab_out, int* kmin_out,
                   int64_t* nslow_out) {
    const TableSpec ts = table_spec(lam);
    const int W = ts.W, kmin = ts.kmin;
    std::vector<uint32_t> tabv(W + 2);
    uint32_t* tab = tabv.data() + 1;
    int lg = 0;
    while ((2 << lg) <= W) ++lg;
    emu_build_table(lam, tab, kmin, W);
    if (tab_out) for (int j = 0; j < W; ++j) tab_out[j] = tab[j];
    if (kmin_out) *kmin_out = kmin;
    DrawKey key;
    key.k0 = (uint32_t)seed; key.k1 = (uint32_t)(seed >> 32); key.stream = 2;
    int64_t nslow = 0;
    for (int64_t i = 0; i < n; ++i) {
        key.real = (uint32_t)i;
        const uint32_t cell = 77u + (uint32_t)(i >> 32);
        Philox4 hi = group_bits(key, cell, 3, PURPOSE_GROUP_HI);
        int nidx;
        double v = draw_table_ladder(tabv.data(), 1u, kmin, W, lg, hi.v[1], &nidx);
        {
            int n2;
            const double v2 = draw_table_fast(tab, kmin, W, hi.v[1], &n2);   // the loop form agrees with the ladder
            if (n2 != nidx || v2 != v) v = -7.0;
        }
        if (v < 0.0 || mode == 1) {
            if (v != -7.0) v = table_resolve_keyed(lam, tab, kmin, W, nidx, hi.v[1], key, cell, 3, 1);
            ++nslow;
        }
        out[i] = v;
    }
    if (nslow_out) *nslow_out = nslow;
    return W;
}

// n realizations of a superposition group with member expectations lam[0..K): out is (n, K) counts
void emu_draw_group_members(const double* lam, int K, int64_t n, uint64_t seed, double* out) {
    std::vector<double> gcum(K);
    double c = 0.0;
    for (int k = 0; k < K; ++k) { c += lam[k]; gcum[k] = c; }
    const TableSpec ts = table_spec(c);
    std::vector<uint32_t> tab(ts.W + 2);
    emu_build_table(c, tab.data() + 1, ts.kmin, ts.W, 32);
    int lg = 0;
    while ((2 << lg) <= ts.W) ++lg;
    DrawKey key;
    key.k0 = (uint32_t)seed; key.k1 = (uint32_t)(seed >> 32); key.stream = 1;
    for (int64_t i = 0; i < n; ++i) {
        key.real = (uint32_t)i;
        double* row = out + i * K;
        for (int k = 0; k < K; ++k) row[k] = 0.0;
        draw_group(tab.data(), 1u, ts.kmin, ts.W, lg, c, gcum.data(), K, 4242u, 3u, key, [&](int member) { row[member] += 1.0; });
    }
}

void emu_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    Philox4 r = philox4x32_10(c0, c1, c2, c3, k0, k1);
    for (int i = 0; i < 4; ++i) out[i] = r.v[i];
}

}  // extern "C"

// ---- eccentric GWB helpers (holo_eccen_math.cuh) -----------------------------------------------
#include "../../holodeck_b200/csrc/holo_eccen_math.cuh"

extern "C" {

void emu_bessel_pair(int nn, double x, double* out2) { bessel_pair(nn, x, &out2[0], &out2[1]); }

double emu_gw_freq_dist_func(int nn, double ee) { return gw_freq_dist_func(nn, ee); }

// continuous eccentric GWB, same factorisation as the CUDA kernels (q-sum, then (f,n) sum over (z,M))
void emu_eccen_gwb(double gw_dadt_sep_const, double gw_src_const, const double* ndens, const double* mtot_log10,
                   const double* mrat, const double* redz, const double* dcom, const double* gwfobs,
                   const double* sepa_evo, const double* eccen_evo, int M, int Q, int Z, int F, int E, int H,
                   double* gwb) {
    EccGeom g{ndens, mtot_log10, mrat, redz, dcom, gwfobs, sepa_evo, eccen_evo, M, Q, Z, F, E, H};
    EccConsts cc{gw_dadt_sep_const, gw_src_const};
    std::vector<double> frst_pref(E), bsum((size_t)Z * M);
    for (int i = 0; i < E; ++i) frst_pref[i] = ((1.0 / (2.0 * CY_PI)) * sqrt(CY_NWTG)) / pow(sepa_evo[i], 1.5);
    const double four_pi_c_mpc = 4 * CY_PI * (CY_SPLC / CY_MPC);
    for (int kk = 0; kk < Z; ++kk)
        for (int ii = 0; ii < M; ++ii) {
            double kw, kdx, iw, idx_;
            trapz_grid_weight(kk, Z, redz, &kw, &kdx);
            trapz_grid_weight(ii, M, mtot_log10, &iw, &idx_);
            const double zterm = 1.0 + redz[kk], dc_mpc = dcom[kk], dc_cm = dc_mpc * CY_MPC;
            const double dc_term = four_pi_c_mpc * pow(dc_mpc, 2.0);
            const double mt = pow(10.0, mtot_log10[ii]);
            const double weight_ik = idx_ * kdx / (iw * kw);
            double acc = 0.0;
            for (int jj = 0; jj < Q; ++jj) {
                double jw, jdx;
                trapz_grid_weight(jj, Q, mrat, &jw, &jdx);
                const double weight = weight_ik * (jdx / jw);
                const double q = mrat[jj], m1 = mt / (1.0 + q), m2 = mt - m1;
                const double mchirp = mt * pow(q, 3.0 / 5.0) / pow(1 + q, 6.0 / 5.0);
                double hp = ndens[((int64_t)ii * Q + jj) * Z + kk] * dc_term * zterm;
                hp *= pow(gw_src_const * mchirp * pow(2.0 * mchirp, 2.0 / 3.0) / dc_cm, 2.0);
                acc += weight * hp / (m1 * m2);
            }
            bsum[(size_t)kk * M + ii] = acc;
        }
    for (int ff = 0; ff < F; ++ff)
        for (int nh = 1; nh <= H; ++nh) {
            double acc = 0.0;
            for (int p = 0; p < Z * M; ++p) {
                if (bsum[p] == 0.0) continue;
                double afac, tf, hf;
                if (eccen_factor(g, cc, frst_pref.data(), p / M, p % M, ff, nh, &afac, &tf, &hf)) acc += afac * bsum[p];
            }
            gwb[(size_t)ff * H + nh - 1] = acc;
        }
}

}  // extern "C"
