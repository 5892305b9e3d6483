#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the ORACLE (run in the build container).

Sources of truth:
  * the compiled reference (oracle/_ref: holodeck's own cyutils.pyx / sam_cyutils.pyx) for every
    quantity the reference computes natively (norm, dbn, integrate, sam_poisson_gwb, loudest_*, ss_bg_*,
    eccentric GWB), with its RNG seeded (oracle/build_ref.py) for the realised ones;
  * oracle/glue.py (numpy restatement of the reference's Python glue, independent quadrature
    cosmology) for density, strain and the params arrays.

Usage:  python tests/golden/make_golden.py        (writes tests/golden/*.npz; small, committed)
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import glue as G   # noqa: E402

OUT = Path(__file__).resolve().parent


def cosmo_tables(oc, size=200):
    """The interpolation tables handed to the kernels (same construction as holodeck_b200.cosmology,
    values from the oracle's quadrature)."""
    z_pnts = [1000.0, 10.0, 4.0, 2.0, 1.0, 0.5, 0.1, 0.01]
    num = size // len(z_pnts)
    z0 = z_pnts[0]
    segs = []
    for z1 in z_pnts[1:]:
        segs.append(np.logspace(*np.log10([z0, z1]), num=num, endpoint=False))
        z0 = z1
    segs.append(np.linspace(z0, 0.0, num=num))
    zg = np.concatenate(segs)
    return zg, oc.comoving_distance(zg), oc.age(zg)


def sam_case(name, shape, nfreq, pta_dur_yr, kind, seed, nreals, nloud, hard='2pwl'):
    pp = dict(G.PS_CLASSIC_DEFAULTS)
    oc = G.OracleCosmo()
    M, Q, Z = shape
    mtot = np.logspace(*np.log10([1.0e4*G.MSOL, 1.0e12*G.MSOL]), M)
    mrat = np.logspace(*np.log10([1e-3, 1.0]), Q)
    redz = np.logspace(*np.log10([1e-3, 10.0]), Z)
    out = dict(mtot=mtot, mrat=mrat, redz=redz, kind=kind, hard=hard)

    if kind == 'classic':   # GSMF_Schechter + GPF + GMT, PS_Classic defaults
        mmb = G.MMBulge('KH2013', mamp_log10=pp['mmb_mamp_log10'], mplaw=pp['mmb_plaw'], scatter_dex=0.0)
        gsmf = lambda m, z: G.gsmf_schechter(m, z, phi0=pp['gsmf_phi0_log10'], phiz=pp['gsmf_phiz'], mchar0_log10=pp['gsmf_mchar0_log10'], mcharz=pp['gsmf_mcharz'], alpha0=pp['gsmf_alpha0'], alphaz=pp['gsmf_alphaz'])   # noqa
        gpf = lambda m, q, z: G.gpf_power_law(m, q, z, frac_norm_allq=pp['gpf_frac_norm_allq'], malpha=pp['gpf_malpha'], qgamma=pp['gpf_qgamma'], zbeta=pp['gpf_zbeta'], max_frac=pp['gpf_max_frac'])   # noqa
        gmt = lambda m, q, z: G.gmt_power_law(m, q, z, oc.h, time_norm=pp['gmt_norm']*G.GYR, malpha=pp['gmt_malpha'], qgamma=pp['gmt_qgamma'], zbeta=pp['gmt_zbeta'])   # noqa
        dd = G.static_binary_density(mtot, mrat, redz, oc, gsmf, mmb, gpf=gpf, gmt=gmt, scatter=False)
    elif kind == 'default':   # Semi_Analytic_Model() defaults: GSMF_Schechter + GMR_Illustris, KH2013, no GMT
        mmb = G.MMBulge('KH2013', scatter_dex=0.0)
        dd = G.static_binary_density(mtot, mrat, redz, oc, G.gsmf_schechter, mmb, gmr=G.gmr_illustris, scatter=False)
    elif kind == 'double':   # GSMF_Double_Schechter + GPF + GMT defaults, MM2013
        mmb = G.MMBulge('MM2013', scatter_dex=0.0)
        gmt = lambda m, q, z: G.gmt_power_law(m, q, z, oc.h)   # noqa
        dd = G.static_binary_density(mtot, mrat, redz, oc, G.gsmf_double_schechter, mmb, gpf=G.gpf_power_law, gmt=gmt, scatter=False)
    else:
        raise ValueError(kind)
    out['dens'] = dd['dens']
    if dd['gmt_time'] is not None:
        out['gmt_time'] = dd['gmt_time']
        out['redz_prime'] = dd['redz_prime']

    zg, dcg, ageg = cosmo_tables(oc)
    out.update(grid_z=zg, grid_dcom=dcg, grid_age=ageg)
    fobs_cents, fobs_edges = G.pta_freqs(pta_dur_yr*G.YR, nfreq)
    out.update(fobs_cents=fobs_cents, fobs_edges=fobs_edges)
    fo_c = fobs_cents / 2.0
    fo_e = fobs_edges / 2.0

    sam = G.StubSam(mtot, mrat, redz, dd['dens'], dd['gmt_time'], dd['redz_prime'])
    tabs = G.StubCosmoTables(zg, dcg, ageg)
    if hard == '2pwl':
        hp = dict(time=pp['hard_time']*G.GYR, sepa_init=pp['hard_sepa_init']*G.PC, rchar=pp['hard_rchar']*G.PC,
                  gamma_inner=pp['hard_gamma_inner'], gamma_outer=pp['hard_gamma_outer'], nsteps=300)
        norm_log10 = G.ref_find_norm(hp['time'], mtot, mrat, hp['sepa_init'], hp['rchar'], hp['gamma_inner'], hp['gamma_outer'], hp['nsteps'])
        out['norm_log10'] = norm_log10
        out['hard_params'] = np.array([hp['time'], hp['sepa_init'], hp['rchar'], hp['gamma_inner'], hp['gamma_outer'], hp['nsteps']])
        _, scy, _ = G.ref()
        nsel = [0, norm_log10.size // 2, norm_log10.size - 1]
        mt2, mr2 = np.meshgrid(mtot, mrat, indexing='ij')
        out['lifetime_idx'] = np.array(nsel)
        out['lifetime'] = np.array([scy.integrate_binary_evolution_2pwl(norm_log10.flat[ii], mt2.flat[ii], mr2.flat[ii], hp['sepa_init'], hp['rchar'], hp['gamma_inner'], hp['gamma_outer'], hp['nsteps']) for ii in nsel])
        rz, dn = G.ref_dbn(fo_c, sam, tabs, '2pwl', 10.0**norm_log10, hp['sepa_init'], hp['rchar'], hp['gamma_inner'], hp['gamma_outer'], hp['nsteps'])
    else:
        rz, dn = G.ref_dbn(fo_c, sam, tabs, 'gw')
    rz = np.asarray(rz)
    dn = np.asarray(dn)
    out.update(redz_final=rz, diff_num=dn)
    edges = [mtot, mrat, redz, fo_e]
    number = np.asarray(G.ref_integrate(edges, dn))
    out['number'] = number
    h2fdf = G.char_strain_sq_from_bin_edges_redz(edges, rz, oc.comoving_distance)
    out['h2fdf'] = h2fdf
    out['h2fdf_noredz'] = G.char_strain_sq_from_bin_edges(edges, oc.comoving_distance)
    zf, dcf, sep, ang = G.ss_params_arrays(edges, rz, oc.comoving_distance)
    out.update(par_redz=zf, par_dcom=dcf, par_sepa=sep, par_angs=ang)
    out['hc2_expect'] = np.sum(h2fdf * number, axis=(0, 1, 2))

    # ---- realised quantities: the seeded reference + its draws in supplied-count layout
    cy, _, _ = G.ref()
    order, msort, qsort, zsort = G.rank_order(h2fdf, 'stable')
    out.update(order=order.astype(np.int32), seed=seed, nreals=nreals, nloud=nloud)
    mt_c, mr_c, rz_c = G.midpoints(mtot), G.midpoints(mrat), G.midpoints(redz)

    cy.ORACLE_SEED = seed
    out['gwb_ref'] = np.asarray(cy.sam_poisson_gwb(number, h2fdf, nreals))
    out['counts_gwb'] = G.counts_sam_poisson_gwb(number, nreals, seed)

    cnt = G.counts_loudest(number, order, nreals, seed)
    out['counts_loud'] = cnt
    cy.ORACLE_SEED = seed
    a, b = cy.loudest_hc_from_sorted(number, h2fdf, nreals, nloud, msort, qsort, zsort)
    out.update(l1_hc2ss=np.asarray(a), l1_hc2bg=np.asarray(b))
    cy.ORACLE_SEED = seed
    a, b, c, d, e = cy.loudest_hc_and_par_from_sorted(number, h2fdf, nreals, nloud, mt_c, mr_c, rz_c, msort, qsort, zsort)
    out.update(l2_hc2ss=np.asarray(a), l2_hc2bg=np.asarray(b), l2_lspar=np.asarray(c), l2_bgpar=np.asarray(d), l2_ssidx=np.asarray(e))
    cy.ORACLE_SEED = seed
    a, b, c, d = cy.loudest_hc_and_par_from_sorted_redz(number, h2fdf, nreals, nloud, mt_c, mr_c, rz_c, zf, dcf, sep, ang, msort, qsort, zsort)
    out.update(l3_hc2ss=np.asarray(a), l3_hc2bg=np.asarray(b), l3_sspar=np.asarray(c), l3_bgpar=np.asarray(d))

    out['counts_ssbg'] = G.counts_ss_bg(number, nreals, seed)
    cy.ORACLE_SEED = seed
    a, b, c = cy.ss_bg_hc(number, h2fdf, nreals)
    out.update(s1_hc2ss=np.asarray(a), s1_hc2bg=np.asarray(b), s1_ssidx=np.asarray(c))
    cy.ORACLE_SEED = seed
    a, b, c, d, e = cy.ss_bg_hc_and_par(number, h2fdf, nreals, mt_c, mr_c, rz_c)
    out.update(s2_hc2ss=np.asarray(a), s2_hc2bg=np.asarray(b), s2_ssidx=np.asarray(c), s2_bgpar=np.asarray(d), s2_sspar=np.asarray(e))
    cy.ORACLE_SEED = None

    fname = OUT / f"{name}.npz"
    np.savez_compressed(fname, **out)
    print(f"{fname.name}: {fname.stat().st_size/1e6:.2f} MB; number>0: {(number>0).mean():.2f}, max N {number.max():.2e}, "
          f"occupied draws/realization ~ {np.minimum(number,1).sum():.0f}")




def eccen_case(name="eccen_small"):
    """Eccentric GWB (cyutils.sam_calc_gwb_single_eccen[_discrete]) on a small synthetic SAM."""
    cy, _, _ = G.ref()
    M, Q, Z, F, H, E = 9, 8, 10, 5, 20, 123
    mtot = np.logspace(np.log10(1e7*G.MSOL), np.log10(1e11*G.MSOL), M)
    mrat = np.logspace(-2, 0, Q)
    redz = np.logspace(-2, 0.7, Z)
    rng = np.random.default_rng(1)
    ndens = rng.uniform(0, 1e-3, (M, Q, Z))
    ndens[rng.uniform(size=ndens.shape) < 0.2] = 0.0
    oc = G.OracleCosmo()
    dcom = oc.comoving_distance(redz) / G.MPC
    fobs, _ = G.pta_freqs(16.03*G.YR, F)
    out = dict(ndens=ndens, mtot=mtot, mrat=mrat, redz=redz, dcom=dcom, fobs=fobs, nharms=H)
    for tag, a0 in (("a", 0.05*G.PC), ("b", 10.0*G.PC)):
        sepa, ecc = G.evolve_eccen_uniform_single(mtot, 0.95, a0, E)
        out[f"sepa_{tag}"] = sepa
        out[f"eccen_{tag}"] = ecc
        out[f"gwb_{tag}"] = np.asarray(cy.sam_calc_gwb_single_eccen(ndens, np.log10(mtot), mrat, redz, dcom, fobs, sepa, ecc, H))
    # discrete: a seeded reference run; only its moments / quantiles are compared (different RNG)
    cy.ORACLE_SEED = 31415
    R = 400
    scale = 3e7   # boost the densities so that the Poisson numbers are O(1-100), not all zero
    out["disc_scale"] = scale
    out["disc_R"] = R
    out["gwb_disc_a"] = np.asarray(cy.sam_calc_gwb_single_eccen_discrete(
        ndens*scale, np.log10(mtot), mrat, redz, dcom, fobs, out["sepa_a"], out["eccen_a"], H, R))
    cy.ORACLE_SEED = None
    fname = OUT / f"{name}.npz"
    np.savez_compressed(fname, **out)
    print(f"{fname.name}: {fname.stat().st_size/1e6:.2f} MB")


def eccen_case_h100(name="eccen_h100"):
    """BASELINE configs[3] harmonic count: `sam_calc_gwb_single_eccen` at H = 100 on a 21^3 grid, 8 frequencies
    (11 s of the compiled reference per track)."""
    cy, _, _ = G.ref()
    M, Q, Z, F, H, E = 21, 21, 21, 8, 100, 123
    mtot = np.logspace(np.log10(1e6*G.MSOL), np.log10(1e11*G.MSOL), M)
    mrat = np.logspace(-2, 0, Q)
    redz = np.logspace(-2, 0.7, Z)
    rng = np.random.default_rng(11)
    ndens = rng.uniform(0, 1e-3, (M, Q, Z))
    ndens[rng.uniform(size=ndens.shape) < 0.2] = 0.0
    oc = G.OracleCosmo()
    dcom = oc.comoving_distance(redz) / G.MPC
    fobs, _ = G.pta_freqs(16.03*G.YR, F)
    out = dict(ndens=ndens, mtot=mtot, mrat=mrat, redz=redz, dcom=dcom, fobs=fobs, nharms=H)
    for tag, a0 in (("a", 0.05*G.PC), ("b", 10.0*G.PC)):
        sepa, ecc = G.evolve_eccen_uniform_single(mtot, 0.95, a0, E)
        out[f"sepa_{tag}"] = sepa
        out[f"eccen_{tag}"] = ecc
        out[f"gwb_{tag}"] = np.asarray(cy.sam_calc_gwb_single_eccen(ndens, np.log10(mtot), mrat, redz, dcom, fobs, sepa, ecc, H))
    fname = OUT / f"{name}.npz"
    np.savez_compressed(fname, **out)
    print(f"{fname.name}: {fname.stat().st_size/1e6:.2f} MB")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "eccen_h100":
        eccen_case_h100()
        sys.exit(0)
    sam_case("classic_2pwl", (13, 11, 15), 6, 16.03, 'classic', seed=12345, nreals=6, nloud=3)
    sam_case("default_gw", (10, 11, 12), 5, 10.0, 'default', seed=777, nreals=5, nloud=2, hard='gw')
    sam_case("double_2pwl", (9, 8, 10), 4, 16.03, 'double', seed=99, nreals=4, nloud=5)
    eccen_case()
    eccen_case_h100()
