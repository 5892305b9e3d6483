"""The reference's CPU path for ``sam.gwb``, end to end -- TEST / BASELINE INFRASTRUCTURE ONLY.

``reference_chain`` strings together, in the order of ``holodeck/sams/sam.py:872-947``:

    static_binary_density (numpy restatement, oracle/glue.py)  ->  Fixed_Time_2PL_SAM norm (compiled
    reference ``find_2pwl_hardening_norm``)  ->  ``dynamic_binary_number_at_fobs`` (compiled reference)
    ->  ``integrate_differential_number_3dx1d`` (compiled reference)  ->  ``char_strain_sq_from_bin_edges_redz``
    + argsort (numpy restatement)  ->  ``loudest_hc_from_sorted`` (compiled reference).

It is what ``bench.py`` times as ``cpu_baseline`` and as ``--impl reference``; nothing under
``holodeck_b200/`` is imported here.
"""
import time

import numpy as np

from oracle import glue as G


def classic_workload(shape=(91, 81, 101), nfreqs=40, pta_dur_yr=16.03, scatter_dex=0.0):
    """BASELINE.json configs[1]: PS_Classic-default SAM (librarian/param_spaces_classic.py:13-42),
    Fixed_Time_2PL_SAM(3 Gyr, 1e4 pc, 100 pc, -1, +2.5, 300 steps), pta_freqs(16.03 yr, 40).  `scatter_dex` is the
    M-Mbulge scatter (PS_Classic's own default is 0.3, param_spaces_classic.py:41; 0 skips `add_scatter_to_masses`)."""
    pp = dict(G.PS_CLASSIC_DEFAULTS)
    M, Q, Z = shape
    wl = dict(params=pp, shape=tuple(shape), nfreqs=nfreqs, pta_dur_yr=pta_dur_yr, scatter_dex=float(scatter_dex))
    wl["mtot"] = np.logspace(*np.log10([1.0e4*G.MSOL, 1.0e12*G.MSOL]), M)
    wl["mrat"] = np.logspace(*np.log10([1e-3, 1.0]), Q)
    wl["redz"] = np.logspace(*np.log10([1e-3, 10.0]), Z)
    wl["fobs_cents"], wl["fobs_edges"] = G.pta_freqs(pta_dur_yr*G.YR, nfreqs)
    wl["hard"] = dict(time=pp['hard_time']*G.GYR, sepa_init=pp['hard_sepa_init']*G.PC, rchar=pp['hard_rchar']*G.PC,
                      gamma_inner=pp['hard_gamma_inner'], gamma_outer=pp['hard_gamma_outer'], nsteps=300)
    return wl


_SCATTER_JOB = None


def _scatter_slices(zsel):
    mtot, mrat, dens, dex = _SCATTER_JOB
    return G.add_scatter_to_masses(mtot, mrat, np.ascontiguousarray(dens[:, :, zsel]), dex)


def scatter_timed(mtot, mrat, dens, dex, nproc=1, sample_slices=None):
    """`add_scatter_to_masses` (sam.py:1291-1394) on the oracle.  The reference treats the redshift slices one after the
    other and independently (sam.py:1358-1392), so (a) the checker may spread them over `nproc` forked processes and
    (b) a BOUNDED timing sample is `sample_slices` evenly spaced slices on one core, linear in the slice count.
    Returns (dens_scattered or None, seconds for all Z slices on ONE core [measured or extrapolated], info)."""
    global _SCATTER_JOB
    Z = dens.shape[2]
    if sample_slices is not None and sample_slices < Z:
        zsel = np.unique(np.linspace(0, Z - 1, max(2, int(sample_slices))).astype(int))
        # the first slice of a call also builds the Delaunay triangulation that the others reuse (sam.py:1362-1370):
        # time one slice alone, then the sample, and extrapolate  t_first + (Z - 1) * t_other
        t0 = time.perf_counter()
        G.add_scatter_to_masses(mtot, mrat, np.ascontiguousarray(dens[:, :, zsel[:1]]), dex)
        t_first = time.perf_counter() - t0
        t0 = time.perf_counter()
        G.add_scatter_to_masses(mtot, mrat, np.ascontiguousarray(dens[:, :, zsel]), dex)
        dt = time.perf_counter() - t0
        t_other = max(dt - t_first, 0.0) / (zsel.size - 1)
        return None, t_first + (Z - 1) * t_other, dict(slices_timed=int(zsel.size) + 1, slices=Z,
                                                       seconds_timed=round(dt + t_first, 3))
    t0 = time.perf_counter()
    if nproc <= 1:
        out = G.add_scatter_to_masses(mtot, mrat, dens, dex)
        dt = time.perf_counter() - t0
        return out, dt, dict(slices_timed=Z, slices=Z, seconds_timed=round(dt, 3))
    import multiprocessing as mp
    _SCATTER_JOB = (mtot, mrat, dens, dex)
    parts = np.array_split(np.arange(Z), nproc)
    with mp.get_context("fork").Pool(nproc) as pool:
        outs = pool.map(_scatter_slices, parts)
    _SCATTER_JOB = None
    out = np.concatenate(outs, axis=2)
    wall = time.perf_counter() - t0
    return out, wall, dict(slices_timed=Z, slices=Z, seconds_timed=round(wall, 3), processes=nproc)


def reference_density(wl, cosmo_tables=None, nproc=1, scatter_sample=None):
    """static_binary_density of the workload.  With `scatter_dex > 0` the M-Mbulge scatter runs too: on `nproc`
    processes (checker), or -- `scatter_sample` = number of slices -- only timed on a bounded sample, in which case
    the returned density is the UNSCATTERED one (used where only the cost of the later stages matters)."""
    pp = wl["params"]
    oc = G.OracleCosmo(closed_form=True)
    dex = float(wl.get("scatter_dex", 0.0))
    mmb = G.MMBulge('KH2013', mamp_log10=pp['mmb_mamp_log10'], mplaw=pp['mmb_plaw'], scatter_dex=dex)
    gsmf = lambda m, z: G.gsmf_schechter(m, z, phi0=pp['gsmf_phi0_log10'], phiz=pp['gsmf_phiz'], mchar0_log10=pp['gsmf_mchar0_log10'], mcharz=pp['gsmf_mcharz'], alpha0=pp['gsmf_alpha0'], alphaz=pp['gsmf_alphaz'])   # noqa
    gpf = lambda m, q, z: G.gpf_power_law(m, q, z, frac_norm_allq=pp['gpf_frac_norm_allq'], malpha=pp['gpf_malpha'], qgamma=pp['gpf_qgamma'], zbeta=pp['gpf_zbeta'], max_frac=pp['gpf_max_frac'])   # noqa
    gmt = lambda m, q, z: G.gmt_power_law(m, q, z, oc.h, time_norm=pp['gmt_norm']*G.GYR, malpha=pp['gmt_malpha'], qgamma=pp['gmt_qgamma'], zbeta=pp['gmt_zbeta'])   # noqa
    dd = G.static_binary_density(wl["mtot"], wl["mrat"], wl["redz"], oc, gsmf, mmb, gpf=gpf, gmt=gmt, scatter=False)
    dd["scatter_s"] = 0.0
    dd["scatter_info"] = None
    if dex > 0.0:
        # static_binary_density zeroes the stalled bins AFTER the scatter (sam.py:368-394): same order here
        raw = dd["dens_raw"]
        scat, secs, info = scatter_timed(wl["mtot"], wl["mrat"], raw, dex, nproc=nproc, sample_slices=scatter_sample)
        dd["scatter_s"], dd["scatter_info"] = secs, info
        if scat is not None:
            scat = scat.copy()
            scat[dd["redz_prime"] < 0.0] = 0.0
            dd["dens"] = scat
    return oc, dd


def make_cosmo_tables(oc, size=200):
    z_pnts = [1000.0, 10.0, 4.0, 2.0, 1.0, 0.5, 0.1, 0.01]
    num = size // len(z_pnts)
    z0 = z_pnts[0]
    segs = []
    for z1 in z_pnts[1:]:
        segs.append(np.logspace(*np.log10([z0, z1]), num=num, endpoint=False))
        z0 = z1
    segs.append(np.linspace(z0, 0.0, num=num))
    zg = np.concatenate(segs)
    return G.StubCosmoTables(zg, oc.comoving_distance(zg), oc.age(zg))


def reference_deterministic(wl, nproc=1, scatter_sample=None):
    """Everything of `sam.gwb` that does not depend on the realization count.  Returns (state, timings).
    `nproc` / `scatter_sample`: see `reference_density` (M-Mbulge scatter of a `scatter_dex > 0` workload)."""
    tt = {}
    t0 = time.perf_counter()
    oc, dd = reference_density(wl, nproc=nproc, scatter_sample=scatter_sample)
    tt["density"] = time.perf_counter() - t0
    if dd["scatter_info"] is not None:
        # report the scatter separately, as ONE-core seconds for all slices (measured, or extrapolated from the sample)
        tt["density"] -= dd["scatter_info"]["seconds_timed"]
        tt["scatter"] = dd["scatter_s"] if nproc <= 1 or scatter_sample is not None else dd["scatter_info"]["seconds_timed"]
    tabs = make_cosmo_tables(oc)
    hp = wl["hard"]
    t0 = time.perf_counter()
    norm_log10 = G.ref_find_norm(hp["time"], wl["mtot"], wl["mrat"], hp["sepa_init"], hp["rchar"], hp["gamma_inner"],
                                 hp["gamma_outer"], hp["nsteps"])
    tt["norm"] = time.perf_counter() - t0
    sam = G.StubSam(wl["mtot"], wl["mrat"], wl["redz"], dd["dens"], dd["gmt_time"], dd["redz_prime"])
    fo_c = wl["fobs_cents"] / 2.0
    fo_e = wl["fobs_edges"] / 2.0
    t0 = time.perf_counter()
    rz, dn = G.ref_dbn(fo_c, sam, tabs, '2pwl', 10.0**norm_log10, hp["sepa_init"], hp["rchar"], hp["gamma_inner"],
                       hp["gamma_outer"], hp["nsteps"])
    rz, dn = np.asarray(rz), np.asarray(dn)
    tt["dbn"] = time.perf_counter() - t0
    edges = [wl["mtot"], wl["mrat"], wl["redz"], fo_e]
    t0 = time.perf_counter()
    number = np.asarray(G.ref_integrate(edges, dn))
    tt["integrate"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    h2fdf = G.char_strain_sq_from_bin_edges_redz(edges, rz, oc.comoving_distance)
    _, msort, qsort, zsort = G.rank_order(h2fdf, 'stable')
    tt["strain_sort"] = time.perf_counter() - t0
    state = dict(edges=edges, redz_final=rz, diff_num=dn, number=number, h2fdf=h2fdf,
                 msort=msort, qsort=qsort, zsort=zsort, norm_log10=norm_log10, dens=dd["dens"],
                 gmt_time=dd["gmt_time"], redz_prime=dd["redz_prime"], dens_noscatter=dd["dens_noscatter"],
                 dens_raw=dd.get("dens_raw"), scatter_info=dd["scatter_info"])
    return state, tt


def reference_realize(state, nreals, loudest, seed=None):
    """`loudest_hc_from_sorted` (the realised stage of `sam.gwb`, params=False) on the CPU."""
    cy, _, _ = G.ref()
    cy.ORACLE_SEED = seed
    t0 = time.perf_counter()
    hc2ss, hc2bg = cy.loudest_hc_from_sorted(state["number"], state["h2fdf"], int(nreals), int(loudest),
                                             state["msort"], state["qsort"], state["zsort"])
    dt = time.perf_counter() - t0
    cy.ORACLE_SEED = None
    return np.sqrt(np.asarray(hc2ss)), np.sqrt(np.asarray(hc2bg)), dt
