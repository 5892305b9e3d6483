"""CPU tests of the kernels' per-element math (holodeck_b200/csrc/*.cuh compiled for the host by
tests/hostemu) against the golden fixtures, i.e. against the compiled reference.  This checks the
arithmetic the CUDA kernels run -- not the launch machinery, which the `-m gpu` tests cover."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import rel_err, load_golden
from _stubs import edges_orb

HERE = Path(__file__).resolve().parent
SO = HERE / "hostemu" / "libhostemu.so"


@pytest.fixture(scope="module")
def emu():
    src = HERE / "hostemu" / "hostemu.cpp"
    if (not SO.exists()) or SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(SO), str(src)], check=True)
    return C.CDLL(str(SO))


def P(arr):
    return None if arr is None else arr.ctypes.data_as(C.c_void_p)


def test_emu_norm_and_dbn(emu, golden):
    from holodeck_b200 import _lib as L
    gg = golden
    cc = L.cy_consts()
    M, Q, Z = gg["mtot"].size, gg["mrat"].size, gg["redz"].size
    fo = np.ascontiguousarray(gg["fobs_cents"] / 2.0)
    F = fo.size
    rz = np.zeros((M, Q, Z, F))
    dn = np.zeros((M, Q, Z, F))
    if str(gg["hard"]) == "2pwl":
        hp = gg["hard_params"]
        mt, mr = np.meshgrid(gg["mtot"], gg["mrat"], indexing="ij")
        mt, mr = np.ascontiguousarray(mt.flatten()), np.ascontiguousarray(mr.flatten())
        out = np.zeros(M * Q)
        emu.emu_norm_2pwl.argtypes = [L.CyConsts, C.c_double, C.c_void_p, C.c_void_p, C.c_int] + [C.c_double] * 4 + \
            [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        emu.emu_norm_2pwl(cc, hp[0], P(mt), P(mr), M * Q, hp[1], hp[2], hp[3], hp[4], int(hp[5]), 0, None, P(out))
        assert np.max(np.abs(out - gg["norm_log10"].flatten())) < 1e-12
        norm = np.ascontiguousarray(10.0 ** gg["norm_log10"])
        emu.emu_dbn_2pwl.argtypes = [L.CyConsts, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p] + [C.c_double] * 3 + \
            [C.c_void_p] * 5 + [C.c_int] * 3 + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 2
        emu.emu_dbn_2pwl(cc, P(fo), F, hp[1], int(hp[5]), P(norm), hp[2], hp[3], hp[4], P(gg["dens"]), P(gg["mtot"]),
                         P(gg["mrat"]), P(gg["redz"]), P(gg["gmt_time"]), M, Q, Z, P(gg["grid_z"]), P(gg["grid_dcom"]),
                         P(gg["grid_age"]), gg["grid_z"].size, P(rz), P(dn))
    else:
        zp = gg["redz_prime"] if "redz_prime" in gg else np.ascontiguousarray(np.broadcast_to(gg["redz"], (M, Q, Z)))
        emu.emu_dbn_gw.argtypes = [L.CyConsts, C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_void_p] * 2 + \
            [C.c_int] + [C.c_void_p] * 2
        emu.emu_dbn_gw(cc, P(fo), F, P(gg["dens"]), P(gg["mtot"]), P(gg["mrat"]), P(zp), M, Q, Z, P(gg["grid_z"]),
                       P(gg["grid_dcom"]), gg["grid_z"].size, P(rz), P(dn))
    assert np.array_equal(rz == -1, gg["redz_final"] == -1)
    assert rel_err(rz, gg["redz_final"]) < 1e-13
    assert np.array_equal(dn == 0, gg["diff_num"] == 0)
    assert rel_err(dn, gg["diff_num"]) < 1e-12


def test_emu_dbn_named_size(emu):
    """The K1b arithmetic (csrc/holo_math.cuh, compiled for the host) on ALL 29,778,840 edge cells of the named
    91x81x101 x 40 grid against the compiled reference (oracle/_ref `dynamic_binary_number_at_fobs`,
    sam_cyutils.pyx:510-781): no sentinel / zero-pattern mismatch -- i.e. none of the threshold branches
    (`time_left > age_universe` :667, bracket ties :709, table-end extrapolation :683-687) flips -- and 1e-12 relative.
    (~35 s: the reference's dbn is 11 s, the emulation 10 s.)"""
    from oracle import chain, glue
    from holodeck_b200 import _lib as L
    wl = chain.classic_workload()
    st, _ = chain.reference_deterministic(wl)
    hp = wl["hard"]
    M, Q, Z = wl["shape"]
    fo = np.ascontiguousarray(wl["fobs_cents"] / 2.0)
    F = fo.size
    tabs = chain.make_cosmo_tables(glue.OracleCosmo(closed_form=True))
    rz = np.zeros((M, Q, Z, F))
    dn = np.zeros((M, Q, Z, F))
    norm = np.ascontiguousarray(10.0 ** st["norm_log10"])
    emu.emu_dbn_2pwl.argtypes = [L.CyConsts, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p] + [C.c_double] * 3 + \
        [C.c_void_p] * 5 + [C.c_int] * 3 + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 2
    emu.emu_dbn_2pwl(L.cy_consts(), P(fo), F, hp["sepa_init"], int(hp["nsteps"]), P(norm), hp["rchar"], hp["gamma_inner"],
                     hp["gamma_outer"], P(st["dens"]), P(wl["mtot"]), P(wl["mrat"]), P(wl["redz"]), P(st["gmt_time"]), M, Q, Z,
                     P(tabs._grid_z), P(tabs._grid_dcom), P(tabs._grid_age), tabs._grid_z.size, P(rz), P(dn))
    assert np.count_nonzero((rz == -1) != (st["redz_final"] == -1)) == 0
    assert np.count_nonzero((dn == 0) != (st["diff_num"] == 0)) == 0
    assert rel_err(rz, st["redz_final"]) < 1e-12
    assert rel_err(dn, st["diff_num"]) < 1e-12
    # ... and the fused K2 + K2b arithmetic on the reference's grids: `number` vs the compiled reference
    assert np.count_nonzero(st["redz_final"] != -1) > 4_000_000


def test_emu_integrate_and_strain(emu, golden):
    from holodeck_b200 import _lib as L, cosmo, utils
    from holodeck_b200.constants import NWTG
    gg = golden
    edges = edges_orb(gg)
    M, Q, Z = gg["mtot"].size, gg["mrat"].size, gg["redz"].size
    F = gg["fobs_cents"].size
    shape = (M - 1, Q - 1, Z - 1, F)
    numb, h2, zm, dc, se, an = [np.zeros(shape) for _ in range(6)]
    l10m = np.log10(edges[0])
    dlnf = np.diff(np.log(edges[3]))
    mtm, mrm = utils.midpoints(edges[0]), utils.midpoints(edges[1])
    fc = utils.midpoints(edges[3])
    fdf = fc / np.diff(edges[3])
    cp = L.cosmo_params(cosmo)
    emu.emu_integrate_and_strain.argtypes = [C.c_void_p, C.c_double, C.c_double] + [C.c_void_p] * 11 + [C.c_int] * 4 + [C.c_void_p] * 6
    emu.emu_integrate_and_strain(C.byref(cp), utils._GW_SRC_CONST, NWTG, P(l10m), P(edges[1]), P(edges[2]), P(dlnf),
                                 P(gg["diff_num"]), P(gg["redz_final"]), None, P(mtm), P(mrm), P(fc), P(fdf), M, Q, Z, F,
                                 P(numb), P(h2), P(zm), P(dc), P(se), P(an))
    assert np.array_equal(numb, gg["number"])
    assert np.array_equal(h2 == 0, gg["h2fdf"] == 0)
    assert rel_err(h2, gg["h2fdf"]) < 1e-12
    assert np.max(np.abs(zm - gg["par_redz"])) < 1e-14
    fin = np.isfinite(gg["par_dcom"])
    assert np.array_equal(np.isfinite(dc), fin) and rel_err(dc[fin], gg["par_dcom"][fin]) < 1e-12
    fin = np.isfinite(gg["par_angs"])
    assert rel_err(an[fin], gg["par_angs"][fin]) < 1e-12 and rel_err(se[fin], gg["par_sepa"][fin]) < 1e-13


def test_emu_density(emu, golden_classic):
    from holodeck_b200 import _lib as L, cosmo, utils
    from holodeck_b200.constants import MSOL, GYR
    gg = golden_classic
    M, Q, Z = gg["mtot"].size, gg["mrat"].size, gg["redz"].size
    sp = L.SamParams()
    sp.gsmf_kind, sp.use_gmr, sp.has_gmt = 0, 0, 1
    for ii, vv in enumerate([-2.77, -0.6, MSOL * np.power(10.0, 11.24), 0.11, -1.21, -0.03]):
        sp.gsmf[ii] = vv
    for ii, vv in enumerate([0.025 / ((1.0 - 0.25) / 1.0), MSOL * np.power(10.0, 11.0), 0.0, 1.0, 0.0, 1.0]):
        sp.gpf[ii] = vv
    for ii, vv in enumerate([0.5 * GYR, 1.0e11 * MSOL * (0.4 / cosmo.h), 0.0, -0.5, -1.0]):
        sp.gmt[ii] = vv
    for ii, vv in enumerate([MSOL * np.power(10.0, 8.69), 1.10, 1.0e11 * MSOL, 0.615]):
        sp.mmb[ii] = vv
    sp.hubble_time, sp.om0, sp.age_universe = cosmo.hubble_time, cosmo.Om0, utils._AGE_UNIVERSE_GYR * GYR
    dens, gt, zp = [np.zeros((M, Q, Z)) for _ in range(3)]
    age_z, dtdz = cosmo.age(gg["redz"]), cosmo.dtdz(gg["redz"])
    emu.emu_sam_density.argtypes = [C.c_void_p] * 5 + [C.c_int] * 3 + [C.c_void_p] * 5
    emu.emu_sam_density(P(gg["mtot"]), P(gg["mrat"]), P(gg["redz"]), P(age_z), P(dtdz), M, Q, Z, C.byref(sp), None,
                        P(dens), P(gt), P(zp))
    dens[zp < 0] = 0.0
    assert np.array_equal(zp == -1, gg["redz_prime"] == -1)
    assert rel_err(dens, gg["dens"]) < 1e-11
    assert rel_err(gt, gg["gmt_time"]) < 1e-13


def test_emu_density_bf_sigmoid(emu):
    """K0 arithmetic with the sigmoid bulge fraction (`BF_Sigmoid`, host_relations.py:198-331; the M-Mbulge relation
    of the `PS_Astro_Strong_*` spaces, param_spaces.py:246-257): the kernel evaluates the closed-form sigmoid and the
    two quadratic interpolants scipy builds (handed over as piecewise polynomials by `Semi_Analytic_Model._kernel_params`)
    -- against the oracle's restatement of the reference procedure (scipy interp1d), double-Schechter GSMF +
    Illustris merger rate, three parameter sets incl. a narrow and a wide transition."""
    from oracle import glue
    from holodeck_b200 import sams, host_relations, cosmo
    emu.emu_sam_density.argtypes = [C.c_void_p] * 5 + [C.c_int] * 3 + [C.c_void_p] * 5
    for (flo, fhi, mc, width) in ((0.4, 0.8, 11.0, 1.0), (0.1, 1.0, 10.5, 0.5), (0.35, 0.6, 11.4, 1.5)):
        bf = host_relations.BF_Sigmoid(bulge_frac_lo=flo, bulge_frac_hi=fhi, mstar_char_log10=mc, width_dex=width)
        mmb = host_relations.MMBulge_KH2013(mamp_log10=8.69, mplaw=1.17, scatter_dex=0.0, bulge_frac=bf)
        sam = sams.Semi_Analytic_Model(gsmf=sams.GSMF_Double_Schechter, gmr=sams.GMR_Illustris, mmbulge=mmb, shape=(23, 19, 17))
        par = sam._kernel_params()
        assert par.bf_kind == 1 and par.bf_n > 100 and sam._bf_tables_host.size == 2 * (4 * par.bf_n + 1)
        M, Q, Z = sam.shape
        dens = np.zeros((M, Q, Z))
        age_z, dtdz = cosmo.age(sam.redz), cosmo.dtdz(sam.redz)
        emu.emu_sam_density(P(sam.mtot), P(sam.mrat), P(sam.redz), P(age_z), P(dtdz), M, Q, Z, C.byref(par),
                            P(sam._bf_tables_host), P(dens), None, None)
        oc = glue.OracleCosmo(closed_form=True)
        omm = glue.MMBulge('KH2013', mamp_log10=8.69, mplaw=1.17, scatter_dex=0.0,
                           bulge_frac=glue.BFSigmoid(bulge_frac_lo=flo, bulge_frac_hi=fhi, mstar_char_log10=mc, width_dex=width))
        want = glue.static_binary_density(sam.mtot, sam.mrat, sam.redz, oc, glue.gsmf_double_schechter, omm,
                                          gmr=glue.gmr_illustris, scatter=False)["dens"]
        assert np.array_equal(dens == 0, want == 0)
        assert rel_err(dens, want) < 1e-10, (flo, fhi, mc, width, rel_err(dens, want))
        # ... and the host mirror of the class itself against the oracle's
        ms = np.logspace(7.0, 13.0, 400) * 1.988409870698051e+33
        assert rel_err(bf.bulge_frac(ms), omm._bfrac.bulge_frac(ms.copy())) < 1e-15
        mb = bf.mbulge_from_mstar(ms)
        assert rel_err(bf.mstar_from_mbulge(mb), omm._bfrac.mstar_from_mbulge(mb.copy())) < 1e-14
        assert rel_err(bf.dmstar_dmbulge(mb), omm._bfrac.dmstar_dmbulge(mb.copy())) < 1e-14


def test_emu_samplers_are_exact_poisson(emu):
    """Chi-square of the Philox/inversion/PTRS samplers against the exact pmf, every class."""
    import scipy.stats as st
    emu.emu_draw_elements.argtypes = [C.c_double, C.c_int64, C.c_uint64, C.c_double, C.c_uint64, C.c_void_p]
    N = 300000
    for ii, lam in enumerate([2e-3, 0.5, 4.0, 9.99, 10.0, 14.0, 80.0, 2.5e3, 7e5, 3e9]):
        dd = np.zeros(N)
        emu.emu_draw_elements(lam, N, 900 + ii, 1e10, 5 + ii, P(dd))
        assert np.all(dd == np.floor(dd)) and np.all(dd >= 0)
        assert abs(dd.mean() - lam) < 5.0 * np.sqrt(lam / N)
        if lam >= 0.5:
            lo, hi = int(st.poisson.ppf(1e-3, lam)), int(st.poisson.ppf(1 - 1e-3, lam)) + 1
            nb = min(60, hi - lo + 1)
            edges = np.unique(np.linspace(lo, hi + 1, nb + 1).astype(np.int64))     # bin i = [edges[i], edges[i+1])
            which = np.searchsorted(edges, dd.astype(np.int64), side="right") - 1
            inside = (dd >= edges[0]) & (dd < edges[-1])
            obs = np.bincount(which[inside], minlength=edges.size - 1).astype(float)
            exp = N * (st.poisson.cdf(edges[1:] - 1, lam) - st.poisson.cdf(edges[:-1] - 1, lam))
            sel = exp > 20
            chi2 = np.sum((obs[sel] - exp[sel])**2 / exp[sel])
            assert st.chi2.sf(chi2, sel.sum()) > 1e-4, (lam, chi2, sel.sum())
    # the normal branch (lam > thresh) is not floored
    dd = np.zeros(20000)
    emu.emu_draw_elements(5e10, 20000, 1, 1e10, 0, P(dd))
    assert np.any(dd != np.floor(dd)) and abs(dd.mean() / 5e10 - 1) < 1e-7 and abs(dd.std() / np.sqrt(5e10) - 1) < 0.03


def test_emu_table_sampler(emu):
    """TABLE class of the realization kernel (32-bit CDF thresholds + exact slow path): the thresholds are
    floor(cdf 2^32) to within one unit, fast and slow path agree draw by draw, and both are exact Poisson."""
    import scipy.stats as st
    emu.emu_draw_table.argtypes = [C.c_double, C.c_int64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    emu.emu_draw_table.restype = C.c_int
    N = 300000
    for ii, lam in enumerate([2.0**-8, 0.05, 0.7, 3.0, 31.0, 33.0, 86.0, 400.0, 1500.0, 4000.0]):
        dd = np.zeros(N)
        tab = np.zeros(1024, dtype=np.uint32)
        kmin, nslow = C.c_int(0), C.c_int64(0)
        W = emu.emu_draw_table(lam, N, 77 + ii, 0, P(dd), P(tab), C.byref(kmin), C.byref(nslow))
        assert 1 <= W <= 1024
        kk = kmin.value + np.arange(W)
        cdf = st.poisson.cdf(kk, lam)
        assert np.all(np.abs(tab[:W].astype(np.float64) - np.minimum(np.floor(cdf * 2.0**32), 2.0**32 - 1)) <= 1.0)
        assert st.poisson.cdf(kmin.value - 1, lam) < 2.0**-64 and st.poisson.sf(kk[-1], lam) < 1e-9
        assert nslow.value <= 5 + 50 * N * 3 * W / 2.0**32
        # every draw through the exact slow path gives the same counts
        d1 = np.zeros(30000)
        emu.emu_draw_table(lam, 30000, 77 + ii, 1, P(d1), None, None, None)
        assert np.array_equal(d1, dd[:30000])
        assert abs(dd.mean() - lam) < 5.0 * np.sqrt(lam / N)
        if lam >= 0.5:
            lo, hi = int(st.poisson.ppf(1e-3, lam)), int(st.poisson.ppf(1 - 1e-3, lam)) + 1
            nb = min(60, hi - lo + 1)
            edges = np.unique(np.linspace(lo, hi + 1, nb + 1).astype(np.int64))
            which = np.searchsorted(edges, dd.astype(np.int64), side="right") - 1
            inside = (dd >= edges[0]) & (dd < edges[-1])
            obs = np.bincount(which[inside], minlength=edges.size - 1).astype(float)
            exp = N * (st.poisson.cdf(edges[1:] - 1, lam) - st.poisson.cdf(edges[:-1] - 1, lam))
            sel = exp > 20
            chi2 = np.sum((obs[sel] - exp[sel])**2 / exp[sel])
            assert st.chi2.sf(chi2, sel.sum()) > 1e-4, (lam, chi2, sel.sum())
        else:
            p1 = -np.expm1(-lam)
            assert abs(np.mean(dd >= 1) - p1) < 5.0 * np.sqrt(p1 / N)


def test_emu_superposition_group(emu):
    """`draw_group` (all elements of a pass with lam < 0.25 drawn as one Poisson process, events assigned to
    members with probability lam_k / Lambda): member counts are independent Poisson(lam_k)."""
    import scipy.stats as st
    emu.emu_draw_group_members.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_void_p]
    rng = np.random.default_rng(5)
    N = 200000
    for K, scale in [(4, 0.25), (64, 0.25), (400, 0.2), (512, 1e-5)]:
        lam = rng.uniform(0.0, 1.0, K)**3 * scale
        lam[0] = scale * 0.999
        out = np.zeros((N, K))
        emu.emu_draw_group_members(P(lam), K, N, 31 + K, P(out))
        assert np.all(out == np.floor(out)) and np.all(out >= 0)
        tot = out.sum(axis=1)
        Lam = lam.sum()
        assert abs(tot.mean() - Lam) < 5.0 * np.sqrt(Lam / N)
        assert abs(tot.var() - Lam) < 6.0 * np.sqrt((Lam + 2 * Lam**2) / N) + 1e-12
        # member marginals: mean, P(n >= 1), P(n >= 2)
        big = lam * N > 50
        if not big.any():
            continue
        zz = (out.mean(axis=0)[big] - lam[big]) / np.sqrt(lam[big] / N)
        assert np.max(np.abs(zz)) < 5.0
        p1 = -np.expm1(-lam[big])
        z1 = ((out[:, big] >= 1).mean(axis=0) - p1) / np.sqrt(p1 / N)
        assert np.max(np.abs(z1)) < 5.0
        p2 = st.poisson.sf(1, lam[0])
        assert abs((out[:, 0] >= 2).mean() - p2) < 5.0 * np.sqrt(p2 / N) + 1e-9
        # independence of the two largest members: covariance ~ 0
        jj = np.argsort(lam)[-2:]
        cov = np.mean(out[:, jj[0]] * out[:, jj[1]]) - out[:, jj[0]].mean() * out[:, jj[1]].mean()
        assert abs(cov) < 5.0 * np.sqrt(lam[jj[0]] * lam[jj[1]] / N) + 1e-12


def test_emu_philox_known_answers(emu):
    """Random123 known-answer vectors for Philox4x32-10."""
    emu.emu_philox.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
    out = np.zeros(4, dtype=np.uint32)
    emu.emu_philox(0, 0, 0, 0, 0, 0, P(out))
    assert [int(vv) for vv in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    emu.emu_philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, P(out))
    assert [int(vv) for vv in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    emu.emu_philox(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0, P(out))
    assert [int(vv) for vv in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_emu_eccentric_gwb(emu):
    from holodeck_b200 import _lib as L
    gg = load_golden("eccen_small")
    M, Q, Z, F, H = gg["mtot"].size, gg["mrat"].size, gg["redz"].size, gg["fobs"].size, int(gg["nharms"])
    l10 = np.log10(gg["mtot"])
    emu.emu_eccen_gwb.argtypes = [C.c_double, C.c_double] + [C.c_void_p] * 8 + [C.c_int] * 6 + [C.c_void_p]
    for tag, tol in (("a", 1e-12), ("b", 1e-8)):
        out = np.zeros((F, H))
        sepa, ecc = gg[f"sepa_{tag}"], gg[f"eccen_{tag}"]
        emu.emu_eccen_gwb(L.cy_consts().gw_dadt_sep_const, L.cy_gw_src_const(), P(gg["ndens"]), P(l10), P(gg["mrat"]),
                          P(gg["redz"]), P(gg["dcom"]), P(gg["fobs"]), P(sepa), P(ecc), M, Q, Z, F, sepa.size, H, P(out))
        assert np.array_equal(out == 0, gg[f"gwb_{tag}"] == 0)
        assert rel_err(out, gg[f"gwb_{tag}"]) < tol
