"""CPU tests that PIN THE ORACLE: the reference's own golden vectors and truth functions for this
path, the independent-cosmology cross checks, and the committed golden fixtures against a fresh run of
the compiled reference (oracle/_ref)."""
import numpy as np
import pytest

from oracle import glue as G
from conftest import rel_err, load_golden
from _stubs import edges_orb, sort_indices

YR = G.YR
MSOL = G.MSOL


# ---- holodeck/tests/test_utils.py:52-94  (Test_GW_Methods.test_with_fixed_values; np.isclose rtol 1e-5)

FIXED = dict(
    m1=[1.57413313e+41, 8.14164709e+41, 3.60895311e+41, 9.19991375e+39, 1.06186683e+41],
    m2=[1.62059681e+42, 9.26879287e+39, 3.91354504e+40, 3.44058558e+40, 3.10034300e+40],
    aa=[5.24456784e+17, 2.78829325e+16, 7.93618648e+15, 3.68590010e+14, 2.72097813e+16],
    ee=[7.25868969e-01, 3.75506309e-01, 3.42479721e-01, 5.43473856e-01, 9.99000000e-01],
    dc=[6.63399191e+25, 1.76236980e+24, 5.54714846e+22, 3.78302329e+23, 1.50096058e+20],
    mc=[3.92683830e+41, 5.54007398e+40, 9.32292666e+40, 1.48710242e+40, 4.81973673e+40],
    hs=[3.72246602e-15, 5.35750864e-15, 4.05245287e-13, 2.78780956e-15, 4.98735430e-11],
    gwlum=[9.75931974e+47, 1.42668479e+45, 8.08693051e+45, 1.77996894e+43, 8.96782668e+44],
    dedt=[-8.50812477e-18, -1.48341135e-15, -1.73541089e-13, -2.71156486e-10, -1.60176941e-09],
    dade=[2.16041096e+18, 7.42617422e+16, 2.15759501e+16, 1.05091558e+15, 2.72192326e+19],
    dadt=[-1.83810460e+01, -1.10160711e+02, -3.74431389e+03, -2.84962577e+05, -4.35989342e+10],
    dfdt=[2.92125254e-18, 7.34264884e-21, 1.50683392e-20, 2.20635454e-21, 2.94579604e-11],
    tau=[4.03507842e+11, 1.05535297e+13, 4.43271589e+12, 9.44834061e+13, 1.33112288e+13],
)


def _check_fixed(mod):
    ff = {kk: np.array(vv) for kk, vv in FIXED.items()}
    freq = 1.0 / YR
    got = dict(
        mc=mod.chirp_mass(ff["m1"], ff["m2"]),
        hs=mod.gw_strain_source(ff["mc"], ff["dc"], freq),
        gwlum=mod.gw_lum_circ(ff["mc"], freq),
        dedt=mod.gw_dedt(ff["m1"], ff["m2"], ff["aa"], ff["ee"]),
        dade=mod.gw_dade(ff["aa"], ff["ee"]),
        dadt=mod.gw_hardening_rate_dadt(ff["m1"], ff["m2"], ff["aa"], ff["ee"]),
        dfdt=mod.gw_hardening_rate_dfdt(ff["m1"], ff["m2"], freq, ff["ee"])[0],
        tau=mod.gw_hardening_timescale_freq(ff["mc"], freq),
    )
    for kk, vv in got.items():
        assert np.all(np.isclose(ff[kk], vv)), f"{kk} did not match the reference's cached values"


def test_reference_gw_known_answers_oracle():
    _check_fixed(G)


def test_reference_gw_known_answers_product_host_utils():
    from holodeck_b200 import utils
    _check_fixed(utils)


# ---- holodeck/tests/test_host_relations__mmbulge.py:165-228 truth functions

def _truth_mm2013(mbulge):
    return np.power(10.0, 8.46 + 1.05 * np.log10(mbulge / (1e11 * MSOL))) * MSOL


def _truth_kh2013(mbulge):
    return (10.0 ** 8.69) * MSOL * np.power(mbulge / (1e11 * MSOL), 1.17)


@pytest.mark.parametrize("kind,truth", [("MM2013", _truth_mm2013), ("KH2013", _truth_kh2013)])
def test_mmbulge_truth(kind, truth):
    from holodeck_b200 import host_relations
    mbulge = np.logspace(8, 13, 11) * MSOL
    for rel in (G.MMBulge(kind), getattr(host_relations, f"MMBulge_{kind}")()):
        vals = rel.mbh_from_mbulge(mbulge)
        assert np.allclose(vals, truth(mbulge))
        assert np.allclose(mbulge, rel.mbulge_from_mbh(vals))
    # dmstar_dmbh vs finite difference (test_host_relations__mmbulge.py:341-385)
    for rel in (G.MMBulge(kind), getattr(host_relations, f"MMBulge_{kind}")()):
        mstar = np.logspace(9, 12.5, 30) * MSOL
        dd = 1e-6
        lo, hi = mstar * (1 - dd), mstar * (1 + dd)
        deriv = (hi - lo) / (rel.mbh_from_mstar(hi) - rel.mbh_from_mstar(lo))
        assert np.allclose(rel.dmstar_dmbh(mstar), deriv, rtol=1e-5)


# ---- cosmology: product closed forms / GL quadrature vs the oracle's adaptive quadrature

def test_cosmology_against_independent_quadrature():
    from holodeck_b200 import cosmo
    oc = G.OracleCosmo()
    zz = np.array([0.0, 1e-6, 1e-3, 0.02, 0.3, 1.0, 2.5, 6.0, 10.0, 100.0, 1000.0])
    assert rel_err(cosmo.age(zz), oc.age(zz)) < 1e-12
    assert rel_err(cosmo.comoving_distance(zz[1:]), oc.comoving_distance(zz[1:])) < 1e-12
    assert rel_err(cosmo.dtdz(zz), oc.dtdz(zz)) < 1e-14
    tt = oc.age(zz[1:])
    assert np.max(np.abs(cosmo.tage_to_z(tt) - zz[1:]) / (1 + zz[1:])) < 1e-12
    assert np.max(np.abs(oc.tage_to_z(tt) - zz[1:]) / (1 + zz[1:])) < 1e-12
    fast = G.OracleCosmo(closed_form=True)
    zs = np.logspace(-4, 1, 200)
    assert rel_err(fast.comoving_distance(zs), oc.comoving_distance(zs)) < 1e-11
    assert rel_err(fast.age(zs), oc.age(zs)) < 1e-12
    # interpolation tables: decreasing z ending at 0, increasing age ending at the age of the universe
    assert cosmo._grid_z[0] == 1000.0 and cosmo._grid_z[-1] == 0.0 and np.all(np.diff(cosmo._grid_z) < 0)
    assert np.all(np.diff(cosmo._grid_age) > 0) and cosmo._grid_age[-1] == cosmo.age_universe
    assert cosmo._grid_dcom[-1] == 0.0 and cosmo._grid_z.size == 200
    assert abs(cosmo.age_universe / G.GYR - 13.7527) < 1e-3


# ---- the committed fixtures are what the compiled reference produces today

def test_golden_fixtures_match_live_reference(golden):
    cy, scy, holo = G.ref()
    gg = golden
    sam = G.StubSam(gg["mtot"], gg["mrat"], gg["redz"], gg["dens"], gg.get("gmt_time"), gg.get("redz_prime"))
    tabs = G.StubCosmoTables(gg["grid_z"], gg["grid_dcom"], gg["grid_age"])
    fo = gg["fobs_cents"] / 2.0
    if str(gg["hard"]) == "2pwl":
        hp = gg["hard_params"]
        nl = G.ref_find_norm(hp[0], gg["mtot"], gg["mrat"], hp[1], hp[2], hp[3], hp[4], int(hp[5]))
        assert np.array_equal(nl, gg["norm_log10"])
        rz, dn = G.ref_dbn(fo, sam, tabs, "2pwl", 10.0**nl, hp[1], hp[2], hp[3], hp[4], int(hp[5]))
    else:
        rz, dn = G.ref_dbn(fo, sam, tabs, "gw")
    assert np.array_equal(np.asarray(rz), gg["redz_final"]) and np.array_equal(np.asarray(dn), gg["diff_num"])
    number = np.asarray(G.ref_integrate(edges_orb(gg), gg["diff_num"]))
    assert np.array_equal(number, gg["number"])
    ms, qs, zs = sort_indices(gg)
    R, L, seed = int(gg["nreals"]), int(gg["nloud"]), int(gg["seed"])
    cy.ORACLE_SEED = seed
    a, b = cy.loudest_hc_from_sorted(gg["number"], gg["h2fdf"], R, L, ms, qs, zs)
    cy.ORACLE_SEED = None
    assert np.array_equal(np.asarray(a), gg["l1_hc2ss"]) and np.array_equal(np.asarray(b), gg["l1_hc2bg"])
    # and the rank order stored in the fixture is the stable argsort of -h2fdf[...,0]
    order = np.argsort(-gg["h2fdf"][..., 0].flatten(), kind="stable")
    assert np.array_equal(order, gg["order"])


def test_golden_density_matches_glue(golden_classic):
    """dens / gmt_time / redz_prime in the fixture == oracle/glue.py with the quadrature cosmology
    (a subset of z, the quadrature inverse is slow) and with the closed-form cosmology (all)."""
    from oracle import chain
    gg = golden_classic
    wl = chain.classic_workload(shape=(gg["mtot"].size, gg["mrat"].size, gg["redz"].size), nfreqs=gg["fobs_cents"].size)
    _, dd = chain.reference_density(wl)
    assert rel_err(dd["dens"], gg["dens"]) < 1e-11
    assert rel_err(dd["gmt_time"], gg["gmt_time"]) < 1e-13
    assert np.max(np.abs(dd["redz_prime"] - gg["redz_prime"])) < 1e-11
