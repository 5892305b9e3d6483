"""GPU parity of the eccentric GWB kernels (K5) against the compiled reference's outputs."""
import numpy as np
import pytest

from conftest import rel_err, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gg():
    return load_golden("eccen_small")


def test_sam_calc_gwb_single_eccen(gg):
    from holodeck_b200 import cyutils
    for tag in ("a", "b"):
        gwb = cyutils.sam_calc_gwb_single_eccen(gg["ndens"], np.log10(gg["mtot"]), gg["mrat"], gg["redz"], gg["dcom"],
                                                gg["fobs"], gg[f"sepa_{tag}"], gg[f"eccen_{tag}"], int(gg["nharms"]))
        want = gg[f"gwb_{tag}"]
        assert gwb.shape == want.shape
        assert np.array_equal(gwb == 0, want == 0)
        # track "b" reaches e ~ 1e-9 where the reference's upward Bessel recursion amplifies rounding noise
        assert rel_err(gwb, want) < (1e-10 if tag == "a" else 1e-8), tag


def test_sam_calc_gwb_single_eccen_discrete(gg):
    from holodeck_b200 import cyutils
    R = int(gg["disc_R"])
    scale = float(gg["disc_scale"])
    H = int(gg["nharms"])
    args = (gg["ndens"] * scale, np.log10(gg["mtot"]), gg["mrat"], gg["redz"], gg["dcom"], gg["fobs"], gg["sepa_a"], gg["eccen_a"])
    got = cyutils.sam_calc_gwb_single_eccen_discrete(*args, H, R, seed=5)
    ref = gg["gwb_disc_a"]
    assert got.shape == ref.shape == (gg["fobs"].size, H, R)
    # expectation of the discretised sum == the continuous integral
    cont = cyutils.sam_calc_gwb_single_eccen(*args, H)
    tot_got = got.sum(axis=1)      # (F, R): sum over harmonics
    tot_ref = ref.sum(axis=1)
    sem = tot_got.std(axis=1) / np.sqrt(R)
    assert np.all(np.abs(tot_got.mean(axis=1) - cont.sum(axis=1)) < 6 * sem + 1e-9 * cont.sum(axis=1))
    # and the realised distribution matches the reference's (numpy RNG) within Monte-Carlo error
    qg = np.percentile(tot_got, [16, 50, 84], axis=1)
    qr = np.percentile(tot_ref, [16, 50, 84], axis=1)
    rng = np.random.default_rng(0)
    boots = np.array([np.percentile(tot_ref[:, rng.integers(0, R, R)], [16, 50, 84], axis=1) for _ in range(200)])
    sig = boots.std(axis=0) * np.sqrt(2.0) + 1e-3 * qr
    assert np.all(np.abs(qg - qr) < 5 * sig), (np.abs(qg - qr) / sig).max()
    # partition independence of the realizations
    lo = cyutils.sam_calc_gwb_single_eccen_discrete(*args, H, R // 2, seed=5, r0=0)
    hi = cyutils.sam_calc_gwb_single_eccen_discrete(*args, H, R // 2, seed=5, r0=R // 2)
    assert np.array_equal(got, np.concatenate([lo, hi], axis=2))


def test_eccen_through_gravwaves_api():
    import holodeck_b200 as holo
    from holodeck_b200 import gravwaves, host_relations
    from holodeck_b200.constants import PC, YR
    sam = holo.sams.Semi_Analytic_Model(shape=(12, 9, 11), mmbulge=host_relations.MMBulge_KH2013(scatter_dex=0.0))
    sepa, eccen = holo.sams.evolve_eccen_uniform_single(sam, 0.9, 0.05*PC, 80)
    assert sepa.shape == eccen.shape == (80,) and eccen[0] == 0.9 and np.all(np.diff(eccen) <= 0)
    fobs, _ = holo.utils.pta_freqs(16.03*YR, 6)
    gwb = gravwaves.sam_calc_gwb_single_eccen(fobs, sam, sepa, eccen, nharms=15)
    assert gwb.shape == (6, 15) and np.all(gwb >= 0) and np.any(gwb > 0)
    disc = gravwaves.sam_calc_gwb_single_eccen_discrete(fobs, sam, sepa, eccen, nharms=15, nreals=8, seed=1)
    assert disc.shape == (6, 15, 8)
