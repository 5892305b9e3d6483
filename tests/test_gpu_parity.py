"""GPU parity tests: every CUDA entry point, called through the drop-in Python shims (i.e. through
the C ABI), against the golden fixtures produced by the oracle (compiled reference + numpy glue).

Tolerances: deterministic fp64 quantities rel. 1e-10 (BASELINE.json north_star); quantities that
are copies of inputs (hc2ss, sspar, ssidx, sentinels) must be bit-exact in supplied-count mode.
"""
import numpy as np
import pytest

from conftest import rel_err
from _stubs import GoldenSam, GoldenCosmo, golden_hard, edges_orb, sort_indices

pytestmark = pytest.mark.gpu

RTOL = 1e-10


@pytest.fixture(scope="module")
def holo():
    import holodeck_b200
    return holodeck_b200


def test_native_library_is_loaded(holo):
    from holodeck_b200 import _lib
    lib = _lib.require_gpu()
    assert lib.holo_device_count() >= 1


def test_integrate_differential_number(holo, golden):
    from holodeck_b200.sams import sam_cyutils
    numb = sam_cyutils.integrate_differential_number_3dx1d(edges_orb(golden), golden["diff_num"])
    assert numb.shape == golden["number"].shape
    assert rel_err(numb, golden["number"]) < 1e-14
    assert np.array_equal(numb == 0, golden["number"] == 0)
    with pytest.raises(AssertionError):
        sam_cyutils.integrate_differential_number_3dx1d(edges_orb(golden), golden["diff_num"][:-1])


def test_integrate_random_dense(holo):
    from holodeck_b200.sams import sam_cyutils
    from oracle import glue
    rng = np.random.default_rng(3)
    edges = [np.sort(rng.uniform(1e40, 1e45, 7)), np.sort(rng.uniform(0, 1, 5)), np.sort(rng.uniform(0, 5, 6)),
             np.sort(rng.uniform(1e-9, 1e-7, 9))]
    dnum = rng.uniform(0, 10, (7, 5, 6, 8))
    want = np.asarray(glue.ref_integrate(edges, dnum))
    got = sam_cyutils.integrate_differential_number_3dx1d(edges, dnum)
    assert rel_err(got, want) < 1e-14


def test_find_2pwl_hardening_norm(holo, golden):
    if str(golden["hard"]) != "2pwl":
        pytest.skip("GW-only fixture")
    from holodeck_b200.sams import sam_cyutils
    hp = golden["hard_params"]
    mt, mr = np.meshgrid(golden["mtot"], golden["mrat"], indexing="ij")
    got = sam_cyutils.find_2pwl_hardening_norm(hp[0], mt.flatten(), mr.flatten(), hp[1], hp[2], hp[3], hp[4], int(hp[5]))
    want = golden["norm_log10"].flatten()
    diff = np.abs(got - want)
    # the root is only defined to xtol=1e-3 dex; following scipy's brentq step-for-step we expect ~1e-13,
    # and tolerate isolated branch flips
    assert np.sum(diff > 1e-10) <= max(1, diff.size // 200), f"{np.sum(diff > 1e-10)} of {diff.size} roots differ, max {diff.max():.3e}"
    assert diff.max() < 2e-3
    # lifetimes at the golden norms
    for ii, want_t in zip(golden["lifetime_idx"], golden["lifetime"]):
        got_t = sam_cyutils.integrate_binary_evolution_2pwl(want[ii], mt.flat[ii], mr.flat[ii], hp[1], hp[2], hp[3], hp[4], int(hp[5]))
        assert abs(got_t / want_t - 1.0) < 1e-12


def test_dynamic_binary_number(holo, golden):
    from holodeck_b200.sams import sam_cyutils
    sam = GoldenSam(golden)
    hard = golden_hard(golden, holo)
    redz_final, diff_num = sam_cyutils.dynamic_binary_number_at_fobs(golden["fobs_cents"] / 2.0, sam, hard, GoldenCosmo(golden))
    want_rz, want_dn = golden["redz_final"], golden["diff_num"]
    assert redz_final.shape == want_rz.shape
    mism = np.sum((redz_final == -1.0) != (want_rz == -1.0))
    assert mism == 0, f"{mism} cells differ in the reached/unreached (-1) pattern"
    assert rel_err(redz_final, want_rz) < RTOL
    assert np.array_equal(diff_num == 0, want_dn == 0)
    assert rel_err(diff_num, want_dn) < RTOL


def test_unknown_hardening_raises(holo, golden):
    from holodeck_b200.sams import sam_cyutils
    with pytest.raises(ValueError):
        sam_cyutils.dynamic_binary_number_at_fobs(golden["fobs_cents"] / 2.0, GoldenSam(golden), object(), GoldenCosmo(golden))


def test_char_strain_and_params(holo, golden):
    from holodeck_b200 import gravwaves
    edges = edges_orb(golden)
    h2 = gravwaves.char_strain_sq_from_bin_edges_redz(edges, golden["redz_final"])
    assert np.array_equal(h2 == 0, golden["h2fdf"] == 0)
    assert rel_err(h2, golden["h2fdf"]) < RTOL
    h2n = gravwaves.char_strain_sq_from_bin_edges(edges)
    assert rel_err(h2n, golden["h2fdf_noredz"]) < RTOL
    out = gravwaves._char_strain_sq(edges, golden["redz_final"], params=True, dnum=golden["diff_num"])
    assert rel_err(out["number"].cpu().numpy(), golden["number"]) < 1e-14
    assert rel_err(out["h2fdf"].cpu().numpy(), golden["h2fdf"]) < RTOL
    zmid = out["zmid"].cpu().numpy()
    assert np.array_equal(zmid == -1.0, golden["par_redz"] == -1.0)
    assert np.max(np.abs(zmid - golden["par_redz"])) < 1e-13
    for key, name in (("dcom", "par_dcom"), ("sepa", "par_sepa"), ("angs", "par_angs")):
        got = out[key].cpu().numpy()
        want = golden[name]
        assert np.array_equal(np.isfinite(got), np.isfinite(want)), key
        sel = np.isfinite(want)
        assert rel_err(got[sel], want[sel]) < RTOL, key


def test_gwb_expectation(holo, golden):
    from holodeck_b200 import gravwaves
    hc = gravwaves._gws_from_number_grid_integrated_redz(edges_orb(golden), golden["redz_final"], golden["number"], False)
    assert rel_err(hc**2, golden["hc2_expect"]) < RTOL


def test_sam_poisson_gwb_supplied_counts(holo, golden):
    from holodeck_b200 import cyutils
    R = int(golden["nreals"])
    gwb = cyutils.sam_poisson_gwb(golden["number"], golden["h2fdf"], R, counts=golden["counts_gwb"])
    assert gwb.shape == golden["gwb_ref"].shape
    assert rel_err(gwb, golden["gwb_ref"]) < 1e-12


def test_loudest_hc_from_sorted_supplied_counts(holo, golden):
    from holodeck_b200 import cyutils
    R, L = int(golden["nreals"]), int(golden["nloud"])
    ms, qs, zs = sort_indices(golden)
    hc2ss, hc2bg = cyutils.loudest_hc_from_sorted(golden["number"], golden["h2fdf"], R, L, ms, qs, zs,
                                                  counts=golden["counts_loud"])
    assert np.array_equal(hc2ss, golden["l1_hc2ss"])
    assert rel_err(hc2bg, golden["l1_hc2bg"]) < 1e-12


def test_loudest_hc_and_par_from_sorted_supplied_counts(holo, golden):
    from holodeck_b200 import cyutils
    R, L = int(golden["nreals"]), int(golden["nloud"])
    ms, qs, zs = sort_indices(golden)
    mt, mr, rz = [0.5 * (golden[kk][1:] + golden[kk][:-1]) for kk in ("mtot", "mrat", "redz")]
    hc2ss, hc2bg, lspar, bgpar, ssidx = cyutils.loudest_hc_and_par_from_sorted(
        golden["number"], golden["h2fdf"], R, L, mt, mr, rz, ms, qs, zs, counts=golden["counts_loud"])
    assert np.array_equal(hc2ss, golden["l2_hc2ss"])
    assert np.array_equal(ssidx, golden["l2_ssidx"])
    assert rel_err(hc2bg, golden["l2_hc2bg"]) < 1e-12
    assert rel_err(bgpar, golden["l2_bgpar"]) < 1e-10
    assert rel_err(lspar, golden["l2_lspar"]) < 1e-12


def test_loudest_redz_supplied_counts(holo, golden):
    from holodeck_b200 import cyutils
    R, L = int(golden["nreals"]), int(golden["nloud"])
    ms, qs, zs = sort_indices(golden)
    mt, mr, rz = [0.5 * (golden[kk][1:] + golden[kk][:-1]) for kk in ("mtot", "mrat", "redz")]
    hc2ss, hc2bg, sspar, bgpar = cyutils.loudest_hc_and_par_from_sorted_redz(
        golden["number"], golden["h2fdf"], R, L, mt, mr, rz, golden["par_redz"], golden["par_dcom"],
        golden["par_sepa"], golden["par_angs"], ms, qs, zs, counts=golden["counts_loud"])
    assert np.array_equal(hc2ss, golden["l3_hc2ss"])
    assert np.array_equal(sspar, golden["l3_sspar"])
    assert rel_err(hc2bg, golden["l3_hc2bg"]) < 1e-12
    assert np.array_equal(np.isnan(bgpar), np.isnan(golden["l3_bgpar"]))
    assert rel_err(bgpar, golden["l3_bgpar"]) < 1e-10


def test_ss_bg_hc_supplied_counts(holo, golden):
    from holodeck_b200 import cyutils
    R = int(golden["nreals"])
    mt, mr, rz = [0.5 * (golden[kk][1:] + golden[kk][:-1]) for kk in ("mtot", "mrat", "redz")]
    hc2ss, hc2bg, ssidx = cyutils.ss_bg_hc(golden["number"], golden["h2fdf"], R, counts=golden["counts_ssbg"])
    assert np.array_equal(hc2ss, golden["s1_hc2ss"])
    assert np.array_equal(ssidx, golden["s1_ssidx"])
    assert rel_err(hc2bg, golden["s1_hc2bg"]) < 1e-10
    hc2ss, hc2bg, ssidx, bgpar, sspar = cyutils.ss_bg_hc_and_par(golden["number"], golden["h2fdf"], R, mt, mr, rz,
                                                                counts=golden["counts_ssbg"])
    assert np.array_equal(hc2ss, golden["s2_hc2ss"])
    assert np.array_equal(ssidx, golden["s2_ssidx"])
    assert np.array_equal(sspar, golden["s2_sspar"])
    assert rel_err(hc2bg, golden["s2_hc2bg"]) < 1e-10
    assert rel_err(bgpar, golden["s2_bgpar"]) < 1e-8


def test_realizations_do_not_depend_on_partition(holo, golden_classic):
    """Philox is keyed on the GLOBAL realization index: R=8 in one launch == two launches of 4 (r0=0,4)."""
    from holodeck_b200 import cyutils
    gg = golden_classic
    ms, qs, zs = sort_indices(gg)
    full = cyutils.loudest_hc_from_sorted(gg["number"], gg["h2fdf"], 8, 3, ms, qs, zs, seed=42)
    lo = cyutils.loudest_hc_from_sorted(gg["number"], gg["h2fdf"], 4, 3, ms, qs, zs, seed=42, r0=0)
    hi = cyutils.loudest_hc_from_sorted(gg["number"], gg["h2fdf"], 4, 3, ms, qs, zs, seed=42, r0=4)
    assert np.array_equal(full[0], np.concatenate([lo[0], hi[0]], axis=1))
    assert np.array_equal(full[1], np.concatenate([lo[1], hi[1]], axis=1))
    g1 = cyutils.sam_poisson_gwb(gg["number"], gg["h2fdf"], 8, seed=7)
    g2 = np.concatenate([cyutils.sam_poisson_gwb(gg["number"], gg["h2fdf"], 4, seed=7, r0=rr) for rr in (0, 4)], axis=1)
    assert np.array_equal(g1, g2)
    assert not np.array_equal(g1, cyutils.sam_poisson_gwb(gg["number"], gg["h2fdf"], 8, seed=8))
    # ... and across the lock-step slot counts the kernel picks from R (600 -> 4 slots, 300 -> 2, 150 -> 1)
    big = cyutils.loudest_hc_from_sorted(gg["number"], gg["h2fdf"], 600, 2, ms, qs, zs, seed=9)
    halves = [cyutils.loudest_hc_from_sorted(gg["number"], gg["h2fdf"], 300, 2, ms, qs, zs, seed=9, r0=rr) for rr in (0, 300)]
    quarters = [cyutils.sam_poisson_gwb(gg["number"], gg["h2fdf"], 150, seed=9, r0=rr) for rr in (0, 150, 300, 450)]
    assert np.array_equal(big[0], np.concatenate([hh[0] for hh in halves], axis=1))
    assert np.array_equal(big[1], np.concatenate([hh[1] for hh in halves], axis=1))
    assert np.array_equal(cyutils.sam_poisson_gwb(gg["number"], gg["h2fdf"], 600, seed=9), np.concatenate(quarters, axis=1))


def test_parameter_variant_does_not_depend_on_partition(holo, golden_classic):
    """The parameter variant (`loudest_hc_and_par_from_sorted_redz`) gives a thread a QUAD of four consecutive GLOBAL
    realizations (Philox block keyed on (element, quad)): splits that cut through quads (r0 = 3, 10, 301) and the two
    CTA sizes (<= 64 quads: 64 threads; more: 256) must reproduce the single launch bit for bit."""
    from holodeck_b200 import cyutils
    gg = golden_classic
    ms, qs, zs = sort_indices(gg)
    mt, mr, rz = [0.5 * (gg[kk][1:] + gg[kk][:-1]) for kk in ("mtot", "mrat", "redz")]
    pars = (mt, mr, rz, gg["par_redz"], gg["par_dcom"], gg["par_sepa"], gg["par_angs"], ms, qs, zs)
    for R, cuts in ((17, (0, 3, 10, 17)), (700, (0, 301, 700))):
        full = cyutils.loudest_hc_and_par_from_sorted_redz(gg["number"], gg["h2fdf"], R, 3, *pars, seed=21)
        parts = [cyutils.loudest_hc_and_par_from_sorted_redz(gg["number"], gg["h2fdf"], b - a, 3, *pars, seed=21, r0=a)
                 for a, b in zip(cuts[:-1], cuts[1:])]
        for ii, axis in enumerate((1, 1, 2, 2)):        # hc2ss (F,R,L), hc2bg (F,R), sspar (4,F,R,L), bgpar (7,F,R)
            assert np.array_equal(full[ii], np.concatenate([pp[ii] for pp in parts], axis=axis), equal_nan=True), (R, ii)
    # a global offset shifts the realizations, it does not redraw them
    a = cyutils.loudest_hc_and_par_from_sorted_redz(gg["number"], gg["h2fdf"], 12, 2, *pars, seed=5, r0=0)
    b = cyutils.loudest_hc_and_par_from_sorted_redz(gg["number"], gg["h2fdf"], 7, 2, *pars, seed=5, r0=5)
    assert np.array_equal(a[1][:, 5:], b[1]) and np.array_equal(a[3][:, :, 5:], b[3], equal_nan=True)


def _same_distribution(got, ref, nboot=40, nsig=7.0):
    """per-frequency 5 / 50 / 95 % quantiles of two independent samples agree within bootstrap Monte-Carlo error"""
    R = ref.shape[1]
    for qq in (0.05, 0.5, 0.95):
        a, b = np.quantile(got, qq, axis=1), np.quantile(ref, qq, axis=1)
        boot = np.std([np.quantile(ref[:, np.random.default_rng(ii).integers(0, R, R)], qq, axis=1) for ii in range(nboot)], axis=0)
        if not np.all(np.abs(a - b) <= nsig * boot * np.sqrt(1.0 + R / got.shape[1]) + 1e-12 * np.abs(b)):
            return False
    return True


def test_fused_gwb_slots_of_the_loudest_pass(holo, golden_classic):
    """`gwb_nreals=`: the loudest pass also draws an independent realised GWB from the same staged records.  The
    loudest products must not change at all, whatever the number of fused slots and however they fall onto threads,
    warps and lock-step slots; the fused GWB is an independent sample of what `sam_poisson_gwb` draws (same Philox
    stream and keys, but the loudest variants stage longer passes, hence other superposition groups), does not depend
    on the loudest realizations it rides with, and honours its own global realization offset."""
    from holodeck_b200 import cyutils
    gg = golden_classic
    ms, qs, zs = sort_indices(gg)
    F = gg["number"].shape[-1]
    lam, hh = gg["number"].reshape(-1, F), gg["h2fdf"].reshape(-1, F)
    ref = cyutils.sam_poisson_gwb(gg["number"], gg["h2fdf"], 1500, seed=77)
    for R, Rg in ((5, 5), (100, 100), (96, 300), (700, 40)):      # 128-thread CTAs, mixed warps, 2 and 4 slots per thread
        alone = cyutils.loudest_hc_from_sorted(gg["number"], gg["h2fdf"], R, 3, ms, qs, zs, seed=11, r0=2)
        fused = cyutils.loudest_hc_from_sorted(gg["number"], gg["h2fdf"], R, 3, ms, qs, zs, seed=11, r0=2,
                                               gwb_nreals=Rg, gwb_seed=12, gwb_r0=1)
        assert len(fused) == 3 and fused[2].shape == (F, Rg)
        assert np.array_equal(fused[0], alone[0]) and np.array_equal(fused[1], alone[1])
    # the fused sample: independent of the loudest realizations riding along, offset-consistent, correctly distributed
    big = cyutils.loudest_hc_from_sorted(gg["number"], gg["h2fdf"], 4, 3, ms, qs, zs, seed=1, gwb_nreals=1500, gwb_seed=12)[2]
    other = cyutils.loudest_hc_from_sorted(gg["number"], gg["h2fdf"], 300, 2, ms, qs, zs, seed=2, gwb_nreals=200, gwb_seed=12,
                                           gwb_r0=1000)[2]
    assert np.array_equal(other, big[:, 1000:1200])
    assert _same_distribution(big, ref)
    mean = np.sum(lam * hh, axis=0)
    sig = np.sqrt(np.sum(lam * hh * hh, axis=0) / big.shape[1])
    assert np.all(np.abs(big.mean(axis=1) - mean) < 6 * sig + 1e-300)
    mt, mr, rz = [0.5 * (gg[kk][1:] + gg[kk][:-1]) for kk in ("mtot", "mrat", "redz")]
    pars = (mt, mr, rz, gg["par_redz"], gg["par_dcom"], gg["par_sepa"], gg["par_angs"], ms, qs, zs)
    R, Rg = 64, 1500
    alone = cyutils.loudest_hc_and_par_from_sorted_redz(gg["number"], gg["h2fdf"], R, 2, *pars, seed=5)
    fused = cyutils.loudest_hc_and_par_from_sorted_redz(gg["number"], gg["h2fdf"], R, 2, *pars, seed=5,
                                                        gwb_nreals=Rg, gwb_seed=6)
    for aa, ff in zip(alone, fused[:4]):
        assert np.array_equal(aa, ff, equal_nan=True)
    assert _same_distribution(fused[4], ref)
