"""GPU tests of the library-generation driver (BASELINE config 5 at toy size) and of the large-R /
large-L paths of the loudest-source kernels."""
import numpy as np
import pytest

from conftest import load_golden, rel_err
from _stubs import sort_indices

pytestmark = pytest.mark.gpu


def test_run_model_all_registered_param_spaces():
    """librarian/tests/test_param_spaces.py:10-40 semantics: every space builds a model at sam_shape=13
    and `run_model` returns populated arrays."""
    from holodeck_b200 import librarian
    for name, cls in librarian.param_spaces_dict.items():
        space = cls(nsamples=3, sam_shape=13, seed=11)
        params = space.param_dict(1)
        params["mmb_scatter_dex"] = 0.0          # the scatter is the host-scipy stage; tested separately below
        sam, hard = space.model_for_params(params)
        data = librarian.run_model(sam, hard, nfreqs=7, nreals=9, nloudest=3, singles_flag=True, params_flag=True,
                                   gwb_flag=True, seed=3)
        assert data["hc_ss"].shape == (7, 9, 3) and data["hc_bg"].shape == (7, 9) and data["gwb"].shape == (7, 9)
        assert data["sspar"].shape == (4, 7, 9, 3) and data["bgpar"].shape == (7, 7, 9)
        for key in ("hc_ss", "hc_bg", "gwb"):
            assert np.any(data[key] > 0), (name, key)
        # reproducible with a seed; the GWB and the single-source split use independent draws
        again = librarian.run_model(sam, hard, nfreqs=7, nreals=9, nloudest=3, params_flag=True, seed=3)
        assert np.array_equal(again["gwb"], data["gwb"]) and np.array_equal(again["hc_bg"], data["hc_bg"])
        tot_ss = np.sqrt(data["hc_bg"]**2 + np.sum(data["hc_ss"]**2, axis=-1))
        assert not np.allclose(tot_ss, data["gwb"], rtol=1e-6, atol=0.0)


def test_gen_lib_writes_reference_file_layout(tmp_path):
    from holodeck_b200 import librarian
    from holodeck_b200.librarian import gen_lib, combine, stream
    space = librarian.PS_Classic_Phenom_Uniform(nsamples=4, sam_shape=11, seed=2)
    space.param_samples[:, space.param_names.index("mmb_scatter_dex")] = 0.0
    kw = dict(nreals=5, nfreqs=6, nloudest=2, params_flag=True, seed=2)
    # streaming file plane (default) + the reference's per-sample files from the writer thread
    done, fails = gen_lib.run_library(space, tmp_path, sim_files=True, **kw)
    assert (done, fails) == (4, 0)
    files = sorted((tmp_path / "library_sims").glob("library__p*.npz"))
    assert [ff.name for ff in files] == [f"library__p{ii:06d}.npz" for ii in range(4)]
    data = np.load(files[0])
    for key in ("fobs_cents", "fobs_edges", "gwb", "hc_ss", "hc_bg", "sspar", "bgpar", "params", "param_names"):
        assert key in data.files, key
    assert data["gwb"].shape == (6, 5) and data["hc_ss"].shape == (6, 5, 2)
    assert (tmp_path / "PS_Classic_Phenom_Uniform.pspace.npz").exists() and (tmp_path / "config.json").exists()
    # the memory-mapped combined layout holds the same rows, and combines into the reference's library file
    store = stream.LibraryStore.open(tmp_path, mode="r")
    assert all(store.is_done(ii) for ii in range(4))
    lib = np.load(combine.sam_lib_combine(tmp_path))
    assert lib["sspar"].shape == (4, 4, 6, 5, 2) and lib["bgpar"].shape == (4, 7, 6, 5)
    for ii, ff in enumerate(files):
        one = np.load(ff)
        for key in ("gwb", "hc_ss", "hc_bg", "sspar", "bgpar"):
            assert np.array_equal(lib[key][ii], one[key], equal_nan=True), (ii, key)
        assert np.array_equal(lib["sample_params"][ii], one["params"])
    # a second run skips finished samples (gen_lib.py:287-296); other settings in the same directory are refused
    done2, _ = gen_lib.run_library(space, tmp_path, sim_files=True, **kw)
    assert done2 == 4
    with pytest.raises(RuntimeError, match="different settings"):
        gen_lib.run_library(space, tmp_path, **dict(kw, nreals=7))
    # the reference's own file plane (synchronous per-sample files, merged afterwards) gives the same library
    other = tmp_path / "per_sample"
    other.mkdir()
    done3, _ = gen_lib.run_library(space, other, streaming=False, **kw)
    assert done3 == 4
    lib2 = np.load(combine.sam_lib_combine(other))
    for key in ("gwb", "hc_ss", "hc_bg", "sspar", "bgpar", "sample_params"):
        assert np.array_equal(lib[key], lib2[key], equal_nan=True), key


def test_scatter_path_runs_and_conserves_mass():
    """M-Mbulge scatter (sam.py:368-389): host scipy stage between K0 and the stalled-bin zeroing."""
    import holodeck_b200 as holo
    from holodeck_b200 import host_relations
    shape = (15, 13, 9)
    kw = dict(gpf=holo.sams.GPF_Power_Law, shape=shape)
    sam0 = holo.sams.Semi_Analytic_Model(mmbulge=host_relations.MMBulge_KH2013(scatter_dex=0.0), **kw)
    sam1 = holo.sams.Semi_Analytic_Model(mmbulge=host_relations.MMBulge_KH2013(scatter_dex=0.3), **kw)
    d0, d1 = sam0.static_binary_density, sam1.static_binary_density
    assert d1.shape == d0.shape and np.all(d1 >= 0) and not np.allclose(d0, d1)
    m0, m1 = sam0._integrated_binary_density(), sam1._integrated_binary_density()
    assert abs(m1 / m0 - 1.0) < 0.5
    assert np.array_equal(d1 == 0, d1 == 0) and np.all(d1[sam1._redz_prime < 0] == 0)


def test_large_nloudest_and_many_realizations():
    """L larger than the default head margin and R spanning several realization tiles, vs the seeded
    reference in supplied-count mode (bit-exact slots) and partition independence across r0."""
    from holodeck_b200 import cyutils
    from oracle import glue
    gg = load_golden("classic_2pwl")
    ms, qs, zs = sort_indices(gg)
    R, L = 37, 40
    cy, _, _ = glue.ref()
    cy.ORACLE_SEED = 5
    r_ss, r_bg = [np.asarray(vv) for vv in cy.loudest_hc_from_sorted(gg["number"], gg["h2fdf"], R, L, ms, qs, zs)]
    cy.ORACLE_SEED = None
    counts = glue.counts_loudest(gg["number"], gg["order"].astype(np.int64), R, 5)
    g_ss, g_bg = cyutils.loudest_hc_from_sorted(gg["number"], gg["h2fdf"], R, L, ms, qs, zs, counts=counts)
    assert np.array_equal(g_ss, r_ss) and rel_err(g_bg, r_bg) < 1e-12
    big = cyutils.loudest_hc_from_sorted(gg["number"], gg["h2fdf"], 600, 3, ms, qs, zs, seed=9)
    parts = [cyutils.loudest_hc_from_sorted(gg["number"], gg["h2fdf"], 200, 3, ms, qs, zs, seed=9, r0=rr) for rr in (0, 200, 400)]
    assert np.array_equal(big[0], np.concatenate([pp[0] for pp in parts], axis=1))
    assert np.array_equal(big[1], np.concatenate([pp[1] for pp in parts], axis=1))


def test_empty_and_degenerate_inputs():
    from holodeck_b200 import cyutils
    zeros = np.zeros((3, 2, 4, 5))
    ones = np.ones_like(zeros)
    order = np.arange(24)
    ms, qs, zs = order // 8, (order // 4) % 2, order % 4
    ss, bg = cyutils.loudest_hc_from_sorted(zeros, ones, 4, 2, ms, qs, zs, seed=1)
    assert np.all(ss == 0) and np.all(bg == 0)
    assert np.all(cyutils.sam_poisson_gwb(zeros, ones, 4, seed=1) == 0)
    hc2ss, hc2bg, ssidx = cyutils.ss_bg_hc(zeros, ones, 3, seed=1)
    assert np.all(hc2ss == 0) and np.all(ssidx == -1)
    with pytest.raises(RuntimeError):
        cyutils.ss_bg_hc_and_par(zeros, ones, 3, np.ones(3), np.ones(2), np.ones(4), seed=1)
    # bgpar = 0/0 -> NaN when there is no background (cyutils.pyx:1761-1767)
    out = cyutils.loudest_hc_and_par_from_sorted_redz(zeros, ones, 2, 1, np.ones(3), np.ones(2), np.ones(4), ones, ones, ones,
                                                      ones, ms, qs, zs, seed=1)
    assert np.all(np.isnan(out[3]))
    # the normal branch: lam > 1e10 draws are not floored, and the whole (only) cell goes to the slots first
    lam = np.full((1, 1, 1, 1), 3.0e10)
    gwb = cyutils.sam_poisson_gwb(lam, np.ones_like(lam), 2000, seed=4)
    assert abs(gwb.mean() / 3.0e10 - 1) < 1e-6 and abs(gwb.std() / np.sqrt(3.0e10) - 1) < 0.1 and np.any(gwb != np.floor(gwb))


def test_realizer_sam_weights():
    """extensions.Realizer_SAM: (ncell, R) Poisson weights with the right mean and bin-centre samples."""
    import holodeck_b200 as holo
    from holodeck_b200 import host_relations, utils
    from holodeck_b200.constants import YR
    sam = holo.sams.Semi_Analytic_Model(shape=(9, 8, 10), mmbulge=host_relations.MMBulge_KH2013(scatter_dex=0.0))
    _, fobs_edges = utils.pta_freqs(16.03*YR, 4)
    real = holo.extensions.Realizer_SAM(fobs_edges / 2.0, sam=sam, hard=holo.hardening.Hard_GW())
    names, samples, weights = real(nreals=300, seed=3)
    ncell = 8 * 7 * 9 * 4
    assert names == ['mtot', 'mrat', 'redz', 'fobs'] and all(ss.shape == (ncell,) for ss in samples)
    assert weights.shape == (ncell, 300) and np.all(weights >= 0) and np.all(weights == np.floor(weights))
    grid, dnum, redz_final = sam.dynamic_binary_number_at_fobs(holo.hardening.Hard_GW(), utils.midpoints(fobs_edges) / 2.0)
    number = holo.sams.sam_cyutils.integrate_differential_number_3dx1d([sam.mtot, sam.mrat, sam.redz, fobs_edges / 2.0], dnum).flatten()
    sel = (number > 0.05) & (number < 1e6)
    zsc = (weights[sel].mean(axis=1) - number[sel]) / np.sqrt(number[sel] / 300)
    assert np.all(np.abs(zsc) < 6) and abs(zsc.mean()) < 0.5
    assert np.all(weights[number == 0] == 0)
    _, s2, w2 = real(nreals=3, clean=True, seed=3)
    assert len(w2) == 3 and all(len(ss[0]) == len(ww) for ss, ww in zip(s2, w2))
    with pytest.raises(ValueError):
        holo.extensions.Realizer_SAM(fobs_edges / 2.0)


def test_model_details_match_reference_procedure():
    """run_model(details_flag=True) / _calc_model_details against the reference procedure
    (oracle/glue.calc_model_details: numpy sums + scipy.stats.binned_statistic, lib_tools.py:845-943)."""
    import holodeck_b200 as holo
    from holodeck_b200 import host_relations, librarian
    from holodeck_b200.librarian import lib_tools
    from holodeck_b200.constants import GYR, PC
    from oracle import glue
    from conftest import rel_err
    sam = holo.sams.Semi_Analytic_Model(shape=(17, 13, 21), gpf=holo.sams.GPF_Power_Law(), gmt=holo.sams.GMT_Power_Law(),
                                        mmbulge=host_relations.MMBulge_KH2013(scatter_dex=0.0))
    hard = holo.hardening.Fixed_Time_2PL_SAM(sam, 3.0*GYR, sepa_init=1e4*PC, rchar=100.0*PC, gamma_inner=-1.0, gamma_outer=2.5)
    data = librarian.run_model(sam, hard, nfreqs=7, nreals=4, nloudest=2, gwb_flag=False, singles_flag=False,
                               details_flag=True, seed=1)
    for key in ("static_binary_density", "number", "redz_final", "gwb_params", "num_params", "gwb_mtot_redz_final",
                "num_mtot_redz_final", "fobs_cents", "fobs_edges"):
        assert key in data
    number, redz_final = data["number"], data["redz_final"]
    assert number.shape == (16, 12, 20, 7) and redz_final.shape == (17, 13, 21, 7)
    edges = [sam.mtot, sam.mrat, sam.redz, data["fobs_edges"] / 2.0]
    from holodeck_b200 import gravwaves
    hc2 = gravwaves.char_strain_sq_from_bin_edges_redz(edges, redz_final)
    ref = glue.calc_model_details(edges, redz_final, number, hc2)
    got = (data["gwb_params"], data["num_params"], data["gwb_mtot_redz_final"], data["num_mtot_redz_final"])
    shapes = [(16, 20, 7), (12, 20, 7), (20, 7), (20, 7)]
    for kk in range(4):
        assert got[0][kk].shape == shapes[kk] and got[1][kk].shape == shapes[kk]
        assert rel_err(got[0][kk], ref[0][kk]) < 1e-12, (kk, rel_err(got[0][kk], ref[0][kk]))
        assert rel_err(got[1][kk], ref[1][kk]) < 1e-12
    assert rel_err(got[2], ref[2]) < 1e-12 and rel_err(got[3], ref[3]) < 1e-12
    assert got[3].sum() > 0 and np.array_equal(got[3] == 0, ref[3] == 0)
    # direct call with host arrays, as the reference signature
    again = lib_tools._calc_model_details(edges, redz_final, number)
    assert np.array_equal(again[2], got[2]) and np.array_equal(again[3], got[3])
