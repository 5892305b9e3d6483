"""CPU tests of the host-side mirror of the reference API (no kernels are launched)."""
import numpy as np
import pytest

import holodeck_b200 as holo
from holodeck_b200 import utils, sams, host_relations, hardening
from holodeck_b200.constants import MSOL, YR, GYR


def test_semi_analytic_model_grid_and_components():
    sam = sams.Semi_Analytic_Model()
    assert sam.shape == (91, 81, 101)
    assert np.isclose(sam.mtot[0], 1e4 * MSOL) and np.isclose(sam.mtot[-1], 1e12 * MSOL)
    assert np.isclose(sam.mrat[0], 1e-3) and sam.mrat[-1] == 1.0 and np.isclose(sam.redz[-1], 10.0)
    # gpf=None -> GMR_Illustris and no GMT (sam.py:168-174)
    assert isinstance(sam._gmr, sams.GMR_Illustris) and sam._gpf is None and sam._gmt is None
    assert isinstance(sam._mmbulge, host_relations.MMBulge_KH2013)
    sam2 = sams.Semi_Analytic_Model(gpf=sams.GPF_Power_Law, shape=(7, None, 9))
    assert sam2.shape == (7, 81, 9) and isinstance(sam2._gmt, sams.GMT_Power_Law) and sam2._gmr is None
    with pytest.raises(ValueError):
        sams.Semi_Analytic_Model(gpf=sams.GPF_Power_Law, gmr=sams.GMR_Illustris)
    with pytest.raises(ValueError):
        sams.Semi_Analytic_Model(bogus=1)
    with pytest.raises(ValueError):
        sams.Semi_Analytic_Model(gsmf=object())
    mstar_pri, mstar_rat, mstar_tot, redz = sam2.mass_stellar()
    assert mstar_pri.shape == sam2.shape and np.all(mstar_rat <= 1.0 + 1e-12) and np.all(mstar_tot >= mstar_pri)


def test_components_are_callable_like_the_reference():
    """sams/tests/test_components.py:52-133 -- every component evaluates on random inputs."""
    rng = np.random.default_rng(0)
    mstar = 10.0 ** rng.uniform(9, 12, 50) * MSOL
    mrat = 10.0 ** rng.uniform(-2, 0, 50)
    redz = rng.uniform(0, 5, 50)
    for gsmf in (sams.GSMF_Schechter(), sams.GSMF_Double_Schechter()):
        vals = gsmf(mstar, redz)
        assert vals.shape == mstar.shape and np.all(vals > 0)
    assert np.all(sams.GPF_Power_Law()(mstar, mrat, redz) <= 1.0)
    assert np.all(sams.GMT_Power_Law()(mstar, mrat, redz) > 0)
    assert np.all(sams.GMR_Illustris()(mstar, mrat, redz) > 0)
    with pytest.raises(TypeError):
        sams.components._Galaxy_Stellar_Mass_Function()
    with pytest.raises(ValueError):
        sams.GPF_Power_Law(max_frac=1.5)
    zp, tau = sams.GMT_Power_Law().zprime(mstar, mrat, redz)
    assert np.all((zp < redz) | (zp == -1.0)) and np.all(tau > 0)
    # frac_norm from frac_norm_allq (components.py:532-539)
    gpf = sams.GPF_Power_Law(frac_norm_allq=0.03, qgamma=0.5)
    assert np.isclose(gpf._frac_norm, 0.03 / ((1.0 - 0.25**1.5) / 1.5))


def test_unfusable_components_are_rejected_not_emulated():
    class MyGSMF(sams.GSMF_Schechter):
        def __call__(self, mstar, redz):
            return np.ones_like(mstar)
    sam = sams.Semi_Analytic_Model(gsmf=MyGSMF, shape=6)
    with pytest.raises(NotImplementedError):
        sam._kernel_params()
    # overriding a HELPER the host __call__ uses would be ignored by the kernel (it reads `_kernel_params()` only)
    class HelperGSMF(sams.GSMF_Schechter):
        def _phi_func(self, redz):
            return 1.0 + 0.0 * redz
    with pytest.raises(NotImplementedError):
        sams.Semi_Analytic_Model(gsmf=HelperGSMF, shape=6)._kernel_params()

    # ... while a subclass that only sets new defaults is the same closed form
    class MyDefaults(sams.GSMF_Schechter):
        def __init__(self):
            super().__init__(phi0=-2.5)
    assert sams.Semi_Analytic_Model(gsmf=MyDefaults, shape=6)._kernel_params().gsmf[0] == -2.5
    ok = sams.Semi_Analytic_Model(gsmf=sams.GSMF_Double_Schechter, gpf=sams.GPF_Power_Law, shape=6)._kernel_params()
    assert ok.gsmf_kind == 1 and ok.use_gmr == 0 and ok.has_gmt == 1 and ok.mmb[1] == 1.17


def test_pta_freqs_and_utils():
    cents, edges = utils.pta_freqs(16.03 * YR, 40)
    assert cents.size == 40 and edges.size == 41
    assert np.allclose(cents, np.arange(1, 41) / (16.03 * YR)) and np.allclose(utils.midpoints(edges), cents)
    assert utils.isinteger(5) and utils.isinteger(np.int64(5)) and not utils.isinteger(5.0)
    m1, m2 = utils.m1m2_from_mtmr(10.0, 0.25)
    assert np.isclose(m1, 8.0) and np.isclose(m2, 2.0)
    assert utils.redz_after(1e30, redz=1.0) == -1.0
    zz = utils.redz_after(np.array([0.0, 1.0 * GYR]), redz=np.array([1.0, 1.0]))
    assert np.isclose(zz[0], 1.0) and 0 < zz[1] < 1.0
    with pytest.raises(ValueError):
        utils.redz_after(1.0)
    assert utils.get_subclass_instance(None, sams.GSMF_Schechter, sams.components._Galaxy_Stellar_Mass_Function) is not None
    with pytest.raises(ValueError):
        utils.get_subclass_instance(3, None, sams.components._Galaxy_Stellar_Mass_Function)


def test_hard_gw_and_gwb_ideal_host_paths():
    dadt = hardening.Hard_GW.dadt(1e9 * MSOL, 0.5, 3.0e16)
    assert dadt < 0
    assert hardening.Hard_GW.deda(1e17, 0.5) > 0
    with pytest.raises(NotImplementedError):
        sams.Semi_Analytic_Model(shape=5).dynamic_binary_number_at_fobs(hardening.Hard_GW(), np.array([1e-9]), use_cython=False)


def test_param_space_and_sharding():
    from holodeck_b200 import librarian, dist
    sp1 = librarian.PS_Classic_Phenom_Uniform(nsamples=16, sam_shape=9, seed=7)
    sp2 = librarian.PS_Classic_Phenom_Uniform(nsamples=16, sam_shape=9, seed=7)
    assert np.array_equal(sp1.param_samples, sp2.param_samples) and sp1.param_samples.shape == (16, 6)
    # latin hypercube: one sample per stratum in each dimension
    strata = np.sort(np.floor(sp1._uniform_samples * 16), axis=0)
    assert np.array_equal(strata, np.tile(np.arange(16.0)[:, None], (1, 6)))
    lo, hi = sp1.extrema[:, 0], sp1.extrema[:, 1]
    assert np.all(sp1.param_samples >= lo) and np.all(sp1.param_samples <= hi)
    assert set(sp1.default_params()) == set(sp1.param_names)
    pp = sp1.normalized_params(0.5)
    assert np.isclose(pp["hard_time"], 5.55)
    # sample sharding: a partition of all samples, identical on every rank
    parts = [dist.sample_indices(37, seed=3, rank=rr, size=4) for rr in range(4)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(37)) and max(map(len, parts)) - min(map(len, parts)) <= 1
    slices = [dist.realization_slice(1000, rr, 8) for rr in range(8)]
    assert slices[0] == (0, 125) and slices[-1] == (875, 125)
    ragged = [dist.realization_slice(10, rr, 4) for rr in range(4)]
    assert [ss[1] for ss in ragged] == [3, 3, 2, 2] and [ss[0] for ss in ragged] == [0, 3, 6, 8]
    ext = librarian.PS_Classic_Phenom_Astro_Extended(nsamples=4, seed=1)
    assert ext.param_names[2] == "gsmf_phi0_log10"     # renamed from `gsmf_phi0` (lib_tools.py:24-26)
    with pytest.raises(RuntimeError):
        librarian.run_model(None, None, gwb_flag=False, singles_flag=False)


def test_library_combine_contract(tmp_path):
    """combine.sam_lib_combine: per-sample npz files -> one library with the reference's datasets;
    failure files become NaN rows (combine.py:204-205, 404-417)."""
    import holodeck_b200 as holo
    from holodeck_b200.librarian import combine, lib_tools, DIRNAME_LIBRARY_SIMS
    space = holo.librarian.PS_Classic_Phenom_Uniform(nsamples=5, seed=3)
    sims = tmp_path / DIRNAME_LIBRARY_SIMS
    sims.mkdir()
    space.save(tmp_path)
    rng = np.random.default_rng(0)
    F, R, L = 6, 4, 3
    fc = np.arange(1, F + 1) / 5e8
    fe = (np.arange(F + 1) + 0.5) / 5e8
    want = {}
    for pnum in range(5):
        fname = lib_tools._get_sim_fname(sims, pnum)
        if pnum == 2:
            np.savez(fname, fail="boom")
            continue
        want[pnum] = dict(gwb=rng.uniform(size=(F, R)), hc_ss=rng.uniform(size=(F, R, L)), hc_bg=rng.uniform(size=(F, R)),
                          sspar=rng.uniform(size=(4, F, R, L)), bgpar=rng.uniform(size=(7, F, R)))
        np.savez(fname, fobs_cents=fc, fobs_edges=fe, params=space.param_samples[pnum], param_names=space.param_names, **want[pnum])
    lib_path = combine.sam_lib_combine(tmp_path, holo.log)
    assert lib_path.exists() and lib_path.stem == "sam-library"
    if lib_path.suffix == ".npz":
        lib = np.load(lib_path)
        names = [nn.decode() for nn in lib["attrs/param_names"]]
    else:
        import h5py
        lib = {kk: vv[()] for kk, vv in h5py.File(lib_path, "r").items()}
        names = [nn.decode() for nn in h5py.File(lib_path, "r").attrs["param_names"]]
    assert names == list(space.param_names)
    assert lib["gwb"].shape == (5, F, R) and lib["hc_ss"].shape == (5, F, R, L) and lib["hc_bg"].shape == (5, F, R)
    assert lib["sspar"].shape == (5, 4, F, R, L) and lib["bgpar"].shape == (5, 7, F, R)
    assert np.array_equal(lib["fobs_cents"], fc) and np.array_equal(lib["fobs_edges"], fe)
    for pnum, dd in want.items():
        for kk, vv in dd.items():
            assert np.array_equal(lib[kk][pnum], vv)
        assert np.array_equal(lib["sample_params"][pnum], space.param_samples[pnum])
    assert np.all(np.isnan(lib["gwb"][2])) and np.all(np.isnan(lib["hc_ss"][2])) and np.all(np.isnan(lib["sample_params"][2]))
    assert combine.sam_lib_combine(tmp_path, holo.log) is None                      # exists: not recreated
    assert combine.sam_lib_combine(tmp_path, holo.log, recreate=True, gwb_only=True).stem == "sam-library_gwb-only"
    (sims / "library__p000004.npz").unlink()
    with pytest.raises(ValueError):
        combine.sam_lib_combine(tmp_path, holo.log, recreate=True)


def test_comoving_distance_table_against_independent_quadrature():
    """The Hermite table the strain kernel reads (`_lib.dc_table_host`, include/holo_b200.h `dc_table`) against the
    oracle's adaptive quadrature: the reference takes d_c from astropy/cosmopy (gravwaves.py:718); 1e-13 here."""
    from holodeck_b200 import _lib, cosmo
    from oracle import glue
    tab, n, wmax = _lib.dc_table_host(float(cosmo.Om0))
    assert tab.shape == (n + 1, 2) and 10.0 < (1.0 / (1.0 - wmax))**2 - 1.0 <= 20.0 + 1e-9
    hh = wmax / n
    rng = np.random.default_rng(3)
    zs = np.concatenate([10**rng.uniform(-6, np.log10(19.9), 300), [1e-3, 10.0]])
    sq = np.sqrt(1 + zs)
    ww = zs / (sq * (sq + 1))
    uu = ww / hh
    ii = uu.astype(int)
    tt = uu - ii
    r0, d0, r1, d1 = tab[ii, 0], tab[ii, 1], tab[ii + 1, 0], tab[ii + 1, 1]
    omt = 1 - tt
    rr = ((1 + 2*tt) * omt * omt) * r0 + (tt * omt * omt) * d0 + (tt * tt * (3 - 2*tt)) * r1 - (tt * tt * omt) * d1
    got = cosmo.hubble_distance * ww * rr
    oc = glue.OracleCosmo()
    want = np.array([oc.comoving_distance(zz) for zz in zs])
    assert np.max(np.abs(got / want - 1.0)) < 1e-13


def test_loudest_retry_schedule_fits_the_resolver():
    """the bucket of a retry is sized from its head margin and never exceeds what `holo_loudest` accepts
    (RES_WARPS * 2 * cap * 16 B <= 200 KB -> cap <= 1600)."""
    from holodeck_b200 import cyutils
    for L in (1, 5, 10, 100, 400):
        prev = 0.0
        for attempt in range(1, cyutils._MAX_RETRY):
            margin, cap = cyutils._retry_schedule(L, attempt)
            assert cap <= 1600 and 4 * 2 * cap * 16 <= 200 * 1024
            assert margin >= prev
            prev = margin
            if cap < cyutils._MAX_BUCKET_CAP:
                assert cap >= 2.0 * (L + margin) + 64.0       # the expected L + margin events fit with slack
            else:
                assert L + margin <= cap                      # clamped: the head is cut to what the bucket holds


def test_streaming_store_and_combine(tmp_path):
    """N2: rows written through `AsyncSampleWriter` into the memory-mapped combined layout come back from
    `sam_lib_combine` exactly like rows merged from per-sample files; failures are NaN rows; a store with a different
    layout is refused (resume must not mix shapes)."""
    import holodeck_b200 as holo
    from holodeck_b200.librarian import combine, stream
    S, F, R, L = 6, 5, 4, 3
    space = holo.librarian.PS_Classic_Phenom_Uniform(nsamples=S, seed=3)
    space.save(tmp_path)
    fc = np.arange(1, F + 1) / 5e8
    fe = (np.arange(F + 1) + 0.5) / 5e8
    store = stream.LibraryStore.create(tmp_path, S, F, R, L, True, True, True, fc, fe)
    assert stream.LibraryStore.exists(tmp_path) and not store.is_done(0)
    with pytest.raises(RuntimeError):
        stream.LibraryStore.create(tmp_path, S, F, R + 1, L, True, True, True, fc, fe)
    rng = np.random.default_rng(1)
    want = {}
    writer = stream.AsyncSampleWriter(stream.LibraryStore.open(tmp_path), nslots=2)
    for pnum in (4, 0, 3, 1, 5):
        want[pnum] = dict(gwb=rng.uniform(size=(F, R)), hc_ss=rng.uniform(size=(F, R, L)), hc_bg=rng.uniform(size=(F, R)),
                          sspar=rng.uniform(size=(4, F, R, L)), bgpar=rng.uniform(size=(7, F, R)))
        writer.submit(pnum, dict(want[pnum], fobs_cents=fc))
    writer.close()
    with pytest.raises(ValueError, match="sample number 2"):
        combine.sam_lib_combine(tmp_path, holo.log)                 # sample 2 has not run
    writer = stream.AsyncSampleWriter(stream.LibraryStore.open(tmp_path))
    writer.submit_failure(2, "boom")
    writer.close()
    again = stream.LibraryStore.open(tmp_path)
    assert again.is_done(4) and not again.is_done(2)                # a failure is re-attempted by the next run
    lib_path = combine.sam_lib_combine(tmp_path, holo.log)
    assert lib_path.suffix == ".npz"
    lib = np.load(lib_path)
    for pnum, dd in want.items():
        for kk, vv in dd.items():
            assert np.array_equal(lib[kk][pnum], vv), (pnum, kk)
        assert np.array_equal(lib["sample_params"][pnum], space.param_samples[pnum])
    assert np.all(np.isnan(lib["gwb"][2])) and np.all(np.isnan(lib["sspar"][2])) and np.all(np.isnan(lib["sample_params"][2]))
    assert np.array_equal(lib["fobs_edges"], fe)
    assert "boom" in (tmp_path / stream.DIRNAME_LIBRARY_STORE / "failures.log").read_text()


def test_gen_lib_config_is_saved_and_checked_on_resume(tmp_path):
    from holodeck_b200.librarian import gen_lib
    cfg = dict(param_space="PS_Classic_Phenom_Uniform", nsamples=4, nreals=3, nfreqs=5, nloudest=2, pta_dur=16.03,
               sam_shape=None, gwb_flag=True, ss_flag=True, params_flag=False, seed=9)
    fname = gen_lib._check_config(tmp_path, cfg)
    assert fname.name == "config.json" and fname.exists()
    gen_lib._check_config(tmp_path, dict(cfg))                      # same settings: a resume
    with pytest.raises(RuntimeError, match="nreals"):
        gen_lib._check_config(tmp_path, dict(cfg, nreals=7))
    gen_lib._check_config(tmp_path, dict(cfg, nreals=7), resume_ok=False)     # --recreate starts over
