"""M-Mbulge scatter: the restatement of scipy's Clough-Tocher machinery that the K6 kernels
implement is pinned against the installed scipy (the reference reaches it through
`sp.interpolate.CloughTocher2DInterpolator`, holodeck/sams/sam.py:1362-1367), and the whole device pipeline
against the reference procedure `add_scatter_to_masses` (oracle/glue.py restatement of sam.py:1291-1394)."""
import numpy as np
import pytest

from conftest import rel_err


def small_case(M=14, Q=11, Z=5, seed=0):
    from holodeck_b200.constants import MSOL
    rng = np.random.default_rng(seed)
    mtot = np.logspace(*np.log10([1.0e5*MSOL, 1.0e11*MSOL]), M)
    mrat = np.logspace(*np.log10([1e-3, 1.0]), Q)
    lm = np.log10(mtot / MSOL)[:, None, None]
    lq = np.log10(mrat)[None, :, None]
    zz = np.linspace(0.1, 3.0, Z)[None, None, :]
    dens = 1e-3 * np.exp(-0.5*((lm - 8.0 - 0.3*zz)/0.9)**2) * (1.0 + 0.3*lq) * (1 + zz)**-1.5 * rng.uniform(0.7, 1.3, (M, Q, Z))
    dens[:2] = 0.0                       # an exactly-empty corner, as in real SAMs
    dens[:, :, 0] *= 1e4                 # one slice with gradients large enough to need several sweeps
    return mtot, mrat, np.ascontiguousarray(dens)


def test_geometry_and_restatement_against_scipy():
    import scipy.interpolate
    from holodeck_b200.sams import scatter
    from oracle import scatter_port as port
    mtot, mrat, dens = small_case()
    geo = scatter.scatter_geometry(mtot, mrat, refine=4)
    npts, G = geo["npts"], geo["G"]
    assert G == 4 * mtot.size and geo["geo"].shape == (G * G,)
    # dependency levels: no two vertices of a level are neighbours, every lower neighbour sits in a lower level
    level = np.empty(npts, dtype=int)
    for lv in range(geo["level_ptr"].size - 1):
        level[geo["order"][geo["level_ptr"][lv]:geo["level_ptr"][lv+1]]] = lv
    for ip in range(npts):
        nb = geo["indices"][geo["indptr"][ip]:geo["indptr"][ip+1]]
        assert np.all(level[nb] != level[ip]) and np.all(level[nb[nb < ip]] < level[ip])
    gx, gy = np.meshgrid(geo["mgrid_log10"], geo["mgrid_log10"], indexing='ij')
    niters = []
    for zz in range(dens.shape[2]):
        data = dens[:, :, zz].ravel()
        interp = scipy.interpolate.CloughTocher2DInterpolator(geo["tri"], data)
        g_seq, n_seq = port.gradients_sequential(geo, data)
        g_lev, n_lev = port.gradients_levels(geo, data)
        niters.append(n_seq)
        assert n_seq == n_lev and np.array_equal(g_seq, g_lev)             # level schedule == sequential sweep
        scale = np.abs(interp.grad).max() + 1e-300
        assert np.abs(g_seq - interp.grad[:, 0, :]).max() <= 1e-13 * scale  # == scipy's estimator
        ref = interp((gx, gy)).ravel()
        got = port.clough_tocher(geo, data, g_lev)
        assert np.array_equal(np.isnan(ref), np.isnan(got))
        ok = ~np.isnan(ref)
        assert np.abs(got[ok] - ref[ok]).max() <= 1e-12 * np.abs(ref[ok]).max()
    assert max(niters) > 1          # the multi-sweep branch is exercised


def test_port_pipeline_matches_reference_procedure():
    import scipy.stats
    from holodeck_b200.sams import scatter
    from oracle import glue, scatter_port as port
    mtot, mrat, dens = small_case(M=10, Q=9, Z=3, seed=1)
    geo = scatter.scatter_geometry(mtot, mrat, refine=4)
    weights = scatter._get_rolled_weights(geo["mgrid_log10"], scipy.stats.norm(loc=0.0, scale=0.3))
    assert rel_err(weights, glue._get_rolled_weights(geo["mgrid_log10"], scipy.stats.norm(loc=0.0, scale=0.3))) == 0
    got = port.add_scatter_port(geo, weights, dens)
    ref = glue.add_scatter_to_masses(mtot, mrat, dens, 0.3)
    assert rel_err(got, ref) < 1e-11


@pytest.mark.gpu
def test_device_scatter_matches_reference_procedure():
    from holodeck_b200.sams import scatter
    from oracle import glue
    for (M, Q, Z, seed, dex) in [(14, 11, 5, 0, 0.3), (23, 17, 9, 2, 0.15)]:
        mtot, mrat, dens = small_case(M, Q, Z, seed)
        ref = glue.add_scatter_to_masses(mtot, mrat, dens, dex)
        got = scatter.add_scatter_to_masses(mtot, mrat, dens, dex)
        assert got.shape == ref.shape
        assert rel_err(got, ref) < 1e-10, rel_err(got, ref)
        # total "mass" moves by a small fraction only (sam.py:381-389 logs it)
        assert abs(got.sum() / dens.sum() - 1.0) < 0.2
    with pytest.raises(ValueError):
        bad = dens.copy()
        bad[3, 3, 1] = np.nan
        scatter.add_scatter_to_masses(mtot, mrat, bad, dex)
