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


def _run_step_program(prog, data, maxiter=400, tol=1e-6):
    """numpy emulation of `ct_gradients_kernel` consuming the step program (neighbour l summed with l + 4, then an
    xor-butterfly over four lanes; doubled edge vectors and the negated inverse matrix, as the records hold them)"""
    from holodeck_b200.sams.scatter import GS_LANES, GS_SLOTS, step_edges
    npts = data.size
    yy = np.zeros((npts, 2))
    steps = [step_edges(rec) for rec in prog]
    for it in range(maxiter):
        err = 0.0
        acc = np.zeros((GS_SLOTS, GS_LANES, 2))
        for vip, flags, nb, ee, qq in steps:
            vv = np.maximum(vip, 0)[:, None]
            df2x2 = -ee[..., 0] * yy[nb, 0] - ee[..., 1] * yy[nb, 1]
            num = 6 * (data[vv] - data[nb]) - df2x2
            pp = np.stack([num * ee[..., 2], num * ee[..., 3]], axis=-1)
            acc = pp if flags & 1 else acc + pp
            if flags & 2:
                tt = acc.copy()
                for off in (4, 2, 1):                                   # xor butterfly, as the shuffles
                    tt = tt + tt[:, np.arange(GS_LANES) ^ off, :]
                tot = tt[:, 0, :]
                for jj, vx in enumerate(vip):
                    if vx < 0:
                        continue
                    r0 = qq[jj, 0] * tot[jj, 0] + qq[jj, 1] * tot[jj, 1]
                    r1 = qq[jj, 2] * tot[jj, 0] + qq[jj, 3] * tot[jj, 1]
                    change = max(abs(yy[vx, 0] - r0), abs(yy[vx, 1] - r1))
                    yy[vx] = (r0, r1)
                    err = max(err, change / max(1.0, abs(r0), abs(r1)))
        if err < tol:
            return yy, it + 1
    return yy, 0


def test_step_program_reproduces_the_sequential_sweep():
    """The flattened step program the kernel streams with TMA (geometry only) executes the same Gauss-Seidel sweep as
    scipy's sequential loop: same sweep counts, gradients to rounding -- including a vertex with more neighbours than
    lanes (several rounds) and levels wider than one step."""
    from holodeck_b200.sams import scatter
    from oracle import scatter_port as port
    for (M, Q, Z, seed) in ((14, 11, 4, 0), (40, 37, 2, 3)):
        mtot, mrat, dens = small_case(M, Q, Z, seed)
        geo = scatter.scatter_geometry(mtot, mrat, refine=4)
        prog = geo["program"]
        deg = np.diff(geo["indptr"])
        assert prog.dtype.itemsize == 11264 and prog["ids"][0, 0, 3] & 1 and prog["ids"][-1, 0, 3] & 2
        # every directed edge appears exactly once, under its own vertex
        pairs = set()
        for rec in prog:
            vip, flags, nb, ee, qq = scatter.step_edges(rec)
            ip = np.repeat(vip[:, None], scatter.GS_LANES, axis=1)
            real = (ip >= 0) & (nb != ip)
            assert not ee[~real].any()                                     # absent edges carry zero coefficients
            for aa, bb in zip(ip[real], nb[real]):
                assert (int(aa), int(bb)) not in pairs
                pairs.add((int(aa), int(bb)))
        assert len(pairs) == geo["indices"].size
        if M == 40:
            assert deg.max() > scatter.GS_LANES                       # the multi-round path is exercised
        for zz in range(dens.shape[2]):
            data = dens[:, :, zz].ravel()
            g_seq, n_seq = port.gradients_sequential(geo, data)
            g_prog, n_prog = _run_step_program(prog, data)
            assert n_prog == n_seq
            assert np.abs(g_prog - g_seq).max() <= 1e-12 * (np.abs(g_seq).max() + 1e-300)


def test_scatter_product_blocks_cover_what_the_back_interpolation_reads():
    """`_gemm_blocks`: the row/column blocks for which the scatter product is evaluated contain every corner the
    bilinear back-interpolation touches (anything else of the scattered grid stays unwritten), for any block count."""
    from holodeck_b200.sams import scatter
    for (M, Q) in ((14, 11), (40, 37)):
        mtot, mrat, _ = small_case(M, Q, 2, 0)
        geo = scatter.scatter_geometry(mtot, mrat, refine=4)
        G, i0, i1 = geo["G"], np.asarray(geo["i0"]), np.asarray(geo["i1"])
        assert i0.max() <= G - 2 and i1.max() <= G - 2 and i0.min() >= 0 and i1.min() >= 0
        for nblk in (1, 3, 4, 9):
            blocks = scatter._gemm_blocks(i0, i1, G, nblk)
            cover = np.zeros((G, G), dtype=bool)
            for k0, k1, b0, b1 in blocks:
                assert 0 <= k0 < k1 <= G and 0 <= b0 < b1 <= G
                assert not cover[:, b0:b1].any()                       # column blocks do not overlap
                cover[k0:k1, b0:b1] = True
            for da in (0, 1):
                for db in (0, 1):
                    assert cover[i0 + da, i1 + db].all()
        # the point of the exercise: the needed region is a small part of the grid
        assert cover.mean() < 0.6


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
def test_gradients_kernel_against_the_step_program_emulation_at_sweep_limits():
    """`ct_gradients_kernel` alone, with the sweep limit set to 1, 2, 3 and 400 on grids with odd and even step counts:
    gradients and sweep counts must follow the numpy execution of the same step program.  This pins the kernel's ring
    protocol where it is most delicate -- the speculative fetch of the next sweep's first record, the producer warp's stop
    at convergence or at the limit, the register sets swapping roles on an odd step count -- and the `niter = 0` (not
    converged) report."""
    import torch
    from holodeck_b200 import _lib
    from holodeck_b200.sams import scatter
    lib = _lib.require_gpu()
    seen = set()
    for (M, Q, Z, seed) in ((14, 11, 4, 0), (15, 11, 3, 1), (23, 17, 2, 2)):
        mtot, mrat, dens = small_case(M, Q, Z, seed)
        geo = scatter.scatter_geometry(mtot, mrat, refine=4)
        dev = scatter._device_geometry(mtot, mrat, 4)
        npts = geo["npts"]
        seen.add(int(geo["program"].size) % 2)
        data = _lib.to_dev(dens).reshape(npts, Z)
        for maxiter in (1, 2, 3, 400):
            grad = _lib.empty((npts, 2, Z))
            niter = torch.full((Z,), -7, dtype=torch.int32, device=data.device)
            rc = lib.holo_scatter_gradients(npts, Z, _lib.ptr(dev["program"]), dev["nsteps"], _lib.ptr(data), maxiter, 1e-6,
                                            _lib.ptr(grad), _lib.ptr(niter), _lib.stream())
            _lib.check(rc, "gradients")
            got, gn = _lib.to_host(grad), niter.cpu().numpy()
            for zz in range(Z):
                want, wn = _run_step_program(geo["program"], dens.reshape(npts, Z)[:, zz], maxiter=maxiter)
                assert gn[zz] == wn, (M, Q, zz, maxiter, gn[zz], wn)
                assert np.abs(got[:, :, zz] - want).max() <= 1e-12 * (np.abs(want).max() + 1e-300)
    assert seen == {0, 1}, "both parities of the step count must be exercised"


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
