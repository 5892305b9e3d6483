"""M-Mbulge scatter of the binary density on the device (reference ``sams/sam.py:1291-1394``).

``add_scatter_to_masses`` keeps the reference's signature and procedure:

1. interpolate every redshift slice of ``dens`` from the irregular ``(log10 m1, log10 m2)`` images of the
   ``(mtot, mrat)`` grid onto a regular ``G x G`` grid (Clough-Tocher, ``G = refine * M``), replacing NaN /
   negative values by the nearest data point's value;
2. redistribute with binned normal weights (``utils._get_rolled_weights``; as the reference's two
   ``_scatter_with_weights`` calls actually do: twice along the primary-mass axis, see below);
3. interpolate (bilinear) back to the original grid points.

What depends on the *geometry* only -- the Delaunay triangulation, the location of the regular-grid points
in it, nearest vertices, the normal matrices of scipy's gradient estimator and the dependency levels of its
Gauss-Seidel sweep, the scatter weights -- is computed once per ``(mtot, mrat, refine, scatter)`` on the
host (scipy.spatial, as the reference does inside ``CloughTocher2DInterpolator``) and cached: every SAM of
a library shares it.  All data-dependent arithmetic runs on the GPU for the Z slices at once
(``csrc/holo_scatter.cu``: K6a gradients, K6b Clough-Tocher + fill, K6c bilinear; the scatter product
is one cuBLAS DGEMM on the ``(G, G*Z)`` layout).  There is no CPU path for the data.
"""
import ctypes as C

import numpy as np

from .. import _lib, utils

__all__ = ["add_scatter_to_masses", "scatter_geometry"]

_GEO_CACHE = {}
_GEO_DTYPE = np.dtype([("simplex", "i4"), ("nearest", "i4"), ("v", "i4", 3), ("pad", "i4"),
                       ("b", "f8", 3), ("e", "f8", 6), ("g", "f8", 3)])


def _roll_rows(arr, roll_num):
    """utils.roll_rows, utils.py:382-413"""
    roll = np.asarray(roll_num)
    nrows, ncols = arr.shape
    arr_roll = arr[:, [*range(ncols), *range(ncols-1)]].copy()
    strd_0, strd_1 = arr_roll.strides
    result = np.lib.stride_tricks.as_strided(arr_roll, (nrows, ncols, ncols), (strd_0, strd_1, strd_1))
    return result[np.arange(nrows), (ncols - roll) % ncols]


def _get_scatter_weights(uniform_cents, dist):
    """utils.get_scatter_weights, utils.py:416-452"""
    num = uniform_cents.size
    dx = np.diff(uniform_cents)
    if not np.allclose(dx, dx[0]):
        raise ValueError("`get_scatter_weights` only works if `uniform_cents` are uniformly spaced!")
    dx = dx[0]
    dx = dx/2.0 + np.arange(num) * dx
    dx = np.concatenate([-dx[::-1], dx])
    return np.diff(dist.cdf(dx))


def _get_rolled_weights(log_cents, dist):
    """utils._get_rolled_weights, utils.py:464-488"""
    num = log_cents.size
    weights = _get_scatter_weights(log_cents, dist)
    weights = weights[np.newaxis, :] * np.ones((num, weights.size))
    roll = 1 - num + np.arange(num)
    weights = _roll_rows(weights, roll)
    return weights[:, :num]


def _dependency_levels(indptr, indices, npts):
    """Level of every vertex in the dependency graph of an index-ordered Gauss-Seidel sweep: a vertex can be
    updated once all its lower-numbered neighbours have been.  Returns (order, level_ptr)."""
    level = np.zeros(npts, dtype=np.int64)
    for ip in range(npts):
        nb = indices[indptr[ip]:indptr[ip+1]]
        lo = nb[nb < ip]
        level[ip] = 0 if lo.size == 0 else level[lo].max() + 1
    order = np.argsort(level, kind="stable").astype(np.int32)      # within a level: ascending vertex index
    nlev = int(level.max()) + 1
    level_ptr = np.zeros(nlev + 1, dtype=np.int32)
    np.cumsum(np.bincount(level, minlength=nlev), out=level_ptr[1:])
    return order, level_ptr


GS_LANES = 8          # neighbours of one vertex handled per step
GS_SLOTS = 32         # vertices of a level handled per step
#: one step of the Gauss-Seidel step program, byte for byte the `GsRec` of csrc/holo_scatter.cu: what consumer thread
#: ``t = 4 slot + sub`` needs lies at ``16 t`` of each plane (conflict-free 16-byte loads).  The thread owns edges
#: ``sub`` (a) and ``sub + 4`` (b) of the vertex in ``slot``: planes ``e[0..3]`` = a's (2 ex, 2 ey), a's (ex, ey)/L^3,
#: b's (2 ex, 2 ey), b's (ex, ey)/L^3; ``ids[t]`` = (a's neighbour, b's neighbour, the vertex or -1, step flags);
#: ``qinv[0..1][slot]`` = the two rows of MINUS the inverse normal matrix.
GS_THREADS = GS_SLOTS * (GS_LANES // 2)
_STEP_DTYPE = np.dtype([("e", "f8", (4, GS_THREADS, 2)), ("qinv", "f8", (2, GS_SLOTS, 2)), ("ids", "i4", (GS_THREADS, 4))])


def _step_program(indptr, indices, edge4, qinv, order, level_ptr):
    """Flatten the level schedule of the Gauss-Seidel sweep into fixed-size step records (geometry only).

    A step handles up to ``GS_SLOTS`` vertices of ONE level, ``GS_LANES`` neighbours each in scipy's neighbour order
    (summed as the round-1 kernel did: neighbour ``l`` with ``l + 4``, then an xor-butterfly); vertices with more than
    8 neighbours (one hull corner of the named grid has 82) span several consecutive steps ("rounds"), the first
    flagged (bit 0) to start the sums, the last (bit 1) to apply the update.  The kernel streams these records with
    TMA bulk copies.  An absent edge points at the vertex itself with zero coefficients (it adds exactly nothing,
    without a branch).  Two exact rescalings shorten the kernel's dependent fp64 chain: the record holds ``2 ex, 2 ey``
    (the edge term is ``6 (f1 - f2) - 2 (-ex y0 - ey y1)``) and the NEGATED inverse normal matrix (the update stores
    ``-(Q^-1 s)``).
    """
    deg = np.diff(indptr)
    half = GS_LANES // 2
    steps = []
    for lv in range(level_ptr.size - 1):
        verts = order[level_ptr[lv]:level_ptr[lv + 1]]
        for c0 in range(0, verts.size, GS_SLOTS):
            chunk = verts[c0:c0 + GS_SLOTS]
            rounds = max(1, -(-int(deg[chunk].max()) // GS_LANES))
            for rr in range(rounds):
                rec = np.zeros((), dtype=_STEP_DTYPE)
                flags = (1 if rr == 0 else 0) | (2 if rr == rounds - 1 else 0)
                ids = np.zeros((GS_SLOTS, half, 4), dtype=np.int32)              # (empty slots: vertex 0, e = 0)
                ids[:, :, 2] = -1
                ids[:, :, 3] = flags
                ids[:chunk.size, :, 2] = chunk[:, None]
                kk = rr * GS_LANES + np.arange(GS_LANES)[None, :]                 # (1, lanes) neighbour number
                has = kk < deg[chunk][:, None]                                    # (slots, lanes)
                jp = np.where(has, indptr[chunk][:, None] + kk, 0)
                nb = np.where(has, indices[jp], chunk[:, None])                   # absent edge: the vertex itself, e = 0
                ee = np.where(has[..., None], edge4[jp], 0.0)                     # (slots, lanes, 4)
                ee[..., :2] *= 2.0                                                # the edge term uses 2 (e . y): exact
                ids[:chunk.size, :, 0] = nb[:, :half]
                ids[:chunk.size, :, 1] = nb[:, half:]
                epl = np.zeros((4, GS_SLOTS, half, 2))
                epl[0, :chunk.size] = ee[:, :half, 0:2]
                epl[1, :chunk.size] = ee[:, :half, 2:4]
                epl[2, :chunk.size] = ee[:, half:, 0:2]
                epl[3, :chunk.size] = ee[:, half:, 2:4]
                rec["e"] = epl.reshape(4, GS_THREADS, 2)
                rec["ids"] = ids.reshape(GS_THREADS, 4)
                qq = -qinv[chunk]                                                 # the update is y = -(Q^-1 s)
                rec["qinv"][0, :chunk.size] = qq[:, 0:2]
                rec["qinv"][1, :chunk.size] = qq[:, 2:4]
                steps.append(rec)
    return np.array(steps, dtype=_STEP_DTYPE)


def step_edges(rec):
    """A step record back in per-vertex form: ``(vip (slots,), flags, nb (slots, 8), e (slots, 8, 4), qinv (slots, 4))``
    (tests and the numpy emulation of the kernel)."""
    half = GS_LANES // 2
    ids = rec["ids"].reshape(GS_SLOTS, half, 4)
    ee = rec["e"].reshape(4, GS_SLOTS, half, 2)
    nb = np.concatenate([ids[:, :, 0], ids[:, :, 1]], axis=1)
    e8 = np.concatenate([np.concatenate([ee[0], ee[1]], axis=-1), np.concatenate([ee[2], ee[3]], axis=-1)], axis=1)
    qq = np.concatenate([rec["qinv"][0], rec["qinv"][1]], axis=-1)
    return ids[:, 0, 2].copy(), int(ids[0, 0, 3]), nb, e8, qq


def scatter_geometry(mtot, mrat, refine=4):
    """Host-side, data-independent set-up for one ``(mtot, mrat)`` grid (cached by the caller)."""
    import scipy.spatial
    mtot = np.asarray(mtot, dtype=np.float64)
    mrat = np.asarray(mrat, dtype=np.float64)
    # primary / secondary masses of the grid points and the regular grid, sam.py:1339-1354
    m1, m2 = utils.m1m2_from_mtmr(mtot[:, np.newaxis], mrat[np.newaxis, :])
    grid_size = m1.shape[0] * refine
    mextr = utils.minmax([0.9*mtot[0]*mrat[0]/(1.0 + mrat[0]), mtot[-1]*(1.0 + mrat[0])/mrat[0]])
    mgrid_log10 = np.log10(np.logspace(*np.log10(mextr), grid_size))
    pts = np.stack([np.log10(m1.flatten()), np.log10(m2.flatten())], axis=1)
    npts = pts.shape[0]

    # the triangulation CloughTocher2DInterpolator builds (scipy.spatial.Delaunay of the unscaled points)
    tri = scipy.spatial.Delaunay(pts)
    indptr, indices = tri.vertex_neighbor_vertices
    indptr = np.asarray(indptr, dtype=np.int32)
    indices = np.asarray(indices, dtype=np.int32)

    # gradient estimator (interpnd.pyx `_estimate_gradients_2d_global`): per directed edge ex, ey, L^3 and per
    # vertex the 2x2 normal matrix, accumulated in the neighbour order scipy uses
    src = np.repeat(np.arange(npts), np.diff(indptr))
    ex = pts[indices, 0] - pts[src, 0]
    ey = pts[indices, 1] - pts[src, 1]
    L = np.sqrt(ex**2 + ey**2)
    L3 = L*L*L
    edge = np.ascontiguousarray(np.stack([ex, ey, L3], axis=1))
    qmat = np.zeros((npts, 4))
    q0, q1, q3 = 4*ex*ex / L3, 4*ex*ey / L3, 4*ey*ey / L3
    for ip in range(npts):          # sequential sums, as the C loop
        a, b = indptr[ip], indptr[ip+1]
        s0 = s1 = s3 = 0.0
        for jp in range(a, b):
            s0 += q0[jp]
            s1 += q1[jp]
            s3 += q3[jp]
        qmat[ip] = (s0, s1, s3, s0*s3 - s1*s1)
    order, level_ptr = _dependency_levels(indptr, indices, npts)

    # regular-grid points: simplex, barycentric coordinates, nearest data point
    gx, gy = np.meshgrid(mgrid_log10, mgrid_log10, indexing='ij')
    xi = np.stack([gx.ravel(), gy.ravel()], axis=1)
    isimp = tri.find_simplex(xi)
    inside = isimp >= 0
    geo = np.zeros(xi.shape[0], dtype=_GEO_DTYPE)
    geo["simplex"] = isimp
    geo["nearest"] = scipy.spatial.cKDTree(pts).query(xi)[1]
    sidx = isimp[inside]
    T = tri.transform[sidx]
    c01 = np.einsum('tij,tj->ti', T[:, :2, :], xi[inside] - T[:, 2, :])
    geo["b"][inside] = np.concatenate([c01, 1.0 - c01.sum(axis=1, keepdims=True)], axis=1)
    S = tri.simplices
    P = tri.points
    geo["v"][inside] = S[sidx]
    e12 = P[S[:, 1]] - P[S[:, 0]]
    e23 = P[S[:, 2]] - P[S[:, 1]]
    e31 = P[S[:, 0]] - P[S[:, 2]]
    geo["e"][inside] = np.concatenate([e12, e23, e31], axis=1)[sidx]
    # affine-invariant edge parameters from the neighbours' centroids (interpnd.pyx `_clough_tocher_2d_single`)
    gpar = np.full((S.shape[0], 3), -0.5)
    cent = (P[S[:, 0]] + P[S[:, 1]] + P[S[:, 2]]) / 3
    for kk in range(3):
        itri = tri.neighbors[:, kk]
        ok = itri != -1
        dd = cent[itri[ok]] - tri.transform[ok, 2, :]
        cc = np.einsum('tij,tj->ti', tri.transform[ok, :2, :], dd)
        cc = np.concatenate([cc, 1.0 - cc.sum(axis=1, keepdims=True)], axis=1)
        i1, i2 = [(2, 1), (0, 2), (1, 0)][kk]
        gpar[ok, kk] = (2*cc[:, i1] + cc[:, i2] - 1) / (2 - 3*cc[:, i1] - 3*cc[:, i2])
    geo["g"][inside] = gpar[sidx]

    # RegularGridInterpolator(method='linear') back to the data points: cell index and normalised distance
    def find(xx):
        ii = np.searchsorted(mgrid_log10, xx) - 1
        ii = np.clip(ii, 0, grid_size - 2)
        return ii.astype(np.int32), (xx - mgrid_log10[ii]) / (mgrid_log10[ii+1] - mgrid_log10[ii])
    if (pts.min() < mgrid_log10[0]) or (pts.max() > mgrid_log10[-1]):
        raise ValueError("One of the requested xi is out of bounds")     # RegularGridInterpolator(bounds_error=True)
    i0, y0 = find(pts[:, 0])
    i1, y1 = find(pts[:, 1])
    # what the kernel reads: per-edge quotients and the inverse normal matrices (geometry only)
    edge4 = np.ascontiguousarray(np.stack([ex, ey, ex / L3, ey / L3], axis=1))
    det = qmat[:, 3]
    qinv = np.ascontiguousarray(np.stack([qmat[:, 2] / det, -qmat[:, 1] / det, -qmat[:, 1] / det, qmat[:, 0] / det], axis=1))
    program = _step_program(indptr, indices, edge4, qinv, order, level_ptr)
    return dict(npts=npts, G=grid_size, mgrid_log10=mgrid_log10, points=pts, indptr=indptr, indices=indices, edge=edge,
                qmat=qmat, edge4=edge4, qinv=qinv, order=order, level_ptr=level_ptr, geo=geo, i0=i0, i1=i1, y0=y0, y1=y1, tri=tri,
                program=program)


#: geometry arrays kept in the on-disk cache (everything `_device_geometry` uploads)
_DISK_KEYS = ("mgrid_log10", "geo", "program", "i0", "i1", "y0", "y1")


def _cache_dir():
    import os
    from pathlib import Path
    root = os.environ.get("HOLO_B200_CACHE")
    return Path(root) if root else Path.home() / ".cache" / "holodeck_b200"


def _cached_geometry(mtot, mrat, refine):
    """`scatter_geometry` through an on-disk cache keyed by the grid.  The set-up is data-independent but costs 1.7 s
    of host time at the named grid (scipy's point location of the 132,496 regular-grid points, 88 % of them outside
    the hull) -- as much as a hundred library samples -- and every rank of every run of a library needs the same one.
    ``HOLO_B200_CACHE`` moves the directory; ``HOLO_B200_CACHE=off`` disables the cache."""
    import hashlib
    import os
    if os.environ.get("HOLO_B200_CACHE", "").lower() in ("off", "0", "none"):
        return scatter_geometry(mtot, mrat, refine)
    hh = hashlib.sha1()
    for part in (np.ascontiguousarray(mtot, dtype=np.float64).tobytes(), np.ascontiguousarray(mrat, dtype=np.float64).tobytes(),
                 str(int(refine)).encode(), str(_STEP_DTYPE).encode(), str(_GEO_DTYPE).encode(), b"v6"):
        hh.update(part)
    fname = _cache_dir() / f"scatter_geometry_{hh.hexdigest()[:20]}.npz"
    try:
        with np.load(fname) as dd:
            gg = {kk: dd[kk] for kk in _DISK_KEYS}
        gg["geo"] = gg["geo"].view(_GEO_DTYPE).reshape(-1)
        gg["program"] = gg["program"].view(_STEP_DTYPE).reshape(-1)
        gg["npts"] = int(np.size(mtot) * np.size(mrat))
        gg["G"] = int(gg["mgrid_log10"].size)
        return gg
    except (OSError, KeyError, ValueError):
        pass
    gg = scatter_geometry(mtot, mrat, refine)
    try:
        fname.parent.mkdir(parents=True, exist_ok=True)
        tmp = fname.with_suffix(f".{os.getpid()}.tmp.npz")
        np.savez(tmp, **{kk: (gg[kk].view(np.uint8) if kk in ("geo", "program") else gg[kk]) for kk in _DISK_KEYS})
        os.replace(tmp, fname)             # atomic: concurrent ranks write the same bytes
    except OSError:
        pass
    return gg


def _gemm_blocks(i0, i1, G, nblk=4):
    """Row / column blocks of the scattered regular grid that the back-interpolation reads.

    `RegularGridInterpolator` touches the four corners ``(i0 + {0,1}, i1 + {0,1})`` of each data point only, and in
    (log m1, log m2) the data points fill the band ``1e-3 m1 <= m2 <= m1``: about a tenth of the ``G x G`` grid (10.6 % at
    the named grid).  The scatter product is therefore evaluated per block of secondary-mass columns, for the contiguous
    range of primary-mass rows that block needs.  Measured on B200 at the named grid (device time): the full product
    0.34 ms; 4 blocks (24 % of the flops) 0.118 ms, 8 blocks (18 %) 0.142 ms, 16 blocks (15 %) 0.183 ms -- the small
    products are launch- and tile-bound, so few blocks win; the values are bit-identical to the full product's."""
    need = np.zeros((G, G), dtype=bool)
    for da in (0, 1):
        for db in (0, 1):
            need[np.minimum(i0 + da, G - 1), np.minimum(i1 + db, G - 1)] = True
    cols = np.flatnonzero(need.any(axis=0))
    edges = np.unique(np.linspace(cols.min(), cols.max() + 1, nblk + 1).astype(int))
    blocks = []
    for b0, b1 in zip(edges[:-1], edges[1:]):
        rows = np.flatnonzero(need[:, b0:b1].any(axis=1))
        if rows.size:
            blocks.append((int(rows.min()), int(rows.max()) + 1, int(b0), int(b1)))
    return blocks


def _device_geometry(mtot, mrat, refine):
    import torch
    key = (np.asarray(mtot).tobytes(), np.asarray(mrat).tobytes(), int(refine), torch.cuda.current_device())
    hit = _GEO_CACHE.get(key)
    if hit is None:
        gg = _cached_geometry(mtot, mrat, refine)
        lib = _lib.load()
        assert lib.holo_scatter_geo_bytes() == _GEO_DTYPE.itemsize
        assert lib.holo_scatter_step_bytes() == _STEP_DTYPE.itemsize
        dev = dict(npts=gg["npts"], G=gg["G"], mgrid_log10=gg["mgrid_log10"], nsteps=int(gg["program"].size))
        for name in ("i0", "i1"):
            dev[name] = _lib.to_dev(gg[name], dtype=torch.int32)
        for name in ("y0", "y1"):
            dev[name] = _lib.to_dev(gg[name])
        dev["gemm_blocks"] = _gemm_blocks(np.asarray(gg["i0"]), np.asarray(gg["i1"]), int(gg["G"]))
        dev["geo"] = torch.from_numpy(gg["geo"].view(np.uint8).copy()).to(_lib.device())
        dev["program"] = torch.from_numpy(gg["program"].view(np.uint8).copy()).to(_lib.device())
        if len(_GEO_CACHE) > 8:
            _GEO_CACHE.clear()
        _GEO_CACHE[key] = hit = dev
    return hit


def _bad_values_error(log=None):
    err = "After 0th order interpolation, bad values remain!"        # sam.py:1376-1380
    if log is not None:
        log.exception(err)
    return ValueError(err)


def add_scatter_to_masses(mtot, mrat, dens, scatter, refine=4, log=None, *, _defer_check=None):
    """Add the given scatter [dex] to masses m1 and m2 of a ``(M, Q, Z)`` density grid (``sam.py:1291-1394``).

    ``dens`` may be a numpy array (numpy is returned, as the reference) or a CUDA tensor (a CUDA tensor is returned).
    """
    import scipy.stats
    import torch
    lib = _lib.require_gpu()
    on_dev = _lib.is_device_array(dens)
    assert dens.ndim == 3
    assert tuple(dens.shape[:2]) == (np.size(mtot), np.size(mrat))
    M, Q, Z = (int(ss) for ss in dens.shape)
    gg = _device_geometry(mtot, mrat, refine)
    npts, G = gg["npts"], gg["G"]
    data = _lib.to_dev(dens).reshape(npts, Z)
    # The two calls `_scatter_with_weights(.., axis=0)` and `(.., axis=1)` of the reference (sam.py:1383-1384) both
    # contract the PRIMARY-mass axis: `np.einsum("j...,jk...")` is in implicit mode, so its output is ordered
    # (..., k) -- the first call returns the scattered array transposed, the `moveaxis(.., 1, 0)` of the second call
    # brings the primary axis back to the front, and it is scattered again (utils.py:455-461).  The net operator
    # is  F = (W W)^T A  along axis 0, nothing along axis 1; reproduced as one DGEMM with the cached product.
    wkey = ("weights", float(scatter))
    if wkey not in gg:
        dist = scipy.stats.norm(loc=0.0, scale=scatter)
        ww = _lib.to_dev(_get_rolled_weights(gg["mgrid_log10"], dist))
        # (W W)^T on the device: a library samples a new scatter value for every model, and under torchrun
        # (OMP_NUM_THREADS=1) this 364^3 product was 10 ms of single-threaded host BLAS per sample
        stale = [kk for kk in gg if isinstance(kk, tuple) and kk[0] == "weights"]
        for kk in stale[:-3]:
            del gg[kk]                   # keep the few most recent ones only
        gg[wkey] = torch.matmul(ww, ww).T.contiguous()
    w2t = gg[wkey]

    grad = _lib.empty((npts, 2, Z))
    niter = torch.empty(Z, dtype=torch.int32, device=data.device)
    rc = lib.holo_scatter_gradients(npts, Z, _lib.ptr(gg["program"]), gg["nsteps"], _lib.ptr(data), 400, 1e-6,
                                    _lib.ptr(grad), _lib.ptr(niter), _lib.stream())
    _lib.check(rc, "add_scatter_to_masses (gradients)")
    grid = _lib.empty((G, G, Z))
    flags = torch.zeros(1, dtype=torch.int32, device=data.device)
    rc = lib.holo_scatter_ct_eval(G * G, Z, _lib.ptr(gg["geo"]), _lib.ptr(data), _lib.ptr(grad), _lib.ptr(grid),
                                  _lib.ptr(flags), _lib.stream())
    _lib.check(rc, "add_scatter_to_masses (interpolation)")
    # cuBLAS DGEMMs, new[k, b, z] = sum_j (WW)[j, k] A[j, b, z], for the blocks the back-interpolation reads only
    # (the rest of `scat` stays unwritten and is never read)
    a2 = grid.reshape(G, G * Z)
    scat = _lib.empty((G, G * Z))
    for k0, k1, b0, b1 in gg["gemm_blocks"]:
        torch.mm(w2t[k0:k1], a2[:, b0 * Z:b1 * Z], out=scat[k0:k1, b0 * Z:b1 * Z])
    grid = scat
    out = _lib.empty((npts, Z))
    rc = lib.holo_scatter_bilinear(npts, G, Z, _lib.ptr(gg["i0"]), _lib.ptr(gg["i1"]), _lib.ptr(gg["y0"]), _lib.ptr(gg["y1"]),
                                   _lib.ptr(grid), _lib.ptr(out), _lib.stream())
    _lib.check(rc, "add_scatter_to_masses (back-interpolation)")
    if _defer_check is not None:
        _defer_check.append(flags)      # the caller reads the flag at its next synchronisation point (no stall here)
    elif int(flags.item()) != 0:
        raise _bad_values_error(log)
    out = out.reshape(M, Q, Z)
    return out if on_dev else _lib.to_host(out)
