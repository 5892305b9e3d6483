"""numpy restatement of the device scatter pipeline (holodeck_b200/csrc/holo_scatter.cu) -- TEST INFRASTRUCTURE ONLY.

It consumes the same host-prepared geometry as the kernels (``holodeck_b200.sams.scatter.scatter_geometry``) and
restates, operation for operation, scipy 1.18.1's ``interpnd.pyx`` routines the reference reaches through
``CloughTocher2DInterpolator`` (``holodeck/sams/sam.py:1362-1367``): ``_estimate_gradients_2d_global`` and
``_clough_tocher_2d_single``.  scipy ships no source for them in this image; the restatement is pinned against the
compiled scipy by ``tests/test_scatter.py`` (gradients and interpolated values agree to ~1e-15).
"""
import numpy as np


def gradients_sequential(geo, data, maxiter=400, tol=1e-6):
    """`_estimate_gradients_2d_global`: index-ordered Gauss-Seidel sweeps (pure-Python loops; small cases only)."""
    indptr, indices, edge, qmat = geo["indptr"], geo["indices"], geo["edge"], geo["qmat"]
    npts = geo["npts"]
    yy = np.zeros((npts, 2))
    for it in range(maxiter):
        err = 0.0
        for ip in range(npts):
            s0 = s1 = 0.0
            f1 = data[ip]
            for jp in range(indptr[ip], indptr[ip+1]):
                ip2 = indices[jp]
                ex, ey, L3 = edge[jp]
                df2 = -ex*yy[ip2, 0] - ey*yy[ip2, 1]
                num = 6*(f1 - data[ip2]) - 2*df2
                s0 += num*ex/L3
                s1 += num*ey/L3
            Q0, Q1, Q3, det = qmat[ip]
            r0 = (Q3*s0 - Q1*s1)/det
            r1 = (-Q1*s0 + Q0*s1)/det
            change = max(abs(yy[ip, 0] + r0), abs(yy[ip, 1] + r1))
            yy[ip] = (-r0, -r1)
            change /= max(1.0, max(abs(r0), abs(r1)))
            err = max(err, change)
        if err < tol:
            return yy, it + 1
    return yy, 0


def gradients_levels(geo, data, maxiter=400, tol=1e-6):
    """The same sweeps executed level by level of the dependency graph, all vertices of a level at once (what the
    kernel does): identical iterates because same-level vertices are never neighbours."""
    indptr, indices, edge, qmat = geo["indptr"], geo["indices"], geo["edge"], geo["qmat"]
    order, level_ptr = geo["order"], geo["level_ptr"]
    npts = geo["npts"]
    deg = np.diff(indptr)
    yy = np.zeros((npts, 2))
    for it in range(maxiter):
        err = 0.0
        for lv in range(level_ptr.size - 1):
            vv = order[level_ptr[lv]:level_ptr[lv+1]]
            s0 = np.zeros(vv.size)
            s1 = np.zeros(vv.size)
            for kk in range(deg[vv].max()):          # neighbour by neighbour: the sequential summation order
                act = deg[vv] > kk
                jp = indptr[vv[act]] + kk
                ip2 = indices[jp]
                ex, ey, L3 = edge[jp, 0], edge[jp, 1], edge[jp, 2]
                df2 = -ex*yy[ip2, 0] - ey*yy[ip2, 1]
                num = 6*(data[vv[act]] - data[ip2]) - 2*df2
                s0[act] += num*ex/L3
                s1[act] += num*ey/L3
            Q0, Q1, Q3, det = qmat[vv].T
            r0 = (Q3*s0 - Q1*s1)/det
            r1 = (-Q1*s0 + Q0*s1)/det
            change = np.maximum(np.abs(yy[vv, 0] + r0), np.abs(yy[vv, 1] + r1))
            yy[vv, 0] = -r0
            yy[vv, 1] = -r1
            change = change / np.maximum(1.0, np.maximum(np.abs(r0), np.abs(r1)))
            err = max(err, float(change.max()))
        if err < tol:
            return yy, it + 1
    return yy, 0


def clough_tocher(geo, data, grad):
    """`_clough_tocher_2d_single` at every regular-grid point inside the hull (NaN outside); (G*G,) values."""
    gp = geo["geo"]
    out = np.full(gp.shape[0], np.nan)
    ins = gp["simplex"] >= 0
    vv = gp["v"][ins]
    bb = gp["b"][ins]
    ee = gp["e"][ins]
    gg = gp["g"][ins]
    f1, f2, f3 = data[vv[:, 0]], data[vv[:, 1]], data[vv[:, 2]]
    d0, d1, d2 = grad[vv[:, 0]], grad[vv[:, 1]], grad[vv[:, 2]]
    e12x, e12y, e23x, e23y, e31x, e31y = ee.T
    df12 = +(d0[:, 0]*e12x + d0[:, 1]*e12y)
    df21 = -(d1[:, 0]*e12x + d1[:, 1]*e12y)
    df23 = +(d1[:, 0]*e23x + d1[:, 1]*e23y)
    df32 = -(d2[:, 0]*e23x + d2[:, 1]*e23y)
    df31 = +(d2[:, 0]*e31x + d2[:, 1]*e31y)
    df13 = -(d0[:, 0]*e31x + d0[:, 1]*e31y)
    c3000 = f1
    c2100 = (df12 + 3*c3000)/3
    c2010 = (df13 + 3*c3000)/3
    c0300 = f2
    c1200 = (df21 + 3*c0300)/3
    c0210 = (df23 + 3*c0300)/3
    c0030 = f3
    c1020 = (df31 + 3*c0030)/3
    c0120 = (df32 + 3*c0030)/3
    c2001 = (c2100 + c2010 + c3000)/3
    c0201 = (c1200 + c0300 + c0210)/3
    c0021 = (c1020 + c0120 + c0030)/3
    c0111 = (gg[:, 0]*(-c0300 + 3*c0210 - 3*c0120 + c0030) + (-c0300 + 2*c0210 - c0120 + c0021 + c0201))/2
    c1011 = (gg[:, 1]*(-c0030 + 3*c1020 - 3*c2010 + c3000) + (-c0030 + 2*c1020 - c2010 + c2001 + c0021))/2
    c1101 = (gg[:, 2]*(-c3000 + 3*c2100 - 3*c1200 + c0300) + (-c3000 + 2*c2100 - c1200 + c2001 + c0201))/2
    c1002 = (c1101 + c1011 + c2001)/3
    c0102 = (c1101 + c0111 + c0201)/3
    c0012 = (c1011 + c0111 + c0021)/3
    c0003 = (c1002 + c0102 + c0012)/3
    minval = bb.min(axis=1)
    b1, b2, b3, b4 = bb[:, 0] - minval, bb[:, 1] - minval, bb[:, 2] - minval, 3*minval
    out[ins] = (b1**3*c3000 + 3*b1**2*b2*c2100 + 3*b1**2*b3*c2010 + 3*b1**2*b4*c2001 + 3*b1*b2**2*c1200 +
                6*b1*b2*b4*c1101 + 3*b1*b3**2*c1020 + 6*b1*b3*b4*c1011 + 3*b1*b4**2*c1002 + b2**3*c0300 +
                3*b2**2*b3*c0210 + 3*b2**2*b4*c0201 + 3*b2*b3**2*c0120 + 6*b2*b3*b4*c0111 + 3*b2*b4**2*c0102 +
                b3**3*c0030 + 3*b3**2*b4*c0021 + 3*b3*b4**2*c0012 + b4**3*c0003)
    return out


def add_scatter_port(geo, weights, dens):
    """The whole device pipeline in numpy for a (M, Q, Z) density (weights: utils._get_rolled_weights, (G, G))."""
    G, npts = geo["G"], geo["npts"]
    Z = dens.shape[2]
    data = dens.reshape(npts, Z)
    out = np.zeros_like(data)
    for zz in range(Z):
        grad, _ = gradients_levels(geo, data[:, zz])
        ww = clough_tocher(geo, data[:, zz], grad)
        bad = np.isnan(ww) | (ww < 0.0)
        ww[bad] = data[geo["geo"]["nearest"][bad], zz]
        grid = ww.reshape(G, G)
        grid = np.einsum("j...,jk...", grid, weights)
        grid = np.moveaxis(np.einsum("j...,jk...", np.moveaxis(grid, 1, 0), weights), 0, 1)
        i0, i1, y0, y1 = geo["i0"], geo["i1"], geo["y0"], geo["y1"]
        out[:, zz] = (grid[i0, i1]*(1 - y0)*(1 - y1) + grid[i0, i1 + 1]*(1 - y0)*y1 +
                      grid[i0 + 1, i1]*y0*(1 - y1) + grid[i0 + 1, i1 + 1]*y0*y1)
    return out.reshape(dens.shape)
