"""Plain numpy / pure-Python restatement of the reference's Cython loops -- TEST INFRASTRUCTURE ONLY.

Second, independent leg of the oracle next to the compiled reference (``oracle/_ref``): every
function restates one ``cdef`` loop of ``holodeck/cyutils.pyx`` / ``holodeck/sams/sam_cyutils.pyx`` in
the reference's own loop order (so that ``np.cumsum`` -- a strictly sequential sum -- reproduces its
floating-point accumulation bit for bit), taking the Poisson draws as an input (``counts``) instead
of drawing them.  ``tests/test_port.py`` checks each against outputs of the compiled reference.

Only ``tests/`` may import this module.
"""
import numpy as np

# ---- the Cython constant set (sam_cyutils.pyx:30-42)
MY_NWTG = 6.6742999e-08
MY_SPLC = 29979245800.0
MY_MPC = 3.08567758e+24
MY_SCHW = 1.4852320538237328e-28
GW_DADT_SEP_CONST = - 64.0 * pow(MY_NWTG, 3) / 5.0 / pow(MY_SPLC, 5)
KEPLER_CONST_FREQ = (1.0 / (2.0*np.pi)) * np.sqrt(MY_NWTG)
KEPLER_CONST_SEPA = pow(MY_NWTG, 1.0/3.0) / pow(2.0*np.pi, 2.0/3.0)
FOUR_PI_SPLC_OVER_MPC = 4 * np.pi * MY_SPLC / MY_MPC


def integrate_differential_number_3dx1d(edges, dnum):
    """sam_cyutils.pyx:115-216: 8-corner sum in (ii, jj, kk) order, then * dm dq dz * dlnf / 8."""
    l10m = np.log10(edges[0])
    dlnf = np.diff(np.log(edges[3]))
    temp = np.zeros(tuple(ss - 1 for ss in dnum.shape[:3]) + (dnum.shape[3],))
    for ii in range(2):
        for jj in range(2):
            for kk in range(2):
                temp = temp + dnum[ii:dnum.shape[0]-1+ii, jj:dnum.shape[1]-1+jj, kk:dnum.shape[2]-1+kk, :]
    dm = np.diff(l10m)[:, None, None]
    dmdq = dm * np.diff(edges[1])[None, :, None]
    dmdqdz = dmdq * np.diff(edges[2])[None, None, :]
    return temp * dmdqdz[..., None] * dlnf / 8.0


def _seqsum(arr, axis=-1):
    """Strictly sequential (left-to-right) floating-point sum, like a C accumulation loop."""
    if arr.shape[axis] == 0:
        return np.zeros(np.delete(arr.shape, axis))
    return np.take(np.cumsum(arr, axis=axis), -1, axis=axis)


def sam_poisson_gwb(number, hc2, counts):
    """cyutils.pyx:862-897 with supplied draws: gwb[f, r] += num * hc2 over cells in (m,q,z) order."""
    F = number.shape[-1]
    h = hc2.reshape(-1, F)                          # (ncell, F)
    contrib = counts * h.T[None, :, :]              # (R, F, ncell)
    return _seqsum(contrib, axis=2).T               # (F, R)


def _loudest_core(number, h2fdf, nloud, order, counts, skip_zero_h):
    """Shared walk of cyutils.pyx:1318-1342 / 1475-1508 / 1701-1752 for all (r, f).

    Returns per (r, f): arrays in RANK order of `take` (slots taken per cell), `rem` (binaries left for
    the background; 0 for skipped cells), `cur`, and the start slot of each cell."""
    F = number.shape[-1]
    h = h2fdf.reshape(-1, F)[order, :]              # (ncell, F) rank order
    n = counts[:, :, order]                         # (R, F, ncell)
    cur = np.broadcast_to(h.T[None, :, :], n.shape)
    elig = n >= 1.0                                 # `if (num < 1): continue`
    if skip_zero_h:
        elig = elig & (cur != 0.0)                  # cyutils.pyx:1727
    ne = np.where(elig, n, 0.0)
    want = np.ceil(ne)                              # slots a cell would take: `while ... (num > 0): num -= 1`
    cum = np.cumsum(want, axis=2)
    start = cum - want
    take = np.clip(nloud - start, 0.0, want)
    rem = ne - take
    return cur, elig, take, rem, start


def _fill_slots(values_rank, take, start, nloud, dtype=float):
    """Scatter per-cell `values` into the L slots they occupy.  values_rank/take/start: (R, F, ncell)."""
    R, F, ncell = take.shape
    out = np.zeros((F, R, nloud), dtype=dtype)
    rr, ff, cc = np.nonzero(take > 0)
    for r_, f_, c_ in zip(rr, ff, cc):
        s0 = int(start[r_, f_, c_])
        out[f_, r_, s0:s0 + int(take[r_, f_, c_])] = values_rank[r_, f_, c_]
    return out


def loudest_hc_from_sorted(number, h2fdf, nloud, order, counts):
    """cyutils.pyx:1266-1344"""
    cur, elig, take, rem, start = _loudest_core(number, h2fdf, nloud, order, counts, False)
    hc2ss = _fill_slots(cur, take, start, nloud)
    # `sum += num * cur` only for non-skipped cells, in rank order
    hc2bg = _seqsum(np.where(elig, rem * cur, 0.0)[:, :, :], axis=2).T
    # skipped cells add nothing at all (not even +0.0 matters)
    return hc2ss, hc2bg


def _unravel(order, shape):
    zz = order % shape[2]
    mq = order // shape[2]
    return mq // shape[1], mq % shape[1], zz


def loudest_hc_and_par_from_sorted(number, h2fdf, nloud, mt, mr, rz, order, counts):
    """cyutils.pyx:1409-1538"""
    shape = number.shape[:3]
    mm, qq, zz = _unravel(order.astype(np.int64), shape)
    cur, elig, take, rem, start = _loudest_core(number, h2fdf, nloud, order, counts, False)
    out = dict(hc2ss=_fill_slots(cur, take, start, nloud))
    bcast = lambda vv: np.broadcast_to(vv[None, None, :], take.shape)   # noqa: E731
    out["ssidx"] = np.stack([_fill_slots(bcast(ii), take, start, nloud, dtype=np.int64) for ii in (mm, qq, zz)])
    nc = np.where(elig, rem * cur, 0.0)
    sum_bg = _seqsum(nc, axis=2)
    out["hc2bg"] = sum_bg.T
    out["bgpar"] = np.stack([(_seqsum(nc * bcast(vv), axis=2) / sum_bg).T for vv in (mt[mm], mr[qq], rz[zz])])
    # loudest-source sums: one `+= cur` per slot (cyutils.pyx:1497-1500), i.e. `take` repeats per cell
    R, F, ncell = take.shape
    sum_ls = np.zeros((F, R))
    par_ls = np.zeros((3, F, R))
    pars = (mt[mm], mr[qq], rz[zz])
    rr_, ff_, cc_ = np.nonzero(take > 0)
    for r_, f_, c_ in zip(rr_, ff_, cc_):
        for _ in range(int(take[r_, f_, c_])):
            sum_ls[f_, r_] += cur[r_, f_, c_]
            for kk in range(3):
                par_ls[kk, f_, r_] += cur[r_, f_, c_] * pars[kk][c_]
    with np.errstate(invalid="ignore", divide="ignore"):
        out["lspar"] = par_ls / sum_ls[None]
    return out


def loudest_hc_and_par_from_sorted_redz(number, h2fdf, nloud, mt, mr, rz, redz_final, dcom_final, sepa, angs, order, counts):
    """cyutils.pyx:1615-1767"""
    shape = number.shape[:3]
    F = number.shape[-1]
    mm, qq, zz = _unravel(order.astype(np.int64), shape)
    cur, elig, take, rem, start = _loudest_core(number, h2fdf, nloud, order, counts, True)
    bcast = lambda vv: np.broadcast_to(vv[None, None, :], take.shape)   # noqa: E731
    per_f = lambda arr: np.broadcast_to(arr.reshape(-1, F)[order, :].T[None], take.shape)   # noqa: E731
    out = dict(hc2ss=_fill_slots(cur, take, start, nloud))
    out["sspar"] = np.stack([_fill_slots(vv, take, start, nloud) for vv in
                             (bcast(mt[mm]), bcast(mr[qq]), bcast(rz[zz]), per_f(redz_final))])
    with np.errstate(invalid="ignore"):
        nc = np.where(elig, rem * cur, 0.0)
        sum_bg = _seqsum(nc, axis=2)
        out["hc2bg"] = sum_bg.T
        pars = [bcast(mt[mm]), bcast(mr[qq]), bcast(rz[zz]), per_f(redz_final), per_f(dcom_final), per_f(sepa), per_f(angs)]
        with np.errstate(divide="ignore"):
            out["bgpar"] = np.stack([(_seqsum(np.where(elig, nc * pp, 0.0), axis=2) / sum_bg).T for pp in pars])
    return out


def ss_bg_hc_and_par(number, h2fdf, mt, mr, rz, counts):
    """cyutils.pyx:935-1014 and 1017-1178 (arg-max by value over occupied cells, natural order)."""
    shape = number.shape[:3]
    F = number.shape[-1]
    h = h2fdf.reshape(-1, F)                            # (ncell, F)
    cur = np.broadcast_to(h.T[None], counts.shape)      # (R, F, ncell)
    cand = np.where(counts > 0, cur, 0.0)
    imax = np.argmax(cand, axis=2)                      # first occurrence of the maximum == strict `>` walk
    vmax = np.take_along_axis(cand, imax[..., None], axis=2)[..., 0]
    found = vmax > 0
    nc = counts * cur
    tot = _seqsum(nc, axis=2)
    out = dict(hc2ss=vmax.T, hc2bg=(tot - vmax).T)
    cells = np.where(found, imax, -1)
    zz = np.where(found, imax % shape[2], -1)
    mq = imax // shape[2]
    qq = np.where(found, mq % shape[1], -1)
    mm = np.where(found, mq // shape[1], -1)
    out["ssidx"] = np.stack([mm.T, qq.T, zz.T]).astype(np.int64)
    midx = np.arange(h.shape[0])
    m_all, q_all, z_all = _unravel(midx, shape)
    pars = (mt[m_all], mr[q_all], rz[z_all])
    bg, ss = [], []
    for kk, pp in enumerate(pars):
        avg = _seqsum(nc * pp[None, None, :], axis=2)
        pm = pp[np.clip(cells, 0, None)]
        with np.errstate(invalid="ignore", divide="ignore"):
            bg.append(((avg - vmax * pm) / (tot - vmax)).T)
        ss.append(pm.T)
    out["bgpar"] = np.stack(bg)
    out["sspar"] = np.stack(ss)
    return out


# ==================================================================================================
# dynamic binary number: step-major pure-Python loops for a handful of (M, q) rows
# ==================================================================================================

def _hard_gw(mt, mr, sepa):
    return GW_DADT_SEP_CONST * pow(mt, 3) * mr / pow(sepa, 3) / pow(1 + mr, 2)


def _hard_2pwl_gw(mt, mr, sepa, norm, rchar, gi, go):
    xx = sepa / rchar
    dadt = - norm * pow(1.0 + xx, -go + gi) / pow(xx, gi - 1)
    return dadt + _hard_gw(mt, mr, sepa)


def _kep_freq(mt, sepa):
    return KEPLER_CONST_FREQ * np.sqrt(mt) / pow(sepa, 1.5)


def _kep_sepa(mt, freq):
    return KEPLER_CONST_SEPA * pow(mt, 1.0/3.0) / pow(freq, 2.0/3.0)


def _ww_increasing(start, size, val, edges):
    """sam_cyutils.pyx:66-84"""
    index = start
    while (index < size - 2) and (edges[index+1] < val):
        index += 1
    while (index > 0) and (edges[index] > val):
        index -= 1
    return index


def _ww_decreasing(start, size, val, edges):
    """sam_cyutils.pyx:89-107"""
    index = start
    while (index < size - 1) and (edges[index+1] > val):
        index += 1
    while (index > 0) and (edges[index-1] < val):
        index -= 1
    return index


def _interp(idx, xnew, xold, yold):
    """cyutils.pyx:326-358"""
    xl, xr, yl, yr = xold[idx], xold[idx+1], yold[idx], yold[idx+1]
    return yl + (yr - yl) * (xnew - xl) / (xr - xl)


def dynamic_binary_number_rows(gg, rows):
    """`_dynamic_binary_number_at_fobs_2pwl` (sam_cyutils.pyx:510-781) or `_gw` (:788-899) for the given
    (ii, jj) rows of a golden fixture; returns lists of (Z, F) arrays (redz_final, diff_num)."""
    fobs = gg["fobs_cents"] / 2.0
    redz, mtot, mrat = gg["redz"], gg["mtot"], gg["mrat"]
    gz, gdc, gage = gg["grid_z"], gg["grid_dcom"], gg["grid_age"]
    n_interp = gz.size
    Z, F = redz.size, fobs.size
    out_rz, out_dn = [], []
    if str(gg["hard"]) == "gw":
        for (ii, jj) in rows:
            rzf = -np.ones((Z, F))
            dnf = np.zeros((Z, F))
            mt, mr = mtot[ii], mrat[jj]
            fisco = _kep_freq(mt, 3.0 * MY_SCHW * mt)
            idx = 0
            for kk in range(Z - 1, -1, -1):
                rzp = gg["redz_prime"][ii, jj, kk] if "redz_prime" in gg else redz[kk]
                if rzp <= 0.0:
                    continue
                for ff in range(F):
                    rzf[kk, ff] = rzp
                    frst = fobs[ff] * (1.0 + rzp)
                    if frst > fisco:
                        rzf[kk, ff+1:] = rzp
                        break
                    idx = _ww_decreasing(idx, n_interp, rzp, gz)
                    dcom = _interp(idx, rzp, gz, gdc)
                    sepa = _kep_sepa(mt, frst)
                    tres = - (2.0/3.0) * sepa / _hard_gw(mt, mr, sepa)
                    dnf[kk, ff] = gg["dens"][ii, jj, kk] * tres * (FOUR_PI_SPLC_OVER_MPC * (1.0 + rzp) * pow(dcom / MY_MPC, 2))
            out_rz.append(rzf)
            out_dn.append(dnf)
        return out_rz, out_dn

    hp = gg["hard_params"]
    sepa_init, rchar, gi, go, nsteps = hp[1], hp[2], hp[3], hp[4], int(hp[5])
    age_universe = gage[n_interp - 1]
    redz_age = np.zeros(Z)
    idx = 0
    for kk in range(Z):
        rev = Z - 1 - kk
        while (gz[idx+1] > redz[rev]) and (idx < n_interp - 1):
            idx += 1
        redz_age[rev] = _interp(idx, redz[rev], gz, gage)
    for (ii, jj) in rows:
        rzf = -np.ones((Z, F))
        dnf = np.zeros((Z, F))
        mt, mr = mtot[ii], mrat[jj]
        norm = 10.0 ** gg["norm_log10"][ii, jj]
        dx = (np.log10(sepa_init) - np.log10(3.0 * MY_SCHW * mt)) / nsteps
        sepa_log10 = np.log10(sepa_init)
        sepa_left = pow(10.0, sepa_log10)
        dadt_left = _hard_2pwl_gw(mt, mr, sepa_left, norm, rchar, gi, go)
        frst_left = _kep_freq(mt, sepa_left)
        time_evo = 0.0
        il = 0
        for step in range(nsteps):
            sepa_log10 -= dx
            sepa_right = pow(10.0, sepa_log10)
            frst_right = _kep_freq(mt, sepa_right)
            dadt_right = _hard_2pwl_gw(mt, mr, sepa_right, norm, rchar, gi, go)
            dt = 2.0 * (sepa_right - sepa_left) / (dadt_left + dadt_right)
            time_evo += dt
            for kk in range(Z - 1, -1, -1):
                time_right = time_evo + gg["gmt_time"][ii, jj, kk] + redz_age[kk]
                time_left = time_right - dt
                if time_left > age_universe:
                    continue
                il = _ww_increasing(il, n_interp, time_left, gage)
                redz_left = _interp(il, time_left, gage, gz)
                if redz_left < 0.0:
                    continue
                ir = _ww_increasing(il, n_interp, time_right, gage)
                redz_right = _interp(ir, time_right, gage, gz)
                if redz_right < 0.0:
                    redz_right = 0.0
                fl = frst_left / (1.0 + redz_left)
                fr = frst_right / (1.0 + redz_right)
                for ff in range(F):
                    ft = fobs[ff]
                    if (ft < fl) or (fr < ft):
                        continue
                    new_time = time_left + (time_right - time_left) * (ft - fl) / (fr - fl)
                    if new_time > age_universe:
                        break
                    ni = _ww_increasing(il, n_interp, new_time, gage)
                    new_redz = _interp(ni, new_time, gage, gz)
                    dcom = _interp(ni, new_time, gage, gdc)
                    rzf[kk, ff] = new_redz
                    sepa = _kep_sepa(mt, ft * (1.0 + new_redz))
                    dadt = _hard_2pwl_gw(mt, mr, sepa, norm, rchar, gi, go)
                    tres = - (2.0/3.0) * sepa / dadt
                    dnf[kk, ff] = gg["dens"][ii, jj, kk] * tres * (FOUR_PI_SPLC_OVER_MPC * (1.0 + new_redz) * pow(dcom / MY_MPC, 2))
            dadt_left = dadt_right
            sepa_left = sepa_right
            frst_left = frst_right
        out_rz.append(rzf)
        out_dn.append(dnf)
    return out_rz, out_dn
