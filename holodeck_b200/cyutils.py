"""Drop-in for the reference's compiled module ``holodeck.cyutils`` (hot-path subset).

Same callables and positional signatures as ``holodeck/cyutils.pyx``; the loops run as sm_100a
kernels in ``libholo_b200.so`` (``include/holo_b200.h``).  Out of scope:
``gamma_of_rho_interp``, ``snr_ss``, ``Sh_rest`` (detection statistics) and ``interp_2d``.

Keyword-only additions (all optional, defaults reproduce the reference's behaviour):

``seed``     integer seed of the counter-based Philox generator.  ``None`` (default) draws a fresh
             seed from the OS, like the reference's unseeded ``PCG64()`` (cyutils.pyx:875,1315,...).
``counts``   "supplied-count mode": an ``(R, F, M*Q*Z)`` array of draws to use instead of random
             numbers (bit-exact parity testing against the seeded reference).
``r0``       global index of the first realization (realization sharding over GPUs; the random
             numbers of realization ``r0 + i`` do not depend on how realizations are partitioned).
``device``   return CUDA ``torch`` tensors instead of numpy arrays.
"""
import ctypes as C
import os

import numpy as np

from holodeck_b200 import _lib

__all__ = [
    "sam_poisson_gwb", "loudest_hc_from_sorted", "loudest_hc_and_par_from_sorted",
    "loudest_hc_and_par_from_sorted_redz", "ss_bg_hc", "ss_bg_hc_and_par",
    "sam_calc_gwb_single_eccen", "sam_calc_gwb_single_eccen_discrete",
]

_MAX_RETRY = 5

#: diagnostics: how often `holo_loudest` had to be re-run with a larger head margin / event bucket
STATS = {"loudest_calls": 0, "loudest_retries": 0}


#: the resolver sorts a bucket in shared memory: RES_WARPS * 2 * cap * 16 B <= 200 KB (holo_realize.cu) -> cap <= 1600
_MAX_BUCKET_CAP = 1536


def _retry_schedule(L, attempt):
    """(head_margin, bucket_cap) of retry number ``attempt`` >= 1 of `holo_loudest` after HOLO_ERR_OVERFLOW.

    The margin grows 4x per attempt from the library's default ``8 sqrt(L) + 24``; the bucket must hold the
    ``L + margin`` events the head is cut for, with the same slack as the library's own choice
    (``auto_cap``: next power of two >= 2 (L + margin) + 64).  Both stop growing where the bucket reaches the
    resolver's shared-memory limit: from there on a retry keeps the largest head that still fits.
    """
    margin = (8.0 * np.sqrt(L) + 24.0) * (4.0 ** attempt)
    margin = min(margin, max((_MAX_BUCKET_CAP - 64) / 2.0 - L, 8.0 * np.sqrt(L) + 24.0))
    cap = 64
    while cap < 2.0 * (L + margin) + 64.0:
        cap *= 2
    return float(margin), int(min(cap, _MAX_BUCKET_CAP))


def _seed(seed):
    if seed is None:
        return int.from_bytes(os.urandom(8), "little")
    return int(seed) & 0xFFFFFFFFFFFFFFFF


def _workspace(nbytes):
    import torch
    return torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=_lib.device())


def _out(tensor, device):
    return tensor if device else _lib.to_host(tensor)


def _counts(counts, R, F, ncell):
    if counts is None:
        return None
    cc = _lib.to_dev(counts)
    assert tuple(cc.shape) == (R, F, ncell), f"`counts` must be shaped (R, F, M*Q*Z) = {(R, F, ncell)}, got {tuple(cc.shape)}"
    return cc


def _order(msort, qsort, zsort, shape):
    """(msort, qsort, zsort) -> flat int32 cell indices, loudest first (single_sources.py:89-93)."""
    import torch
    _, Qb, Zb = shape
    if _lib.is_device_array(msort):
        flat = (msort.to(torch.int64) * Qb + qsort.to(torch.int64)) * Zb + zsort.to(torch.int64)
        return flat.to(torch.int32).contiguous()
    flat = (np.asarray(msort, dtype=np.int64) * Qb + np.asarray(qsort, dtype=np.int64)) * Zb + np.asarray(zsort, dtype=np.int64)
    return _lib.to_dev(flat.astype(np.int32), dtype=torch.int32)


# ==================================================================================================
# K3
# ==================================================================================================

def sam_poisson_gwb(dist, hc2, nreals, normal_threshold=1e10, *, seed=None, counts=None, r0=0, device=False):
    """GWB from `nreals` Poisson realizations of the binary-number grid (cyutils.pyx:854-897).

    ``dist``, ``hc2`` : (M, Q, Z, F) expected numbers and per-source hc^2;  returns ``gwb`` (F, R).
    Bins with ``dist > int(normal_threshold)`` use an un-floored normal draw (cyutils.pyx:886-891).
    """
    lib = _lib.require_gpu()
    number = _lib.to_dev(dist)
    h2 = _lib.to_dev(hc2)
    assert number.shape == h2.shape and number.dim() == 4
    F = number.shape[3]
    ncell = number.numel() // F
    R = int(nreals)
    cnt = _counts(counts, R, F, ncell)
    gwb = _lib.empty((F, R))
    ws = _workspace(lib.holo_realize_workspace_bytes(0, ncell, F, R))
    rc = lib.holo_sam_poisson_gwb(_lib.ptr(number), _lib.ptr(h2), ncell, F, R, int(r0), _seed(seed),
                                  float(int(normal_threshold)), _lib.ptr(cnt), _lib.ptr(gwb), _lib.ptr(ws),
                                  ws.numel(), _lib.stream())
    _lib.check(rc, "sam_poisson_gwb")
    return _out(gwb, device)


# ==================================================================================================
# K4
# ==================================================================================================

def _loudest(variant, number, h2fdf, nreals, nloudest, msort, qsort, zsort, normal_threshold,
             mt=None, mr=None, rz=None, redz_final=None, dcom_final=None, sepa=None, angs=None,
             seed=None, counts=None, r0=0, gwb_nreals=None, gwb_seed=None, gwb_r0=0, order=None, defer=None):
    import torch
    lib = _lib.require_gpu()
    number = _lib.to_dev(number)
    h2fdf = _lib.to_dev(h2fdf)
    assert number.shape == h2fdf.shape and number.dim() == 4
    Mb, Qb, Zb, F = [int(ss) for ss in number.shape]
    ncell = Mb * Qb * Zb
    R, L = int(nreals), int(nloudest)
    # `order` (keyword-only addition): the flat int32 rank order on the device, as `_order` would build it
    order = _order(msort, qsort, zsort, (Mb, Qb, Zb)) if order is None else order
    assert order.numel() == ncell, "msort/qsort/zsort must list every (M,Q,Z) bin exactly once"
    cnt = _counts(counts, R, F, ncell)

    keep = [number, h2fdf, order, cnt]
    args = _lib.LoudestArgs()
    args.variant = variant
    args.Mb, args.Qb, args.Zb, args.F, args.R, args.L = Mb, Qb, Zb, F, R, L
    args.r0 = int(r0)
    args.seed = _seed(seed)
    args.normal_threshold = float(int(normal_threshold))     # `long thresh` (cyutils.pyx:1268)
    args.number = number.data_ptr()
    args.h2fdf = h2fdf.data_ptr()
    args.order = order.data_ptr()
    args.counts = cnt.data_ptr() if cnt is not None else None

    out = {}
    out["hc2ss"] = _lib.empty((F, R, L))
    out["hc2bg"] = _lib.empty((F, R))
    args.hc2ss = out["hc2ss"].data_ptr()
    args.hc2bg = out["hc2bg"].data_ptr()
    if variant != 1:
        mt, mr, rz = [_lib.to_dev(vv) for vv in (mt, mr, rz)]
        assert mt.numel() == Mb and mr.numel() == Qb and rz.numel() == Zb
        keep += [mt, mr, rz]
        args.mt, args.mr, args.rz = mt.data_ptr(), mr.data_ptr(), rz.data_ptr()
    if variant == 2:
        out["lspar"] = _lib.empty((3, F, R))
        out["bgpar"] = _lib.empty((3, F, R))
        out["ssidx"] = _lib.empty((3, F, R, L), dtype=torch.int64)
        args.lspar, args.bgpar, args.ssidx = out["lspar"].data_ptr(), out["bgpar"].data_ptr(), out["ssidx"].data_ptr()
    if variant == 3:
        extra = [_lib.to_dev(vv) for vv in (redz_final, dcom_final, sepa, angs)]
        for ee in extra:
            assert ee.shape == number.shape
        keep += extra
        args.redz_final, args.dcom_final, args.sepa, args.angs = [ee.data_ptr() for ee in extra]
        out["sspar"] = _lib.empty((4, F, R, L))
        out["bgpar"] = _lib.empty((7, F, R))
        args.sspar, args.bgpar = out["sspar"].data_ptr(), out["bgpar"].data_ptr()

    # fused product: an independently drawn realised GWB of the same grid from the same pass (lib_tools.run_model)
    Rg = 0 if gwb_nreals is None else int(gwb_nreals)
    if Rg > 0:
        assert cnt is None, "the fused GWB is not available in supplied-count mode"
        out["gwb"] = _lib.empty((F, Rg))
        args.gwb = out["gwb"].data_ptr()
        args.gwb_R, args.gwb_r0, args.gwb_seed = Rg, int(gwb_r0), _seed(gwb_seed)

    if defer is not None:
        # One attempt with the library's default head / bucket, the overflow flags left on the device: `defer` (a list)
        # receives an int32 view of them; the caller checks it at its next synchronisation point and repeats the call
        # without `defer` if either is set (never observed at the named configurations).
        args.bucket_cap, args.head_margin, args.defer_check = 0, 0.0, 1
        nbytes = lib.holo_loudest_workspace_bytes(variant, ncell, F, R, L, 0)
        if Rg > 0:
            nbytes += lib.holo_realize_workspace_bytes(0, ncell, F, Rg)
        ws = _workspace(nbytes)
        args.workspace, args.workspace_bytes = ws.data_ptr(), ws.numel()
        rc = lib.holo_loudest(C.byref(args), _lib.stream())
        STATS["loudest_calls"] += 1
        _lib.check(rc, "loudest")
        defer.append(ws[:8].view(torch.int32))
        return out

    cap, margin = 0, 0.0
    for attempt in range(_MAX_RETRY):
        args.bucket_cap = cap
        args.head_margin = margin
        nbytes = lib.holo_loudest_workspace_bytes(variant, ncell, F, R, L, cap)
        if Rg > 0:
            nbytes += lib.holo_realize_workspace_bytes(0, ncell, F, Rg)
        ws = _workspace(nbytes)
        args.workspace = ws.data_ptr()
        args.workspace_bytes = ws.numel()
        rc = lib.holo_loudest(C.byref(args), _lib.stream())
        STATS["loudest_calls"] += 1
        if rc != 3:
            break
        STATS["loudest_retries"] += 1
        # bucket overflow / head too short (HOLO_ERR_OVERFLOW): enlarge the head margin and redo the draws, with the
        # bucket sized FROM the margin (about L + margin events are expected per bucket)
        margin, cap = _retry_schedule(L, attempt + 1)
    _lib.check(rc, "loudest")
    del keep
    return out


def loudest_hc_from_sorted(number, h2fdf, nreals, nloudest, msort, qsort, zsort, normal_threshold=1e10, *,
                           seed=None, counts=None, r0=0, device=False, gwb_nreals=None, gwb_seed=None, gwb_r0=0,
                           order=None, _defer=None):
    """Characteristic strain of the `nloudest` loudest single sources and of the background of all
    other sources (cyutils.pyx:1220-1344).

    Bins are visited in the order given by (msort, qsort, zsort) -- loudest first at the FIRST
    frequency (single_sources.py:89) -- and the first L binaries found take the L slots (a bin
    holding n binaries takes up to n slots); bins whose draw is < 1 are skipped entirely.

    Returns ``hc2ss`` (F, R, L), ``hc2bg`` (F, R).  Keyword-only addition ``gwb_nreals``: the same pass over the
    grid also draws that many independent realizations of the total GWB (what ``sam_poisson_gwb`` computes, own
    ``gwb_seed``), appended to the result as ``gwb`` (F, gwb_nreals).
    """
    out = _loudest(1, number, h2fdf, nreals, nloudest, msort, qsort, zsort, normal_threshold,
                   seed=seed, counts=counts, r0=r0, gwb_nreals=gwb_nreals, gwb_seed=gwb_seed, gwb_r0=gwb_r0, order=order,
                   defer=_defer)
    res = (_out(out["hc2ss"], device), _out(out["hc2bg"], device))
    return res + ((_out(out["gwb"], device),) if "gwb" in out else ())


def loudest_hc_and_par_from_sorted(number, h2fdf, nreals, nloudest, mt, mr, rz, msort, qsort, zsort,
                                   normal_threshold=1e10, *, seed=None, counts=None, r0=0, device=False):
    """As :func:`loudest_hc_from_sorted`, plus hc^2-weighted mean (M, q, z) of the loudest sources
    (``lspar``) and of the background (``bgpar``), and the grid indices of the loudest sources
    (cyutils.pyx:1347-1538).

    Returns ``hc2ss`` (F,R,L), ``hc2bg`` (F,R), ``lspar`` (3,F,R), ``bgpar`` (3,F,R), ``ssidx`` (3,F,R,L) int64.
    """
    out = _loudest(2, number, h2fdf, nreals, nloudest, msort, qsort, zsort, normal_threshold,
                   mt=mt, mr=mr, rz=rz, seed=seed, counts=counts, r0=r0)
    return tuple(_out(out[kk], device) for kk in ("hc2ss", "hc2bg", "lspar", "bgpar", "ssidx"))


def loudest_hc_and_par_from_sorted_redz(number, h2fdf, nreals, nloudest, mt, mr, rz, redz_final, dcom_final,
                                        sepa, angs, msort, qsort, zsort, normal_threshold=1e10, *,
                                        seed=None, counts=None, r0=0, device=False, gwb_nreals=None, gwb_seed=None,
                                        gwb_r0=0, order=None, _defer=None):
    """As :func:`loudest_hc_from_sorted` for self-consistent hardening: per-source parameters
    ``sspar`` = (M, q, z_initial, z_final) and hc^2-weighted background means ``bgpar`` =
    (M, q, z_initial, z_final, d_c, a, theta) (cyutils.pyx:1541-1767).  Bins with ``h2fdf == 0`` are
    skipped (cyutils.pyx:1727).

    Returns ``hc2ss`` (F,R,L), ``hc2bg`` (F,R), ``sspar`` (4,F,R,L), ``bgpar`` (7,F,R) (and ``gwb`` (F, gwb_nreals)
    with the keyword-only addition ``gwb_nreals``, see :func:`loudest_hc_from_sorted`).
    """
    out = _loudest(3, number, h2fdf, nreals, nloudest, msort, qsort, zsort, normal_threshold,
                   mt=mt, mr=mr, rz=rz, redz_final=redz_final, dcom_final=dcom_final, sepa=sepa, angs=angs,
                   seed=seed, counts=counts, r0=r0, gwb_nreals=gwb_nreals, gwb_seed=gwb_seed, gwb_r0=gwb_r0, order=order,
                   defer=_defer)
    return tuple(_out(out[kk], device) for kk in ("hc2ss", "hc2bg", "sspar", "bgpar") + (("gwb",) if "gwb" in out else ()))


def _ss_bg(number, h2fdf, nreals, normal_threshold, mt=None, mr=None, rz=None, seed=None, counts=None, r0=0):
    import torch
    lib = _lib.require_gpu()
    number = _lib.to_dev(number)
    h2fdf = _lib.to_dev(h2fdf)
    assert number.shape == h2fdf.shape and number.dim() == 4
    Mb, Qb, Zb, F = [int(ss) for ss in number.shape]
    ncell = Mb * Qb * Zb
    R = int(nreals)
    cnt = _counts(counts, R, F, ncell)
    par = mt is not None
    hc2ss = _lib.empty((F, R))
    hc2bg = _lib.empty((F, R))
    ssidx = _lib.empty((3, F, R), dtype=torch.int64)
    bgpar = sspar = None
    if par:
        mt, mr, rz = [_lib.to_dev(vv) for vv in (mt, mr, rz)]
        bgpar = _lib.empty((3, F, R))
        sspar = _lib.empty((3, F, R))
    ws = _workspace(lib.holo_realize_workspace_bytes(5 if par else 4, ncell, F, R))
    rc = lib.holo_ss_bg_hc(
        _lib.ptr(number), _lib.ptr(h2fdf), Mb, Qb, Zb, F, R, int(r0), _seed(seed), float(int(normal_threshold)),
        _lib.ptr(cnt), _lib.ptr(mt), _lib.ptr(mr), _lib.ptr(rz), _lib.ptr(hc2ss), _lib.ptr(hc2bg),
        _lib.ptr(ssidx), _lib.ptr(bgpar), _lib.ptr(sspar), _lib.ptr(ws), ws.numel(), _lib.stream())
    if rc == 3:
        # the reference executes a bare `raise` here (cyutils.pyx:1157-1158)
        raise RuntimeError("No active exception to reraise")
    _lib.check(rc, "ss_bg_hc")
    return hc2ss, hc2bg, ssidx, bgpar, sspar


def ss_bg_hc(number, h2fdf, nreals, normal_threshold=1e10, *, seed=None, counts=None, r0=0, device=False):
    """Loudest single source (arg-max of h2fdf over occupied bins) and background of the rest
    (cyutils.pyx:900-1014).  Returns ``hc2ss`` (F,R), ``hc2bg`` (F,R), ``ssidx`` (3,F,R) (-1 if none)."""
    hc2ss, hc2bg, ssidx, _, _ = _ss_bg(number, h2fdf, nreals, normal_threshold, seed=seed, counts=counts, r0=r0)
    return _out(hc2ss, device), _out(hc2bg, device), _out(ssidx, device)


def ss_bg_hc_and_par(number, h2fdf, nreals, mt, mr, rz, normal_threshold=1e10, *, seed=None, counts=None, r0=0,
                     device=False):
    """As :func:`ss_bg_hc` plus background-average and single-source (M, q, z) (cyutils.pyx:1017-1178).
    Returns ``hc2ss, hc2bg, ssidx, bgpar, sspar``."""
    vals = _ss_bg(number, h2fdf, nreals, normal_threshold, mt=mt, mr=mr, rz=rz, seed=seed, counts=counts, r0=r0)
    return tuple(_out(vv, device) for vv in vals)


# ==================================================================================================
# K5
# ==================================================================================================

def _eccen(ndens, mtot_log10, mrat, redz, dcom, gwfobs, sepa_evo, eccen_evo, nharms, nreals, seed, r0, device):
    lib = _lib.require_gpu()
    ndens = _lib.to_dev(ndens)
    M, Q, Z = [int(ss) for ss in ndens.shape]
    arrs = [_lib.to_dev(np.asarray(vv, dtype=float)) for vv in (mtot_log10, mrat, redz, dcom, gwfobs, sepa_evo, eccen_evo)]
    assert arrs[0].numel() == M and arrs[1].numel() == Q and arrs[2].numel() == Z and arrs[3].numel() == Z
    F = arrs[4].numel()
    E = arrs[5].numel()
    assert arrs[6].numel() == E
    H = int(nharms)
    R = int(nreals)
    gwb = _lib.empty((F, H) if R == 0 else (F, H, R))
    ws = _workspace(lib.holo_eccen_workspace_bytes(M, Q, Z, F, H, R))
    rc = lib.holo_sam_calc_gwb_single_eccen(
        _lib.cy_consts(), _lib.cy_gw_src_const(),
        _lib.ptr(ndens), *[_lib.ptr(aa) for aa in arrs], M, Q, Z, F, E, H, R, int(r0), _seed(seed),
        _lib.ptr(gwb), _lib.ptr(ws), ws.numel(), _lib.stream())
    _lib.check(rc, "sam_calc_gwb_single_eccen")
    return _out(gwb, device)


def sam_calc_gwb_single_eccen(ndens, mtot_log10, mrat, redz, dcom, gwfobs, sepa_evo, eccen_evo, nharms=100, *,
                              device=False):
    """GWB hc^2 per (frequency, harmonic) for a single eccentricity track e(a) (cyutils.pyx:361-597).
    ``dcom`` in [Mpc]; returns ``gwb`` (F, H)."""
    return _eccen(ndens, mtot_log10, mrat, redz, dcom, gwfobs, sepa_evo, eccen_evo, nharms, 0, 0, 0, device)


def sam_calc_gwb_single_eccen_discrete(ndens, mtot_log10, mrat, redz, dcom, gwfobs, sepa_evo, eccen_evo, nharms,
                                       nreals, *, seed=None, r0=0, device=False):
    """Poisson-discretised version of :func:`sam_calc_gwb_single_eccen` (cyutils.pyx:600-851);
    returns ``gwb`` (F, H, R)."""
    return _eccen(ndens, mtot_log10, mrat, redz, dcom, gwfobs, sepa_evo, eccen_evo, nharms, int(nreals), seed, r0,
                  device)
