"""Single-source / background split (hot-path subset of ``holodeck/single_sources.py``).

* :func:`ss_gws_redz` (``single_sources.py:40-173``) -- what ``sam.gwb`` and ``librarian.run_model`` call
* :func:`ss_gws`      (``single_sources.py:177-275``) -- same without final redshifts

Arrays may be numpy or CUDA ``torch`` tensors; the per-bin strain, rank ordering, Poisson draws and
loudest-source selection all run on the device.  The plotting / example helpers and the older
``loudest_by_cython`` / ``ss_by_*`` variants of the reference are out of scope.
"""
import numpy as np

import holodeck_b200 as holo
from holodeck_b200 import _lib, utils, gravwaves, cyutils


def _rank_order(h2fdf, shape):
    """Bins sorted from largest to smallest h2fdf at the FIRST frequency (``single_sources.py:89-93``).

    The reference uses ``np.argsort`` (unstable quicksort) on ``-h2fdf[...,0]``; ties (the many bins
    with h=0) come out in an implementation-defined order there.  Here ties keep ascending bin index
    (stable sort), which is one of the orders the reference may produce.
    Returns (msort, qsort, zsort) as CUDA int64 tensors.
    """
    import torch
    indices = _rank_order_flat(h2fdf).to(dtype=torch.int64)
    _, Qb, Zb = shape
    zsort = indices % Zb
    mq = indices // Zb
    return mq // Qb, mq % Qb, zsort


def _rank_order_flat(h2fdf):
    """The same order as :func:`_rank_order` as flat int32 bin indices ``(m*Q + q)*Z + z`` on the device -- what the
    loudest kernels consume (the index triple of the reference's signature is only built when a caller asks)."""
    import torch
    key = -h2fdf[..., 0].reshape(-1)
    return torch.sort(key, stable=True).indices.to(torch.int32)


class DeferredChecks:
    """The host-side checks of one deferred `ss_gws_redz` call: device flags that are looked at later, at the caller's
    next synchronisation point (`librarian.gen_lib` pipelines one model behind).  `verify()` raises what the
    synchronous call would have raised; `overflow()` says whether the loudest split has to be redone with a larger
    head (HOLO_ERR_OVERFLOW)."""

    def __init__(self):
        self.loudest_flags = []      # int32[2] views: event-bucket overflow, head too short
        self.bad_redz = None         # strain kernel's `redz < 0 and != -1` flag (int32[1]) or None
        self.bad_sspar = None        # bool tensor: sspar[3] < 0 and != -1

    def overflow(self):
        return any(bool(ff.any().item()) for ff in self.loudest_flags)

    def verify(self):
        if self.bad_redz is not None and int(self.bad_redz.item()) != 0:
            raise ValueError("redz < 0 and !=-1 found in redz, in ss_gws_redz()")
        if self.bad_sspar is not None and bool(self.bad_sspar.any().item()):
            raise ValueError("check 1: sspar[3] values are negative and not -1 in sings.ss_gws_redz()")


def ss_gws_redz(edges, redz, number, realize, loudest=1, params=False, *, seed=None, r0=0, device=False,
                _precomputed=None, _gwb=None, _deferred=None):
    """Strain of the `loudest` loudest single sources and of the background, per frequency and realization.

    Parameters mirror ``single_sources.ss_gws_redz`` (``single_sources.py:40-85``):
    ``edges`` (4,) list of edge arrays (M), (Q), (Z), (F+1) in orbital observer-frame frequency;
    ``redz`` (M,Q,Z,F) final redshifts at grid edges; ``number`` (M-1,Q-1,Z-1,F) binaries per bin;
    ``realize`` integer number of realizations.

    Returns ``hc_ss`` (F,R,L), ``hc_bg`` (F,R) and, if ``params``, ``sspar`` (4,F,R,L), ``bgpar`` (7,F,R)
    (numpy arrays; CUDA tensors with the keyword-only addition ``device=True``).
    ``_gwb=(nreals, seed)`` (used by ``librarian.run_model``) appends ``gwb`` (F, nreals): the characteristic strain
    of an independently drawn realised GWB of the same grid, produced by the same pass of the realization kernel.
    """
    import torch
    _lib.require_gpu()
    host = (lambda tt: tt) if device else _lib.to_host
    gkw = {} if _gwb is None else dict(gwb_nreals=int(_gwb[0]), gwb_seed=_gwb[1])
    edges_np = [np.asarray(ee.cpu()) if _lib.is_device_array(ee) else np.asarray(ee, dtype=float) for ee in edges]
    # All other bin midpoints
    mt = utils.midpoints(edges_np[0])   #: total mass
    mr = utils.midpoints(edges_np[1])   #: mass ratio
    rz = utils.midpoints(edges_np[2])   #: initial redshift
    shape = (mt.size, mr.size, rz.size)

    redz_d = _lib.to_dev(redz)
    number_d = _lib.to_dev(number)

    # hsfdf = hsamp^2 * f/df  (and the params=True glue of single_sources.py:112-139, same kernel)
    if _precomputed is not None:
        strain = _precomputed
    else:
        strain = gravwaves._char_strain_sq(edges_np, redz_d, params=bool(params))
    h2fdf = strain["h2fdf"]

    # indices of bins sorted by h2fdf, just for the first frequency
    order = _rank_order_flat(h2fdf)

    # `redz` must be non-negative or the -1 sentinel (single_sources.py:95-99).  The strain kernel raises a flag
    # while it reads the values (strains computed elsewhere are checked in separate passes); the flag is read after
    # the draws, whose call synchronises anyway, so the check costs no extra pass and no extra stall.
    def check_redz():
        if _deferred is not None and strain.get("bad_redz") is not None:
            _deferred.bad_redz = strain["bad_redz"]
            return
        flag = strain.get("bad_redz")
        if (flag is not None and int(flag.item()) != 0) or \
                (flag is None and bool(torch.any(torch.logical_and(redz_d < 0, redz_d != -1)))):
            err = int(torch.sum(torch.logical_and(redz_d < 0, redz_d != -1)))
            err = f"{err} redz < 0 and !=-1 found in redz, in ss_gws_redz()"
            raise ValueError(err)

    if not utils.isinteger(realize):
        raise Exception("`realize` ({}) must be an integer!")

    if params is True or params:
        hc2ss, hc2bg, sspar, bgpar, *extra = cyutils.loudest_hc_and_par_from_sorted_redz(
            number_d, h2fdf, realize, loudest,
            mt, mr, rz, strain["zmid"], strain["dcom"], strain["sepa"], strain["angs"],
            None, None, None, seed=seed, r0=r0, device=True, order=order,
            _defer=None if _deferred is None else _deferred.loudest_flags, **gkw)
        check_redz()
        extra = tuple(host(torch.sqrt(ee)) for ee in extra)
        hc_ss = host(torch.sqrt(hc2ss))
        hc_bg = host(torch.sqrt(hc2bg))
        sspar = host(sspar)
        bgpar = host(bgpar)
        # check that all final redshifts are positive or -1
        bad = (sspar[3] < 0) & (sspar[3] != -1)
        if _deferred is not None and device:
            _deferred.bad_sspar = bad
        elif bool(bad.any()):
            err = int(bad.sum())
            err = f"check 1: {err} out of {int(np.prod(sspar[3].shape))} sspar[3] are negative and not -1 in sings.ss_gws_redz()"
            raise ValueError(err)
        return (hc_ss, hc_bg, sspar, bgpar) + extra

    hc2ss, hc2bg, *extra = cyutils.loudest_hc_from_sorted(number_d, h2fdf, realize, loudest, None, None, None,
                                                          seed=seed, r0=r0, device=True, order=order,
                                                          _defer=None if _deferred is None else _deferred.loudest_flags, **gkw)
    check_redz()
    hc_ss = host(torch.sqrt(hc2ss))
    hc_bg = host(torch.sqrt(hc2bg))
    return (hc_ss, hc_bg) + tuple(host(torch.sqrt(ee)) for ee in extra)


def ss_gws(edges, number, realize, loudest=1, params=False, *, seed=None, r0=0):
    """As :func:`ss_gws_redz` with every bin at its initial redshift (``single_sources.py:177-275``).

    With ``params=True`` returns ``hc_ss, hc_bg, sspar (3,F,R,L), bgpar (3,F,R)`` where ``sspar`` holds
    the (M, q, z) bin-centre values of each loud source (``single_sources.py:255-263``).
    """
    import torch
    _lib.require_gpu()
    edges_np = [np.asarray(ee.cpu()) if _lib.is_device_array(ee) else np.asarray(ee, dtype=float) for ee in edges]
    mt = utils.midpoints(edges_np[0])
    mr = utils.midpoints(edges_np[1])
    rz = utils.midpoints(edges_np[2])
    shape = (mt.size, mr.size, rz.size)
    number_d = _lib.to_dev(number)
    h2fdf = gravwaves._char_strain_sq(edges_np, None)["h2fdf"]
    msort, qsort, zsort = _rank_order(h2fdf, shape)
    if not utils.isinteger(realize):
        raise Exception("`realize` ({}) must be an integer!")
    if params:
        hc2ss, hc2bg, lspar, bgpar, ssidx = cyutils.loudest_hc_and_par_from_sorted(
            number_d, h2fdf, realize, loudest, mt, mr, rz, msort, qsort, zsort, seed=seed, r0=r0, device=True)
        ssidx = _lib.to_host(ssidx)
        hc_ss = _lib.to_host(torch.sqrt(hc2ss))
        hc_bg = _lib.to_host(torch.sqrt(hc2bg))
        sspar = np.array([mt[ssidx[0]], mr[ssidx[1]], rz[ssidx[2]]])
        return hc_ss, hc_bg, sspar, _lib.to_host(bgpar)
    hc2ss, hc2bg = cyutils.loudest_hc_from_sorted(number_d, h2fdf, realize, loudest, msort, qsort, zsort,
                                                  seed=seed, r0=r0, device=True)
    return _lib.to_host(torch.sqrt(hc2ss)), _lib.to_host(torch.sqrt(hc2bg))
