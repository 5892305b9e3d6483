"""GW strain from gridded SAM populations (hot-path subset of ``holodeck/gravwaves.py``).

Public functions keep the reference's names and positional signatures:

* :func:`char_strain_sq_from_bin_edges_redz`   (``gravwaves.py:694-725``)
* :func:`char_strain_sq_from_bin_edges`        (``gravwaves.py:760-783``)
* :func:`_gws_from_number_grid_integrated_redz` (``gravwaves.py:470-542``)
* :func:`_gws_from_number_grid_integrated`     (``gravwaves.py:545-616``)
* :func:`poisson_as_needed`                    (``gravwaves.py:666-691``)
* :func:`gwb_ideal`                            (``gravwaves.py:619-663``; host, analytic check value)
* :func:`sam_calc_gwb_single_eccen[_discrete]` (``gravwaves.py:924-981``)

Arrays may be numpy (results come back as numpy) or CUDA ``torch`` tensors (results stay on the
device).  All O(grid) work runs in ``libholo_b200.so``; there is no CPU fallback.
"""
import ctypes as C

import numpy as np

import holodeck_b200 as holo
from holodeck_b200 import _lib, cosmo, log, utils
from holodeck_b200.constants import NWTG, SPLC, MPC


def _np(xx):
    return np.asarray(xx.cpu()) if _lib.is_device_array(xx) else np.asarray(xx, dtype=float)


def _strain_inputs(edges):
    """Host-side 1-D preparations shared by the strain kernels: bin centres, fc, fc/df."""
    assert len(edges) == 4
    edges = [_np(ee) for ee in edges]
    assert np.all([np.ndim(ee) == 1 for ee in edges])
    foo = edges[-1]                     #: observer-frame orbital-frequency bin edges
    df = np.diff(foo)                   #: frequency bin widths
    fc = utils.midpoints(foo)           #: frequency-bin centers
    mt = utils.midpoints(edges[0])
    mr = utils.midpoints(edges[1])
    rz = utils.midpoints(edges[2])
    return edges, mt, mr, rz, fc, fc / df


def _char_strain_sq(edges, redz, params=False, device=False, dnum=None):
    """Launch K2b (optionally fused with K2 when `dnum` is given).

    Returns a dict with ``h2fdf`` and, if requested, ``number``, ``zmid``, ``dcom``, ``sepa``, ``angs``
    (all CUDA tensors shaped (M-1, Q-1, Z-1, F)).
    """
    lib = _lib.require_gpu()
    edges, mt, mr, rz, fc, fdf = _strain_inputs(edges)
    M, Q, Z = [len(ee) for ee in edges[:3]]
    F = fc.size
    shape = (M - 1, Q - 1, Z - 1, F)
    redz_d = None
    if redz is not None:
        redz_d = _lib.to_dev(redz)
        assert tuple(redz_d.shape) == (M, Q, Z, F), f"`redz` shape {tuple(redz_d.shape)} != {(M, Q, Z, F)}"
    cp = _lib.cosmo_params(cosmo)
    dc_tab, dc_n, dc_wmax = _lib.dc_table(cosmo)
    mt_d, mr_d, rz_d, fc_d, fdf_d = [_lib.to_dev(vv) for vv in (mt, mr, rz, fc, fdf)]
    import torch
    out = dict(h2fdf=_lib.empty(shape))
    # set by the kernel when `redz` holds a negative value other than the -1 sentinel (checked by ss_gws_redz)
    out["bad_redz"] = torch.zeros(1, dtype=torch.int32, device=_lib.device()) if redz_d is not None else None
    names = ("zmid", "dcom", "sepa", "angs")
    for nn in names:
        out[nn] = _lib.empty(shape) if params else None
    if dnum is None:
        rc = lib.holo_char_strain_sq(
            C.byref(cp), utils._GW_SRC_CONST, NWTG, _lib.ptr(redz_d), _lib.ptr(rz_d), _lib.ptr(mt_d), _lib.ptr(mr_d),
            _lib.ptr(fc_d), _lib.ptr(fdf_d), M, Q, Z, F, _lib.ptr(out["h2fdf"]),
            *[_lib.ptr(out[nn]) for nn in names], _lib.ptr(out["bad_redz"]), _lib.ptr(dc_tab), dc_n, dc_wmax, _lib.stream())
        _lib.check(rc, "char_strain_sq")
    else:
        dnum_d = _lib.to_dev(dnum)
        assert tuple(dnum_d.shape) == (M, Q, Z, F)
        l10m, mrat_d, redz_e, dlnf = [_lib.to_dev(vv) for vv in
                                      (np.log10(edges[0]), edges[1], edges[2], np.diff(np.log(edges[3])))]
        out["number"] = _lib.empty(shape)
        rc = lib.holo_integrate_and_strain(
            C.byref(cp), utils._GW_SRC_CONST, NWTG, _lib.ptr(l10m), _lib.ptr(mrat_d), _lib.ptr(redz_e), _lib.ptr(dlnf),
            _lib.ptr(dnum_d), _lib.ptr(redz_d), _lib.ptr(mt_d), _lib.ptr(mr_d), _lib.ptr(fc_d), _lib.ptr(fdf_d),
            M, Q, Z, F, _lib.ptr(out["number"]), _lib.ptr(out["h2fdf"]),
            *[_lib.ptr(out[nn]) for nn in names], _lib.ptr(out["bad_redz"]), _lib.ptr(dc_tab), dc_n, dc_wmax, _lib.stream())
        _lib.check(rc, "integrate_and_strain")
    return out


def char_strain_sq_from_bin_edges_redz(edges, redz, device=None):
    """hc^2 = hs^2 * f/df of one binary in each grid bin, using each bin's *final* redshift.

    ``edges`` : (4,) list of (M,), (Q,), (Z,), (F+1,) edge arrays (orbital, observer-frame frequency);
    ``redz``  : (M, Q, Z, F) redshift at each grid EDGE point (``-1`` where the binary never reaches
    that frequency).  Bin-centre redshifts are the 8-corner means (sentinels included); bins whose
    mean is <= 0 get infinite distance, i.e. zero strain.  Mirrors ``gravwaves.py:694-725``.
    """
    on_dev = _lib.is_device_array(redz) if device is None else device
    out = _char_strain_sq(edges, redz)["h2fdf"]
    return out if on_dev else _lib.to_host(out)


def char_strain_sq_from_bin_edges(edges, device=False):
    """As above but every bin sits at its initial (bin-centre) redshift (``gravwaves.py:760-783``)."""
    out = _char_strain_sq(edges, None)["h2fdf"]
    return out if device else _lib.to_host(out)


def poisson_as_needed(values, thresh=1e10, *, seed=None, device=None):
    """Poisson draws of ``values``; floor(Normal(v, sqrt(v))) above ``thresh`` (``gravwaves.py:666-691``)."""
    from holodeck_b200.cyutils import _seed
    lib = _lib.require_gpu()
    on_dev = _lib.is_device_array(values) if device is None else device
    lam = _lib.to_dev(values)
    out = _lib.empty(tuple(lam.shape))
    rc = lib.holo_poisson_as_needed(_lib.ptr(lam), lam.numel(), _seed(seed), 0, float(thresh), _lib.ptr(out),
                                    _lib.stream())
    _lib.check(rc, "poisson_as_needed")
    return out if on_dev else _lib.to_host(out)


def _expectation(number, hc2):
    lib = _lib.require_gpu()
    F = number.shape[-1]
    out = _lib.empty((F,))
    ws = _lib.empty((lib.holo_gwb_expectation_workspace_bytes(F) // 8,))
    rc = lib.holo_gwb_expectation(_lib.ptr(number), _lib.ptr(hc2), number.numel() // F, F, _lib.ptr(out), _lib.ptr(ws),
                                  ws.numel() * 8, _lib.stream())
    _lib.check(rc, "gwb_expectation")
    return out


def _gws_from_hc2(hc2, number, realize, sum, seed, r0, on_dev):
    """Shared tail of the two `_gws_from_number_grid_integrated*` functions (gravwaves.py:502-542)."""
    import torch
    from holodeck_b200 import cyutils
    number = _lib.to_dev(number)
    assert number.shape == hc2.shape
    # Create a single realization
    if realize is True:
        hc2 = hc2 * poisson_as_needed(number, seed=seed, device=True)
        if sum:
            hc2 = torch.sum(hc2, dim=(0, 1, 2))
    # Do not create a discrete realization, use the expectation values directly
    elif realize in [None, False]:
        if sum:
            hc2 = _expectation(number, hc2)
        else:
            hc2 = hc2 * number
    # Create multiple discrete realizations
    elif utils.isinteger(realize):
        if sum:
            hc2 = cyutils.sam_poisson_gwb(number, hc2, realize, seed=seed, r0=r0, device=True)
        else:
            log.warning(f"`sum`={sum} :: this requires a large amount of memory!")
            shape = tuple(number.shape) + (int(realize),)
            draws = poisson_as_needed(number[..., None].expand(shape).contiguous(), seed=seed, device=True)
            hc2 = hc2[..., None] * draws
    else:
        err = "`realize` ({}) must be one of {{True, False, integer}}!".format(realize)
        log.error(err)
        raise ValueError(err)
    hc = torch.sqrt(hc2)
    return hc if on_dev else _lib.to_host(hc)


def _gws_from_number_grid_integrated_redz(edges, redz, number, realize, sum=True, *, seed=None, r0=0, device=None):
    """Characteristic strain of the GWB from a grid of binary numbers and final redshifts.

    ``realize``: ``False``/``None`` -> expectation value; ``True`` -> one Poisson realization;
    integer R -> R realizations through ``cyutils.sam_poisson_gwb`` (returns (F, R)).
    Mirrors ``gravwaves.py:470-542``.
    """
    on_dev = _lib.is_device_array(number) if device is None else device
    hc2 = _char_strain_sq(edges, redz)["h2fdf"]
    return _gws_from_hc2(hc2, number, realize, sum, seed, r0, on_dev)


def _gws_from_number_grid_integrated(edges, number, realize, sum=True, *, seed=None, r0=0, device=None):
    """As :func:`_gws_from_number_grid_integrated_redz` with bins at their initial redshifts
    (``gravwaves.py:545-616``)."""
    on_dev = _lib.is_device_array(number) if device is None else device
    hc2 = _char_strain_sq(edges, None)["h2fdf"]
    return _gws_from_hc2(hc2, number, realize, sum, seed, r0, on_dev)


def gwb_ideal(fobs_gw, ndens, mtot, mrat, redz, dlog10, sum=True):
    """Idealised GWB amplitude, [Phinney2001]_ Eq.5 (``gravwaves.py:619-663``).  Host numpy: this is an
    O(M Q Z) analytic check value, not part of the realised-GWB hot path."""
    const = ((4.0 * np.pi) / (3 * SPLC**2))
    mc = utils.chirp_mass_mtmr(mtot, mrat)
    mc = np.power(NWTG * mc, 5.0/3.0)
    rz = np.power(1 + redz, -1.0/3.0)
    fogw = np.power(np.pi * fobs_gw, -4.0/3.0)
    integ = ndens * mc * rz
    redz = redz * np.ones_like(integ)
    integ[redz <= 0.0] = 0.0
    arguments = [mtot, mrat, redz]
    if dlog10:
        arguments[0] = np.log10(arguments[0])
    for ax, xx in enumerate(arguments):
        integ = np.moveaxis(integ, ax, 0)
        xx = np.moveaxis(xx, ax, 0)
        try:
            integ = 0.5 * (integ[:-1] + integ[1:]) * np.diff(xx, axis=0)
        except ValueError:
            for jj in range(1, len(arguments)):
                sh = np.shape(xx)[jj]
                if (sh == 1) or (sh == np.shape(integ)[jj]):
                    continue
                xx = np.moveaxis(xx, jj, 0)
                xx = 0.5 * (xx[:-1] + xx[1:])
                xx = np.moveaxis(xx, 0, jj)
            integ = 0.5 * (integ[:-1] + integ[1:]) * np.diff(xx, axis=0)
        integ = np.moveaxis(integ, 0, ax)
    gwb = const * fogw
    gwb = gwb * np.sum(integ) if sum else gwb * integ
    return np.sqrt(gwb)


def sam_calc_gwb_single_eccen(gwfobs, sam, sepa_evo, eccen_evo, nharms=100):
    """Eccentric-SAM GWB: hc^2 per (frequency, harmonic) (``gravwaves.py:924-933``)."""
    from holodeck_b200 import cyutils
    dens_of = getattr(sam, "_static_binary_density_device", None)
    ndens = dens_of() if dens_of is not None else sam.static_binary_density
    return cyutils.sam_calc_gwb_single_eccen(
        ndens, np.log10(sam.mtot), sam.mrat, sam.redz, cosmo.comoving_distance(sam.redz) / MPC,
        gwfobs, sepa_evo, eccen_evo, nharms
    )


def sam_calc_gwb_single_eccen_discrete(gwfobs, sam, sepa_evo, eccen_evo, nharms=100, nreals=None, *, seed=None):
    """Discretised (Poisson) eccentric-SAM GWB: (F, H, R); squeezed to (F, H) if ``nreals`` is None
    (``gravwaves.py:936-981``)."""
    from holodeck_b200 import cyutils
    dens_of = getattr(sam, "_static_binary_density_device", None)
    ndens = dens_of() if dens_of is not None else sam.static_binary_density
    if nreals is None:
        nreals = 1
        squeeze = True
    else:
        squeeze = False
    gwb = cyutils.sam_calc_gwb_single_eccen_discrete(
        ndens, np.log10(sam.mtot), sam.mrat, sam.redz, cosmo.comoving_distance(sam.redz) / MPC,
        gwfobs, sepa_evo, eccen_evo, nharms, nreals, seed=seed
    )
    if squeeze:
        gwb = gwb.squeeze()
    return np.asarray(gwb)
