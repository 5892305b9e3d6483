"""Drop-in for the reference's compiled module ``holodeck.sams.sam_cyutils``.

Same callables, same argument meaning and error behaviour as ``holodeck/sams/sam_cyutils.pyx``;
the loops run as sm_100a kernels in ``libholo_b200.so`` (``include/holo_b200.h``) instead of Cython.

Array convention: numpy in -> numpy out (drop-in).  If the array arguments are CUDA ``torch``
tensors the results stay on the device (this is what ``Semi_Analytic_Model.gwb`` uses internally so
the 238 MB ``(M,Q,Z,F)`` grids never cross PCIe).  Extra keyword-only arguments (``device``) are
additions; positional signatures are the reference's.
"""
import numpy as np

import holodeck_b200 as holo
from holodeck_b200 import _lib

__all__ = [
    "hard_gw", "integrate_differential_number_3dx1d", "hard_func_2pwl_gw", "find_2pwl_hardening_norm",
    "integrate_binary_evolution_2pwl", "dynamic_binary_number_at_fobs",
]


def _out(tensor, like_device):
    """Return `tensor` as numpy unless the caller passed device arrays."""
    if like_device:
        return tensor
    return _lib.to_host(tensor)


def hard_gw(mtot, mrat, sepa):
    """``cpdef double hard_gw`` (sam_cyutils.pyx:45-49): scalar GW hardening rate da/dt [cm/s]."""
    lib = _lib.require_gpu()
    args = [_lib.to_dev(np.atleast_1d(np.asarray(aa, dtype=float))) for aa in (mtot, mrat, sepa)]
    zero = _lib.to_dev(np.zeros(1))
    out = _lib.empty((1,))
    # norm = 0 switches the phenomenological term off: dadt = -0 * ... + hard_gw
    rc = lib.holo_hard_func_2pwl_gw(_lib.cy_consts(), _lib.ptr(args[0]), _lib.ptr(args[1]), _lib.ptr(args[2]),
                                    _lib.ptr(zero), 1.0, 0.0, 0.0, 1, _lib.ptr(out), _lib.stream())
    _lib.check(rc, "hard_gw")
    return float(out.cpu()[0])


def integrate_differential_number_3dx1d(edges, dnum):
    """Integrate the differential number of binaries over each grid bin (sam_cyutils.pyx:115-163).

    Trapezoid over the first 3 dims (log10 mtot, mrat, redz), Riemann over ln(freq).

    * ``edges`` : (4,) list of arrays of lengths M, Q, Z, F+1 (mtot and freq NOT in log space)
    * ``dnum``  : (M, Q, Z, F)
    * returns ``numb`` : (M-1, Q-1, Z-1, F)
    """
    lib = _lib.require_gpu()
    on_dev = _lib.is_device_array(dnum)
    edges = [np.asarray(ee.cpu()) if _lib.is_device_array(ee) else np.asarray(ee) for ee in edges]
    # each edge should have the same length as the corresponding dimension of `dnum`
    shape = [len(ee) for ee in edges]
    err = f"Shape of edges={shape} does not match dnum={tuple(dnum.shape)}"
    # except the last edge (freq), where `dnum` should be 1-shorter
    shape[-1] -= 1
    assert tuple(dnum.shape) == tuple(shape), err
    M, Q, Z, F = shape
    new_shape = (M - 1, Q - 1, Z - 1, F)

    # Convert from  mtot => log10(mtot)  and  freq ==> ln(freq)   (pyx:159)
    l10m = _lib.to_dev(np.log10(edges[0]))
    mrat = _lib.to_dev(edges[1])
    redz = _lib.to_dev(edges[2])
    dlnf = _lib.to_dev(np.diff(np.log(edges[3])))
    dnum_d = _lib.to_dev(dnum)
    numb = _lib.empty(new_shape)
    rc = lib.holo_integrate_differential_number_3dx1d(
        _lib.ptr(l10m), _lib.ptr(mrat), _lib.ptr(redz), _lib.ptr(dlnf), _lib.ptr(dnum_d), _lib.ptr(numb),
        M, Q, Z, F, _lib.stream())
    _lib.check(rc, "integrate_differential_number_3dx1d")
    return _out(numb, on_dev)


def hard_func_2pwl_gw(mtot, mrat, sepa, norm, rchar, gamma_inner, gamma_outer):
    """Total (phenomenological 2-power-law + GW) hardening rate da/dt (sam_cyutils.pyx:271-286).

    All arguments broadcast against each other; ``rchar, gamma_inner, gamma_outer`` must broadcast
    to scalars per element (the reference flattens them too) -- here they must be scalars, which is
    how every reference call site uses them (hardening.py:1447-1450).
    """
    lib = _lib.require_gpu()
    args = np.broadcast_arrays(mtot, mrat, sepa, norm)
    shape = args[0].shape
    mtot, mrat, sepa, norm = [_lib.to_dev(np.ascontiguousarray(aa, dtype=float).reshape(-1)) for aa in args]
    rchar, gamma_inner, gamma_outer = [float(np.asarray(vv).reshape(-1)[0]) for vv in (rchar, gamma_inner, gamma_outer)]
    dadt = _lib.empty((mtot.numel(),))
    rc = lib.holo_hard_func_2pwl_gw(_lib.cy_consts(), _lib.ptr(mtot), _lib.ptr(mrat), _lib.ptr(sepa), _lib.ptr(norm),
                                    rchar, gamma_inner, gamma_outer, mtot.numel(), _lib.ptr(dadt), _lib.stream())
    _lib.check(rc, "hard_func_2pwl_gw")
    return _lib.to_host(dadt).reshape(shape)


def find_2pwl_hardening_norm(time, mtot, mrat, sepa_init, rchar, gamma_inner, gamma_outer, nsteps, device=False):
    """log10 of the 2PL hardening normalisation giving total lifetime ``time`` (sam_cyutils.pyx:289-306).

    One Brent root-find per (mtot, mrat) pair, scipy ``brentq`` semantics (xtol=1e-3, rtol=1e-5,
    maxiter=100 on [-20, +20]).
    """
    assert np.ndim(time) == 0
    assert np.ndim(mtot) == 1
    assert np.shape(mtot) == np.shape(mrat)
    lib = _lib.require_gpu()
    mt = _lib.to_dev(mtot)
    mr = _lib.to_dev(mrat)
    out = _lib.empty((mt.numel(),))
    rc = lib.holo_find_2pwl_hardening_norm(
        _lib.cy_consts(), float(time), _lib.ptr(mt), _lib.ptr(mr), mt.numel(), float(sepa_init), float(rchar),
        float(gamma_inner), float(gamma_outer), int(nsteps), _lib.ptr(out), _lib.stream())
    _lib.check(rc, "find_2pwl_hardening_norm")
    return _out(out, device)


def integrate_binary_evolution_2pwl(norm_log10, mtot, mrat, sepa_init, rchar, gamma_inner, gamma_outer, nsteps):
    """Binary lifetime [s] for the given log10-normalisation (sam_cyutils.pyx:401-413); scalars in, float out."""
    lib = _lib.require_gpu()
    nl = _lib.to_dev(np.atleast_1d(np.asarray(norm_log10, dtype=float)))
    mt = _lib.to_dev(np.atleast_1d(np.asarray(mtot, dtype=float)))
    mr = _lib.to_dev(np.atleast_1d(np.asarray(mrat, dtype=float)))
    out = _lib.empty((mt.numel(),))
    rc = lib.holo_binary_lifetime_2pwl(
        _lib.cy_consts(), _lib.ptr(nl), _lib.ptr(mt), _lib.ptr(mr), mt.numel(), float(sepa_init), float(rchar),
        float(gamma_inner), float(gamma_outer), int(nsteps), _lib.ptr(out), _lib.stream())
    _lib.check(rc, "integrate_binary_evolution_2pwl")
    res = _lib.to_host(out)
    return float(res[0]) if np.ndim(norm_log10) == 0 else res


def dynamic_binary_number_at_fobs(fobs_orb, sam, hard, cosmo, device=False):
    """Differential number of binaries d^4N/[dlog10M dq dz dlnf] at the given orbital frequencies.

    Mirrors ``sam_cyutils.dynamic_binary_number_at_fobs`` (sam_cyutils.pyx:421-504): dispatches on the
    hardening class, reads ``sam.static_binary_density``, ``sam._gmt_time`` / ``sam._redz_prime``,
    ``hard._norm`` etc. and the cosmology interpolation tables ``cosmo._grid_z/_grid_dcom/_grid_age``.

    Returns ``(redz_final, diff_num)``, both (M, Q, Z, F); unreached cells hold -1 / 0.
    """
    lib = _lib.require_gpu()
    on_dev = bool(device)
    dens_of = getattr(sam, "_static_binary_density_device", None)
    nden = dens_of() if dens_of is not None else sam.static_binary_density
    nden = _lib.to_dev(nden)

    fobs = _lib.to_dev(np.asarray(fobs_orb.cpu() if _lib.is_device_array(fobs_orb) else fobs_orb, dtype=float))
    F = fobs.numel()
    M, Q, Z = sam.shape
    shape = tuple(sam.shape) + (F,)
    mtot = _lib.to_dev(sam.mtot)
    mrat = _lib.to_dev(sam.mrat)
    redz = _lib.to_dev(sam.redz)
    grid_z = _lib.to_dev(cosmo._grid_z)
    grid_dcom = _lib.to_dev(cosmo._grid_dcom)
    redz_final = _lib.empty(shape)
    diff_num = _lib.empty(shape)
    cc = _lib.cy_consts()

    # ---- Fixed_Time_2pwl_SAM
    if isinstance(hard, holo.hardening.Fixed_Time_2PL_SAM):
        gmt_time = getattr(sam, "_gmt_time_device", None)
        gmt_time = gmt_time() if gmt_time is not None else sam._gmt_time
        # if `sam` is using galaxy merger rate (GMR), then `gmt_time` will be `None`
        if gmt_time is None:
            sam._log.info("`gmt_time` not calculated in SAM.  Setting to zeros.")
            gmt_time = np.zeros(sam.shape)
        gmt_time = _lib.to_dev(gmt_time)
        norm = hard._norm_device() if hasattr(hard, "_norm_device") else hard._norm
        norm = _lib.to_dev(norm)
        assert tuple(norm.shape) == (M, Q), f"hard._norm shape {tuple(norm.shape)} != {(M, Q)}"
        grid_age = _lib.to_dev(cosmo._grid_age)
        rc = lib.holo_dbn_2pwl(
            cc, _lib.ptr(fobs), F, float(hard._sepa_init), int(hard._num_steps), _lib.ptr(norm),
            float(hard._rchar), float(hard._gamma_inner), float(hard._gamma_outer),
            _lib.ptr(nden), _lib.ptr(mtot), _lib.ptr(mrat), _lib.ptr(redz), _lib.ptr(gmt_time), M, Q, Z,
            _lib.ptr(grid_z), _lib.ptr(grid_dcom), _lib.ptr(grid_age), grid_z.numel(),
            _lib.ptr(redz_final), _lib.ptr(diff_num), _lib.stream())
        _lib.check(rc, "dynamic_binary_number_at_fobs[2pwl]")

    # ---- Hard_GW
    elif isinstance(hard, holo.hardening.Hard_GW) or (isinstance(hard, type) and issubclass(hard, holo.hardening.Hard_GW)):
        redz_prime = getattr(sam, "_redz_prime_device", None)
        redz_prime = redz_prime() if redz_prime is not None else sam._redz_prime
        # if `sam` doesn't use a galaxy merger time (GMT), then `redz_prime` will be `None`,
        # set to initial redshift values instead
        if redz_prime is None:
            sam._log.info("`redz_prime` not calculated in SAM.  Setting to `redz` (initial) values.")
            redz_prime = np.asarray(sam.redz)[np.newaxis, np.newaxis, :] * np.ones(sam.shape)
        redz_prime = _lib.to_dev(redz_prime)
        rc = lib.holo_dbn_gw(
            cc, _lib.ptr(fobs), F, _lib.ptr(nden), _lib.ptr(mtot), _lib.ptr(mrat), _lib.ptr(redz),
            _lib.ptr(redz_prime), M, Q, Z, _lib.ptr(grid_z), _lib.ptr(grid_dcom), grid_z.numel(),
            _lib.ptr(redz_final), _lib.ptr(diff_num), _lib.stream())
        _lib.check(rc, "dynamic_binary_number_at_fobs[gw]")

    # ---- OTHER
    else:
        raise ValueError(f"Unexpected `hard` value {hard}!")

    return _out(redz_final, on_dev), _out(diff_num, on_dev)
