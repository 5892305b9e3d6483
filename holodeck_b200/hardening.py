"""Binary hardening models usable with the SAM GW-background path.

``sam.gwb`` / ``librarian.run_model`` accept exactly two models (``holodeck/sams/sam.py:910-916``):

* :class:`Hard_GW`             -- GW-only evolution            (``holodeck/hardening.py:87-212``)
* :class:`Fixed_Time_2PL_SAM`  -- phenomenological double power law with a fixed total lifetime
                                  (``holodeck/hardening.py:1371-1452``)

The other reference classes (CBD torques, stellar scattering, dynamical friction, ``Fixed_Time_2PL``)
drive the *discrete* populations and are out of scope.
"""
import abc

import numpy as np

import holodeck_b200 as holo
from holodeck_b200 import utils, _lib
from holodeck_b200.constants import GYR, PC


class _Hardening(abc.ABC):
    """Base class for binary-hardening models (hardening.py:58-80)."""

    CONSISTENT = None

    @abc.abstractmethod
    def dadt_dedt(self, evo, step, *args, **kwargs):
        pass

    def dadt(self, *args, **kwargs):
        rv_dadt, _dedt = self.dadt_dedt(*args, **kwargs)
        return rv_dadt

    def dedt(self, *args, **kwargs):
        _dadt, rv_dedt = self.dadt_dedt(*args, **kwargs)
        return rv_dedt


class Hard_GW(_Hardening):
    """Gravitational-wave driven binary hardening (hardening.py:87-212)."""

    CONSISTENT = False

    @staticmethod
    def dadt_dedt(evo, step):
        m1, m2 = evo.mass[:, step, :].T
        sepa = evo.sepa[:, step]
        eccen = evo.eccen[:, step] if (evo.eccen is not None) else None
        dadt = utils.gw_hardening_rate_dadt(m1, m2, sepa, eccen=eccen)
        dedt = None if eccen is None else utils.gw_dedt(m1, m2, sepa, eccen)
        return dadt, dedt

    @staticmethod
    def dadt(mtot, mrat, sepa, eccen=None):
        m1, m2 = utils.m1m2_from_mtmr(mtot, mrat)
        return utils.gw_hardening_rate_dadt(m1, m2, sepa, eccen=eccen)

    @staticmethod
    def dedt(mtot, mrat, sepa, eccen=None):
        if eccen is None:
            return np.zeros_like(mtot)
        m1, m2 = utils.m1m2_from_mtmr(mtot, mrat)
        return utils.gw_dedt(m1, m2, sepa, eccen=eccen)

    @staticmethod
    def deda(sepa, eccen):
        return 1.0 / utils.gw_dade(sepa, eccen)

    @property
    def consistent(self):
        return False


class Fixed_Time_2PL_SAM(_Hardening):
    """SAM-optimised double power-law hardening with a fixed total binary lifetime.

    Mirrors ``hardening.py:1371-1452``.  Construction solves, for every (mtot, mrat) grid edge, for
    the normalisation that makes the binary lifetime equal ``time`` -- one Brent root-find per pair
    with a 300-step trapezoid inside, done by the CUDA kernel K1a
    (``sam_cyutils.find_2pwl_hardening_norm``).  ``_norm`` is kept on the device and only copied to
    the host on first access.
    """

    CONSISTENT = True

    def __init__(self, sam, time, sepa_init=1.0e3*PC, rchar=10.0*PC, gamma_inner=-1.0, gamma_outer=+1.5, num_steps=300):
        from holodeck_b200.sams import sam_cyutils

        assert np.ndim(time) == 0
        assert np.ndim(rchar) == 0
        assert np.ndim(gamma_inner) == 0
        assert np.ndim(gamma_outer) == 0

        mtot, mrat = np.meshgrid(sam.mtot, sam.mrat, indexing='ij')
        shape = mtot.shape
        mt, mr = [mm.flatten() for mm in [mtot, mrat]]
        norm_log10 = sam_cyutils.find_2pwl_hardening_norm(
            time, mt, mr,
            sepa_init, rchar, gamma_inner, gamma_outer, num_steps, device=True,
        )
        # (M*Q,) ==> (M, Q)
        norm_log10 = norm_log10.reshape(shape)

        self._target_time = time
        self._norm_dev = 10.0 ** norm_log10
        self._norm_host = None
        self._num_steps = num_steps
        self._sepa_init = sepa_init
        self._rchar = rchar
        self._gamma_inner = gamma_inner
        self._gamma_outer = gamma_outer

    @property
    def _norm(self):
        """(M, Q) hardening-rate normalisation [cm/s] (numpy; ``hardening.py:1416``)."""
        if self._norm_host is None:
            self._norm_host = _lib.to_host(self._norm_dev)
        return self._norm_host

    def _norm_device(self):
        return self._norm_dev

    def __str__(self):
        return (
            f"{super().__str__()} :: "
            f"target_time/Gyr={self._target_time/GYR:.2e} num_steps={self._num_steps} "
            f"sepa_init/pc={self._sepa_init/PC:.2e} rchar/pc={self._rchar/PC:.2e} "
            f"gamma_inner={self._gamma_inner:.2e} gamma_outer={self._gamma_outer:.2e} "
        )

    def dadt_dedt(self, evo, step, *args, **kwargs):
        raise NotImplementedError()

    def dadt(self, mtot, mrat, sepa, norm=None):
        from holodeck_b200.sams import sam_cyutils
        if norm is None:
            norm = self._norm
        args = np.broadcast_arrays(mtot, mrat, sepa, norm)
        shape = args[0].shape
        mtot, mrat, sepa, norm = [aa.flatten() for aa in args]
        dadt_vals = sam_cyutils.hard_func_2pwl_gw(
            mtot, mrat, sepa, norm,
            self._rchar, self._gamma_inner, self._gamma_outer
        )
        return dadt_vals.reshape(shape)
