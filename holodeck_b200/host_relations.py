"""MBH--host scaling relations on the SAM path (subset of ``holodeck/host_relations.py``).

In scope: the power-law M-Mbulge relations ``MMBulge_Standard`` /
``MMBulge_KH2013`` / ``MMBulge_MM2013`` with a constant bulge fraction ``BF_Constant`` or the sigmoid
``BF_Sigmoid`` (used by the ``PS_Astro_Strong_*`` parameter spaces).  The density kernel (K0,
``csrc/holo_math.cuh``) evaluates them on the device -- closed forms, plus, for the sigmoid's numerically
inverted relations, the same quadratic splines scipy builds here, handed over as piecewise polynomials;
the numpy methods are the public callables the reference exposes, for host-side use.

Out of scope: ``MMBulge_Redshift*``, M-sigma and stellar-mass halo-mass relations -- not reachable from
the named configurations.
"""
import abc

import numpy as np

from holodeck_b200 import log, utils
from holodeck_b200.constants import MSOL


# ---- bulge fractions (host_relations.py:89-195)

class _Bulge_Frac(abc.ABC):

    def mbulge_from_mstar(self, mstar, redz=None, **kwargs):
        return mstar * self.bulge_frac(mstar=mstar, redz=redz, **kwargs)

    @abc.abstractmethod
    def bulge_frac(self, mstar=None, redz=None, mhalo=None, **kwargs):
        return

    @abc.abstractmethod
    def dmstar_dmbulge(self, mbulge=None, redz=None, mhalo=None, **kwargs):
        return

    def mstar_from_mbulge(self, mbulge, redz=None, **kwargs):
        raise NotImplementedError(f"``mstar_from_mbulge`` is not implemented in {self}!")


class BF_Constant(_Bulge_Frac):
    """Constant stellar-bulge mass fraction (host_relations.py:166-195)."""

    def __init__(self, bulge_frac=0.69):
        assert (0.0 < bulge_frac) and (bulge_frac <= 1.0)
        self._bulge_mass_frac = bulge_frac

    def bulge_frac(self, *args, **kwargs):
        return self._bulge_mass_frac

    def mstar_from_mbulge(self, mbulge, redz=None, **kwargs):
        return mbulge / self.bulge_frac()

    def dmstar_dmbulge(self, mbulge, redz=None, **kwargs):
        return 1.0 / self.bulge_frac()


class BF_Sigmoid(_Bulge_Frac):
    r"""Stellar-bulge mass fraction rising from ``bulge_frac_lo`` to ``bulge_frac_hi`` with stellar mass
    (``host_relations.py:198-331``):

    .. math:: f_b(m < m_c) = f_l + (f_h - f_l) / [1 + ((m / m_c)^{-1} - 1)^k],  \qquad  f_b(m \ge m_c) = f_h.

    The inverse relations (stellar mass and its derivative as functions of BULGE mass) have no closed form; like
    the reference they are quadratic ``scipy.interpolate.interp1d`` interpolants over a 210-point grid that is
    dense around the transition.  ``_kernel_tables`` exports the two splines as piecewise polynomials for K0.
    """

    _INTERP_GRID_SIZE = 200
    _DERIV_DELTA = 1.0e-6

    def __init__(self, bulge_frac_lo=0.5, bulge_frac_hi=1.0, mstar_char_log10=11.0, width_dex=1.0):
        import scipy.interpolate
        assert (0.0 < bulge_frac_lo) and (bulge_frac_lo <= 1.0), f"{bulge_frac_lo=} must be in (0.0, 1.0]!"
        assert (0.0 < bulge_frac_hi) and (bulge_frac_hi <= 1.0), f"{bulge_frac_hi=} must be in (0.0, 1.0]!"
        assert (bulge_frac_lo < bulge_frac_hi), f"{bulge_frac_lo=} must be less than {bulge_frac_hi=} !"
        assert not (width_dex < 0.0), f"{width_dex=} must be non-negative!"
        self._bulge_frac_lo, self._bulge_frac_hi = bulge_frac_lo, bulge_frac_hi
        self._mstar_char = (10.0 ** mstar_char_log10) * MSOL
        self._width_dex = width_dex
        # stellar-mass grid: half of the points in [x_c - w/2, x_c + 1/2] dex, half over the ten decades below,
        # ten more over the ten decades above (host_relations.py:251-263)
        xc = np.log10(self._mstar_char)
        lo, hi = xc - 0.5 * width_dex, xc + 0.5
        half = self._INTERP_GRID_SIZE // 2
        ms = np.concatenate([np.logspace(lo - 10.0, lo, half, endpoint=False), np.logspace(lo, hi, half, endpoint=False),
                             np.logspace(hi, hi + 10.0, 10)])
        mb = self.mbulge_from_mstar(ms)
        dd = self._DERIV_DELTA
        ms_lo, ms_hi = ms * (1.0 - dd / 2.0), ms * (1.0 + dd / 2.0)
        dms_dmb = (ms_hi - ms_lo) / (self.mbulge_from_mstar(ms_hi) - self.mbulge_from_mstar(ms_lo))     # central difference
        self._interp_mstar_from_mbulge = scipy.interpolate.interp1d(mb, ms, kind='quadratic', fill_value='extrapolate')
        self._interp_dmstar_dmbulge_from_mbulge = scipy.interpolate.interp1d(mb, dms_dmb, kind='quadratic',
                                                                             fill_value='extrapolate')

    def bulge_frac(self, mstar, redz=None, **kwargs):
        mm = np.minimum(np.asarray(mstar, dtype=float) / self._mstar_char, 1.0)
        flo, fhi = self._bulge_frac_lo, self._bulge_frac_hi
        frac = flo + (fhi - flo) / (1.0 + ((1.0 / mm) - 1.0) ** self._width_dex)
        return np.where((mm >= 1.0) | (frac > fhi), fhi, frac)

    def mstar_from_mbulge(self, mbulge, redz=None, **kwargs):
        mbulge = np.asarray(mbulge, dtype=float)
        mstar = mbulge / self._bulge_frac_hi                    # right wherever this lands above the characteristic mass
        below = (mstar / self._mstar_char) < 1.0
        return np.where(below, self._interp_mstar_from_mbulge(mbulge), mstar)

    def dmstar_dmbulge(self, mbulge, redz=None, **kwargs):
        mbulge = np.asarray(mbulge, dtype=float)
        below = (mbulge / self._mstar_char) < self._bulge_frac_hi
        return np.where(below, self._interp_dmstar_dmbulge_from_mbulge(mbulge), 1.0 / self._bulge_frac_hi)

    def _kernel_tables(self):
        """The two interpolants as piecewise quadratics for the density kernel: ``(breaks (n+1,), coef (3, n))`` each,
        value = c0 dx^2 + c1 dx + c2 with dx = x - breaks[i]; first / last piece extrapolate, as interp1d does."""
        out = []
        for itp in (self._interp_mstar_from_mbulge, self._interp_dmstar_dmbulge_from_mbulge):
            spl = itp._spline                                   # scipy BSpline, k = 2
            knots = np.unique(spl.t)                            # piece boundaries (the end knots are repeated)
            left = knots[:-1]
            # Taylor coefficients of each quadratic piece at its left end (evaluated just inside the piece)
            mid = 0.5 * (left + knots[1:])
            d2 = np.ravel(spl(mid, nu=2))
            d1 = np.ravel(spl(mid, nu=1)) - d2 * (mid - left)
            d0 = np.ravel(spl(mid)) - d1 * (mid - left) - 0.5 * d2 * (mid - left) ** 2
            out.append((np.ascontiguousarray(knots), np.ascontiguousarray(np.stack([0.5 * d2, d1, d0]))))
        return out


# ---- M-Mbulge relations (host_relations.py:334-799)

class _BH_Host_Relation(abc.ABC):
    pass


def _add_scatter(vals, eps):
    """host_relations.py:1181-1206"""
    if (eps is None) or (eps is False) or (eps == 0.0):
        return vals
    return vals + np.random.normal(0.0, eps, size=np.shape(vals))


def _log10_relation(xx, amp, plaw, eps_dex, x0=1.0):
    """y = amp * (x/x0)^plaw with optional log-normal scatter (host_relations.py:1102-1134)."""
    yy = np.log10(xx/x0) * plaw
    yy = _add_scatter(yy, eps_dex)
    return amp * np.power(10.0, yy)


def _log10_relation_reverse(yy, amp, plaw, eps_dex, x0=1.0):
    """Inverse of :func:`_log10_relation` (host_relations.py:1137-1178)."""
    xx = np.log10(yy/amp)
    xx = _add_scatter(xx, eps_dex)
    xx = (1.0/plaw) * xx
    return x0 * np.power(10.0, xx)


class _MMBulge_Relation(_BH_Host_Relation):
    """Base class of Mbh-Mbulge relations (host_relations.py:410-571)."""

    def __init__(self, bulge_frac=None, bulge_mfrac=None):
        if bulge_mfrac is not None:
            log.warning("Parameter ``bulge_mfrac`` is deprecated!  Please use ``bulge_frac`` instead!")
            if bulge_frac is not None:
                err = "Cannot provide both a ``bulge_mfrac`` and ``bulge_frac``!"
                log.exception(err)
                raise ValueError(err)
            bulge_frac = BF_Constant(bulge_mfrac)
        if bulge_frac is None:
            bulge_frac = BF_Constant(self.BULGE_MASS_FRAC)
        self._bulge_frac = bulge_frac

    def dmstar_dmbh(self, mstar, redz=None, **bfkwargs):
        mbulge = self._bulge_frac.mbulge_from_mstar(mstar, redz=redz)
        dmstar_dmbulge = self._bulge_frac.dmstar_dmbulge(mbulge, redz=redz, **bfkwargs)
        dmbulge_dmbh = self.dmbulge_dmbh(mbulge, redz=redz)
        return dmstar_dmbulge * dmbulge_dmbh

    def mbh_from_mstar(self, mstar, redz=None, scatter=None):
        mbulge = self._bulge_frac.mbulge_from_mstar(mstar, redz=redz)
        return self.mbh_from_mbulge(mbulge, redz=redz, scatter=scatter)

    @abc.abstractmethod
    def mbh_from_mbulge(self, mbulge, redz=None, scatter=None, **kwargs):
        return

    @abc.abstractmethod
    def dmbulge_dmbh(self, mbulge, redz=None):
        return

    def mbulge_from_mbh(self, *args, **kwargs):
        raise NotImplementedError(f"``mbulge_from_mbh`` has not been implemented in {self}!")

    def mstar_from_mbh(self, mbh, redz=None):
        mbulge = self.mbulge_from_mbh(mbh)
        return self._bulge_frac.mstar_from_mbulge(mbulge, redz=redz)


class MMBulge_Standard(_MMBulge_Relation):
    """Simple power-law Mbh-Mbulge relation (host_relations.py:624-771)."""

    MASS_AMP_LOG10 = 8.17
    MASS_PLAW = 1.01
    MASS_REF = 1.0e11 * MSOL
    SCATTER_DEX = 0.3
    BULGE_MASS_FRAC = 0.615

    def __init__(self, mamp_log10=None, mplaw=None, mref=None, scatter_dex=None, bulge_frac=None,
                 bulge_mfrac=None, **kwargs):
        if 'mamp' in kwargs:
            log.warning("The `mamp` parameter has been deprecated!  Use `mamp_log10`!")
            if mamp_log10 is not None:
                err = "Both `mamp` (deprecated!) and `mamp_log10` have been given!  Cannot correct."
                log.exception(err)
                raise ValueError(err)
            mamp_log10 = np.log10(kwargs.pop('mamp') / MSOL)
        super().__init__(bulge_frac=bulge_frac, bulge_mfrac=bulge_mfrac)
        if mamp_log10 is None:
            mamp_log10 = self.MASS_AMP_LOG10
        mamp = MSOL * np.power(10.0, mamp_log10)
        if mplaw is None:
            mplaw = self.MASS_PLAW
        if mref is None:
            mref = self.MASS_REF
        if scatter_dex is None:
            scatter_dex = self.SCATTER_DEX
        self._mamp = mamp
        self._mplaw = mplaw
        self._mref = mref
        self._scatter_dex = scatter_dex
        if len(kwargs) > 0:
            log.warning(f"Unused parameters passed to {self}!  kwargs={kwargs}")

    def mbh_from_mbulge(self, mbulge, redz=None, scatter=None):
        scatter_dex = self._scatter_dex if scatter else None
        return _log10_relation(mbulge, self._mamp, self._mplaw, scatter_dex, x0=self._mref)

    def dmbulge_dmbh(self, mbulge, redz=None, **bfkwargs):
        plaw = self._mplaw
        mbh = self.mbh_from_mbulge(mbulge, redz=redz, scatter=False)
        return mbulge / (plaw * mbh)

    def mbulge_from_mbh(self, mbh, redz=None, scatter=None):
        scatter_dex = self._scatter_dex if scatter else None
        return _log10_relation_reverse(mbh, self._mamp, self._mplaw, scatter_dex, x0=self._mref)

    def mstar_from_mbh(self, mbh, redz=None, scatter=None, **kwargs):
        mbulge = self.mbulge_from_mbh(mbh, redz=redz, scatter=scatter)
        return self._bulge_frac.mstar_from_mbulge(mbulge, redz=redz, **kwargs)


class MMBulge_KH2013(MMBulge_Standard):
    """[KH2013]_ Eq.10 (host_relations.py:774-784)."""
    MASS_AMP_LOG10 = 8.69
    MASS_REF = MSOL * 1e11
    MASS_PLAW = 1.17
    SCATTER_DEX = 0.28


class MMBulge_MM2013(MMBulge_Standard):
    """[MM2013]_ (host_relations.py:787-799)."""
    MASS_AMP_LOG10 = 8.46
    MASS_REF = MSOL * 1e11
    MASS_PLAW = 1.05
    SCATTER_DEX = 0.34


def get_mmbulge_relation(mmbulge=None):
    """host_relations.py:888-907"""
    return utils.get_subclass_instance(mmbulge, MMBulge_KH2013, _MMBulge_Relation)
