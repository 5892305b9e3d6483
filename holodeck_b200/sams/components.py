"""SAM components: galaxy stellar-mass function, pair fraction, merger time and merger rate.

Same classes, constructor arguments and call signatures as ``holodeck/sams/components.py``.
Instances remain callable on the host (numpy, closed forms); ``Semi_Analytic_Model`` does not call
them for the density grid -- it passes their parameters to the fused CUDA density kernel (K0), which
evaluates the same formulas on the device (``csrc/holo_math.cuh``).  ``_kernel_params()`` is the
hand-off.  User-defined subclasses that override ``__call__`` cannot be fused and are rejected by
``Semi_Analytic_Model.static_binary_density`` with ``NotImplementedError`` (no CPU fallback).
"""
import abc

import numpy as np

from holodeck_b200 import cosmo, utils, log
from holodeck_b200.constants import GYR, MSOL


# ----    Galaxy Stellar-Mass Function    ----

class _Galaxy_Stellar_Mass_Function(abc.ABC):
    """GSMF base class: number density of galaxies per log10 stellar-mass (components.py:34-107)."""

    @abc.abstractmethod
    def __init__(self, *args, **kwargs):
        return

    @abc.abstractmethod
    def __call__(self, mstar, redz):
        return

    def mbh_mass_func(self, mbh, redz, mmbulge, scatter=None):
        """GSMF -> MBH mass function via an M-Mbulge relation (components.py:62-107, no-scatter branch)."""
        if scatter in [None, True]:
            scatter = mmbulge._scatter_dex
        mstar = mmbulge.mstar_from_mbh(mbh, scatter=False)
        ndens = self(mstar, redz)
        dmstar_dmbh = mmbulge.dmstar_dmbh(mstar)
        ndens = ndens * (mbh/mstar) * dmstar_dmbh
        if scatter is not False and scatter > 0.0:
            raise NotImplementedError("scatter_redistribute_densities is outside the SAM-GWB hot path")
        return ndens


class GSMF_Schechter(_Galaxy_Stellar_Mass_Function):
    r"""Single Schechter function GSMF, $\Phi = dn / d\log_{10}(M)$ (components.py:110-172)."""

    def __init__(self, phi0=-2.77, phiz=-0.27, mchar0_log10=11.24, mchar0=None, mcharz=0.0, alpha0=-1.24, alphaz=-0.03):
        mchar0, _ = utils._parse_val_log10_val_pars(
            mchar0, mchar0_log10, val_units=MSOL, name='mchar0', only_one=True
        )
        self._phi0 = phi0
        self._phiz = phiz
        self._mchar0 = mchar0
        self._mcharz = mcharz
        self._alpha0 = alpha0
        self._alphaz = alphaz

    def __call__(self, mstar, redz):
        phi = self._phi_func(redz)
        mchar = self._mchar_func(redz)
        alpha = self._alpha_func(redz)
        xx = mstar / mchar
        return np.log(10.0) * phi * np.power(xx, 1.0 + alpha) * np.exp(-xx)

    def _phi_func(self, redz):
        return np.power(10.0, self._phi0 + self._phiz * redz)

    def _mchar_func(self, redz):
        return self._mchar0 + self._mcharz * redz

    def _alpha_func(self, redz):
        return self._alpha0 + self._alphaz * redz

    def _kernel_params(self):
        return 0, [self._phi0, self._phiz, self._mchar0, self._mcharz, self._alpha0, self._alphaz]


class _GSMF_Single_Schechter(_Galaxy_Stellar_Mass_Function):
    """Schechter function with quadratic-in-redshift parameters (components.py:175-273)."""

    def __init__(self, log10_phi_terms, log10_mstar_terms, alpha):
        self._log10_phi_terms = log10_phi_terms
        self._log10_mstar_terms = log10_mstar_terms
        self._alpha = alpha

    def __call__(self, mstar, redz):
        phi = self._phi_func(redz)
        mchar = self._mstar_func(redz)
        xx = mstar / mchar
        return np.log(10.0) * phi * np.power(xx, 1.0 + self._alpha) * np.exp(-xx)

    def _phi_func(self, redz):
        cc = self._log10_phi_terms
        return np.power(10.0, cc[0] + cc[1] * redz + cc[2] * redz**2)

    def _mstar_func(self, redz):
        cc = self._log10_mstar_terms
        return MSOL * np.power(10.0, cc[0] + cc[1] * redz + cc[2] * redz**2)


class GSMF_Double_Schechter(_Galaxy_Stellar_Mass_Function):
    """Sum of two Schechter functions, [Leja2020]_ defaults (components.py:276-329)."""

    def __init__(
        self,
        log10_phi1=[-2.383, -0.264, -0.107],
        log10_phi2=[-2.818, -0.368, +0.046],
        log10_mstar=[+10.767, +0.124, -0.033],
        alpha1=-0.28,
        alpha2=-1.48
    ):
        self._gsmf_one = _GSMF_Single_Schechter(log10_phi_terms=log10_phi1, log10_mstar_terms=log10_mstar, alpha=alpha1)
        self._gsmf_two = _GSMF_Single_Schechter(log10_phi_terms=log10_phi2, log10_mstar_terms=log10_mstar, alpha=alpha2)

    def __call__(self, mstar, redz):
        vals = self._gsmf_one(mstar, redz)
        vals = vals + self._gsmf_two(mstar, redz)
        return vals

    def _kernel_params(self):
        one, two = self._gsmf_one, self._gsmf_two
        return 1, [*one._log10_phi_terms, *two._log10_phi_terms, *one._log10_mstar_terms,
                   one._alpha, two._alpha, MSOL]


# ----    Galaxy Merger Rate    ----

class _Galaxy_Merger_Rate(abc.ABC):
    """components.py:335-377"""

    @abc.abstractmethod
    def __init__(self, *args, **kwargs):
        return

    @abc.abstractmethod
    def __call__(self, mass, mrat, redz):
        return


class GMR_Illustris(_Galaxy_Merger_Rate):
    """Galaxy merger rate from Illustris, [Rodriguez-Gomez2015]_ (components.py:380-482)."""

    def __init__(self, norm0_log10=None, normz=None, malpha0=None, malphaz=None, mdelta0=None, mdeltaz=None,
                 qgamma0=None, qgammaz=None, qgammam=None):
        if norm0_log10 is None:
            norm0_log10 = -2.2287
        if normz is None:
            normz = +2.4644
        if malpha0 is None:
            malpha0 = +0.2241
        if malphaz is None:
            malphaz = -1.1759
        if mdelta0 is None:
            mdelta0 = +0.7668
        if mdeltaz is None:
            mdeltaz = -0.4695
        if qgamma0 is None:
            qgamma0 = -1.2595
        if qgammaz is None:
            qgammaz = +0.0611
        if qgammam is None:
            qgammam = -0.0477
        self._norm0 = (10.0 ** norm0_log10) / GYR   # [1/sec]
        self._normz = normz
        self._malpha0 = malpha0
        self._malphaz = malphaz
        self._mdelta0 = mdelta0
        self._mdeltaz = mdeltaz
        self._qgamma0 = qgamma0
        self._qgammaz = qgammaz
        self._qgammam = qgammam
        self._mref_delta = 2.0e11 * MSOL
        self._mref = 1.0e10 * MSOL

    def _get_norm(self, redz):
        return self._norm0 * np.power(1.0 + redz, self._normz)

    def _get_malpha(self, redz):
        return self._malpha0 * np.power(1.0 + redz, self._malphaz)

    def _get_mdelta(self, redz):
        return self._mdelta0 * np.power(1.0 + redz, self._mdeltaz)

    def _get_qgamma(self, redz, mtot):
        qgamma = self._qgamma0 * np.power(1.0 + redz, self._qgammaz)
        return qgamma + self._qgammam * np.log10(mtot/self._mref)

    def __call__(self, mtot, mrat, redz):
        norm = self._get_norm(redz)
        malpha = self._get_malpha(redz)
        mdelta = self._get_mdelta(redz)
        qgamma = self._get_qgamma(redz, mtot)
        xx = (mtot/self._mref)
        mt = np.power(xx, malpha)
        yy = mtot/self._mref_delta
        mp1t = np.power(1.0 + yy, mdelta)
        qt = np.power(mrat, qgamma)
        return norm * mt * mp1t * qt

    def _kernel_params(self):
        return [self._norm0, self._normz, self._malpha0, self._malphaz, self._mdelta0, self._mdeltaz,
                self._qgamma0, self._qgammaz, self._qgammam, self._mref, self._mref_delta]


# ----    Galaxy Pair Fraction    ----

class _Galaxy_Pair_Fraction(abc.ABC):
    """components.py:488-517"""

    @abc.abstractmethod
    def __init__(self, *args, **kwargs):
        return

    @abc.abstractmethod
    def __call__(self, mass, mrat, redz):
        return


class GPF_Power_Law(_Galaxy_Pair_Fraction):
    """Power-law galaxy pair fraction, [Chen2019]_ Eq.6 (components.py:520-583)."""

    def __init__(self, frac_norm_allq=0.025, frac_norm=None, mref=None, mref_log10=11.0,
                 malpha=0.0, zbeta=0.8, qgamma=0.0, obs_conv_qlo=0.25, max_frac=1.0):
        mref, _ = utils._parse_val_log10_val_pars(
            mref, mref_log10, val_units=MSOL, name='mref', only_one=True
        )
        # If the pair-fraction integrated over all mass-ratios is given (f0), convert to regular (f0-prime)
        if frac_norm is None:
            if frac_norm_allq is None:
                raise ValueError("If `frac_norm` is not given, `frac_norm_allq` is requried!")
            pow = qgamma + 1.0
            qlo = obs_conv_qlo
            qhi = 1.00
            pair_norm = (qhi**pow - qlo**pow) / pow
            frac_norm = frac_norm_allq / pair_norm
        self._frac_norm = frac_norm
        self._malpha = malpha
        self._zbeta = zbeta
        self._qgamma = qgamma
        if (max_frac < 0.0) or (1.0 < max_frac):
            err = f"Given `max_frac`={max_frac:.4f} must be between [0.0, 1.0]!"
            log.exception(err)
            raise ValueError(err)
        self._max_frac = max_frac
        self._mref = mref

    def __call__(self, mass, mrat, redz):
        f0p = self._frac_norm
        am0 = self._mref
        rv = f0p * np.power(mass/am0, self._malpha) * np.power(1.0 + redz, self._zbeta) * np.power(mrat, self._qgamma)
        return np.clip(rv, None, self._max_frac)

    def _kernel_params(self):
        return [self._frac_norm, self._mref, self._malpha, self._zbeta, self._qgamma, self._max_frac]


# ----    Galaxy Merger Time    ----

class _Galaxy_Merger_Time(abc.ABC):
    """components.py:589-626"""

    @abc.abstractmethod
    def __init__(self, *args, **kwargs):
        return

    @abc.abstractmethod
    def __call__(self, mass, mrat, redz):
        return

    def zprime(self, mass, mrat, redz, **kwargs):
        """Redshift after the galaxy-merger time has elapsed; -1 where past z=0 (components.py:620-626)."""
        tau0 = self(mass, mrat, redz, **kwargs)
        redz_prime = utils.redz_after(tau0, redz=redz)
        return redz_prime, tau0


class GMT_Power_Law(_Galaxy_Merger_Time):
    """Power-law galaxy merger time, [Chen2019]_ Eq.18 (components.py:629-675)."""

    def __init__(self, time_norm=0.55*GYR, mref0=1.0e11*MSOL, malpha=0.0, zbeta=-0.5, qgamma=0.0):
        self._time_norm = time_norm
        self._malpha = malpha
        self._zbeta = zbeta
        self._qgamma = qgamma
        # NOTE: this is `b * M_0 = 0.4e11 Msol / h0` in [Chen2019]_
        self._mref = mref0 * (0.4 / cosmo.h)

    def __call__(self, mass, mrat, redz):
        tau0 = self._time_norm
        bm0 = self._mref
        return tau0 * np.power(mass/bm0, self._malpha) * np.power(1.0 + redz, self._zbeta) * np.power(mrat, self._qgamma)

    def _kernel_params(self):
        return [self._time_norm, self._mref, self._malpha, self._zbeta, self._qgamma]
