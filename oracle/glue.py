"""CPU oracle for the SAM GW-background hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  Nothing under ``holodeck_b200/`` does.

Two layers:

1. **The reference itself** for everything the reference implements natively: ``oracle/_ref`` holds
   the reference's own ``holodeck/cyutils.pyx`` and ``holodeck/sams/sam_cyutils.pyx`` compiled
   in-container by ``oracle/build_ref.py`` (``ref()`` below imports them).  Parity for K1/K2/K3/K4/K5
   is therefore pinned on outputs of the reference run here.
2. **A numpy restatement of the reference's Python glue** (this file), because the reference's
   Python layer cannot be imported here (it needs cosmopy / astropy / kalepy / h5py, none present):
   each function cites the reference file:line it follows.  Third-party arithmetic that is absent
   from ``/root/reference`` -- ``cosmopy`` (unpinned, ``requirements.txt:2``) and astropy's
   ``comoving_distance`` / ``age`` -- is restated from the published flat-LambdaCDM model with
   *independent* numerics (scipy adaptive quadrature, not the closed forms the product uses).
   No reference test pins a cosmology value, so that part is **parity unpinned** (DESIGN.md).

The golden vectors the reference's own tests hold for this path (``holodeck/tests/test_utils.py:52-94``
and ``tests/test_host_relations__mmbulge.py``) are checked against this file in
``tests/test_oracle_golden.py``.
"""
import os
import sys
from pathlib import Path

import numpy as np
import scipy as sp
import scipy.integrate   # noqa
import scipy.interpolate   # noqa
import scipy.stats   # noqa

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"

# ---- constants: astropy CODATA-2018 values, holodeck/constants.py:23-57 --------------------------
NWTG = 6.6743e-08
SPLC = 29979245800.0
MSOL = 1.988409870698051e+33
PC = 3.0856775814913674e+18
YR = 31557600.0
GYR = 1.0e9 * YR
MPC = 1.0e6 * PC
SCHW = 2 * NWTG / (SPLC * SPLC)
# utils.py:39-41
_GW_SRC_CONST = 8 * np.power(NWTG, 5/3) * np.power(np.pi, 2/3) / np.sqrt(10) / np.power(SPLC, 4)
_GW_DADT_SEP_CONST = - 64 * np.power(NWTG, 3) / 5 / np.power(SPLC, 5)
_GW_DEDT_ECC_CONST = - 304 * np.power(NWTG, 3) / 15 / np.power(SPLC, 5)
_GW_LUM_CONST = (32.0 / 5.0) * np.power(NWTG, 7.0/3.0) * np.power(SPLC, -5.0)


def ref():
    """Import the compiled reference (``oracle/_ref``): returns ``(cyutils, sam_cyutils, holodeck_stub)``."""
    if not (REF_DIR / "holodeck" / "__init__.py").exists():
        raise RuntimeError("oracle/_ref is not built: run `python oracle/build_ref.py`")
    # the stub package is named `holodeck` (the compiled Cython hard-codes that import); keep it from
    # shadowing anything else by inserting the path only for the import
    prev = sys.modules.get("holodeck")
    if prev is not None and not str(getattr(prev, "__file__", "")).startswith(str(REF_DIR)):
        raise RuntimeError("a different `holodeck` package is already imported")
    sys.path.insert(0, str(REF_DIR))
    try:
        import holodeck
        import holodeck.cyutils as cyutils
        import holodeck.sams.sam_cyutils as sam_cyutils
    finally:
        sys.path.remove(str(REF_DIR))
    return cyutils, sam_cyutils, holodeck


# ==================================================================================================
# Cosmology: flat LambdaCDM, no radiation (holodeck/__init__.py:48-85 -> cosmopy.Cosmology(h=0.6933,
# Om0=0.288, Ob0=0.0472)); independent numerics (adaptive quadrature + root finding)
# ==================================================================================================

class OracleCosmo:
    """``closed_form=False`` (default): every quantity by adaptive quadrature / Newton inversion --
    slow, fully independent of the product's formulas; used for the golden fixtures.
    ``closed_form=True``: the textbook flat-LCDM closed forms for ``age`` / ``tage_to_z`` and a dense
    spline for ``comoving_distance`` -- what the CPU-baseline chain uses at the 91x81x101 grid
    (checked against the quadrature versions in tests/test_oracle.py)."""

    def __init__(self, h=0.6933, Om0=0.2880, closed_form=False):
        self.h = h
        self.Om0 = Om0
        self.H0_cgs = 100.0 * h * 1.0e5 / MPC
        self.hubble_time = 1.0 / self.H0_cgs
        self.hubble_distance = SPLC / self.H0_cgs
        if closed_form:
            ol = 1.0 - Om0
            self.age = lambda zz: (2.0 / 3.0) * self.hubble_time / np.sqrt(ol) * np.arcsinh(
                np.sqrt(ol / Om0) * np.power(1.0 + np.asarray(zz, dtype=float), -1.5))
            self.tage_to_z = lambda tt: np.power(
                np.sqrt(ol / Om0) / np.sinh(1.5 * np.sqrt(ol) * np.asarray(tt, dtype=float) / self.hubble_time), 2.0 / 3.0) - 1.0
            self.comoving_distance = self.comoving_distance_fast

    def efunc(self, zz):
        return np.sqrt(self.Om0 * (1.0 + zz)**3 + (1.0 - self.Om0))

    def dtdz(self, zz):
        return self.hubble_time / ((1.0 + zz) * self.efunc(zz))

    def _age_scalar(self, zz):
        # t(z) = int_z^inf dz' / ((1+z') H(z')) ; substitute a = 1/(1+z)
        val, _ = sp.integrate.quad(
            lambda aa: 1.0 / (aa * np.sqrt(self.Om0 / aa**3 + 1.0 - self.Om0)), 0.0, 1.0 / (1.0 + zz),
            epsabs=0.0, epsrel=1e-13, limit=400)
        return self.hubble_time * val

    def age(self, zz):
        return np.vectorize(self._age_scalar)(np.asarray(zz, dtype=float))

    def _dcom_scalar(self, zz):
        val, _ = sp.integrate.quad(lambda xx: 1.0 / self.efunc(xx), 0.0, zz, epsabs=0.0, epsrel=1e-13, limit=400)
        return self.hubble_distance * val

    def comoving_distance(self, zz):
        """[cm]; astropy `cosmo.comoving_distance(z).cgs.value` (gravwaves.py:718)."""
        return np.vectorize(self._dcom_scalar)(np.asarray(zz, dtype=float))

    def tage_to_z(self, age):
        """Invert age(z) (cosmopy `tage_to_z`, utils.py:1799,1806): Newton iterations on the
        quadrature-based `age`, started from a coarse table."""
        age = np.asarray(age, dtype=float)
        flat = age.reshape(-1)
        ztab = np.concatenate([[-0.9], -np.logspace(-1, -6, 11), [0.0], np.logspace(-6, 4, 400)])
        ttab = self.age(ztab)
        zz = np.interp(flat, ttab[::-1], ztab[::-1])
        for _ in range(12):
            resid = self.age(zz) - flat
            step = resid / self.dtdz(zz)        # dt/dz = -dtdz  =>  z_new = z + resid/dtdz
            zz = zz + step
            if np.all(np.abs(step) <= 2e-15 * (1.0 + np.abs(zz))):
                break
        return zz.reshape(age.shape)

    def comoving_distance_fast(self, zz, npts=20001, zmax=None):
        """Dense-table + cubic-spline version for large arrays (error << 1e-11, checked in tests)."""
        zz = np.asarray(zz, dtype=float)
        if zz.size == 0:
            return np.zeros_like(zz)
        zmax = max(float(np.max(zz)), 1e-3) if zmax is None else zmax
        xg = np.linspace(0.0, np.log1p(zmax), npts)
        zg = np.expm1(xg)
        # cumulative integral in x = ln(1+z): d_c = D_H int (1+z)/E dx, integrand smooth -> Simpson-exact spline
        fg = (1.0 + zg) / self.efunc(zg)
        spl = sp.interpolate.CubicSpline(xg, fg)
        cum = spl.antiderivative()
        return self.hubble_distance * (cum(np.log1p(zz)) - cum(0.0))


# ==================================================================================================
# utils.py restatements
# ==================================================================================================

def midpoints(vals, axis=-1):
    """utils.midpoints utils.py:703-711 (kalepy.utils.midpoints at the call sites, linear)."""
    mm = np.moveaxis(vals, axis, 0)
    mm = 0.5 * (mm[1:] + mm[:-1])
    return np.moveaxis(mm, 0, axis)


def pta_freqs(dur=16.03*YR, num=40):
    """utils.pta_freqs utils.py:835-874."""
    fmin = 1.0 / dur
    cents = np.arange(1, num+2) * fmin
    edges = cents - fmin / 2.0
    cents = cents[:-1]
    return cents, edges


def m1m2_from_mtmr(mt, mr):
    """utils.py:1620-1642"""
    mt = np.asarray(mt)
    mr = np.asarray(mr)
    m1 = mt / (1.0 + mr)
    m2 = mt - m1
    return np.array([m1, m2])


def chirp_mass(m1, m2):
    """utils.py:1951-1975"""
    return np.power(m1 * m2, 3.0/5.0)/np.power(m1 + m2, 1.0/5.0)


def chirp_mass_mtmr(mt, mr):
    """utils.py:1978-1997"""
    return mt * np.power(mr, 3.0/5.0) / np.power(1 + mr, 6.0/5.0)


def gw_strain_source(mchirp, dcom, freq_rest_orb):
    """utils.py:2260-2285"""
    return _GW_SRC_CONST * mchirp * np.power(2*mchirp*freq_rest_orb, 2/3) / dcom


def _gw_ecc_func(eccen):
    """utils.py:2421-2441"""
    e2 = eccen*eccen
    return (1 + (73/24)*e2 + (37/96)*e2*e2) / np.power(1 - e2, 7/2)


def gw_hardening_rate_dadt(m1, m2, sepa, eccen=None):
    """utils.py:2153-2183"""
    dadt = _GW_DADT_SEP_CONST * m1 * m2 * (m1 + m2) / np.power(sepa, 3)
    if eccen is not None:
        dadt = dadt * _gw_ecc_func(eccen)
    return dadt


def gw_dedt(m1, m2, sepa, eccen):
    """utils.py:2045-2074"""
    cc = _GW_DEDT_ECC_CONST
    e2 = eccen**2
    dedt = cc * m1 * m2 * (m1 + m2) / np.power(sepa, 4)
    dedt = dedt * (1.0 + e2*121.0/304.0) * eccen / np.power(1 - e2, 5.0/2.0)
    return dedt


def gw_dade(sepa, eccen):
    """utils.py:2077-2102"""
    e2 = eccen**2
    num = (1 + (73.0/24.0)*e2 + (37.0/96.0)*e2*e2)
    den = (1 - e2) * (1.0 + (121.0/304.0)*e2)
    return (12.0 / 19.0) * (sepa / eccen) * (num / den)


def gw_lum_circ(mchirp, freq_orb_rest):
    """utils.py:2225-2257"""
    return _GW_LUM_CONST * np.power(2.0*np.pi*freq_orb_rest*mchirp, 10.0/3.0)


def dfdt_from_dadt(dadt, sepa, mtot=None, frst_orb=None):
    """utils.py:1554-1587"""
    if frst_orb is None:
        frst_orb = kepler_freq_from_sepa(mtot, sepa)
    dfdt = -1.5 * (frst_orb / sepa) * dadt
    return dfdt, frst_orb


def gw_hardening_rate_dfdt(m1, m2, frst_orb, eccen=None):
    """utils.py:2186-2211"""
    m1, m2, frst_orb = [np.asarray(vv) for vv in (m1, m2, frst_orb)]
    sepa = kepler_sepa_from_freq(m1+m2, frst_orb)
    dfdt = gw_hardening_rate_dadt(m1, m2, sepa, eccen=None if eccen is None else np.asarray(eccen))
    dfdt, _ = dfdt_from_dadt(dfdt, sepa, frst_orb=frst_orb)
    return dfdt, frst_orb


def gw_hardening_timescale_freq(mchirp, frst):
    """utils.py:2214-2234"""
    mchirp, frst = np.asarray(mchirp), np.asarray(frst)
    return (5.0 / 96.0) * np.power(NWTG*mchirp/SPLC**3, -5.0/3.0) * np.power(2*np.pi*frst, -8.0/3.0)


def kepler_freq_from_sepa(mass, sepa):
    """utils.py:1685-1702"""
    return (1.0/(2.0*np.pi))*np.sqrt(NWTG*mass)/np.power(sepa, 1.5)


def kepler_sepa_from_freq(mass, freq):
    """utils.py:1705-1724"""
    return np.power(NWTG*mass/np.square(2.0*np.pi*freq), 1.0/3.0)


def rk4_step(func, x0, y0, dx):
    """utils.py:1018-1043"""
    k1 = dx * func(x0, y0)
    k2 = dx * func(x0 + dx/2.0, y0 + k1/2.0)
    k3 = dx * func(x0 + dx/2.0, y0 + k2/2.0)
    k4 = dx * func(x0 + dx, y0 + k3)
    y1 = y0 + (1.0/6.0) * (k1 + 2*k2 + 2*k3 + k4)
    return x0 + dx, y1


# ==================================================================================================
# SAM density: sams/sam.py:250-398 with sams/components.py and host_relations.py callables
# ==================================================================================================

PS_CLASSIC_DEFAULTS = dict(   # librarian/param_spaces_classic.py:13-42
    hard_time=3.0, hard_sepa_init=1e4, hard_rchar=100.0, hard_gamma_inner=-1.0, hard_gamma_outer=+2.5,
    gsmf_phi0_log10=-2.77, gsmf_phiz=-0.6, gsmf_mchar0_log10=11.24, gsmf_mcharz=0.11,
    gsmf_alpha0=-1.21, gsmf_alphaz=-0.03,
    gpf_frac_norm_allq=0.025, gpf_malpha=0.0, gpf_qgamma=0.0, gpf_zbeta=1.0, gpf_max_frac=1.0,
    gmt_norm=0.5, gmt_malpha=0.0, gmt_qgamma=-1.0, gmt_zbeta=-0.5,
    mmb_mamp_log10=8.69, mmb_plaw=1.10, mmb_scatter_dex=0.3,
)


def gsmf_schechter(mstar, redz, phi0=-2.77, phiz=-0.27, mchar0_log10=11.24, mcharz=0.0, alpha0=-1.24, alphaz=-0.03):
    """GSMF_Schechter components.py:110-172"""
    mchar0 = MSOL * np.power(10.0, mchar0_log10)       # utils._parse_val_log10_val_pars utils.py:1299
    phi = np.power(10.0, phi0 + phiz * redz)
    mchar = mchar0 + mcharz * redz
    alpha = alpha0 + alphaz * redz
    xx = mstar / mchar
    return np.log(10.0) * phi * np.power(xx, 1.0 + alpha) * np.exp(-xx)


def gsmf_double_schechter(mstar, redz, log10_phi1=(-2.383, -0.264, -0.107), log10_phi2=(-2.818, -0.368, +0.046),
                          log10_mstar=(+10.767, +0.124, -0.033), alpha1=-0.28, alpha2=-1.48):
    """GSMF_Double_Schechter components.py:276-329 (+ _GSMF_Single_Schechter :175-270)"""
    def single(cc_phi, alpha):
        phi = np.power(10.0, cc_phi[0] + cc_phi[1] * redz + cc_phi[2] * redz**2)
        mchar = MSOL * np.power(10.0, log10_mstar[0] + log10_mstar[1] * redz + log10_mstar[2] * redz**2)
        xx = mstar / mchar
        return np.log(10.0) * phi * np.power(xx, 1.0 + alpha) * np.exp(-xx)
    vals = single(log10_phi1, alpha1)
    vals += single(log10_phi2, alpha2)
    return vals


def gpf_power_law(mass, mrat, redz, frac_norm_allq=0.025, frac_norm=None, mref_log10=11.0, malpha=0.0, zbeta=0.8,
                  qgamma=0.0, obs_conv_qlo=0.25, max_frac=1.0):
    """GPF_Power_Law components.py:520-583"""
    mref = MSOL * np.power(10.0, mref_log10)
    if frac_norm is None:
        pow = qgamma + 1.0
        qlo = obs_conv_qlo
        qhi = 1.00
        pair_norm = (qhi**pow - qlo**pow) / pow
        frac_norm = frac_norm_allq / pair_norm
    rv = frac_norm * np.power(mass/mref, malpha) * np.power(1.0 + redz, zbeta) * np.power(mrat, qgamma)
    return np.clip(rv, None, max_frac)


def gmt_power_law(mass, mrat, redz, cosmo_h, time_norm=0.55*GYR, mref0=1.0e11*MSOL, malpha=0.0, zbeta=-0.5, qgamma=0.0):
    """GMT_Power_Law components.py:629-675"""
    mref = mref0 * (0.4 / cosmo_h)
    return time_norm * np.power(mass/mref, malpha) * np.power(1.0 + redz, zbeta) * np.power(mrat, qgamma)


GMR_ILLUSTRIS_DEFAULTS = dict(norm0_log10=-2.2287, normz=+2.4644, malpha0=+0.2241, malphaz=-1.1759, mdelta0=+0.7668,
                              mdeltaz=-0.4695, qgamma0=-1.2595, qgammaz=+0.0611, qgammam=-0.0477)


def gmr_illustris(mtot, mrat, redz, **kw):
    """GMR_Illustris components.py:380-482"""
    pp = dict(GMR_ILLUSTRIS_DEFAULTS)
    pp.update(kw)
    norm0 = (10.0 ** pp['norm0_log10']) / GYR
    mref_delta = 2.0e11 * MSOL
    mref = 1.0e10 * MSOL
    norm = norm0 * np.power(1.0 + redz, pp['normz'])
    malpha = pp['malpha0'] * np.power(1.0 + redz, pp['malphaz'])
    mdelta = pp['mdelta0'] * np.power(1.0 + redz, pp['mdeltaz'])
    qgamma = pp['qgamma0'] * np.power(1.0 + redz, pp['qgammaz'])
    qgamma = qgamma + pp['qgammam'] * np.log10(mtot/mref)
    xx = (mtot/mref)
    mt = np.power(xx, malpha)
    yy = mtot/mref_delta
    mp1t = np.power(1.0 + yy, mdelta)
    qt = np.power(mrat, qgamma)
    return norm * mt * mp1t * qt


class BFSigmoid:
    """BF_Sigmoid host_relations.py:198-331: sigmoid bulge fraction with scipy-interp1d inverses (the reference's own
    construction: grid :251-263, central differences :266-274, quadratic interpolants :276-283)."""
    _INTERP_GRID_SIZE = 200
    _DERIV_DELTA = 1.0e-6

    def __init__(self, bulge_frac_lo=0.5, bulge_frac_hi=1.0, mstar_char_log10=11.0, width_dex=1.0):
        self._bulge_frac_lo = bulge_frac_lo
        self._bulge_frac_hi = bulge_frac_hi
        self._mstar_char = (10.0 ** mstar_char_log10) * MSOL
        self._width_dex = width_dex
        mc = self._mstar_char
        xx = np.log10(mc)
        xbreak1 = xx - 0.5*width_dex
        xbreak2 = xx + 0.5
        _ms_lo = np.logspace(xbreak1 - 10.0, xbreak1, self._INTERP_GRID_SIZE//2, endpoint=False)
        ms = np.logspace(xbreak1, xbreak2, self._INTERP_GRID_SIZE//2, endpoint=False)
        _ms_hi = np.logspace(xbreak2, xbreak2 + 10.0, 10)
        ms = np.concatenate([_ms_lo, ms, _ms_hi])
        mb = self.mbulge_from_mstar(ms)
        dd = self._DERIV_DELTA
        ms_lo = ms * (1.0 - dd/2.0)
        ms_hi = ms * (1.0 + dd/2.0)
        mb_lo = self.mbulge_from_mstar(ms_lo)
        mb_hi = self.mbulge_from_mstar(ms_hi)
        dms_dmb = (ms_hi - ms_lo) / (mb_hi - mb_lo)
        self._interp_mstar_from_mbulge = sp.interpolate.interp1d(mb, ms, kind='quadratic', fill_value='extrapolate')
        self._interp_dmstar_dmbulge_from_mbulge = sp.interpolate.interp1d(mb, dms_dmb, kind='quadratic', fill_value='extrapolate')

    def bulge_frac(self, mstar):
        """host_relations.py:286-295"""
        mm = mstar / self._mstar_char
        steep = self._width_dex
        flo = self._bulge_frac_lo
        fhi = self._bulge_frac_hi
        mm[mm > 1.0] = 1.0
        frac = flo + (fhi - flo) / (1.0 + ((1.0 / mm) - 1.0)**steep)
        frac[(mm >= 1.0) | (frac > fhi)] = fhi
        return frac

    def mbulge_from_mstar(self, mstar):
        """_Bulge_Frac.mbulge_from_mstar host_relations.py:113-131"""
        return mstar * self.bulge_frac(np.array(mstar, dtype=float))

    def mstar_from_mbulge(self, mbulge):
        """host_relations.py:297-308"""
        fhi = self._bulge_frac_hi
        mstar = np.ones_like(mbulge) * mbulge / fhi
        sel = (mstar/self._mstar_char) < 1.0
        mstar[sel] = self._interp_mstar_from_mbulge(mbulge[sel])
        return mstar

    def dmstar_dmbulge(self, mbulge):
        """host_relations.py:310-321"""
        fhi = self._bulge_frac_hi
        dms_dmb = np.ones_like(mbulge) / fhi
        sel = (mbulge/self._mstar_char) < fhi
        dms_dmb[sel] = self._interp_dmstar_dmbulge_from_mbulge(mbulge[sel])
        return dms_dmb


class MMBulge:
    """MMBulge_Standard / KH2013 / MM2013 with BF_Constant (or, `bulge_frac=BFSigmoid(...)`, BF_Sigmoid):
    host_relations.py:624-799, 166-195, 198-331."""
    KINDS = {   # MASS_AMP_LOG10, MASS_PLAW, SCATTER_DEX, BULGE_MASS_FRAC   host_relations.py:640-644, 774-799
        'Standard': (8.17, 1.01, 0.3, 0.615),
        'KH2013': (8.69, 1.17, 0.28, 0.615),
        'MM2013': (8.46, 1.05, 0.34, 0.615),
    }

    def __init__(self, kind='KH2013', mamp_log10=None, mplaw=None, mref=None, scatter_dex=None, bulge_frac=None):
        amp, plaw, scat, bfrac = self.KINDS[kind]
        self._mamp = MSOL * np.power(10.0, amp if mamp_log10 is None else mamp_log10)
        self._mplaw = plaw if mplaw is None else mplaw
        self._mref = 1.0e11 * MSOL if mref is None else mref
        self._scatter_dex = scat if scatter_dex is None else scatter_dex
        self._bfrac = bfrac if bulge_frac is None else bulge_frac

    def mbh_from_mbulge(self, mbulge):
        """host_relations.py:696-718 -> _log10_relation :1102-1134 (scatter off)"""
        yy = np.log10(mbulge/self._mref) * self._mplaw
        return self._mamp * np.power(10.0, yy)

    def mbulge_from_mbh(self, mbh):
        """host_relations.py:745-765 -> _log10_relation_reverse :1137-1178"""
        xx = np.log10(mbh/self._mamp)
        xx = (1.0/self._mplaw) * xx
        return self._mref * np.power(10.0, xx)

    def mstar_from_mbh(self, mbh):
        """host_relations.py:768-771 ; BF_Constant.mstar_from_mbulge :190-192 / BF_Sigmoid :297-308"""
        if isinstance(self._bfrac, BFSigmoid):
            return self._bfrac.mstar_from_mbulge(np.array(self.mbulge_from_mbh(mbh), dtype=float))
        return self.mbulge_from_mbh(mbh) / self._bfrac

    def mbh_from_mstar(self, mstar):
        """host_relations.py:514-536"""
        if isinstance(self._bfrac, BFSigmoid):
            return self.mbh_from_mbulge(self._bfrac.mbulge_from_mstar(mstar))
        return self.mbh_from_mbulge(mstar * self._bfrac)

    def dmbulge_dmbh(self, mbulge):
        """host_relations.py:720-743"""
        mbh = self.mbh_from_mbulge(mbulge)
        return mbulge / (self._mplaw * mbh)

    def dmstar_dmbh(self, mstar):
        """host_relations.py:483-512"""
        if isinstance(self._bfrac, BFSigmoid):
            mbulge = self._bfrac.mbulge_from_mstar(mstar)
            return self._bfrac.dmstar_dmbulge(mbulge) * self.dmbulge_dmbh(mbulge)
        mbulge = mstar * self._bfrac
        dmstar_dmbulge = 1.0 / self._bfrac
        return dmstar_dmbulge * self.dmbulge_dmbh(mbulge)


def redz_after(time, redz, cosmo, age_universe):
    """utils.redz_after utils.py:1772-1808 (array branch)."""
    age = cosmo.age(redz)
    new_age = age + time
    new_redz = -1.0 * np.ones_like(new_age)
    idx = (new_age < age_universe)
    new_redz[idx] = cosmo.tage_to_z(new_age[idx])
    return new_redz


def static_binary_density(mtot, mrat, redz, cosmo, gsmf, mmbulge, gpf=None, gmt=None, gmr=None, scatter=True):
    """Semi_Analytic_Model.static_binary_density, sams/sam.py:280-398 (module switches *_USES_MTOT = False).

    `gsmf(mstar, redz)`, `gpf(mass, mrat, redz)`, `gmt(mass, mrat, redz)`, `gmr(mtot, mrat, redz)` are
    callables (closures over the functions above).  Returns dict(dens, gmt_time, redz_prime, dens_noscatter).
    """
    # ---- mass_stellar sam.py:250-278
    rz = redz[np.newaxis, np.newaxis, :]
    masses = m1m2_from_mtmr(mtot[:, np.newaxis], mrat[np.newaxis, :])
    mbh_pri, mbh_sec, rz = np.broadcast_arrays(masses[0][..., np.newaxis], masses[1][..., np.newaxis], rz)
    mstar_pri = mmbulge.mstar_from_mbh(mbh_pri)
    mstar_sec = mmbulge.mstar_from_mbh(mbh_sec)
    mstar_rat = mstar_sec / mstar_pri
    mstar_tot = mstar_pri + mstar_sec
    mass_gsmf = mstar_pri                                   # sam.py:314 (GSMF_USES_MTOT False)

    gmt_time = None
    zprime = None
    idx_stalled = None
    if gmt is not None:                                     # sam.py:318-328
        gmt_time = gmt(mstar_pri, mstar_rat, rz)
        zprime = redz_after(gmt_time, rz, cosmo, cosmo.age(0.0))
        idx_stalled = (zprime < 0.0)

    if gmr is None:                                         # sam.py:335-344
        gal_merger_rate = gpf(mstar_pri, mstar_rat, rz) / gmt_time
    else:
        gal_merger_rate = gmr(mstar_tot, mstar_rat, rz)

    dens = gsmf(mass_gsmf, rz) * gal_merger_rate * cosmo.dtdz(rz)   # sam.py:347
    mplaw = mmbulge._mplaw
    dqbh_dqgal = mplaw * np.power(mstar_rat, mplaw - 1.0)
    dmstar_dmbh_pri = mmbulge.dmstar_dmbh(mstar_pri)
    qterm = (1.0 + mstar_rat) / (1.0 + mrat[np.newaxis, :, np.newaxis])
    dmstar_dmbh = dmstar_dmbh_pri * qterm
    dens = dens * ((mtot[:, np.newaxis, np.newaxis] / mstar_tot) * (dmstar_dmbh / dqbh_dqgal))   # sam.py:365
    dens_noscatter = dens.copy()
    dens_raw = dens.copy()                                  # what the scatter is applied to (stalled bins not zeroed yet)

    if scatter and (mmbulge._scatter_dex > 0.0):            # sam.py:368-389
        dens = add_scatter_to_masses(mtot, mrat, dens, mmbulge._scatter_dex)

    if idx_stalled is not None:                             # sam.py:392-394
        dens = dens.copy()
        dens[idx_stalled] = 0.0
        dens_noscatter[idx_stalled] = 0.0
    return dict(dens=dens, gmt_time=gmt_time, redz_prime=zprime, dens_noscatter=dens_noscatter, dens_raw=dens_raw)


# ---- M-Mbulge scatter: sams/sam.py:1291-1394 with utils.py:382-488 helpers (scipy, as in the reference)

def roll_rows(arr, roll_num):
    """utils.roll_rows utils.py:382-413"""
    roll = np.asarray(roll_num)
    nrows, ncols = arr.shape
    arr_roll = arr[:, [*range(ncols), *range(ncols-1)]].copy()
    strd_0, strd_1 = arr_roll.strides
    result = np.lib.stride_tricks.as_strided(arr_roll, (nrows, ncols, ncols), (strd_0, strd_1, strd_1))
    return result[np.arange(nrows), (ncols - roll) % ncols]


def get_scatter_weights(uniform_cents, dist):
    """utils.get_scatter_weights utils.py:416-452"""
    num = uniform_cents.size
    dx = np.diff(uniform_cents)
    assert np.allclose(dx, dx[0])
    dx = dx[0]
    dx = dx/2.0 + np.arange(num) * dx
    dx = np.concatenate([-dx[::-1], dx])
    return np.diff(dist.cdf(dx))


def _get_rolled_weights(log_cents, dist):
    """utils._get_rolled_weights utils.py:464-488"""
    num = log_cents.size
    weights = get_scatter_weights(log_cents, dist)
    weights = weights[np.newaxis, :] * np.ones((num, weights.size))
    roll = 1 - num + np.arange(num)
    weights = roll_rows(weights, roll)
    return weights[:, :num]


def _scatter_with_weights(dens, weights, axis=0):
    """utils._scatter_with_weights utils.py:455-461"""
    dens = np.moveaxis(dens, axis, 0)
    dens_new = np.einsum("j...,jk...", dens, weights)
    return np.moveaxis(dens_new, 0, axis)


def add_scatter_to_masses(mtot, mrat, dens, scatter, refine=4):
    """add_scatter_to_masses sams/sam.py:1291-1394"""
    dist = sp.stats.norm(loc=0.0, scale=scatter)
    output = np.zeros_like(dens)
    m1, m2 = m1m2_from_mtmr(mtot[:, np.newaxis], mrat[np.newaxis, :])
    grid_size = m1.shape[0] * refine
    mextr = [0.9*mtot[0]*mrat[0]/(1.0 + mrat[0]), mtot[-1]*(1.0 + mrat[0])/mrat[0]]
    mextr = [np.min(mextr), np.max(mextr)]
    _mgrid = np.logspace(*np.log10(mextr), grid_size)
    mgrid_log10 = np.log10(_mgrid)
    pts = tuple([np.log10(mm.flatten()) for mm in (m1, m2)])
    m1m2_grid = np.meshgrid(mgrid_log10, mgrid_log10, indexing='ij')
    dlay = None
    weights = _get_rolled_weights(mgrid_log10, dist)
    for ii in range(dens.shape[2]):
        dens_redz = dens[:, :, ii]
        points = pts if dlay is None else dlay
        interp = sp.interpolate.CloughTocher2DInterpolator(points, dens_redz.flatten())
        m1m2_dens = interp(tuple(m1m2_grid))
        if dlay is None:
            dlay = interp.tri
        bads = np.isnan(m1m2_dens) | (m1m2_dens < 0.0)
        if np.any(bads):
            temp = sp.interpolate.NearestNDInterpolator(points, dens_redz.flatten())(tuple(m1m2_grid))
            m1m2_dens[bads] = temp[bads]
        m1m2_dens = _scatter_with_weights(m1m2_dens, weights, axis=0)
        m1m2_dens = _scatter_with_weights(m1m2_dens, weights, axis=1)
        interp = sp.interpolate.RegularGridInterpolator((mgrid_log10, mgrid_log10), m1m2_dens)
        output[:, :, ii] = interp(pts, method='linear').reshape(m1.shape)
    return output


# ==================================================================================================
# Stub objects for the reference's duck-typed arguments (sam_cyutils.pyx:457-497)
# ==================================================================================================

class StubLog:
    def info(self, *args, **kwargs):
        pass


class StubSam:
    def __init__(self, mtot, mrat, redz, dens, gmt_time=None, redz_prime=None):
        self.mtot = np.ascontiguousarray(mtot, dtype=float)
        self.mrat = np.ascontiguousarray(mrat, dtype=float)
        self.redz = np.ascontiguousarray(redz, dtype=float)
        self.static_binary_density = np.ascontiguousarray(dens, dtype=float)
        self._gmt_time = None if gmt_time is None else np.ascontiguousarray(gmt_time, dtype=float)
        self._redz_prime = None if redz_prime is None else np.ascontiguousarray(redz_prime, dtype=float)
        self.shape = (self.mtot.size, self.mrat.size, self.redz.size)
        self._log = StubLog()


class StubCosmoTables:
    def __init__(self, grid_z, grid_dcom, grid_age):
        self._grid_z = np.ascontiguousarray(grid_z, dtype=float)
        self._grid_dcom = np.ascontiguousarray(grid_dcom, dtype=float)
        self._grid_age = np.ascontiguousarray(grid_age, dtype=float)


def ref_find_norm(time, mtot_edges, mrat_edges, sepa_init, rchar, gamma_inner, gamma_outer, nsteps):
    """Fixed_Time_2PL_SAM.__init__ hardening.py:1404-1416 through the compiled reference."""
    _, scy, _ = ref()
    mt, mr = np.meshgrid(mtot_edges, mrat_edges, indexing='ij')
    shape = mt.shape
    norm_log10 = scy.find_2pwl_hardening_norm(time, mt.flatten(), mr.flatten(), sepa_init, rchar,
                                              gamma_inner, gamma_outer, nsteps)
    return np.reshape(norm_log10, shape)


def ref_dbn(fobs_orb, sam, cosmo_tables, hard_kind, norm=None, sepa_init=None, rchar=None, gamma_inner=None,
            gamma_outer=None, num_steps=None):
    """sam_cyutils.dynamic_binary_number_at_fobs (pyx:421-504) through the compiled reference."""
    _, scy, holo = ref()
    if hard_kind == '2pwl':
        hard = holo.hardening.Fixed_Time_2PL_SAM(np.ascontiguousarray(norm), sepa_init, rchar, gamma_inner,
                                                 gamma_outer, num_steps)
    else:
        hard = holo.hardening.Hard_GW()
    redz_final, diff_num = scy.dynamic_binary_number_at_fobs(np.ascontiguousarray(fobs_orb), sam, hard, cosmo_tables)
    return redz_final, diff_num


def ref_integrate(edges, dnum):
    _, scy, _ = ref()
    return scy.integrate_differential_number_3dx1d(edges, dnum)


# ==================================================================================================
# gravwaves.py / single_sources.py glue
# ==================================================================================================

def char_strain_sq_from_bin_edges_redz(edges, redz, dcom_func):
    """gravwaves.char_strain_sq_from_bin_edges_redz gravwaves.py:694-725"""
    foo = edges[-1]
    df = np.diff(foo)
    fc = midpoints(foo)
    for dd in range(3):
        redz = np.moveaxis(redz, dd, 0)
        redz = midpoints(redz, axis=0)
        redz = np.moveaxis(redz, 0, dd)
    mt = midpoints(edges[0])
    mr = midpoints(edges[1])
    mc = chirp_mass_mtmr(mt[:, np.newaxis], mr[np.newaxis, :])
    mc = mc[:, :, np.newaxis, np.newaxis]
    dc = +np.inf * np.ones_like(redz)
    sel = (redz > 0.0)
    dc[sel] = dcom_func(redz[sel])
    fr = fc[np.newaxis, np.newaxis, np.newaxis, :] * (1.0 + redz)
    hs = gw_strain_source(mc, dc, fr)
    return (hs ** 2) * (fc / df)


def char_strain_sq_from_bin_edges(edges, dcom_func):
    """gravwaves.char_strain_sq_from_bin_edges gravwaves.py:760-783"""
    foo = edges[-1]
    df = np.diff(foo)
    fc = midpoints(foo)
    mt = midpoints(edges[0])
    mr = midpoints(edges[1])
    rz = midpoints(edges[2])
    mc = chirp_mass_mtmr(mt[:, np.newaxis], mr[np.newaxis, :])
    mc = mc[:, :, np.newaxis, np.newaxis]
    dc = dcom_func(rz)
    dc = dc[np.newaxis, np.newaxis, :, np.newaxis]
    fr = fc[np.newaxis, :] * (1.0 + rz[:, np.newaxis])
    fr = fr[np.newaxis, np.newaxis, :, :]
    hs = gw_strain_source(mc, dc, fr)
    return (hs ** 2) * (fc / df)


def rank_order(h2fdf, kind='stable'):
    """single_sources.py:89-93 -- ``np.argsort(-h2fdf[...,0].flatten())``.

    The reference uses numpy's default (unstable) quicksort; the many exact ties (cells with h=0)
    make its order implementation-defined.  Tests pin the order with ``kind='stable'`` and pass the
    *same* (msort, qsort, zsort) to both sides.
    """
    shape = h2fdf.shape[:3]
    indices = np.argsort(-h2fdf[..., 0].flatten(), kind=kind)
    unraveled = np.array(np.unravel_index(indices, shape))
    return indices, unraveled[0, :], unraveled[1, :], unraveled[2, :]


def ss_params_arrays(edges, redz, dcom_func):
    """params=True glue of ss_gws_redz, single_sources.py:112-139: returns (redz, dcom_final, sepa, angs)."""
    mt = midpoints(edges[0])
    for dd in range(3):
        redz = np.moveaxis(redz, dd, 0)
        redz = midpoints(redz, axis=0)
        redz = np.moveaxis(redz, 0, dd)
    dcom_final = +np.inf*np.ones_like(redz)
    sel = (redz > 0.0)
    redz[~sel] = -1.0
    redz[redz < 0] = -1.0
    dcom_final[sel] = dcom_func(redz[sel])
    fobs_orb_cents = midpoints(edges[-1])
    frst_orb_cents = fobs_orb_cents[np.newaxis, np.newaxis, np.newaxis, :] * (1.0 + redz)
    with np.errstate(divide='ignore', invalid='ignore'):
        sepa = kepler_sepa_from_freq(mt[:, np.newaxis, np.newaxis, np.newaxis], frst_orb_cents)
        dang = dcom_final / (1.0 + redz)       # utils.angs_from_sepa utils.py:1897-1917
        angs = sepa / dang
    return redz, dcom_final, sepa, angs


def ss_gws_redz(edges, redz, number, realize, loudest=1, params=False, dcom_func=None, seed=None, order_kind='stable'):
    """single_sources.ss_gws_redz single_sources.py:40-173 driving the compiled reference kernels."""
    cy, _, _ = ref()
    mt = midpoints(edges[0])
    mr = midpoints(edges[1])
    rz = midpoints(edges[2])
    h2fdf = char_strain_sq_from_bin_edges_redz(edges, redz, dcom_func)
    _, msort, qsort, zsort = rank_order(h2fdf, order_kind)
    if np.any(np.logical_and(redz < 0, redz != -1)):
        raise ValueError("redz < 0 and !=-1 found in redz, in ss_gws_redz()")
    cy.ORACLE_SEED = seed
    if params:
        rzf, dcom_final, sepa, angs = ss_params_arrays(edges, redz, dcom_func)
        hc2ss, hc2bg, sspar, bgpar = cy.loudest_hc_and_par_from_sorted_redz(
            number, h2fdf, realize, loudest, mt, mr, rz, rzf, dcom_final, sepa, angs, msort, qsort, zsort)
        return np.sqrt(hc2ss), np.sqrt(hc2bg), np.asarray(sspar), np.asarray(bgpar)
    hc2ss, hc2bg = cy.loudest_hc_from_sorted(number, h2fdf, realize, loudest, msort, qsort, zsort)
    return np.sqrt(hc2ss), np.sqrt(hc2bg)


def gws_from_number_grid_integrated_redz(edges, redz, number, realize, dcom_func, seed=None):
    """gravwaves._gws_from_number_grid_integrated_redz gravwaves.py:470-542 (sum=True branches)."""
    cy, _, _ = ref()
    hc2 = char_strain_sq_from_bin_edges_redz(edges, redz, dcom_func)
    if realize in [None, False]:
        hc2 = np.sum(hc2 * number, axis=(0, 1, 2))
    else:
        cy.ORACLE_SEED = seed
        hc2 = np.asarray(cy.sam_poisson_gwb(number, hc2, int(realize)))
    return np.sqrt(hc2)


# ---- supplied-count helpers: reproduce the seeded reference's draws in its own draw order -------

def _draws_in_order(lam, seed, thresh=1e10):
    """Draw, in C order of `lam`, exactly what the seeded reference kernels draw: `random_poisson(lam)`
    or, where ``lam > int(thresh)``, `random_normal(lam, sqrt(lam))` from one PCG64 stream
    (cyutils.pyx:886-895, 1326-1332).  numpy's Generator methods bind the same C routines."""
    gen = np.random.Generator(np.random.PCG64(seed))
    lam = np.ascontiguousarray(lam, dtype=float)
    big = lam > int(thresh)
    if not np.any(big):
        return gen.poisson(lam).astype(float)
    out = np.empty(lam.size)
    flat = lam.reshape(-1)
    bigf = big.reshape(-1)
    # vectorise the runs between normal-branch elements (the stream is strictly sequential)
    idx = np.flatnonzero(bigf)
    beg = 0
    for ii in idx:
        if ii > beg:
            out[beg:ii] = gen.poisson(flat[beg:ii])
        out[ii] = gen.normal(flat[ii], np.sqrt(flat[ii]))
        beg = ii + 1
    if beg < flat.size:
        out[beg:] = gen.poisson(flat[beg:])
    return out.reshape(lam.shape)


def counts_sam_poisson_gwb(number, nreals, seed, thresh=1e10):
    """Draws of `_sam_poisson_gwb` (cyutils.pyx:881-895, order m,q,z,f then r) -> (R, F, ncell) doubles."""
    F = number.shape[-1]
    lam = number.reshape(-1, F)
    cnt = _draws_in_order(np.repeat(lam.reshape(-1, 1), nreals, axis=1), seed, thresh)   # ((cell,f), r)
    cnt = cnt.reshape(lam.shape[0], F, nreals)
    return np.ascontiguousarray(np.transpose(cnt, (2, 1, 0)))


def counts_loudest(number, order, nreals, seed, thresh=1e10):
    """Draws of the `_loudest_*_from_sorted` kernels (cyutils.pyx:1318-1332: r, then f, then rank order)
    -> (R, F, ncell) doubles indexed by *natural* flat cell index."""
    F = number.shape[-1]
    lam = number.reshape(-1, F)
    ncell = lam.shape[0]
    lam_sorted = lam[order, :].T                                         # (F, ncell) in rank order
    draws = _draws_in_order(np.broadcast_to(lam_sorted, (nreals, F, ncell)), seed, thresh)
    out = np.empty((nreals, F, ncell))
    out[:, :, order] = draws
    return out


def counts_ss_bg(number, nreals, seed, thresh=1e10):
    """Draws of `_ss_bg_hc[_and_par]` (cyutils.pyx:981-1001: r, f, then natural m,q,z order)."""
    F = number.shape[-1]
    lam = number.reshape(-1, F)
    return _draws_in_order(np.broadcast_to(lam.T, (nreals, F, lam.shape[0])), seed, thresh)


# ==================================================================================================
# Eccentric evolution feeder: sams/sam.py:1235-1288
# ==================================================================================================

def evolve_eccen_uniform_single(mtot_edges, eccen_init, sepa_init, nsteps):
    eccen = np.zeros(nsteps)
    eccen[0] = eccen_init
    sepa_coal = SCHW * mtot_edges * 3
    sepa_min = sepa_coal.min()
    sepa = np.logspace(*np.log10([sepa_init, sepa_min]), nsteps)
    for step in range(1, nsteps):
        a0 = sepa[step-1]
        a1 = sepa[step]
        da = (a1 - a0)
        e0 = eccen[step-1]
        _, e1 = rk4_step(lambda aa, ee: 1.0 / gw_dade(aa, ee), x0=a0, y0=e0, dx=da)   # Hard_GW.deda hardening.py:185-208
        e1 = np.clip(e1, 0.0, None)
        eccen[step] = e1
    return sepa, eccen


# ---- librarian "details": lib_tools._calc_model_details, librarian/lib_tools.py:845-943

def calc_model_details(edges, redz_final, number, hc2):
    """lib_tools._calc_model_details lib_tools.py:866-943 given hc2 = char_strain_sq_from_bin_edges_redz(edges, redz_final)"""
    redz = edges[2]
    nmbins = len(edges[0]) - 1
    nzbins = len(redz) - 1
    nfreqs = len(edges[3]) - 1
    hc2_num = hc2 * number
    denom = np.sum(hc2_num, axis=(0, 1, 2))
    gwb_pars = []
    num_pars = []
    for ii in range(3):
        margins = [0, 1]
        if ii in margins:
            del margins[ii]
        margins = tuple(margins)
        gwb_pars.append(np.sum(hc2_num, axis=margins) / denom)
        num_pars.append(np.sum(number, axis=margins))
    rz = redz_final.copy()
    for ii in range(3):
        rz = midpoints(rz, axis=ii)
    gwb_mtot_redz_final = np.zeros((nmbins, nzbins, nfreqs))
    num_mtot_redz_final = np.zeros((nmbins, nzbins, nfreqs))
    gwb_rz = np.zeros((nzbins, nfreqs))
    num_rz = np.zeros((nzbins, nfreqs))
    for ii in range(nfreqs):
        rz_flat = rz[:, :, :, ii].flatten()
        numer, *_ = sp.stats.binned_statistic(rz_flat, hc2_num[:, :, :, ii].flatten(), bins=redz, statistic='sum')
        gwb_rz[:, ii] = numer / denom[ii]
        tpar, *_ = sp.stats.binned_statistic(rz_flat, number[:, :, :, ii].flatten(), bins=redz, statistic='sum')
        num_rz[:, ii] = tpar
        for mm in range(nmbins):
            rz_flat = rz[mm, :, :, ii].flatten()
            numer, *_ = sp.stats.binned_statistic(rz_flat, hc2_num[mm, :, :, ii].flatten(), bins=redz, statistic='sum')
            gwb_mtot_redz_final[mm, :, ii] = numer / denom[ii]
            tpar, *_ = sp.stats.binned_statistic(rz_flat, number[mm, :, :, ii].flatten(), bins=redz, statistic='sum')
            num_mtot_redz_final[mm, :, ii] = tpar
    gwb_pars.append(gwb_rz)
    num_pars.append(num_rz)
    return gwb_pars, num_pars, gwb_mtot_redz_final, num_mtot_redz_final
