"""Host-side utilities on the SAM GW-background path (subset of ``holodeck/utils.py``).

Only the functions reachable from ``Semi_Analytic_Model.gwb`` / ``librarian.run_model`` are
provided; each cites the reference lines it mirrors.  These are O(grid
edge) closed-form helpers that run on the host; everything O(grid cells) runs in CUDA.
"""
import inspect
import numbers

import numpy as np

from holodeck_b200 import log, cosmo
from holodeck_b200.constants import NWTG, SCHW, SPLC, YR, GYR

# utils.py:39-44
_GW_SRC_CONST = 8 * np.power(NWTG, 5/3) * np.power(np.pi, 2/3) / np.sqrt(10) / np.power(SPLC, 4)
_GW_DADT_SEP_CONST = - 64 * np.power(NWTG, 3) / 5 / np.power(SPLC, 5)
_GW_DEDT_ECC_CONST = - 304 * np.power(NWTG, 3) / 15 / np.power(SPLC, 5)
_GW_LUM_CONST = (32.0 / 5.0) * np.power(NWTG, 7.0/3.0) * np.power(SPLC, -5.0)

_AGE_UNIVERSE_GYR = cosmo.age_universe / GYR   # utils.py:46


def get_subclass_instance(value, default, superclass, allow_none=False):
    """Convert a class, instance or `None` into an instance of ``superclass`` (utils.py:320-367)."""
    if (value is None) and (default is not None):
        value = default
    if (value is None) and allow_none:
        return value
    if inspect.isclass(value):
        value = value()
    if not isinstance(value, superclass):
        err = f"argument ({value}) must be an instance or subclass of `{superclass}`!"
        log.error(err)
        raise ValueError(err)
    return value


def isnumeric(val):
    """utils.py:674-700"""
    try:
        float(str(val))
    except ValueError:
        return False
    return True


def isinteger(val):
    """utils.py:656-671"""
    return isnumeric(val) and isinstance(val, numbers.Integral)


def midpoints(vals, axis=-1, log=False):
    """Midpoints between adjacent values along ``axis`` (utils.py:703-711)."""
    mm = np.moveaxis(vals, axis, 0)
    if log:
        mm = np.log10(mm)
    mm = 0.5 * (mm[1:] + mm[:-1])
    if log:
        mm = 10.0 ** mm
    return np.moveaxis(mm, 0, axis)


def minmax(vals, filter=False):
    """utils.py:720-742"""
    vv = np.asarray(vals)
    if filter:
        vv = vv[np.isfinite(vv)]
    return np.array([np.min(vv), np.max(vv)])


def pta_freqs(dur=16.03*YR, num=40, cad=None):
    """Nyquist-sampled PTA frequency bin centers and edges (utils.py:835-874)."""
    fmin = 1.0 / dur
    if cad is not None:
        num = dur / (2.0 * cad)
        num = int(np.floor(num))
    cents = np.arange(1, num+2) * fmin
    edges = cents - fmin / 2.0
    cents = cents[:-1]
    return cents, edges


def rk4_step(func, x0, y0, dx, args=None, check_nan=0, check_nan_max=5):
    """One 4th-order Runge-Kutta step (utils.py:1018-1043)."""
    if args is None:
        k1 = dx * func(x0, y0)
        k2 = dx * func(x0 + dx/2.0, y0 + k1/2.0)
        k3 = dx * func(x0 + dx/2.0, y0 + k2/2.0)
        k4 = dx * func(x0 + dx, y0 + k3)
    else:
        k1 = dx * func(x0, y0, *args)
        k2 = dx * func(x0 + dx/2.0, y0 + k1/2.0, *args)
        k3 = dx * func(x0 + dx/2.0, y0 + k2/2.0, *args)
        k4 = dx * func(x0 + dx, y0 + k3, *args)
    y1 = y0 + (1.0/6.0) * (k1 + 2*k2 + 2*k3 + k4)
    x1 = x0 + dx
    if check_nan > 0 and not np.isfinite(y1):
        if check_nan > check_nan_max:
            raise RuntimeError("Failed to find finite step!  `check_nan` = {}!".format(check_nan))
        rk4_step(func, x0, y0, dx / 2.0, check_nan=check_nan+1, check_nan_max=check_nan_max)
    return x1, y1


def trapz(yy, xx, axis=-1, cumsum=True):
    """Trapezoid rule along ``axis`` (utils.py:1099-1136)."""
    xx = np.asarray(xx)
    if np.ndim(xx) == 1:
        pass
    elif np.ndim(xx) == np.ndim(yy):
        xx = xx[axis]
    else:
        err = f"Bad shape for `xx` (xx.shape={np.shape(xx)}, yy.shape={np.shape(yy)})!"
        log.error(err)
        raise ValueError(err)
    ct = np.moveaxis(yy, axis, 0)
    ct = 0.5 * (ct[1:] + ct[:-1])
    ct = np.moveaxis(ct, 0, -1)
    ct = ct * np.diff(xx)
    if cumsum:
        ct = np.cumsum(ct, axis=-1)
    return np.moveaxis(ct, -1, axis)


def _parse_val_log10_val_pars(val, val_log10, val_units=1.0, name='value', only_one=True):
    """utils.py:1299-1337"""
    both_or_neither = (val_log10 is not None) == (val is not None)
    if only_one and both_or_neither:
        err = f"One of {name} OR {name}_log10 must be provided!  {name}={val}, {name}_log10={val_log10}"
        log.exception(err)
        raise ValueError(err)
    if val is None:
        val = val_units * np.power(10.0, val_log10)
    if val_log10 is None:
        val_log10 = np.log10(val / val_units)
    return val, val_log10


def _integrate_grid_differential_number(edges, dnum, freq=False):
    """Host (numpy) trapezoid integration of dN over the grid (utils.py:1340-1373)."""
    number = trapz(dnum, np.log10(edges[0]), axis=0, cumsum=False)
    number = trapz(number, edges[1], axis=1, cumsum=False)
    number = trapz(number, edges[2], axis=2, cumsum=False)
    if freq:
        number = trapz(number, np.log(edges[3]), axis=3, cumsum=False)
    return number


def m1m2_from_mtmr(mt, mr):
    """Total mass and mass ratio -> (m1, m2), m2 <= m1 (utils.py:1620-1642)."""
    mt = np.asarray(mt)
    mr = np.asarray(mr)
    m1 = mt / (1.0 + mr)
    m2 = mt - m1
    return np.array([m1, m2])


def frst_from_fobs(fobs, redz):
    """utils.py:1645-1662"""
    return fobs * (1.0 + redz)


def fobs_from_frst(frst, redz):
    """utils.py:1665-1682"""
    return frst / (1.0 + redz)


def kepler_freq_from_sepa(mass, sepa):
    """utils.py:1685-1702"""
    return (1.0/(2.0*np.pi))*np.sqrt(NWTG*mass)/np.power(sepa, 1.5)


def kepler_sepa_from_freq(mass, freq):
    """utils.py:1705-1724"""
    mass = np.asarray(mass)
    freq = np.asarray(freq)
    return np.power(NWTG*mass/np.square(2.0*np.pi*freq), 1.0/3.0)


def schwarzschild_radius(mass):
    """utils.py:1811-1827"""
    return SCHW * mass


def rad_isco(m1, m2=0.0, factor=3.0):
    """utils.py:1727-1749"""
    return factor * schwarzschild_radius(m1+m2)


def redz_after(time, redz=None, age=None):
    """Redshift after ``time`` [s] has elapsed from ``redz`` (or ``age``); -1 past z=0 (utils.py:1772-1808)."""
    if (redz is None) == (age is None):
        raise ValueError("One of `redz` and `age` must be provided (and not both)!")
    if redz is not None:
        age = cosmo.age(redz)
    new_age = age + time
    if np.isscalar(new_age):
        if new_age < _AGE_UNIVERSE_GYR * GYR:
            new_redz = float(cosmo.tage_to_z(new_age))
        else:
            new_redz = -1.0
    else:
        new_redz = -1.0 * np.ones_like(new_age)
        idx = (new_age < _AGE_UNIVERSE_GYR * GYR)
        new_redz[idx] = cosmo.tage_to_z(new_age[idx])
    return new_redz


def angs_from_sepa(sepa, dcom, redz):
    """Angular separation [rad] (utils.py:1897-1917)."""
    dang = dcom / (1.0 + redz)
    return sepa / dang


def chirp_mass(m1, m2=None):
    """utils.py:1951-1975"""
    m1 = np.asarray(m1)
    if m2 is None:
        m1, m2 = np.moveaxis(m1, -1, 0)
    m2 = np.asarray(m2)
    return np.power(m1 * m2, 3.0/5.0)/np.power(m1 + m2, 1.0/5.0)


def chirp_mass_mtmr(mt, mr):
    """utils.py:1978-1997"""
    mt = np.asarray(mt)
    mr = np.asarray(mr)
    return mt * np.power(mr, 3.0/5.0) / np.power(1 + mr, 6.0/5.0)


def _gw_ecc_func(eccen):
    """GW hardening-rate eccentricity dependence F(e) (utils.py:2421-2441)."""
    e2 = eccen*eccen
    num = 1 + (73/24)*e2 + (37/96)*e2*e2
    den = np.power(1 - e2, 7/2)
    return num / den


def gw_dedt(m1, m2, sepa, eccen):
    """utils.py:2045-2074"""
    m1, m2, sepa, eccen = [np.asarray(vv) for vv in (m1, m2, sepa, eccen)]
    cc = _GW_DEDT_ECC_CONST
    e2 = eccen**2
    dedt = cc * m1 * m2 * (m1 + m2) / np.power(sepa, 4)
    dedt *= (1.0 + e2*121.0/304.0) * eccen / np.power(1 - e2, 5.0/2.0)
    return dedt


def gw_dade(sepa, eccen):
    """da/de due to GW emission (utils.py:2077-2102)."""
    sepa = np.asarray(sepa)
    eccen = np.asarray(eccen)
    e2 = eccen**2
    num = (1 + (73.0/24.0)*e2 + (37.0/96.0)*e2*e2)
    den = (1 - e2) * (1.0 + (121.0/304.0)*e2)
    return (12.0 / 19.0) * (sepa / eccen) * (num / den)


def gw_hardening_rate_dadt(m1, m2, sepa, eccen=None):
    """GW hardening rate da/dt [cm/s] (utils.py:2153-2183)."""
    m1, m2, sepa = [np.asarray(vv) for vv in (m1, m2, sepa)]
    dadt = _GW_DADT_SEP_CONST * m1 * m2 * (m1 + m2) / np.power(sepa, 3)
    if eccen is not None:
        dadt = dadt * _gw_ecc_func(np.asarray(eccen))
    return dadt


def dfdt_from_dadt(dadt, sepa, mtot=None, frst_orb=None):
    """Convert a hardening rate in separation to one in frequency (utils.py:1554-1587)."""
    if (mtot is None) and (frst_orb is None):
        err = "Either `mtot` or `frst_orb` must be provided!"
        log.exception(err)
        raise ValueError(err)
    if frst_orb is None:
        frst_orb = kepler_freq_from_sepa(mtot, sepa)
    dfdt = -1.5 * (frst_orb / sepa) * dadt
    return dfdt, frst_orb


def gw_hardening_rate_dfdt(m1, m2, frst_orb, eccen=None):
    """GW hardening rate in frequency (utils.py:2186-2211)."""
    m1, m2, frst_orb = [np.asarray(vv) for vv in (m1, m2, frst_orb)]
    sepa = kepler_sepa_from_freq(m1+m2, frst_orb)
    dfdt = gw_hardening_rate_dadt(m1, m2, sepa, eccen=eccen)
    dfdt, _ = dfdt_from_dadt(dfdt, sepa, frst_orb=frst_orb)
    return dfdt, frst_orb


def gw_hardening_timescale_freq(mchirp, frst):
    """GW hardening timescale f/(df/dt) of a circular binary (utils.py:2214-2234)."""
    mchirp, frst = np.asarray(mchirp), np.asarray(frst)
    return (5.0 / 96.0) * np.power(NWTG*mchirp/SPLC**3, -5.0/3.0) * np.power(2*np.pi*frst, -8.0/3.0)


def gw_lum_circ(mchirp, freq_orb_rest):
    """utils.py:2225-2257"""
    return _GW_LUM_CONST * np.power(2.0*np.pi*np.asarray(freq_orb_rest)*np.asarray(mchirp), 10.0/3.0)


def gw_strain_source(mchirp, dcom, freq_rest_orb):
    """Sky/polarisation-averaged strain of a circular binary (utils.py:2260-2285)."""
    mchirp, dcom, freq_rest_orb = [np.asarray(vv) for vv in (mchirp, dcom, freq_rest_orb)]
    return _GW_SRC_CONST * mchirp * np.power(2*mchirp*freq_rest_orb, 2/3) / dcom


def stats(vals, percs=None, prec=2, weights=None):
    """Short string of percentiles, for log messages (utils.py:1376-1420, simplified formatting)."""
    if percs is None:
        percs = [0, 16, 50, 84, 100]
    vals = np.asarray(vals)
    if vals.size == 0:
        return "[]"
    qq = np.percentile(vals[np.isfinite(vals)], percs) if np.isfinite(vals).any() else [np.nan]*len(percs)
    return ", ".join(f"{vv:.{prec}e}" for vv in qq)


def frac_str(vals, prec=2):
    """utils.py:582-601"""
    vals = np.asarray(vals)
    num = np.count_nonzero(vals)
    den = vals.size
    frc = num / den if den > 0 else np.nan
    return f"{num:.{prec}e}/{den:.{prec}e} = {frc:.{prec}e}"
