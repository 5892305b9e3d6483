"""Flat-LCDM cosmology: host-side stand-in for the reference's ``cosmopy.Cosmology`` object.

The reference builds ``holodeck.cosmo = cosmopy.Cosmology(h=0.6933, Om0=0.288, Ob0=0.0472,
size=200)`` (``holodeck/__init__.py:48-85``).  ``cosmopy`` is an un-vendored, un-pinned
dependency that is not present in this image, and no reference test pins any cosmology value, so
this is a *restatement of the published model* (flat LambdaCDM, no radiation -- astropy's
``FlatLambdaCDM`` with ``Tcmb0=0``), **parity unpinned** at this boundary:

* ``E(z) = sqrt(Om0 (1+z)^3 + 1 - Om0)``
* ``age(z) = 2/(3 H0 sqrt(OL)) asinh(sqrt(OL/Om0) (1+z)^-3/2)``   (closed form)
* ``tage_to_z`` is the closed-form inverse of ``age``
* ``dtdz(z) = 1 / (H0 (1+z) E(z))``
* ``comoving_distance(z) = c/H0 * int_0^z dz'/E(z')`` -- evaluated with the substitution
  ``s = (1+z)^-1/2`` (``d_c = c/H0 * int_s^1 2 ds / sqrt(Om0 + OL s^6)``, a smooth integrand) by
  fixed-order Gauss-Legendre.  The CUDA strain kernel uses the *same* quadrature
  (``csrc/holo_cosmo.cuh``) so host and device agree to rounding.

The attributes the native kernels consume are exactly those the reference's Cython reads
(``sam_cyutils.pyx:476,494``): ``_grid_z`` (decreasing, ending at 0), ``_grid_dcom`` [cm],
``_grid_age`` [s] (increasing, last element = age of the universe).
"""
import numpy as np

from .constants import MPC, KMPERSEC, SPLC, GYR

GL_ORDER = 16   #: Gauss-Legendre order of the comoving-distance quadrature (host and device)

_GL_X, _GL_W = np.polynomial.legendre.leggauss(GL_ORDER)


class Cosmology:
    """WMAP9-default flat LCDM (see module docstring).  Times in [s], distances in [cm]."""

    #: z=0 is appended automatically; log-spaced between consecutive values, linear to zero
    _Z_GRID = [1000.0, 10.0, 4.0, 2.0, 1.0, 0.5, 0.1, 0.01]

    def __init__(self, h=0.6933, Om0=0.2880, Ob0=0.0472, size=200):
        self.h = float(h)
        self.H0 = 100.0 * self.h                    #: [km/s/Mpc]
        self.Om0 = float(Om0)
        self.Ob0 = float(Ob0)
        self.Ode0 = 1.0 - self.Om0
        self._H0_cgs = self.H0 * KMPERSEC / MPC     #: [1/s]
        self.hubble_time = 1.0 / self._H0_cgs       #: [s]
        self.hubble_distance = SPLC / self._H0_cgs  #: [cm]
        self._size = int(size)

        zgrid = self._init_interp_grid(self._Z_GRID, max(self._size // len(self._Z_GRID), 2))
        self._grid_z = np.ascontiguousarray(zgrid)
        self._grid_age = np.ascontiguousarray(self.age(zgrid))
        self._grid_dcom = np.ascontiguousarray(self.comoving_distance(zgrid))
        return

    @staticmethod
    def _init_interp_grid(z_pnts, num_pnts):
        z0 = z_pnts[0]
        segs = []
        for z1 in z_pnts[1:]:
            segs.append(np.logspace(*np.log10([z0, z1]), num=num_pnts, endpoint=False))
            z0 = z1
        segs.append(np.linspace(z0, 0.0, num=num_pnts))
        return np.concatenate(segs)

    # ---- closed-form background quantities

    def efunc(self, zz):
        zp1 = 1.0 + np.asarray(zz, dtype=float)
        return np.sqrt(self.Om0 * zp1 * zp1 * zp1 + self.Ode0)

    def dtdz(self, zz):
        """|dt/dz| in [s] (``sam.py:347`` multiplies the density by this)."""
        zz = np.asarray(zz, dtype=float)
        return self.hubble_time / ((1.0 + zz) * self.efunc(zz))

    def age(self, zz):
        """Age of the universe at redshift ``zz`` in [s]."""
        zp1 = 1.0 + np.asarray(zz, dtype=float)
        sq = np.sqrt(self.Ode0)
        arg = np.sqrt(self.Ode0 / self.Om0) * np.power(zp1, -1.5)
        return (2.0 / 3.0) * self.hubble_time / sq * np.arcsinh(arg)

    def tage_to_z(self, age):
        """Redshift at which the universe has the given age [s] (closed-form inverse of `age`)."""
        age = np.asarray(age, dtype=float)
        sq = np.sqrt(self.Ode0)
        sh = np.sinh(1.5 * sq * age / self.hubble_time)
        zp1 = np.power(np.sqrt(self.Ode0 / self.Om0) / sh, 2.0 / 3.0)
        return zp1 - 1.0

    def comoving_distance(self, zz):
        """Line-of-sight comoving distance in [cm] (Gauss-Legendre, see module docstring)."""
        zz = np.asarray(zz, dtype=float)
        sq = np.sqrt(1.0 + zz)
        # (1 - s0)/2 with s0 = 1/sqrt(1+z), written without cancellation at small z
        half = 0.5 * zz / (sq * (sq + 1.0))
        mid = 1.0 - half
        tot = np.zeros_like(zz)
        for xx, ww in zip(_GL_X, _GL_W):
            ss = mid + half * xx
            s2 = ss * ss
            tot = tot + ww * (2.0 / np.sqrt(self.Om0 + self.Ode0 * s2 * s2 * s2))
        return self.hubble_distance * half * tot

    def luminosity_distance(self, zz):
        zz = np.asarray(zz, dtype=float)
        return (1.0 + zz) * self.comoving_distance(zz)

    @property
    def age_universe(self):
        return float(self.age(0.0))

    def __repr__(self):
        return f"Cosmology(h={self.h}, Om0={self.Om0}, Ob0={self.Ob0}, size={self._size})"
