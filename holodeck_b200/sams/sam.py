"""Semi-Analytic Model of MBH-binary populations: same public API as ``holodeck/sams/sam.py``.

``Semi_Analytic_Model(mtot, mrat, redz, shape, gsmf, gpf, gmt, gmr, mmbulge)`` builds the same
log-spaced (M, q, z) grid as the reference (``sam.py:105-234``).  The heavy methods run on the GPU:

=================================  ==========================================  =======================
method                             reference                                   device kernel(s)
=================================  ==========================================  =======================
``static_binary_density``          ``sam.py:280-398``                          K0 ``holo_sam_density``
``dynamic_binary_number_at_fobs``  ``sam.py:400-466`` -> ``sam_cyutils``       K1b / K1c
``gwb``                            ``sam.py:872-947``                          K1 -> K2+K2b -> K4
``gwb_new``                        ``sam.py:787-813``                          K1 -> K2+K2b -> K3
=================================  ==========================================  =======================

Device-resident intermediates (density, merger times, the (M,Q,Z,F) grids) never leave HBM inside
``gwb``; numpy copies of ``static_binary_density``, ``_gmt_time`` and ``_redz_prime`` are made lazily
on first attribute access so user code that reads them keeps working.

The M-Mbulge scatter (``add_scatter_to_masses``, ``sam.py:1291-1394``; row N1 of SURVEY.md section 8f) runs on
the device too (``sams/scatter.py`` + K6): only its data-independent geometry (the Delaunay triangulation scipy
builds inside ``CloughTocher2DInterpolator``) is prepared on the host, once per grid.
"""
import ctypes as C
from datetime import datetime

import numpy as np

import holodeck_b200 as holo
from holodeck_b200 import _lib, cosmo, utils, log, host_relations
from holodeck_b200.constants import SPLC, MSOL, MPC
from holodeck_b200.sams.scatter import add_scatter_to_masses   # noqa: F401  (module-level name, as in the reference)
from holodeck_b200.sams.components import (
    _Galaxy_Pair_Fraction, _Galaxy_Stellar_Mass_Function, _Galaxy_Merger_Time, _Galaxy_Merger_Rate,
    GSMF_Schechter, GSMF_Double_Schechter, GPF_Power_Law, GMT_Power_Law, GMR_Illustris
)

REDZ_SAMPLE_VOLUME = True    #: get redshifts by sampling uniformly in 3D spatial volume, and converting

GSMF_USES_MTOT = False       #: the mass used in the GSMF is interpretted as M=m1+m2, otherwise use primary m1
GPF_USES_MTOT = False        #: the mass used in the GPF  is interpretted as M=m1+m2, otherwise use primary m1
GMT_USES_MTOT = False        #: the mass used in the GMT  is interpretted as M=m1+m2, otherwise use primary m1


class Semi_Analytic_Model:
    """Semi-Analytic Model (SAM) of MBH Binary populations (see module docstring)."""

    def __init__(
        self,
        mtot=(1.0e4*MSOL, 1.0e12*MSOL, 91),
        mrat=(1e-3, 1.0, 81),
        redz=(1e-3, 10.0, 101),
        shape=None,
        log=None,
        gsmf=GSMF_Schechter,
        gpf=None,
        gmt=None,
        gmr=None,
        mmbulge=host_relations.MMBulge_KH2013,
        **kwargs
    ):
        if log is None:
            log = holo.log
        self._log = log

        # Process deprecated or unexpected kwargs   (sam.py:150-162)
        deprecated_keys = ['ZERO_DYNAMIC_STALLED_SYSTEMS', 'ZERO_GMT_STALLED_SYSTEMS']
        for key, val in kwargs.items():
            if key in deprecated_keys:
                log.error(f"Using deprecated kwarg: {key}: {val}!  In the future this will raise an error.")
            else:
                err = f"Unexpected kwarg {key=}: {val=}!"
                log.exception(err)
                raise ValueError(err)

        # Sanitize input classes/instances   (sam.py:164-186)
        gsmf = utils.get_subclass_instance(gsmf, None, _Galaxy_Stellar_Mass_Function)
        mmbulge = utils.get_subclass_instance(mmbulge, None, host_relations._MMBulge_Relation)
        # if GPF is None, then we must use a GMR
        if gpf is None:
            log.info("No galaxy pair-fraction given, using galaxy merger-rate.")
            gmr = utils.get_subclass_instance(gmr, GMR_Illustris, _Galaxy_Merger_Rate)
            # if GMR is used, `gmt` can still be used (for calculating stalling), but doesn't have to be
            gmt = utils.get_subclass_instance(gmt, None, _Galaxy_Merger_Time, allow_none=True)
        # if GPF is given, GMR must not be, and a GMT is required
        else:
            if gmr is not None:
                err = "Can only use one of `gpf` and `gmr`!"
                log.exception(err)
                raise ValueError(err)
            log.info("Galaxy pair-fraction provided, using galaxy pair-fraction and merger-time.")
            gmt = utils.get_subclass_instance(gmt, GMT_Power_Law, _Galaxy_Merger_Time)
            gpf = utils.get_subclass_instance(gpf, GPF_Power_Law, _Galaxy_Pair_Fraction)

        self._gsmf = gsmf             #: Galaxy Stellar-Mass Function (`_Galaxy_Stellar_Mass_Function` instance)
        self._gmr = gmr               #: Galaxy Merger Rate (`_Galaxy_Merger_Rate` instance)
        self._gpf = gpf               #: Galaxy Pair Fraction (`_Galaxy_Pair_Fraction` instance)
        self._gmt = gmt               #: Galaxy Merger Time (`_Galaxy_Merger_Time` instance)
        self._mmbulge = mmbulge       #: Mbh-Mbulge relation (`host_relations._MMBulge_Relation` instance)
        log.debug(f"{gsmf=}, {gmr=}, {gpf=}, {gmt=}, {mmbulge=}")

        # ---- Create SAM grid edges   (sam.py:190-222)
        if shape is not None:
            if np.isscalar(shape):
                shape = [shape for ii in range(3)]

        params = [mtot, mrat, redz]
        param_names = ['mtot', 'mrat', 'redz']
        for ii, (par, name) in enumerate(zip(params, param_names)):
            if not isinstance(par, tuple) and (len(par) == 3):
                err = (
                    f"{name} (type={type(par)}, len={len(par)}) must be a (3,) tuple specifying a log-spacing, "
                    "or ndarray of grid edges!"
                )
                log.exception(err)
                raise ValueError(err)
            par = [pp for pp in par]
            if shape is not None:
                if shape[ii] is not None:
                    par[2] = shape[ii]
            params[ii] = np.logspace(*np.log10(par[:2]), par[2])
            log.debug(f"{name}: [{params[ii][0]}, {params[ii][-1]}] {params[ii].size}")

        mtot, mrat, redz = params
        self.mtot = mtot
        self.mrat = mrat
        self.redz = redz

        # These values are calculated as needed by the class when the corresponding methods are called
        self._density_host = None     #: numpy copy of the binary comoving number-density (lazy)
        self._density_dev = None      #: the same on the device
        self._shape = None            #: Shape of the parameter-space domain (mtot, mrat, redz)
        self._gmt_time_dev = None
        self._redz_prime_dev = None
        self._gmt_time_host = None
        self._redz_prime_host = None
        return

    # ---- grid

    @property
    def edges(self):
        """The grid edges defining the domain (list of: [`mtot`, `mrat`, `redz`])"""
        return [self.mtot, self.mrat, self.redz]

    @property
    def shape(self):
        """Shape of the parameter space domain (number of edges in each dimension), (3,) tuple"""
        if self._shape is None:
            self._shape = tuple([len(ee) for ee in self.edges])
        return self._shape

    def mass_stellar(self):
        """Stellar masses for each MBH based on the M-MBulge relation (``sam.py:250-278``; host numpy).

        Returns ``mstar_pri, mstar_rat, mstar_tot, redz``, each (M, Q, Z).
        """
        redz = self.redz[np.newaxis, np.newaxis, :]
        masses = utils.m1m2_from_mtmr(self.mtot[:, np.newaxis], self.mrat[np.newaxis, :])
        mbh_pri = masses[0]
        mbh_sec = masses[1]
        args = [mbh_pri[..., np.newaxis], mbh_sec[..., np.newaxis], redz]
        mbh_pri, mbh_sec, redz = np.broadcast_arrays(*args)
        mstar_pri = self._mmbulge.mstar_from_mbh(mbh_pri, redz=redz, scatter=False)
        mstar_sec = self._mmbulge.mstar_from_mbh(mbh_sec, redz=redz, scatter=False)
        mstar_rat = mstar_sec / mstar_pri
        mstar_tot = mstar_pri + mstar_sec
        return mstar_pri, mstar_rat, mstar_tot, redz

    # ---- density (K0)

    def _kernel_params(self):
        """Pack the component parameters for the fused density kernel (``holo_sam_params``)."""
        par = _lib.SamParams()
        gsmf, gpf, gmt, gmr, mmb = self._gsmf, self._gpf, self._gmt, self._gmr, self._mmbulge

        def fusable(obj, classes, what):
            # The kernel evaluates the closed form from `_kernel_params()` alone: a subclass that overrides ANY method
            # (`__call__`, `_phi_func`, `zprime`, ...) would be silently ignored on the device, so only the exact
            # classes -- or subclasses that add nothing but new default parameters through `__init__` -- are accepted.
            ok = type(obj) in classes
            if not ok and isinstance(obj, classes):
                base = next(cc for cc in classes if isinstance(obj, cc))
                added = set()
                for klass in type(obj).__mro__:
                    if klass is base:
                        break
                    added |= {nn for nn, vv in vars(klass).items() if callable(vv) or isinstance(vv, property)}
                ok = added <= {"__init__"}
            if not ok:
                raise NotImplementedError(
                    f"{what} {obj!r}: only the closed-form reference classes {[cc.__name__ for cc in classes]} "
                    "(or subclasses that override nothing but `__init__`) can be fused into the CUDA density "
                    "kernel; holodeck_b200 has no CPU fallback for user-defined components.")

        fusable(gsmf, (GSMF_Schechter, GSMF_Double_Schechter), "gsmf")
        kind, vals = gsmf._kernel_params()
        par.gsmf_kind = kind
        for ii, vv in enumerate(vals):
            par.gsmf[ii] = vv
        par.gsmf_uses_mtot = int(GSMF_USES_MTOT)
        par.gpf_uses_mtot = int(GPF_USES_MTOT)
        par.gmt_uses_mtot = int(GMT_USES_MTOT)

        par.use_gmr = int(gmr is not None)
        if gmr is not None:
            fusable(gmr, (GMR_Illustris,), "gmr")
            for ii, vv in enumerate(gmr._kernel_params()):
                par.gmr[ii] = vv
        else:
            fusable(gpf, (GPF_Power_Law,), "gpf")
            for ii, vv in enumerate(gpf._kernel_params()):
                par.gpf[ii] = vv
        par.has_gmt = int(gmt is not None)
        if gmt is not None:
            fusable(gmt, (GMT_Power_Law,), "gmt")
            for ii, vv in enumerate(gmt._kernel_params()):
                par.gmt[ii] = vv

        bfrac = mmb._bulge_frac
        if type(mmb) not in (host_relations.MMBulge_Standard, host_relations.MMBulge_KH2013, host_relations.MMBulge_MM2013) or \
                type(bfrac) not in (host_relations.BF_Constant, host_relations.BF_Sigmoid):
            raise NotImplementedError(
                f"mmbulge {mmb!r}: only the power-law `MMBulge_Standard` relations with `BF_Constant` or `BF_Sigmoid` "
                "bulge fractions are fused into the CUDA density kernel (SURVEY.md section 2a row 6).")
        par.mmb[0] = mmb._mamp
        par.mmb[1] = mmb._mplaw
        par.mmb[2] = mmb._mref
        self._bf_tables_host = None
        if isinstance(bfrac, host_relations.BF_Constant):
            par.bf_kind = 0
            par.mmb[3] = bfrac.bulge_frac()
        else:
            par.bf_kind = 1
            (x0, c0), (x1, c1) = bfrac._kernel_tables()
            assert x0.size == x1.size
            par.bf_n = int(x0.size - 1)
            for ii, vv in enumerate((bfrac._bulge_frac_lo, bfrac._bulge_frac_hi, bfrac._mstar_char, bfrac._width_dex)):
                par.bf[ii] = vv
            self._bf_tables_host = np.concatenate([x0, c0.ravel(), x1, c1.ravel()])
        par.hubble_time = cosmo.hubble_time
        par.om0 = cosmo.Om0
        par.age_universe = utils._AGE_UNIVERSE_GYR * holo.constants.GYR
        return par

    def _compute_density(self):
        lib = _lib.require_gpu()
        log = self._log
        M, Q, Z = self.shape
        par = self._kernel_params()
        mtot, mrat, redz = [_lib.to_dev(vv) for vv in (self.mtot, self.mrat, self.redz)]
        age_z = _lib.to_dev(cosmo.age(self.redz))
        dtdz_z = _lib.to_dev(cosmo.dtdz(self.redz))
        dens = _lib.empty((M, Q, Z))
        has_gmt = self._gmt is not None
        gmt_time = _lib.empty((M, Q, Z)) if has_gmt else None
        zprime = _lib.empty((M, Q, Z)) if has_gmt else None
        bf_tab = None if self._bf_tables_host is None else _lib.to_dev(self._bf_tables_host)
        rc = lib.holo_sam_density(_lib.ptr(mtot), _lib.ptr(mrat), _lib.ptr(redz), _lib.ptr(age_z), _lib.ptr(dtdz_z),
                                  M, Q, Z, C.byref(par), _lib.ptr(bf_tab), _lib.ptr(dens), _lib.ptr(gmt_time), _lib.ptr(zprime),
                                  _lib.stream())
        _lib.check(rc, "static_binary_density")
        if has_gmt:
            self._gmt_time_dev = gmt_time
            self._redz_prime_dev = zprime
        else:
            log.info("No GMT was provided, cannot calculate Galaxy-Merger based stalling.")

        # ---- Add scatter from the M-Mbulge relation   (sam.py:368-389)
        scatter = self._mmbulge._scatter_dex
        log.debug(f"mmbulge scatter = {scatter}")
        if scatter > 0.0:
            log.info(f"Adding MMbulge scatter ({scatter:.4e})")
            mass_bef = self._integrated_binary_density_device(dens)
            self._dens_bef_dev = dens
            pending = []
            dens = add_scatter_to_masses(self.mtot, self.mrat, dens, scatter, log=log, _defer_check=pending)   # device (K6)
            self._scatter_flags_dev = pending[0]
            self._dens_aft_dev = dens.clone() if has_gmt else dens    # (the stalled bins are zeroed in place below)
            mass_aft = self._integrated_binary_density_device(dens)
            # The reference logs the change of the integrated number here (sam.py:381-389).  The two integrals stay on
            # the device: reading them back now would stall the launch queue twice per SAM; the message is emitted
            # the next time the host looks at this model's results anyway (`_report_scatter_mass`).
            self._scatter_mass_dev = (mass_bef, mass_aft)

        # set values after redshift zero to have zero density   (sam.py:392-394)
        if has_gmt:
            rc = lib.holo_zero_stalled(_lib.ptr(dens), _lib.ptr(zprime), dens.numel(), _lib.stream())
            _lib.check(rc, "zero_stalled")
        self._density_dev = dens

    def _static_binary_density_device(self):
        if self._density_dev is None:
            self._compute_density()
        return self._density_dev

    def _report_scatter_mass(self):
        """Deferred log of the change of the integrated number by the M-Mbulge scatter (``sam.py:381-389``)."""
        pair = getattr(self, "_scatter_mass_dev", None)
        if pair is None:
            return
        self._scatter_mass_dev = None
        flags = getattr(self, "_scatter_flags_dev", None)
        if flags is not None and int(flags.item()) != 0:
            from holodeck_b200.sams.scatter import _bad_values_error
            raise _bad_values_error(self._log)                # sam.py:1376-1380
        mass_bef, mass_aft = (float(vv.item()) for vv in pair)
        dm = (mass_aft - mass_bef) / mass_bef
        msg = f"mass: {mass_bef:.2e} ==> {mass_aft:.2e} || change = {dm:.4e}"
        self._log.info(f"Scatter added\t{msg}")
        if np.fabs(dm) > 0.2:
            self._log.error(f"Warning, significant change in number-mass!  {msg}")

    @property
    def static_binary_density(self):
        """Number-density of binaries d^3 n / [dlog10(M) dq dz] in [Mpc^-3], shape (M, Q, Z).

        Cached after the first access, like the reference (``sam.py:280-398``).
        """
        if self._density_host is None:
            self._density_host = _lib.to_host(self._static_binary_density_device())
            self._report_scatter_mass()
        return self._density_host

    @property
    def _density(self):
        return self._density_host

    def _gmt_time_device(self):
        return self._gmt_time_dev

    def _redz_prime_device(self):
        return self._redz_prime_dev

    @property
    def _gmt_time(self):
        """(M, Q, Z) galaxy-merger time [s]; `None` until the density exists or if there is no GMT."""
        if (self._gmt_time_host is None) and (self._gmt_time_dev is not None):
            self._gmt_time_host = _lib.to_host(self._gmt_time_dev)
        return self._gmt_time_host

    @property
    def _redz_prime(self):
        """(M, Q, Z) redshift after the galaxy merger (-1 if after z=0); `None` as for `_gmt_time`."""
        if (self._redz_prime_host is None) and (self._redz_prime_dev is not None):
            self._redz_prime_host = _lib.to_host(self._redz_prime_dev)
        return self._redz_prime_host

    def _integrated_binary_density_device(self, dens):
        """`_integrated_binary_density(dens, sum=True)` for a device array (trapezoid over the three grid axes)."""
        import torch
        integ = torch.trapezoid(dens, _lib.to_dev(np.log10(self.mtot)), dim=0)
        integ = torch.trapezoid(integ, _lib.to_dev(self.mrat), dim=0)
        integ = torch.trapezoid(integ, _lib.to_dev(self.redz), dim=0)
        return integ                      # 0-d device tensor

    @property
    def _dens_bef(self):
        """Density before the M-Mbulge scatter was applied (``sam.py:374``), numpy."""
        return None if getattr(self, "_dens_bef_dev", None) is None else _lib.to_host(self._dens_bef_dev)

    @property
    def _dens_aft(self):
        """Density after the M-Mbulge scatter was applied (``sam.py:377``), numpy."""
        return None if getattr(self, "_dens_aft_dev", None) is None else _lib.to_host(self._dens_aft_dev)

    def _integrated_binary_density(self, ndens=None, sum=True):
        """Integrate the binary number-density over the grid (``sam.py:678-703``; host numpy)."""
        if ndens is None:
            ndens = self.static_binary_density
        integ = utils.trapz(ndens, np.log10(self.mtot), axis=0, cumsum=False)
        integ = utils.trapz(integ, self.mrat, axis=1, cumsum=False)
        integ = utils.trapz(integ, self.redz, axis=2, cumsum=False)
        if sum:
            integ = integ.sum()
        return integ

    # ---- number (K1)

    def dynamic_binary_number_at_fobs(self, hard, fobs_orb, use_cython=True, **kwargs):
        """Differential number of binaries at the given observer-frame orbital frequencies.

        Returns ``(grid, dnum, redz_final)`` as ``sam.py:400-466``.  ``use_cython=True`` (default)
        selects the native path -- here the CUDA kernels.  The reference's pure-numpy alternatives
        (``_dynamic_binary_number_at_fobs_consistent/_inconsistent``, ``sam.py:468-676``) exist there
        only as slow cross-checks and are not provided: ``use_cython=False`` raises.
        """
        if not use_cython:
            raise NotImplementedError(
                "holodeck_b200 only provides the native (CUDA) `dynamic_binary_number_at_fobs`; the reference's "
                "numpy cross-check implementations (sam.py:468-676) are not part of the hot path.")
        from holodeck_b200.sams import sam_cyutils
        redz_final, dnum = sam_cyutils.dynamic_binary_number_at_fobs(fobs_orb, self, hard, cosmo)
        grid = [self.mtot, self.mrat, self.redz, fobs_orb]
        return grid, dnum, redz_final

    # ---- GWB drivers

    def _number_and_strain(self, fobs_gw_edges, hard, params):
        """K1 -> fused K2+K2b on the device; returns (edges, redz_final, strain dict incl. `number`)."""
        from holodeck_b200.sams import sam_cyutils
        from holodeck_b200 import gravwaves
        fobs_gw_edges = np.asarray(fobs_gw_edges, dtype=float)
        fobs_gw_cents = utils.midpoints(fobs_gw_edges)
        fobs_orb_edges = fobs_gw_edges / 2.0
        fobs_orb_cents = fobs_gw_cents / 2.0
        redz_final, diff_num = sam_cyutils.dynamic_binary_number_at_fobs(fobs_orb_cents, self, hard, cosmo, device=True)
        edges = [self.mtot, self.mrat, self.redz, fobs_orb_edges]
        strain = gravwaves._char_strain_sq(edges, redz_final, params=params, dnum=diff_num)
        return edges, redz_final, strain

    def gwb_new(self, fobs_gw_edges, hard=None, realize=100, *, seed=None):
        """GWB (no single-source split) through ``sam_poisson_gwb`` (``sam.py:787-813``); returns hc (F, R)."""
        from holodeck_b200 import gravwaves
        if hard is None:
            hard = holo.hardening.Hard_GW()
        assert isinstance(hard, (holo.hardening.Fixed_Time_2PL_SAM, holo.hardening.Hard_GW))
        edges, redz_final, strain = self._number_and_strain(fobs_gw_edges, hard, params=False)
        return gravwaves._gws_from_hc2(strain["h2fdf"], strain["number"], realize, True, seed, 0, False)

    def gwb_old(self, fobs_gw_edges, hard=None, realize=100, *, seed=None):
        """GWB through ``dynamic_binary_number_at_fobs`` + per-axis trapezoids (``sam.py:815-835``); returns hc (F, R).

        The reference integrates ``dnum`` with three successive ``utils.trapz`` calls (``utils.py:1340-1373``) and
        multiplies by ``diff(ln fobs_gw_edges)``: the product of the three one-dimensional trapezoid rules is the
        8-corner mean times the bin volume that ``integrate_differential_number_3dx1d`` (K2) evaluates in one pass, and
        ``diff(ln f_gw) == diff(ln f_orb)``; the two agree to rounding (1e-15), so this is the ``gwb_new`` device
        path under the reference's older name.  ``hard`` may be a class (the reference's default is the class
        ``Hard_GW``) or an instance."""
        if hard is None:
            hard = holo.hardening.Hard_GW
        if isinstance(hard, type):
            hard = hard()
        return self.gwb_new(np.atleast_1d(fobs_gw_edges), hard=hard, realize=realize, seed=seed)

    def gwb_ideal(self, fobs_gw, sum=True, redz_prime=None):
        """Idealized, continuous GWB amplitude, [Phinney2001]_ Eq.5 (``sam.py:837-870``; host)."""
        from holodeck_b200 import gravwaves
        mstar_pri, mstar_rat, mstar_tot, redz = self.mass_stellar()
        # default to using `redz_prime` values if a GMT instance is stored
        if redz_prime is None:
            redz_prime = (self._gmt is not None)
        elif redz_prime and (self._gmt is None):
            err = "No `GMT` instance stored, cannot use `redz_prime` values!"
            self._log.exception(err)
            raise AttributeError(err)
        rz = self.redz
        if redz_prime:
            gmt_mass = mstar_tot if GMT_USES_MTOT else mstar_pri
            rz, _ = self._gmt.zprime(gmt_mass, mstar_rat, rz)
        ndens = self.static_binary_density / (MPC**3)
        mt = self.mtot[:, np.newaxis, np.newaxis]
        mr = self.mrat[np.newaxis, :, np.newaxis]
        return gravwaves.gwb_ideal(fobs_gw, ndens, mt, mr, rz, dlog10=True, sum=sum)

    def gwb(self, fobs_gw_edges, hard=None, realize=100, loudest=1, params=False, *, seed=None, r0=0, device=False):
        """Calculate the (smooth/semi-analytic) GWB and CWs at the given observed GW-frequencies.

        Parameters
        ----------
        fobs_gw_edges : (F+1,) array_like of scalar,
            Observer-frame GW-frequency bin edges [1/sec].
        hard : `Hard_GW` or `Fixed_Time_2PL_SAM` class or instance
            Hardening mechanism to apply over the range of `fobs_gw`.
        realize : int
            Number of discrete realizations to construct.
        loudest : int
            Number of loudest single sources to distinguish from the background.
        params : bool
            Whether or not to return astrophysical parameters of the binaries.
        seed, r0, device : keyword-only additions (see ``holodeck_b200.cyutils``); ``device=True`` returns CUDA tensors.

        Returns
        -------
        hc_ss : (F, R, L) characteristic strain of the L loudest single sources at each frequency.
        hc_bg : (F, R) characteristic strain of the background.
        sspar : (4, F, R, L), bgpar : (7, F, R) -- only if ``params``.

        Mirrors ``sam.py:872-947``.
        """
        from holodeck_b200 import single_sources
        if hard is None:
            hard = holo.hardening.Hard_GW()
        if isinstance(hard, type) and issubclass(hard, holo.hardening.Hard_GW):
            hard = hard()
        if not isinstance(hard, (holo.hardening.Fixed_Time_2PL_SAM, holo.hardening.Hard_GW)):
            err = (
                "`sam_cyutils` methods only work with `Fixed_Time_2PL_SAM` or `Hard_GW` hardening models!  "
                "Use `gwb_only` for alternative classes!"
            )
            self._log.exception(err)
            raise ValueError(err)

        edges, redz_final, strain = self._number_and_strain(fobs_gw_edges, hard, params=bool(params))
        number = strain["number"]
        ret_vals = single_sources.ss_gws_redz(edges, redz_final, number, realize=realize, loudest=loudest,
                                              params=params, seed=seed, r0=r0, device=device, _precomputed=strain)
        hc_ss = ret_vals[0]
        hc_bg = ret_vals[1]
        self._report_scatter_mass()        # (the draws have synchronised: reading two scalars back is free now)
        if params:
            return hc_ss, hc_bg, ret_vals[2], ret_vals[3]
        return hc_ss, hc_bg


# ===========================================
# ====    Evolution & Utility Functions    ====
# ===========================================

def evolve_eccen_uniform_single(sam, eccen_init, sepa_init, nsteps):
    """Evolve binary eccentricity from an initial value along a range of separations (``sam.py:1235-1288``).

    A 1-D, `nsteps`-long RK4 recurrence: host-side feeder of the eccentric GWB kernel (K5).
    Returns ``sepa`` (E,), ``eccen`` (E,).
    """
    assert (0.0 <= eccen_init) and (eccen_init < 1.0)
    eccen = np.zeros(nsteps)
    eccen[0] = eccen_init
    sepa_max = sepa_init
    sepa_coal = utils.schwarzschild_radius(sam.mtot) * 3
    sepa_min = sepa_coal.min()
    sepa = np.logspace(*np.log10([sepa_max, sepa_min]), nsteps)
    for step in range(1, nsteps):
        a0 = sepa[step-1]
        a1 = sepa[step]
        da = (a1 - a0)
        e0 = eccen[step-1]
        _, e1 = utils.rk4_step(holo.hardening.Hard_GW.deda, x0=a0, y0=e0, dx=da)
        e1 = np.clip(e1, 0.0, None)
        eccen[step] = e1
    return sepa, eccen
