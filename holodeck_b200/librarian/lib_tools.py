"""Parameters, parameter spaces and the per-sample model driver (``holodeck/librarian/lib_tools.py``).

* :class:`_Param_Space` (``lib_tools.py:34-465``): Latin-hypercube parameter samples -> ``(sam, hard)``
* :class:`_Param_Dist` and the ``PD_*`` distributions (``lib_tools.py:467-712``)
* :func:`run_model` (``lib_tools.py:714-842``): one library sample = one pass of the SAM hot path
  (K0/K1/K2 once, then the loudest split K4 and the GWB K3 with *independent* draws, as in the
  reference).  Everything stays on the device between stages.
"""
import abc
from pathlib import Path

import numpy as np
import scipy as sp
import scipy.stats   # noqa

import holodeck_b200 as holo
from holodeck_b200 import utils, cosmo, _lib
from holodeck_b200.constants import YR
from holodeck_b200.librarian import (
    DEF_NUM_FBINS, DEF_NUM_LOUDEST, DEF_NUM_REALS, DEF_PTA_DUR, PSPACE_FILE_SUFFIX, FNAME_LIBRARY_SIM_FILE,
)

#: parameter names the reference refuses (name, message) and renames (old -> (new, converter or None)),
#: ``lib_tools.py:25-31``
PARAM_NAMES__ERROR = []
PARAM_NAMES_REPLACE = {
    "gsmf_phi0": ["gsmf_phi0_log10", None],
}


def _canonical(name, value=None, log=None, convert=False):
    """Apply the reference's rename / refuse rules to one parameter name (and, with ``convert``, its value)."""
    for bad, msg in PARAM_NAMES__ERROR:
        if bad == name:
            err = f"Found '{name}' in parameters: {msg}"
            if log is not None:
                log.exception(err)
            raise ValueError(err)
    if name in PARAM_NAMES_REPLACE:
        new_name, func = PARAM_NAMES_REPLACE[name]
        if log is not None:
            log.error(f"Found '{name}' in parameters, should be '{new_name}'!")
        if func is not None:
            if not convert:
                err = f"CANNOT replace '{name}' ==> '{new_name}'!"
                if log is not None:
                    log.exception(err)
                raise ValueError(err)
            value = func(value)
        name = new_name
    return name, value


# ==================================================================================================
# Parameter distributions: maps of the unit interval onto parameter values (``lib_tools.py:467-712``)
# ==================================================================================================

class _Param_Dist(abc.ABC):
    """A named one-dimensional distribution, evaluated through its quantile function on [0, 1]."""

    def __init__(self, name, default=None, clip=None):
        if clip is not None and len(clip) != 2:
            raise AssertionError("`clip` must be (lo, hi)")
        self._name, self._default, self._clip = name, default, clip

    @abc.abstractmethod
    def _dist_func(self, xx):
        """quantile function"""

    def __call__(self, xx):
        vals = self._dist_func(xx)
        return vals if self._clip is None else np.clip(vals, *self._clip)

    name = property(lambda self: self._name)
    extrema = property(lambda self: self(np.asarray([0.0, 1.0])))
    default = property(lambda self: self(0.5) if self._default is None else self._default)


class PD_Uniform(_Param_Dist):
    def __init__(self, name, lo, hi, **kwargs):
        super().__init__(name, **kwargs)
        self._lo, self._hi = lo, hi

    def _dist_func(self, xx):
        return self._lo + (self._hi - self._lo) * xx


class PD_Uniform_Log(_Param_Dist):
    def __init__(self, name, lo, hi, **kwargs):
        super().__init__(name, **kwargs)
        assert lo > 0.0 and hi > 0.0
        self._lo_log10, self._hi_log10 = np.log10(lo), np.log10(hi)

    def _dist_func(self, xx):
        return np.power(10.0, self._lo_log10 + (self._hi_log10 - self._lo_log10) * xx)


class PD_Normal(_Param_Dist):
    def __init__(self, name, mean, stdev, clip=None, **kwargs):
        assert stdev > 0.0
        super().__init__(name, clip=clip, **kwargs)
        self._mean, self._stdev = mean, stdev
        self._frozen_dist = sp.stats.norm(loc=mean, scale=stdev)

    def _dist_func(self, xx):
        return self._frozen_dist.ppf(xx)


# ==================================================================================================
# Parameter spaces (``lib_tools.py:34-465``)
# ==================================================================================================

class _Param_Space(abc.ABC):
    """A library's parameter space: named distributions, their Latin-hypercube samples, and the recipe that turns
    one sample into a ``(sam, hard)`` pair.  Same attributes / methods / file format as the reference's class;
    subclasses give ``DEFAULTS`` and the two builders ``_init_sam(sam_shape, settings)``, ``_init_hard(sam, settings)``
    (the concrete spaces of this package get them from ``librarian.recipes``)."""

    __version__ = "0.0"
    _SAVED_ATTRIBUTES = ["sam_shape", "param_names", "_uniform_samples", "param_samples", "_nsamples", "_nparameters"]
    DEFAULTS = {}

    def __init__(self, parameters, log=None, nsamples=None, sam_shape=None, seed=None, random_state=None):
        log = holo.log if log is None else log
        # the reference also seeds numpy's legacy global generator here (lib_tools.py:93-98)
        if random_state is None:
            np.random.seed(seed)
            random_state = np.random.get_state()
        else:
            np.random.set_state(random_state)
        if not (hasattr(parameters, "__len__") and len(parameters) > 0):
            log.exception("`parameters` must be a list of `_Param_Dist` subclasses!")
            raise TypeError("`parameters` must be a non-empty list of `_Param_Dist` objects")
        for par in parameters:
            if not isinstance(par, _Param_Dist):
                err = f"{getattr(par, 'name', par)}: {par} is not a `_Param_Dist` object!"
                log.exception(err)
                raise ValueError(err)
        self._log = log
        self._parameters = parameters
        self.param_names = [_canonical(par.name, log=log)[0] for par in parameters]
        self._nparameters = len(parameters)
        self._nsamples = nsamples
        self._seed, self._random_state = seed, random_state
        self.sam_shape = sam_shape
        self._uniform_samples = self.param_samples = None
        if nsamples is None:
            log.info(f"{self}: {nsamples=} - cannot generate parameter samples.")
        else:
            # (S, D) points of a strength-1 Latin hypercube, pushed through each parameter's quantile function
            self._uniform_samples = sp.stats.qmc.LatinHypercube(d=self._nparameters, strength=1, seed=seed).random(n=nsamples)
            self.param_samples = np.column_stack([par(col) for par, col in zip(parameters, self._uniform_samples.T)])

    # ---- models

    def model_for_params(self, params, sam_shape=None):
        """``(sam, hard)`` for a dict of parameter values on top of ``DEFAULTS`` (``lib_tools.py:156-212``)."""
        settings = dict(self.DEFAULTS)
        for name, value in params.items():
            name, value = _canonical(name, value, log=self._log, convert=True)
            settings[name] = value
        sam = self._init_sam(self.sam_shape if sam_shape is None else sam_shape, settings)
        return sam, self._init_hard(sam, settings)

    def model_for_sample_number(self, samp_num, sam_shape=None):
        return self.model_for_params(self.param_dict(samp_num), sam_shape)

    @classmethod
    @abc.abstractmethod
    def _init_sam(cls, sam_shape, params):
        raise NotImplementedError

    @classmethod
    @abc.abstractmethod
    def _init_hard(cls, sam, params):
        raise NotImplementedError

    # ---- samples

    def param_dict(self, samp_num):
        return dict(zip(self.param_names, self.param_samples[samp_num]))

    def default_params(self):
        return {par.name: par.default for par in self._parameters}

    def normalized_params(self, vals):
        """Parameter dict from quantiles in [0, 1]; ``None`` entries take the parameter's default."""
        if np.ndim(vals) == 0:
            vals = [vals] * self.nparameters
        assert len(vals) == self.nparameters
        assert all((vv is None) or np.isfinite(vv) for vv in vals), f"Not all `vals` are finite!  {vals}"
        return {name: (par.default if vv is None else par(vv)) for name, par, vv in zip(self.param_names, self._parameters, vals)}

    extrema = property(lambda self: np.asarray([par.extrema for par in self._parameters]))
    name = property(lambda self: type(self).__name__)
    lib_shape = property(lambda self: self.param_samples.shape)
    nsamples = property(lambda self: self._nsamples)
    nparameters = property(lambda self: self._nparameters)

    # ---- files (``<name>.pspace.npz``, the reference's keys)

    def save(self, path_output):
        path_output = Path(path_output)
        if not path_output.is_dir():
            err = f"save path {path_output} does not exist, or is not a directory!"
            self._log.exception(err)
            raise ValueError(err)
        fname = path_output / f"{self.name}{PSPACE_FILE_SUFFIX}"
        np.savez(fname, class_name=self.name, class_vers=self.__version__, librarian_version=holo.librarian.__version__,
                 **{key: getattr(self, key) for key in self._SAVED_ATTRIBUTES})
        return fname

    @classmethod
    def from_save(cls, fname, log=None):
        log = holo.log if log is None else log
        data = np.load(fname, allow_pickle=True)
        space_class = holo.librarian.param_spaces_dict.get(str(data['class_name'][()]))
        if space_class is None:
            log.warning(f"pspace file {fname} has class_name={data['class_name'][()]!r}, not in `param_spaces_dict`!")
            space_class = cls
        samples = data['param_samples'][()]
        nsamples = None if samples is None else samples.shape[0]
        space = space_class(nsamples=nsamples, log=log)
        if list(data['param_names']) != list(space.param_names):
            err = f"Mismatch between loaded parameter names ({data['param_names']}) and class parameter names ({space.param_names})!"
            log.exception(err)
            raise RuntimeError(err)
        fallback = {'_nsamples': nsamples, '_nparameters': None if samples is None else samples.shape[1]}
        for key in space._SAVED_ATTRIBUTES:
            setattr(space, key, data[key][()] if key in data.files else fallback[key])
        return space


def run_model(
    sam, hard,
    pta_dur=DEF_PTA_DUR, nfreqs=DEF_NUM_FBINS, nreals=DEF_NUM_REALS, nloudest=DEF_NUM_LOUDEST,
    gwb_flag=True, singles_flag=True, details_flag=False, params_flag=False, log=None, *, seed=None, device=False,
    deferred=None,
):
    """Run the given SAM + hardening model to produce GW signals (``lib_tools.py:714-842``).

    Returns a dict with ``fobs_cents, fobs_edges`` and, depending on the flags, ``hc_ss (F,R,L)``,
    ``hc_bg (F,R)``, ``sspar (4,F,R,L)``, ``bgpar (7,F,R)``, ``gwb (F,R)``.  As in the reference the
    single-source split and the GWB use independent Poisson draws of the same number grid.
    ``details_flag`` adds ``static_binary_density, number, redz_final, gwb_params, num_params, gwb_mtot_redz_final,
    num_mtot_redz_final`` (``_calc_model_details``, K7).  Keyword-only additions: ``seed``; ``device=True`` leaves
    ``hc_ss, hc_bg, sspar, bgpar, gwb`` on the GPU as CUDA tensors (the streaming library writer copies them out
    asynchronously, ``librarian/stream.py``); ``deferred`` (a ``single_sources.DeferredChecks``, with ``device=True``)
    postpones every host-side look at a device flag, so the call returns with its kernels still in flight.
    """
    from holodeck_b200.sams import sam_cyutils
    from holodeck_b200 import gravwaves, single_sources

    if not any([gwb_flag, details_flag, singles_flag, params_flag]):
        err = f"No flags set!  {gwb_flag=} {details_flag=} {singles_flag=} {params_flag=}"
        if log is not None:
            log.exception(err)
        raise RuntimeError(err)
    data = {}
    fobs_cents, fobs_edges = utils.pta_freqs(dur=pta_dur*YR, num=nfreqs)
    # convert from GW to orbital frequencies
    fobs_orb_cents = fobs_cents / 2.0
    fobs_orb_edges = fobs_edges / 2.0
    data['fobs_cents'] = fobs_cents
    data['fobs_edges'] = fobs_edges

    if not isinstance(hard, (holo.hardening.Fixed_Time_2PL_SAM, holo.hardening.Hard_GW)):
        err = f"`holo.hardening.Fixed_Time_2PL_SAM` must be used here!  Not {hard}!"
        if log is not None:
            log.exception(err)
        raise RuntimeError(err)

    redz_final, diff_num = sam_cyutils.dynamic_binary_number_at_fobs(fobs_orb_cents, sam, hard, cosmo, device=True)
    use_redz = redz_final
    edges = [sam.mtot, sam.mrat, sam.redz, fobs_orb_edges]
    # K2 + K2b in one pass: `number` and the strain (plus the params arrays when needed)
    strain = gravwaves._char_strain_sq(edges, use_redz, params=bool(params_flag), dnum=diff_num)
    number = strain["number"]
    sub = None if seed is None else np.random.SeedSequence(seed).generate_state(2, dtype=np.uint64)
    if details_flag:
        data['static_binary_density'] = sam.static_binary_density
        data['number'] = _lib.to_host(number)
        data['redz_final'] = _lib.to_host(redz_final)
        gwb_pars, num_pars, gwb_mtot_redz_final, num_mtot_redz_final = _calc_model_details(
            edges, redz_final, number, _h2fdf=strain["h2fdf"])
        data['gwb_params'] = gwb_pars
        data['num_params'] = num_pars
        data['gwb_mtot_redz_final'] = gwb_mtot_redz_final
        data['num_mtot_redz_final'] = num_mtot_redz_final

    # calculate single sources and/or binary parameters
    fused_gwb = None
    if singles_flag or params_flag:
        nloudest = nloudest if singles_flag else 1
        # the reference draws the GWB below independently of the loudest split (lib_tools.py:801-832); here the
        # same pass of the realization kernel produces both, with independent Philox streams and seeds
        vals = single_sources.ss_gws_redz(
            edges, use_redz, number, realize=nreals, loudest=nloudest, params=params_flag,
            seed=None if sub is None else int(sub[0]), _precomputed=strain, device=bool(device),
            _deferred=deferred if device else None,
            _gwb=(nreals, None if sub is None else int(sub[1])) if gwb_flag else None,
        )
        if gwb_flag:
            vals, fused_gwb = vals[:-1], vals[-1]
        if params_flag:
            hc_ss, hc_bg, sspar, bgpar = vals
            data['sspar'] = sspar
            data['bgpar'] = bgpar
        else:
            hc_ss, hc_bg = vals
        if singles_flag:
            data['hc_ss'] = hc_ss
            data['hc_bg'] = hc_bg

    if gwb_flag:
        if fused_gwb is not None:
            gwb = fused_gwb
        else:
            gwb = gravwaves._gws_from_hc2(strain["h2fdf"], number, nreals, True,
                                          None if sub is None else int(sub[1]), 0, bool(device))
        data['gwb'] = gwb

    return data


def _calc_model_details(edges, redz_final, number, _h2fdf=None):
    """Derived properties of a population (``lib_tools.py:845-943``): strain-weighted and number-weighted
    marginals over (mtot, mrat, redz) and their distributions over the *final* redshift.

    Returns ``gwb_pars`` = [(M-1,Z-1,F), (Q-1,Z-1,F), (Z-1,F), (Z-1,F)], ``num_pars`` (same shapes),
    ``gwb_mtot_redz_final (M-1,Z-1,F)``, ``num_mtot_redz_final (M-1,Z-1,F)`` as numpy arrays.  The marginals are
    axis sums on the device; the ``2 F (1 + (M-1))`` ``scipy.stats.binned_statistic`` calls of the reference
    are one histogram kernel (K7, ``holo_model_details_hist``).
    """
    import torch
    from holodeck_b200 import gravwaves
    lib = _lib.require_gpu()
    redz = np.asarray(edges[2], dtype=np.float64)
    rzf = _lib.to_dev(redz_final)
    num = _lib.to_dev(number)
    M, Q, Z, F = (int(ss) for ss in rzf.shape)
    assert tuple(num.shape) == (M - 1, Q - 1, Z - 1, F)
    # (M-1, Q-1, Z-1, F) characteristic-strain squared for each bin
    hc2 = _h2fdf if _h2fdf is not None else gravwaves._char_strain_sq(edges, rzf, params=False)["h2fdf"]
    hc2_num = hc2 * num                                    # strain-squared weighted number of binaries
    denom = torch.sum(hc2_num, dim=(0, 1, 2))             # (F,) total GWB in each frequency bin
    gwb_pars, num_pars = [], []
    for margins in ((1,), (0,), (0, 1)):                  # lib_tools.py:876-900
        gwb_pars.append(_lib.to_host(torch.sum(hc2_num, dim=margins) / denom))
        num_pars.append(_lib.to_host(torch.sum(num, dim=margins)))
    gwb_hist = _lib.empty((M - 1, Z - 1, F))
    num_hist = _lib.empty((M - 1, Z - 1, F))
    rc = lib.holo_model_details_hist(_lib.ptr(_lib.to_dev(redz)), _lib.ptr(rzf), _lib.ptr(num), _lib.ptr(hc2), M, Q, Z, F,
                                     _lib.ptr(gwb_hist), _lib.ptr(num_hist), _lib.stream())
    _lib.check(rc, "_calc_model_details")
    gwb_pars.append(_lib.to_host(torch.sum(gwb_hist, dim=0) / denom))       # all mass bins together, lib_tools.py:913-923
    num_pars.append(_lib.to_host(torch.sum(num_hist, dim=0)))
    return gwb_pars, num_pars, _lib.to_host(gwb_hist / denom), _lib.to_host(num_hist)


def _get_sim_fname(path, pnum, library=True):
    """``lib_tools.py:1037-1048`` (library files only)."""
    return Path(path).joinpath(FNAME_LIBRARY_SIM_FILE.format(pnum=pnum))
