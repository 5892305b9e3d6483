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

PARAM_NAMES__ERROR = []
PARAM_NAMES_REPLACE = {
    "gsmf_phi0": ["gsmf_phi0_log10", None],
}


class _Param_Space(abc.ABC):
    """Base class for generating holodeck libraries.  Defines the parameter space and settings."""

    __version__ = "0.0"
    _SAVED_ATTRIBUTES = ["sam_shape", "param_names", "_uniform_samples", "param_samples", "_nsamples", "_nparameters"]
    DEFAULTS = {}

    def __init__(self, parameters, log=None, nsamples=None, sam_shape=None, seed=None, random_state=None):
        if log is None:
            log = holo.log
        log.debug(f"seed = {seed}")
        if random_state is None:
            np.random.seed(seed)
            random_state = np.random.get_state()
        else:
            np.random.set_state(random_state)

        try:
            nparameters = len(parameters)
            assert nparameters > 0
        except (TypeError, AssertionError) as err:
            log.exception("`parameters` must be a list of `_Param_Dist` subclasses!")
            raise err

        param_names = []
        for param in parameters:
            name = param.name
            if not isinstance(param, _Param_Dist):
                err = f"{name}: {param} is not a `_Param_Dist` object!"
                log.exception(err)
                raise ValueError(err)
            for pname, msg in PARAM_NAMES__ERROR:
                if pname == name:
                    err = f"Found '{name}' in parameters: {msg}"
                    log.exception(err)
                    raise ValueError(err)
            for pname, replace in PARAM_NAMES_REPLACE.items():
                if pname != name:
                    continue
                new_name, new_func = replace
                log.error(f"Found '{name}' in parameters, should be '{new_name}'!")
                if new_func is None:
                    name = new_name
                else:
                    err = f"CANNOT replace '{name}' ==> '{new_name}'!"
                    log.exception(err)
                    raise ValueError(err)
            param_names.append(name)

        if (nsamples is None) or (nparameters == 0):
            log.info(f"{self}: {nsamples=} {nparameters=} - cannot generate parameter samples.")
            uniform_samples = None
            param_samples = None
        else:
            # strength = 1 : basic latin hypercube
            lhc = sp.stats.qmc.LatinHypercube(d=nparameters, strength=1, seed=seed)
            # (S, D) samples `S` and parameters `D`
            uniform_samples = lhc.random(n=nsamples)
            param_samples = np.zeros_like(uniform_samples)
            for ii, param in enumerate(parameters):
                param_samples[:, ii] = param(uniform_samples[:, ii])

        self._log = log
        self._nparameters = nparameters
        self._nsamples = nsamples
        self._seed = seed
        self._random_state = random_state
        self.sam_shape = sam_shape
        self.param_names = param_names
        self.param_samples = param_samples
        self._parameters = parameters
        self._uniform_samples = uniform_samples

    def model_for_params(self, params, sam_shape=None):
        """Construct ``(sam, hard)`` for a dict of parameter values (``lib_tools.py:156-212``)."""
        if sam_shape is None:
            sam_shape = self.sam_shape
        settings = self.DEFAULTS.copy()
        for name, value in params.items():
            for pname, replace in PARAM_NAMES_REPLACE.items():
                if pname != name:
                    continue
                new_name, new_func = replace
                self._log.error(f"Found '{name}' in parameters, should be '{new_name}'!")
                name = new_name
                value = value if new_func is None else new_func(value)
            for pname, msg in PARAM_NAMES__ERROR:
                if pname == name:
                    err = f"Found '{name}' in parameters: {msg}"
                    self._log.exception(err)
                    raise ValueError(err)
            settings[name] = value
        sam = self._init_sam(sam_shape, settings)
        hard = self._init_hard(sam, settings)
        return sam, hard

    @classmethod
    @abc.abstractmethod
    def _init_sam(cls, sam_shape, params):
        raise

    @classmethod
    @abc.abstractmethod
    def _init_hard(cls, sam, params):
        raise

    def save(self, path_output):
        """Save the generated samples and parameter-space info into a single ``.pspace.npz`` file."""
        path_output = Path(path_output)
        if not path_output.exists() or not path_output.is_dir():
            err = f"save path {path_output} does not exist, or is not a directory!"
            self._log.exception(err)
            raise ValueError(err)
        fname = path_output.joinpath(f"{self.name}{PSPACE_FILE_SUFFIX}")
        data = {key: getattr(self, key) for key in self._SAVED_ATTRIBUTES}
        np.savez(fname, class_name=self.name, class_vers=self.__version__,
                 librarian_version=holo.librarian.__version__, **data)
        return fname

    @classmethod
    def from_save(cls, fname, log=None):
        """Create a new parameter-space instance loaded from the given save file (``lib_tools.py:257-377``)."""
        if log is None:
            log = holo.log
        data = np.load(fname, allow_pickle=True)
        class_name = data['class_name'][()]
        pspace_class = holo.librarian.param_spaces_dict.get(str(class_name), None)
        if pspace_class is None:
            log.warning(f"pspace file {fname} has {class_name=}, not found in `holo.param_spaces_dict`!")
            pspace_class = cls
        nsamples = None if data['param_samples'][()] is None else data['param_samples'].shape[0]
        space = pspace_class(nsamples=nsamples, log=log)
        param_names = data['param_names']
        if not all(pl == pc for pl, pc in zip(param_names, space.param_names)):
            err = f"Mismatch between loaded parameter names ({param_names}) and class parameter names ({space.param_names})!"
            log.exception(err)
            raise RuntimeError(err)
        for key in space._SAVED_ATTRIBUTES:
            try:
                val = data[key][()]
            except KeyError:
                if key == '_nsamples':
                    val = nsamples
                elif key == '_nparameters':
                    val = None if nsamples is None else data['param_samples'].shape[1]
                else:
                    raise
            setattr(space, key, val)
        return space

    def param_dict(self, samp_num):
        return {nn: pp for nn, pp in zip(self.param_names, self.param_samples[samp_num])}

    @property
    def extrema(self):
        return np.asarray([dd.extrema for dd in self._parameters])

    @property
    def name(self):
        return self.__class__.__name__

    @property
    def lib_shape(self):
        return self.param_samples.shape

    @property
    def nsamples(self):
        return self._nsamples

    @property
    def nparameters(self):
        return self._nparameters

    def model_for_sample_number(self, samp_num, sam_shape=None):
        params = self.param_dict(samp_num)
        self._log.debug(f"params {samp_num} :: {params}")
        return self.model_for_params(params, sam_shape)

    def normalized_params(self, vals):
        """Params dict from normalized [0,1] values (``None`` -> parameter default) (``lib_tools.py:412-449``)."""
        if np.ndim(vals) == 0:
            vals = self.nparameters * [vals]
        assert len(vals) == self.nparameters
        assert np.all([(vv is None) or np.isfinite(vv) for vv in vals]), f"Not all `vals` are finite!  {vals}"
        params = {}
        for ii, pname in enumerate(self.param_names):
            param = self._parameters[ii]
            params[pname] = param.default if vals[ii] is None else param(vals[ii])
        return params

    def default_params(self):
        return {param.name: param.default for param in self._parameters}


class _Param_Dist(abc.ABC):
    """Parameter distribution: maps [0, 1] to parameter values (``lib_tools.py:467-518``)."""

    def __init__(self, name, default=None, clip=None):
        if clip is not None:
            assert len(clip) == 2
        self._clip = clip
        self._name = name
        self._default = default

    def __call__(self, xx):
        rv = self._dist_func(xx)
        if self._clip is not None:
            rv = np.clip(rv, *self._clip)
        return rv

    @abc.abstractmethod
    def _dist_func(self, *args, **kwargs):
        pass

    @property
    def extrema(self):
        return self(np.asarray([0.0, 1.0]))

    @property
    def name(self):
        return self._name

    @property
    def default(self):
        if self._default is not None:
            return self._default
        return self(0.5)


class PD_Uniform(_Param_Dist):
    """``lib_tools.py:520-531``"""

    def __init__(self, name, lo, hi, **kwargs):
        super().__init__(name, **kwargs)
        self._lo = lo
        self._hi = hi

    def _dist_func(self, xx):
        return self._lo + (self._hi - self._lo) * xx


class PD_Uniform_Log(_Param_Dist):
    """``lib_tools.py:533-545``"""

    def __init__(self, name, lo, hi, **kwargs):
        super().__init__(name, **kwargs)
        assert lo > 0.0 and hi > 0.0
        self._lo_log10 = np.log10(lo)
        self._hi_log10 = np.log10(hi)

    def _dist_func(self, xx):
        return np.power(10.0, self._lo_log10 + (self._hi_log10 - self._lo_log10) * xx)


class PD_Normal(_Param_Dist):
    """Normal distribution mapped from [0,1] through the ppf (``lib_tools.py:547-569``)."""

    def __init__(self, name, mean, stdev, clip=None, **kwargs):
        assert stdev > 0.0
        super().__init__(name, clip=clip, **kwargs)
        self._mean = mean
        self._stdev = stdev
        self._frozen_dist = sp.stats.norm(loc=mean, scale=stdev)

    def _dist_func(self, xx):
        return self._frozen_dist.ppf(xx)


def run_model(
    sam, hard,
    pta_dur=DEF_PTA_DUR, nfreqs=DEF_NUM_FBINS, nreals=DEF_NUM_REALS, nloudest=DEF_NUM_LOUDEST,
    gwb_flag=True, singles_flag=True, details_flag=False, params_flag=False, log=None, *, seed=None, device=False,
):
    """Run the given SAM + hardening model to produce GW signals (``lib_tools.py:714-842``).

    Returns a dict with ``fobs_cents, fobs_edges`` and, depending on the flags, ``hc_ss (F,R,L)``,
    ``hc_bg (F,R)``, ``sspar (4,F,R,L)``, ``bgpar (7,F,R)``, ``gwb (F,R)``.  As in the reference the
    single-source split and the GWB use independent Poisson draws of the same number grid.
    ``details_flag`` adds ``static_binary_density, number, redz_final, gwb_params, num_params, gwb_mtot_redz_final,
    num_mtot_redz_final`` (``_calc_model_details``, K7).  Keyword-only additions: ``seed``; ``device=True`` leaves
    ``hc_ss, hc_bg, sspar, bgpar, gwb`` on the GPU as CUDA tensors (the streaming library writer copies them out
    asynchronously, ``librarian/stream.py``).
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
