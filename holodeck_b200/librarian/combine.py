"""Combine the per-sample files of a library into one file (``holodeck/librarian/combine.py:86-468``).

Same on-disk contract as the reference: per-sample ``library_sims/library__pNNNNNN.npz`` files with keys
``fobs_cents, fobs_edges, gwb, hc_ss, hc_bg, sspar, bgpar`` (or a single ``fail`` key) are merged into the
datasets ``fobs_cents (F,)``, ``fobs_edges (F+1,)``, ``sample_params (S, P)``, ``gwb (S, F, R)``,
``hc_ss (S, F, R, L)``, ``hc_bg (S, F, R)``, ``sspar (S, 4, F, R, L)``, ``bgpar (S, 7, F, R)`` plus the attributes
``param_names, parameter_space_class_name, holodeck_version, holodeck_git_hash, holodeck_librarian_version``.
Failed samples become NaN rows, and their parameters NaN (``combine.py:204-205, 404-417``).

The reference writes hdf5 through h5py.  h5py is an optional dependency here: when it is importable the
output is ``sam-library.hdf5`` exactly as the reference's; otherwise the same datasets (attributes as extra
0-d / string arrays) go into ``sam-library.npz``.  Pure host I/O: no kernels involved.
"""
from pathlib import Path

import numpy as np

import holodeck_b200 as holo
from holodeck_b200.librarian import DIRNAME_LIBRARY_SIMS, PSPACE_FILE_SUFFIX, lib_tools

FNAME_LIBRARY_COMBINED_FILE = "sam-library"    # do NOT include file suffix


class DomainNotLibraryError(Exception):
    def __init__(self, message="This looks like a 'domain' not a 'library'!"):
        self.message = message
        super().__init__(self.message)


def _have_h5py():
    try:
        import h5py   # noqa: F401
        return True
    except ImportError:
        return False


def get_sam_lib_fname(path, gwb_only, library=True):
    """``lib_tools.get_sam_lib_fname`` ``lib_tools.py:1029-1040`` (suffix .npz when h5py is unavailable)."""
    if not library:
        raise NotImplementedError("'domain' explorations are outside the hot path (SURVEY section 8)")
    fname = FNAME_LIBRARY_COMBINED_FILE
    if gwb_only:
        fname += "_gwb-only"
    return Path(path).joinpath(fname).with_suffix(".hdf5" if _have_h5py() else ".npz")


def load_pspace_from_path(path, space_class=None, log=None):
    """``lib_tools.load_pspace_from_path`` ``lib_tools.py:946-1008``."""
    path = Path(path).absolute().resolve()
    if not path.exists():
        raise RuntimeError(f"path {path} does not exist!")
    if path.is_dir():
        pattern = "*" + PSPACE_FILE_SUFFIX
        space_fname = list(path.glob(pattern))
        if len(space_fname) != 1:
            raise FileNotFoundError(f"found {len(space_fname)} matches to {pattern} in output {path}!")
        space_fname = space_fname[0]
    else:
        space_fname = path
    if space_class is None:
        try:
            space_class = holo.librarian.param_spaces_dict[str(np.load(space_fname, allow_pickle=True)['class_name'])]
        except Exception as err:   # noqa: BLE001  (as the reference: fall back to the file name)
            if log is not None:
                log.error(f"Could not load `class_name` from save file '{space_fname}'.")
                log.error(str(err))
            space_class = holo.librarian.param_spaces_dict[space_fname.name.split(".")[0]]
    return space_class.from_save(space_fname, log=log), space_fname


def sam_lib_combine(path_output, log=None, path_pspace=None, recreate=False, gwb_only=False, library=True):
    """Combine individual simulation files into a single library file; returns its path (``combine.py:86-272``)."""
    log = holo.log if log is None else log
    if not library:
        raise NotImplementedError("'domain' explorations are outside the hot path (SURVEY section 8)")
    path_output = Path(path_output)
    log.info(f"Path output = {path_output}")
    path_sims = path_output.joinpath(DIRNAME_LIBRARY_SIMS)

    lib_path = get_sam_lib_fname(path_output, gwb_only, library=library)
    if lib_path.exists():
        log.warning(f"combined library already exists: {lib_path}, run with `recreate` to recreate.")
        if not recreate:
            return None
        log.warning("re-combining data into new file")

    if path_pspace is None:
        path_pspace = path_output
    from holodeck_b200.librarian import stream
    if stream.LibraryStore.exists(path_output):
        return _combine_from_store(path_output, path_pspace, lib_path, gwb_only, log)
    pspace, pspace_fname = load_pspace_from_path(path_pspace, log=log)
    log.info(f"loaded param space: {pspace} from '{pspace_fname}'")
    param_names = pspace.param_names
    param_samples = pspace.param_samples[()]
    if param_samples is None:
        raise DomainNotLibraryError(f"`library` is True, but {path_output} looks like it's a domain.")
    param_samples = np.array(param_samples, dtype=np.float64)
    nsamp_all, ndim = param_samples.shape

    log.info(f"checking that all {nsamp_all} files exist")
    fobs_cents, fobs_edges, nreals, nloudest, has_gwb, has_ss, has_params = _check_files_and_load_shapes(
        log, path_sims, nsamp_all, library)
    if not has_gwb and gwb_only:
        err = f"Combining with {gwb_only=}, but received {has_gwb=} from `_check_files_and_load_shapes`!"
        log.exception(err)
        raise RuntimeError(err)
    if (fobs_cents is None) or (nreals is None):
        err = f"After checking files, {fobs_cents=} and {nreals=}!"
        log.exception(err)
        raise ValueError(err)
    nfreqs = fobs_cents.size

    gwb = np.zeros((nsamp_all, nfreqs, nreals)) if has_gwb else None
    hc_ss = hc_bg = sspar = bgpar = None
    if (not gwb_only) and has_ss:
        hc_ss = np.zeros((nsamp_all, nfreqs, nreals, nloudest))
        hc_bg = np.zeros((nsamp_all, nfreqs, nreals))
    if (not gwb_only) and has_params:
        sspar = np.zeros((nsamp_all, 4, nfreqs, nreals, nloudest))
        bgpar = np.zeros((nsamp_all, 7, nfreqs, nreals))
    gwb, hc_ss, hc_bg, sspar, bgpar, param_samples, bad_files = _load_library_from_all_files(
        path_sims, gwb, hc_ss, hc_bg, sspar, bgpar, param_samples, log, library)
    param_samples[bad_files] = np.nan

    datasets = dict(fobs_cents=fobs_cents, fobs_edges=fobs_edges, sample_params=param_samples)
    if gwb is not None:
        datasets['gwb'] = gwb
    if not gwb_only:
        if has_ss:
            datasets['hc_ss'] = hc_ss
            datasets['hc_bg'] = hc_bg
        if has_params:
            datasets['sspar'] = sspar
            datasets['bgpar'] = bgpar
    attrs = dict(param_names=np.array(param_names).astype('S'), parameter_space_class_name=pspace.name,
                 holodeck_version=holo.__version__, holodeck_git_hash="None",
                 holodeck_librarian_version=holo.librarian.__version__)
    _write_library(lib_path, datasets, attrs, log)
    assert np.all(fobs_cents > 0.0)
    return lib_path


def _write_library(lib_path, datasets, attrs, log):
    log.info(f"Writing collected data to file {lib_path}")
    if lib_path.suffix == ".hdf5":
        import h5py
        with h5py.File(lib_path, 'w') as h5:
            for key, val in datasets.items():
                h5.create_dataset(key, data=val)
            for key, val in attrs.items():
                h5.attrs[key] = val
    else:
        np.savez(lib_path, **datasets, **{f"attrs/{key}": np.asarray(val) for key, val in attrs.items()})


def _combine_from_store(path_output, path_pspace, lib_path, gwb_only, log):
    """The library was generated through the streaming file plane (``librarian/stream.py``): the combined arrays
    already exist as memory maps, rows of failed samples are NaN.  Same datasets / attributes / checks as the
    per-sample-file path above (``combine.py:86-272``)."""
    from holodeck_b200.librarian import stream
    store = stream.LibraryStore.open(path_output, mode="r")
    pspace, pspace_fname = load_pspace_from_path(path_pspace, log=log)
    log.info(f"loaded param space: {pspace} from '{pspace_fname}'")
    param_samples = pspace.param_samples[()]
    if param_samples is None:
        raise DomainNotLibraryError(f"`library` is True, but {path_output} looks like it's a domain.")
    param_samples = np.array(param_samples, dtype=np.float64)
    status = np.asarray(store.status)
    todo = np.flatnonzero(status == stream.STATUS_TODO)
    if todo.size:
        err = f"Missing at least sample number {int(todo[0])} out of {status.size} samples!  ({todo.size} not run yet)"
        log.exception(err)
        raise ValueError(err)
    bad = status == stream.STATUS_FAIL
    log.info(f"{int(bad.sum())}/{bad.size} files are failures")
    param_samples[bad] = np.nan
    has_gwb = 'gwb' in store.maps
    if gwb_only and not has_gwb:
        raise RuntimeError(f"Combining with {gwb_only=}, but the library holds no `gwb`!")
    fobs_cents, fobs_edges = store.fobs
    datasets = dict(fobs_cents=fobs_cents, fobs_edges=fobs_edges, sample_params=param_samples)
    for key in ('gwb', 'hc_ss', 'hc_bg', 'sspar', 'bgpar'):
        if key in store.maps and (key == 'gwb' or not gwb_only):
            datasets[key] = store.maps[key]
    attrs = dict(param_names=np.array(pspace.param_names).astype('S'), parameter_space_class_name=pspace.name,
                 holodeck_version=holo.__version__, holodeck_git_hash="None",
                 holodeck_librarian_version=holo.librarian.__version__)
    _write_library(lib_path, datasets, attrs, log)
    assert np.all(fobs_cents > 0.0)
    return lib_path


def _check_files_and_load_shapes(log, path_sims, nsamp, library):
    """All `nsamp` files must exist; array shapes come from the first usable one (``combine.py:272-363``)."""
    fobs_edges = fobs_cents = nreals = nloudest = None
    has_gwb = has_ss = has_params = False
    log.info(f"Checking {nsamp} files in {path_sims}")
    for ii in range(nsamp):
        temp_fname = lib_tools._get_sim_fname(path_sims, ii, library=library)
        if not temp_fname.exists():
            err = f"Missing at least file number {ii} out of {nsamp} files!  {temp_fname}"
            log.exception(err)
            raise ValueError(err)
        if (fobs_cents is not None) and (nreals is not None) and (nloudest is not None):
            continue
        temp = np.load(temp_fname)
        data_keys = list(temp.keys())
        if 'fail' in data_keys:
            log.error(f"File {ii=} is a failed simulation file.  {temp_fname=}: {temp['fail']}")
            continue
        if fobs_cents is None:
            if temp.get('fobs', None) is not None:
                err = "Found `fobs` in data, expected only `fobs_cents` and `fobs_edges`!"
                log.exception(err)
                raise ValueError(err)
            fobs_cents = temp['fobs_cents']
            fobs_edges = temp['fobs_edges']
        has_gwb = has_gwb or ('gwb' in data_keys)
        if (not has_ss) and ('hc_ss' in data_keys):
            assert 'hc_bg' in data_keys
            has_ss = True
        if (not has_params) and ('sspar' in data_keys):
            assert 'bgpar' in data_keys
            has_params = True
        if nreals is None:
            nreals_1 = temp['gwb'].shape[-1] if 'gwb' in data_keys else None
            nreals_2 = temp['hc_bg'].shape[-1] if 'hc_bg' in data_keys else None
            nreals = nreals_2 if nreals_2 is not None else nreals_1
            if (nreals_1 is not None) and (nreals_2 is not None):
                assert nreals_1 == nreals_2
        if (nloudest is None) and ('hc_ss' in data_keys):
            nloudest = temp['hc_ss'].shape[-1]
    return fobs_cents, fobs_edges, nreals, nloudest, has_gwb, has_ss, has_params


def _load_library_from_all_files(path_sims, gwb, hc_ss, hc_bg, sspar, bgpar, param_samples, log, library):
    """Fill the combined arrays from all files; failure files become NaN (``combine.py:366-441``)."""
    if hc_bg is not None:
        nsamp_all = hc_bg.shape[0]
    elif gwb is not None:
        nsamp_all = gwb.shape[0]
    else:
        err = "Unable to get shape from either `hc_bg` or `gwb`!"
        log.exception(err)
        raise RuntimeError(err)
    bad_files = np.zeros(nsamp_all, dtype=bool)
    for pnum in range(nsamp_all):
        fname = lib_tools._get_sim_fname(path_sims, pnum, library=library)
        temp = np.load(fname, allow_pickle=True)
        if 'fail' in temp:
            log.info(f"file {pnum=:06d} is a failure file, setting values to NaN ({fname})")
            if gwb is not None:
                gwb[pnum, :, :] = np.nan
            if hc_ss is not None:
                hc_ss[pnum, :, :, :] = np.nan
                hc_bg[pnum, :, :] = np.nan
            bad_files[pnum] = True
            continue
        if gwb is not None:
            gwb[pnum, :, :] = temp['gwb'][...]
        if hc_ss is not None:
            hc_ss[pnum, :, :, :] = temp['hc_ss'][...]
            hc_bg[pnum, :, :] = temp['hc_bg'][...]
        if bgpar is not None:
            sspar[pnum, :, :, :, :] = temp['sspar'][...]
            bgpar[pnum, :, :, :] = temp['bgpar'][...]
    log.info(f"{int(bad_files.sum())}/{bad_files.size} files are failures")
    return gwb, hc_ss, hc_bg, sspar, bgpar, param_samples, bad_files
