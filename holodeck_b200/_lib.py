"""ctypes binding of ``libholo_b200.so`` (the C ABI declared in ``include/holo_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` (``nvcc -gencode
arch=compute_100a,code=sm_100a``).  There is **no CPU fallback**: if the shared object is missing,
or no CUDA device is usable, every product entry point raises ``HoloNativeError``.
"""
import ctypes as C
import math
import os
from pathlib import Path

import numpy as np

GL_ORDER = 16

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libholo_b200.so"


class HoloNativeError(RuntimeError):
    """The native CUDA library is missing, failed to load, or a kernel call failed."""


class CyConsts(C.Structure):
    """``holo_cy_consts``: derived constants of ``sam_cyutils.pyx:30-42`` evaluated with libm."""
    _fields_ = [
        ("gw_dadt_sep_const", C.c_double),
        ("kepler_const_freq", C.c_double),
        ("kepler_const_sepa", C.c_double),
        ("four_pi_c_over_mpc", C.c_double),
    ]


class CosmoParams(C.Structure):
    _fields_ = [
        ("hubble_distance", C.c_double),
        ("hubble_time", C.c_double),
        ("om0", C.c_double),
        ("age_universe", C.c_double),
        ("gl_x", C.c_double * GL_ORDER),
        ("gl_w", C.c_double * GL_ORDER),
    ]


class SamParams(C.Structure):
    _fields_ = [
        ("gsmf_kind", C.c_int),
        ("use_gmr", C.c_int),
        ("has_gmt", C.c_int),
        ("gsmf_uses_mtot", C.c_int),
        ("gpf_uses_mtot", C.c_int),
        ("gmt_uses_mtot", C.c_int),
        ("bf_kind", C.c_int),
        ("bf_n", C.c_int),
        ("gsmf", C.c_double * 12),
        ("gpf", C.c_double * 6),
        ("gmt", C.c_double * 5),
        ("gmr", C.c_double * 11),
        ("mmb", C.c_double * 4),
        ("hubble_time", C.c_double),
        ("om0", C.c_double),
        ("age_universe", C.c_double),
        ("bf", C.c_double * 4),
    ]


class LoudestArgs(C.Structure):
    _fields_ = [
        ("variant", C.c_int),
        ("Mb", C.c_int), ("Qb", C.c_int), ("Zb", C.c_int), ("F", C.c_int),
        ("R", C.c_int), ("L", C.c_int),
        ("r0", C.c_int64),
        ("seed", C.c_uint64),
        ("normal_threshold", C.c_double),
        ("number", C.c_void_p),
        ("h2fdf", C.c_void_p),
        ("order", C.c_void_p),
        ("mt", C.c_void_p),
        ("mr", C.c_void_p),
        ("rz", C.c_void_p),
        ("redz_final", C.c_void_p),
        ("dcom_final", C.c_void_p),
        ("sepa", C.c_void_p),
        ("angs", C.c_void_p),
        ("counts", C.c_void_p),
        ("hc2ss", C.c_void_p),
        ("hc2bg", C.c_void_p),
        ("sspar", C.c_void_p),
        ("bgpar", C.c_void_p),
        ("lspar", C.c_void_p),
        ("ssidx", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_int64),
        ("bucket_cap", C.c_int),
        ("head_margin", C.c_double),
        ("gwb", C.c_void_p),
        ("gwb_R", C.c_int),
        ("gwb_r0", C.c_int64),
        ("gwb_seed", C.c_uint64),
        ("defer_check", C.c_int),
    ]


def cy_consts():
    """Evaluate the reference's module-level Cython constants exactly as it does (libm, same order).

    ``sam_cyutils.pyx:30-42``: ``GW_DADT_SEP_CONST = - 64.0 * pow(MY_NWTG, 3) / 5.0 / pow(MY_SPLC, 5)``
    etc.  ``math.pow`` / ``math.sqrt`` are the same glibc routines Cython's ``libc.math`` binds.
    """
    nwtg = 6.6742999e-08
    splc = 29979245800.0
    mpc = 3.08567758e+24
    cc = CyConsts()
    cc.gw_dadt_sep_const = - 64.0 * math.pow(nwtg, 3) / 5.0 / math.pow(splc, 5)
    cc.kepler_const_freq = (1.0 / (2.0*math.pi)) * math.sqrt(nwtg)
    cc.kepler_const_sepa = math.pow(nwtg, 1.0/3.0) / math.pow(2.0*math.pi, 2.0/3.0)
    cc.four_pi_c_over_mpc = 4 * math.pi * splc / mpc
    return cc


def cy_gw_src_const():
    """``GW_SRC_CONST`` of ``cyutils.pyx:48`` evaluated with libm in the same order."""
    nwtg = 6.6742999e-08
    splc = 29979245800.0
    return 8.0 * math.pow(nwtg, 5.0/3.0) * math.pow(math.pi, 2.0/3.0) / math.sqrt(10.0) / math.pow(splc, 4.0)


def cosmo_params(cosmo):
    from .cosmology import _GL_X, _GL_W
    cp = CosmoParams()
    cp.hubble_distance = cosmo.hubble_distance
    cp.hubble_time = cosmo.hubble_time
    cp.om0 = cosmo.Om0
    cp.age_universe = cosmo.age_universe
    for ii in range(GL_ORDER):
        cp.gl_x[ii] = _GL_X[ii]
        cp.gl_w[ii] = _GL_W[ii]
    return cp


DC_TABLE_N = 2048          # intervals of the comoving-distance table
DC_TABLE_ZMAX = 20.0       # beyond it the kernels fall back to the quadrature
_DC_TABLES = {}


def dc_table_host(om0, n=DC_TABLE_N, zmax=DC_TABLE_ZMAX):
    """Nodes ``(R_i, h R'_i)`` of ``R(w) = (1/w) int_{1-w}^{1} g(s) ds``, ``g = 2 / sqrt(Om0 + (1-Om0) s^6)``, at
    ``w_i = i h`` (``include/holo_b200.h``, `dc_table`): flat LCDM without radiation, the same integrand as
    ``cosmology.comoving_distance`` (``gravwaves.py:718`` via astropy/cosmopy in the reference).  The node integrals are
    running sums of 16-point Gauss-Legendre panels, one per interval (error << 1e-15)."""
    from .cosmology import _GL_X, _GL_W
    ol = 1.0 - om0
    wmax = 1.0 - 1.0 / math.sqrt(1.0 + zmax)
    hh = wmax / n

    def gg(ss):
        return 2.0 / np.sqrt(om0 + ol * ss**6)
    wn = np.arange(n + 1) * hh
    mid = 0.5 * (wn[1:] + wn[:-1])
    vv = mid[:, None] + 0.5 * hh * np.asarray(_GL_X)[None, :]
    panels = 0.5 * hh * np.sum(np.asarray(_GL_W)[None, :] * gg(1.0 - vv), axis=1)
    ff = np.concatenate([[0.0], np.cumsum(panels)])
    rr = np.empty(n + 1)
    rp = np.empty(n + 1)
    rr[1:] = ff[1:] / wn[1:]
    rr[0] = gg(1.0)
    rp[1:] = (gg(1.0 - wn[1:]) - rr[1:]) / wn[1:]
    rp[0] = 3.0 * ol                               # R'(0) = -g'(1)/2,  g'(s) = -6 OL s^5 (Om0 + OL s^6)^(-3/2)
    return np.ascontiguousarray(np.stack([rr, hh * rp], axis=1)), n, wmax


def dc_table(cosmo):
    """The table on the current device, built once per (Om0, device)."""
    import torch
    key = (float(cosmo.Om0), torch.cuda.current_device())
    if key not in _DC_TABLES:
        tab, n, wmax = dc_table_host(float(cosmo.Om0))
        _DC_TABLES[key] = (to_dev(tab), n, wmax)
    return _DC_TABLES[key]


_P = C.c_void_p
_D = C.c_double
_I = C.c_int
_L = C.c_int64
_U = C.c_uint64

#: name -> argtypes ; every exported function returns int unless listed in `_RESTYPES`
SIGNATURES = {
    "holo_abi_version": [],
    "holo_last_error": [],
    "holo_device_count": [],
    "holo_launch_count": [],
    "holo_set_profiling": [_I],
    "holo_get_profile": [_P, _I],
    "holo_sam_density": [_P, _P, _P, _P, _P, _I, _I, _I, C.POINTER(SamParams), _P, _P, _P, _P, _P],
    "holo_zero_stalled": [_P, _P, _L, _P],
    "holo_find_2pwl_hardening_norm": [CyConsts, _D, _P, _P, _I, _D, _D, _D, _D, _I, _P, _P],
    "holo_binary_lifetime_2pwl": [CyConsts, _P, _P, _P, _I, _D, _D, _D, _D, _I, _P, _P],
    "holo_hard_func_2pwl_gw": [CyConsts, _P, _P, _P, _P, _D, _D, _D, _L, _P, _P],
    "holo_dbn_2pwl": [CyConsts, _P, _I, _D, _I, _P, _D, _D, _D, _P, _P, _P, _P, _P, _I, _I, _I,
                      _P, _P, _P, _I, _P, _P, _P],
    "holo_dbn_gw": [CyConsts, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _I, _P, _P, _P],
    "holo_integrate_differential_number_3dx1d": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "holo_char_strain_sq": [C.POINTER(CosmoParams), _D, _D, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I,
                            _P, _P, _P, _P, _P, _P, _P, _I, _D, _P],
    "holo_integrate_and_strain": [C.POINTER(CosmoParams), _D, _D, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                  _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _D, _P],
    "holo_gwb_expectation_workspace_bytes": [_I],
    "holo_gwb_expectation": [_P, _P, _L, _I, _P, _P, _L, _P],
    "holo_sam_poisson_gwb": [_P, _P, _L, _I, _I, _L, _U, _D, _P, _P, _P, _L, _P],
    "holo_loudest_workspace_bytes": [_I, _L, _I, _I, _I, _I],
    "holo_loudest": [C.POINTER(LoudestArgs), _P],
    "holo_ss_bg_hc": [_P, _P, _I, _I, _I, _I, _I, _L, _U, _D, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                      _P, _L, _P],
    "holo_realize_workspace_bytes": [_I, _L, _I, _I],
    "holo_poisson_as_needed": [_P, _L, _U, _U, _D, _P, _P],
    "holo_sam_calc_gwb_single_eccen": [CyConsts, _D, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I,
                                       _L, _U, _P, _P, _L, _P],
    "holo_eccen_workspace_bytes": [_I, _I, _I, _I, _I, _I],
    "holo_scatter_step_bytes": [],
    "holo_scatter_gradients": [_I, _I, _P, _I, _P, _I, _D, _P, _P, _P],
    "holo_scatter_ct_eval": [_L, _I, _P, _P, _P, _P, _P, _P],
    "holo_scatter_bilinear": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    "holo_scatter_geo_bytes": [],
    "holo_model_details_hist": [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P],
}
_RESTYPES = {
    "holo_last_error": C.c_char_p,
    "holo_launch_count": C.c_int64,
    "holo_loudest_workspace_bytes": C.c_int64,
    "holo_realize_workspace_bytes": C.c_int64,
    "holo_eccen_workspace_bytes": C.c_int64,
    "holo_gwb_expectation_workspace_bytes": C.c_int64,
}

_lib = None


def load(path=None):
    """Open the shared library and bind every symbol of ``include/holo_b200.h`` (no GPU needed)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = Path(path) if path is not None else Path(os.environ.get("HOLO_B200_LIB", LIB_PATH))
    if not path.exists():
        raise HoloNativeError(
            f"native library {path} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)"
        )
    try:
        lib = C.CDLL(str(path))
    except OSError as err:
        raise HoloNativeError(f"could not load {path}: {err}") from err
    for name, argtypes in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as err:
            raise HoloNativeError(f"{path} does not export `{name}`") from err
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.holo_abi_version() != 3:
        raise HoloNativeError(f"ABI mismatch: library reports version {lib.holo_abi_version()}")
    _lib = lib
    return lib


def check(rc, what=""):
    """Translate a non-zero status into the exception type the reference would raise."""
    if rc == 0:
        return
    msg = load().holo_last_error().decode()
    if rc == 1:
        raise ValueError(f"{what}: {msg}")
    raise HoloNativeError(f"{what}: status {rc}: {msg}")


def require_gpu():
    """Fail loudly when there is no usable CUDA device (the product path has no CPU fallback)."""
    import torch
    if not torch.cuda.is_available():
        raise HoloNativeError("no CUDA device available: holodeck_b200 has no CPU fallback")
    lib = load()
    if lib.holo_device_count() <= 0:
        raise HoloNativeError("libholo_b200 sees no CUDA device: " + lib.holo_last_error().decode())
    return lib


# ---- device-buffer plumbing (torch tensors are only used as owning handles for HBM) -------------

def device(device=None):
    import torch
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def to_dev(arr, dtype=None, dev=None):
    """numpy array / torch tensor -> contiguous CUDA tensor of the given dtype (default float64)."""
    import torch
    dtype = torch.float64 if dtype is None else dtype
    if isinstance(arr, torch.Tensor):
        tt = arr
        if not tt.is_cuda:
            TRAFFIC["h2d"] += tt.numel() * tt.element_size()
            tt = tt.to(device(dev), non_blocking=True)
        return tt.to(dtype).contiguous()
    arr = np.ascontiguousarray(arr)
    tt = torch.from_numpy(arr)
    TRAFFIC["h2d"] += arr.nbytes
    return tt.to(device=device(dev), dtype=dtype, non_blocking=True).contiguous()


def empty(shape, dtype=None, dev=None):
    import torch
    return torch.empty(shape, dtype=torch.float64 if dtype is None else dtype, device=device(dev))


#: bytes moved across PCIe by the shims (bench.py reads these for `h2d_bytes_per_step` / `d2h_bytes_per_step`)
TRAFFIC = {"h2d": 0, "d2h": 0}


#: device-to-host copies of at least this many bytes land in page-locked memory (torch's caching host allocator
#: recycles the blocks): the drop-in shims hand 230 MB grids back to numpy callers, and a pageable read-back runs
#: at a fraction of the PCIe rate
PINNED_D2H_MIN_BYTES = 1 << 20


def to_host(tensor):
    """CUDA tensor -> numpy array (counted device-to-host copy)."""
    import torch
    nbytes = tensor.numel() * tensor.element_size()
    TRAFFIC["d2h"] += nbytes
    if nbytes >= PINNED_D2H_MIN_BYTES and tensor.is_cuda:
        src = tensor.contiguous()
        out = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
        out.copy_(src, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out.numpy()             # (the array keeps the pinned block alive; it returns to the cache with it)
    return tensor.cpu().numpy()


def ptr(tensor):
    return C.c_void_p(0) if tensor is None else C.c_void_p(tensor.data_ptr())


def stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def is_device_array(xx):
    try:
        import torch
    except ImportError:   # pragma: no cover
        return False
    return isinstance(xx, torch.Tensor) and xx.is_cuda
