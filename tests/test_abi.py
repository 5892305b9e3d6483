"""The C-ABI shared library loads without a GPU and exports every symbol include/holo_b200.h declares;
the product path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_functions():
    text = (ROOT / "include" / "holo_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+char\*|int64_t|int|void)\s+(holo_\w+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_header_symbols_are_exported_and_bound():
    from holodeck_b200 import _lib
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/holo_b200.h but not exported"
    # every declared function has a ctypes signature, and nothing undeclared is bound
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert lib.holo_abi_version() == 3
    assert isinstance(lib.holo_last_error(), bytes)
    assert lib.holo_launch_count() >= 0


def test_struct_layouts_match_the_header():
    from holodeck_b200 import _lib
    assert C.sizeof(_lib.CyConsts) == 4 * 8
    assert C.sizeof(_lib.CosmoParams) == (4 + 2 * _lib.GL_ORDER) * 8
    assert C.sizeof(_lib.SamParams) == 8 * 4 + (12 + 6 + 5 + 11 + 4 + 3 + 4) * 8      # + bf[4] (BF_Sigmoid)
    assert _lib.LoudestArgs.number.offset % 8 == 0 and _lib.LoudestArgs.workspace_bytes.offset % 8 == 0
    cc = _lib.cy_consts()
    assert cc.gw_dadt_sep_const < 0 and abs(cc.kepler_const_sepa / 1.19128405e-3 - 1) < 1e-6


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import holodeck_b200 as holo
    from holodeck_b200 import _lib, cyutils
    with pytest.raises(_lib.HoloNativeError):
        holo.sams.Semi_Analytic_Model(shape=8).static_binary_density
    with pytest.raises(_lib.HoloNativeError):
        cyutils.sam_poisson_gwb(np.ones((2, 2, 2, 2)), np.ones((2, 2, 2, 2)), 3)
    with pytest.raises(_lib.HoloNativeError):
        holo.sams.sam_cyutils.integrate_differential_number_3dx1d([np.arange(1, 4.0)] * 3 + [np.arange(1, 5.0)], np.ones((3, 3, 3, 3)))


def test_product_never_imports_the_oracle():
    """Nothing under holodeck_b200/ may reference oracle/ (the oracle is test infrastructure)."""
    for path in (ROOT / "holodeck_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path
