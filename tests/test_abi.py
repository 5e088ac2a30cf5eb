"""The C-ABI library loads and exports every symbol include/bilby_b200.h declares (no compute calls:
there is no GPU in the build container and the library has no CPU path)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared():
    text = open(os.path.join(ROOT, "include", "bilby_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bilby_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/bilby_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)
    assert lib.bb_abi_version() == 1


def test_enums_match_python_packing():
    from bilby_b200.gw import _params
    text = open(os.path.join(ROOT, "include", "bilby_b200.h")).read()
    for name, val in re.findall(r"BB_([A-Z0-9_]+) = (\d+)", text):
        if hasattr(_params, name):
            assert getattr(_params, name) == int(val), name
    assert _params.NPARAM == int(re.search(r"#define BB_NPARAM (\d+)", text).group(1))


def test_no_cpu_path_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from bilby_b200 import _lib
    with pytest.raises(_lib.BilbyB200Error):
        _lib.Handle()
    lib = _lib.load()
    ptr = ctypes.c_void_p()
    assert lib.bb_create(0, ctypes.byref(ptr)) != 0
    assert b"no CUDA device" in lib.bb_last_error()


def test_torch_library_shim_registers_the_ops():
    """csrc/bb_torch.cpp: TORCH_LIBRARY(bilby_b200) over the C ABI loads without a GPU and refuses CPU tensors loudly."""
    import torch
    from bilby_b200 import _lib
    ops = _lib.torch_ops()
    for name in ("log_likelihood_ratio", "log_likelihood_ratio_cal", "inner_products", "likelihood_from_inner_products"):
        assert hasattr(ops, name), name
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.log_likelihood_ratio(1, torch.zeros((2, 16), dtype=torch.float64))
