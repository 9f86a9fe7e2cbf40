"""CPU: the C-ABI shared library loads and exports every symbol include/kgwas_b200.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "kgwas_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"KGB_API[^;(]*?\b(kgb_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from kgwas_b200 import _lib
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_declares_symbols():
    syms = _declared_symbols()
    assert "kgb_spmm" in syms and "kgb_csr_build" in syms and len(syms) >= 15


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in kgwas_b200.h but not exported: {missing}"


def test_binding_covers_every_declared_symbol():
    from kgwas_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_identity_calls(lib):
    assert lib.kgb_sm_arch() == 100
    assert lib.kgb_version() >= 100
    lib.kgb_last_error.restype = ctypes.c_char_p
    assert lib.kgb_last_error() is not None


def test_sass_is_sm100a_only():
    import subprocess
    from kgwas_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback():
    """A CPU tensor must raise: there is no CPU / PyTorch fallback on the product path."""
    import torch
    from kgwas_b200 import _lib
    x = torch.zeros(4, 32)
    with pytest.raises(_lib.KgbError):
        _lib.relu_bwd(x, x)
