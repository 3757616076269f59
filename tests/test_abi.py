"""The C-ABI shared library loads and exports every symbol include/colosseum_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "colosseum_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(crl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from colosseumrl_b200 import build, _lib
    build.build()
    lib = ctypes.CDLL(build.LIB)
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    # the ctypes signature table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names
    _lib.declare(lib)
    assert lib.crl_version() >= 100


def test_no_cpu_fallback_without_device():
    import torch
    from colosseumrl_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.crl_init(0) != 0            # fails loudly: no device
    assert b"CUDA" in lib.crl_last_error() or b"device" in lib.crl_last_error()
    with pytest.raises(_lib.CrlError):
        _lib.init(0)
