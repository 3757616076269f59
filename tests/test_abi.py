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


def test_argument_validation_is_host_side():
    """Bad arguments are refused before anything is launched (works without a GPU): error code + message, no crash."""
    import ctypes as C
    from colosseumrl_b200 import _lib
    lib = _lib.load()
    assert lib.crl_tron_state_bytes(19, 4, 65536) == 208 * 65536
    assert lib.crl_tron_state_bytes(25, 4, 1) == (4 * 20 + 9) * 4                                # N > 19: the wide layout
    assert lib.crl_tron_state_bytes(19, 5, 1) == (5 * 12 + 11) * 4                               # P > 4: the wide layout
    assert lib.crl_tron_state_bytes(65, 4, 1) < 0 and b"board size" in lib.crl_last_error()
    assert lib.crl_tron_state_bytes(19, 9, 1) < 0 and b"player count" in lib.crl_last_error()
    assert (lib.crl_tron_action_stride(19, 4), lib.crl_tron_result_bytes(19, 4)) == (4, 8)
    assert (lib.crl_tron_action_stride(21, 4), lib.crl_tron_result_bytes(11, 6)) == (8, 16)
    assert lib.crl_blokus_state_bytes(16384) == 352 * 16384
    h, d = (C.c_int32 * 4)(), (C.c_int32 * 4)()
    assert lib.crl_tron_start_positions(19, 4, h, d) == 0 and list(h) == [30, 226, 330, 134] and list(d) == [2, 3, 0, 1]
    assert lib.crl_tron_start_positions_at(19, 4, 9, 0, h, d) != 0          # no such ring
    assert lib.crl_tron_start_positions_at(19, 4, -1, 0, h, d) != 0
    assert lib.crl_tron_reset(None, None, 8, 19, 4, None) != 0 and b"null" in lib.crl_last_error()
    assert lib.crl_tron_step(None, None, None, None, None, 8, 19, 4, 0, None) != 0
    assert lib.crl_ttt_reset(None, None, 8, 4, None) != 0
    assert lib.crl_ttt_reset(None, None, 8, 5, None) != 0                    # no 5-player variant
    assert lib.crl_blokus_reset(None, None, 8, None) != 0
    assert lib.crl_blokus_legal(None, 0, None, None, 4096, None, 8, 0, None) != 0
    assert lib.crl_blokus_is_valid(None, 0, None, None, 8, 0, None) != 0
    assert lib.crl_blokus_step(None, None, None, None, None, 8, 0, None) != 0
    assert lib.crl_ttt_cells(2) == 9 and lib.crl_ttt_cells(3) == 15 and lib.crl_ttt_cells(4) == 27
