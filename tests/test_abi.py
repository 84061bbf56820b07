"""The C-ABI library loads and exports every symbol include/slmgs.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import CUDA_LIB, ROOT
from slmsuite_b200 import _lib


def _declared():
    text = open(os.path.join(ROOT, "include", "slmgs.h")).read()
    return sorted(set(re.findall(r"SLMGS_API[^;(]*?\b(slmgs_\w+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    names = _declared()
    assert len(names) >= 30
    for must in ("slmgs_create", "slmgs_destroy", "slmgs_run", "slmgs_forward", "slmgs_constrain_inverse",
                 "slmgs_update_weights", "slmgs_update_weights_spot", "slmgs_get_phase", "slmgs_last_error"):
        assert must in names


def test_bindings_cover_the_header():
    assert sorted(_lib.SIGNATURES) == _declared()


def test_cuda_library_exports_every_symbol(cuda_library):
    lib = ctypes.CDLL(cuda_library)
    for name in _declared():
        assert hasattr(lib, name), name
    lib.slmgs_version.restype = ctypes.c_int
    assert lib.slmgs_version() >= 100


def test_emulation_library_exports_every_symbol(emu_library):
    lib = ctypes.CDLL(emu_library)
    for name in _declared():
        assert hasattr(lib, name), name


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "DEFAULT_LIBRARY", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_params_struct_layout_matches_header():
    # slmgs_params: 3 ints, 2 floats, 2 ints, 1 float, 3 ints, 1 float -> 48 bytes, no padding
    assert ctypes.sizeof(_lib.Params) == 48
    assert _lib.Params.zero_factor.offset == 44
    assert _lib.Params.mraf_factor.offset == 28
    assert _lib.Params.spot_width.offset == 36
