"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/c2v_b200.h declares; the ctypes binding and the header agree."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from code2vec_b200 import _lib

HEADER = os.path.join(ROOT, "include", "c2v_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(c2v_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in c2v_b200.h but not exported by libc2v_b200.so"
    assert sorted(_lib.SYMBOLS) == names, "ctypes binding and header disagree"
    assert lib.c2v_abi_version() == _lib.ABI_VERSION


def test_header_cites_reference_lines():
    src = open(HEADER).read()
    for cite in ("model.py:44-88", "model.py:83", "model.py:71-80", "main.py:251-264", "main.py:285"):
        assert cite in src


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.Dims) == 40
    assert ctypes.sizeof(_lib.Params) == 64
    assert ctypes.sizeof(_lib.Grads) == 48
    assert ctypes.sizeof(_lib.Dropout) == 16
    assert ctypes.sizeof(_lib.DeviceInfo) == 32


def test_size_queries_and_argument_checks_without_a_gpu():
    lib = _lib.load()
    d = _lib.Dims(1000, 800, 64, 128, 128, 128, 0)
    n = lib.c2v_encode_workspace_bytes(ctypes.byref(d), 1024, 200)
    assert n > 0 and n % 1024 == 0
    assert lib.c2v_encode_workspace_bytes(ctypes.byref(d), 0, 200) == 0
    assert lib.c2v_encode_supports_tcgen05(ctypes.byref(d)) == 1
    d2 = _lib.Dims(1000, 800, 64, 100, 100, 100, 0)        # the reference's defaults (main.py:56-58)
    assert lib.c2v_encode_supports_tcgen05(ctypes.byref(d2)) == 1
    for bad in (_lib.Dims(1000, 800, 64, 100, 96, 100, 0),      # terminal_embed != path_embed
                _lib.Dims(1000, 800, 64, 256, 256, 128, 0),     # embed > 128 needs encode_size 256
                _lib.Dims(1000, 800, 64, 260, 260, 256, 0),     # > 256
                _lib.Dims(1000, 800, 64, 128, 128, 64, 0),      # encode_size not in {100, 128}
                _lib.Dims(10_000_000, 800, 64, 128, 128, 128, 0)):   # table > 4 GB (32-bit row offsets)
        assert lib.c2v_encode_supports_tcgen05(ctypes.byref(bad)) == 0
    assert lib.c2v_encode_supports_tcgen05(ctypes.byref(_lib.Dims(1000, 800, 64, 256, 256, 256, 0))) == 1   # BASELINE configs[3]
    # NULL pointers are rejected before any CUDA call
    rc = lib.c2v_encode_forward(ctypes.byref(d), None, None, None, None, 4, 7, None, None, None, None, 0, 0, None)
    assert rc == _lib.C2V_EINVAL
    assert b"NULL" in lib.c2v_last_error()
    bad = _lib.Dims(0, 800, 64, 128, 128, 128, 0)
    assert lib.c2v_encode_workspace_bytes(ctypes.byref(bad), 4, 7) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.C2VError, match="no CPU"):
        _lib.load()
