"""The C-ABI library loads without a GPU, exports every symbol include/breakfast_b200.h declares,
and refuses to compute without a device (no CPU fallback).  No compute calls here."""
import ctypes as C
import re
from pathlib import Path

import pytest

from breakfast_b200 import _native

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "breakfast_b200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(bf_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_native.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert lib.bf_abi_version() == 1


def test_header_cites_the_reference_interface():
    text = HEADER.read_text()
    for cite in ("breakfast.py:279-340", "breakfast.py:223-276", "breakfast.py:93-113", "cache.py:51-71"):
        assert cite in text


def test_invalid_arguments_are_reported_not_crashed():
    lib = _native.load()
    assert lib.bf_device_count(None) == _native.BF_ERR_INVALID
    assert b"null" in lib.bf_last_error()
    assert lib.bf_ctx_create(0, None, None) == _native.BF_ERR_INVALID


def test_no_device_means_error_not_fallback():
    if _native.device_count() > 0:
        pytest.skip("a GPU is present")
    lib = _native.load()
    ctx = C.c_void_p()
    assert lib.bf_ctx_create(0, None, C.byref(ctx)) == _native.BF_ERR_NO_DEVICE
    assert not ctx.value
    with pytest.raises(_native.NativeError) as e:
        _native.require_device()
    assert e.value.code == _native.BF_ERR_NO_DEVICE
    g = C.c_double()
    assert lib.bf_measure_peak(0, b"popc32", C.byref(g)) == _native.BF_ERR_NO_DEVICE
