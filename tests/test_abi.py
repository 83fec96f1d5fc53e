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


def test_csr16_encode_round_trip_and_limits():
    """bf_csr16_encode (pure host code, no GPU needed): the compact form decodes back to the plain CSR; matrices that
    do not fit it are refused"""
    import numpy as np
    import pytest
    from breakfast_b200 import _native
    rng = np.random.default_rng(5)
    for n_cols in (300, 65536, 65537, 131072):
        rows = [np.sort(rng.choice(n_cols, size=int(rng.integers(0, 120)), replace=False)) for _ in range(400)]
        rows[7] = np.zeros(0, np.int64)
        if n_cols > 65536:
            rows[3] = np.unique(np.array([0, 65535, 65536, n_cols - 1]))
        indptr = np.concatenate(([0], np.cumsum([len(r) for r in rows]))).astype(np.int64)
        indices = np.concatenate(rows).astype(np.int32)
        ip32, split, lo = _native.csr16_encode(indptr, indices, n_cols)
        assert ip32.dtype == np.uint32 and lo.dtype == np.uint16 and np.array_equal(ip32, indptr)
        assert (split is None) == (n_cols <= 65536)
        pos = np.arange(indices.size) - np.repeat(indptr[:-1], np.diff(indptr))
        hi = np.zeros(indices.size, np.int64) if split is None else (pos >= np.repeat(split.astype(np.int64), np.diff(indptr))) * 65536
        assert np.array_equal(lo.astype(np.int64) + hi, indices)
    indptr, indices = np.array([0, 2], np.int64), np.array([1, 200000], np.int32)
    with pytest.raises(_native.NativeError):
        _native.csr16_encode(indptr, indices, 200001)                       # too many columns
    with pytest.raises(_native.NativeError):
        _native.csr16_encode(np.array([0, 70000], np.int64), np.arange(70000, dtype=np.int32), 70000)   # row too long
    with pytest.raises(_native.NativeError):
        _native.csr16_encode(np.array([0, 2], np.int64), np.array([5, 5], np.int32), 10)                # not strictly ascending
