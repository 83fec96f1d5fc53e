"""Native fast path for the host side of the pipeline: filter_features + collapse_duplicates +
sparse_feature_matrix (+ the binary CSR for the device) in one pass over the profiles.

Semantics are those of the Python functions in breakfast.py (which stay the reference-shaped API and are
what the tests of the reference call): the per-occurrence work (split, intern, filter, dedup, CSR) runs
in csrc/host_parse.cpp, the classification of the DISTINCT tokens runs here with the same regular
expressions as filter_features.  tests/test_host_cpu.py checks both paths against each other.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np
import pandas as pd

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(str(_build.build_host()))
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        lib.bfh_tokenise.restype = vp
        lib.bfh_tokenise.argtypes = [C.c_char_p, i64, C.c_char, C.c_char_p, i32, i64]
        lib.bfh_tokenise_arrow.restype = vp
        lib.bfh_tokenise_arrow.argtypes = [vp, vp, i32, i64, C.c_char_p, i32, i32]
        lib.bfh_tokenise_arrow_chunks.restype = vp
        lib.bfh_tokenise_arrow_chunks.argtypes = [i32, vp, vp, vp, i32, C.c_char_p, i32, i32]
        for name in ("bfh_n_tokens", "bfh_distinct_bytes", "bfh_n_unique", "bfh_n_invalid", "bfh_token_nnz",
                     "bfh_binary_nnz", "bfh_string_bytes"):
            getattr(lib, name).restype = i64
            getattr(lib, name).argtypes = [vp]
        for name in ("bfh_n_distinct", "bfh_n_vocab", "bfh_n_cols"):
            getattr(lib, name).restype = i32
            getattr(lib, name).argtypes = [vp]
        lib.bfh_get_distinct.argtypes = [vp, vp, vp]
        lib.bfh_classify_dna.restype = i32
        lib.bfh_classify_dna.argtypes = [vp, i32, i32, i32, i64, i64, vp]
        lib.bfh_build.argtypes = [vp, vp, C.c_int]
        lib.bfh_get_results.argtypes = [vp] + [vp] * 9
        lib.bfh_free.argtypes = [vp]
        _lib = lib
    return _lib


_warned = False


def available() -> bool:
    """True when the native host parser can be used.  A missing compiler means "not available" (the Python
    functions of breakfast.py then do the same work, about three times slower - said once on stderr); a library
    that was built but does not load or lacks a symbol is a broken installation and raises."""
    global _warned
    try:
        _build.build_host()
    except RuntimeError as exc:
        if not _warned:
            print(f"breakfast_b200: native host parser unavailable ({str(exc).splitlines()[0]}); "
                  f"using the Python host path", file=sys.stderr)
            _warned = True
        return False
    _load()   # OSError / AttributeError propagate: the .so exists but is unusable
    return True


_NATIVE_TYPES = {"covsonar_dna": 0, "nextclade_dna": 1}   # --var-type values bfh_classify_dna knows
_ASK_PYTHON = 255


def _fits_i64(v) -> bool:
    return -(2 ** 63) <= int(v) < 2 ** 63


def host_threads() -> int:
    """threads for the native host pass: BREAKFAST_B200_HOST_THREADS, else min(cores, 16)"""
    env = os.environ.get("BREAKFAST_B200_HOST_THREADS", "")
    if env.strip():
        return max(1, int(env))
    return max(1, min(os.cpu_count() or 1, 16))


def _arrow_strings(series):
    """(arrow chunks, data pointers, offset pointers, rows per chunk, offset width) of a pandas string column - read in
    place, chunk by chunk (the Arrow CSV reader hands over hundreds of chunks per million lines; gluing them together
    would copy the whole column) - or None when the column holds missing values or is not an Arrow string column"""
    import pyarrow as pa
    try:
        arr = pa.array(series)
    except (pa.ArrowInvalid, pa.ArrowTypeError, TypeError):
        return None
    chunks = list(arr.chunks) if isinstance(arr, pa.ChunkedArray) else [arr]
    if not chunks:
        return None
    kind = chunks[0].type
    if not (pa.types.is_string(kind) or pa.types.is_large_string(kind)):
        return None
    width = 8 if pa.types.is_large_string(kind) else 4
    data_ptrs, off_ptrs, rows = [], [], []
    for ch in chunks:
        if ch.null_count or ch.type != kind:
            return None
        _, offsets, data = ch.buffers()
        data_ptrs.append(data.address if data is not None else 0)
        off_ptrs.append(offsets.address + ch.offset * width)
        rows.append(len(ch))
    return chunks, data_ptrs, off_ptrs, rows, width


def prepare(meta, feature_sep, feature_type, skip_ins, skip_del, trim_start, trim_end, reference_length, threads=None):
    """meta (DataFrame[id, feature], raw profiles) -> (meta_nodups, csr): meta_nodups exactly as
    collapse_duplicates(meta with feature = filter_features(...)) would build it, csr = the token CSR and the
    binary CSR of the unique profiles (a plain dict handed to breakfast.cluster(..., pre=csr); it is NOT stored
    in DataFrame.attrs, which pandas deep-copies on every column access and pickles into the cache).  Returns
    None when the input cannot take the fast path (missing profiles); the caller then uses the Python functions.
    The profile column is read in place through its Arrow buffers and the unique profile strings come back the same
    way; the native pass runs on `threads` host threads (host_threads() by default)."""
    import pyarrow as pa
    from .breakfast import _INVALID, _token_classifier

    n_seq = len(meta)
    sep_b = feature_sep.encode("utf-8")
    if n_seq == 0 or not sep_b:
        return None
    view = _arrow_strings(meta["feature"])
    if view is None:
        return None
    arr, data_ptrs, off_ptrs, rows, width = view   # `arr` keeps the buffers alive until bfh_build is done with them
    lib = _load()
    n_chunks = len(rows)
    state = lib.bfh_tokenise_arrow_chunks(n_chunks, (C.c_void_p * n_chunks)(*data_ptrs), (C.c_void_p * n_chunks)(*off_ptrs),
                                          (C.c_int64 * n_chunks)(*rows), width, sep_b, len(sep_b), int(threads or host_threads()))
    if not state:
        return None
    try:
        nd = lib.bfh_n_distinct(state)
        tok_bytes = C.create_string_buffer(max(1, lib.bfh_distinct_bytes(state)))
        tok_off = np.empty(nd + 1, dtype=np.int64)
        lib.bfh_get_distinct(state, tok_bytes, tok_off.ctypes.data)
        raw = tok_bytes.raw

        def token(i):
            return raw[tok_off[i]:tok_off[i + 1]].decode("utf-8")

        filter_active = bool(skip_del or skip_ins or trim_start > 0 or trim_end > 0)
        verdict = np.zeros(max(nd, 1), dtype=np.uint8)
        if filter_active and feature_type != "raw":
            classify = _token_classifier(feature_type, skip_ins, skip_del, trim_start, trim_end, reference_length)
            upper_cut = reference_length - trim_end
            if feature_type in _NATIVE_TYPES and _fits_i64(trim_start) and _fits_i64(upper_cut):
                # the two DNA notations are judged natively (byte scanners that restate the reference's expressions for
                # ASCII tokens); whatever is not plain ASCII comes back as _ASK_PYTHON and gets the expressions themselves
                left = lib.bfh_classify_dna(state, _NATIVE_TYPES[feature_type], int(bool(skip_ins)), int(bool(skip_del)),
                                            int(trim_start), int(upper_cut), verdict.ctypes.data)
                if left:
                    for i in np.flatnonzero(verdict[:nd] == _ASK_PYTHON).tolist():
                        verdict[i] = classify(token(i))
            else:
                verdict[:nd] = [classify(token(i)) for i in range(nd)]
        lib.bfh_build(state, verdict.ctypes.data, int(filter_active))

        n_unique = lib.bfh_n_unique(state)
        codes = np.empty(n_seq, dtype=np.int32)
        first_seq = np.empty(n_unique, dtype=np.int32)
        invalid = np.empty(lib.bfh_n_invalid(state), dtype=np.int32)
        u_ptr = np.empty(n_unique + 1, dtype=np.int64)
        u_idx = np.empty(lib.bfh_token_nnz(state), dtype=np.int32)
        b_ptr = np.empty(n_unique + 1, dtype=np.int64)
        b_idx = np.empty(lib.bfh_binary_nnz(state), dtype=np.int32)
        s_off = np.empty(n_unique + 1, dtype=np.int64)
        s_bytes = np.empty(max(1, lib.bfh_string_bytes(state)), dtype=np.uint8)
        lib.bfh_get_results(state, codes.ctypes.data, first_seq.ctypes.data, invalid.ctypes.data, u_ptr.ctypes.data,
                            u_idx.ctypes.data, b_ptr.ctypes.data, b_idx.ctypes.data, s_off.ctypes.data, s_bytes.ctypes.data)
        n_vocab, n_cols = lib.bfh_n_vocab(state), lib.bfh_n_cols(state)
    finally:
        lib.bfh_free(state)
    del arr

    if filter_active and feature_type != "raw":
        assert _INVALID == 2
        for t in invalid.tolist():
            print(f"Skipping invalid feature: '{token(t)}'")
    print(f"Number of duplicates: {n_seq - n_unique}")
    # ids of every unique profile as a tuple, in first-appearance order of the profiles
    ids = meta["id"].to_numpy(dtype=object)
    if n_unique == n_seq:
        grouped = list(zip(ids.tolist()))
    else:
        order = np.argsort(codes, kind="stable")
        counts = np.bincount(codes, minlength=n_unique)
        bounds = np.concatenate(([0], np.cumsum(counts)))
        ids_sorted = ids[order]
        grouped = list(zip(ids_sorted[bounds[:-1]].tolist()))          # right for the single-sequence profiles
        for g in np.flatnonzero(counts > 1).tolist():
            grouped[g] = tuple(ids_sorted[bounds[g]:bounds[g + 1]])
    # the unique profile strings as an Arrow string column built on the native buffers (no per-string decode)
    strings = pa.LargeStringArray.from_buffers(n_unique, pa.py_buffer(s_off), pa.py_buffer(s_bytes))
    nodups = pd.DataFrame({"id": pd.Series(grouped, dtype=object),
                           "feature": pd.Series(strings, dtype=meta["feature"].dtype)})
    print(f"Number of unique sequences: {n_unique}")
    csr = dict(n=n_unique, token_indptr=u_ptr, token_indices=u_idx, n_vocab=int(n_vocab),
               bin_indptr=b_ptr, bin_indices=b_idx, n_cols=int(n_cols))
    return nodups, csr
