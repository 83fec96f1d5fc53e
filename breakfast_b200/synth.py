"""Synthetic SARS-CoV-2-shaped mutation profiles; wraps csrc/synth.c.

Benchmark/test input only.  One event code = pos*1024 + slot (see synth.c).  The generator is
deterministic in (n, seed, flags) and independent of numpy's RNG.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import build as _build

_BASES = "ACGT"
_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(str(_build.build_synth()))
        lib.bfsynth_create.restype = C.c_void_p
        lib.bfsynth_create.argtypes = [C.c_int64, C.c_uint64, C.c_int]
        lib.bfsynth_nnz.restype = C.c_int64
        lib.bfsynth_nnz.argtypes = [C.c_void_p]
        lib.bfsynth_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.bfsynth_free.argtypes = [C.c_void_p]
        lib.bfsynth_ref_base.restype = C.c_int
        lib.bfsynth_ref_base.argtypes = [C.c_int]
        _lib = lib
    return _lib


def ref_base(pos: int) -> str:
    return _BASES[(((pos * 2654435761) & 0xFFFFFFFF) >> 7) & 3]


def _ins_seq(k: int) -> str:
    """k in [0, 84): the inserted strings of length 1, 2, 3 over ACGT, in that order."""
    if k < 4:
        return _BASES[k]
    if k < 20:
        k -= 4
        return _BASES[k >> 2] + _BASES[k & 3]
    k -= 20
    return _BASES[k >> 4] + _BASES[(k >> 2) & 3] + _BASES[k & 3]


def token(code: int, var_type: str = "covsonar_dna") -> str:
    pos, slot = code >> 10, code & 1023
    if slot < 4:
        return f"{ref_base(pos)}{pos}{_BASES[slot]}"
    if slot < 34:
        length = slot - 3
        if var_type == "nextclade_dna":
            return str(pos) if length == 1 else f"{pos}-{pos + length - 1}"
        return f"del:{pos}:{length}"
    seq = _ins_seq(slot - 34)
    if var_type == "nextclade_dna":
        return f"{pos}:{seq}"
    rb = ref_base(pos)
    return f"{rb}{pos}{rb}{seq}"


@dataclass
class SynthProfiles:
    indptr: np.ndarray  # int64 [n+1]
    codes: np.ndarray   # int32 [nnz], ascending inside each row
    mult: np.ndarray    # int32 [n], sequences per profile

    @property
    def n(self) -> int:
        return len(self.indptr) - 1

    def csr(self, subs_only: bool = True):
        """Binary CSR (indptr int64, indices int32 ascending per row, n_cols) of the profiles as the
        default filters (--skip-del --skip-ins, no trimmed positions are ever generated) leave them."""
        codes, indptr = self.codes, self.indptr
        if subs_only:
            keep = (codes & 1023) < 4
            csum = np.concatenate(([0], np.cumsum(keep, dtype=np.int64)))
            indptr = csum[indptr]
            codes = codes[keep]
        # dense column ids in ascending code order (a lookup table beats np.unique at 10^8 entries)
        present = np.zeros(int(codes.max()) + 1 if codes.size else 1, dtype=bool)
        present[codes] = True
        col_of = np.cumsum(present, dtype=np.int32) - 1
        return np.ascontiguousarray(indptr, dtype=np.int64), col_of[codes], int(present.sum())

    def features(self, var_type: str = "covsonar_dna", sep: str = " ") -> list:
        # the distinct codes in ascending order and every entry's rank among them - what np.unique(return_inverse=True)
        # returns, through a presence table instead of an argsort of 10^8 entries
        present = np.zeros(int(self.codes.max()) + 1 if self.codes.size else 1, dtype=bool)
        present[self.codes] = True
        uniq = np.flatnonzero(present)
        inv = (np.cumsum(present, dtype=np.int64) - 1)[self.codes]
        toks = np.array([token(int(c), var_type) for c in uniq], dtype=object)
        flat = toks[inv]
        ip = self.indptr
        return [sep.join(flat[ip[i]:ip[i + 1]]) for i in range(self.n)]

    def table(self, var_type: str = "covsonar_dna", sep: str = " ", id_col: str = "accession",
              feature_col: str = "dna_profile", shuffle_seed: int | None = 0):
        """A pandas DataFrame with one line per *sequence* (profiles repeated `mult` times), shuffled."""
        import pandas as pd
        feats = np.array(self.features(var_type, sep), dtype=object)
        rep = np.repeat(np.arange(self.n), self.mult)
        if shuffle_seed is not None:
            rep = rep[np.random.default_rng(shuffle_seed).permutation(len(rep))]
        ids = [f"seq{i:08d}" for i in range(len(rep))]
        return pd.DataFrame({id_col: ids, feature_col: feats[rep]})


def generate(n: int, seed: int = 1, with_mult: bool = False, unique_on_all_events: bool = False) -> SynthProfiles:
    lib = _load()
    flags = (1 if with_mult else 0) | (2 if unique_on_all_events else 0)
    h = lib.bfsynth_create(int(n), int(seed), flags)
    if not h:
        raise MemoryError("bfsynth_create failed")
    try:
        nnz = lib.bfsynth_nnz(h)
        indptr = np.empty(n + 1, dtype=np.int64)
        codes = np.empty(nnz, dtype=np.int32)
        mult = np.empty(n, dtype=np.int32)
        lib.bfsynth_copy(h, indptr.ctypes.data, codes.ctypes.data, mult.ctypes.data)
    finally:
        lib.bfsynth_free(h)
    return SynthProfiles(indptr, codes, mult)
