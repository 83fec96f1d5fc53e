"""CPU oracle for the breakfast distance-and-clustering path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import
this package; it is the checker, never the product.  breakfast_b200 does not import it.

  oracle.c       plain-C restatement of the distance + threshold + components core
  hashjoin.py    second, independent oracle for max_dist <= 2 at full scale (deletion-neighbourhood hash joins,
                 every match verified exactly)
  ref_port.py    Python restatement of the whole reference pipeline, structured like the reference
                 (per-cardinality batches through scikit-learn's pairwise_distances_chunked, networkx
                 components), used as the CPU baseline ("port")

Parity pin: tests/test_oracle_golden.py checks both against the reference's own golden files and
against outputs produced by running the reference itself (tests/golden/make_golden.py).
The reference is pure Python, so there is nothing to compile into oracle/_ref.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB = _DIR / "liboracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src = _DIR / "oracle.c"
    if force or not _LIB.exists() or _LIB.stat().st_mtime < src.stat().st_mtime:
        cmd = ["gcc", "-O2", "-fPIC", "-fopenmp", "-std=c11", "-shared", str(src), "-o", str(_LIB)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(str(build()))
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        lib.orc_distance.restype = i64
        lib.orc_distance.argtypes = [vp, vp, i64, i64]
        lib.orc_edges.restype = i64
        lib.orc_edges.argtypes = [vp, vp, i64, i32, vp, i64, C.POINTER(vp), C.POINTER(vp)]
        lib.orc_free.argtypes = [vp]
        lib.orc_free.restype = None
        lib.orc_components.restype = C.c_int
        lib.orc_components.argtypes = [i64, vp, vp, i64, vp, vp, i64, vp]
        lib.orc_cluster.restype = i64
        lib.orc_cluster.argtypes = [vp, vp, i64, i32, vp]
        lib.orc_within_batch.restype = None
        lib.orc_within_batch.argtypes = [vp, vp, vp, vp, i64, i64, vp]
        lib.orc_join_two_deletions.restype = i64
        lib.orc_join_two_deletions.argtypes = [vp, vp, vp, i64, vp, vp, vp, i32, C.POINTER(vp), C.POINTER(vp)]
        _lib = lib
    return _lib


def _csr(indptr, indices):
    return np.ascontiguousarray(indptr, dtype=np.int64), np.ascontiguousarray(indices, dtype=np.int32)


def distance(indptr, indices, a: int, b: int) -> int:
    indptr, indices = _csr(indptr, indices)
    return int(_load().orc_distance(indptr.ctypes.data, indices.ctypes.data, a, b))


def edges(indptr, indices, max_dist: int, queries=None):
    """(src, dst) int32, src < dst, sorted; pairs with at least one endpoint in `queries` (None = all)."""
    indptr, indices = _csr(indptr, indices)
    n = indptr.size - 1
    q = None if queries is None else np.ascontiguousarray(queries, dtype=np.int32)
    s, d = C.c_void_p(), C.c_void_p()
    lib = _load()
    ne = lib.orc_edges(indptr.ctypes.data, indices.ctypes.data, n, int(max_dist),
                       None if q is None else q.ctypes.data, 0 if q is None else q.size, C.byref(s), C.byref(d))
    if ne < 0:
        raise MemoryError("orc_edges failed")
    src = np.ctypeslib.as_array(C.cast(s, C.POINTER(C.c_int32)), shape=(max(ne, 1),))[:ne].copy()
    dst = np.ctypeslib.as_array(C.cast(d, C.POINTER(C.c_int32)), shape=(max(ne, 1),))[:ne].copy()
    lib.orc_free(s)
    lib.orc_free(d)
    return src, dst


def components(n: int, src=None, dst=None, list_indptr=None, list_members=None) -> np.ndarray:
    src = np.zeros(0, np.int32) if src is None else np.ascontiguousarray(src, dtype=np.int32)
    dst = np.zeros(0, np.int32) if dst is None else np.ascontiguousarray(dst, dtype=np.int32)
    li = np.zeros(1, np.int64) if list_indptr is None else np.ascontiguousarray(list_indptr, dtype=np.int64)
    lm = np.zeros(0, np.int32) if list_members is None else np.ascontiguousarray(list_members, dtype=np.int32)
    labels = np.empty(n, dtype=np.int32)
    rc = _load().orc_components(n, src.ctypes.data, dst.ctypes.data, src.size, li.ctypes.data, lm.ctypes.data,
                                li.size - 1, labels.ctypes.data)
    if rc != 0:
        raise MemoryError("orc_components failed")
    return labels


def cluster(indptr, indices, max_dist: int):
    """labels (smallest row index per component) and the number of edges of the full graph."""
    indptr, indices = _csr(indptr, indices)
    n = indptr.size - 1
    labels = np.empty(n, dtype=np.int32)
    ne = _load().orc_cluster(indptr.ctypes.data, indices.ctypes.data, n, int(max_dist), labels.ctypes.data)
    if ne < 0:
        raise MemoryError("orc_cluster failed")
    return labels, int(ne)
