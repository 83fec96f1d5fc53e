"""ctypes binding of libbreakfast_b200.so (C ABI: include/breakfast_b200.h).

This is the only door from the Python host into the CUDA kernels.  There is no CPU fallback:
a missing library or a missing GPU raises NativeError.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import build as _build

ENGINE_SKETCH = 0
ENGINE_FULL = 1
ENGINE_HASHJOIN = 2
_ENGINES = {"sketch": ENGINE_SKETCH, "full": ENGINE_FULL, "hashjoin": ENGINE_HASHJOIN, 0: 0, 1: 1, 2: 2}

BF_OK = 0
BF_ERR_INVALID = -1
BF_ERR_NO_DEVICE = -2
BF_ERR_CUDA = -3
BF_ERR_OOM = -4
BF_ERR_OVERFLOW = -5
BF_ERR_STATE = -6


class NativeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libbreakfast_b200 error {code}: {message}")
        self.code = code


class Stats(C.Structure):
    _int_fields = ("n_rows", "n_query", "n_cols", "nnz", "bits_per_row", "pairs_total", "pairs_band",
                   "pairs_evaluated", "tiles_total", "tiles_band", "tiles_rank", "n_candidates", "n_edges",
                   "n_components")
    _dbl_fields = ("ms_h2d", "ms_sort", "ms_pack", "ms_pairs", "ms_verify", "ms_cc", "ms_merge", "ms_d2h",
                   "ms_total")
    _fields_ = ([(n, C.c_int64) for n in _int_fields] + [(n, C.c_double) for n in _dbl_fields]
                + [("runs_since_sync", C.c_int64), ("kernel_launches", C.c_int64),
                   ("ms_pairs_sum", C.c_double), ("ms_total_sum", C.c_double),
                   ("l2_warp_items", C.c_int64), ("popc32_executed", C.c_int64),
                   ("ms_l1_sum", C.c_double)])

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None

# every symbol include/breakfast_b200.h declares
EXPORTS = (
    "bf_abi_version", "bf_last_error", "bf_device_count", "bf_ctx_create", "bf_ctx_destroy", "bf_ctx_set_option",
    "bf_upload_csr", "bf_upload_csr_async", "bf_csr16_encode", "bf_upload_csr16_async", "bf_adopt_csr_device", "bf_run", "bf_comm_unique_id", "bf_ctx_comm_init_rank", "bf_comm_init_all", "bf_ctx_comm_destroy", "bf_labels_to_device", "bf_merge_labels_device", "bf_merge_labels_host",
    "bf_union_lists", "bf_sync", "bf_download_labels", "bf_edge_count", "bf_download_edges", "bf_cluster_csr",
    "bf_neighbours_csr", "bf_edges_copy", "bf_edges_free", "bf_components", "bf_pinned_alloc", "bf_pinned_free",
    "bf_measure_peak",
)


def library_path() -> Path:
    return _build.LIB_PATH


def load() -> C.CDLL:
    """Load the native library, (re)building it in-tree with nvcc when it is absent or older than its sources
    (build.needs_build: a stale library after an edit to kernels.cuh would otherwise be loaded silently)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    lib = C.CDLL(str(path))
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.bf_abi_version.restype = C.c_int
    lib.bf_last_error.restype = C.c_char_p
    lib.bf_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.bf_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    lib.bf_ctx_destroy.argtypes = [vp]
    lib.bf_ctx_destroy.restype = None
    lib.bf_ctx_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.bf_upload_csr.argtypes = [vp, vp, vp, i64, i32, vp, i64]
    lib.bf_upload_csr_async.argtypes = [vp, vp, vp, i64, i32]
    lib.bf_csr16_encode.argtypes = [vp, vp, i64, i32, vp, vp, vp]
    lib.bf_upload_csr16_async.argtypes = [vp, vp, vp, vp, i64, i32]
    lib.bf_adopt_csr_device.argtypes = [vp, vp, vp, i64, i32, i64]
    lib.bf_run.argtypes = [vp, i32, i32, i32]
    lib.bf_comm_unique_id.argtypes = [vp]
    lib.bf_ctx_comm_init_rank.argtypes = [vp, vp, i32, i32]
    lib.bf_comm_init_all.argtypes = [C.POINTER(vp), i32]
    lib.bf_ctx_comm_destroy.argtypes = [vp]
    lib.bf_labels_to_device.argtypes = [vp, vp]
    lib.bf_merge_labels_device.argtypes = [vp, vp, i32]
    lib.bf_merge_labels_host.argtypes = [vp, vp, i32]
    lib.bf_union_lists.argtypes = [vp, vp, vp, i64]
    lib.bf_sync.argtypes = [vp, C.POINTER(Stats)]
    lib.bf_download_labels.argtypes = [vp, vp]
    lib.bf_edge_count.argtypes = [vp, C.POINTER(i64)]
    lib.bf_download_edges.argtypes = [vp, vp, vp]
    lib.bf_cluster_csr.argtypes = [vp, vp, i64, i32, i32, i32, i32, vp, C.POINTER(Stats)]
    lib.bf_neighbours_csr.argtypes = [vp, vp, i64, i32, vp, i64, i32, i32, i32, C.POINTER(vp), C.POINTER(i64),
                                      C.POINTER(Stats)]
    lib.bf_edges_copy.argtypes = [vp, vp, vp]
    lib.bf_edges_free.argtypes = [vp]
    lib.bf_edges_free.restype = None
    lib.bf_components.argtypes = [i64, vp, vp, i64, vp, vp, i64, i32, vp, C.POINTER(i64)]
    lib.bf_pinned_alloc.argtypes = [i64, C.POINTER(vp)]
    lib.bf_pinned_free.argtypes = [vp]
    lib.bf_pinned_free.restype = None
    lib.bf_measure_peak.argtypes = [i32, C.c_char_p, C.POINTER(C.c_double)]
    if lib.bf_abi_version() != 1:
        raise NativeError(BF_ERR_STATE, f"ABI version mismatch: library says {lib.bf_abi_version()}, binding wants 1")
    _lib = lib
    return lib


def _ck(rc: int) -> None:
    if rc != BF_OK:
        raise NativeError(rc, load().bf_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return None if a is None else a.ctypes.data


def _csr_args(indptr, indices):
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    if indptr.ndim != 1 or indptr.size < 1:
        raise ValueError("indptr must be a 1-d array of length n_rows + 1")
    return indptr, indices


def device_count() -> int:
    n = C.c_int(0)
    rc = load().bf_device_count(C.byref(n))
    return n.value if rc == BF_OK else 0


def require_device() -> int:
    n = C.c_int(0)
    _ck(load().bf_device_count(C.byref(n)))
    if n.value < 1:
        raise NativeError(BF_ERR_NO_DEVICE, "no CUDA device visible; breakfast_b200 has no CPU fallback")
    return n.value


def measure_peak(name: str, device: int = 0) -> float:
    g = C.c_double(0)
    _ck(load().bf_measure_peak(device, name.encode(), C.byref(g)))
    return g.value


class Context:
    """Device-resident, asynchronous API (bf_ctx_*): upload once, run many times, merge across ranks."""

    def __init__(self, device: int = 0, stream: int | None = None, engine="sketch", **options):
        self._lib = load()
        self._h = C.c_void_p()
        _ck(self._lib.bf_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.n_rows = 0
        self.comm = None
        self.set_option("engine", _ENGINES[engine])
        for k, v in options.items():
            self.set_option(k, v)

    def close(self):
        if self._h:
            self._lib.bf_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int):
        _ck(self._lib.bf_ctx_set_option(self._h, key.encode(), int(value)))

    def upload_csr(self, indptr, indices, n_cols: int, query_rows=None):
        indptr, indices = _csr_args(indptr, indices)
        q = None
        if query_rows is not None:
            q = np.ascontiguousarray(query_rows, dtype=np.int32)
        self.n_rows = indptr.size - 1
        _ck(self._lib.bf_upload_csr(self._h, _ptr(indptr), _ptr(indices), self.n_rows, int(n_cols), _ptr(q),
                                    0 if q is None else q.size))

    def upload_csr_ptr(self, indptr_ptr: int, indices_ptr: int, n_rows: int, n_cols: int):
        """Upload from raw host addresses (e.g. pinned buffers)."""
        self.n_rows = n_rows
        _ck(self._lib.bf_upload_csr(self._h, indptr_ptr, indices_ptr, n_rows, int(n_cols), None, 0))

    def upload_csr_async_ptr(self, indptr_ptr: int, indices_ptr: int, n_rows: int, n_cols: int):
        """Pinned host buffers -> idle device slot on the context's copy stream (overlaps a running pass)."""
        self.n_rows = n_rows
        _ck(self._lib.bf_upload_csr_async(self._h, indptr_ptr, indices_ptr, n_rows, int(n_cols)))

    def upload_csr16_async_ptr(self, indptr32_ptr: int, split_ptr: int | None, lo_ptr: int, n_rows: int, n_cols: int):
        """Compact host form (csr16_encode) from pinned buffers -> idle device slot, decoded on the device."""
        self.n_rows = n_rows
        _ck(self._lib.bf_upload_csr16_async(self._h, indptr32_ptr, split_ptr, lo_ptr, n_rows, int(n_cols)))

    def adopt_csr_device(self, indptr_dev: int, indices_dev: int, n_rows: int, n_cols: int, nnz: int):
        """Use caller-owned device memory as the CSR (no copy)."""
        self.n_rows = n_rows
        _ck(self._lib.bf_adopt_csr_device(self._h, C.c_void_p(indptr_dev), C.c_void_p(indices_dev), n_rows,
                                          int(n_cols), int(nnz)))

    def comm_init_rank(self, unique_id: bytes, rank: int, world: int):
        """Join the library's own NCCL communicator (multi-process jobs: one context per process)."""
        if len(unique_id) != 128:
            raise ValueError("the communicator id has 128 bytes (comm_unique_id())")
        _ck(self._lib.bf_ctx_comm_init_rank(self._h, C.c_char_p(unique_id), int(rank), int(world)))
        self.comm = (int(rank), int(world))

    def comm_destroy(self):
        _ck(self._lib.bf_ctx_comm_destroy(self._h))
        self.comm = None

    def run(self, max_dist: int, rank: int = 0, world: int = 1):
        _ck(self._lib.bf_run(self._h, int(max_dist), int(rank), int(world)))

    def sync(self) -> Stats:
        st = Stats()
        _ck(self._lib.bf_sync(self._h, C.byref(st)))
        return st

    def run_sync(self, max_dist: int, rank: int = 0, world: int = 1) -> Stats:
        """run + sync; on BF_ERR_OVERFLOW bf_sync has raised the capacity of whatever overflowed (work list, level-2
        queue, candidate buffer - one can follow the other), so the pass is simply run again."""
        for _ in range(6):
            self.run(max_dist, rank, world)
            st = Stats()
            rc = self._lib.bf_sync(self._h, C.byref(st))
            if rc == BF_ERR_OVERFLOW:
                continue
            _ck(rc)
            return st
        raise NativeError(BF_ERR_OVERFLOW, "buffer overflow persisted")

    def labels_to_device(self, device_ptr: int):
        _ck(self._lib.bf_labels_to_device(self._h, C.c_void_p(device_ptr)))

    def merge_labels_device(self, gathered_device_ptr: int, world: int):
        _ck(self._lib.bf_merge_labels_device(self._h, C.c_void_p(gathered_device_ptr), int(world)))

    def merge_labels_host(self, gathered: np.ndarray):
        g = np.ascontiguousarray(gathered, dtype=np.int32)
        if g.ndim != 2 or g.shape[1] != self.n_rows:
            raise ValueError("gathered labels must have shape [world, n_rows]")
        _ck(self._lib.bf_merge_labels_host(self._h, _ptr(g), g.shape[0]))

    def union_lists(self, list_indptr, list_members):
        li = np.ascontiguousarray(list_indptr, dtype=np.int64)
        lm = np.ascontiguousarray(list_members, dtype=np.int32)
        _ck(self._lib.bf_union_lists(self._h, _ptr(li), _ptr(lm), li.size - 1))

    def download_labels(self) -> np.ndarray:
        out = np.empty(self.n_rows, dtype=np.int32)
        _ck(self._lib.bf_download_labels(self._h, _ptr(out)))
        return out

    def download_edges(self):
        n = C.c_int64(0)
        _ck(self._lib.bf_edge_count(self._h, C.byref(n)))
        src = np.empty(n.value, dtype=np.int32)
        dst = np.empty(n.value, dtype=np.int32)
        _ck(self._lib.bf_download_edges(self._h, _ptr(src), _ptr(dst)))
        return src, dst


def comm_unique_id() -> bytes:
    """128 bytes that identify a new NCCL communicator; rank 0 creates them, every rank passes them to comm_init_rank."""
    buf = C.create_string_buffer(128)
    _ck(load().bf_comm_unique_id(buf))
    return buf.raw


def comm_init_all(contexts) -> None:
    """One process, one Context per distinct device: rank i = contexts[i] (ncclCommInitAll inside the library)."""
    arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    _ck(load().bf_comm_init_all(arr, len(contexts)))
    for i, c in enumerate(contexts):
        c.comm = (i, len(contexts))


def csr16_encode(indptr, indices, n_cols: int, out=None):
    """Plain CSR -> compact host form (indptr32 uint32[n+1], split uint16[n] or None, lo uint16[nnz]).  `out` may hold
    three preallocated arrays (e.g. views of pinned memory).  Raises NativeError(BF_ERR_INVALID) when the matrix is not
    representable (n_cols > 131072, a row longer than 65535, nnz >= 2^32)."""
    indptr, indices = _csr_args(indptr, indices)
    n = indptr.size - 1
    if out is None:
        out = (np.empty(n + 1, np.uint32), np.empty(n, np.uint16) if n_cols > 65536 else None, np.empty(indices.size, np.uint16))
    ip32, split, lo = out
    _ck(load().bf_csr16_encode(_ptr(indptr), _ptr(indices), n, int(n_cols), _ptr(ip32), _ptr(split), _ptr(lo)))
    return ip32, split, lo


def cluster_csr(indptr, indices, n_cols: int, max_dist: int, device: int = 0, engine="sketch"):
    """One-shot bf_cluster_csr: labels[i] = smallest row index of row i's component."""
    indptr, indices = _csr_args(indptr, indices)
    n = indptr.size - 1
    labels = np.empty(n, dtype=np.int32)
    st = Stats()
    _ck(load().bf_cluster_csr(_ptr(indptr), _ptr(indices), n, int(n_cols), int(max_dist), int(device),
                              _ENGINES[engine], _ptr(labels), C.byref(st)))
    return labels, st


def neighbours_csr(indptr, indices, n_cols: int, max_dist: int, query_rows=None, device: int = 0, engine="sketch"):
    """One-shot bf_neighbours_csr: (src, dst, stats), each unordered pair once, src < dst, sorted."""
    indptr, indices = _csr_args(indptr, indices)
    n = indptr.size - 1
    q = None if query_rows is None else np.ascontiguousarray(query_rows, dtype=np.int32)
    handle = C.c_void_p()
    ne = C.c_int64(0)
    st = Stats()
    lib = load()
    _ck(lib.bf_neighbours_csr(_ptr(indptr), _ptr(indices), n, int(n_cols), _ptr(q), 0 if q is None else q.size,
                              int(max_dist), int(device), _ENGINES[engine], C.byref(handle), C.byref(ne),
                              C.byref(st)))
    try:
        src = np.empty(ne.value, dtype=np.int32)
        dst = np.empty(ne.value, dtype=np.int32)
        _ck(lib.bf_edges_copy(handle, _ptr(src), _ptr(dst)))
    finally:
        lib.bf_edges_free(handle)
    return src, dst, st


def components(n_rows: int, src=None, dst=None, list_indptr=None, list_members=None, device: int = 0):
    """bf_components: GPU union-find over explicit edges and/or member lists."""
    src = np.zeros(0, np.int32) if src is None else np.ascontiguousarray(src, dtype=np.int32)
    dst = np.zeros(0, np.int32) if dst is None else np.ascontiguousarray(dst, dtype=np.int32)
    if src.size != dst.size:
        raise ValueError("src and dst differ in length")
    li = np.zeros(1, np.int64) if list_indptr is None else np.ascontiguousarray(list_indptr, dtype=np.int64)
    lm = np.zeros(0, np.int32) if list_members is None else np.ascontiguousarray(list_members, dtype=np.int32)
    labels = np.empty(n_rows, dtype=np.int32)
    ncomp = C.c_int64(0)
    _ck(load().bf_components(int(n_rows), _ptr(src), _ptr(dst), src.size, _ptr(li), _ptr(lm), li.size - 1,
                             int(device), _ptr(labels), C.byref(ncomp)))
    return labels, ncomp.value
