"""Second, independent CPU oracle for max_dist <= 2 at full scale (TEST INFRASTRUCTURE, like everything under oracle/).

oracle.c compares all band pairs — O(N^2), minutes at 10^5 rows, hours at 10^6.  For small distances the edge set has a
closed form.  Two DIFFERENT binary rows A, B (sorted unique column ids) are at distance
    0  iff  A == B,
    1  iff  one is the other minus exactly one column,
    2  iff  |A| == |B| and A minus some column x equals B minus some column y                      ("swap"), or
            one is the other minus exactly two columns                                               ("superset").
With an additive 64-bit row hash H(A) = sum of g(col) (mod 2^64) each case is an equi-join (the "deletion neighbourhood"
join of SURVEY.md section 8(f) row 4):
    distance 1:        H(A) - g(x) == H(B),           |B| == |A| - 1,  x in A
    distance 2, swap:  H(A) - g(x) == H(B) - g(y),    |B| == |A|,      x in A, y in B
    distance 2, super: H(A) - g(x) - g(y) == H(B),    |B| == |A| - 2,  x < y in A
A true edge always satisfies its join, so none can be lost; every match is then verified with the exact two-pointer
distance of oracle.c, so a hash collision cannot add one.  Same definition as oracle.edges (breakfast.py:223-276 +
sklearn _pairwise_fast.pyx:83-107 restated): an edge is an unordered pair of distinct rows with |A xor B| <= max_dist.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import oracle

_TABLE_BITS = 27   # membership table over the low hash bits: rejects ~99 % of the probes with one lookup


def _g(cols: np.ndarray) -> np.ndarray:
    """splitmix64 of the column id"""
    with np.errstate(over="ignore"):
        x = cols.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def _pairs_of_runs(starts, lengths, members):
    """all unordered index pairs inside every run [start, start + length) of `members`"""
    out_a, out_b = [], []
    for L in np.unique(lengths):
        s = starts[lengths == L]
        i, j = np.triu_indices(int(L), k=1)
        out_a.append(members[(s[:, None] + i[None, :]).ravel()])
        out_b.append(members[(s[:, None] + j[None, :]).ravel()])
    if not out_a:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.concatenate(out_a).astype(np.int64), np.concatenate(out_b).astype(np.int64)


def _runs(sorted_keys):
    """(starts, lengths) of the runs of equal values with length >= 2"""
    n = sorted_keys.size
    if n < 2:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    starts = np.flatnonzero(np.r_[True, sorted_keys[1:] != sorted_keys[:-1]])
    lengths = np.diff(np.r_[starts, n])
    keep = lengths > 1
    return starts[keep], lengths[keep]


def edges(indptr, indices, max_dist: int, chunk_rows: int = 100_000):
    """(src, dst) int32, src < dst, sorted lexicographically — the contract of oracle.edges(indptr, indices, max_dist)
    for max_dist in {0, 1, 2}; rows must hold sorted unique column ids."""
    if max_dist not in (0, 1, 2):
        raise ValueError("the joins cover max_dist 0, 1 and 2 only")
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    n = indptr.size - 1
    if n < 2:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    card = np.diff(indptr)
    g = _g(indices)
    with np.errstate(over="ignore"):
        gsum = np.concatenate([np.zeros(1, np.uint64), np.cumsum(g, dtype=np.uint64)])
        H = gsum[indptr[1:]] - gsum[indptr[:-1]]              # wraps mod 2^64; empty rows -> 0
    del gsum
    order = np.lexsort((card, H))                              # rows by (H, card)
    Hs, cs = H[order], card[order]
    cand_a, cand_b = [], []

    # ---- distance 0: runs of equal (H, card)
    same_as_prev = (Hs[1:] == Hs[:-1]) & (cs[1:] == cs[:-1])
    starts = np.flatnonzero(np.r_[True, ~same_as_prev])
    lengths = np.diff(np.r_[starts, n])
    a, b = _pairs_of_runs(starts[lengths > 1], lengths[lengths > 1], order)
    cand_a.append(a)
    cand_b.append(b)

    mask = np.uint64((1 << _TABLE_BITS) - 1)
    if max_dist >= 1:
        # ---- distance 1: H(A) - g(x) == H(B), |B| == |A| - 1
        present = np.zeros(1 << _TABLE_BITS, dtype=bool)
        present[H & mask] = True
        for r0 in range(0, n, chunk_rows):
            r1 = min(n, r0 + chunk_rows)
            lo, hi = indptr[r0], indptr[r1]
            if hi == lo:
                continue
            rows = np.repeat(np.arange(r0, r1), card[r0:r1])
            with np.errstate(over="ignore"):
                key = H[rows] - g[lo:hi]
            maybe = np.flatnonzero(present[key & mask])
            key, rows = key[maybe], rows[maybe]
            first = np.searchsorted(Hs, key, "left")
            hit = np.flatnonzero(Hs[np.minimum(first, n - 1)] == key)   # ~ number of edges, not nnz
            first, key, rows = first[hit], key[hit], rows[hit]
            width = np.searchsorted(Hs, key, "right") - first
            for w in range(int(width.max()) if hit.size else 0):        # w-th row of every equal-hash run
                sel = width > w
                p = first[sel] + w
                ok = cs[p] == card[rows[sel]] - 1
                cand_a.append(rows[sel][ok])
                cand_b.append(order[p][ok])
        del present

    if max_dist >= 2:
        # ---- distance 2, swap: equal deletion keys among rows of equal cardinality.  The nnz keys are grouped in
        # 16 partitions by their top bits (memory), each sorted by (key, cardinality)
        rows_all = np.repeat(np.arange(n, dtype=np.int32), card)
        with np.errstate(over="ignore"):
            keys_all = H[rows_all] - g
        top = (keys_all >> np.uint64(60)).astype(np.uint8)
        for part in range(16):
            sel = np.flatnonzero(top == part)
            if sel.size < 2:
                continue
            k, r = keys_all[sel], rows_all[sel]
            c = card[r]
            o = np.lexsort((c, k))
            k, r, c = k[o], r[o], c[o]
            tag = np.r_[True, (k[1:] != k[:-1]) | (c[1:] != c[:-1])]
            st = np.flatnonzero(tag)
            ln = np.diff(np.r_[st, k.size])
            a, b = _pairs_of_runs(st[ln > 1], ln[ln > 1], r)
            cand_a.append(a)
            cand_b.append(b)
        del rows_all, keys_all, top
        # ---- distance 2, superset: H(A) - g(x) - g(y) == H(B), |B| == |A| - 2 (pairs of columns: done in C)
        table = np.zeros((1 << _TABLE_BITS) // 64, dtype=np.uint64)
        slots = np.unique(H & mask)
        np.bitwise_or.at(table, (slots >> np.uint64(6)).astype(np.int64), np.uint64(1) << (slots & np.uint64(63)))
        sorted_row = order.astype(np.int32)
        pa, pb = C.c_void_p(), C.c_void_p()
        lib = oracle._load()
        m = lib.orc_join_two_deletions(indptr.ctypes.data, g.ctypes.data, H.ctypes.data, n, Hs.ctypes.data,
                                       sorted_row.ctypes.data, table.ctypes.data, _TABLE_BITS, C.byref(pa), C.byref(pb))
        if m < 0:
            raise MemoryError("orc_join_two_deletions failed")
        cand_a.append(np.ctypeslib.as_array(C.cast(pa, C.POINTER(C.c_int32)), shape=(max(m, 1),))[:m].astype(np.int64))
        cand_b.append(np.ctypeslib.as_array(C.cast(pb, C.POINTER(C.c_int32)), shape=(max(m, 1),))[:m].astype(np.int64))
        lib.orc_free(pa)
        lib.orc_free(pb)

    # ---- canonical form, exact verification
    a = np.concatenate([np.asarray(x, dtype=np.int64) for x in cand_a])
    b = np.concatenate([np.asarray(x, dtype=np.int64) for x in cand_b])
    keep = a != b
    lo, hi = np.minimum(a[keep], b[keep]), np.maximum(a[keep], b[keep])
    packed = np.unique((lo << np.int64(32)) | hi)              # sorted lexicographically by (lo, hi)
    src = (packed >> np.int64(32)).astype(np.int32)
    dst = (packed & np.int64(0xFFFFFFFF)).astype(np.int32)
    ok = np.zeros(src.size, dtype=np.uint8)
    if src.size:
        oracle._load().orc_within_batch(indptr.ctypes.data, indices.ctypes.data, src.ctypes.data, dst.ctypes.data,
                                        src.size, int(max_dist), ok.ctypes.data)
    ok = ok.astype(bool)
    return src[ok].copy(), dst[ok].copy()


def edges_d1(indptr, indices, max_dist: int = 1):
    """kept name of the first version of this module (max_dist 0 or 1)"""
    if max_dist not in (0, 1):
        raise ValueError("edges_d1 covers max_dist 0 and 1 only")
    return edges(indptr, indices, max_dist)


def cluster(indptr, indices, max_dist: int):
    """labels (smallest row index per component) and the number of edges — oracle.cluster's contract"""
    src, dst = edges(indptr, indices, max_dist)
    return oracle.components(len(indptr) - 1, src, dst), int(src.size)


def cluster_d1(indptr, indices, max_dist: int = 1):
    return cluster(indptr, indices, max_dist)
