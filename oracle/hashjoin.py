"""Second, independent CPU oracle for max_dist = 1 at full scale (TEST INFRASTRUCTURE, like everything under oracle/).

oracle.c compares all band pairs — O(N^2), minutes at 10^5 rows, hours at 10^6.  For max_dist <= 1 the edge set has a
closed form: two DIFFERENT binary rows are at distance 1 iff one is the other minus exactly one column, and at distance 0
iff they are equal.  With an additive 64-bit row hash H(A) = sum of g(col) the candidates are an equi-join
    H(A) - g(x) == H(B),  |B| == |A| - 1,  x in A
(the "deletion neighbourhood" join of SURVEY.md section 8(f) row 4), found with one sort, a membership table over the
low hash bits for all nnz (row, column) entries and binary searches for the few that pass it; every candidate is then verified with the exact two-pointer distance of oracle.c, so hash
collisions cannot add an edge, and a true edge always satisfies the join, so none can be lost.
Follows the same definition as oracle.edges (breakfast.py:223-276 + sklearn _pairwise_fast.pyx:83-107 restated): an edge
is an unordered pair of distinct rows with |A xor B| <= max_dist.
"""
from __future__ import annotations

import numpy as np

import oracle


def _g(cols: np.ndarray) -> np.ndarray:
    """splitmix64 of the column id"""
    x = cols.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def edges_d1(indptr, indices, max_dist: int = 1, chunk_rows: int = 100_000):
    """(src, dst) int32, src < dst, sorted lexicographically — same contract as oracle.edges(..., max_dist) for
    max_dist in {0, 1}; rows must hold sorted unique column ids."""
    if max_dist not in (0, 1):
        raise ValueError("the join covers max_dist 0 and 1 only")
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    n = indptr.size - 1
    card = np.diff(indptr)
    with np.errstate(over="ignore"):
        gsum = np.concatenate([np.zeros(1, np.uint64), np.cumsum(_g(indices), dtype=np.uint64)])
        H = gsum[indptr[1:]] - gsum[indptr[:-1]]              # wraps mod 2^64; empty rows -> 0
    order = np.lexsort((card, H))                              # sorted by (H, card)
    Hs, cs = H[order], card[order]
    pairs = []
    # distance 0: equal (H, card) runs
    same = np.flatnonzero((Hs[1:] == Hs[:-1]) & (cs[1:] == cs[:-1]))
    run_start = None
    if same.size:
        starts = np.flatnonzero(np.r_[True, (Hs[1:] != Hs[:-1]) | (cs[1:] != cs[:-1])])
        ends = np.r_[starts[1:], n]
        for s, e in zip(starts[ends - starts > 1], ends[ends - starts > 1]):
            members = order[s:e]
            for i in range(len(members)):
                for j in range(i + 1, len(members)):
                    pairs.append((int(members[i]), int(members[j])))
    # distance 1: H(A) - g(x) == H(B) and |B| == |A| - 1
    if max_dist == 1 and n:
        mask = np.uint64((1 << 27) - 1)                           # membership table over the low hash bits
        present = np.zeros(1 << 27, dtype=bool)
        present[H & mask] = True
        for r0 in range(0, n, chunk_rows):
            r1 = min(n, r0 + chunk_rows)
            lo, hi = indptr[r0], indptr[r1]
            if hi == lo:
                continue
            rows = np.repeat(np.arange(r0, r1), card[r0:r1])
            with np.errstate(over="ignore"):
                key = H[rows] - _g(indices[lo:hi])
            maybe = np.flatnonzero(present[key & mask])           # one table lookup per entry rejects ~99 %
            key, rows = key[maybe], rows[maybe]
            a = np.searchsorted(Hs, key, "left")
            hit = np.flatnonzero(Hs[np.minimum(a, n - 1)] == key)   # ~ number of edges, not nnz
            a, key_hit, rows_hit = a[hit], key[hit], rows[hit]
            width = np.searchsorted(Hs, key_hit, "right") - a
            for w in range(int(width.max()) if hit.size else 0):   # w-th row of every equal-hash run (runs are short)
                sel = width > w
                p = a[sel] + w
                ok = cs[p] == card[rows_hit[sel]] - 1
                pairs.extend(zip(rows_hit[sel][ok].tolist(), order[p][ok].tolist()))
    # exact verification + canonical form
    if not pairs:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    arr = np.array(pairs, dtype=np.int64)
    arr = np.unique(np.sort(arr[arr[:, 0] != arr[:, 1]], axis=1), axis=0)
    keep = np.fromiter((oracle.distance(indptr, indices, int(x), int(y)) <= max_dist for x, y in arr), dtype=bool,
                       count=len(arr))
    arr = arr[keep].astype(np.int32)                            # np.unique left the pairs sorted lexicographically
    return arr[:, 0].copy(), arr[:, 1].copy()


def cluster_d1(indptr, indices, max_dist: int = 1):
    """labels (smallest row index per component) and the number of edges — oracle.cluster's contract"""
    src, dst = edges_d1(indptr, indices, max_dist)
    return oracle.components(len(indptr) - 1, src, dst), int(src.size)
