"""The second CPU oracle (oracle/hashjoin.py: exact edge set for max_dist <= 2 by deletion-neighbourhood hash joins)
against the first (oracle/oracle.c: all band pairs with the two-pointer distance)."""
import numpy as np
import pytest

import oracle
from oracle import hashjoin
from breakfast_b200 import synth


def _csr(rows, n_cols):
    indptr = np.zeros(len(rows) + 1, dtype=np.int64)
    indptr[1:] = np.cumsum([len(r) for r in rows])
    flat = [c for r in rows for c in sorted(r)]
    return indptr, np.asarray(flat, dtype=np.int32), n_cols


def _same(indptr, indices, d):
    a, b = hashjoin.edges(indptr, indices, d), oracle.edges(indptr, indices, d)
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("d", [0, 1, 2])
def test_small_adversarial_rows(d):
    rows = [[], [], [1], [1], [1, 2], [2], [1, 2, 3], [5, 6, 7], [5, 6], [5, 7], [6, 7], [5, 6, 7, 8], [9], [1, 2, 3],
            [3, 4], [4, 5], [1, 2, 3, 4, 5], [2, 3, 4]]
    indptr, indices, _ = _csr(rows, 10)
    assert _same(indptr, indices, d)


@pytest.mark.parametrize("n,seed", [(1, 1), (2, 2), (500, 3), (3000, 5), (20000, 13)])
@pytest.mark.parametrize("d", [0, 1, 2])
def test_synthetic_profiles(n, seed, d):
    indptr, indices, _ = synth.generate(n, seed=seed).csr()
    assert _same(indptr, indices, d)
    labels, n_edges = hashjoin.cluster(indptr, indices, d)
    want, want_edges = oracle.cluster(indptr, indices, d)
    assert n_edges == want_edges and np.array_equal(labels, want)


def test_random_subsets_with_duplicates():
    rng = np.random.default_rng(4)
    rows = []
    for _ in range(400):
        base = sorted(rng.choice(60, size=int(rng.integers(0, 12)), replace=False).tolist())
        rows.append(base)
        for _ in range(int(rng.integers(0, 4))):
            r = list(base)
            if r and rng.random() < 0.5:
                r.pop(int(rng.integers(len(r))))
            elif rng.random() < 0.7:
                r = sorted(set(r) | {int(rng.integers(60))})
            rows.append(r)
    indptr, indices, _ = _csr(rows, 60)
    for d in (0, 1, 2):
        assert _same(indptr, indices, d)


def test_config2_size_at_distance_two():
    """100 000 profiles (the size of BASELINE config 2), max_dist 2: both oracles list the same 194 154 edges"""
    indptr, indices, _ = synth.generate(100000, seed=2).csr()
    src, dst = hashjoin.edges(indptr, indices, 2)
    ws, wd = oracle.edges(indptr, indices, 2)
    assert src.size == 194154 and np.array_equal(src, ws) and np.array_equal(dst, wd)


def test_rejects_larger_distances():
    indptr, indices, _ = _csr([[1], [2]], 3)
    with pytest.raises(ValueError):
        hashjoin.edges(indptr, indices, 3)
    with pytest.raises(ValueError):
        hashjoin.edges_d1(indptr, indices, 2)
