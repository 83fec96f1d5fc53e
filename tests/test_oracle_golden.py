"""Pins the CPU oracle (oracle/ref_port.py, oracle/oracle.c) against the reference: its own
expected_clusters_*.tsv golden files and outputs produced by running the unmodified reference on
seeded inputs (tests/golden/make_golden.py).  CPU only."""
from io import StringIO

import numpy as np
import pandas as pd
import pytest

import oracle
from oracle import ref_port
from tests import helpers
from tests.helpers import GOLDEN


def _kwargs(case_opts):
    kw = dict(case_opts)
    return kw


def _check(case, expected_rel, text):
    helpers.assert_matches(case, expected_rel, text)


@pytest.mark.parametrize("case", helpers.cases("plain"), ids=lambda c: c["name"])
def test_ref_port_plain(case):
    text, _ = ref_port.run_file(GOLDEN / case["input"], **_kwargs(case["opts"]))
    _check(case, case["expected"], text)


@pytest.mark.parametrize("case", helpers.cases("plain"), ids=lambda c: c["name"])
def test_c_core_matches_goldens(case):
    """the same pipeline with oracle.c doing the distance part (used for 1e5-scale parity tests)"""
    text, _ = ref_port.run_file(GOLDEN / case["input"], core="c", **_kwargs(case["opts"]))
    _check(case, case["expected"], text)


@pytest.mark.parametrize("case", [c for c in helpers.cases("cached") if "cache_from" in c], ids=lambda c: c["name"])
def test_ref_port_cached(case):
    first = case["cache_from"]
    _, cache = ref_port.run_file(GOLDEN / first["input"], want_cache=True, **_kwargs(first["opts"]))
    assert cache is not None
    text, _ = ref_port.run_file(GOLDEN / case["input"], cache=cache, **_kwargs(case["opts"]))
    _check(case, case["expected"], text)


@pytest.mark.parametrize("case", helpers.cases("chain"), ids=lambda c: c["name"])
def test_ref_port_cache_chain(case):
    cache = None
    for step in case["chain"]:
        text, cache = ref_port.run_file(GOLDEN / step["input"], cache=cache, want_cache=True, **_kwargs(case["opts"]))
        _check(case, step["expected"], text)


def test_ghost_lists_are_visible_in_the_goldens():
    """The cached run of step 1 must differ from a fresh run of the same input: that is
    what makes the cache goldens a real test of the ghost-list semantics."""
    syn = GOLDEN / "synthetic"
    assert (syn / "cache_d1_step1.expected.tsv").read_text() != (syn / "cache_d1_step1_fresh.expected.tsv").read_text()
    cached = pd.read_table(syn / "cache_d1_step1.expected.tsv").set_index("id")["cluster_id"]
    fresh = pd.read_table(syn / "cache_d1_step1_fresh.expected.tsv").set_index("id")["cluster_id"]
    assert cached["ghostA1"] == cached["ghostB1"]      # chained through the vanished profile X
    assert fresh["ghostA1"] != fresh["ghostB1"]        # distance 2 apart without it


@pytest.mark.parametrize("max_dist", [1, 2, 3])
def test_c_oracle_matches_ref_port(max_dist):
    """oracle.c (merge + union-find) == ref_port (scikit-learn + networkx) on seeded profiles."""
    from breakfast_b200 import synth
    prof = synth.generate(700, seed=21 + max_dist, with_mult=False)
    feats = prof.features("covsonar_dna", " ")
    uniq, codes, mult = ref_port.dedup(feats)
    X = ref_port.count_matrix(uniq, " ")
    lists = ref_port.neighbour_lists(X, max_dist)
    want = ref_port.components(len(uniq), lists)
    indptr, indices, n_cols = ref_port.binary_csr(uniq, " ")
    got, n_edges = oracle.cluster(indptr, indices, max_dist)
    assert np.array_equal(got, want)
    # the edge set too: every list is {row} + its neighbours
    src, dst = oracle.edges(indptr, indices, max_dist)
    D = {}
    for a, b in zip(src.tolist(), dst.tolist()):
        D[(a, b)] = oracle.distance(indptr, indices, a, b)
        assert D[(a, b)] <= max_dist
    from sklearn.metrics import pairwise_distances
    dense = pairwise_distances(X, metric="manhattan")
    iu = np.triu_indices(len(uniq), 1)
    want_edges = set(zip(iu[0][dense[iu] <= max_dist].tolist(), iu[1][dense[iu] <= max_dist].tolist()))
    assert set(D) == want_edges
    assert n_edges == len(want_edges)


def test_c_oracle_repeated_tokens_are_counts():
    """'A1T A1T B' vs 'A1T B': L1 on counts = 1; thermometer coding keeps that."""
    uniq = ["A1T A1T B", "A1T B", "B A1T", "A1T A1T A1T B"]
    indptr, indices, n_cols = ref_port.binary_csr(uniq, " ")
    assert oracle.distance(indptr, indices, 0, 1) == 1
    assert oracle.distance(indptr, indices, 1, 2) == 0
    assert oracle.distance(indptr, indices, 0, 3) == 1
    assert oracle.distance(indptr, indices, 1, 3) == 2
    from sklearn.metrics import pairwise_distances
    dense = pairwise_distances(ref_port.count_matrix(uniq, " "), metric="manhattan")
    for a in range(4):
        for b in range(4):
            assert dense[a, b] == oracle.distance(indptr, indices, a, b)


def test_c_oracle_rectangle_and_lists():
    from breakfast_b200 import synth
    indptr, indices, _ = synth.generate(500, seed=33).csr()
    src, dst = oracle.edges(indptr, indices, 2)
    q = np.arange(0, 500, 3, dtype=np.int32)
    qs = set(q.tolist())
    s2, d2 = oracle.edges(indptr, indices, 2, queries=q)
    assert set(zip(s2.tolist(), d2.tolist())) == {(a, b) for a, b in zip(src.tolist(), dst.tolist()) if a in qs or b in qs}
    # lists chain their members
    lab = oracle.components(6, None, None, np.array([0, 3, 5]), np.array([5, 1, 3, 0, 2]))
    assert lab.tolist() == [0, 1, 0, 1, 4, 1]
