"""Python restatement ("port") of the reference pipeline, structured like the reference so that it
exercises the same third-party kernels.  TEST INFRASTRUCTURE / CPU BASELINE ONLY — the product never
imports this.

Restated from rki-mf1/breakfast v0.4.6 (paths relative to /root/reference):
  read            src/breakfast/breakfast.py:16-29
  filter          src/breakfast/breakfast.py:116-190
  dedup           src/breakfast/breakfast.py:72-79
  count matrix    src/breakfast/breakfast.py:193-215
  neighbours      src/breakfast/breakfast.py:223-276, loop 314-318 — scikit-learn 1.9.0
                  pairwise_distances_chunked(metric="manhattan") -> _sparse_manhattan
                  (sklearn/metrics/_pairwise_fast.pyx:34-107), the reference's hot loop
  cache           src/breakfast/cache.py:51-112 + breakfast.py:294-304 (ghost lists included)
  components      src/breakfast/breakfast.py:93-113,325-326 (networkx 3.6.1)
  size filter     src/breakfast/breakfast.py:329-338
  output          src/breakfast/breakfast.py:32-69

Pinned by tests/test_oracle_golden.py against the reference's expected_clusters_*.tsv files and
against outputs of the reference itself on synthetic inputs (tests/golden/make_golden.py).
"""
from __future__ import annotations

import re
from itertools import chain

import numpy as np

RULES = {
    "covsonar_dna": (r"^[A-Z](\d+)[A-Z]$", r"^.*[A-Z][A-Z]$", r"^del:\d+:\d+$"),
    "covsonar_aa": (r"^[a-zA-Z0-9]+:[A-Z]\d+[A-Z]$", r"^[a-zA-Z0-9]+:[A-Z]\d+[A-Z][A-Z]+$", r"^[a-zA-Z0-9]+:del:\d+:\d+$"),
    "nextclade_dna": (r"^[A-Z](\d+)[A-Z]$", r"^\d+:[A-Z]+$", r"^\d+(-\d+)?$"),
    "nextclade_aa": (r"^[a-zA-Z0-9]+:[A-Z]\d+[A-Z*]$", r"^$", r"^[a-zA-Z0-9]+:[A-Z]\d+-$"),
}


def read(path, sep="\t", id_col="accession", feature_col="dna_profile"):
    import pandas as pd
    t = pd.read_table(path, sep=sep, usecols=[id_col, feature_col], dtype={id_col: str, feature_col: str})
    ids = t[id_col].tolist()
    seen = set()
    for i in ids:
        if i in seen:
            raise ValueError("Duplicate sequence identifiers found")
        seen.add(i)
    feats = ["" if isinstance(f, float) else f for f in t[feature_col].tolist()]
    return ids, feats


def filter_profiles(feats, sep, var_type, skip_ins, skip_del, trim_start, trim_end, ref_len):
    if not (skip_del or skip_ins or trim_start > 0 or trim_end > 0):
        return list(feats)
    if var_type != "raw":
        sub, ins, dele = (re.compile(p) for p in RULES[var_type])
    out = []
    for f in feats:
        keep = []
        for tok in f.split(sep):
            if var_type != "raw":
                m = sub.match(tok)
                if m:
                    if m.lastindex:
                        pos = int(m.group(1))
                        if pos <= trim_start or pos >= ref_len - trim_end:
                            continue
                elif ins.match(tok):
                    if skip_ins:
                        continue
                elif dele.match(tok):
                    if skip_del:
                        continue
                else:
                    continue
            if tok:
                keep.append(tok)
        out.append(sep.join(keep))
    return out


def dedup(feats):
    """unique strings in first-appearance order, code of every sequence, sequences per unique string"""
    first = {}
    codes = np.empty(len(feats), dtype=np.int64)
    for i, f in enumerate(feats):
        codes[i] = first.setdefault(f, len(first))
    uniq = list(first)
    return uniq, codes, np.bincount(codes, minlength=len(uniq))


def count_matrix(uniq, sep):
    """scipy CSR of token COUNTS (duplicates summed), vocabulary by first appearance."""
    from scipy.sparse import csr_matrix
    vocab, indptr, indices = {}, [0], []
    for f in uniq:
        for tok in f.split(sep):
            if tok:
                indices.append(vocab.setdefault(tok, len(vocab)))
        indptr.append(len(indices))
    X = csr_matrix((np.ones(len(indices), dtype=np.int64), np.array(indices, dtype=np.int64),
                    np.array(indptr, dtype=np.int64)), shape=(len(uniq), max(len(vocab), 1)))
    X.sum_duplicates()
    return X


def binary_csr(uniq, sep):
    """Strictly binary CSR whose set distance equals L1 on counts (k-th repeat of a token = own column)."""
    vocab, indptr, indices = {}, [0], []
    for f in uniq:
        seen = {}
        row = []
        for tok in f.split(sep):
            if tok:
                k = seen.get(tok, 0)
                seen[tok] = k + 1
                row.append(vocab.setdefault((tok, k), len(vocab)))
        indices.extend(sorted(row))
        indptr.append(len(indices))
    return np.array(indptr, dtype=np.int64), np.array(indices, dtype=np.int32), len(vocab)


def neighbour_lists(X, max_dist, select=None, stats=None):
    """The reference's neighbour search: for every distinct row cardinality q, the band
    |card - q| <= max_dist of the batch against the band of all rows, L1 via scikit-learn."""
    from sklearn.metrics import pairwise_distances_chunked
    card = np.asarray(X.sum(axis=1)).ravel()
    if select is None:
        batch_rows = np.arange(X.shape[0])
    else:
        batch_rows = np.asarray(select, dtype=np.int64)
        if batch_rows.size == 0:
            return []
    qs = list(dict.fromkeys(card[batch_rows].tolist()))
    lists = []
    for q in qs:
        band_all = np.flatnonzero(np.isclose(card, q, atol=max_dist))
        band_batch = batch_rows[np.isclose(card[batch_rows], q, atol=max_dist)]
        if stats is not None:
            stats["evaluations"] = stats.get("evaluations", 0) + int(band_batch.size) * int(band_all.size)
        gen = pairwise_distances_chunked(
            X=X[band_batch, :], Y=X[band_all, :], metric="manhattan", n_jobs=1,
            reduce_func=lambda D, start: [np.flatnonzero(d <= max_dist) for d in D])
        for hits in chain.from_iterable(gen):
            lists.append(band_all[hits])
    return lists


def components(n, lists):
    """labels[i] = smallest row of i's component; rows that appear in no list get -1 (the reference
    never sees them as graph nodes)."""
    import networkx
    G = networkx.Graph()
    for part in lists:
        part = [int(x) for x in part]
        G.add_nodes_from(part)
        G.add_edges_from(zip(part[:-1], part[1:]))
    labels = np.full(n, -1, dtype=np.int64)
    for comp in networkx.connected_components(G):
        members = sorted(comp)
        labels[members] = members[0]
    return labels


def remap_cached_lists(cache, uniq):
    """cache = {"features": [...], "neigh": [...]} -> (re-indexed surviving lists, rows new to the cache)."""
    where = {f: i for i, f in enumerate(uniq)}
    c2n = [where.get(f) for f in cache["features"]]
    kept = []
    for lst in cache["neigh"]:
        m = [c2n[int(x)] for x in lst]
        m = [x for x in m if x is not None]
        if m:
            kept.append(m)
    known = set(cache["features"])
    new_rows = [i for i, f in enumerate(uniq) if f not in known]
    return kept, new_rows


def lists_from_edges(n, rows, src, dst):
    """neighbour lists (ascending, the row itself included - what the reference's radius query returns,
    breakfast.py:226-228) of the rows `rows` (ascending) from an undirected edge list"""
    rows = np.asarray(rows, dtype=np.int64)
    a = np.concatenate([src, dst, rows]).astype(np.int64)
    b = np.concatenate([dst, src, rows]).astype(np.int64)
    wanted = np.zeros(n, dtype=bool)
    wanted[rows] = True
    keep = wanted[a]
    a, b = a[keep], b[keep]
    order = np.lexsort((b, a))
    a, b = a[order], b[order]
    bounds = np.searchsorted(a, np.concatenate([rows, [n]]))
    return [b[bounds[i]:bounds[i + 1]] for i in range(len(rows))]


def components_of_lists(n, lists):
    """components(n, lists) without networkx (for 10^6 lists): every list chains its members (breakfast.py:103-113);
    label = smallest row of the component; every row must occur in some list"""
    import oracle
    lengths = np.fromiter((len(l) for l in lists), dtype=np.int64, count=len(lists))
    li = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
    lm = np.concatenate([np.asarray(l, dtype=np.int32) for l in lists]) if len(lists) else np.zeros(0, np.int32)
    seen = np.zeros(n, dtype=bool)
    seen[lm] = True
    assert seen.all(), "a row occurs in no neighbour list"
    return oracle.components(n, None, None, li, lm).astype(np.int64)


def cluster_table(ids, feats, sep2=" ", max_dist=1, min_cluster_size=2, cache=None, want_cache=False, core="sklearn"):
    """ids/filtered profile strings -> (text of clusters.tsv, cache dict or None).
    core="sklearn": the reference's own neighbour search (slow, hours beyond ~1e5 profiles);
    core="c": oracle.c brute force for the distance part (no cache support) — same results, pinned by
    tests/test_oracle_golden.py::test_c_oracle_matches_ref_port and test_c_core_matches_goldens;
    core="hashjoin": oracle/hashjoin.py for the distance part (max_dist <= 2, 10^6 profiles in a minute), cache
    supported: cached lists re-indexed as in cache.py:51-71, new rows queried against all rows (breakfast.py:294-304)."""
    uniq, codes, mult = dedup(feats)
    n = len(uniq)
    if max_dist == 0:
        labels = np.arange(n)
        lists = None
    elif core == "hashjoin":
        from oracle import hashjoin
        indptr, indices, _ = binary_csr(uniq, sep2)
        src, dst = hashjoin.edges(indptr, indices, max_dist)
        if cache is not None and cache["max_dist"] == max_dist:
            kept, new_rows = remap_cached_lists(cache, uniq)
            lists = kept + lists_from_edges(n, sorted(new_rows), src, dst)
        else:
            lists = lists_from_edges(n, np.arange(n), src, dst)
        labels = components_of_lists(n, lists)
    elif core == "c":
        import oracle
        assert cache is None and not want_cache
        indptr, indices, _ = binary_csr(uniq, sep2)
        labels, _ = oracle.cluster(indptr, indices, max_dist)
        labels = labels.astype(np.int64)
        lists = None
    else:
        X = count_matrix(uniq, sep2)
        if cache is not None and cache["max_dist"] == max_dist:
            kept, new_rows = remap_cached_lists(cache, uniq)
            lists = kept + neighbour_lists(X, max_dist, select=np.array(sorted(new_rows), dtype=np.int64))
        else:
            lists = neighbour_lists(X, max_dist)
        labels = components(n, lists)
    size = np.zeros(n + 1, dtype=np.int64)
    np.add.at(size, np.where(labels >= 0, labels, n), mult)
    seq_label = labels[codes]
    ok = (seq_label >= 0) & (size[np.where(seq_label >= 0, seq_label, n)] >= min_cluster_size)
    order, out = {}, ["id\tcluster_id"]
    for i, sid in enumerate(ids):
        if ok[i]:
            cid = order.setdefault(int(seq_label[i]), len(order) + 1)
            out.append(f"{sid}\t{cid}")
        else:
            out.append(f"{sid}\t")
    text = "\n".join(out) + "\n"
    new_cache = None
    if want_cache and lists is not None:
        new_cache = {"max_dist": max_dist, "features": uniq, "neigh": [list(map(int, l)) for l in lists]}
    return text, new_cache


def run_file(path, sep="\t", id_col="accession", clust_col="dna_profile", var_type="covsonar_dna", sep2=" ",
             max_dist=1, min_cluster_size=2, trim_start=264, trim_end=228, reference_length=29903,
             skip_del=True, skip_ins=True, cache=None, want_cache=False, core="sklearn"):
    """The CLI's five steps (console.py:152-170) on one input table."""
    if var_type not in ("covsonar_dna", "nextclade_dna"):
        trim_start = trim_end = 0
        skip_del = skip_ins = False
    ids, feats = read(path, sep, id_col, clust_col)
    feats = filter_profiles(feats, sep2, var_type, skip_ins, skip_del, trim_start, trim_end, reference_length)
    return cluster_table(ids, feats, sep2, max_dist, min_cluster_size, cache, want_cache, core)
