"""Host-side mirror of the reference module src/breakfast/breakfast.py (same callables, same
argument meaning, same error behaviour), with the distance / neighbour / connected-component work
delegated to the CUDA library (breakfast_b200._native).

Function-by-function correspondence (reference file:line):
  read_input                  breakfast.py:16-29
  write_output                breakfast.py:32-69
  collapse_duplicates         breakfast.py:72-79
  cluster                     breakfast.py:82-89
  filter_features             breakfast.py:116-190
  sparse_feature_matrix       breakfast.py:193-215
  cluster_features            breakfast.py:279-340  (get_neighbours_batch 223-276, _to_graph 93-113 and
                                                     networkx connected_components run on the GPU)
  cluster_identical_features  breakfast.py:343-364
"""
from __future__ import annotations

import os
import re
import sys
from itertools import chain

import numpy as np
import pandas as pd
from scipy.sparse import csr_matrix

from . import cache as ca
from . import engine

import time

# seconds the last cluster_features call spent inside the engine (H2D, kernels, D2H) - the rest of its time is host work
LAST_ENGINE_TIMINGS: dict = {}


# --------------------------------------------------------------------------------------------
# I/O
# --------------------------------------------------------------------------------------------
def read_input(input_file, sep, id_col, feature_col):
    """TSV -> DataFrame[id, feature] (both str).  Duplicate ids are an error; a missing profile
    cell becomes the empty profile.  pandas' default NA parsing applies, as in the reference."""
    options = dict(sep=sep, usecols=[id_col, feature_col], dtype={id_col: str, feature_col: str})
    table = None
    if os.environ.get("BREAKFAST_B200_READER", "pyarrow") == "pyarrow":
        # multi-threaded Arrow reader behind the same pandas call (same NA rules, same dtypes; ten times faster on
        # a million lines); anything it refuses - or parses into a different shape - goes to pandas' own C parser
        try:
            table = pd.read_table(input_file, engine="pyarrow", **options)
            if list(table.columns) != [c for c in table.columns if c in (id_col, feature_col)] or table.shape[1] != 2:
                table = None
        except Exception:
            table = None
    if table is None:
        table = pd.read_table(input_file, **options)
    table = table.rename(columns={id_col: "id", feature_col: "feature"})
    repeated = table["id"].duplicated()
    if repeated.any():
        names = ", ".join(table.loc[repeated, "id"].unique())
        raise ValueError(f"Duplicate sequence identifiers found: {names}")
    table["feature"] = table["feature"].fillna("")
    print(f"Number of sequences: {table.shape[0]}")
    return table


def write_output(meta_nodups, meta_original, outdir):
    """One line per input sequence, input order, clusters renumbered 1..K by first appearance,
    unclustered sequences get an empty field; written to <outdir>/clusters.tsv."""
    id_groups = meta_nodups["id"].tolist()
    group_sizes = np.fromiter(map(len, id_groups), dtype=np.int64, count=len(id_groups))
    per_profile = np.empty(len(id_groups), dtype=object)
    per_profile[:] = meta_nodups["cluster_id"].tolist()
    flat_ids = list(chain.from_iterable(id_groups))
    per_sequence = np.repeat(per_profile, group_sizes)
    original_ids = meta_original["id"]
    # back to the order of the input file.  Without repeated profiles the groups are already in that order (first
    # appearance): then the input's own id column is written and the join on the ids is skipped.
    if flat_ids == original_ids.tolist():
        ids_out, raw, expanded = original_ids, per_sequence, None
    else:
        expanded = pd.DataFrame({"id": flat_ids, "cluster_id": per_sequence})
        expanded = expanded.set_index("id").reindex(index=original_ids).reset_index()
        ids_out, raw = expanded["id"], expanded["cluster_id"].to_numpy(dtype=object)

    # engine labels -> 1..K in order of first appearance in the input
    clustered = ~pd.isna(raw)
    renumbered = np.empty(raw.size, dtype=object)
    renumbered[:] = pd.NA
    if clustered.any():
        codes, _ = pd.factorize(raw[clustered], sort=False)
        renumbered[clustered] = (codes + 1).tolist()

    if raw.size != meta_original.shape[0]:
        raise RuntimeError("Output row count differs from input row count")

    outdir.mkdir(parents=True, exist_ok=True)
    if not _write_clusters_arrow(ids_out, renumbered, clustered, outdir / "clusters.tsv"):
        frame = pd.DataFrame({"id": ids_out.to_numpy() if expanded is None else expanded["id"], "cluster_id": renumbered})
        frame[["id", "cluster_id"]].to_csv(outdir / "clusters.tsv", sep="\t", index=False)


def _write_clusters_arrow(ids, renumbered, clustered, path) -> bool:
    """clusters.tsv through pyarrow's CSV writer - the bytes pandas' to_csv(sep="\\t", index=False) writes (reference
    breakfast.py:64-69), several times faster at 10^6 rows.  Only for the plain case: string ids without missing values
    and without a character that the csv module would quote (tab, double quote, line break); anything else returns False
    and the caller takes the pandas path, which is the reference's own."""
    try:
        import pyarrow as pa
        import pyarrow.compute as pc
        import pyarrow.csv as pcsv
        arr = pa.array(ids)
        if isinstance(arr, pa.ChunkedArray):
            arr = arr.combine_chunks()
        if not (pa.types.is_string(arr.type) or pa.types.is_large_string(arr.type)) or arr.null_count:
            return False
        if pc.any(pc.match_substring_regex(arr, '["\t\n\r]')).as_py():
            return False
        values = np.zeros(len(renumbered), dtype=np.int64)
        values[clustered] = renumbered[clustered].astype(np.int64)
        col = pa.array(values, mask=~clustered)
        table = pa.table({"id": arr, "cluster_id": col})
        with open(path, "wb") as fh:
            fh.write(b"id\tcluster_id\n")
            pcsv.write_csv(table, fh, pcsv.WriteOptions(include_header=False, delimiter="\t", quoting_style="none"))
        return True
    except Exception:   # any surprise in the fast writer: the pandas path decides
        return False


# --------------------------------------------------------------------------------------------
# dedup
# --------------------------------------------------------------------------------------------
def collapse_duplicates(meta):
    """Group sequences with an identical (filtered) profile string; first-appearance order; the id
    column becomes a tuple of the member ids."""
    features = meta["feature"]
    print(f"Number of duplicates: {features.duplicated().sum()}")
    codes, _ = pd.factorize(features, sort=False)
    present = codes >= 0  # a NaN profile belongs to no group (pandas groupby drops it as well)
    positions = np.flatnonzero(present)
    order = positions[np.argsort(codes[present], kind="stable")]
    counts = np.bincount(codes[present])
    bounds = np.concatenate(([0], np.cumsum(counts)))
    ids = meta["id"].to_numpy(dtype=object)[order]
    grouped_ids = np.empty(len(counts), dtype=object)
    for g in range(len(counts)):
        grouped_ids[g] = tuple(ids[bounds[g]:bounds[g + 1]])
    first_rows = order[bounds[:-1]]
    meta_nodups = pd.DataFrame(
        {
            "id": pd.Series(grouped_ids, dtype=object),
            "feature": features.iloc[first_rows].reset_index(drop=True),
        }
    )
    print(f"Number of unique sequences: {meta_nodups.shape[0]}")
    return meta_nodups


# --------------------------------------------------------------------------------------------
# feature filtering
# --------------------------------------------------------------------------------------------
# (substitution, insertion, deletion) per --var-type.  A substitution pattern with a group captures
# the genome position; amino-acid substitutions have none and are therefore never trimmed.
_PATTERNS = {
    "covsonar_dna": (r"^[A-Z](\d+)[A-Z]$", r"^.*[A-Z][A-Z]$", r"^del:\d+:\d+$"),
    "covsonar_aa": (r"^[a-zA-Z0-9]+:[A-Z]\d+[A-Z]$", r"^[a-zA-Z0-9]+:[A-Z]\d+[A-Z][A-Z]+$",
                    r"^[a-zA-Z0-9]+:del:\d+:\d+$"),
    "nextclade_dna": (r"^[A-Z](\d+)[A-Z]$", r"^\d+:[A-Z]+$", r"^\d+(-\d+)?$"),
    "nextclade_aa": (r"^[a-zA-Z0-9]+:[A-Z]\d+[A-Z*]$", r"^$", r"^[a-zA-Z0-9]+:[A-Z]\d+-$"),
}
_KEEP, _DROP, _INVALID = 0, 1, 2


def _token_classifier(feature_type, skip_ins, skip_del, trim_start, trim_end, reference_length):
    """token -> _KEEP / _DROP / _INVALID under the reference's rules (breakfast.py:135-184)."""
    if feature_type not in _PATTERNS:
        print(f"The feature type (--var-type) you chose is not supported: '{feature_type}'")
        sys.exit(1)
    sub_re, ins_re, del_re = (re.compile(p) for p in _PATTERNS[feature_type])
    upper_cut = reference_length - trim_end

    def classify(tok):
        hit = sub_re.match(tok)
        if hit:
            if hit.lastindex:
                pos = int(hit.group(1))
                if pos <= trim_start or pos >= upper_cut:
                    return _DROP
            return _KEEP
        if ins_re.match(tok):
            return _DROP if skip_ins else _KEEP
        if del_re.match(tok):
            return _DROP if skip_del else _KEEP
        return _INVALID

    return classify


def filter_features(
    features,
    feature_sep,
    feature_type,
    skip_ins,
    skip_del,
    trim_start,
    trim_end,
    reference_length,
):
    """Drop trimmed substitutions, skipped indels, unparsable and empty tokens from every profile.

    With no filter active the input is handed back untouched (so unparsable tokens survive), which
    is also what happens for every non-DNA --var-type coming from the CLI.  Classification order is
    substitution, insertion, deletion; trimming applies to substitutions only and is inclusive at
    both ends.  The verdict per distinct token is memoised: a data set has ~10^5 distinct tokens
    but ~10^2 tokens per profile.
    """
    if not (skip_del or skip_ins or trim_start > 0 or trim_end > 0):
        return features

    is_raw = feature_type == "raw"
    classify = None if is_raw else _token_classifier(feature_type, skip_ins, skip_del, trim_start, trim_end,
                                                     reference_length)

    verdicts: dict = {}
    out = []
    for profile in features:
        kept = []
        for tok in profile.split(feature_sep):
            if not is_raw:
                v = verdicts.get(tok)
                if v is None:
                    v = verdicts[tok] = classify(tok)
                if v == _INVALID:
                    print(f"Skipping invalid feature: '{tok}'")
                    continue
                if v == _DROP:
                    continue
            if tok:
                kept.append(tok)
        out.append(feature_sep.join(kept))
    return out


# --------------------------------------------------------------------------------------------
# vectorisation
# --------------------------------------------------------------------------------------------
def sparse_feature_matrix(features, feature_sep):
    """Profiles -> scipy CSR (sequences x vocabulary, int64 ones; a token repeated inside a profile
    appears twice).  Vocabulary ids follow first appearance; empty tokens are ignored."""
    indptr, indices, vocab = engine.tokenise(features, feature_sep)
    if not vocab:
        # the reference lets scipy infer the width and scipy refuses an all-empty matrix
        raise ValueError("unable to infer matrix dimensions")
    data = np.ones(indices.size, dtype=np.int64)
    return csr_matrix((data, indices, indptr), shape=(len(indptr) - 1, len(vocab)), dtype=int)


# --------------------------------------------------------------------------------------------
# clustering
# --------------------------------------------------------------------------------------------
def cluster(meta_nodups, sep2, max_dist, min_cluster_size, input_cache, output_cache, pre=None):
    """`pre` (not in the reference): the CSR dict of hostfast.prepare for exactly these rows, so that the profiles
    are not tokenised a second time."""
    if max_dist == 0:
        return cluster_identical_features(meta_nodups, min_cluster_size)
    return cluster_features(meta_nodups, sep2, max_dist, min_cluster_size, input_cache, output_cache, pre=pre)


def _assign_cluster_ids(meta, labels, min_cluster_size):
    """Component labels (smallest member row) -> running cluster ids for components whose total
    number of *sequences* reaches min_cluster_size (breakfast.py:329-338).  Components are numbered
    by their smallest row; write_output renumbers by first appearance in the input anyway."""
    n = len(meta)
    seq_counts = np.fromiter((len(t) for t in meta["id"].tolist()), dtype=np.int64, count=n)
    comp_size = np.bincount(labels, weights=seq_counts, minlength=n).astype(np.int64)
    big_enough = comp_size[labels] >= min_cluster_size
    is_root = labels == np.arange(n)
    kept_roots = is_root & big_enough
    new_id = np.cumsum(kept_roots)  # id of a kept root = its rank among kept roots
    column = np.empty(n, dtype=object)
    column[:] = pd.NA
    column[big_enough] = new_id[labels[big_enough]].tolist()
    meta["cluster_id"] = column
    return int(kept_roots.sum())


def cluster_features(meta, feature_sep, max_dist, min_cluster_size, input_cache, output_cache, pre=None):
    if pre is not None and pre["n"] == len(meta):
        # the native host path (hostfast.prepare) already tokenised the unique profiles
        if pre["n_vocab"] == 0:
            raise ValueError("unable to infer matrix dimensions")   # what scipy says to the reference here
        meta["n_features"] = np.diff(pre["token_indptr"])
        n = len(meta)
        indptr, indices, n_cols = pre["bin_indptr"], pre["bin_indices"], pre["n_cols"]
    else:
        feat_matrix = sparse_feature_matrix(meta["feature"], feature_sep)
        meta["n_features"] = feat_matrix.sum(axis=1)
        n = feat_matrix.shape[0]
        # strictly binary rows for the device (repeated tokens thermometer-coded)
        indptr, indices, n_cols = engine.thermometer_binarise(
            feat_matrix.indptr.astype(np.int64), feat_matrix.indices.astype(np.int64), feat_matrix.shape[1]
        )

    cached = None
    try:
        cache = ca.load(input_cache, max_dist)
        feature_map = ca.map_features(cache["meta"]["feature"], meta["feature"])
        list_indptr, list_members = ca.update_neighbours_csr(cache["neigh"], feature_map)
        new_rows = np.array(ca.find_new(feature_map)).astype(int)
        cached = (list_indptr, list_members, new_rows)
    except (UnboundLocalError, TypeError):
        print(
            "Imported cached results are not available. "
            "Distance matrix of complete dataset will be calculated."
        )

    want_edges = bool(output_cache)
    t_engine = time.perf_counter()
    if cached is None:
        result = engine.components_full(indptr, indices, n_cols, max_dist, want_edges=want_edges)
        list_rows = np.arange(n)
        kept_lists = []
    else:
        list_indptr, list_members, new_rows = cached
        result = engine.components_incremental(
            indptr, indices, n_cols, max_dist, new_rows, list_indptr, list_members, want_edges=want_edges
        )
        list_rows = new_rows
        kept_lists = [list_members[list_indptr[i]:list_indptr[i + 1]].astype(np.int64)
                      for i in range(len(list_indptr) - 1)] if want_edges else []
    LAST_ENGINE_TIMINGS.clear()
    LAST_ENGINE_TIMINGS.update(engine=time.perf_counter() - t_engine)

    if output_cache:
        src, dst = result.edges
        neigh = kept_lists + engine.adjacency_lists(n, list_rows, src, dst)
        ca.save(output_cache, neigh, meta, max_dist)

    print("Create graph and recover connected components")
    print("Save clusters")
    n_clusters = _assign_cluster_ids(meta, result.labels.astype(np.int64), min_cluster_size)
    print(f"Number of clusters found: {n_clusters}")
    return meta


def cluster_identical_features(meta, min_cluster_size):
    """max_dist == 0: every unique profile string is its own cluster (if it has enough sequences).
    No matrix, no distances, no cache — host only, like the reference."""
    print("Skip sparse matrix calculation since max-dist = 0")
    n = len(meta)
    seq_counts = np.fromiter((len(t) for t in meta["id"].tolist()), dtype=np.int64, count=n)
    keep = seq_counts >= min_cluster_size
    column = np.empty(n, dtype=object)
    column[:] = pd.NA
    column[keep] = np.arange(1, int(keep.sum()) + 1).tolist()
    meta["cluster_id"] = column
    print(f"Number of clusters found: {int(keep.sum())}")
    return meta
