"""Result cache — host-side mirror of the reference module src/breakfast/cache.py.

File format (unchanged, so caches are interchangeable with the reference): gzip + pickle protocol 2
of {"max_dist": int, "version": str, "neigh": list of int arrays, "meta": DataFrame[id, feature]}
(cache.py:18-32).  Semantics kept bit-for-bit, including the "ghost list" behaviour: a cached
neighbour list does not record its query row, so when that profile disappears its surviving
members stay chained together (cache.py:51-71 + breakfast.py:93-113).

The re-indexing is array code instead of the reference's per-element Python loops.
"""
from __future__ import annotations

import _pickle as cPickle
import gzip

import numpy as np
import pandas as pd

from . import __version__


def load(input_file, max_dist):
    """Read and validate a cache written by `save` (or by the reference).  `input_file=None`
    raises TypeError from gzip, which the caller treats as "no cache" exactly like the reference."""
    with gzip.open(input_file, "rb") as handle:
        print("Import from pickle file")
        cache = cPickle.load(handle)
    validate(cache, max_dist, __version__)
    return cache


def save(output_file, neigh, meta, max_dist):
    try:
        print("Export results as pickle")
        frame = meta[["id", "feature"]]
        frame.attrs = {}   # the file holds what the reference's holds: no private metadata rides along
        payload = {
            "max_dist": max_dist,
            "version": __version__,
            "neigh": neigh,
            "meta": frame,
        }
        output_file.parent.mkdir(parents=True, exist_ok=True)
        # gzip level 6 instead of the module's default 9: the same format (the reference reads it with gzip.open), a file
        # 1 % larger, written 2.5 times faster - at 10^6 profiles the difference is a minute of the run
        with gzip.open(output_file, "wb", compresslevel=6) as handle:
            cPickle.dump(payload, handle, 2)
    except TypeError:
        print("Export of pickle was not succesfull")


def validate(cache, max_dist, version):
    """A cache computed for another max_dist is useless: signal it the way the reference does
    (UnboundLocalError, caught by cluster_features -> full recomputation).  A version mismatch only
    warns."""
    cached_dist = cache["max_dist"]
    if max_dist != cached_dist:
        print("WARNING: Cached results were created using a differnt max-dist paramter")
        print(f"Current max-dist parameter: {max_dist}")
        print(f"Cached max-dist parameter: {cached_dist}")
        raise UnboundLocalError()
    cached_version = cache["version"]
    if cached_version != version:
        print(f"WARNING: Cached results were created using breakfast version {cached_version}")


def map_features(cached_feats, new_feats):
    """Outer join of cached and current profiles on the filtered profile *string*: one row per
    distinct string, sorted, with its row number in the cache (`idx_cache`) and in the current data
    (`idx_new`); NaN where absent."""
    cached_index = pd.Index(np.asarray(cached_feats, dtype=object))
    new_index = pd.Index(np.asarray(new_feats, dtype=object))
    every = cached_index.union(new_index)
    at_cache = cached_index.get_indexer(every).astype(float)
    at_new = new_index.get_indexer(every).astype(float)
    at_cache[at_cache < 0] = np.nan
    at_new[at_new < 0] = np.nan
    return pd.DataFrame({"idx_cache": at_cache, "idx_new": at_new}, index=every.rename("feature"))


def _cache_to_new(fmap):
    """int64 lookup cache row -> current row (-1 = profile no longer present)."""
    cached_rows = fmap["idx_cache"].to_numpy()
    known = ~np.isnan(cached_rows)
    size = int(cached_rows[known].max()) + 1 if known.any() else 0
    lookup = np.full(size, -1, dtype=np.int64)
    both = known & ~np.isnan(fmap["idx_new"].to_numpy())
    lookup[cached_rows[both].astype(np.int64)] = fmap["idx_new"].to_numpy()[both].astype(np.int64)
    return lookup


def update_neighbours_csr(neigh, fmap):
    """Cached neighbour lists re-indexed to the current rows, members of vanished profiles dropped,
    empty lists dropped — as (list_indptr int64, members int32) CSR."""
    lookup = _cache_to_new(fmap)
    if len(neigh) == 0:
        return np.zeros(1, np.int64), np.zeros(0, np.int32)
    lengths = np.fromiter((len(x) for x in neigh), dtype=np.int64, count=len(neigh))
    flat = np.concatenate([np.asarray(x, dtype=np.int64).ravel() for x in neigh]) if lengths.sum() else np.zeros(0, np.int64)
    owner = np.repeat(np.arange(len(neigh)), lengths)
    inside = (flat >= 0) & (flat < lookup.size)
    mapped = np.full(flat.size, -1, dtype=np.int64)
    mapped[inside] = lookup[flat[inside]]
    alive = mapped >= 0
    new_lengths = np.bincount(owner[alive], minlength=len(neigh))
    new_lengths = new_lengths[new_lengths > 0]
    list_indptr = np.concatenate(([0], np.cumsum(new_lengths))).astype(np.int64)
    return list_indptr, mapped[alive].astype(np.int32)


def update_neighbours(neigh, fmap):
    """Same result as the reference function of this name: a list of lists of current row numbers."""
    list_indptr, members = update_neighbours_csr(neigh, fmap)
    return [members[list_indptr[i]:list_indptr[i + 1]].tolist() for i in range(len(list_indptr) - 1)]


def find_deleted(fmap):
    """Cache rows whose profile is absent from the current data."""
    return fmap[fmap["idx_new"].isna().tolist()]["idx_cache"]


def find_new(fmap):
    """Current rows whose profile is absent from the cache (these are the only rows that need
    distance work)."""
    return fmap[fmap["idx_cache"].isna().tolist()]["idx_new"]
