"""Parity at the shapes BASELINE.json lists as configs 2-5 (config 1 = the reference's own goldens,
tests/test_gpu_parity.py).  Full-size where the oracle can keep up, otherwise size-independent
properties plus exact checks on row samples.  GPU only."""
import numpy as np
import pandas as pd
import pytest

import oracle
from oracle import ref_port
from breakfast_b200 import _native, breakfast, cache, engine, synth
from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    _native.require_device()


def _write_table(df, path):
    df.to_csv(path, sep="\t", index=False)
    return path


def _run_cli_file(path, opts, outdir):
    import click.testing
    from breakfast_b200 import console
    res = click.testing.CliRunner().invoke(console.main, ["--input-file", str(path), "--outdir", str(outdir)] + helpers.cli_args(opts))
    assert res.exit_code == 0, res.output[-2000:]
    return (outdir / "clusters.tsv").read_text()


def test_config2_100k_covsonar_cli_bytes_equal_oracle(tmp_path):
    """config 2: synthetic 100k covsonar_dna profiles, --max-dist 1 --skip-del, 1 GPU, through the CLI;
    clusters.tsv must be byte-identical to the oracle pipeline (oracle.c for the distance part)."""
    prof = synth.generate(100_000, seed=2, with_mult=False)
    path = _write_table(prof.table("covsonar_dna", " "), tmp_path / "c2.tsv")
    opts = dict(max_dist=1, skip_del=True)
    got = _run_cli_file(path, opts, tmp_path / "out")
    want, _ = ref_port.run_file(path, core="c", **opts)
    assert got == want
    t = pd.read_table(tmp_path / "out" / "clusters.tsv")
    assert len(t) == 100_000 and t["cluster_id"].notna().sum() > 50_000


def test_config4_nextclade_indels_trimming_cli_and_4_ranks(tmp_path):
    """config 4 shape: nextclade_dna with indels in the clustered column, --skip-ins --skip-del and
    trimmed ends (wide trims so that substitutions really are dropped and profiles collapse), through
    the CLI; then the same rows split over 4 ranks at the C ABI."""
    prof = synth.generate(60_000, seed=4, with_mult=True, unique_on_all_events=True)
    path = _write_table(prof.table("nextclade_dna", ",", id_col="seqName", feature_col="substitutions"), tmp_path / "c4.tsv")
    opts = dict(max_dist=1, sep2=",", id_col="seqName", clust_col="substitutions", var_type="nextclade_dna",
                skip_ins=True, skip_del=True, trim_start=3000, trim_end=3000)
    got = _run_cli_file(path, opts, tmp_path / "out")
    want, _ = ref_port.run_file(path, core="c", **opts)
    assert got == want
    # 4 ranks on the filtered, deduplicated rows
    ids, feats = ref_port.read(path, "\t", "seqName", "substitutions")
    feats = ref_port.filter_profiles(feats, ",", "nextclade_dna", True, True, 3000, 3000, 29903)
    uniq, codes, mult = ref_port.dedup(feats)
    assert len(uniq) < len(set(prof.features("nextclade_dna", ",")))      # trimming merged profiles
    indptr, indices, n_cols = engine.binary_csr(uniq, ",")
    single, st1 = _native.cluster_csr(indptr, indices, n_cols, 1)
    gathered = np.empty((4, len(uniq)), dtype=np.int32)
    with _native.Context() as ctx:
        ctx.upload_csr(indptr, indices, n_cols)
        for r in range(4):
            ctx.run_sync(1, rank=r, world=4)
            gathered[r] = ctx.download_labels()
        ctx.merge_labels_host(gathered)
        assert np.array_equal(ctx.download_labels(), single)
    want_labels, _ = oracle.cluster(indptr, indices, 1)
    assert np.array_equal(single, want_labels)


@pytest.fixture(scope="module")
def million_mult():
    prof = synth.generate(1_000_000, seed=3, with_mult=True)
    return prof.csr() + (prof.mult.copy(),)


def test_config3_1m_dist2_mincluster5_8_ranks(million_mult):
    """config 3: 1M unique profiles, --max-dist 2 --min-cluster-size 5, tile space over 8 ranks with
    label merge.  Ranks are emulated one after the other on this GPU; checks: merged == single-rank,
    exact neighbour sets on a row sample, labels == CPU union-find of the reported edges, and the
    weighted size filter of the host on top."""
    indptr, indices, n_cols, mult = million_mult
    n = len(indptr) - 1
    with _native.Context(want_edges=1) as ctx:
        ctx.upload_csr(indptr, indices, n_cols)
        st = ctx.run_sync(2)
        single = ctx.download_labels()
        src, dst = ctx.download_edges()
    assert st.pairs_band > 2.5e10 and src.size == st.n_edges
    assert np.array_equal(oracle.components(n, src, dst), single)
    # the whole edge set, bit for bit, against the second CPU oracle (hash joins + exact verification, about a minute)
    from oracle import hashjoin
    hs, hd = hashjoin.edges(indptr, indices, 2)
    o = np.lexsort((np.maximum(src, dst), np.minimum(src, dst)))
    assert np.array_equal(np.minimum(src, dst)[o], hs) and np.array_equal(np.maximum(src, dst)[o], hd)
    assert np.array_equal(single, oracle.components(n, hs, hd))
    rng = np.random.default_rng(3)
    q = np.sort(rng.choice(n, size=800, replace=False)).astype(np.int32)
    ws, wd = oracle.edges(indptr, indices, 2, queries=q)
    mask = np.isin(src, q) | np.isin(dst, q)
    assert set(zip(src[mask].tolist(), dst[mask].tolist())) == set(zip(ws.tolist(), wd.tolist()))
    gathered = np.empty((8, n), dtype=np.int32)
    tiles = 0
    with _native.Context() as ctx:
        ctx.upload_csr(indptr, indices, n_cols)
        for r in range(8):
            s = ctx.run_sync(2, rank=r, world=8)
            tiles += s.tiles_rank
            gathered[r] = ctx.download_labels()
        assert tiles == st.tiles_band
        ctx.merge_labels_host(gathered)
        assert np.array_equal(ctx.download_labels(), single)
    # host: clusters need >= 5 sequences (multiplicities count), ids by smallest member
    sizes = np.bincount(single, weights=mult, minlength=n)
    keep = sizes[single] >= 5
    frame = pd.DataFrame({"id": pd.Series([(0,) * int(m) for m in mult], dtype=object)})
    k = breakfast._assign_cluster_ids(frame, single.astype(np.int64), 5)
    col = frame["cluster_id"].to_numpy(dtype=object)
    assert k == int((sizes[np.flatnonzero(single == np.arange(n))] >= 5).sum())
    assert np.array_equal(~pd.isna(col), keep)
    ids_kept = np.array([c for c in col[keep]], dtype=np.int64)
    assert len(set(zip(single[keep].tolist(), ids_kept.tolist()))) == k      # one id per kept component


def test_config5_incremental_1m_cached_plus_delta(million_mult):
    """config 5 shape: a cache from a 1M-profile run, then a delta (new profiles, deleted profiles incl.
    ones that leave ghost lists); only the new x all block is evaluated.  Expected labels = CPU
    union-find over (re-indexed cached lists chained) + (exact new x all edges from the oracle)."""
    indptr, indices, n_cols, _ = million_mult
    n_old = 1_000_000 - 6_000                       # the last 6000 rows play the "new" profiles
    old_ip, old_ix = indptr[: n_old + 1], indices[: indptr[n_old]]
    # --- run 1: full run on the old set, neighbour lists as the cache would store them
    with _native.Context(want_edges=1) as ctx:
        ctx.upload_csr(old_ip, old_ix, n_cols)
        ctx.run_sync(1)
        src, dst = ctx.download_edges()
    neigh = engine.adjacency_lists(n_old, np.arange(n_old), src, dst)
    # --- delta: 10k old profiles vanish (among them bridge profiles), 6k new ones arrive, order changes
    rng = np.random.default_rng(55)
    deg = np.bincount(np.concatenate([src, dst]), minlength=n_old)
    bridges = np.flatnonzero(deg >= 2)
    gone = np.union1d(rng.choice(bridges, size=5_000, replace=False), rng.choice(n_old, size=5_000, replace=False))
    alive_old = np.setdiff1d(np.arange(n_old), gone)
    new_order = rng.permutation(np.concatenate([alive_old, np.arange(n_old, 1_000_000)]))
    names = np.array([f"p{i}" for i in range(1_000_000)], dtype=object)     # stand-in profile strings
    fmap = cache.map_features(pd.Series(names[:n_old]), pd.Series(names[new_order]))
    list_indptr, list_members = cache.update_neighbours_csr(neigh, fmap)
    new_rows = np.sort(np.array(cache.find_new(fmap)).astype(np.int32))
    assert new_rows.size == 6_000
    # CSR of the current rows
    lens = np.diff(indptr)[new_order]
    cur_ip = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    starts = indptr[:-1][new_order]
    gather = np.repeat(starts - cur_ip[:-1], lens) + np.arange(int(cur_ip[-1]))
    cur_ix = indices[gather]
    res = engine.components_incremental(cur_ip, cur_ix, n_cols, 1, new_rows, list_indptr, list_members, want_edges=True)
    n_cur = len(new_order)
    # --- expected
    ws, wd = oracle.edges(cur_ip, cur_ix, 1, queries=new_rows)
    assert np.array_equal(res.edges[0], ws) and np.array_equal(res.edges[1], wd)
    want = oracle.components(n_cur, ws, wd, list_indptr, list_members)
    assert np.array_equal(res.labels, want)
    # ghost lists matter here: dropping them changes the partition
    assert not np.array_equal(oracle.components(n_cur, ws, wd), want) or list_members.size == 0
    assert res.stats["n_query"] == 6_000 and res.stats["pairs_total"] < 6_000 * n_cur


# ------------------------------------------------------------------------------------------------------------------
# configs 4 and 5 at BASELINE's stated size, through the CLI; expected digests: tests/golden/configs.json
# (tests/golden/make_config_golden.py: the reference's host steps restated + hash-join oracle, run on the CPU box)
# ------------------------------------------------------------------------------------------------------------------
import hashlib
import json

CONFIGS = json.loads((helpers.GOLDEN / "configs.json").read_text())


def _sha(text):
    return hashlib.sha256(text.encode()).hexdigest()


def _cli(path, opts, outdir, cache_in=None, cache_out=None):
    import click.testing
    from breakfast_b200 import console
    args = ["--input-file", str(path), "--outdir", str(outdir)] + helpers.cli_args(opts)
    if cache_in:
        args += ["--input-cache", str(cache_in)]
    if cache_out:
        args += ["--output-cache", str(cache_out)]
    res = click.testing.CliRunner().invoke(console.main, args)
    assert res.exit_code == 0, f"{res.output[-2000:]}\n{res.exception!r}"
    return (outdir / "clusters.tsv").read_text()


def test_config4_full_size_500k_nextclade_4_devices(tmp_path, monkeypatch):
    """config 4 as BASELINE.json states it: 500 000 unique nextclade_dna profiles (833 990 sequences) with indels and
    trimming, --skip-ins --skip-del, the pairwise work split over 4 ranks (four contexts on this GPU when the box has
    fewer than four: BREAKFAST_B200_DEVICES=0,0,0,0), through console.main; clusters.tsv byte-identical to the CPU
    oracle pipeline's (digest)."""
    want = CONFIGS["config4"]
    text = helpers.config4_table(want["n_profiles"])
    assert _sha(text) == want["table_sha256"], "the generator drifted"
    path = tmp_path / "c4.tsv"
    path.write_text(text)
    del text
    n_dev = _native.device_count()
    monkeypatch.setenv("BREAKFAST_B200_DEVICES", "0,1,2,3" if n_dev >= 4 else "0,0,0,0")
    got = _cli(path, want["opts"], tmp_path / "out")
    assert got.count("\n") == want["n_lines"]
    assert _sha(got) == want["sha256"]


def test_config5_full_size_1m_cache_plus_50k_delta_through_the_cli(tmp_path):
    """config 5 as BASELINE.json states it: a first run over 1 000 000 unique profiles (1 667 179 sequences, --max-dist 2
    --min-cluster-size 5) writes the reference-format cache; the second run gets a table with 50 000 sequences added,
    modified or deleted (whole profiles vanish: ghost lists) and --input-cache, so only the new x all block is
    evaluated.  Both outputs byte-identical to the CPU oracle pipeline's (digests); the cached answer differs from a
    fresh run of the same table, as it does for the reference (ghost lists)."""
    want = CONFIGS["config5"]
    text0, text1 = helpers.config5_tables(want["n_profiles"], want["n_delta"])
    assert [_sha(text0), _sha(text1)] == want["table_sha256"], "the generator drifted"
    p0, p1 = tmp_path / "c5_0.tsv", tmp_path / "c5_1.tsv"
    p0.write_text(text0)
    p1.write_text(text1)
    del text0, text1
    cache = tmp_path / "cache" / "c5.cache"
    got0 = _cli(p0, want["opts"], tmp_path / "o0", cache_out=cache)
    assert got0.count("\n") == want["n_lines"][0] and _sha(got0) == want["sha256"][0]
    del got0
    got1 = _cli(p1, want["opts"], tmp_path / "o1", cache_in=cache)
    assert got1.count("\n") == want["n_lines"][1] and _sha(got1) == want["sha256"][1]
    assert want["fresh_step1_sha256"] != want["sha256"][1]
