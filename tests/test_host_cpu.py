"""Host-side logic of the drop-in interface, on the CPU: feature filtering, vectorisation, dedup,
output, cache re-indexing, CLI option rules, and the max-dist-0 path end to end.  Mirrors what the
reference pins in tests/test_filtering.py and the non-distance parts of tests/test_breakfast.py."""
import re
from pathlib import Path

import click.testing
import numpy as np
import pandas as pd
import pytest

from breakfast_b200 import breakfast, cache, console, engine
from tests import helpers
from tests.helpers import GOLDEN

ROOT = Path(__file__).resolve().parent.parent


# ------------------------------------------------------------------ filter_features
@pytest.mark.parametrize("features,args,expected", [
    (["C241T"], (" ", "covsonar_dna", False, False, 0, 0, 1000), ["C241T"]),
    (["C241T"], (" ", "covsonar_dna", False, False, 250, 0, 1000), [""]),
    ([""], (" ", "covsonar_dna", False, False, 250, 0, 1000), [""]),
    (["  "], (" ", "covsonar_dna", False, False, 250, 0, 1000), [""]),
    (["C241T del:10:1 G5343TT"], (" ", "covsonar_dna", False, True, 0, 0, 1000), ["C241T G5343TT"]),
    (["C241T del:10:1 G5343TT"], (" ", "covsonar_dna", True, False, 0, 0, 1000), ["C241T del:10:1"]),
    (["C241T del:10:1 G5343TT"], (" ", "covsonar_dna", True, True, 0, 0, 1000), ["C241T"]),
    (["C241T del:10:1 G5343TT C900T"], (" ", "covsonar_dna", True, True, 250, 150, 1000), [""]),
    (["S:N501Y S:del:69:2 S:A222VV"], (" ", "covsonar_aa", True, True, 0, 0, 1000), ["S:N501Y"]),
    (["S:N501Y S:H69- S:Y144*"], (" ", "nextclade_aa", True, True, 0, 0, 1000), ["S:N501Y S:Y144*"]),
    (["C241T,1000-1010,1200,300:ACG"], (",", "nextclade_dna", True, True, 0, 0, 30000), ["C241T"]),
    (["C241T,1000-1010,1200,300:ACG"], (",", "nextclade_dna", False, True, 0, 0, 30000), ["C241T,300:ACG"]),
    (["a  b x"], (" ", "raw", True, True, 10, 10, 100), ["a b x"]),
])
def test_filter_features(features, args, expected):
    assert breakfast.filter_features(features, *args) == expected


def test_filter_trimming_is_inclusive_and_substitutions_only():
    f = ["C264T C265T C29674T C29675T del:100:5 A100AT"]
    out = breakfast.filter_features(f, " ", "covsonar_dna", False, False, 264, 228, 29903)
    assert out == ["C265T C29674T del:100:5 A100AT"]


def test_filter_noop_returns_the_same_object_and_keeps_invalid_tokens():
    f = pd.Series(["bogus  C241T"])
    assert breakfast.filter_features(f, " ", "covsonar_dna", False, False, 0, 0, 1000) is f


def test_filter_reports_invalid_tokens(capsys):
    out = breakfast.filter_features(["C241T bogus  G1A"], " ", "covsonar_dna", True, True, 0, 0, 1000)
    assert out == ["C241T G1A"]
    printed = capsys.readouterr().out
    assert "Skipping invalid feature: 'bogus'" in printed and "Skipping invalid feature: ''" in printed


def test_filter_unknown_type_exits():
    with pytest.raises(SystemExit):
        breakfast.filter_features(["x"], " ", "nope", True, True, 0, 0, 10)


# ------------------------------------------------------------------ vectorisation
def test_sparse_matrix_ignores_empty_features():
    m = breakfast.sparse_feature_matrix(["", "C241T"], " ")
    assert m.shape == (2, 1)
    assert m[0].count_nonzero() == 0 and m[1].count_nonzero() == 1


def test_sparse_matrix_vocabulary_order_and_repeats():
    m = breakfast.sparse_feature_matrix(pd.Series(["b a b", "a  c", float("nan")]), " ")
    assert m.shape == (3, 3) and m.dtype == np.int64
    assert m.indptr.tolist() == [0, 3, 5, 5]
    assert m.indices.tolist() == [0, 1, 0, 1, 2]          # b=0, a=1, c=2; the repeat is kept
    assert np.asarray(m.sum(axis=1)).ravel().tolist() == [3, 2, 0]


def test_sparse_matrix_all_empty_raises_like_the_reference():
    with pytest.raises(ValueError, match="unable to infer matrix dimensions"):
        breakfast.sparse_feature_matrix(["", ""], " ")


def test_thermometer_coding_equals_l1_on_counts():
    feats = ["A A B", "A B", "B A", "A A A B", "", "C"]
    indptr, indices, n_cols = engine.binary_csr(feats, " ")
    rows = [set(indices[indptr[i]:indptr[i + 1]].tolist()) for i in range(len(feats))]
    for i in range(len(feats)):
        assert sorted(rows[i]) == indices[indptr[i]:indptr[i + 1]].tolist()  # ascending, unique
    from collections import Counter
    counts = [Counter(t for t in f.split(" ") if t) for f in feats]
    for i in range(len(feats)):
        for j in range(len(feats)):
            l1 = sum(abs(counts[i][k] - counts[j][k]) for k in set(counts[i]) | set(counts[j]))
            assert len(rows[i] ^ rows[j]) == l1
    assert n_cols == 3 + 2  # A, B, C + (A,1), (A,2)


# ------------------------------------------------------------------ dedup / output
def test_collapse_duplicates_groups_by_string_in_first_appearance_order(capsys):
    meta = pd.DataFrame({"id": ["s1", "s2", "s3", "s4", "s5"], "feature": ["x y", "z", "x y", "y x", "z"]})
    nd = breakfast.collapse_duplicates(meta)
    assert nd.columns.tolist() == ["id", "feature"]
    assert nd["feature"].tolist() == ["x y", "z", "y x"]   # permutations stay separate profiles
    assert nd["id"].tolist() == [("s1", "s3"), ("s2", "s5"), ("s4",)]
    out = capsys.readouterr().out
    assert "Number of duplicates: 2" in out and "Number of unique sequences: 3" in out


def test_read_input_rejects_duplicate_ids():
    with pytest.raises(ValueError, match="Duplicate sequence identifiers"):
        breakfast.read_input(GOLDEN / "reference" / "duplicate-ids.tsv", "\t", "accession", "dna_profile")


def test_read_input_na_profile_is_the_empty_profile(tmp_path):
    p = tmp_path / "t.tsv"
    p.write_text("accession\tdna_profile\na\tNA\nb\t\nc\tC300T\n")
    t = breakfast.read_input(p, "\t", "accession", "dna_profile")
    assert t["feature"].tolist() == ["", "", "C300T"]


def test_write_output_renumbers_by_first_appearance(tmp_path):
    orig = pd.DataFrame({"id": ["c", "a", "d", "b", "e"], "feature": ["."] * 5})
    col = np.empty(3, dtype=object)
    col[:] = [7, pd.NA, 3]
    nodups = pd.DataFrame({"id": [("a", "b"), ("c",), ("d", "e")], "feature": ["x", "y", "z"], "cluster_id": col})
    breakfast.write_output(nodups, orig, tmp_path / "deep" / "out")
    assert (tmp_path / "deep" / "out" / "clusters.tsv").read_text() == "id\tcluster_id\nc\t\na\t1\nd\t2\nb\t1\ne\t2\n"


# ------------------------------------------------------------------ cache re-indexing
@pytest.mark.parametrize("ids", [
    ["s1", "s2", "s3", "s4", "s5", "s6"],                    # plain: the pyarrow writer
    ["a b", "é", "", "x,y", "007", "s'6"],                   # still plain for the csv module
    ["s\t1", "s2", "s3", "s4", "s5", "s6"],                  # a tab: quoted by the csv module -> pandas path
    ['s"1', "s2", "s3", "s4", "s5", "s6"],                   # a quote
    ["s\n1", "s2", "s3", "s4", "s5", "s6"],                  # a line break
    [11, 12, 13, 14, 15, 16],                                # numeric ids
], ids=["plain", "odd-but-plain", "tab", "quote", "newline", "numeric"])
@pytest.mark.parametrize("labels", [[3, 3, None, 7, 7, 3], [None] * 6, [5, 4, 3, 2, 1, 5]], ids=["mixed", "none", "all"])
def test_fast_clusters_writer_is_byte_identical_to_pandas(ids, labels, tmp_path, monkeypatch):
    """write_output's pyarrow writer against the pandas writer the reference uses (breakfast.py:64-69): same bytes for
    plain ids, and it steps aside (returns False) for anything the csv module would quote or that is not a string"""
    original = pd.DataFrame({"id": ids, "feature": ["f"] * 6})
    nodups = pd.DataFrame({"id": [(i,) for i in ids], "feature": ["f"] * 6,
                           "cluster_id": pd.Series([pd.NA if v is None else v for v in labels], dtype=object)})
    breakfast.write_output(nodups, original, tmp_path / "fast")
    monkeypatch.setattr(breakfast, "_write_clusters_arrow", lambda *a, **k: False)
    breakfast.write_output(nodups, original, tmp_path / "pandas")
    assert (tmp_path / "fast" / "clusters.tsv").read_bytes() == (tmp_path / "pandas" / "clusters.tsv").read_bytes()


def test_cache_map_and_update_with_ghost_list():
    cached = pd.Series(["A", "A g", "A g h", "Q"])
    new = pd.Series(["A g h", "N", "A"])                    # "A g" and "Q" vanished, "N" is new
    fmap = cache.map_features(cached, new)
    assert fmap.index.tolist() == sorted(["A", "A g", "A g h", "Q", "N"])
    assert np.array(cache.find_new(fmap)).astype(int).tolist() == [1]
    assert sorted(np.array(cache.find_deleted(fmap)).astype(int).tolist()) == [1, 3]
    neigh = [np.array([0, 1]), np.array([0, 1, 2]), [1, 2], np.array([3])]
    # list of row 1 ("A g", gone) keeps chaining A and "A g h": the ghost list
    assert cache.update_neighbours(neigh, fmap) == [[2], [2, 0], [0]]
    li, lm = cache.update_neighbours_csr(neigh, fmap)
    assert li.tolist() == [0, 1, 3, 4] and lm.tolist() == [2, 2, 0, 0]


def test_cache_roundtrip_and_validation(tmp_path, capsys):
    meta = pd.DataFrame({"id": [("a",), ("b", "c")], "feature": ["x", "y"], "extra": [1, 2]})
    f = tmp_path / "sub" / "cache.pkl.gz"
    cache.save(f, [np.array([0, 1]), [1]], meta, 2)
    got = cache.load(f, 2)
    assert got["max_dist"] == 2 and got["meta"].columns.tolist() == ["id", "feature"]
    assert [list(map(int, x)) for x in got["neigh"]] == [[0, 1], [1]]
    with pytest.raises(UnboundLocalError):
        cache.load(f, 1)
    with pytest.raises(TypeError):
        cache.load(None, 1)


def test_reference_written_cache_is_readable():
    got = cache.load(GOLDEN / "synthetic" / "ref_testfile_dist1.cache", 1)
    assert set(got) == {"max_dist", "version", "neigh", "meta"}
    assert got["meta"]["feature"].size == 3


# ------------------------------------------------------------------ CLI surface
@pytest.fixture
def runner():
    return click.testing.CliRunner()


def test_cli_help_lists_every_reference_option(runner):
    res = runner.invoke(console.main, ["--help"])
    assert res.exit_code == 0
    for opt in ("--input-file", "--sep", "--outdir", "--max-dist", "--min-cluster-size", "--input-cache",
                "--output-cache", "--id-col", "--clust-col", "--var-type", "--sep2", "--trim-start", "--trim-end",
                "--reference-length", "--skip-del", "--no-skip-del", "--skip-ins", "--no-skip-ins", "--jobs",
                "--version", "--gpus"):
        assert opt in res.output, opt


@pytest.mark.parametrize("extra", [["--clust-col", "somethingmissing"], ["--id-col", "somethingmissing"]])
def test_cli_missing_column_fails(runner, tmp_path, extra):
    res = runner.invoke(console.main, ["--input-file", str(GOLDEN / "reference" / "testfile.tsv"), "--outdir",
                                       str(tmp_path), "--max-dist", "0"] + extra)
    assert res.exit_code != 0


def test_cli_duplicate_ids_fail(runner, tmp_path):
    res = runner.invoke(console.main, ["--input-file", str(GOLDEN / "reference" / "duplicate-ids.tsv"), "--outdir",
                                       str(tmp_path), "--max-dist", "0"])
    assert res.exit_code != 0 and isinstance(res.exception, ValueError)


@pytest.mark.parametrize("extra", [["--trim-start", "10"], ["--trim-end", "10"], ["--skip-del"], ["--skip-ins"]])
def test_cli_non_dna_rejects_explicit_dna_options(runner, tmp_path, extra):
    res = runner.invoke(console.main, ["--input-file", str(GOLDEN / "reference" / "testfile.tsv"), "--outdir",
                                       str(tmp_path), "--max-dist", "0", "--var-type", "raw"] + extra)
    assert res.exit_code != 0 and "non-DNA" in res.output


def test_cli_trim_beyond_reference_fails(runner, tmp_path):
    res = runner.invoke(console.main, ["--input-file", str(GOLDEN / "reference" / "testfile.tsv"), "--outdir",
                                       str(tmp_path), "--max-dist", "0", "--trim-start", "40000"])
    assert res.exit_code != 0


@pytest.mark.parametrize("case", [c for c in helpers.cases("plain") if c["opts"].get("max_dist") == 0],
                         ids=lambda c: c["name"])
def test_cli_max_dist_zero_goldens(case, tmp_path):
    """--max-dist 0 is a host-only path in the reference too (breakfast.py:343-364): no GPU needed."""
    helpers.assert_matches(case, case["expected"], helpers.run_cli(case["input"], case["opts"], tmp_path))


# ------------------------------------------------------------------ native host path == Python host path
def _python_path(meta, sep2, opts):
    m = meta.copy()
    m["feature"] = breakfast.filter_features(m["feature"], sep2, *opts)
    nd = breakfast.collapse_duplicates(m)
    indptr, indices, vocab = engine.tokenise(nd["feature"], sep2)
    return nd, indptr, indices, len(vocab)


@pytest.mark.parametrize("opts", [
    ("covsonar_dna", True, True, 264, 228, 29903), ("covsonar_dna", False, False, 0, 0, 29903),
    ("covsonar_dna", False, True, 3000, 3000, 29903), ("raw", True, True, 264, 228, 29903),
    ("raw", False, False, 0, 0, 29903)], ids=lambda o: "-".join(map(str, o)))
@pytest.mark.parametrize("table", ["synthetic/quirks.tsv", "synthetic/syn_dna.tsv.gz", "reference/testfile.tsv"])
def test_native_host_path_equals_python_path(table, opts, capsys):
    from breakfast_b200 import hostfast
    assert hostfast.available()
    meta = breakfast.read_input(GOLDEN / table, "\t", "accession", "dna_profile")
    capsys.readouterr()
    nd_py, indptr, indices, n_vocab = _python_path(meta, " ", opts)
    out_py = capsys.readouterr().out
    nd_c, pre = hostfast.prepare(meta, " ", *opts)
    out_c = capsys.readouterr().out
    assert out_c == out_py                                        # same messages in the same order
    assert nd_c["id"].tolist() == nd_py["id"].tolist() and nd_c["feature"].tolist() == nd_py["feature"].tolist()
    assert nd_c.attrs == {}                                       # nothing private rides along in the frame
    assert pre["n_vocab"] == n_vocab
    # same token matrix up to the numbering of the vocabulary (the native pass numbers the kept tokens in interning
    # order, the reference by first appearance): the two id sequences are related by a bijection
    assert np.array_equal(pre["token_indptr"], indptr) and pre["token_indices"].size == indices.size
    fwd = dict(zip(pre["token_indices"].tolist(), indices.tolist()))
    assert len(fwd) == len(set(fwd.values())) == n_vocab
    assert np.array_equal(np.array([fwd[x] for x in pre["token_indices"].tolist()], dtype=indices.dtype), indices)
    bi, bx, nc = engine.thermometer_binarise(indptr, indices, n_vocab)
    assert np.array_equal(pre["bin_indptr"], bi) and pre["n_cols"] == nc
    # extra (repeat) columns may be numbered differently: compare the pairwise set distances instead
    rows_c = [set(pre["bin_indices"][pre["bin_indptr"][i]:pre["bin_indptr"][i + 1]].tolist()) for i in range(min(len(nd_c), 60))]
    rows_p = [set(bx[bi[i]:bi[i + 1]].tolist()) for i in range(min(len(nd_c), 60))]
    for i in range(len(rows_c)):
        assert sorted(rows_c[i]) == pre["bin_indices"][pre["bin_indptr"][i]:pre["bin_indptr"][i + 1]].tolist()
        for j in range(len(rows_c)):
            assert len(rows_c[i] ^ rows_c[j]) == len(rows_p[i] ^ rows_p[j])


def test_native_host_path_nextclade_and_multichar_separator(capsys):
    from breakfast_b200 import hostfast
    meta = pd.DataFrame({"id": list("abcde"), "feature": ["C300T; G400A", "G400A; C300T; 12-15", "C300T; G400A; ; 5:ACG",
                                                           "C300T; G400A", "é1; C300T"]})
    opts = ("nextclade_dna", True, True, 10, 10, 29903)
    nd_py, indptr, indices, n_vocab = _python_path(meta, "; ", opts)
    nd_c, pre = hostfast.prepare(meta, "; ", *opts)
    assert nd_c["id"].tolist() == nd_py["id"].tolist() and nd_c["feature"].tolist() == nd_py["feature"].tolist()
    fwd = dict(zip(pre["token_indices"].tolist(), indices.tolist()))
    assert len(fwd) == len(set(fwd.values()))
    assert [fwd[x] for x in pre["token_indices"].tolist()] == indices.tolist()


@pytest.mark.parametrize("kind", ["large_string", "string"])
def test_native_host_path_reads_a_chunked_arrow_column_in_place(kind, capsys):
    """the profile column as the Arrow CSV reader leaves it: many chunks, one of them a slice with a non-zero offset, one
    empty - tokenised in place (bfh_tokenise_arrow_chunks), same result as the Python path on the same strings"""
    import pyarrow as pa
    from breakfast_b200 import hostfast
    meta = breakfast.read_input(GOLDEN / "synthetic/syn_dna.tsv.gz", "\t", "accession", "dna_profile")
    profiles = meta["feature"].tolist()
    typ = pa.large_string() if kind == "large_string" else pa.string()
    cuts = [0, 1, 2, 50, 50, 333, len(profiles)]
    chunks = [pa.array(profiles[a:b], type=typ) for a, b in zip(cuts[:-1], cuts[1:])]
    padded = pa.array(["zzz"] * 7 + profiles[2:50] + ["yyy"], type=typ)
    chunks[2] = padded.slice(7, 48)                       # a chunk whose offsets do not start at its buffer
    column = pa.chunked_array(chunks, type=typ)
    assert column.num_chunks == 6 and column.to_pylist() == profiles
    chunked = pd.DataFrame({"id": meta["id"], "feature": pd.Series(pd.arrays.ArrowStringArray(column), dtype=meta["feature"].dtype)})
    opts = ("covsonar_dna", True, True, 264, 228, 29903)
    capsys.readouterr()
    nd_py, indptr, indices, n_vocab = _python_path(meta, " ", opts)
    out_py = capsys.readouterr().out
    nd_c, pre = hostfast.prepare(chunked, " ", *opts)
    out_c = capsys.readouterr().out
    assert out_c == out_py
    assert nd_c["id"].tolist() == nd_py["id"].tolist() and nd_c["feature"].tolist() == nd_py["feature"].tolist()
    assert pre["n_vocab"] == n_vocab and np.array_equal(pre["token_indptr"], indptr)
    # and with no filter active (profiles compared and returned as raw strings)
    raw_opts = ("covsonar_dna", False, False, 0, 0, 29903)
    nd_py, indptr, _, n_vocab = _python_path(meta, " ", raw_opts)
    nd_c, pre = hostfast.prepare(chunked, " ", *raw_opts)
    assert nd_c["id"].tolist() == nd_py["id"].tolist() and nd_c["feature"].tolist() == nd_py["feature"].tolist()
    assert pre["n_vocab"] == n_vocab and np.array_equal(pre["token_indptr"], indptr)


@pytest.mark.parametrize("var_type", ["covsonar_dna", "nextclade_dna"])
@pytest.mark.parametrize("flags", [(True, True, 264, 228), (False, True, 0, 0), (True, False, 5, 29000), (False, False, 1, 0)],
                         ids=lambda f: "-".join(map(str, f)))
def test_native_token_verdicts_equal_the_reference_expressions(var_type, flags, capsys):
    """bfh_classify_dna (byte scanners) against the reference's regular expressions (breakfast.py:135-184) on tokens
    built to sit on every edge of the patterns - short forms, missing parts, lower case, huge positions, trim
    boundaries, non-ASCII digits and letters, line breaks - plus random strings over the patterns' alphabet"""
    from breakfast_b200 import hostfast
    skip_ins, skip_del, trim_start, trim_end = flags
    reflen = 29903
    rng = np.random.default_rng(5)
    edge = ["A1T", "A12T", "AT", "A", "", "1", "A1", "1A", "a1T", "A1t", "A1TT", "AA1T", "A 1T", "A1.T", "A-1T", "A+1T",
            "A0T", "A00012T", "A264T", "A265T", "A29674T", "A29675T", "A29676T", "A5T", "A6T", "A28999T", "A29000T",
            "A99999999999999999999999999T", "A9223372036854775807T", "A9223372036854775808T",
            "AB", "xAB", "12AB", "A1AB", "ab", "aB", "Ab", "ABc", ":AB", "del:1:1", "del:1:", "del::1", "del:1", "del:1:1:1",
            "del:12:34", "Del:1:1", "del:1:1A", "del:1:1AB", "del:a:1", "del:1:b", "del: 1:1", "12:A", "12:ACGT", "12:", ":A",
            "12:a", "12:A1", "1:AB", "12", "12-13", "12-", "-12", "12--13", "12-13-14", "0", "007", "1-2A", "12:AC GT",
            "A1T\n", "AB\n", "del:1:1\n", "12\n", "\nA1T", "A１２T", "A1Ｔ", "ÉB", "AÉ", "A٣T", "12٣", "del:١:1", "é1"]
    alphabet = list("ACGTNdel:-0123456789 aX.") + ["É", "٣"]
    fuzz = ["".join(rng.choice(alphabet, size=int(rng.integers(0, 9)))) for _ in range(4000)]
    tokens = sorted(set(edge + fuzz) - {""}) + [""]
    tokens = [t for t in tokens if "|" not in t]
    # one profile per token (plus one with all of them), separator "|"
    meta = pd.DataFrame({"id": [f"s{i}" for i in range(len(tokens) + 1)], "feature": tokens + ["|".join(tokens)]})
    opts = (var_type, skip_ins, skip_del, trim_start, trim_end, reflen)
    capsys.readouterr()
    nd_py, indptr, indices, n_vocab = _python_path(meta, "|", opts)
    out_py = capsys.readouterr().out
    nd_c, pre = hostfast.prepare(meta, "|", *opts)
    out_c = capsys.readouterr().out
    assert out_c == out_py
    assert nd_c["id"].tolist() == nd_py["id"].tolist() and nd_c["feature"].tolist() == nd_py["feature"].tolist()
    assert pre["n_vocab"] == n_vocab and np.array_equal(pre["token_indptr"], indptr)


# ------------------------------------------------------------------ no CPU fallback, no oracle in the product
def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|ref_port|liboracle", re.M)
    for f in list((ROOT / "breakfast_b200").rglob("*.py")) + list((ROOT / "breakfast").rglob("*.py")) + \
            list((ROOT / "breakfast_b200" / "csrc").glob("*")):
        assert not pat.search(f.read_text(errors="ignore")), f"{f} refers to the oracle"


def test_cluster_without_gpu_fails_loudly(tmp_path):
    from breakfast_b200 import _native
    if _native.device_count() > 0:
        pytest.skip("a GPU is present")
    meta = pd.DataFrame({"id": [("a",), ("b",)], "feature": ["C300T", "C300T G400A"]})
    with pytest.raises(_native.NativeError):
        breakfast.cluster(meta, " ", 1, 1, None, None)
