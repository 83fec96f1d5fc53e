"""The whole host pipeline (CLI -> parse -> filter -> dedup -> CSR -> [engine] -> cache -> labels ->
clusters.tsv) against every golden, on the CPU, with the CUDA engine replaced by the oracle through a
test-only monkeypatch.  This pins the host logic (incl. cache / ghost-list handling and cache-file
interchange with the reference) independently of the GPU; tests/test_gpu_parity.py runs the same
cases through the real kernels."""
import shutil

import pytest

from tests import helpers
from tests.helpers import GOLDEN, OracleEngine


@pytest.fixture(autouse=True, params=["native", "python"])
def oracle_engine(monkeypatch, request):
    """both host paths: the native tokenise/filter/dedup pass (csrc/host_parse.cpp) and the Python functions"""
    monkeypatch.setenv("BREAKFAST_B200_HOST", request.param)
    OracleEngine.install(monkeypatch)


@pytest.mark.parametrize("case", helpers.cases("plain"), ids=lambda c: c["name"])
def test_plain(case, tmp_path):
    helpers.assert_matches(case, case["expected"], helpers.run_cli(case["input"], case["opts"], tmp_path))


@pytest.mark.parametrize("case", [c for c in helpers.cases("cached") if "cache_from" in c], ids=lambda c: c["name"])
def test_cached_own_cache(case, tmp_path):
    first = case["cache_from"]
    cache = tmp_path / "cachedir" / "cache.pkl"
    helpers.run_cli(first["input"], first["opts"], tmp_path / "first", cache_out=cache)
    assert cache.exists()
    text = helpers.run_cli(case["input"], case["opts"], tmp_path / "second", cache_in=cache)
    helpers.assert_matches(case, case["expected"], text)


@pytest.mark.parametrize("case", [c for c in helpers.cases("cached") if "reference_cache_file" in c],
                         ids=lambda c: c["name"])
def test_cached_reference_written_cache(case, tmp_path):
    """a cache file written by the reference itself is a valid --input-cache for the drop-in"""
    cache = tmp_path / "ref.cache"
    shutil.copyfile(GOLDEN / case["reference_cache_file"], cache)
    text = helpers.run_cli(case["input"], case["opts"], tmp_path / "out", cache_in=cache)
    helpers.assert_matches(case, case["expected"], text)


@pytest.mark.parametrize("case", helpers.cases("chain"), ids=lambda c: c["name"])
def test_cache_chain(case, tmp_path):
    prev = None
    for k, step in enumerate(case["chain"]):
        out_cache = tmp_path / f"cache{k}"
        text = helpers.run_cli(step["input"], case["opts"], tmp_path / f"out{k}", cache_in=prev, cache_out=out_cache)
        helpers.assert_matches(case, step["expected"], text)
        prev = out_cache


def test_cache_max_dist_mismatch_recomputes(tmp_path):
    cache = tmp_path / "c"
    helpers.run_cli("synthetic/cache_step0.tsv.gz", dict(max_dist=1), tmp_path / "a", cache_out=cache)
    text = helpers.run_cli("synthetic/cache_step1.tsv.gz", dict(max_dist=2), tmp_path / "b", cache_in=cache)
    assert text == (GOLDEN / "synthetic" / "cache_d2_step1_fresh.expected.tsv").read_text()


def test_cache_file_carries_no_private_metadata(tmp_path, monkeypatch):
    """the --output-cache pickle holds {max_dist, version, neigh, meta[id, feature]} and nothing else: the native host
    path must not leak its CSR arrays into it (the frame's attrs are empty, the file is as small as the Python path's)"""
    import _pickle
    import gzip
    sizes = {}
    for host in ("native", "python"):
        monkeypatch.setenv("BREAKFAST_B200_HOST", host)
        cache = tmp_path / f"cache_{host}"
        helpers.run_cli("synthetic/cache_step0.tsv.gz", dict(max_dist=1), tmp_path / host, cache_out=cache)
        with gzip.open(cache, "rb") as fh:
            payload = _pickle.load(fh)
        assert sorted(payload) == ["max_dist", "meta", "neigh", "version"]
        assert payload["meta"].attrs == {} and list(payload["meta"].columns) == ["id", "feature"]
        sizes[host] = cache.stat().st_size
    assert abs(sizes["native"] - sizes["python"]) <= 64, sizes
