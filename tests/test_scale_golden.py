"""Scale goldens (tests/golden/scale.json, made by tests/golden/make_scale_golden.py): the unmodified reference was
run on seeded synthetic tables of 41 643 and 66 807 sequences; only the recipe of each table and the SHA-256 of the
reference's clusters.tsv are committed.  Here the tables are regenerated and the product's host pipeline (native and
Python host paths) runs on them with the CPU hash-join oracle as the engine; tests/test_gpu_parity.py runs the same
cases through the CUDA kernels.  Byte-identical output == identical digest."""
import hashlib
import json

import click.testing
import pytest

from breakfast_b200 import console
from oracle.engine_standin import HashJoinEngine, OracleEngine
from tests import helpers

SCALE = json.loads((helpers.GOLDEN / "scale.json").read_text())


build_table = helpers.scale_table


@pytest.fixture(scope="module")
def tables(tmp_path_factory):
    root = tmp_path_factory.mktemp("scale")
    out = {}
    for name, recipe in SCALE["tables"].items():
        text = build_table(recipe)
        assert hashlib.sha256(text.encode()).hexdigest() == recipe["sha256"], f"{name}: the generator drifted"
        out[name] = root / f"{name}.tsv"
        out[name].write_text(text)
    return out


def run_case(case, table_path, outdir, cache_in=None, cache_out=None):
    args = ["--input-file", str(table_path), "--outdir", str(outdir)] + helpers.cli_args(case["opts"])
    if cache_in:
        args += ["--input-cache", str(cache_in)]
    if cache_out:
        args += ["--output-cache", str(cache_out)]
    res = click.testing.CliRunner().invoke(console.main, args)
    assert res.exit_code == 0, f"{res.output}\n{res.exception!r}"
    return (outdir / "clusters.tsv").read_text()


def write_chain(root, name):
    """regenerate the three tables of a cached chain, check them against the committed digests"""
    recipe = SCALE["chains"][name]
    paths = []
    for k, text in enumerate(helpers.scale_chain(recipe)):
        assert hashlib.sha256(text.encode()).hexdigest() == recipe["sha256"][k], f"{name} step {k}: the generator drifted"
        paths.append(root / f"{name}_{k}.tsv")
        paths[-1].write_text(text)
    return paths


def run_chain(case, steps, root):
    """step 0 fresh (writes a cache), step 1 and 2 from the previous cache; plus step 1 without a cache"""
    c0, c1 = root / "c0", root / "c1"
    outs = [run_case(case, steps[0], root / "o0", None, c0), run_case(case, steps[1], root / "o1", c0, c1),
            run_case(case, steps[2], root / "o2", c1, None)]
    return outs, run_case(case, steps[1], root / "o1_fresh")


def check_chain(case, outs, fresh):
    digests = [hashlib.sha256(o.encode()).hexdigest() for o in outs]
    assert [o.count("\n") for o in outs] == case["n_lines"]
    assert digests == case["sha256"], case["name"]
    assert hashlib.sha256(fresh.encode()).hexdigest() == case["fresh_step1_sha256"]
    assert case["fresh_step1_sha256"] != case["sha256"][1]      # the ghost lists of the cache change the answer


@pytest.mark.parametrize("case", SCALE["cases"], ids=[c["name"] for c in SCALE["cases"]])
def test_host_pipeline_reproduces_the_reference_digest(case, tables, tmp_path, monkeypatch):
    HashJoinEngine.install(monkeypatch)
    got = run_case(case, tables[case["table"]], tmp_path)
    assert got.count("\n") == case["n_lines"]
    assert hashlib.sha256(got.encode()).hexdigest() == case["sha256"], case["name"]


def test_python_host_path_gives_the_same_digest(tables, tmp_path, monkeypatch):
    """the pure-Python mirror of the host functions (BREAKFAST_B200_HOST=python), on the smaller table"""
    case = next(c for c in SCALE["cases"] if c["name"] == "scale_nextclade_d1")
    monkeypatch.setenv("BREAKFAST_B200_HOST", "python")
    HashJoinEngine.install(monkeypatch)
    got = run_case(case, tables[case["table"]], tmp_path)
    assert hashlib.sha256(got.encode()).hexdigest() == case["sha256"]


@pytest.mark.parametrize("case", SCALE["chain_cases"], ids=[c["name"] for c in SCALE["chain_cases"]])
def test_cached_chain_reproduces_the_reference_digests(case, tmp_path, monkeypatch):
    """19 k -> 20 k -> 19 k sequences through --output-cache / --input-cache: deleted profiles leave ghost lists, new
    and modified sequences take the incremental (new x all) path; every step equals what the reference wrote"""
    OracleEngine.install(monkeypatch)       # incremental path: exact new x all edges from oracle.c
    HashJoinEngine.install(monkeypatch)     # full path: the hash-join oracle
    steps = write_chain(tmp_path, case["chain"])
    outs, fresh = run_chain(case, steps, tmp_path)
    check_chain(case, outs, fresh)
