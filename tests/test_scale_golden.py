"""Scale goldens (tests/golden/scale.json, made by tests/golden/make_scale_golden.py): the unmodified reference was
run on seeded synthetic tables of 41 643 and 66 807 sequences; only the recipe of each table and the SHA-256 of the
reference's clusters.tsv are committed.  Here the tables are regenerated and the product's host pipeline (native and
Python host paths) runs on them with the CPU hash-join oracle as the engine; tests/test_gpu_parity.py runs the same
cases through the CUDA kernels.  Byte-identical output == identical digest."""
import hashlib
import json

import click.testing
import pytest

from breakfast_b200 import console, synth
from oracle.engine_standin import HashJoinEngine
from tests import helpers

SCALE = json.loads((helpers.GOLDEN / "scale.json").read_text())


def build_table(recipe: dict) -> str:
    prof = synth.generate(recipe["n"], seed=recipe["seed"], with_mult=recipe["with_mult"])
    return prof.table(**recipe["table"]).to_csv(sep="\t", index=False)


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


def run_case(case, table_path, outdir):
    args = ["--input-file", str(table_path), "--outdir", str(outdir)] + helpers.cli_args(case["opts"])
    res = click.testing.CliRunner().invoke(console.main, args)
    assert res.exit_code == 0, f"{res.output}\n{res.exception!r}"
    return (outdir / "clusters.tsv").read_text()


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
