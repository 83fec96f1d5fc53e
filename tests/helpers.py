"""Shared helpers for the parity tests: golden manifest, CLI runner, oracle stand-ins."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pandas as pd

GOLDEN = Path(__file__).resolve().parent / "golden"

_FLAGS = {"id_col": "--id-col", "clust_col": "--clust-col", "var_type": "--var-type", "sep2": "--sep2",
          "max_dist": "--max-dist", "min_cluster_size": "--min-cluster-size", "trim_start": "--trim-start",
          "trim_end": "--trim-end", "reference_length": "--reference-length", "sep": "--sep"}


def load_manifest() -> dict:
    return json.loads((GOLDEN / "manifest.json").read_text())


def cases(kind=None):
    out = []
    for c in load_manifest()["cases"]:
        is_chain = "chain" in c
        needs_cache = is_chain or "cache_from" in c or "reference_cache_file" in c
        k = "chain" if is_chain else ("cached" if needs_cache else "plain")
        if kind is None or k == kind:
            out.append(c)
    return out


def cli_args(opts: dict) -> list:
    a = []
    for k, v in opts.items():
        if k in _FLAGS:
            a += [_FLAGS[k], str(v)]
        elif k == "skip_del":
            a.append("--skip-del" if v else "--no-skip-del")
        elif k == "skip_ins":
            a.append("--skip-ins" if v else "--no-skip-ins")
        else:
            raise KeyError(k)
    return a


def run_cli(input_rel: str, opts: dict, outdir: Path, cache_in=None, cache_out=None) -> str:
    """Run the product CLI (breakfast_b200.console.main) like the reference's tests do."""
    import click.testing
    from breakfast_b200 import console
    args = ["--input-file", str(GOLDEN / input_rel), "--outdir", str(outdir)] + cli_args(opts)
    if cache_in:
        args += ["--input-cache", str(cache_in)]
    if cache_out:
        args += ["--output-cache", str(cache_out)]
    res = click.testing.CliRunner().invoke(console.main, args)
    if res.exit_code != 0:
        raise AssertionError(f"CLI failed ({res.exit_code}) for {args}:\n{res.output}\n{res.exception!r}")
    return (outdir / "clusters.tsv").read_text()


def assert_matches(case: dict, expected_rel: str, got_text: str):
    want_path = GOLDEN / expected_rel
    if case.get("compare") == "table":
        # the reference's own goldens are compared the way its tests compare them
        from io import StringIO
        assert pd.read_table(want_path, sep="\t").equals(pd.read_table(StringIO(got_text), sep="\t")), case["name"]
    else:
        assert got_text == want_path.read_text(), f"{case['name']}: clusters.tsv differs from {expected_rel}"


def canonical_labels(labels: np.ndarray) -> np.ndarray:
    """relabel components by their smallest member (labels may be any partition ids)."""
    labels = np.asarray(labels)
    n = labels.size
    mins = {}
    for i, l in enumerate(labels.tolist()):
        if l not in mins:
            mins[l] = i
    return np.array([mins[l] for l in labels.tolist()], dtype=np.int32) if n else np.zeros(0, np.int32)


def merge_labels_cpu(gathered: np.ndarray) -> np.ndarray:
    """Semantics of the multi-rank label merge, on the CPU, for checking: union(i, gathered[r][i])."""
    import oracle
    world, n = gathered.shape
    src = np.tile(np.arange(n, dtype=np.int32), world)
    dst = np.ascontiguousarray(gathered, dtype=np.int32).ravel()
    return oracle.components(n, src, dst)


from oracle.engine_standin import OracleEngine  # noqa: E402,F401  (re-exported for the tests)


# ---------------------------------------------------------------------------------------------------------------
# scale goldens (tests/golden/scale.json): tables are regenerated from a recipe, only digests are committed
# ---------------------------------------------------------------------------------------------------------------
def scale_table(recipe: dict) -> str:
    """one seeded synthetic table as TSV text"""
    from breakfast_b200 import synth
    prof = synth.generate(recipe["n"], seed=recipe["seed"], with_mult=recipe["with_mult"])
    return prof.table(**recipe["table"]).to_csv(sep="\t", index=False)


def scale_chain(recipe: dict) -> list:
    """three tables for a cached run chain: step 0; step 1 = step 0 minus ~12 % of its sequences (whole profiles
    disappear -> ghost lists, incl. a planted ghost triple) plus new profiles plus modified sequences, shuffled;
    step 2 = step 1 minus ~10 % plus more new profiles"""
    from breakfast_b200 import synth
    rng = np.random.default_rng(recipe["seed"])
    df0 = synth.generate(recipe["n"], seed=recipe["seed"], with_mult=True).table("covsonar_dna", " ")
    a = "C1000T G2000A T3000C A4000G C5000T"
    ghost = pd.DataFrame({"accession": ["ghostA1", "ghostA2", "ghostX1", "ghostB1", "ghostB2"],
                          "dna_profile": [a, a, a + " G6000A", a + " G6000A T7000C", a + " G6000A T7000C"]})
    df0 = pd.concat([df0, ghost], ignore_index=True)
    ids = df0["accession"].to_numpy()
    keep = rng.random(len(df0)) > 0.12
    keep[ids == "ghostX1"] = False
    keep[np.isin(ids, ["ghostA1", "ghostA2", "ghostB1", "ghostB2"])] = True
    df1 = df0[keep].copy()
    extra = synth.generate(recipe["n_new1"], seed=recipe["seed"] + 1, with_mult=True).table("covsonar_dna", " ")
    extra["accession"] = ["new1_" + s for s in extra["accession"]]
    mod = rng.choice(len(df1), size=recipe["n_modified"], replace=False)
    df1.iloc[mod, 1] = [p + " A12345C" if p else "A12345C" for p in df1.iloc[mod, 1]]
    df1 = pd.concat([df1, extra], ignore_index=True).sample(frac=1.0, random_state=3).reset_index(drop=True)
    keep2 = rng.random(len(df1)) > 0.10
    extra2 = synth.generate(recipe["n_new2"], seed=recipe["seed"] + 2, with_mult=False).table("covsonar_dna", " ")
    extra2["accession"] = ["new2_" + s for s in extra2["accession"]]
    df2 = pd.concat([extra2, df1[keep2]], ignore_index=True)
    return [d.to_csv(sep="\t", index=False) for d in (df0, df1, df2)]
