"""Scale goldens: run the UNMODIFIED reference (rki-mf1/breakfast, /root/reference) on seeded synthetic tables that are
too large to commit, and commit only the recipe of each table and the SHA-256 of the clusters.tsv the reference wrote.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_scale_golden.py          # about five minutes

The tests regenerate the tables from the recipe (breakfast_b200.synth is deterministic) and compare digests:
tests/test_scale_golden.py runs the host pipeline with the CPU oracle as the engine, tests/test_gpu_parity.py the
product CLI on the GPU.  Writes tests/golden/scale.json.
"""
from __future__ import annotations

import hashlib
import json
import sys
import tempfile
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import ROOT, run_reference  # noqa: E402  (puts the reference first on sys.path)
from tests.helpers import scale_chain, scale_table  # noqa: E402

NC = dict(sep2=",", id_col="seqName", clust_col="substitutions", var_type="nextclade_dna")
TABLES = {
    "scale_dna": dict(n=40000, seed=21, with_mult=True, table=dict()),
    "scale_nextclade": dict(n=25000, seed=22, with_mult=True,
                            table=dict(var_type="nextclade_dna", sep=",", id_col="seqName", feature_col="substitutions")),
}
CASES = [
    ("scale_dna_d1", "scale_dna", dict()),
    ("scale_dna_d2_m5", "scale_dna", dict(max_dist=2, min_cluster_size=5)),
    ("scale_dna_d2_noskipdel", "scale_dna", dict(max_dist=2, skip_del=False)),
    ("scale_nextclade_d1", "scale_nextclade", dict(max_dist=1, **NC)),
]


CHAINS = {
    "scale_cache": dict(n=15000, seed=31, n_new1=3000, n_new2=1000, n_modified=300),
}
CHAIN_CASES = [("scale_cache_d1", "scale_cache", dict(max_dist=1)), ("scale_cache_d2", "scale_cache", dict(max_dist=2))]


def digest(text: str) -> str:
    return hashlib.sha256(text.encode()).hexdigest()


def main():
    out = {"tables": {}, "cases": [], "chains": {}, "chain_cases": []}
    with tempfile.TemporaryDirectory() as tmp:
        paths = {}
        for name, recipe in TABLES.items():
            text = scale_table(recipe)
            paths[name] = Path(tmp) / f"{name}.tsv"
            paths[name].write_text(text)
            out["tables"][name] = dict(recipe, n_sequences=text.count("\n") - 1,
                                       sha256=hashlib.sha256(text.encode()).hexdigest())
        for case, table, opts in CASES:
            t = time.time()
            got = run_reference(paths[table], opts, None, None)
            n_clusters = len({line.split("\t")[1] for line in got.splitlines()[1:] if line.split("\t")[1]})
            out["cases"].append(dict(name=case, table=table, opts=opts, sha256=hashlib.sha256(got.encode()).hexdigest(),
                                     n_lines=got.count("\n"), n_clusters=n_clusters))
            print(f"{case}: {got.count(chr(10))} lines, {n_clusters} clusters, reference took {time.time() - t:.0f} s")
        # cached run chains: step 0 writes a cache, step 1 reads it and writes the next, step 2 reads that
        for name, recipe in CHAINS.items():
            texts = scale_chain(recipe)
            for k, text in enumerate(texts):
                (Path(tmp) / f"{name}_{k}.tsv").write_text(text)
            out["chains"][name] = dict(recipe, sha256=[digest(t) for t in texts],
                                       n_sequences=[t.count("\n") - 1 for t in texts])
        for case, chain, opts in CHAIN_CASES:
            t = time.time()
            c0, c1 = Path(tmp) / f"{case}.c0", Path(tmp) / f"{case}.c1"
            steps = [Path(tmp) / f"{chain}_{k}.tsv" for k in range(3)]
            got = [run_reference(steps[0], opts, None, c0), run_reference(steps[1], opts, c0, c1),
                   run_reference(steps[2], opts, c1, None)]
            fresh1 = run_reference(steps[1], opts, None, None)
            out["chain_cases"].append(dict(name=case, chain=chain, opts=opts, sha256=[digest(g) for g in got],
                                           n_lines=[g.count("\n") for g in got], fresh_step1_sha256=digest(fresh1)))
            print(f"{case}: cached step 1 differs from a fresh run (ghost lists visible): {got[1] != fresh1}; "
                  f"reference took {time.time() - t:.0f} s")
    (HERE / "scale.json").write_text(json.dumps(out, indent=1) + "\n")
    print(f"wrote {HERE / 'scale.json'} (repo root {ROOT})")


if __name__ == "__main__":
    main()
