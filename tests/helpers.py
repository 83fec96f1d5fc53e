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


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configs 4 and 5 at their stated size (tests/golden/configs.json, made by tests/golden/make_config_golden.py)
# ---------------------------------------------------------------------------------------------------------------
CONFIG4_OPTS = dict(max_dist=1, sep2=",", id_col="seqName", clust_col="substitutions", var_type="nextclade_dna",
                    skip_ins=True, skip_del=True, trim_start=3000, trim_end=3000)
CONFIG5_OPTS = dict(max_dist=2, min_cluster_size=5)


def frame_to_tsv(df) -> str:
    """df.to_csv(sep="\\t", index=False) for a frame of plain strings - the same bytes, several times faster at 10^6 rows
    (the csv module would only quote a field that holds a tab, a double quote or a line break: then pandas writes it)"""
    cols = [df[c].tolist() for c in df.columns]
    plain = all(isinstance(v, str) and not any(ch in v for ch in '\t"\n\r') for col in cols for v in col) and \
        all(isinstance(c, str) and not any(ch in c for ch in '\t"\n\r') for c in df.columns)
    if not plain or not cols or not cols[0]:
        return df.to_csv(sep="\t", index=False)
    return "\t".join(df.columns) + "\n" + "\n".join(map("\t".join, zip(*cols))) + "\n"


def config4_table(n=500_000) -> str:
    """config 4: n unique nextclade_dna profiles with indels in the clustered column (several sequences per profile),
    to be run with --skip-ins --skip-del and wide trims (profiles collapse)"""
    from breakfast_b200 import synth
    prof = synth.generate(n, seed=4, with_mult=True, unique_on_all_events=True)
    return frame_to_tsv(prof.table("nextclade_dna", ",", id_col="seqName", feature_col="substitutions"))


def _new_substitution(rng, profile_tokens):
    from breakfast_b200 import synth
    while True:
        pos = int(rng.integers(265, 29675))
        ref = synth.ref_base(pos)
        alt = "ACGT"[int(rng.integers(0, 4))]
        tok = f"{ref}{pos}{alt}"
        if alt != ref and tok not in profile_tokens:
            return tok


def config5_tables(n=1_000_000, n_delta=50_000, seed=5):
    """config 5: (step 0 table, step 1 table).  Step 0 = n unique covsonar_dna profiles with multiplicities (config-3-like,
    seed 5) plus a planted ghost triple.  Step 1 = step 0 with n_delta sequences touched: 60 % added (half of them new
    unique profiles one or two substitutions away from an existing profile, half further sequences of existing
    profiles), 20 % modified (their profile gains a substitution; a profile whose only sequence is modified vanishes), 20 %
    deleted (most profiles have a single sequence, so whole profiles vanish and leave ghost lists), shuffled."""
    from breakfast_b200 import synth
    rng = np.random.default_rng(seed)
    df0 = synth.generate(n, seed=seed, with_mult=True).table("covsonar_dna", " ")
    a = "C1000T G2000A T3000C A4000G C5000T"
    ghost = pd.DataFrame({"accession": ["ghostA1", "ghostA2", "ghostA3", "ghostA4", "ghostA5", "ghostX1",
                                        "ghostB1", "ghostB2", "ghostB3", "ghostB4", "ghostB5"],
                          "dna_profile": [a] * 5 + [a + " G6000A T7000C"] + [a + " G6000A T7000C A8000G C9000T"] * 5})
    df0 = pd.concat([df0, ghost], ignore_index=True)
    ids = df0["accession"].to_numpy(dtype=object)
    feats = df0["dna_profile"].to_numpy(dtype=object)
    m = len(df0)
    n_add, n_mod, n_del = int(0.6 * n_delta), int(0.2 * n_delta), int(0.2 * n_delta)
    touched = rng.choice(m - len(ghost), size=n_mod + n_del, replace=False)
    mod_rows, del_rows = touched[:n_mod], touched[n_mod:]
    feats1 = feats.copy()
    for r in mod_rows:
        toks = feats1[r].split(" ") if feats1[r] else []
        feats1[r] = " ".join(toks + [_new_substitution(rng, set(toks))])
    keep = np.ones(m, dtype=bool)
    keep[del_rows] = False
    keep[ids == "ghostX1"] = False                      # the bridge of the ghost triple disappears
    src_rows = rng.choice(m - len(ghost), size=n_add, replace=False)
    new_feats = []
    for k, r in enumerate(src_rows):
        if k % 2 == 0:                                  # a new profile at distance 1 or 2 of an existing one
            toks = feats[r].split(" ") if feats[r] else []
            extra = [_new_substitution(rng, set(toks))]
            if k % 4 == 0:
                extra.append(_new_substitution(rng, set(toks) | set(extra)))
            new_feats.append(" ".join(toks + extra))
        else:                                           # one more sequence of an existing profile
            new_feats.append(feats[r])
    added = pd.DataFrame({"accession": [f"add{k:06d}" for k in range(n_add)], "dna_profile": new_feats})
    df1 = pd.concat([pd.DataFrame({"accession": ids[keep], "dna_profile": feats1[keep]}), added], ignore_index=True)
    df1 = df1.iloc[rng.permutation(len(df1))].reset_index(drop=True)
    return frame_to_tsv(df0), frame_to_tsv(df1)
