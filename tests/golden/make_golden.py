"""Generate tests/golden/* by running the UNMODIFIED reference (rki-mf1/breakfast, /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

Writes
  reference/…            the reference's own test tables and expected_clusters_*.tsv (golden vectors the
                          reference's tests hold for this path) — re-checked here against a live reference run
  synthetic/…            small seeded inputs (breakfast_b200.synth + hand-written quirk tables) and the
                          clusters.tsv the reference produced for them
  manifest.json          one entry per case: input, options, expected output, cache chaining
"""
from __future__ import annotations

import gzip
import json
import shutil
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REF / "src"))
sys.path.insert(1, str(ROOT))

import click.testing  # noqa: E402
import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402
from breakfast import console as ref_console  # noqa: E402  (the reference, see sys.path)

assert "/root/reference" in ref_console.__file__, ref_console.__file__
from breakfast_b200 import synth  # noqa: E402

DEFAULTS = dict(sep="\t", id_col="accession", clust_col="dna_profile", var_type="covsonar_dna", sep2=" ",
                max_dist=1, min_cluster_size=2, trim_start=264, trim_end=228, reference_length=29903,
                skip_del=True, skip_ins=True)


def cli_args(opts: dict) -> list:
    """options dict -> CLI arguments (only what differs from the defaults, like a user would type)."""
    a = []
    flag = {"id_col": "--id-col", "clust_col": "--clust-col", "var_type": "--var-type", "sep2": "--sep2",
            "max_dist": "--max-dist", "min_cluster_size": "--min-cluster-size", "trim_start": "--trim-start",
            "trim_end": "--trim-end", "reference_length": "--reference-length", "sep": "--sep"}
    for k, v in opts.items():
        if k in flag:
            a += [flag[k], str(v)]
        elif k == "skip_del":
            a.append("--skip-del" if v else "--no-skip-del")
        elif k == "skip_ins":
            a.append("--skip-ins" if v else "--no-skip-ins")
        else:
            raise KeyError(k)
    return a


def run_reference(input_file: Path, opts: dict, cache_in: Path | None, cache_out: Path | None) -> str:
    with tempfile.TemporaryDirectory() as tmp:
        args = ["--input-file", str(input_file), "--outdir", tmp] + cli_args(opts)
        if cache_in:
            args += ["--input-cache", str(cache_in)]
        if cache_out:
            args += ["--output-cache", str(cache_out)]
        res = click.testing.CliRunner().invoke(ref_console.main, args)
        if res.exit_code != 0:
            raise RuntimeError(f"reference failed on {input_file} {args}: {res.output}\n{res.exception!r}")
        return (Path(tmp) / "clusters.tsv").read_text()


def write_gz(path: Path, text: str):
    with open(path, "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as f:
        f.write(text.encode())


def table_text(df: pd.DataFrame) -> str:
    return df.to_csv(sep="\t", index=False)


def main():
    cases = []
    ref_dir, syn_dir = HERE / "reference", HERE / "synthetic"
    for d in (ref_dir, syn_dir):
        shutil.rmtree(d, ignore_errors=True)
        d.mkdir(parents=True)

    # ---------------------------------------------------------------- A. the reference's own goldens
    for f in sorted((REF / "tests").glob("*.tsv")):
        shutil.copyfile(f, ref_dir / f.name)
    nc = dict(sep2=",", id_col="seqName", clust_col="substitutions", var_type="nextclade_dna")
    own = [
        ("ref_dist0", "testfile.tsv", dict(max_dist=0), "expected_clusters_dist0.tsv"),
        ("ref_dist1", "testfile.tsv", dict(max_dist=1), "expected_clusters_dist1.tsv"),
        ("ref_dist1_noskipdel", "testfile.tsv", dict(max_dist=1, skip_del=False), "expected_clusters_dist1_noskipdel.tsv"),
        ("ref_raw_defaults", "testfile.tsv", dict(max_dist=1, var_type="raw"), "expected_clusters_dist1_noskipdel.tsv"),
        ("ref_raw_notrim", "testfile.tsv", dict(max_dist=1, var_type="raw", trim_start=0, trim_end=0, skip_del=False, skip_ins=False), "expected_clusters_dist1_noskipdel.tsv"),
        ("ref_nextclade_dist0", "testfile_nextclade.tsv", dict(max_dist=0, **nc), "expected_clusters_dist0.tsv"),
        ("ref_nextclade_dist1", "testfile_nextclade.tsv", dict(max_dist=1, **nc), "expected_clusters_dist1.tsv"),
    ]
    for name, inp, opts, exp in own:
        got = run_reference(ref_dir / inp, opts, None, None)
        want = pd.read_table(ref_dir / exp)
        assert want.equals(pd.read_table(pd.io.common.StringIO(got))), name
        cases.append(dict(name=name, input=f"reference/{inp}", opts=opts, expected=f"reference/{exp}", compare="table"))
    # the reference's cache tests: first run writes the cache, second run reads it
    cache0 = syn_dir / "ref_testfile_dist1.cache"
    run_reference(ref_dir / "testfile.tsv", dict(max_dist=1), None, cache0)
    for f in sorted(ref_dir.glob("testfile_caching0*.tsv")):
        num = f.name[len("testfile_caching"):len("testfile_caching") + 2]
        exp = f"expected_clusters_caching{num}_dist1.tsv"
        got = run_reference(f, dict(max_dist=1), cache0, None)
        assert pd.read_table(ref_dir / exp).equals(pd.read_table(pd.io.common.StringIO(got))), f.name
        cases.append(dict(name=f"ref_caching{num}", input=f"reference/{f.name}", opts=dict(max_dist=1),
                          expected=f"reference/{exp}", compare="table",
                          cache_from=dict(input="reference/testfile.tsv", opts=dict(max_dist=1)),
                          reference_cache_file="synthetic/ref_testfile_dist1.cache"))

    # ---------------------------------------------------------------- B. synthetic, reference-produced
    def add(name, text, opts, chain=None, gz=True):
        inp = syn_dir / (f"{name}.tsv.gz" if gz else f"{name}.tsv")
        if not inp.exists():
            write_gz(inp, text) if gz else inp.write_text(text)
        return inp

    prof = synth.generate(420, seed=11, with_mult=True)
    t_dna = table_text(prof.table("covsonar_dna", " "))
    base = add("syn_dna", t_dna, {})
    for name, opts in [
        ("syn_dna_d1", dict(max_dist=1)),
        ("syn_dna_d2_m5", dict(max_dist=2, min_cluster_size=5)),
        ("syn_dna_d3_m1", dict(max_dist=3, min_cluster_size=1)),
        ("syn_dna_d0_m3", dict(max_dist=0, min_cluster_size=3)),
        ("syn_dna_d2_noskip", dict(max_dist=2, skip_del=False, skip_ins=False)),
        ("syn_dna_d1_notrim_noskipdel", dict(max_dist=1, trim_start=0, trim_end=0, skip_del=False)),
        ("syn_dna_d4_raw", dict(max_dist=4, var_type="raw", min_cluster_size=3)),
    ]:
        exp = syn_dir / f"{name}.expected.tsv"
        exp.write_text(run_reference(base, opts, None, None))
        cases.append(dict(name=name, input=f"synthetic/{base.name}", opts=opts, expected=f"synthetic/{exp.name}", compare="bytes"))

    prof_nc = synth.generate(380, seed=12, with_mult=True, unique_on_all_events=True)
    t_nc = table_text(prof_nc.table("nextclade_dna", ",", id_col="seqName", feature_col="substitutions"))
    base_nc = add("syn_nextclade", t_nc, {})
    for name, opts in [("syn_nextclade_d1", dict(max_dist=1, **nc)),
                       ("syn_nextclade_d2_noskip", dict(max_dist=2, skip_del=False, skip_ins=False, **nc))]:
        exp = syn_dir / f"{name}.expected.tsv"
        exp.write_text(run_reference(base_nc, opts, None, None))
        cases.append(dict(name=name, input=f"synthetic/{base_nc.name}", opts=opts, expected=f"synthetic/{exp.name}", compare="bytes"))

    # quirks: repeated tokens (L1 on counts), permuted tokens, empty / NA profiles, double separators,
    # unparsable tokens, trimming boundaries, insertion shapes
    P1, P2 = "C1000T C1100T C1200T C1300T", "T2000A T2100A T2200A T2300A"
    P3, P4 = "G3000T G3100T G3200T", "A4000C A4100C A4200C"
    quirk_rows = [
        ("q01", f"{P1} C300T G400A"), ("q02", f"G400A {P1} C300T"), ("q03", f"{P1} C300T G400A G400A"),
        ("q04", f"{P1} C300T G400A G400A G400A"), ("q05", ""), ("q06", "NA"), ("q07", "C300T"),
        ("q08", f"{P1} C300T  G400A"), ("q09", f"{P1} C300T G400A bogus"),
        ("q10", f"{P2} C264T C265T"), ("q11", f"{P2} C265T"), ("q12", f"{P2} C29674T C29675T"),
        ("q13", f"{P2} C29674T"), ("q14", f"{P1} C300T G400A A500AT"), ("q15", f"{P1} C300T G400A del:600:3"),
        ("q16", f"{P1} C300T G400A T700C"), ("q17", f"{P1} C300T G400A T700C A800G"),
        ("q18", f"{P2} C265T C29674T"), ("q19", f"{P3} del:600:3"), ("q20", f"{P3} A500AT"),
        ("q21", f"{P1} C300T G400A"), ("q22", f"{P4} T900G T900G"), ("q23", f"{P4} T900G"), ("q24", "c300t"),
        ("q25", P3), ("q26", P4), ("q27", f"{P4} T900G T900G T900G T900G"), ("q28", f"{P2} C264T"),
        ("q29", f"{P3} G3300T G3400T G3500T"), ("q30", f"{P3} del:600:3 A500AT"),
    ]
    t_quirk = "accession\tdna_profile\n" + "".join(f"{a}\t{b}\n" for a, b in quirk_rows)
    base_q = add("quirks", t_quirk, {}, gz=False)
    for name, opts in [
        ("quirks_d1", dict(max_dist=1)), ("quirks_d1_m1", dict(max_dist=1, min_cluster_size=1)),
        ("quirks_d2", dict(max_dist=2)), ("quirks_d0", dict(max_dist=0, min_cluster_size=1)),
        ("quirks_d1_nofilter", dict(max_dist=1, trim_start=0, trim_end=0, skip_del=False, skip_ins=False)),
        ("quirks_d1_noskipins", dict(max_dist=1, skip_ins=False)),
        ("quirks_d1_raw", dict(max_dist=1, var_type="raw")),
    ]:
        exp = syn_dir / f"{name}.expected.tsv"
        exp.write_text(run_reference(base_q, opts, None, None))
        cases.append(dict(name=name, input=f"synthetic/{base_q.name}", opts=opts, expected=f"synthetic/{exp.name}", compare="bytes"))

    # cache chain with deletions (ghost lists), additions and modifications
    rng = np.random.default_rng(5)
    prof_c = synth.generate(300, seed=13, with_mult=True)
    df0 = prof_c.table("covsonar_dna", " ")
    # planted ghost triple: A, X = A + g, B = A + g + h ; X disappears in step 1
    A = "C1000T G2000A T3000C A4000G C5000T"
    ghost = pd.DataFrame({"accession": ["ghostA1", "ghostA2", "ghostX1", "ghostB1", "ghostB2"],
                          "dna_profile": [A, A, A + " G6000A", A + " G6000A T7000C", A + " G6000A T7000C"]})
    df0 = pd.concat([df0, ghost], ignore_index=True)
    step0 = add("cache_step0", table_text(df0), {})
    keep = rng.random(len(df0)) > 0.12
    keep[df0["accession"].to_numpy() == "ghostX1"] = False
    keep[np.isin(df0["accession"].to_numpy(), ["ghostA1", "ghostA2", "ghostB1", "ghostB2"])] = True
    df1 = df0[keep].copy()
    extra = synth.generate(90, seed=14, with_mult=True).table("covsonar_dna", " ")
    extra["accession"] = ["new1_" + s for s in extra["accession"]]
    # a few modified sequences: same id, a neighbour's profile plus one substitution
    mod_idx = rng.choice(len(df1), size=12, replace=False)
    df1.iloc[mod_idx, 1] = [p + " A12345C" if p else "A12345C" for p in df1.iloc[mod_idx, 1]]
    df1 = pd.concat([df1, extra], ignore_index=True).sample(frac=1.0, random_state=3).reset_index(drop=True)
    step1 = add("cache_step1", table_text(df1), {})
    keep2 = rng.random(len(df1)) > 0.10
    df2 = df1[keep2].copy()
    extra2 = synth.generate(40, seed=15, with_mult=False).table("covsonar_dna", " ")
    extra2["accession"] = ["new2_" + s for s in extra2["accession"]]
    df2 = pd.concat([extra2, df2], ignore_index=True)
    step2 = add("cache_step2", table_text(df2), {})
    for d in (1, 2):
        opts = dict(max_dist=d)
        with tempfile.TemporaryDirectory() as tmp:
            c0, c1 = Path(tmp) / "c0", Path(tmp) / "c1"
            e0 = syn_dir / f"cache_d{d}_step0.expected.tsv"
            e1 = syn_dir / f"cache_d{d}_step1.expected.tsv"
            e2 = syn_dir / f"cache_d{d}_step2.expected.tsv"
            e1_fresh = syn_dir / f"cache_d{d}_step1_fresh.expected.tsv"
            e0.write_text(run_reference(step0, opts, None, c0))
            e1.write_text(run_reference(step1, opts, c0, c1))
            e2.write_text(run_reference(step2, opts, c1, None))
            e1_fresh.write_text(run_reference(step1, opts, None, None))
            if d == 1:
                shutil.copyfile(c0, syn_dir / "cache_d1_step0.reference.cache")
        cases.append(dict(name=f"cache_d{d}_chain", compare="bytes", opts=opts, chain=[
            dict(input=f"synthetic/{step0.name}", expected=f"synthetic/{e0.name}"),
            dict(input=f"synthetic/{step1.name}", expected=f"synthetic/{e1.name}"),
            dict(input=f"synthetic/{step2.name}", expected=f"synthetic/{e2.name}")]))
        cases.append(dict(name=f"cache_d{d}_step1_fresh", input=f"synthetic/{step1.name}", opts=opts,
                          expected=f"synthetic/{e1_fresh.name}", compare="bytes"))
        differs = e1.read_text() != e1_fresh.read_text()
        print(f"d={d}: cached step1 differs from fresh step1 (ghost lists visible): {differs}")
    cases.append(dict(name="cache_d1_from_reference_file", compare="bytes", opts=dict(max_dist=1),
                      input=f"synthetic/{step1.name}", expected="synthetic/cache_d1_step1.expected.tsv",
                      reference_cache_file="synthetic/cache_d1_step0.reference.cache"))

    (HERE / "manifest.json").write_text(json.dumps({"defaults": DEFAULTS, "cases": cases}, indent=1) + "\n")
    total = sum(f.stat().st_size for f in HERE.rglob("*") if f.is_file())
    print(f"{len(cases)} cases, {total / 1024:.0f} KiB under {HERE}")


if __name__ == "__main__":
    main()
