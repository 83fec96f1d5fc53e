"""`breakfast` command line — drop-in for the reference CLI (src/breakfast/console.py:10-170):
same option names, defaults, types, environment binding (--jobs <-> OMP_NUM_THREADS) and the same
rules for non-DNA feature types; the five pipeline steps run in the same order.  The clustering
step executes on the GPU through libbreakfast_b200.so.
"""
from __future__ import annotations

import os
import pathlib

import click
from click.core import ParameterSource

from . import __version__, breakfast

import time

VAR_TYPES = ("covsonar_dna", "covsonar_aa", "nextclade_dna", "nextclade_aa", "raw")
# wall-clock of the five pipeline steps of the last run (seconds), for benchmarks: read, prepare (filter + dedup
# [+ CSR on the native host path]), cluster (CSR if not done yet, engine, cache, labels), write
LAST_TIMINGS: dict = {}
DNA_TYPES = ("covsonar_dna", "nextclade_dna")


def _options():
    path = click.Path(path_type=pathlib.Path)
    existing = click.Path(exists=True, path_type=pathlib.Path)
    nonneg = click.IntRange(0)
    spec = [
        (("--input-file",), dict(type=existing, required=True, help="Table with one sequence per line")),
        (("--sep",), dict(default="\t", help="Column separator of the input table")),
        (("--outdir",), dict(type=path, default="output", help="Directory that receives clusters.tsv")),
        (("--max-dist",), dict(type=nonneg, default=1, help="Largest pairwise distance that still links two profiles")),
        (("--min-cluster-size",), dict(type=click.IntRange(1), default=2, help="Smallest cluster (in sequences) that is reported")),
        (("--input-cache",), dict(type=existing, help="Cache written by an earlier run (--output-cache)")),
        (("--output-cache",), dict(type=path, help="Where to write the cache of this run")),
        (("--id-col",), dict(default="accession", help="Name of the sequence-id column")),
        (("--clust-col",), dict(default="dna_profile", help="Name of the column holding the mutation profile")),
        (("--var-type",), dict(type=click.Choice(list(VAR_TYPES)), default="covsonar_dna", help="Notation of the mutations")),
        (("--sep2",), dict(default=" ", help="Separator between the mutations of one profile")),
        (("--trim-start",), dict(type=nonneg, default=264, help="Ignore substitutions in the first N bases (0 = off)")),
        (("--trim-end",), dict(type=nonneg, default=228, help="Ignore substitutions in the last N bases (0 = off)")),
        (("--reference-length",), dict(type=nonneg, default=29903, help="Reference genome length (default: NC_045512.2)")),
        (("--skip-del/--no-skip-del",), dict(default=True, help="Ignore deletions")),
        (("--skip-ins/--no-skip-ins",), dict(default=True, help="Ignore insertions")),
        (("--jobs",), dict(type=click.IntRange(1), default=1, envvar="OMP_NUM_THREADS", help="Host threads (kept for compatibility; the distance work runs on the GPU)")),
        # not in the reference: how many GPUs of this box share the pairwise work (default 1)
        (("--gpus",), dict(type=click.IntRange(1), default=1, help="GPUs to partition the pairwise work over")),
    ]
    return [click.Option(list(names), **kw) for names, kw in spec]


def _explicit(ctx, name, *sources_that_do_not_count):
    return ctx.get_parameter_source(name) not in sources_that_do_not_count


def run(input_file, outdir, input_cache, output_cache, id_col, clust_col, var_type, sep, sep2, max_dist,
        min_cluster_size, trim_start, trim_end, reference_length, skip_del, skip_ins, jobs, gpus=1):
    if gpus > 1 and not os.environ.get("BREAKFAST_B200_DEVICES"):
        os.environ["BREAKFAST_B200_DEVICES"] = ",".join(str(d) for d in range(gpus))
    if var_type not in DNA_TYPES:
        # trimming / indel skipping only make sense for nucleotide positions: the DNA-oriented
        # defaults are dropped silently, an explicit request is an error
        ctx = click.get_current_context()
        if trim_start != 0 and _explicit(ctx, "trim_start", ParameterSource.DEFAULT):
            raise click.BadParameter("Can not trim non-DNA features")
        if trim_end != 0 and _explicit(ctx, "trim_end", ParameterSource.DEFAULT):
            raise click.BadParameter("Can not trim non-DNA features")
        if skip_del and ctx.get_parameter_source("skip_del") == ParameterSource.COMMANDLINE:
            raise click.BadParameter("Can not skip indels in non-DNA features")
        if skip_ins and ctx.get_parameter_source("skip_ins") == ParameterSource.COMMANDLINE:
            raise click.BadParameter("Can not skip indels in non-DNA features")
        trim_start = trim_end = 0
        skip_del = skip_ins = False

    if max(trim_start, trim_end) > reference_length:
        raise click.BadParameter("Can not trim more than the reference length")

    banner = [
        ("Input file", input_file),
        ("Input file separator", f"'{sep}'"),
        ("ID column", id_col),
        ("clustering feature type", var_type),
        ("clustering feature column", clust_col),
        ("clustering feature column separator", f"'{sep2}'"),
        ("max dist", max_dist),
        ("minimum cluster size", min_cluster_size),
        ("trim start (bp)", trim_start),
        ("trim end (bp)", trim_end),
        ("reference length (bp)", reference_length),
        ("skip deletions", skip_del),
        ("skip insertions", skip_ins),
        ("Input cache file", input_cache),
        ("Output cache file", output_cache),
    ]
    print("Clustering sequences")
    for label, value in banner:
        print(f"  {label} = {value}")

    os.environ["OMP_NUM_THREADS"] = str(jobs)

    LAST_TIMINGS.clear()
    t0 = time.perf_counter()
    meta = breakfast.read_input(input_file, sep, id_col, clust_col)
    t1 = time.perf_counter()
    meta_nodups, pre = None, None
    if os.environ.get("BREAKFAST_B200_HOST", "native") == "native" and var_type in VAR_TYPES:
        # one native pass (csrc/host_parse.cpp) instead of three Python loops over every token; same result
        from . import hostfast
        if hostfast.available():
            prepared = hostfast.prepare(meta, sep2, var_type, skip_ins, skip_del, trim_start, trim_end, reference_length)
            if prepared is not None:
                meta_nodups, pre = prepared
    if meta_nodups is None:
        meta["feature"] = breakfast.filter_features(
            meta["feature"], sep2, var_type, skip_ins, skip_del, trim_start, trim_end, reference_length
        )
        meta_nodups = breakfast.collapse_duplicates(meta)
    t2 = time.perf_counter()
    meta_clustered = breakfast.cluster(meta_nodups, sep2, max_dist, min_cluster_size, input_cache, output_cache, pre=pre)
    t3 = time.perf_counter()
    breakfast.write_output(meta_clustered, meta, outdir)
    t4 = time.perf_counter()
    LAST_TIMINGS.update(read=t1 - t0, prepare=t2 - t1, cluster=t3 - t2, write=t4 - t3, total=t4 - t0,
                        host_path="native" if pre is not None else "python", **breakfast.LAST_ENGINE_TIMINGS)


main = click.version_option(version=__version__)(
    click.Command("breakfast", params=_options(), callback=run, context_settings={"show_default": True})
)

if __name__ == "__main__":
    main()
