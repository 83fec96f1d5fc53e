"""Times the UNMODIFIED reference (imported from /root/reference/src; build container only, it cannot travel to the GPU
box) on BASELINE configs 1 and 2 in full, with the survey's stage timers, and writes
profiles/r02_reference_cpu_c1_c2.json.  Orientation numbers for DESIGN.md: this container's cores, not the GPU box's.

    python tools/time_reference_cpu.py [--jobs 8]
"""
import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    os.environ["OMP_NUM_THREADS"] = str(args.jobs)          # before scikit-learn is imported (console.py:150)
    from breakfast_b200 import synth                        # table generator only
    sys.path.insert(0, "/root/reference/src")
    import breakfast.breakfast as ref                       # the reference itself
    assert "/root/reference" in ref.__file__
    out = {"jobs": args.jobs, "cpu": open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t")}

    def run(name, path, sep2, max_dist, opts):
        t = {}
        t0 = time.perf_counter()
        meta = ref.read_input(path, "\t", "accession", "dna_profile")
        t["read"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        meta["feature"] = ref.filter_features(meta["feature"], sep2, "covsonar_dna", *opts)
        t["filter"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        nd = ref.collapse_duplicates(meta)
        t["dedup"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        fm = ref.sparse_feature_matrix(nd["feature"], sep2)
        nd["n_features"] = fm.sum(axis=1)
        t["vectorise"] = time.perf_counter() - t0
        import numpy as np
        card = np.asarray(nd["n_features"]).ravel()
        t0 = time.perf_counter()
        neigh, evaluations = [], 0
        for q in nd["n_features"].drop_duplicates():
            evaluations += int(np.isclose(card, q, atol=max_dist).sum()) ** 2
            neigh += ref.get_neighbours_batch(fm, nd["n_features"], q, max_dist, None)
        t["neighbours"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        from networkx import connected_components
        comps = list(connected_components(ref._to_graph(neigh)))
        t["graph_cc"] = time.perf_counter() - t0
        h = np.bincount(card.astype(np.int64))
        band = int(sum(int(c) * (int(c) - 1) // 2 for c in h))
        for k in range(1, max_dist + 1):
            band += int(sum(int(h[c]) * int(h[c + k]) for c in range(len(h) - k)))
        out[name] = {"sequences": int(len(meta)), "unique_profiles": int(len(nd)), "seconds": t, "evaluations": evaluations,
                     "evaluations_per_s": evaluations / t["neighbours"], "candidate_pairs": band,
                     "candidate_pairs_per_s": band / t["neighbours"], "components": len(comps)}
        print(name, json.dumps(out[name]), flush=True)

    run("config1_testfile", ROOT / "tests" / "golden" / "reference" / "testfile.tsv", " ", 1, (True, True, 264, 228, 29903))
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp) / "c2.tsv"
        synth.generate(100_000, seed=2, with_mult=False).table("covsonar_dna", " ").to_csv(path, sep="\t", index=False)
        run("config2_100k_d1_skipdel", path, " ", 1, (True, True, 264, 228, 29903))
    (ROOT / "profiles" / "r02_reference_cpu_c1_c2.json").write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
