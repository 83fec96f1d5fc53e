"""Expected outputs for BASELINE configs 4 and 5 at their stated size -> tests/golden/configs.json.

The reference itself cannot produce these in reasonable time (config 5's first run alone is about 1.5e12 distance
evaluations, SURVEY.md section 6: days on this box), so the expectation comes from the CPU oracle pipeline:
oracle/ref_port.py (the reference's host steps restated: read, filter, dedup, cache re-indexing with ghost lists, size
filter, renumbering, output) with oracle/hashjoin.py + oracle.c for the distance and component part (core="hashjoin").
That combination reproduces every reference-made golden of this repo, cached chains included
(tests/test_oracle_golden.py).  Only recipes and SHA-256 digests are committed; tests/helpers.py regenerates the tables.

    python tests/golden/make_config_golden.py            (about half an hour on 8 cores, 40 GB of RAM)
"""
import hashlib
import json
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_port          # noqa: E402
from tests import helpers            # noqa: E402


def sha(text):
    return hashlib.sha256(text.encode()).hexdigest()


def main():
    out = {"made_by": "tests/golden/make_config_golden.py (oracle/ref_port.py, core='hashjoin')"}
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        t0 = time.time()
        text = helpers.config4_table()
        path = tmp / "c4.tsv"
        path.write_text(text)
        got, _ = ref_port.run_file(path, core="hashjoin", **helpers.CONFIG4_OPTS)
        out["config4"] = {"n_profiles": 500_000, "opts": helpers.CONFIG4_OPTS, "table_sha256": sha(text),
                          "n_lines": got.count("\n"), "sha256": sha(got)}
        print(f"config 4: {got.count(chr(10))} lines, {time.time() - t0:.0f} s", flush=True)

        t0 = time.time()
        text0, text1 = helpers.config5_tables()
        p0, p1 = tmp / "c5_0.tsv", tmp / "c5_1.tsv"
        p0.write_text(text0)
        p1.write_text(text1)
        got0, cache = ref_port.run_file(p0, core="hashjoin", want_cache=True, **helpers.CONFIG5_OPTS)
        print(f"config 5 step 0: {got0.count(chr(10))} lines, {time.time() - t0:.0f} s", flush=True)
        got1, _ = ref_port.run_file(p1, core="hashjoin", cache=cache, **helpers.CONFIG5_OPTS)
        print(f"config 5 step 1 (cached): {got1.count(chr(10))} lines, {time.time() - t0:.0f} s", flush=True)
        del cache
        fresh1, _ = ref_port.run_file(p1, core="hashjoin", **helpers.CONFIG5_OPTS)
        print(f"config 5 step 1 (fresh): {time.time() - t0:.0f} s", flush=True)
        assert sha(fresh1) != sha(got1), "the ghost lists of the cache must change the answer"
        out["config5"] = {"n_profiles": 1_000_000, "n_delta": 50_000, "opts": helpers.CONFIG5_OPTS,
                          "table_sha256": [sha(text0), sha(text1)], "n_lines": [got0.count("\n"), got1.count("\n")],
                          "sha256": [sha(got0), sha(got1)], "fresh_step1_sha256": sha(fresh1)}
    (ROOT / "tests" / "golden" / "configs.json").write_text(json.dumps(out, indent=1) + "\n")
    print("written tests/golden/configs.json")


if __name__ == "__main__":
    main()
