"""Kernel variants side by side on the headline workload (1 M profiles, max-dist 1): phase times of bf_run for
combinations of the tuning options.  Run under `ncu --metrics gpu__time_duration.sum` for per-kernel times.
    python tools/variants.py [n_profiles] [max_dist] [key=v1,v2 ...]"""
import itertools
import sys

sys.path.insert(0, ".")
from breakfast_b200 import _native, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1
grid = {}
for arg in sys.argv[3:]:
    k, v = arg.split("=")
    grid[k] = [int(x) for x in v.split(",")]
if not grid:
    grid = {"resident_csr16": [0, 1]}
ip, ix, nc = synth.generate(n, seed=1).csr()
for combo in itertools.product(*grid.values()):
    opts = dict(zip(grid.keys(), combo))
    with _native.Context(sketch_bits=128, **opts) as ctx:
        ctx.upload_csr(ip, ix, nc)
        for _ in range(3):
            st = ctx.run_sync(d)
        print(opts, {k: round(getattr(st, k), 4) for k in ("ms_total", "ms_sort", "ms_pack", "ms_pairs", "ms_verify", "ms_cc")},
              "edges", st.n_edges, flush=True)
