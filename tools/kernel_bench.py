"""Phase timings of the device pipeline on synthetic profiles, through the C ABI (ctypes).

python tools/kernel_bench.py --n 100000 --seed 2 --dist 1 --bits 128,256,512 --full --reps 20
Clocks are warmed with the pipe microbenchmarks first; every config gets warm-up runs, then
`reps` timed runs; min and median per phase are printed as one JSON line per config.
"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from breakfast_b200 import synth  # noqa: E402
from tools.first_light import Stats, ck, lib, p32, p64  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--dist", type=int, default=1)
    ap.add_argument("--bits", default="128,256,512")
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--bps", type=int, default=0)
    ap.add_argument("--two-level", type=int, default=1)
    ap.add_argument("--level1", type=int, default=0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    prof = synth.generate(args.n, seed=args.seed)
    indptr, indices, n_cols = prof.csr()
    peaks = {}
    for name in ("popc32", "lop3", "iadd3", "xor_popc_add"):
        g = C.c_double()
        ck(lib.bf_measure_peak(0, name.encode(), C.byref(g)))
        peaks[name] = g.value
    print(json.dumps({"peaks_gops": peaks, "n": args.n, "n_cols": n_cols, "nnz": int(len(indices))}))
    configs = [(0, int(b)) for b in args.bits.split(",") if b]
    if args.full:
        configs.append((1, 0))
    results = []
    ref_labels = None
    for engine, bits in configs:
        ctx = C.c_void_p()
        ck(lib.bf_ctx_create(0, None, C.byref(ctx)))
        ck(lib.bf_ctx_set_option(ctx, b"engine", C.c_int64(engine)))
        if bits:
            ck(lib.bf_ctx_set_option(ctx, b"sketch_bits", C.c_int64(bits)))
        ck(lib.bf_ctx_set_option(ctx, b"two_level", C.c_int64(args.two_level)))
        ck(lib.bf_ctx_set_option(ctx, b"level1", C.c_int64(args.level1)))
        if args.bps:
            ck(lib.bf_ctx_set_option(ctx, b"blocks_per_sm", C.c_int64(args.bps)))
        ck(lib.bf_upload_csr(ctx, p64(indptr), p32(indices), C.c_int64(args.n), C.c_int32(n_cols), None, C.c_int64(0)))
        st = Stats()
        reps = args.reps if engine == 0 else max(3, args.reps // 5)
        rows = []
        for rep in range(args.warmup + reps):
            ck(lib.bf_run(ctx, args.dist, 0, 1))
            ck(lib.bf_sync(ctx, C.byref(st)))
            if rep >= args.warmup:
                rows.append(st.as_dict())
        lab = np.empty(args.n, np.int32)
        ck(lib.bf_download_labels(ctx, p32(lab)))
        lib.bf_ctx_destroy(ctx)
        if ref_labels is None:
            ref_labels = lab
        d = dict(rows[-1])
        for k in ("ms_sort", "ms_pack", "ms_pairs", "ms_verify", "ms_cc", "ms_total"):
            v = np.array([r[k] for r in rows])
            d[k] = float(np.median(v))
            d[k + "_min"] = float(v.min())
        d["engine"] = "full" if engine else "sketch"
        d["labels_match_first"] = bool(np.array_equal(lab, ref_labels))
        t = d["ms_pairs_min"] * 1e-3
        d["cand_pairs_per_s"] = d["pairs_band"] / t
        d["popc_gops"] = d["popc32_executed"] / t / 1e9
        d["popc_frac_of_measured_peak"] = d["popc_gops"] / peaks["popc32"]
        results.append(d)
        print(json.dumps(d))
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps({"peaks_gops": peaks, "results": results}, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
