"""tensor-core level 1 (option level1=1) against the oracle and against level1=0; then timing at 1M."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle
from breakfast_b200 import _native, synth

ok = True
for n, d in ((130, 1), (1000, 1), (5000, 1), (5000, 2), (5000, 3), (20000, 1)):
    indptr, indices, n_cols = synth.generate(n, seed=50 + n).csr()
    want, ne = oracle.cluster(indptr, indices, d)
    for bits in (128, 256):
        with _native.Context(sketch_bits=bits, level1=1, want_edges=1) as ctx:
            ctx.upload_csr(indptr, indices, n_cols)
            st = ctx.run_sync(d)
            lab = ctx.download_labels()
            good = np.array_equal(lab, want) and st.n_edges == ne
            print(f"n={n} d={d} bits={bits}: {'OK' if good else 'MISMATCH'} edges={st.n_edges}/{ne} units={st.l2_warp_items} cand={st.n_candidates}", flush=True)
            ok &= good
    q = np.arange(0, n, 3, dtype=np.int32)
    with _native.Context(level1=1, want_edges=1) as ctx:
        ctx.upload_csr(indptr, indices, n_cols, query_rows=q)
        ctx.run_sync(d)
        src, dst = ctx.download_edges()
    ws, wd = oracle.edges(indptr, indices, d, queries=q)
    good = np.array_equal(src, ws) and np.array_equal(dst, wd)
    print(f"n={n} d={d} rectangle: {'OK' if good else 'MISMATCH'}", flush=True)
    ok &= good
print("PARITY", "OK" if ok else "FAILED", flush=True)
if ok and "--big" in sys.argv:
    indptr, indices, n_cols = synth.generate(1_000_000, seed=1).csr()
    res = {}
    for lvl in (0, 1):
        with _native.Context(level1=lvl) as ctx:
            ctx.upload_csr(indptr, indices, n_cols)
            for _ in range(3):
                ctx.run_sync(1)
            ts = []
            for _ in range(10):
                st = ctx.run_sync(1)
                ts.append((st.ms_total, st.ms_pairs, st.ms_l1_sum, st.ms_pack))
            res[lvl] = ctx.download_labels()
            t = np.median(np.array(ts), axis=0)
            print(f"level1={lvl}: total {t[0]:.3f} ms pairs {t[1]:.3f} l1 {t[2]:.3f} pack {t[3]:.3f} units={st.l2_warp_items} cand={st.n_candidates} edges={st.n_edges}", flush=True)
    print("labels equal:", np.array_equal(res[0], res[1]))
sys.exit(0 if ok else 1)
