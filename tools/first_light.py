"""First-light GPU check of libbreakfast_b200.so through ctypes (no torch, no package imports).

Run on the GPU box:  python tools/first_light.py [--big N]
 1. integer-pipe peaks (POPC / LOP3 / IADD3 / fused)
 2. small parity: labels + edges vs a dense numpy brute force, both engines, several max_dist
 3. timing on a larger synthetic set
"""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
import os
lib = C.CDLL(os.environ.get("BF_LIB", str(ROOT / "breakfast_b200" / "libbreakfast_b200.so")))
lib.bf_last_error.restype = C.c_char_p


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "n_rows", "n_query", "n_cols", "nnz", "bits_per_row", "pairs_total", "pairs_band", "pairs_evaluated",
        "tiles_total", "tiles_band", "tiles_rank", "n_candidates", "n_edges", "n_components")] + [
        (n, C.c_double) for n in ("ms_h2d", "ms_sort", "ms_pack", "ms_pairs", "ms_verify", "ms_cc", "ms_merge",
                                  "ms_d2h", "ms_total")] + [
        ("runs_since_sync", C.c_int64), ("kernel_launches", C.c_int64), ("ms_pairs_sum", C.c_double),
        ("ms_total_sum", C.c_double), ("l2_warp_items", C.c_int64), ("popc32_executed", C.c_int64),
        ("ms_l1_sum", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def ck(rc):
    if rc != 0:
        raise RuntimeError(f"bf error {rc}: {lib.bf_last_error().decode()}")


def p64(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def p32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def gen(n, n_cols, n_roots, root_len, rng):
    """lineage-structured binary rows (sorted unique column ids), unique as sets"""
    rows, seen = [], set()
    roots = [np.sort(rng.choice(n_cols, size=root_len + 5 * (i % 8), replace=False)) for i in range(n_roots)]
    for r in roots:
        rows.append(r)
        seen.add(r.tobytes())
    while len(rows) < n:
        par = rows[rng.integers(len(rows))]
        k = rng.geometric(0.7)
        child = set(par.tolist())
        for _ in range(k):
            if rng.random() < 0.9 or len(child) == 0:
                child.add(int(rng.integers(n_cols)))
            else:
                child.discard(int(rng.choice(list(child))))
        c = np.array(sorted(child), dtype=np.int64)
        key = c.tobytes()
        if key in seen:
            continue
        seen.add(key)
        rows.append(c)
    order = rng.permutation(len(rows))
    rows = [rows[i] for i in order]
    indptr = np.zeros(len(rows) + 1, dtype=np.int64)
    indptr[1:] = np.cumsum([len(r) for r in rows])
    indices = np.concatenate(rows).astype(np.int32) if rows else np.zeros(0, np.int32)
    return indptr, indices


def brute(indptr, indices, n_cols, d):
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import connected_components
    n = len(indptr) - 1
    X = csr_matrix((np.ones(len(indices), np.int32), indices, indptr), shape=(n, n_cols))
    card = np.asarray(X.sum(axis=1)).ravel()
    inter = (X @ X.T).toarray()
    D = card[:, None] + card[None, :] - 2 * inter
    adj = D <= d
    ncomp, lab = connected_components(csr_matrix(adj), directed=False)
    mins = np.full(ncomp, n, dtype=np.int64)
    np.minimum.at(mins, lab, np.arange(n))
    iu = np.triu_indices(n, 1)
    e = adj[iu]
    edges = set(zip(iu[0][e].tolist(), iu[1][e].tolist()))
    return mins[lab].astype(np.int32), edges, D


def cluster(indptr, indices, n_cols, d, engine):
    n = len(indptr) - 1
    labels = np.empty(n, dtype=np.int32)
    st = Stats()
    ck(lib.bf_cluster_csr(p64(indptr), p32(indices), C.c_int64(n), C.c_int32(n_cols), C.c_int32(d), 0, engine,
                          p32(labels), C.byref(st)))
    return labels, st


def neighbours(indptr, indices, n_cols, d, engine, query=None):
    n = len(indptr) - 1
    h = C.c_void_p()
    ne = C.c_int64()
    st = Stats()
    q = None if query is None else p32(query)
    nq = 0 if query is None else len(query)
    ck(lib.bf_neighbours_csr(p64(indptr), p32(indices), C.c_int64(n), C.c_int32(n_cols), q, C.c_int64(nq),
                             C.c_int32(d), 0, engine, C.byref(h), C.byref(ne), C.byref(st)))
    src = np.empty(ne.value, np.int32)
    dst = np.empty(ne.value, np.int32)
    ck(lib.bf_edges_copy(h, p32(src), p32(dst)))
    lib.bf_edges_free(h)
    return set(zip(src.tolist(), dst.tolist())), st


def main():
    out = {}
    n = C.c_int()
    ck(lib.bf_device_count(C.byref(n)))
    print("devices:", n.value)
    for name in ("popc32", "lop3", "iadd3", "xor_popc_add"):
        g = C.c_double()
        ck(lib.bf_measure_peak(0, name.encode(), C.byref(g)))
        out[f"peak_{name}_gops"] = g.value
        print(f"peak {name}: {g.value:.1f} Gop/s")

    rng = np.random.default_rng(7)
    ok = True
    for (nn, nc) in ((1, 50), (130, 300), (3000, 5000)):
        indptr, indices = gen(nn, nc, min(nn, 12), 12, rng)
        for d in (1, 2, 3):
            ref_lab, ref_edges, _ = brute(indptr, indices, nc, d)
            for engine in (0, 1):
                lab, st = cluster(indptr, indices, nc, d, engine)
                good = np.array_equal(lab, ref_lab)
                edges, st2 = neighbours(indptr, indices, nc, d, engine)
                good_e = edges == ref_edges
                print(f"N={nn} d={d} engine={engine}: labels {'OK' if good else 'MISMATCH'} edges {'OK' if good_e else 'MISMATCH'} "
                      f"(edges={len(edges)} ref={len(ref_edges)} cand={st.n_candidates} comps={st.n_components} band={st.pairs_band})")
                ok &= good and good_e
            # rectangle: a third of the rows as queries
            q = np.sort(rng.choice(nn, size=max(1, nn // 3), replace=False)).astype(np.int32)
            qs = set(q.tolist())
            want = {e for e in ref_edges if e[0] in qs or e[1] in qs}
            got, st3 = neighbours(indptr, indices, nc, d, 0, q)
            print(f"N={nn} d={d} rectangle: {'OK' if got == want else 'MISMATCH'} (edges={len(got)} want={len(want)} band={st3.pairs_band})")
            ok &= got == want
    out["parity_ok"] = bool(ok)

    big = 100_000
    if "--big" in sys.argv:
        big = int(sys.argv[sys.argv.index("--big") + 1])
    t0 = time.time()
    indptr, indices = gen(big, 88_000, 8, 40, np.random.default_rng(2))
    print(f"generated {big} rows, nnz={len(indices)} in {time.time() - t0:.1f}s")
    ctx = C.c_void_p()
    runs = []
    for engine, bits in ((0, 128), (0, 256), (0, 512), (0, 1024), (1, 0)):
        ck(lib.bf_ctx_create(0, None, C.byref(ctx)))
        ck(lib.bf_ctx_set_option(ctx, b"engine", C.c_int64(engine)))
        if bits:
            ck(lib.bf_ctx_set_option(ctx, b"sketch_bits", C.c_int64(bits)))
        ck(lib.bf_upload_csr(ctx, p64(indptr), p32(indices), C.c_int64(big), C.c_int32(88_000), None, C.c_int64(0)))
        st = Stats()
        labs = []
        for rep in range(3):
            ck(lib.bf_run(ctx, 1, 0, 1))
            ck(lib.bf_sync(ctx, C.byref(st)))
        lab = np.empty(big, np.int32)
        ck(lib.bf_download_labels(ctx, p32(lab)))
        labs.append(lab)
        d = st.as_dict()
        d["engine"] = engine
        d["cand_pairs_per_s"] = d["pairs_band"] / (d["ms_pairs"] * 1e-3) if d["ms_pairs"] > 0 else None
        d["popc_gops"] = d["pairs_evaluated"] * d["bits_per_row"] / 32 / (d["ms_pairs"] * 1e-3) / 1e9 if d["ms_pairs"] > 0 else None
        runs.append(d)
        print(json.dumps(d))
        lib.bf_ctx_destroy(ctx)
        if len(runs) > 1:
            same = np.array_equal(lab, first_lab)
            print("labels equal to first config:", same)
            ok &= same
        else:
            first_lab = lab
    out["runs"] = runs
    out["all_ok"] = bool(ok)
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "first_light.json").write_text(json.dumps(out, indent=1))
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
