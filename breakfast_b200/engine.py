"""Host-side glue between the reference-shaped Python interface and the native CUDA library.

Nothing in here computes distances or components: it tokenises profiles into a strictly binary CSR
(thermometer-coding repeated tokens so that set symmetric difference equals the reference's L1
distance on token counts, breakfast.py:210-212 + sklearn manhattan) and hands it to
libbreakfast_b200.so.  If the library or the GPU is missing the calls raise — there is no fallback.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from . import _native


def default_device() -> int:
    return int(os.environ.get("BREAKFAST_B200_DEVICE", "0"))


def default_engine() -> str:
    return os.environ.get("BREAKFAST_B200_ENGINE", "sketch")


def default_devices() -> list:
    """Devices for one job: BREAKFAST_B200_DEVICES="0,1,2,3" (or --gpus N on the CLI); default = one device."""
    spec = os.environ.get("BREAKFAST_B200_DEVICES", "")
    if spec.strip():
        return [int(x) for x in spec.split(",") if x.strip() != ""]
    return [default_device()]


def _context_options() -> dict:
    opts = {}
    if "BREAKFAST_B200_SKETCH_BITS" in os.environ:
        opts["sketch_bits"] = int(os.environ["BREAKFAST_B200_SKETCH_BITS"])
    if "BREAKFAST_B200_MERGE_CAPACITY" in os.environ:   # entries per rank of the multi-GPU label exchange (tests)
        opts["merge_capacity"] = int(os.environ["BREAKFAST_B200_MERGE_CAPACITY"])
    if "BREAKFAST_B200_SHARD_PACK_FROM" in os.environ:   # rank count from which the sketch pass is sharded (default: never on one box)
        opts["shard_pack_from"] = int(os.environ["BREAKFAST_B200_SHARD_PACK_FROM"])
    return opts


def tokenise(features, sep: str):
    """Token ids per profile, vocabulary by first appearance, duplicates kept, empty tokens skipped —
    exactly the traversal of reference sparse_feature_matrix (breakfast.py:199-213).

    Returns (indptr int64 [n+1], indices int64 [n_tokens], vocabulary dict token -> id).
    """
    vocab: dict = {}
    lookup = vocab.setdefault
    flat: list = []
    extend = flat.extend
    ends = np.empty(len(features) + 1, dtype=np.int64)
    ends[0] = 0
    for r, text in enumerate(features):
        if not isinstance(text, float) and text:  # NaN profiles and "" have no tokens
            extend([lookup(tok, len(vocab)) for tok in text.split(sep) if tok])
        ends[r + 1] = len(flat)
    return ends, np.asarray(flat, dtype=np.int64), vocab


def thermometer_binarise(indptr: np.ndarray, indices: np.ndarray, n_vocab: int):
    """Binary CSR with sorted unique int32 columns whose set distance equals L1 on token counts.

    The k-th repeat (k >= 1) of token t inside one profile becomes the extra feature (t, k); then
    |A xor B| = sum_t |count_A(t) - count_B(t)|, which is what the reference computes because scipy
    sums duplicate CSR entries before sklearn's manhattan kernel sees them.
    """
    n = len(indptr) - 1
    if indices.size == 0:
        return indptr.astype(np.int64), np.zeros(0, np.int32), int(n_vocab)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    order = np.lexsort((indices, rows))
    rows_s, cols_s = rows[order], indices[order]
    same = np.zeros(cols_s.size, dtype=bool)
    same[1:] = (rows_s[1:] == rows_s[:-1]) & (cols_s[1:] == cols_s[:-1])
    if not same.any():
        return indptr.astype(np.int64), cols_s.astype(np.int32), int(n_vocab)
    # rank of every entry inside its run of equal (row, col)
    run_start = np.flatnonzero(~same)
    run_id = np.cumsum(~same) - 1
    rank = np.arange(cols_s.size) - run_start[run_id]
    extra = rank > 0
    keys = cols_s[extra] * (rank.max() + 1) + rank[extra]
    uniq, inv = np.unique(keys, return_inverse=True)
    cols_b = cols_s.copy()
    cols_b[extra] = n_vocab + inv
    order2 = np.lexsort((cols_b, rows_s))
    return indptr.astype(np.int64), cols_b[order2].astype(np.int32), int(n_vocab + uniq.size)


def binary_csr(features, sep: str):
    indptr, indices, vocab = tokenise(features, sep)
    return thermometer_binarise(indptr, indices, len(vocab))


# seconds the last single-device engine call spent in each step (context, H2D, pass, D2H, teardown) - for benchmarks
LAST_TIMINGS: dict = {}


@dataclass
class ClusterResult:
    labels: np.ndarray                       # int32 [n]: smallest row index of the row's component
    stats: dict = field(default_factory=dict)
    edges: tuple | None = None               # (src, dst) int32, src < dst, when requested


def _components_multi_device(indptr, indices, n_cols, max_dist, devices, want_edges, engine, query_rows=None,
                             lists=None) -> ClusterResult:
    """Single-process multi-GPU: one context and one host thread per device, the band work items dealt cyclically
    (rank r of len(devices)).  On distinct devices the contexts share the library's NCCL communicator
    (bf_comm_init_all): the sketch pass is sharded and the ranks' union-finds are merged over NVLink inside bf_run, so
    every rank ends with the final labels and nothing but the first rank's labels crosses the host link.  Repeated
    device ids (several ranks emulated on one GPU, tests) cannot form a communicator: their labels are merged through
    the host (bf_merge_labels_host)."""
    import threading
    world = len(devices)
    use_comm = len(set(devices)) == world
    stats, labels, edges, errors = [None] * world, [None] * world, [None] * world, []
    ctxs = []
    try:
        for r in range(world):
            ctxs.append(_native.Context(device=devices[r], engine=engine, want_edges=int(bool(want_edges)), **_context_options()))
        if use_comm:
            _native.comm_init_all(ctxs)
        barrier = threading.Barrier(world)
        overflowed = [False] * world

        def work(r):
            try:
                ctx = ctxs[r]
                ctx.upload_csr(indptr, indices, n_cols, query_rows=query_rows)
                if not use_comm:
                    stats[r] = ctx.run_sync(max_dist, rank=r, world=world)
                    labels[r] = ctx.download_labels()
                else:
                    # a pass is collective: if any rank's bounded buffers overflowed (bf_sync raised them), all rerun
                    for _ in range(6):
                        ctx.run(max_dist, r, world)
                        try:
                            stats[r], overflowed[r] = ctx.sync(), False
                        except _native.NativeError as exc:
                            if exc.code != _native.BF_ERR_OVERFLOW:
                                raise
                            overflowed[r] = True
                        barrier.wait()
                        again = any(overflowed)
                        barrier.wait()
                        if not again:
                            break
                    else:
                        raise _native.NativeError(_native.BF_ERR_OVERFLOW, "buffer overflow persisted on some device")
                if want_edges:
                    edges[r] = ctx.download_edges()
            except BaseException as exc:  # re-raised in the caller's thread
                errors.append(exc)
                barrier.abort()

        threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            real = [e for e in errors if not isinstance(e, threading.BrokenBarrierError)]
            raise (real or errors)[0]
        ctx0 = ctxs[0]
        if not use_comm:
            ctx0.merge_labels_host(np.stack(labels))
        if lists is not None and len(lists[0]) > 1:
            ctx0.union_lists(*lists)
        st = ctx0.sync().as_dict()
        st["n_edges"] = int(sum(s.n_edges for s in stats))
        st["n_candidates"] = int(sum(s.n_candidates for s in stats))
        st["n_gpus"] = world
        st["label_merge"] = "nccl (in-library, over NVLink)" if use_comm else "host"
        out_labels = ctx0.download_labels()
        out_edges = None
        if want_edges:
            src = np.concatenate([e[0] for e in edges])
            dst = np.concatenate([e[1] for e in edges])
            order = np.lexsort((dst, src))
            out_edges = (src[order], dst[order])
        return ClusterResult(out_labels, st, out_edges)
    finally:
        for ctx in ctxs:
            ctx.close()


def components_full(indptr, indices, n_cols, max_dist, want_edges=False, device=None, engine=None) -> ClusterResult:
    """All-pairs run: radius-neighbour graph at max_dist + connected components (one GPU, or all of
    BREAKFAST_B200_DEVICES with the tile space partitioned over them)."""
    engine = default_engine() if engine is None else engine
    _native.require_device()
    if device is None and len(default_devices()) > 1:
        return _components_multi_device(indptr, indices, n_cols, max_dist, default_devices(), want_edges, engine)
    device = default_device() if device is None else device
    import time
    t0 = time.perf_counter()
    with _native.Context(device=device, engine=engine, want_edges=int(bool(want_edges)), **_context_options()) as ctx:
        t1 = time.perf_counter()
        ctx.upload_csr(indptr, indices, n_cols)
        t2 = time.perf_counter()
        st = ctx.run_sync(max_dist)
        t3 = time.perf_counter()
        labels = ctx.download_labels()
        edges = ctx.download_edges() if want_edges else None
        t4 = time.perf_counter()
    LAST_TIMINGS.clear()
    LAST_TIMINGS.update(context=t1 - t0, upload=t2 - t1, run=t3 - t2, download=t4 - t3, teardown=time.perf_counter() - t4)
    return ClusterResult(labels, st.as_dict(), edges)


def components_incremental(indptr, indices, n_cols, max_dist, new_rows, list_indptr, list_members,
                           want_edges=False, device=None, engine=None) -> ClusterResult:
    """Incremental run: only the new x all block is evaluated (reference: X = new rows, Y = all rows,
    breakfast.py:236-254) and the cached neighbour lists are united on top (breakfast.py:304)."""
    engine = default_engine() if engine is None else engine
    _native.require_device()
    new_rows = np.unique(np.asarray(new_rows, dtype=np.int32))
    if device is None and len(default_devices()) > 1:
        return _components_multi_device(indptr, indices, n_cols, max_dist, default_devices(), want_edges, engine,
                                        query_rows=new_rows, lists=(list_indptr, list_members))
    device = default_device() if device is None else device
    with _native.Context(device=device, engine=engine, want_edges=int(bool(want_edges)), **_context_options()) as ctx:
        ctx.upload_csr(indptr, indices, n_cols, query_rows=new_rows)
        st = ctx.run_sync(max_dist)
        if len(list_indptr) > 1:
            ctx.union_lists(list_indptr, list_members)
            st = ctx.sync()
        labels = ctx.download_labels()
        edges = ctx.download_edges() if want_edges else None
    return ClusterResult(labels, st.as_dict(), edges)


def adjacency_lists(n_rows: int, rows, src, dst):
    """Neighbour lists in the shape the reference caches them (list of ascending int64 arrays, each
    containing the query row itself, breakfast.py:226-228,274-275): one full list per row of
    sorted(unique(rows))."""
    rows = np.unique(np.asarray(rows, dtype=np.int64))
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    a = np.concatenate([src, dst, rows])
    b = np.concatenate([dst, src, rows])
    wanted = np.zeros(n_rows, dtype=bool)
    wanted[rows] = True
    keep = wanted[a]
    a, b = a[keep], b[keep]
    order = np.lexsort((b, a))
    a, b = a[order], b[order]
    counts = np.bincount(a, minlength=n_rows)[rows]
    starts = np.concatenate(([0], np.cumsum(counts)))
    # `rows` is ascending (np.unique), so the groups come out in the same order
    return [b[starts[i]:starts[i + 1]] for i in range(len(rows))]
