"""Parity tests proper: the CUDA path, called through the C ABI (and through the CLI on top of it),
against the CPU oracle and the committed golden files.  Bit-exact — the work is integer/indexing.
Run on the B200 box:  python -m pytest tests -m gpu
"""
import ctypes as C
import shutil

import numpy as np
import pytest

import oracle
from breakfast_b200 import _native, synth
from tests import helpers
from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu

ENGINES = ["sketch", "full"]


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    _native.require_device()  # fail loudly (do not skip): the GPU suite must run on the GPU


def rows_to_csr(rows, n_cols=None):
    indptr = np.zeros(len(rows) + 1, dtype=np.int64)
    indptr[1:] = np.cumsum([len(r) for r in rows])
    indices = np.concatenate([np.asarray(sorted(r), dtype=np.int32) for r in rows]) if indptr[-1] else np.zeros(0, np.int32)
    if n_cols is None:
        n_cols = int(indices.max()) + 1 if indices.size else 1
    return indptr, indices.astype(np.int32), n_cols


def edge_set(src, dst):
    return set(zip(np.asarray(src).tolist(), np.asarray(dst).tolist()))


# ------------------------------------------------------------------ goldens through the CLI
@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("case", helpers.cases("plain"), ids=lambda c: c["name"])
def test_cli_goldens(case, engine, tmp_path, monkeypatch):
    monkeypatch.setenv("BREAKFAST_B200_ENGINE", engine)
    helpers.assert_matches(case, case["expected"], helpers.run_cli(case["input"], case["opts"], tmp_path))


@pytest.mark.parametrize("case", [c for c in helpers.cases("cached") if "cache_from" in c], ids=lambda c: c["name"])
def test_cli_cached_goldens(case, tmp_path):
    first = case["cache_from"]
    cache = tmp_path / "cachedir" / "cache.pkl"
    helpers.run_cli(first["input"], first["opts"], tmp_path / "first", cache_out=cache)
    helpers.assert_matches(case, case["expected"],
                           helpers.run_cli(case["input"], case["opts"], tmp_path / "second", cache_in=cache))


@pytest.mark.parametrize("case", [c for c in helpers.cases("cached") if "reference_cache_file" in c],
                         ids=lambda c: c["name"])
def test_cli_reference_written_cache(case, tmp_path):
    cache = tmp_path / "ref.cache"
    shutil.copyfile(GOLDEN / case["reference_cache_file"], cache)
    helpers.assert_matches(case, case["expected"],
                           helpers.run_cli(case["input"], case["opts"], tmp_path / "out", cache_in=cache))


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("case", helpers.cases("chain"), ids=lambda c: c["name"])
def test_cli_cache_chain(case, engine, tmp_path, monkeypatch):
    monkeypatch.setenv("BREAKFAST_B200_ENGINE", engine)
    prev = None
    for k, step in enumerate(case["chain"]):
        out_cache = tmp_path / f"cache{k}"
        text = helpers.run_cli(step["input"], case["opts"], tmp_path / f"out{k}", cache_in=prev, cache_out=out_cache)
        helpers.assert_matches(case, step["expected"], text)
        prev = out_cache


# ------------------------------------------------------------------ C ABI vs oracle on seeded inputs
@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("max_dist", [1, 2, 3])
@pytest.mark.parametrize("n", [1, 2, 127, 128, 129, 1000, 4097])
def test_cluster_and_edges_match_oracle(n, max_dist, engine):
    indptr, indices, n_cols = synth.generate(n, seed=100 + n).csr()
    want_labels, want_ne = oracle.cluster(indptr, indices, max_dist)
    labels, st = _native.cluster_csr(indptr, indices, n_cols, max_dist, engine=engine)
    assert np.array_equal(labels, want_labels)
    assert st.n_edges == want_ne and st.n_rows == n
    assert st.n_components == len(set(want_labels.tolist()))
    src, dst, st2 = _native.neighbours_csr(indptr, indices, n_cols, max_dist, engine=engine)
    ws, wd = oracle.edges(indptr, indices, max_dist)
    assert np.array_equal(src, ws) and np.array_equal(dst, wd)
    card = np.diff(indptr)
    h = np.bincount(card)
    band = sum(int(h[c]) * (int(h[c]) - 1) // 2 for c in range(len(h)))
    band += sum(int(h[c]) * int(h[c + k]) for k in range(1, max_dist + 1) for c in range(len(h) - k))
    assert st.pairs_band == band and st.pairs_total == n * (n - 1) // 2


@pytest.mark.parametrize("max_dist", [1, 2, 3, 5])
@pytest.mark.parametrize("level1", [0, 1, 2])
def test_both_level1_kernels_are_exact(level1, max_dist):
    """level 1 on the integer pipes (POPC / POPC-free hybrid) and on the tensor cores (int8 mma.sync; 2 = two column
    rows per accumulator, the default)"""
    indptr, indices, n_cols = synth.generate(6000, seed=91).csr()
    want, ne = oracle.cluster(indptr, indices, max_dist)
    with _native.Context(level1=level1, sketch_bits=128) as ctx:
        ctx.upload_csr(indptr, indices, n_cols)
        st = ctx.run_sync(max_dist)
        assert np.array_equal(ctx.download_labels(), want) and st.n_edges == ne


@pytest.mark.parametrize("max_dist", [8, 9, 15, 16, 31, 32, 40])
@pytest.mark.parametrize("level1", [0, 1, 2])
@pytest.mark.parametrize("bits", [128, 256])
def test_large_max_dist_where_the_32_bit_level_cannot_reject(bits, level1, max_dist):
    """from max_dist 16 on the 32-bit level-1 test passes every pair (threshold 32 - 2 d <= 0) and the work falls to
    level 2 and the verification (capacities are raised and the pass rerun where the queues overflow); 8 / 9 straddle
    the switch from the three-key to the two-key schedule.  Still exact."""
    indptr, indices, n_cols = synth.generate(1500, seed=17).csr()
    want, ne = oracle.cluster(indptr, indices, max_dist)
    with _native.Context(level1=level1, sketch_bits=bits) as ctx:
        ctx.upload_csr(indptr, indices, n_cols)
        st = ctx.run_sync(max_dist)
        assert np.array_equal(ctx.download_labels(), want) and st.n_edges == ne


@pytest.mark.parametrize("ctas", [1, 2])
@pytest.mark.parametrize("max_dist", [1, 2, 4])
def test_packed_level1_matches_the_plain_tensor_core_level1(ctas, max_dist):
    """k_pairs_l1_imma2 (two column rows per accumulator, one or two CTAs per SM) against k_pairs_l1_imma: same candidates
    after level 2 (both levels are exact 32-bit fold tests, level 1 may only add the complement-fold alias), same edges,
    same labels; rectangle run included (query rows = every third row)."""
    indptr, indices, n_cols = synth.generate(20000, seed=5).csr()
    q = np.arange(0, 20000, 3, dtype=np.int32)
    res = []
    for level1, opts in ((1, {}), (2, {"l1_ctas": ctas})):
        with _native.Context(level1=level1, sketch_bits=128, want_edges=1, **opts) as ctx:
            ctx.upload_csr(indptr, indices, n_cols)
            st = ctx.run_sync(max_dist)
            labels, edges = ctx.download_labels(), ctx.download_edges()
            ctx.upload_csr(indptr, indices, n_cols, query_rows=q)
            st_q = ctx.run_sync(max_dist)
            res.append((labels, sorted(zip(*edges[:2])), st.n_candidates, st.n_edges, st_q.n_candidates, st_q.n_edges))
    assert np.array_equal(res[0][0], res[1][0])
    assert res[0][1:] == res[1][1:]
    want, ne = oracle.cluster(indptr, indices, max_dist)
    assert np.array_equal(res[1][0], want) and res[1][3] == ne


@pytest.mark.parametrize("bits", [128, 256, 512, 1024, 2048])
def test_every_sketch_width_is_exact(bits):
    indptr, indices, n_cols = synth.generate(3000, seed=77).csr()
    want, _ = oracle.cluster(indptr, indices, 2)
    with _native.Context(engine="sketch", sketch_bits=bits) as ctx:
        ctx.upload_csr(indptr, indices, n_cols)
        st = ctx.run_sync(2)
        assert st.bits_per_row == bits
        assert np.array_equal(ctx.download_labels(), want)


@pytest.mark.parametrize("bits", [128, 256, 512])
@pytest.mark.parametrize("max_dist", [1, 3])
def test_two_level_filter_changes_nothing(bits, max_dist):
    """the 32-bit first-level fold inside the pair kernel is a pure pre-filter: same candidates' edges"""
    indptr, indices, n_cols = synth.generate(5000, seed=31).csr()
    outs = []
    for two_level in (0, 1):
        with _native.Context(sketch_bits=bits, two_level=two_level, want_edges=1) as ctx:
            ctx.upload_csr(indptr, indices, n_cols)
            st = ctx.run_sync(max_dist)
            outs.append((ctx.download_labels(), *ctx.download_edges(), st.n_candidates))
            if two_level:
                assert st.l2_warp_items > 0
                # 128/256-bit sketches: separate level-1 kernel whose test is POPC for half of the pairs when
                # max_dist is 1 or 2; 512 bits: level 1 inside the single kernel, one POPC per pair
                # (default level 1 for 128/256 bits = int8 mma.sync: no POPC at all there)
                if bits <= 256:   # level 2: 32-bit test of the unit's 32 pairs, full width only for those that pass
                    assert 32 * st.l2_warp_items <= st.popc32_executed <= 32 * st.l2_warp_items * (1 + bits // 32)
                else:
                    assert st.popc32_executed == st.pairs_evaluated + st.l2_warp_items * 1024 * (bits // 32)
            else:
                assert st.l2_warp_items == 0 and st.popc32_executed == st.pairs_evaluated * (bits // 32)
    assert all(np.array_equal(a, b) for a, b in zip(outs[0][:3], outs[1][:3])) and outs[0][3] == outs[1][3]
    want, _ = oracle.cluster(indptr, indices, max_dist)
    assert np.array_equal(outs[1][0], want)


@pytest.mark.parametrize("engine", ENGINES)
def test_edge_cases(engine):
    # empty input
    labels, st = _native.cluster_csr(np.zeros(1, np.int64), np.zeros(0, np.int32), 5, 1, engine=engine)
    assert labels.size == 0 and st.n_edges == 0
    # only empty profiles, identical sets, a chain, an isolated row
    rows = [[], [], [3], [3, 4], [3, 4, 5], [3, 4, 5, 6], [10, 11, 12, 13, 14], [3, 4], [20], []]
    indptr, indices, n_cols = rows_to_csr(rows, 32)
    for d in (1, 2, 5):
        want, _ = oracle.cluster(indptr, indices, d)
        got, _ = _native.cluster_csr(indptr, indices, n_cols, d, engine=engine)
        assert np.array_equal(got, want), d
    # ragged: one very long row next to short ones, long rows one feature apart
    big = list(range(0, 9000, 2))
    rows = [big, big + [9001], [1], [1, 3], big[:-1], list(range(50))]
    indptr, indices, n_cols = rows_to_csr(rows, 9100)
    want, _ = oracle.cluster(indptr, indices, 1)
    got, _ = _native.cluster_csr(indptr, indices, n_cols, 1, engine=engine)
    assert np.array_equal(got, want) and want[1] == 0 and want[4] == 0


def test_cardinalities_beyond_the_sort_key_clamp():
    """rows with more than 65535 features share one (clamped) cardinality key; banding must stay sound"""
    base = list(range(0, 140000, 2))              # 70000 features
    rows = [base, base + [1], base[:-1], base[:65535], base[:65534], base[:65535] + [139999], [5]]
    indptr, indices, n_cols = rows_to_csr(rows, 140000)
    for engine in ENGINES:
        want, _ = oracle.cluster(indptr, indices, 1)
        got, _ = _native.cluster_csr(indptr, indices, n_cols, 1, engine=engine)
        assert np.array_equal(got, want)
    assert want.tolist() == [0, 0, 0, 3, 3, 3, 6]


@pytest.mark.parametrize("max_dist", [1, 2, 3])
def test_schedule_counter_matches_the_cpu_mirror(max_dist):
    """k_schedule lists exactly the tile pairs that tools/schedule_sim.py (its CPU restatement, whose soundness
    tests/test_schedule_math.py checks against the oracle) lists"""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import schedule_sim
    indptr, indices, n_cols = synth.generate(20000, seed=13).csr()
    with _native.Context(sketch_bits=128) as ctx:
        ctx.upload_csr(indptr, indices, n_cols)
        st = ctx.run_sync(max_dist)
    assert st.tiles_band == schedule_sim.tile_pairs(indptr, indices, max_dist, 3)
    assert st.tiles_band <= schedule_sim.tile_pairs(indptr, indices, max_dist, 2) < schedule_sim.tile_pairs(indptr, indices, max_dist, 1)


def _near_duplicate_rows(n, card, n_cols, seed, spread):
    """rows of (almost) equal cardinality, planted in families whose members differ by a few columns"""
    rng = np.random.default_rng(seed)
    rows = []
    while len(rows) < n:
        base = set(rng.choice(n_cols, size=card, replace=False).tolist())
        for _ in range(int(rng.integers(1, 6))):
            r = set(base)
            for _ in range(int(rng.integers(0, 3))):          # swap a column: distance 2, same cardinality
                r.remove(next(iter(r)))
                r.add(int(rng.integers(n_cols)))
            for _ in range(int(rng.integers(0, spread + 1))):  # add a column: distance 1, cardinality + 1
                r.add(int(rng.integers(n_cols)))
            rows.append(sorted(r))
    return rows[:n]


@pytest.mark.parametrize("max_dist", [1, 2, 3])
@pytest.mark.parametrize("spread", [0, 2])
@pytest.mark.parametrize("card", [40, 600])
def test_two_key_schedule_on_equal_cardinalities(max_dist, spread, card):
    """every tile holds rows of one cardinality (spread 0) or of a few: the (cardinality, half-cardinality) runs of
    k_schedule decide which tile pairs are evaluated at all — no edge may be lost, for the triangle and the rectangle.
    card 600: both key halves need their high byte (all four radix passes run) and the rows are longer than the
    128 columns the verify kernel handles in registers."""
    rows = _near_duplicate_rows(3000 if card < 100 else 1500, card, 5000, seed=7 + spread, spread=spread)
    indptr, indices, n_cols = rows_to_csr(rows, 5000)
    want, want_edges = oracle.cluster(indptr, indices, max_dist)
    for bits in (128, 256, 512):
        with _native.Context(sketch_bits=bits, want_edges=1) as ctx:
            ctx.upload_csr(indptr, indices, n_cols)
            st = ctx.run_sync(max_dist)
            assert np.array_equal(ctx.download_labels(), want)
            assert st.n_edges == want_edges
            if bits <= 256 and spread == 0:   # the runs prune: fewer tile pairs than the triangle of all tiles
                tiles = (len(rows) + 127) // 128
                assert st.pairs_evaluated < tiles * (tiles + 1) // 2 * 128 * 128
    query = np.arange(0, len(rows), 7, dtype=np.int32)
    src, dst, _ = _native.neighbours_csr(indptr, indices, n_cols, max_dist, query_rows=query)
    ws, wd = oracle.edges(indptr, indices, max_dist, queries=query)
    assert np.array_equal(src, ws) and np.array_equal(dst, wd)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("max_dist", [1, 2])
def test_incremental_rectangle_matches_oracle(max_dist, engine):
    indptr, indices, n_cols = synth.generate(3000, seed=5).csr()
    rng = np.random.default_rng(1)
    for nq in (0, 1, 40, 1500, 3000):
        q = np.sort(rng.choice(3000, size=nq, replace=False)).astype(np.int32)
        src, dst, st = _native.neighbours_csr(indptr, indices, n_cols, max_dist, query_rows=q, engine=engine)
        if nq == 0:
            assert src.size == 0
            continue
        ws, wd = oracle.edges(indptr, indices, max_dist, queries=q)
        assert np.array_equal(src, ws) and np.array_equal(dst, wd)
        assert st.n_query == nq


def test_components_api_matches_oracle():
    rng = np.random.default_rng(3)
    n = 5000
    src = rng.integers(0, n, 3000).astype(np.int32)
    dst = rng.integers(0, n, 3000).astype(np.int32)
    lens = rng.integers(0, 6, 400)
    li = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    lm = rng.integers(0, n, int(li[-1])).astype(np.int32)
    got, ncomp = _native.components(n, src, dst, li, lm)
    want = oracle.components(n, src, dst, li, lm)
    assert np.array_equal(got, want) and ncomp == len(set(want.tolist()))
    got2, _ = _native.components(7)
    assert got2.tolist() == list(range(7))


@pytest.mark.parametrize("world", [2, 3, 8])
def test_multi_rank_partition_and_merge(world):
    """every rank takes its cyclic share of the band tiles; merged labels == single-rank labels"""
    indptr, indices, n_cols = synth.generate(6000, seed=8).csr()
    single, st1 = _native.cluster_csr(indptr, indices, n_cols, 2)
    gathered = np.empty((world, 6000), dtype=np.int32)
    tiles = edges = 0
    ctxs = []
    for r in range(world):
        ctx = _native.Context(engine="sketch")
        ctx.upload_csr(indptr, indices, n_cols)
        st = ctx.run_sync(2, rank=r, world=world)
        gathered[r] = ctx.download_labels()
        tiles += st.tiles_rank
        edges += st.n_edges
        assert st.tiles_band == st1.tiles_band
        ctxs.append(ctx)
    assert tiles == st1.tiles_band and edges == st1.n_edges
    assert not np.array_equal(gathered[0], single)      # a single share is not the answer
    for ctx in ctxs:
        ctx.merge_labels_host(gathered)
        assert np.array_equal(ctx.download_labels(), single)
        assert ctx.sync().n_components == st1.n_components
        ctx.close()
    assert np.array_equal(helpers.merge_labels_cpu(gathered), single)


def test_single_process_multi_device_engine(monkeypatch, tmp_path):
    """the CLI's --gpus path: one host thread + context per device, cyclic work partition, merge on the first
    device.  With one GPU in the box the same device is listed several times, which exercises the whole path."""
    from breakfast_b200 import engine
    indptr, indices, n_cols = synth.generate(8000, seed=17).csr()
    want, ne = oracle.cluster(indptr, indices, 2)
    monkeypatch.setenv("BREAKFAST_B200_DEVICES", "0,0,0")
    res = engine.components_full(indptr, indices, n_cols, 2, want_edges=True)
    assert np.array_equal(res.labels, want) and res.stats["n_edges"] == ne and res.stats["n_gpus"] == 3
    ws, wd = oracle.edges(indptr, indices, 2)
    assert np.array_equal(res.edges[0], ws) and np.array_equal(res.edges[1], wd)
    # incremental + cached lists over several devices
    q = np.arange(0, 8000, 9, dtype=np.int32)
    li, lm = np.array([0, 2, 5], np.int64), np.array([1, 7000, 3, 4, 7999], np.int32)
    res = engine.components_incremental(indptr, indices, n_cols, 1, q, li, lm, want_edges=True)
    s1, d1 = oracle.edges(indptr, indices, 1, queries=q)
    assert np.array_equal(res.edges[0], s1) and np.array_equal(res.edges[1], d1)
    assert np.array_equal(res.labels, oracle.components(8000, s1, d1, li, lm))
    # and through the CLI option
    case = [c for c in helpers.cases("plain") if c["name"] == "syn_dna_d2_m5"][0]
    monkeypatch.delenv("BREAKFAST_B200_DEVICES")
    import click.testing
    from breakfast_b200 import console
    monkeypatch.setenv("BREAKFAST_B200_DEVICES", "0,0")
    r = click.testing.CliRunner().invoke(console.main, ["--input-file", str(GOLDEN / case["input"]), "--outdir", str(tmp_path),
                                                        "--gpus", "2"] + helpers.cli_args(case["opts"]))
    assert r.exit_code == 0, r.output
    helpers.assert_matches(case, case["expected"], (tmp_path / "clusters.tsv").read_text())


@pytest.mark.parametrize("shard_pack", [False, True], ids=["replicated-sketch-pass", "sharded-sketch-pass"])
@pytest.mark.parametrize("merge_capacity", [0, 64], ids=["auto", "tiny-exchange-buffer"])
def test_in_library_communicator_on_distinct_devices(merge_capacity, shard_pack, monkeypatch):
    """two (or four) real GPUs in one process: bf_comm_init_all gives the contexts the library's NCCL communicator, the
    union-finds are exchanged inside bf_run and - with shard_pack_from = 2 - the sketch pass is sharded and its shares
    all-gathered (the default replicates it on one box: measured faster); labels and edges equal the single-GPU answer,
    for the full run and for the incremental rectangle, with a deliberately small exchange capacity (overflow -> rerun)"""
    from breakfast_b200 import engine
    n_dev = _native.device_count()
    if n_dev < 2:
        pytest.skip("needs at least two GPUs")
    devices = list(range(4 if n_dev >= 4 else 2))
    if merge_capacity:
        monkeypatch.setenv("BREAKFAST_B200_MERGE_CAPACITY", str(merge_capacity))
    if shard_pack:
        monkeypatch.setenv("BREAKFAST_B200_SHARD_PACK_FROM", "2")
    indptr, indices, n_cols = synth.generate(30000, seed=17).csr()
    want_labels, _ = oracle.cluster(indptr, indices, 2)
    ws, wd = oracle.edges(indptr, indices, 2)
    res = engine._components_multi_device(indptr, indices, n_cols, 2, devices, True, "sketch")
    assert res.stats["label_merge"].startswith("nccl")
    assert np.array_equal(res.labels, want_labels)
    assert np.array_equal(res.edges[0], ws) and np.array_equal(res.edges[1], wd)
    q = np.arange(0, 30000, 7, dtype=np.int32)
    qs, qd = oracle.edges(indptr, indices, 1, queries=q)
    res = engine._components_multi_device(indptr, indices, n_cols, 1, devices, True, "sketch", query_rows=q)
    assert np.array_equal(res.edges[0], qs) and np.array_equal(res.edges[1], qd)
    assert np.array_equal(res.labels, oracle.components(30000, qs, qd))


def test_merge_on_device_with_torch_tensors():
    """the torchrun path: labels -> torch device tensor -> (all-gather) -> merge on device"""
    torch = pytest.importorskip("torch")
    assert torch.cuda.is_available()
    indptr, indices, n_cols = synth.generate(3000, seed=4).csr()
    single, _ = _native.cluster_csr(indptr, indices, n_cols, 1)
    tstream = torch.cuda.Stream()          # the library enqueues on this torch stream
    with torch.cuda.stream(tstream):
        parts = []
        ctxs = [_native.Context(stream=tstream.cuda_stream) for _ in range(2)]
        for r, ctx in enumerate(ctxs):
            ctx.upload_csr(indptr, indices, n_cols)
            ctx.run(1, r, 2)
            t = torch.empty(3000, dtype=torch.int32, device="cuda")
            ctx.labels_to_device(t.data_ptr())
            parts.append(t)
        gathered = torch.stack(parts).contiguous()
        for ctx in ctxs:
            ctx.merge_labels_device(gathered.data_ptr(), 2)
            ctx.sync()
            assert np.array_equal(ctx.download_labels(), single)
            ctx.close()


def test_candidate_overflow_is_reported_and_recovered():
    indptr, indices, n_cols = synth.generate(4000, seed=6).csr()
    want, _ = oracle.cluster(indptr, indices, 2)
    with _native.Context(cand_capacity=16) as ctx:
        ctx.upload_csr(indptr, indices, n_cols)
        ctx.run(2)
        with pytest.raises(_native.NativeError) as e:
            ctx.sync()
        assert e.value.code == _native.BF_ERR_OVERFLOW
        st = ctx.run_sync(2)                               # grows the buffer and reruns
        assert st.n_candidates > 16
        assert np.array_equal(ctx.download_labels(), want)


@pytest.mark.parametrize("option", ["units_capacity", "items_capacity"])
def test_queue_and_work_list_overflow_are_recovered(option):
    """the level-2 queue and the work list are bounded too: bf_sync reports the overflow, raises the capacity, and the
    rerun (which may have to grow the next buffer of the chain as well) gives the exact result"""
    indptr, indices, n_cols = synth.generate(6000, seed=8).csr()
    want, want_edges = oracle.cluster(indptr, indices, 2)
    with _native.Context(want_edges=1, cand_capacity=64, **{option: 8}) as ctx:
        ctx.upload_csr(indptr, indices, n_cols)
        ctx.run(2)
        with pytest.raises(_native.NativeError) as e:
            ctx.sync()
        assert e.value.code == _native.BF_ERR_OVERFLOW
        st = ctx.run_sync(2)
        assert st.n_edges == want_edges and np.array_equal(ctx.download_labels(), want)
    labels, st = _native.cluster_csr(indptr, indices, n_cols, 2)
    assert np.array_equal(labels, want)


def test_call_order_and_argument_errors():
    with _native.Context() as ctx:
        with pytest.raises(_native.NativeError) as e:
            ctx.run(1)
        assert e.value.code == _native.BF_ERR_STATE
        with pytest.raises(_native.NativeError):
            ctx.set_option("sketch_bits", 100)
        with pytest.raises(_native.NativeError):
            ctx.upload_csr(np.array([0, 2, 1], np.int64), np.zeros(2, np.int32), 4)
        indptr, indices, n_cols = synth.generate(10, seed=1).csr()
        ctx.upload_csr(indptr, indices, n_cols)
        with pytest.raises(_native.NativeError):
            ctx.run(1, rank=2, world=2)


def test_determinism_and_pinned_memory():
    indptr, indices, n_cols = synth.generate(20000, seed=12).csr()
    lib = _native.load()
    p1, p2 = C.c_void_p(), C.c_void_p()
    assert lib.bf_pinned_alloc(indptr.nbytes, C.byref(p1)) == 0 and lib.bf_pinned_alloc(indices.nbytes, C.byref(p2)) == 0
    C.memmove(p1, indptr.ctypes.data, indptr.nbytes)
    C.memmove(p2, indices.ctypes.data, indices.nbytes)
    outs = []
    with _native.Context(want_edges=1) as ctx:
        ctx.upload_csr_ptr(p1.value, p2.value, len(indptr) - 1, n_cols)
        for _ in range(3):
            ctx.run_sync(1)
            outs.append((ctx.download_labels(), *ctx.download_edges()))
    lib.bf_pinned_free(p1)
    lib.bf_pinned_free(p2)
    for o in outs[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(o, outs[0]))
    want, _ = oracle.cluster(indptr, indices, 1)
    assert np.array_equal(outs[0][0], want)


def test_async_double_buffered_upload_and_adopted_device_csr():
    """bf_upload_csr_async alternates two device slots (different batches back to back give each its
    own answer) and bf_adopt_csr_device runs on caller-owned device memory."""
    torch = pytest.importorskip("torch")
    batches = [synth.generate(n, seed=s).csr() for n, s in ((9000, 1), (7000, 2), (9000, 3))]
    wants = [oracle.cluster(ip, ix, 1)[0] for ip, ix, _ in batches]
    lib = _native.load()
    pinned = []
    for ip, ix, _ in batches:
        a, b = C.c_void_p(), C.c_void_p()
        assert lib.bf_pinned_alloc(ip.nbytes, C.byref(a)) == 0 and lib.bf_pinned_alloc(ix.nbytes, C.byref(b)) == 0
        C.memmove(a, ip.ctypes.data, ip.nbytes)
        C.memmove(b, ix.ctypes.data, ix.nbytes)
        pinned.append((a, b))
    with _native.Context() as ctx:
        ctx.upload_csr_async_ptr(pinned[0][0].value, pinned[0][1].value, len(batches[0][0]) - 1, batches[0][2])
        for k in range(6):
            cur = k % 3
            ctx.run(1)
            nxt = (k + 1) % 3
            ctx.upload_csr_async_ptr(pinned[nxt][0].value, pinned[nxt][1].value, len(batches[nxt][0]) - 1, batches[nxt][2])
            ctx.n_rows = len(batches[cur][0]) - 1
            assert np.array_equal(ctx.download_labels(), wants[cur]), k
        with pytest.raises(_native.NativeError):        # one pending upload at a time
            ctx.upload_csr_async_ptr(pinned[0][0].value, pinned[0][1].value, len(batches[0][0]) - 1, batches[0][2])
        ctx.run(1)
        ctx.sync()
    for a, b in pinned:
        lib.bf_pinned_free(a)
        lib.bf_pinned_free(b)
    ip, ix, nc = batches[1]
    ts = torch.cuda.Stream()
    with torch.cuda.stream(ts):
        d_ip, d_ix = torch.from_numpy(ip).cuda(), torch.from_numpy(ix).cuda()
        with _native.Context(stream=ts.cuda_stream) as ctx:
            ctx.adopt_csr_device(d_ip.data_ptr(), d_ix.data_ptr(), len(ip) - 1, nc, ix.size)
            ctx.run_sync(1)
            assert np.array_equal(ctx.download_labels(), wants[1])


# ------------------------------------------------------------------ hash-join engine (max_dist <= 2)
@pytest.mark.parametrize("max_dist", [0, 1, 2])
@pytest.mark.parametrize("n", [1, 2, 129, 1000, 4097, 30000])
def test_hashjoin_engine_matches_oracle(n, max_dist):
    """the third engine (equi-joins on an additive row hash + exact verification): labels, edges and counters of oracle.c"""
    indptr, indices, n_cols = synth.generate(n, seed=300 + n).csr()
    want_labels, want_ne = oracle.cluster(indptr, indices, max_dist)
    labels, st = _native.cluster_csr(indptr, indices, n_cols, max_dist, engine="hashjoin")
    assert np.array_equal(labels, want_labels) and st.n_edges == want_ne
    src, dst, st2 = _native.neighbours_csr(indptr, indices, n_cols, max_dist, engine="hashjoin")
    ws, wd = oracle.edges(indptr, indices, max_dist)
    assert np.array_equal(src, ws) and np.array_equal(dst, wd)
    _, st_sketch = _native.cluster_csr(indptr, indices, n_cols, max_dist, engine="sketch")
    assert st.pairs_band == st_sketch.pairs_band and st.pairs_total == st_sketch.pairs_total


def test_hashjoin_engine_identical_rows_rectangle_ranks_and_limits():
    rows = [{1, 2, 3}, {1, 2, 3}, {1, 2}, {1, 2, 3, 9}, {5, 6, 7, 8}, {5, 6, 7, 9}, {5, 6}, set(), set(), {70000, 5, 6}]
    indptr, indices, n_cols = rows_to_csr(rows, 70001)
    for d in (0, 1, 2):
        want, _ = oracle.cluster(indptr, indices, d)
        got, _ = _native.cluster_csr(indptr, indices, n_cols, d, engine="hashjoin")
        assert np.array_equal(got, want), d
        ws, wd = oracle.edges(indptr, indices, d)
        src, dst, _ = _native.neighbours_csr(indptr, indices, n_cols, d, engine="hashjoin")
        assert np.array_equal(src, ws) and np.array_equal(dst, wd), d
    with pytest.raises(_native.NativeError):
        _native.cluster_csr(indptr, indices, n_cols, 3, engine="hashjoin")
    # rectangle (incremental path): pairs with at least one endpoint among the query rows
    indptr, indices, n_cols = synth.generate(20000, seed=31).csr()
    q = np.arange(3, 20000, 11, dtype=np.int32)
    for d in (1, 2):
        ws, wd = oracle.edges(indptr, indices, d, queries=q)
        src, dst, _ = _native.neighbours_csr(indptr, indices, n_cols, d, query_rows=q, engine="hashjoin")
        assert np.array_equal(src, ws) and np.array_equal(dst, wd)
    # three ranks (emulated one after the other): the probes are dealt by row, the merged labels are the single-rank ones
    want, _ = oracle.cluster(indptr, indices, 2)
    gathered = np.empty((3, 20000), dtype=np.int32)
    edges = 0
    with _native.Context(engine="hashjoin") as ctx:
        ctx.upload_csr(indptr, indices, n_cols)
        for r in range(3):
            edges += ctx.run_sync(2, rank=r, world=3).n_edges
            gathered[r] = ctx.download_labels()
        ctx.merge_labels_host(gathered)
        assert np.array_equal(ctx.download_labels(), want)
    assert edges == oracle.edges(indptr, indices, 2)[0].size


@pytest.mark.parametrize("case", [c for c in helpers.cases("plain") if c["opts"].get("max_dist", 1) in (1, 2)], ids=lambda c: c["name"])
def test_cli_goldens_with_the_hashjoin_engine(case, tmp_path, monkeypatch):
    monkeypatch.setenv("BREAKFAST_B200_ENGINE", "hashjoin")
    helpers.assert_matches(case, case["expected"], helpers.run_cli(case["input"], case["opts"], tmp_path))


def _pinned_copy(lib, arr):
    q = C.c_void_p()
    assert lib.bf_pinned_alloc(max(arr.nbytes, 1), C.byref(q)) == 0
    C.memmove(q, arr.ctypes.data, arr.nbytes)
    return q


@pytest.mark.parametrize("wide", [False, True], ids=["cols<=65536", "cols>65536"])
def test_compact_csr16_upload_gives_the_plain_answer(wide):
    """bf_csr16_encode + bf_upload_csr16_async (16-bit columns + per-row split, decoded on the device) alternating with
    the plain async upload on the two slots: every batch gets the oracle's labels and edges"""
    lib = _native.load()
    batches = []
    for n, seed in ((9000, 21), (12000, 22)):
        ip, ix, nc = synth.generate(n, seed=seed).csr()
        if wide:   # spread the columns over 0 .. 131071 (order kept), so that most rows straddle the 65536 boundary
            nc2 = 131072
            ix = (ix.astype(np.int64) * (nc2 - 1) // max(nc - 1, 1)).astype(np.int32)
            nc = nc2
        batches.append((ip, ix, nc))
    if wide:
        assert any((batches[0][1] < 65536).any() and (batches[0][1] >= 65536).any() for _ in [0])
    wants = [oracle.cluster(ip, ix, 1)[0] for ip, ix, _ in batches]
    pinned = []
    for ip, ix, nc in batches:
        ip32, split, lo = _native.csr16_encode(ip, ix, nc)
        assert (split is not None) == wide
        pinned.append((_pinned_copy(lib, ip32), _pinned_copy(lib, split) if split is not None else None, _pinned_copy(lib, lo),
                       _pinned_copy(lib, ip), _pinned_copy(lib, ix)))
    with _native.Context(want_edges=1) as ctx:
        for k in range(5):
            cur = k % 2
            ip, ix, nc = batches[cur]
            q = pinned[cur]
            if k == 3:     # plain form in between: the slots alternate whatever the form
                ctx.upload_csr_async_ptr(q[3].value, q[4].value, len(ip) - 1, nc)
            else:
                ctx.upload_csr16_async_ptr(q[0].value, q[1].value if q[1] is not None else None, q[2].value, len(ip) - 1, nc)
            ctx.run(1)
            ctx.sync()
            assert np.array_equal(ctx.download_labels(), wants[cur]), k
        src, dst = ctx.download_edges()
        ws, wd = oracle.edges(batches[0][0], batches[0][1], 1)
        assert np.array_equal(src, ws) and np.array_equal(dst, wd)
    for q in pinned:
        for x in q:
            if x is not None:
                lib.bf_pinned_free(x)


def _run_ctx(indptr, indices, n_cols, max_dist, query_rows=None, **options):
    with _native.Context(want_edges=1, **options) as ctx:
        ctx.upload_csr(indptr, indices, n_cols, query_rows=query_rows)
        st = ctx.run_sync(max_dist)
        labels = ctx.download_labels()
        src, dst = ctx.download_edges()
    return labels, src, dst, st


@pytest.mark.parametrize("shape", ["cols<=65536", "cols>65536", "cols>131072", "long_rows", "ragged"])
@pytest.mark.parametrize("max_dist", [1, 2, 3, 5])
def test_resident_compact_form_changes_nothing(shape, max_dist):
    """bf_upload_csr derives the compact resident form (k_csr16_encode) and the sketch pass (k_pack_sketch_rows16) and,
    with option verify_csr16 = 1, the verification (k_verify_unite<., RowStore16>) then read it: same labels and edges
    as with the plain CSR (option resident_csr16 = 0) and as the oracle - also where the matrix is not representable
    and the plain form has to stay in charge"""
    rng = np.random.default_rng(7 + max_dist)
    if shape in ("cols<=65536", "cols>65536", "cols>131072"):
        ip, ix, nc = synth.generate(6000, seed=31).csr()
        target = {"cols<=65536": 65536, "cols>65536": 131072, "cols>131072": 400000}[shape]
        ix = (ix.astype(np.int64) * (target - 1) // max(nc - 1, 1)).astype(np.int32)
        nc = target
    elif shape == "long_rows":
        # 700 rows of about 600 columns: a block of 128 rows spans several staging chunks, rows span several
        # verification chunks; clusters of near-identical rows
        base = [np.sort(rng.choice(90000, size=600, replace=False)) for _ in range(70)]
        rows = []
        for b in base:
            for _ in range(10):
                r = set(b.tolist())
                for _ in range(int(rng.integers(0, 3))):
                    r ^= {int(rng.integers(0, 90000))}
                rows.append(sorted(r))
        ip, ix, nc = rows_to_csr(rows, 90000)
    else:
        big = list(range(0, 120000, 3))
        rows = [big, big + [120001], [], [1], [1, 70000], big[:-1], [], list(range(65530, 65545)), list(range(65531, 65545)), [70000]]
        ip, ix, nc = rows_to_csr(rows, 130000)
    want_labels, _ = oracle.cluster(ip, ix, max_dist)
    ws, wd = oracle.edges(ip, ix, max_dist)
    for resident, verify16 in ((1, 0), (1, 1), (0, 0)):
        labels, src, dst, st = _run_ctx(ip, ix, nc, max_dist, resident_csr16=resident, verify_csr16=verify16)
        assert np.array_equal(labels, want_labels), (resident, verify16)
        assert np.array_equal(src, ws) and np.array_equal(dst, wd), (resident, verify16)


def test_resident_compact_form_rectangle_and_unsorted_rows():
    """query rows (the cache path's rectangle) on the compact form; a matrix whose rows are not ascending is refused by
    the encoder and handled by the plain path exactly as before"""
    ip, ix, nc = synth.generate(5000, seed=77).csr()
    q = np.arange(0, 5000, 9, dtype=np.int32)
    want = None
    for resident in (1, 0):
        labels, src, dst, _ = _run_ctx(ip, ix, nc, 2, query_rows=q, resident_csr16=resident)
        if want is None:
            want = (labels, src, dst)
            ws, wd = oracle.edges(ip, ix, 2, queries=q)
            assert np.array_equal(src, ws) and np.array_equal(dst, wd)
        assert np.array_equal(labels, want[0]) and np.array_equal(src, want[1]) and np.array_equal(dst, want[2])
    # descending rows: same answer with and without the option (the encoder flags the matrix, nothing changes)
    ix2 = ix.copy()
    for r in range(0, 50):
        ix2[ip[r]:ip[r + 1]] = ix2[ip[r]:ip[r + 1]][::-1]
    a = _run_ctx(ip, ix2, nc, 1, resident_csr16=1)
    b = _run_ctx(ip, ix2, nc, 1, resident_csr16=0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


@pytest.mark.parametrize("engine,options", [("hashjoin", {}), ("full", {}), ("sketch", {"sketch_bits": 512}), ("sketch", {"sketch_bits": 256})])
def test_compact_upload_then_engines_that_want_the_plain_csr(engine, options):
    """after bf_upload_csr16_async only the compact form is resident; engines and sketch widths whose kernels read the
    plain CSR get it decoded inside bf_run"""
    lib = _native.load()
    ip, ix, nc = synth.generate(7000, seed=55).csr()
    want, _ = oracle.cluster(ip, ix, 2)
    ip32, split, lo = _native.csr16_encode(ip, ix, nc)
    pins = [_pinned_copy(lib, ip32), _pinned_copy(lib, split) if split is not None else None, _pinned_copy(lib, lo)]
    with _native.Context(engine=engine, **options) as ctx:
        for _ in range(3):
            ctx.upload_csr16_async_ptr(pins[0].value, pins[1].value if pins[1] is not None else None, pins[2].value, len(ip) - 1, nc)
            ctx.run_sync(2)
            assert np.array_equal(ctx.download_labels(), want)
    for x in pins:
        if x is not None:
            lib.bf_pinned_free(x)


def test_measured_pipe_peaks_are_plausible():
    assert _native.measure_peak("popc32") > 1000.0
    assert _native.measure_peak("lop3") > _native.measure_peak("popc32")
    # int8 tensor rates: tcgen05.mma kind::i8 (TMEM accumulators) is about four times mma.sync m16n8k32
    imma, umma = _native.measure_peak("imma_s8"), _native.measure_peak("umma_i8")
    assert imma > 1e5 and 2.5 * imma < umma < 6 * imma, (imma, umma)


# ------------------------------------------------------------------ scale goldens (reference run on 42 k / 67 k sequences)
@pytest.fixture(scope="module")
def scale_tables(tmp_path_factory):
    import hashlib
    from tests import test_scale_golden as sg
    root = tmp_path_factory.mktemp("scale_gpu")
    out = {}
    for name, recipe in sg.SCALE["tables"].items():
        text = sg.build_table(recipe)
        assert hashlib.sha256(text.encode()).hexdigest() == recipe["sha256"]
        out[name] = root / f"{name}.tsv"
        out[name].write_text(text)
    return out


def _scale_cases():
    import json
    return json.loads((GOLDEN / "scale.json").read_text())["cases"]


@pytest.mark.parametrize("case", _scale_cases(), ids=[c["name"] for c in _scale_cases()])
def test_scale_goldens_through_the_cli(case, scale_tables, tmp_path):
    """clusters.tsv of the product CLI on the GPU == clusters.tsv the unmodified reference wrote (SHA-256 committed by
    tests/golden/make_scale_golden.py), on seeded tables of 41 643 / 66 807 sequences: d = 1 and 2, min-cluster-size 5,
    deletions kept, nextclade notation"""
    import hashlib
    from tests import test_scale_golden as sg
    got = sg.run_case(case, scale_tables[case["table"]], tmp_path)
    assert got.count("\n") == case["n_lines"]
    assert hashlib.sha256(got.encode()).hexdigest() == case["sha256"], case["name"]


def _scale_chain_cases():
    import json
    return json.loads((GOLDEN / "scale.json").read_text())["chain_cases"]


@pytest.mark.parametrize("case", _scale_chain_cases(), ids=[c["name"] for c in _scale_chain_cases()])
def test_scale_cached_chain_through_the_cli(case, tmp_path):
    """three cached runs (19 k -> 20 k -> 19 k sequences, ghost lists, new x all rectangle on the GPU): every
    clusters.tsv == what the unmodified reference wrote for the same chain (SHA-256 in scale.json)"""
    from tests import test_scale_golden as sg
    steps = sg.write_chain(tmp_path, case["chain"])
    outs, fresh = sg.run_chain(case, steps, tmp_path)
    sg.check_chain(case, outs, fresh)


# ------------------------------------------------------------------ BASELINE-size properties (1M profiles)
@pytest.fixture(scope="module")
def million():
    prof = synth.generate(1_000_000, seed=1)
    return prof.csr()


def test_million_profiles_properties(million):
    """At the headline size the oracle cannot run the full job in seconds, so check size-independent
    properties: exact neighbour sets on a row sample (completeness + soundness), every reported edge
    is a true edge, labels are canonical fixed points consistent with the edges, two sketch widths
    agree, and a rerun is identical."""
    indptr, indices, n_cols = million
    n = len(indptr) - 1
    with _native.Context(want_edges=1, sketch_bits=128) as ctx:
        ctx.upload_csr(indptr, indices, n_cols)
        st = ctx.run_sync(1)
        labels = ctx.download_labels()
        src, dst = ctx.download_edges()
        ctx.run_sync(1)
        assert np.array_equal(ctx.download_labels(), labels)
    assert st.n_edges == src.size and st.pairs_total == n * (n - 1) // 2
    # canonical labels: fixed points, never larger than the row, constant along every edge
    assert (labels <= np.arange(n)).all() and np.array_equal(labels[labels], labels)
    assert np.array_equal(labels[src], labels[dst])
    # components of exactly these edges (CPU union-find) == labels
    assert np.array_equal(oracle.components(n, src, dst), labels)
    # soundness on a sample of reported edges
    rng = np.random.default_rng(0)
    for k in rng.choice(src.size, size=min(3000, src.size), replace=False):
        assert oracle.distance(indptr, indices, int(src[k]), int(dst[k])) <= 1
    # completeness: exact neighbour sets of 1500 sampled rows, brute force on the CPU
    q = np.sort(rng.choice(n, size=1500, replace=False)).astype(np.int32)
    ws, wd = oracle.edges(indptr, indices, 1, queries=q)
    mask = np.isin(src, q) | np.isin(dst, q)
    assert edge_set(src[mask], dst[mask]) == edge_set(ws, wd)
    # a different sketch width gives the same answer
    with _native.Context(sketch_bits=512) as ctx:
        ctx.upload_csr(indptr, indices, n_cols)
        st2 = ctx.run_sync(1)
        assert np.array_equal(ctx.download_labels(), labels) and st2.n_edges == st.n_edges
    # and the WHOLE answer, bit for bit: at max_dist 1 the edge set has a closed form (rows that differ by one
    # column), which the second CPU oracle computes by a hash join + exact verification in seconds
    from oracle import hashjoin
    ws, wd = hashjoin.edges(indptr, indices, 1)
    order = np.lexsort((np.maximum(src, dst), np.minimum(src, dst)))
    assert np.array_equal(np.minimum(src, dst)[order], ws) and np.array_equal(np.maximum(src, dst)[order], wd)
    assert np.array_equal(labels, oracle.components(n, ws, wd))
