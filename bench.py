#!/usr/bin/env python
"""bench.py — candidate pairs/s of the breakfast distance-and-clustering hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Workload (BASELINE.json metric): 1,000,000 unique synthetic SARS-CoV-2 profiles (seed 1), --max-dist 1,
strictly binary CSR of the filtered profiles.  One step = one pass of the hot path on that batch:
cardinality sort -> bit-pack -> band tile schedule -> tiled XOR/POPC pair kernel -> exact verify ->
union-find -> labels (N > 1: the band tile pairs are dealt cyclically to the ranks, then the labels are
all-gathered over NCCL and re-united on every rank).  `value` = candidate pairs (||A|-|B|| <= max_dist,
SURVEY 8d) per second with the CSR already resident in HBM; `e2e` = the same through the C ABI with
pinned HOST buffers (H2D of the CSR and D2H of the labels inside the timed region, every step).

The line also carries `roofline` (the pair kernel against the POPC pipe peak measured in this run) and
`cpu_baseline` (the oracle port — the reference's per-cardinality scikit-learn path — on a bounded sample
of the same workload, on this box's host cores).  `--impl reference` times only that CPU arm.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_PROFILES = 1_000_000
SEED = 1
MAX_DIST = 1
CPU_SAMPLE = 16_000          # profiles in the bounded CPU sample (about 10-20 s of scikit-learn work)
METRIC = "candidate_pairs_per_s"
UNIT = "pairs/s"
WORKLOAD = f"synthetic {N_PROFILES} unique SARS-CoV-2 covsonar_dna profiles (seed {SEED}), --max-dist {MAX_DIST}"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def pairs_band_of(card, max_dist):
    import numpy as np
    h = np.bincount(card).astype(object)
    total = sum(int(c) * (int(c) - 1) // 2 for c in h)
    for k in range(1, max_dist + 1):
        total += sum(int(h[c]) * int(h[c + k]) for c in range(len(h) - k))
    return int(total)


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (scikit-learn pairwise_distances_chunked per cardinality)
# ------------------------------------------------------------------------------------------------
def cpu_sample_setup(n_sample):
    import numpy as np
    from scipy.sparse import csr_matrix
    from breakfast_b200 import synth
    prof = synth.generate(N_PROFILES, seed=SEED)
    indptr, indices, n_cols = prof.csr()
    ip = indptr[: n_sample + 1]
    ix = indices[: ip[-1]]
    X = csr_matrix((np.ones(ix.size, dtype=np.int64), ix.astype(np.int64), ip), shape=(n_sample, n_cols))
    band = pairs_band_of(np.diff(ip), MAX_DIST)
    return X, band


def cpu_step(X):
    from oracle import ref_port
    stats = {}
    t0 = time.perf_counter()
    lists = ref_port.neighbour_lists(X, MAX_DIST, stats=stats)
    n_links = sum(len(l) for l in lists)
    return time.perf_counter() - t0, stats.get("evaluations", 0), n_links


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)   # before scikit-learn is imported (console.py:150)
    X, band = cpu_sample_setup(CPU_SAMPLE)
    for _ in range(args.warmup):
        cpu_step(X)
    times, evals = [], 0
    for _ in range(args.steps):
        t, evals, _ = cpu_step(X)
        times.append(t)
    total = sum(times)
    value = band * args.steps / total
    sample = (f"first {CPU_SAMPLE} profiles of the workload per step: {band} candidate pairs, {evals} ordered "
              f"distance evaluations by the reference's per-cardinality batches")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "max_dist": MAX_DIST, "sample_profiles": CPU_SAMPLE},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 50 ms from before the warm-up; only samples that fall inside the timed
    region are reported (all samples are used, and flagged, if the region was shorter than one poll)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None
        self.nvml_rows, self.nvml_stop, self.nvml_thread = [], threading.Event(), None

    def _nvml_loop(self):
        """NVML polled every 2 ms next to nvidia-smi: a timed region of a few milliseconds still holds several samples.
        Any NVML trouble just ends this loop - the nvidia-smi rows remain."""
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.nvml_stop.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                except Exception:
                    mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.nvml_rows.append((time.time(), int(sm), int(mx), int(mask)))
                time.sleep(0.002)
            pynvml.nvmlShutdown()
        except Exception:
            pass

    def start(self):
        try:
            self.nvml_thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.nvml_thread.start()
        except Exception:
            self.nvml_thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        self.thread.join(timeout=2)
        self.nvml_stop.set()
        if self.nvml_thread is not None:
            self.nvml_thread.join(timeout=2)
        fine = [r for r in self.nvml_rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or r[0])]
        if len(fine) >= 3:
            # bits of nvmlClocksEventReasons: 0x4 sw_power_cap, 0x8 hw_slowdown, 0x20 sw_thermal_slowdown, 0x40 hw_thermal_slowdown
            bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            sm = sorted(r[1] for r in fine)
            reasons = sorted({name for r in fine for bit, name in bits.items() if r[3] & bit})
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(r[2] for r in fine), "reasons": reasons, "samples": len(sm),
                    "scope": "timed region", "source": "NVML polled every 2 ms (nvidia-smi -lms 50 beside it: "
                                                       f"{len(self.rows)} rows over warm-up + timed region)"}
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.05]
        scope = "timed region"
        if not inside:
            inside, scope = [r for _, r in self.rows], "warm-up + timed region (timed region shorter than one 50 ms poll)"
        sm = sorted(int(r[0]) for r in inside if r and r[0].isdigit())
        mx = [int(r[1]) for r in inside if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in inside if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "scope": scope}



# ------------------------------------------------------------------------------------------------
# extra measurements of the N = 1 line: other configs, CLI wall-clock, FULL-engine POPC roofline
# ------------------------------------------------------------------------------------------------
def time_device_steps(_native, stream, csr, max_dist, query_rows=None, steps=5, warmup=3, **opts):
    """ms per device-resident pass (CUDA events inside the library, summed over `steps` passes) + counters"""
    indptr, indices, n_cols = csr
    with _native.Context(device=0, stream=stream, **opts) as ctx:
        ctx.upload_csr(indptr, indices, n_cols, query_rows=query_rows)
        for _ in range(warmup):
            ctx.run_sync(max_dist)
        for _ in range(steps):
            ctx.run(max_dist)
        st = ctx.sync()
    return st.ms_total_sum / max(1, min(st.runs_since_sync, 128)), st


def measure_configs(_native, stream, np):
    """device step of BASELINE configs 2-5 at their stated shapes (parity of each: tests/test_gpu_configs.py)"""
    from breakfast_b200 import synth
    out = {}
    ms, st = time_device_steps(_native, stream, synth.generate(100_000, seed=2).csr(), 1)
    out["config2_100k_d1"] = {"ms_per_step": ms, "candidate_pairs": st.pairs_band, "edges": st.n_edges}
    c3 = synth.generate(1_000_000, seed=3, with_mult=True).csr()
    ms, st = time_device_steps(_native, stream, c3, 2)
    out["config3_1m_d2"] = {"ms_per_step": ms, "candidate_pairs": st.pairs_band, "edges": st.n_edges, "tiles_band": st.tiles_band}
    ms, st = time_device_steps(_native, stream, synth.generate(500_000, seed=4, with_mult=True, unique_on_all_events=True).csr(), 1)
    out["config4_500k_nextclade_shape_d1"] = {"ms_per_step": ms, "candidate_pairs": st.pairs_band, "edges": st.n_edges}
    q = np.sort(np.random.default_rng(5).choice(1_000_000, size=50_000, replace=False)).astype(np.int32)
    ms, st = time_device_steps(_native, stream, c3, 2, query_rows=q)
    out["config5_rectangle_50k_x_1m_d2"] = {"ms_per_step": ms, "candidate_pairs": st.pairs_band, "edges": st.n_edges}
    return out


def measure_full_engine_roofline(_native, stream, peaks):
    """the literal north-star kernel (FULL engine: every column as a bit, tiled XOR/POPC, executed == algorithmic work)
    against the POPC pipe peak measured in this process, at a size that fits the time budget"""
    from breakfast_b200 import synth
    csr = synth.generate(100_000, seed=1).csr()
    ms, st = time_device_steps(_native, stream, csr, 1, steps=3, warmup=2, engine="full")
    ms_pairs = st.ms_pairs_sum / max(1, min(st.runs_since_sync, 128))
    gpopc = st.popc32_executed / (ms_pairs * 1e-3) / 1e9
    return {"kernel": f"k_pairs<4> (FULL engine, {st.bits_per_row} bits/row, 100 000 profiles)", "bound": "int_pipe_popc",
            "achieved": gpopc, "peak": peaks["popc32"], "unit": "GPOPC32/s", "frac": gpopc / peaks["popc32"], "ms_per_launch": ms_pairs,
            "popc32_per_launch": st.popc32_executed, "pairs_evaluated": st.pairs_evaluated}


def measure_cli_wall(n_profiles):
    """wall-clock of the product CLI (console.main) on the headline table: n unique profiles, one sequence each"""
    import tempfile
    import click.testing
    from breakfast_b200 import console, synth
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        t0 = time.perf_counter()
        table = synth.generate(n_profiles, seed=SEED).table("covsonar_dna", " ")
        path = tmp / "table.tsv"
        table.to_csv(path, sep="\t", index=False)
        t_make = time.perf_counter() - t0
        size = path.stat().st_size
        del table
        res = click.testing.CliRunner().invoke(console.main, ["--input-file", str(path), "--outdir", str(tmp / "out"), "--max-dist", str(MAX_DIST)])
        if res.exit_code != 0:
            return {"error": f"CLI failed: {res.exception!r}"}
        t = dict(console.LAST_TIMINGS)
        from breakfast_b200 import engine as _engine
        engine_split = dict(_engine.LAST_TIMINGS)
        clustered = sum(1 for line in open(tmp / "out" / "clusters.tsv") if not line.rstrip("\n").endswith("\t")) - 1
    host = t["total"] - t.get("engine", 0.0)
    return {"cli_wall_s": t["total"], "host_s": host, "stages_s": {k: t[k] for k in ("read", "prepare", "cluster", "write")},
            "engine_s": t.get("engine"), "engine_split_s": engine_split, "host_path": t.get("host_path"), "sequences": n_profiles, "table_bytes": size,
            "clustered_sequences": clustered, "table_generation_s": t_make,
            "note": "read = pandas read_table; prepare = filter + dedup + CSR (native host pass); cluster = engine (H2D, kernels, "
                    "D2H; engine_s) + labelling; write = clusters.tsv"}


def measure_port_stages(n_sequences=30_000):
    """per-stage host times of the oracle port (the reference's host steps restated) on a bounded sample, for the split
    next to cli_wall_s (BASELINE.md section 3 iii)"""
    import tempfile
    from breakfast_b200 import synth
    from oracle import ref_port
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp) / "t.tsv"
        synth.generate(n_sequences, seed=SEED).table("covsonar_dna", " ").to_csv(path, sep="\t", index=False)
        t = {}
        t0 = time.perf_counter(); ids, feats = ref_port.read(path); t["read"] = time.perf_counter() - t0
        t0 = time.perf_counter(); feats = ref_port.filter_profiles(feats, " ", "covsonar_dna", True, True, 264, 228, 29903); t["filter"] = time.perf_counter() - t0
        t0 = time.perf_counter(); uniq, codes, mult = ref_port.dedup(feats); t["dedup"] = time.perf_counter() - t0
        t0 = time.perf_counter(); ref_port.count_matrix(uniq, " "); t["vectorise"] = time.perf_counter() - t0
    return {"sequences": n_sequences, "stages_s": t, "us_per_sequence": {k: 1e6 * v / n_sequences for k, v in t.items()},
            "kind": "port", "note": "single-threaded Python, as in the reference; the neighbour search is the cpu_baseline"}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sketch-bits", type=int, default=int(os.environ.get("BREAKFAST_B200_SKETCH_BITS", "128")))
    ap.add_argument("--engine", default="sketch", choices=["sketch", "full", "hashjoin"])
    ap.add_argument("--two-level", type=int, default=1, choices=[0, 1])
    ap.add_argument("--level1", type=int, default=2, choices=[0, 1, 2],
                    help="2 = int8 mma.sync level 1 with two column rows per accumulator (default), 1 = one row per accumulator, 0 = integer pipes")
    ap.add_argument("--l1-ctas", type=int, default=0, choices=[0, 1, 2], help="CTAs per SM of the level-1 kernel (0 = library default)")
    ap.add_argument("--profiles", type=int, default=N_PROFILES, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip configs_ms / cli_wall_s / full-engine roofline (N = 1 extras)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    from breakfast_b200 import _native, synth
    from breakfast_b200.dist import RankRunner

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            log(f"--gpus {args.gpus} needs one process per GPU: launch with python -m torch.distributed.run "
                f"--nnodes=1 --nproc-per-node {args.gpus} --master-addr 127.0.0.1 ... bench.py --gpus {args.gpus}")
            return 2
        args.gpus = world
    _native.require_device()
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.profiles
    t0 = time.time()
    prof = synth.generate(n, seed=SEED)
    indptr, indices, n_cols = prof.csr()
    del prof
    if rank == 0:
        log(f"[bench] generated {n} profiles, nnz={indices.size}, n_cols={n_cols} in {time.time() - t0:.1f}s")

    # pinned host copies (the e2e leg copies from these every step)
    lib = _native.load()
    p_indptr, p_indices = C.c_void_p(), C.c_void_p()
    assert lib.bf_pinned_alloc(indptr.nbytes, C.byref(p_indptr)) == 0
    assert lib.bf_pinned_alloc(indices.nbytes, C.byref(p_indices)) == 0
    C.memmove(p_indptr, indptr.ctypes.data, indptr.nbytes)
    C.memmove(p_indices, indices.ctypes.data, indices.nbytes)
    # the compact host form of the same matrix (bf_csr16_encode: 32-bit offsets, 16-bit columns + a split per row), in
    # pinned memory: what the N = 1 e2e leg ships every step when the matrix is representable
    csr16 = None
    if n_cols <= 131072 and int(np.diff(indptr).max(initial=0)) <= 65535:
        sizes = ((n + 1) * 4, n * 2 if n_cols > 65536 else 0, int(indices.size) * 2)
        ptrs = []
        for nb in sizes:
            q = C.c_void_p()
            assert lib.bf_pinned_alloc(max(nb, 1), C.byref(q)) == 0
            ptrs.append(q)
        _native._ck(lib.bf_csr16_encode(indptr.ctypes.data, indices.ctypes.data, n, n_cols, ptrs[0], ptrs[1] if sizes[1] else None, ptrs[2]))
        csr16 = (ptrs, sizes)
    labels_host = np.empty(n, dtype=np.int32)
    p_labels = C.c_void_p()
    assert lib.bf_pinned_alloc(labels_host.nbytes, C.byref(p_labels)) == 0

    peaks = {name: _native.measure_peak(name, local_rank) for name in ("popc32", "lop3", "imma_s8", "umma_i8")}  # also warms the clocks

    # a real (non-default) torch stream: the library enqueues on it, torch events and NCCL order against it
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    ctx = _native.Context(device=local_rank, stream=stream, engine=args.engine, sketch_bits=args.sketch_bits,
                          two_level=args.two_level, level1=args.level1, l1_ctas=args.l1_ctas)
    ctx.upload_csr_ptr(p_indptr.value, p_indices.value, n, n_cols)
    runner = RankRunner(ctx, n, rank, world)   # N > 1: joins the library's own NCCL communicator (NCCL may print its banner on stdout)

    # ---- device-resident leg: `value`
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    runner.run_sync(MAX_DIST)        # settles the bounded buffers (an overflow on any rank reruns every rank)
    for _ in range(args.warmup):
        runner.step(MAX_DIST)
    st = ctx.sync()
    barrier()
    clocks.mark_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        runner.step(MAX_DIST)
    ev1.record()
    barrier()
    clocks.mark_end()
    clock_info = clocks.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    st = ctx.sync()      # counters + per-launch sums over the timed steps
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = st.pairs_band / (ms_per_step * 1e-3)

    # per-launch duration of the dominant kernel over the timed region (CUDA events on the launching stream)
    runs = max(1, min(st.runs_since_sync, 128))
    ms_pairs = st.ms_pairs_sum / runs
    two_kernel = args.engine == "sketch" and args.two_level and st.bits_per_row in (128, 256)
    if two_kernel and args.level1 == 2:
        # dominant kernel = k_pairs_l1_imma2: the column operand packs two rows per int8 row (e1 + 64 e2), so one m16n8k32
        # MMA decides 256 pairs: 16 executed int8 MACs per evaluated pair for the 32 MACs of the plain contraction
        # (DESIGN.md section 3).  `achieved` counts the ALGORITHMIC 32 MACs per pair (2 F int8 ops per
        # evaluated entry, F = 32 fold bits), `frac_executed` the MACs the tensor pipe really performs.
        ms_kernel = st.ms_l1_sum / runs
        kernel_name = f"k_pairs_l1_imma2 (int8 mma.sync, two column rows per accumulator, 32-bit folds of {st.bits_per_row}-bit sketches)"
        bound, unit_r, peak = "tensor", "Gint8-MAC/s (mma.sync m16n8k32)", peaks["imma_s8"]
        ops_per_launch = int(st.pairs_evaluated * 32)
        achieved = ops_per_launch / (ms_kernel * 1e-3) / 1e9
        sm_hz = 148 * 1.965e9
        extra = {"macs_per_pair": 32, "executed_macs_per_pair": 16, "frac_executed": achieved / 2 / peak,
                 "pairs_per_clk_per_sm": st.pairs_evaluated / (ms_kernel * 1e-3) / sm_hz,
                 "level2_units": st.l2_warp_items, "ms_level2": (st.ms_pairs_sum - st.ms_l1_sum) / runs,
                 "tcgen05_int8_peak": peaks["umma_i8"], "frac_of_tcgen05_int8_peak": achieved / peaks["umma_i8"],
                 "note": "frac = algorithmic MACs (32 per evaluated pair) over the int8 rate reachable with register accumulators "
                         "(mma.sync), frac_executed = the MACs really issued (16 per pair: two column rows share an accumulator); "
                         "per pair the kernel also issues half an IMAD and a quarter of a packed 16x2 max, which is what binds "
                         "next (ncu under profiles/).  frac_of_tcgen05_int8_peak = against the tcgen05.mma kind::i8 rate measured "
                         "in this process; the tcgen05 form of this filter was built and instrumented twice "
                         "(profiles/r02_tcgen05_*.log): one TMEM -> register read per pair and >= 1000 cycles per accumulator "
                         "round trip cap it at 12-18 pairs/clk/SM"}
    elif two_kernel and args.level1 == 1:
        # dominant kernel = k_pairs_l1_imma: one m16n8k32 int8 MMA per 128 pairs = 32 int8 MACs per evaluated pair
        # (DESIGN.md section 3); bound = the tensor pipe as reachable through mma.sync, peak measured in this process
        ms_kernel = st.ms_l1_sum / runs
        kernel_name = f"k_pairs_l1_imma (int8 mma.sync on the +-1 expanded 32-bit folds of {st.bits_per_row}-bit sketches)"
        bound, unit_r, peak = "tensor", "Gint8-MAC/s (mma.sync m16n8k32)", peaks["imma_s8"]
        ops_per_launch = int(st.pairs_evaluated * 32)
        achieved = ops_per_launch / (ms_kernel * 1e-3) / 1e9
        extra = {"macs_per_pair": 32, "level2_units": st.l2_warp_items, "ms_level2": (st.ms_pairs_sum - st.ms_l1_sum) / runs,
                 "tcgen05_int8_peak": peaks["umma_i8"], "frac_of_tcgen05_int8_peak": achieved / peaks["umma_i8"],
                 "note": "side by side: frac = of the int8 rate reachable with register accumulators (mma.sync), "
                         "frac_of_tcgen05_int8_peak = of the tcgen05.mma kind::i8 rate measured in this process (UTCIMMA, TMEM "
                         "accumulators).  The tcgen05 form of this kernel was built and instrumented twice (profiles/r02_tcgen05_*.log): "
                         "every pair costs one TMEM -> register read, TMEM holds 65 536 accumulators and one accumulator round trip "
                         "(issue -> MMA -> commit -> wake -> tcgen05.ld -> release -> wake) takes >= 1000 cycles at K <= 128, so it peaks "
                         "at 12-18 pairs/clk/SM against 35 for this kernel"}
    elif two_kernel:
        # dominant kernel = k_pairs_l1<T>.  Per evaluated pair it executes:
        #   T = max_dist in {1,2}: 1 XOR + T/2 AND + 1/2 min on the ALU pipe, 1/2 POPC on the XU pipe, T/2 IMAD (FMA)
        #   otherwise            : 1 XOR + 1/2 min on the ALU pipe, 1 POPC on the XU pipe
        hybrid = MAX_DIST in (1, 2)
        alu_per_pair = 1.0 + (MAX_DIST / 2.0 if hybrid else 0.0) + 0.5
        popc_per_pair = 0.5 if hybrid else 1.0
        ms_kernel = st.ms_l1_sum / runs
        kernel_name = f"k_pairs_l1<{MAX_DIST if hybrid else 0}> (32-bit fold level 1 of {st.bits_per_row}-bit sketches)"
        alu_rate = st.pairs_evaluated * alu_per_pair / (ms_kernel * 1e-3) / 1e9
        xu_rate = st.pairs_evaluated * popc_per_pair / (ms_kernel * 1e-3) / 1e9
        alu_frac, xu_frac = alu_rate / peaks["lop3"], xu_rate / peaks["popc32"]
        if alu_frac >= xu_frac:
            bound, achieved, peak, unit_r = "int_pipe_alu", alu_rate, peaks["lop3"], "G lane-op/s (LOP3-class)"
        else:
            bound, achieved, peak, unit_r = "int_pipe_popc", xu_rate, peaks["popc32"], "GPOPC32/s"
        ops_per_launch = int(st.pairs_evaluated * (alu_per_pair if bound == "int_pipe_alu" else popc_per_pair))
        extra = {"alu_frac": alu_frac, "xu_frac": xu_frac, "alu_ops_per_pair": alu_per_pair, "popc_per_pair": popc_per_pair,
                 "level2_units": st.l2_warp_items, "ms_level2": (st.ms_pairs_sum - st.ms_l1_sum) / runs}
    elif args.engine == "hashjoin":
        # the probes of the hash-join engine: nnz (distance 1) random 8-byte reads of an L2-resident table; algorithmic
        # traffic = one 32-byte sector per probe, reported against the HBM peak for want of a measured L2 figure
        ms_kernel = ms_pairs
        kernel_name = "k_hj_probe_deletions<1> (one probe per stored column against the row table)"
        hbm = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
        bound, unit_r, peak = "hbm", "GB/s (32-byte sector per probe; the table is L2-resident)", hbm
        ops_per_launch = int(indices.size // world) * 32
        achieved = ops_per_launch / (ms_kernel * 1e-3) / 1e9
        extra = {"probes_per_launch": int(indices.size // world)}
    else:
        ms_kernel = ms_pairs
        kernel_name = f"k_pairs<K4> ({args.engine}, {st.bits_per_row} bits/row)"
        bound, unit_r, peak = "int_pipe_popc", "GPOPC32/s", peaks["popc32"]
        ops_per_launch = int(st.popc32_executed)
        achieved = ops_per_launch / (ms_kernel * 1e-3) / 1e9
        extra = {}
    traffic = None
    tf = ROOT / "profiles" / "roofline_traffic.json"
    if tf.exists():
        try:
            key = (("k_pairs_l1_imma2" if args.level1 == 2 else "k_pairs_l1_imma" if args.level1 == 1 else "k_pairs_l1") if two_kernel else "k_pairs") + \
                f"_{args.engine}_{st.bits_per_row}_n{n}_w{world}"
            traffic = json.loads(tf.read_text()).get(key)
        except Exception:
            traffic = None
    launches = st.kernel_launches

    # ---- end-to-end leg through the host-buffer C ABI: every step copies its CSR from pinned host memory
    # and reads its labels back to the host.
    #   N = 1: bf_upload_csr_async double-buffers, so the H2D of step k+1 overlaps the pass of step k.
    #   N > 1: every rank copies 1/N of the column array, the slices are all-gathered over NVLink (NCCL) and
    #          adopted in place (bf_adopt_csr_device); then the pass, the label all-gather + merge, and D2H.
    if csr16 is not None:
        # compact host form; with N ranks every rank copies 1/N of the 16-bit column array over its own host link and the
        # shares are all-gathered over NVLink inside the library (copy stream, its own communicator), then decoded
        (q_ip, q_split, q_lo), sizes = csr16
        h2d_bytes = int(sizes[0] + sizes[1] + -(-sizes[2] // world))

        def upload():
            ctx.upload_csr16_async_ptr(q_ip.value, q_split.value if sizes[1] else None, q_lo.value, n, n_cols)
    else:
        h2d_bytes = int(indptr.nbytes + indices.nbytes)

        def upload():
            ctx.upload_csr_async_ptr(p_indptr.value, p_indices.value, n, n_cols)

    def e2e_loop(k_steps):
        upload()
        for k in range(k_steps):
            runner.step(MAX_DIST)                         # waits for the upload of this step
            if k + 1 < k_steps:
                upload()                                  # overlaps the pass
            _native._ck(lib.bf_download_labels(ctx._h, p_labels))

    e2e_loop(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / args.steps
    st_e2e = ctx.sync()
    C.memmove(labels_host.ctypes.data, p_labels, labels_host.nbytes)
    ok = bool((labels_host <= np.arange(n)).all() and np.array_equal(labels_host[labels_host], labels_host))
    import hashlib
    labels_sha = hashlib.sha256(labels_host.tobytes()).hexdigest()   # canonical labels: identical for every N

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            os.environ["OMP_NUM_THREADS"] = str(cores)
            import numpy as _np
            from scipy.sparse import csr_matrix
            ns = min(CPU_SAMPLE, n)
            ip = indptr[: ns + 1]
            ix = indices[: ip[-1]]
            X = csr_matrix((_np.ones(ix.size, dtype=_np.int64), ix.astype(_np.int64), ip), shape=(ns, n_cols))
            band = pairs_band_of(_np.diff(ip), MAX_DIST)
            cold, _, _ = cpu_step(X)            # first call: imports scikit-learn, spins up its thread pool
            sec, evals, _ = cpu_step(X)         # warm: what the reference arm (--impl reference) also reports
            cpu = {"value": band / sec, "unit": UNIT, "cores": cores, "kind": "port", "cold_first_call_s": cold,
                   "sample": f"first {ns} profiles of the workload: {band} candidate pairs, {evals} ordered distance "
                             f"evaluations (reference batches through scikit-learn), {sec:.1f} s warm ({cold:.1f} s for the cold first call)"}
        extras = {}
        if world == 1 and not args.no_extras and n == N_PROFILES:
            try:
                extras["configs_ms"] = measure_configs(_native, stream, np)
                extras["full_engine_roofline"] = measure_full_engine_roofline(_native, stream, peaks)
                extras["cli"] = measure_cli_wall(n)
                extras["cli"]["port_stages"] = measure_port_stages()
            except Exception as exc:   # the headline numbers above are already taken; say what went wrong
                extras["extras_error"] = repr(exc)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOAD if n == N_PROFILES else f"{n} profiles (override)", "max_dist": MAX_DIST,
                       "engine": args.engine, "bits_per_row": st.bits_per_row, "two_level": bool(args.two_level), "level1": {2: "imma_packed_pairs", 1: "imma", 0: "int_pipes"}[args.level1],
                       "l2_warp_items": st.l2_warp_items, "pairs_evaluated": st.pairs_evaluated, "n_cols": n_cols, "nnz": int(indices.size),
                       "candidate_pairs": st.pairs_band, "pairs_total": st.pairs_total, "tiles_band": st.tiles_band,
                       "edges": None if world > 1 else st.n_edges, "components": st.n_components, "sketch_survivors_rank0": st.n_candidates,
                       "l2_policy": "inputs larger than L2 (CSR 360 MB + per-step rebuilt bitsets); no explicit flush",
                       "parallelism": f"tile-partition x{world}" if world > 1 else "single GPU"},
            "clocks": clock_info,
            "e2e": {"value": st.pairs_band / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(labels_host.nbytes),
                    "input_path": ("compact host form (bf_upload_csr16_async: 32-bit offsets + 16-bit columns + per-row split, decoded on the "
                                   "device), double-buffered: copy and decode of step k+1 overlap pass k"
                                   + ("; every rank copies 1/N of the columns, NCCL all-gather over NVLink in the library" if world > 1 else "")
                                   if csr16 is not None else "double-buffered async H2D of the plain CSR (copy of step k+1 overlaps pass k)"),
                    "labels_sane": ok, "labels_sha256": labels_sha},
            "gpu_launches": int(launches),
            "roofline": dict({"kernel": kernel_name, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit_r,
                              "frac": achieved / peak, "traffic": traffic, "ms_per_launch": ms_kernel,
                              "ops_per_launch": ops_per_launch,
                              "peak_source": "bf_measure_peak('imma_s8'/'lop3'/'popc32') measured in this process; MEASURED_PEAKS.json "
                                             "holds only HBM and bf16 (cuBLAS/tcgen05) peaks, not the int8 mma.sync or integer-pipe rates that bind here",
                              "kernel_share_of_step": ms_kernel / (st.ms_total_sum / runs) if st.ms_total_sum else None}, **extra),
            "cpu_baseline": cpu,
            "phases_ms": {k: getattr(st, k) for k in ("ms_sort", "ms_pack", "ms_pairs", "ms_verify", "ms_cc", "ms_merge")},
        }
        # second kernel of the step: the exact verification, an HBM-bound gather of two CSR rows per candidate
        # (algorithmic bytes: the two rows + their extents; peak: MEASURED_PEAKS.json, else the profiling guide's fallback)
        if args.engine == "sketch" and world == 1 and st.n_candidates > 0:
            peaks_file = ROOT / "MEASURED_PEAKS.json"
            hbm = json.loads(peaks_file.read_text())["hbm_gbs"] if peaks_file.exists() else 6650.0
            row_bytes = 4.0 * indices.size / n + 16.0
            vbytes = st.n_candidates * 2 * row_bytes
            vtraffic = None
            try:
                vtraffic = json.loads(tf.read_text()).get(f"k_verify_unite_{args.engine}_{st.bits_per_row}_n{n}")
            except Exception:
                pass
            line["roofline_verify"] = {"kernel": "k_verify_unite<max_dist> (exact |A xor B| on the CSR rows of the sketch survivors + union-find hook)",
                                       "bound": "hbm", "achieved": vbytes / (st.ms_verify * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                       "frac": vbytes / (st.ms_verify * 1e-3) / 1e9 / hbm, "traffic": vtraffic,
                                       "ms_per_launch": st.ms_verify, "candidates": st.n_candidates,
                                       "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks_file.exists() else "fallback 6650 GB/s",
                                       "note": "scattered 360-byte rows: 12-13 sectors per row at random addresses; more resident warps or "
                                               "half the bytes (compact form) were measured slower (DESIGN.md section 3)"}
        line.update(extras)
        print(json.dumps(line), flush=True)
    ctx.close()
    for p in (p_indptr, p_indices, p_labels) + (tuple(csr16[0]) if csr16 else ()):
        lib.bf_pinned_free(p)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
