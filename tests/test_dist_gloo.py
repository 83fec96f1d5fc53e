"""world_size-2 check of the multi-rank host path on the CPU (gloo): cyclic work partition and the
label all-gather + merge semantics.  The edges each rank "finds" come from the oracle here; on the
GPU box tests/test_gpu_parity.py runs the same partition inside the real kernels."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import oracle
    from breakfast_b200 import dist as bdist
    from breakfast_b200 import synth
    from tests import helpers
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        indptr, indices, _ = synth.generate(600, seed=9).csr()
        n = len(indptr) - 1
        src, dst = oracle.edges(indptr, indices, 2)
        mine = slice(rank, None, world)                      # edge k belongs to rank k % world
        assert len(src[mine]) == bdist.rank_share(len(src), rank, world)
        local = oracle.components(n, src[mine], dst[mine])
        gathered = bdist.gather_labels(torch.from_numpy(local)).numpy()
        assert gathered.shape == (world, n)
        assert np.array_equal(gathered[rank], local)
        merged = helpers.merge_labels_cpu(gathered)
        want = oracle.components(n, src, dst)
        q.put((rank, bool(np.array_equal(merged, want)), int(len(set(local.tolist()))), int(len(set(want.tolist())))))
    finally:
        dist.destroy_process_group()


def test_two_rank_partition_gather_merge():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, n_local, n_full in results:
        assert ok, f"rank {rank}: merged labels differ from the single-rank result"
        assert n_local > n_full      # each rank alone sees a finer partition


class _FakeCtx:
    """stands in for _native.Context on the CPU: `fail_first` syncs report an overflow"""

    def __init__(self, fail_first):
        self.fail_first, self.runs = fail_first, 0

    def run(self, max_dist, rank, world):
        self.runs += 1

    def labels_to_device(self, ptr):
        pass

    def merge_labels_device(self, ptr, world):
        pass

    def sync(self):
        from breakfast_b200 import _native
        if self.runs <= self.fail_first:
            raise _native.NativeError(_native.BF_ERR_OVERFLOW, "overflow (test)")
        return "stats"


def _overflow_worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from breakfast_b200 import dist as bdist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = _FakeCtx(fail_first=2 if rank == 1 else 0)      # only rank 1 overflows, twice
        runner = bdist.RankRunner.__new__(bdist.RankRunner)  # no CUDA here: fill the fields by hand
        runner.ctx, runner.n, runner.rank, runner.world, runner.group = ctx, 4, rank, world, None
        runner.local = torch.zeros(4, dtype=torch.int32)
        runner.gathered = torch.zeros((world, 4), dtype=torch.int32)
        st = runner.run_sync(1)
        q.put((rank, st, ctx.runs))
    finally:
        dist.destroy_process_group()


def test_overflow_on_one_rank_reruns_every_rank():
    """the label exchange is a collective: when one rank has to rerun after an overflow, all of them must"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overflow_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == [(0, "stats", 3), (1, "stats", 3)]


@pytest.mark.parametrize("n_work,world", [(0, 1), (1, 2), (7, 2), (8, 4), (1000003, 8)])
def test_rank_share_partitions_exactly(n_work, world):
    from breakfast_b200.dist import rank_share
    shares = [rank_share(n_work, r, world) for r in range(world)]
    assert sum(shares) == n_work and max(shares) - min(shares) <= 1
