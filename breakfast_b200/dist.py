"""Multi-GPU host logic: one process per GPU (torchrun), torch.distributed for the plumbing.

The band tile pairs of the (upper-triangular) tile space are dealt cyclically to the ranks inside
the pair kernel (work item w belongs to rank w % world, kernels.cuh k_pairs); every rank holds the
whole CSR and runs a local union-find over the edges it found.  Two exchange steps, both inside
bf_run on the library's own NCCL communicator (include/breakfast_b200.h, "multi-GPU"): the shares of the
sketch + sort-key pass are all-gathered (every rank streams 1/N of the rows), and after the verify
step the ranks' union-finds are exchanged as compact (row, root) lists and re-united.  torch.distributed
only carries the 128-byte communicator id (and the benchmark's barriers).  Without the library
communicator (`lib_comm=False`, or a CPU process group in the tests) the older form remains: an
all-gather of int32 labels[n_rows] per rank through torch.distributed + k_uf_merge_labels.
"""
from __future__ import annotations


def rank_share(n_work: int, rank: int, world: int) -> int:
    """How many of n_work cyclically dealt tile pairs rank `rank` owns (host mirror of the kernel's
    `for (w = rank + world*block; w < W; w += world*grid)` partition)."""
    return (n_work - rank + world - 1) // world if n_work > rank else 0


def gather_labels(local, group=None):
    """all-gather equal-length int32 label vectors -> tensor [world, n] on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world, local.numel()), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    else:
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local.contiguous(), group=group)
        out = torch.stack(parts)
    return out


def init_library_comm(ctx, rank: int, world: int, group=None) -> None:
    """Give `ctx` the library's own NCCL communicator: rank 0 draws the id, torch.distributed broadcasts its 128 bytes."""
    import torch
    import torch.distributed as dist
    from . import _native
    dev = torch.device("cuda", torch.cuda.current_device())
    box = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        box.copy_(torch.frombuffer(bytearray(_native.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.comm_init_rank(bytes(box.cpu().numpy().tobytes()), rank, world)


class RankRunner:
    """Drives one rank's context through one pass per step.  With the library communicator (default on GPUs) a step is
    just bf_run: the exchange steps happen inside it.  Otherwise run -> all-gather of the labels -> merge; the context
    must then have been created on torch's current CUDA stream so that kernels and collectives are ordered."""

    def __init__(self, ctx, n_rows: int, rank: int, world: int, group=None, lib_comm: bool = True):
        import torch
        self.ctx, self.n, self.rank, self.world, self.group = ctx, n_rows, rank, world, group
        dev = torch.device("cuda", torch.cuda.current_device())
        self.local = self.gathered = None
        if world > 1 and lib_comm and ctx.comm is None:
            init_library_comm(ctx, rank, world, group)
        if world > 1 and ctx.comm is None:
            self.local = torch.empty(max(n_rows, 1), dtype=torch.int32, device=dev)
            self.gathered = torch.empty((world, max(n_rows, 1)), dtype=torch.int32, device=dev)

    def step(self, max_dist: int):
        """Enqueue one pass; returns nothing — call ctx.sync() for counters."""
        import torch.distributed as dist
        self.ctx.run(max_dist, self.rank, self.world)
        if self.world > 1 and getattr(self.ctx, "comm", None) is None:
            self.ctx.labels_to_device(self.local.data_ptr())
            if self.local.is_cuda:
                dist.all_gather_into_tensor(self.gathered, self.local, group=self.group)
            else:   # CPU tensors over gloo: tests/test_dist_gloo.py drives this class with a stand-in context
                dist.all_gather(list(self.gathered.unbind(0)), self.local, group=self.group)
            self.ctx.merge_labels_device(self.gathered.data_ptr(), self.world)

    def run_sync(self, max_dist: int, attempts: int = 6):
        """One synchronous pass on every rank; returns this rank's counters.  A rank whose bounded device buffers
        overflowed (bf_sync has raised their capacity) must run again - and because the label exchange is a collective,
        so must every other rank: the overflow flag is max-reduced over the group before anyone decides."""
        import torch
        import torch.distributed as dist
        from . import _native
        for _ in range(attempts):
            self.step(max_dist)
            st, over = None, 0
            try:
                st = self.ctx.sync()
            except _native.NativeError as e:
                if e.code != _native.BF_ERR_OVERFLOW:
                    raise
                over = 1
            if self.world > 1:
                dev = self.local.device if self.local is not None else torch.device("cuda", torch.cuda.current_device())
                flag = torch.tensor([over], dtype=torch.int32, device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
                over = int(flag.item())
            if not over:
                return st
        raise _native.NativeError(_native.BF_ERR_OVERFLOW, "buffer overflow persisted on some rank")


class ShardedCsrUploader:
    """End-to-end input path for N ranks: every rank copies only its 1/N slice of the column array
    from pinned host memory (the PCIe links work in parallel), the slices are all-gathered over
    NVLink, and the context adopts the gathered device CSR without another copy.

    Two device slots and a side stream: `prefetch()` enqueues copy + all-gather of the NEXT batch on
    the side stream while the current pass runs; `activate()` makes the compute stream wait for it and
    hands the slot to the context; `release()` (after the pass is enqueued) marks when the slot may be
    refilled.  Call order per step: activate, prefetch (next), run, release — the CSR all-gather is
    issued before the step's label all-gather so NCCL does not serialise it behind the pass."""

    def __init__(self, ctx, indptr, indices, n_cols: int, rank: int, world: int, group=None):
        import torch
        self.ctx, self.n_cols, self.rank, self.world, self.group = ctx, int(n_cols), rank, world, group
        self.n_rows, self.nnz = len(indptr) - 1, int(indices.size)
        chunk = max(1, -(-self.nnz // world))
        lo, hi = min(rank * chunk, self.nnz), min((rank + 1) * chunk, self.nnz)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.h_indptr = torch.from_numpy(indptr).pin_memory()
        self.h_shard = torch.zeros(chunk, dtype=torch.int32).pin_memory()
        self.h_shard[: hi - lo] = torch.from_numpy(indices[lo:hi])
        self.d_indptr = [torch.empty(self.n_rows + 1, dtype=torch.int64, device=dev) for _ in range(2)]
        self.d_shard = [torch.empty(chunk, dtype=torch.int32, device=dev) for _ in range(2)]
        self.d_full = [torch.empty(world * chunk, dtype=torch.int32, device=dev) if world > 1 else None for _ in range(2)]
        self.h2d_bytes = self.h_indptr.numel() * 8 + chunk * 4
        self.main = torch.cuda.current_stream()
        self.side = torch.cuda.Stream()
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [None, None]
        self.next_slot, self.active, self.pending = 0, None, None

    def prefetch(self):
        """copy + all-gather the next batch into the idle slot, on the side stream"""
        import torch
        import torch.distributed as dist
        s = self.next_slot
        with torch.cuda.stream(self.side):
            if self.free[s] is not None:
                self.side.wait_event(self.free[s])
            self.d_indptr[s].copy_(self.h_indptr, non_blocking=True)
            self.d_shard[s].copy_(self.h_shard, non_blocking=True)
            if self.world > 1:
                dist.all_gather_into_tensor(self.d_full[s], self.d_shard[s], group=self.group)
            self.ready[s].record(self.side)
        self.pending = s
        self.next_slot = s ^ 1

    def activate(self):
        """the compute stream waits for the prefetched batch; the context adopts it (no copy)"""
        s = self.pending
        if s is None:
            raise RuntimeError("ShardedCsrUploader.activate() without a prefetch(): nothing has been uploaded")
        self.pending = None
        self.main.wait_event(self.ready[s])
        full = self.d_full[s] if self.world > 1 else self.d_shard[s]
        self.ctx.adopt_csr_device(self.d_indptr[s].data_ptr(), full.data_ptr(), self.n_rows, self.n_cols, self.nnz)
        self.active = s

    def release(self):
        """call after the pass on the active slot has been enqueued"""
        import torch
        ev = torch.cuda.Event()
        ev.record(self.main)
        self.free[self.active] = ev

    def upload(self):
        """unpipelined convenience: prefetch + activate"""
        self.prefetch()
        self.activate()
