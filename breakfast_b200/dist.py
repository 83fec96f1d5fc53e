"""Multi-GPU host logic: one process per GPU (torchrun), torch.distributed for the plumbing.

The band tile pairs of the (upper-triangular) tile space are dealt cyclically to the ranks inside
the pair kernel (work item w belongs to rank w % world, kernels.cuh k_pairs); every rank holds the
whole packed matrix and runs a local union-find over the edges it found.  The only exchange step is
the label merge: an all-gather of int32 labels[n_rows] per rank (NCCL over NVLink on GPUs) followed
by a device-side re-union (k_uf_merge_labels).  torch is used for the collective and for device
memory of the gathered buffer only.
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


class RankRunner:
    """Drives one rank's context through run -> all-gather -> merge.  The context must have been
    created on torch's current CUDA stream so that kernels and collectives are ordered."""

    def __init__(self, ctx, n_rows: int, rank: int, world: int, group=None):
        import torch
        self.ctx, self.n, self.rank, self.world, self.group = ctx, n_rows, rank, world, group
        dev = torch.device("cuda", torch.cuda.current_device())
        self.local = torch.empty(max(n_rows, 1), dtype=torch.int32, device=dev)
        self.gathered = torch.empty((world, max(n_rows, 1)), dtype=torch.int32, device=dev) if world > 1 else None

    def step(self, max_dist: int):
        """Enqueue one pass; returns nothing — call ctx.sync() for counters."""
        import torch.distributed as dist
        self.ctx.run(max_dist, self.rank, self.world)
        if self.world > 1:
            self.ctx.labels_to_device(self.local.data_ptr())
            dist.all_gather_into_tensor(self.gathered, self.local, group=self.group)
            self.ctx.merge_labels_device(self.gathered.data_ptr(), self.world)
