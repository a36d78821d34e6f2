"""Multi-GPU plumbing (one process per GPU, ``torch.distributed`` / NCCL over NVLink).

The replication loop needs no data-path collective: ranks replay disjoint ranges of the spawned
blocks.  Collectives appear only where partial statistics are combined (``_stats_device``) and
here, where every rank receives the full result vector the public API returns.
"""
from __future__ import annotations


def gather_results(local, group):
    """Concatenate the ranks' result shards in rank (= reference block) order on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    n_local = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    width = max(sizes)
    padded = torch.zeros(width, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    parts = [torch.empty(width, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:k] for p, k in zip(parts, sizes)])


def gather_results_to_root(local, group):
    """Rank 0 receives the concatenation (rank order); other ranks get ``None``."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_local = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    width = max(sizes)
    padded = torch.zeros(width, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    parts = [torch.empty(width, dtype=local.dtype, device=local.device) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, parts, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if rank != 0:
        return None
    return torch.cat([p[:k] for p, k in zip(parts, sizes)])
