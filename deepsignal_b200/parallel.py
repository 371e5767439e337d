"""Read sharding across the GPUs of one box: reads are independent, weights are
replicated, there is NO collective on the data path — only a host-side gather of the result text
(``torch.distributed.gather_object`` over gloo/nccl's CPU side) that keeps each read's lines
contiguous and in input order, like the reference's single writer process
(call_modifications.py:283-297)."""
from __future__ import annotations

import os


def world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(n: int, rank: int, world_size: int):
    """Contiguous [lo, hi) of ``n`` reads for ``rank``; sizes differ by at most one."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def batch_owner(batch_index: int, world_size: int) -> int:
    """Round-robin owner of a features batch when the input is a stream (feature file)."""
    return batch_index % world_size


def gather_in_order(local_items, rank: int, world_size: int, group=None):
    """``local_items``: list of (global_index, payload).  Rank 0 returns all payloads sorted by
    global index, other ranks return None."""
    if world_size == 1:
        return [p for _, p in sorted(local_items, key=lambda t: t[0])]
    import torch.distributed as dist
    bucket = [None] * world_size if rank == 0 else None
    dist.gather_object(local_items, bucket, dst=0, group=group)
    if rank != 0:
        return None
    merged = [it for part in bucket for it in part]
    return [p for _, p in sorted(merged, key=lambda t: t[0])]
