"""Tile-level data parallelism: units (tiles / (z, tile) pairs / images) are independent, so ranks take
static contiguous shards and no data-path collective is needed.  The only exchange micro-sam's algorithms ever need is
for stitched results: per-rank *instance tables* (box, score, area, tile id ...) are all-gathered so that cross-tile NMS /
painting (instance_segmentation.py:511-521, util.py:1750-1770) sees every instance, and per-slice id offsets
(multi_dimensional_segmentation.py:401-414) need an exclusive scan of per-slice max ids."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_units: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block partition [lo, hi) of n_units; sizes differ by at most one; covers every unit exactly once."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return (n_units * rank) // world_size, (n_units * (rank + 1)) // world_size


def all_gather_tables(table: torch.Tensor, group=None) -> Tuple[torch.Tensor, List[int]]:
    """All-gather row-tables of different lengths ([n_r, C] per rank, same C/dtype) -> ([sum n_r, C], counts).
    One all_gather of the counts, one all_gather of tables padded to the maximum count (NCCL over NVLink on GPU, gloo on
    CPU).  Without an initialised process group this is the identity."""
    if not (dist.is_available() and dist.is_initialized()):
        return table, [int(table.shape[0])]
    world = dist.get_world_size(group)
    n = torch.tensor([table.shape[0]], dtype=torch.int64, device=table.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts) if counts else 0
    padded = torch.zeros((mx,) + tuple(table.shape[1:]), dtype=table.dtype, device=table.device)
    padded[: table.shape[0]] = table
    parts = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0), counts


def allreduce_average_(tensors: List[torch.Tensor], group=None) -> int:
    """Data-parallel gradient exchange (training): average the given tensors over the ranks IN PLACE with one
    all-reduce of a flat buffer (NCCL over NVLink on GPUs, gloo on CPU); returns the number of elements.  torch DDP's semantics
    (micro_sam/training/training.py:train_sam).  Without an initialised process group / with one rank it only counts."""
    n = sum(int(t.numel()) for t in tensors)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1 or n == 0:
        return n
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()
    return n


def exclusive_id_offsets(local_max_ids: torch.Tensor, group=None) -> torch.Tensor:
    """Per-unit id offsets for slice-wise segmentation: exclusive scan over ALL units (in rank order) of their max ids.
    `local_max_ids` [n_r] int64 -> offsets [n_r] for this rank's units."""
    allv, counts = all_gather_tables(local_max_ids.reshape(-1, 1).to(torch.int64), group)
    allv = allv.reshape(-1)
    scan = torch.cumsum(allv, 0) - allv
    rank = dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0
    lo = sum(counts[:rank])
    return scan[lo: lo + local_max_ids.numel()]


def gather_instance_tables(tables: dict, group=None) -> Tuple[dict, List[int]]:
    """The exchange step of a multi-rank stitched result: every rank contributes the per-instance tensors of
    the tiles it owns (same keys / trailing shapes / dtypes on all ranks, first dimension = its number of instances) and
    receives the concatenation over ranks in rank order -- which is tile order, because ranks own contiguous tile ranges,
    so the gathered list is exactly the list a single process would have built."""
    out, counts = {}, None
    for k in sorted(tables):
        t = tables[k]
        width = 1
        for d in t.shape[1:]:
            width *= int(d)
        flat = t.reshape(t.shape[0], width)   # explicit width: reshape(0, -1) is ambiguous for a rank without instances
        g, c = all_gather_tables(flat.contiguous(), group)
        if counts is not None and c != counts:
            raise RuntimeError(f"gather_instance_tables: inconsistent row counts for {k}: {c} != {counts}")
        counts = c
        out[k] = g.reshape((g.shape[0],) + tuple(t.shape[1:]))
    return out, counts or []
