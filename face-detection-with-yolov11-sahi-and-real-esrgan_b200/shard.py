"""Image-index data parallelism: one process per GPU, image i -> rank i % world_size, no collective on
the hot path; ONE exchange at the end — a variable-length all-gather of detections for evaluation
(NCCL over NVLink on GPUs; the same code runs over gloo on CPU tensors in the world_size-2 tests)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world_size: int) -> List[int]:
    """Round-robin ownership: keeps load balanced when image sizes vary."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(range(rank, n_items, world_size))


def owner_of(index: int, world_size: int) -> int:
    return index % world_size


def gather_detections(image_ids: torch.Tensor, rows: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather per-rank detections.  image_ids [n] int64, rows [n, R] float32 (same device on every rank).
    Returns (ids [N_total], rows [N_total, R]) ordered by (image id, original order) — identical on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        order = torch.argsort(image_ids, stable=True)
        return image_ids[order], rows[order]
    world = dist.get_world_size(group)
    dev = rows.device
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    R = rows.shape[1]
    pad_rows = torch.zeros((cap, R), dtype=rows.dtype, device=dev)
    pad_ids = torch.full((cap,), -1, dtype=torch.int64, device=dev)
    pad_rows[: rows.shape[0]] = rows
    pad_ids[: rows.shape[0]] = image_ids
    all_rows = [torch.empty_like(pad_rows) for _ in range(world)]
    all_ids = [torch.empty_like(pad_ids) for _ in range(world)]
    dist.all_gather(all_rows, pad_rows, group=group)
    dist.all_gather(all_ids, pad_ids, group=group)
    ids = torch.cat([all_ids[r][: counts[r]] for r in range(world)])
    out = torch.cat([all_rows[r][: counts[r]] for r in range(world)])
    order = torch.argsort(ids, stable=True)
    return ids[order], out[order]
