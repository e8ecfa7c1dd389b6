"""Data-parallel sharding of images across the GPUs of one box + the single exchange step.

The reference has no communication on this path; north_star adds ONE collective:
an all-gather of the projected visual tokens onto the rank that holds the LLM batch.  Images (with
all their tiles) are assigned to ranks in contiguous blocks so every rank's output is a contiguous
run of visual rows in splice order; ranks are padded to the largest shard so a plain
all_gather_into_tensor (NCCL over NVLink on GPUs, gloo in the CPU tests) does the exchange.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_images(tiles_per_image: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) image ranges per rank, balanced by tile count (greedy on the prefix)."""
    n = len(tiles_per_image)
    total = sum(tiles_per_image)
    bounds, acc, start = [], 0, 0
    prefix = [0]
    for t in tiles_per_image:
        prefix.append(prefix[-1] + t)
    for r in range(world_size):
        if r == world_size - 1:
            end = n
        else:
            target = total * (r + 1) / world_size
            end = start
            while end < n and abs(prefix[end + 1] - target) <= abs(prefix[end] - target):
                end += 1
            # leave at least one image for each remaining rank when possible
            end = min(end, n - min(world_size - 1 - r, n - end) if n - end >= world_size - 1 - r else end)
        end = max(end, start)
        bounds.append((start, end))
        start = end
    return bounds


def gather_visual_tokens(local_rows: torch.Tensor, rows_per_rank: Sequence[int], group=None) -> torch.Tensor:
    """All-gather [rows_r, D] shards (rows_per_rank known on every rank from the host-side shard
    plan) into the full [sum rows, D] tensor, in rank order.  One collective, padded to max rows."""
    world = dist.get_world_size(group)
    D = local_rows.shape[1]
    max_rows = max(rows_per_rank)
    send = local_rows
    if local_rows.shape[0] != max_rows:
        send = torch.zeros((max_rows, D), dtype=local_rows.dtype, device=local_rows.device)
        send[:local_rows.shape[0]] = local_rows
    recv = torch.empty((world * max_rows, D), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if all(r == max_rows for r in rows_per_rank):
        return recv
    parts = [recv[r * max_rows:r * max_rows + rows_per_rank[r]] for r in range(world)]
    return torch.cat(parts, dim=0)
