"""Data-parallel sharding of images across the GPUs of one box + the single exchange step.

The reference has no communication on this path; north_star adds ONE exchange step:
the projected visual tokens of every rank must reach the rank that holds the LLM batch.  Images (with
all their tiles) are assigned to ranks in contiguous blocks so every rank's output is a contiguous
run of visual rows in splice order.

Two transports:
* PeerGather (default on CUDA when symmetric memory is available): the receive buffer of the
  destination rank is mapped into every rank's address space (torch symmetric memory = cuMem handles
  over NVLink / NVSwitch); each rank's Q-Former writes the rows of its final LayerNorm STRAIGHT into
  that buffer -- the stores of `layernorm_kernel` are the transfer -- followed by one device-side
  barrier.  No staging copy, no collective kernel, only the destination receives data.
* gather_visual_tokens: ranks padded to the largest shard + one all_gather_into_tensor (NCCL on GPUs,
  gloo in the CPU tests); used when peer mapping is unavailable or VZ_PEER_GATHER=0.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import os

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


class PeerGather:
    """Double-buffered symmetric receive buffer [2][rows_total, D] on every rank of `group`.

    Protocol per step: every rank asks `slot(dst, row_begin, n_rows)` for a view of the DESTINATION
    rank's buffer (a peer-mapped tensor), lets its projector write there, then calls `finish(dst)`:
    one device-side barrier on the current stream, after which rank `dst` may read `local(...)`.
    Two slots alternate so that a rank may start writing step s + 1 while `dst` still reads step s;
    the barrier of step s + 1 (which `dst` joins after its reads of step s were enqueued) protects the
    slot's next reuse at step s + 2."""

    def __init__(self, rows_total: int, width: int, dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.rows, self.width, self.dtype = rows_total, width, dtype
        self.group = group if group is not None else dist.group.WORLD
        self.buf = symm_mem.empty((2, rows_total, width), dtype=dtype, device=device)
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        self.step = 0

    def slot(self, dst: int, row_begin: int, n_rows: int) -> torch.Tensor:
        remote = self.handle.get_buffer(dst, (2, self.rows, self.width), self.dtype)
        return remote[self.step % 2, row_begin:row_begin + n_rows]

    def finish(self) -> None:
        self.handle.barrier()

    def local(self, n_rows: int) -> torch.Tensor:
        out = self.buf[self.step % 2, :n_rows]
        self.step += 1
        return out

    def skip_local(self) -> None:
        self.step += 1


_peer_cache = {}


def peer_gather_for(rows_total: int, width: int, dtype, device, group=None):
    """A PeerGather big enough for `rows_total` rows, or None when the peer transport is not usable
    (CPU / gloo group, symmetric memory missing, VZ_PEER_GATHER=0).  Collective: every rank of the
    group must call it with the same arguments (they derive them from the global shard plan)."""
    if os.environ.get("VZ_PEER_GATHER", "1") == "0" or torch.device(device).type != "cuda":
        return None
    key = (id(group), str(device), width, dtype)
    pg = _peer_cache.get(key)
    if pg is False:
        return None
    if pg is None or pg.rows < rows_total:
        try:
            pg = PeerGather(max(rows_total, pg.rows * 2 if pg else 0), width, dtype, device, group)
        except Exception as e:  # no P2P mapping on this box / group: fall back to the collective, once
            import warnings
            warnings.warn(f"vision_zephyr_b200: peer gather unavailable ({type(e).__name__}: {e}); using all_gather")
            _peer_cache[key] = False
            return None
        _peer_cache[key] = pg
    return pg
