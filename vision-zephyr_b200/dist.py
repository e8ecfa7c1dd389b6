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
    """Contiguous [begin, end) image ranges per rank that MINIMISE the largest tile count of any rank (the
    step time is set by the slowest rank).  No rank is left empty while there are at least as many images as
    ranks.  Binary search on the cap + a greedy left-to-right fill, O(n log sum)."""
    n = len(tiles_per_image)
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    tiles = [int(t) for t in tiles_per_image]

    def cut(cap):
        """greedy bounds under `cap` tiles per rank, keeping one image for every later rank; None if infeasible"""
        bounds, start = [], 0
        for r in range(world_size):
            later = world_size - 1 - r
            if r == world_size - 1:
                end = n
            else:
                limit = max(n - later, start) if n - start > later else start   # images this rank may take at most
                end, acc = start, 0
                while end < limit and acc + tiles[end] <= cap:
                    acc += tiles[end]
                    end += 1
                if end == start and n - start > later:        # must take one image but it does not fit the cap
                    return None
            if sum(tiles[start:end]) > cap:
                return None
            bounds.append((start, end))
            start = end
        return bounds

    lo, hi = max(tiles, default=0), max(sum(tiles), 0)
    while lo < hi:
        mid = (lo + hi) // 2
        if cut(mid) is None:
            lo = mid + 1
        else:
            hi = mid
    return cut(lo)


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
TRANSPORT_PEER, TRANSPORT_NCCL = "peer_store", "nccl_all_gather"


def peer_transport_requested() -> str:
    """VZ_PEER_GATHER: '1' = peer stores REQUIRED (fail loudly when unavailable), '0' = all-gather,
    unset / 'auto' = peer stores when every rank can set them up, else all-gather."""
    v = os.environ.get("VZ_PEER_GATHER", "auto").lower()
    return {"1": "require", "0": "off"}.get(v, "auto")


def peer_gather_for(rows_total: int, width: int, dtype, device, group=None):
    """A PeerGather big enough for `rows_total` rows, or None when the all-gather transport is to be used
    (CPU / gloo group, VZ_PEER_GATHER=0, or symmetric memory unavailable on ANY rank).  Collective: every
    rank of the group calls it with the same arguments (derived from the global shard plan).  The ranks
    AGREE on the outcome (all_reduce MIN of a success flag) before the decision is cached, so a failure on
    a subset of ranks can never leave the others waiting in a barrier.  With VZ_PEER_GATHER=1 an unavailable
    peer transport raises instead of falling back."""
    mode = peer_transport_requested()
    if mode == "off" or torch.device(device).type != "cuda":
        if mode == "require" and torch.device(device).type != "cuda":
            raise RuntimeError("VZ_PEER_GATHER=1 but the model is not on a CUDA device")
        return None
    key = (id(group), str(device), width, dtype)
    pg = _peer_cache.get(key)
    if pg is False:
        return None
    if pg is None or pg.rows < rows_total:
        new, err = None, None
        try:
            new = PeerGather(max(rows_total, pg.rows * 2 if pg else 0), width, dtype, device, group)
        except Exception as e:   # no P2P mapping on this box / group
            err = e
        ok = torch.tensor([1 if new is not None else 0], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            why = f"{type(err).__name__}: {err}" if err is not None else "another rank could not map the peer buffer"
            if mode == "require":
                raise RuntimeError(f"vision_zephyr_b200: VZ_PEER_GATHER=1 but the peer-store transport is unavailable ({why})")
            import warnings
            warnings.warn(f"vision_zephyr_b200: peer gather unavailable ({why}); every rank uses all_gather")
            _peer_cache[key] = False
            return None
        pg = new
        _peer_cache[key] = pg
    return pg
