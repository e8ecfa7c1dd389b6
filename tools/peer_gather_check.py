#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/peer_gather_check.py: the sharded path with the peer-memory transport
(final LayerNorm stores straight into the destination rank's buffer) must give bit-identical embeddings to the
all-gather transport and to the unsharded path on one GPU; prints the per-step time of both transports."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_zephyr_b200 as vz
from vision_zephyr_b200.runtime import VisionEmbeddingPath, random_init_

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
path = random_init_(VisionEmbeddingPath(device=dev), seed=0)
lut = vz.clip_lut()
PIN = [[336, 672], [672, 336], [672, 672], [336, 1008], [1008, 336]]
n_img = 4 * world
sizes = [(1000, 900), (637, 336), (900, 1000), (336, 336)] * world     # 5, 3, 5, 3 tiles per rank
tiles = [5, 3, 5, 3] * world
imgs_all = [np.random.default_rng(100 + i).integers(0, 256, (h, w, 3), dtype=np.uint8) for i, (w, h) in enumerate(sizes)]
g = torch.Generator().manual_seed(1)
ids = torch.randint(3, 32000, (n_img, 48), generator=g)
ids[:, 7] = -200
ids = ids.to(dev)
lo, hi = rank * 4, rank * 4 + 4
mine = [torch.from_numpy(x).to(dev) for x in imgs_all[lo:hi]]


def run(peer):
    os.environ["VZ_PEER_GATHER"] = "1" if peer else "0"
    pb = vz.process_any_resolution_images(mine, PIN, lut, out_mode="patches")
    for _ in range(3):
        r = path.prepare_inputs_labels_for_multimodal_sharded(ids, None, None, None, None, pb, tiles, sizes)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        r = path.prepare_inputs_labels_for_multimodal_sharded(ids, None, None, None, None, pb, tiles, sizes)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    return r[4], e0.elapsed_time(e1) / 10


a, ta = run(True)
b, tb = run(False)
if rank == 0:
    from vision_zephyr_b200 import dist as vd
    used_peer = any(v not in (None, False) for v in vd._peer_cache.values())
    pb_all = vz.process_any_resolution_images([torch.from_numpy(x).to(dev) for x in imgs_all], PIN, lut, out_mode="patches")
    c = path.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, pb_all, sizes)[4]
    torch.cuda.synchronize()
    print(f"world {world}: peer transport in use: {used_peer}; peer == all_gather: {torch.equal(a, b)}; "
          f"sharded vs unsharded max diff {(a.float() - c.float()).abs().max().item():.4g}; "
          f"step {ta:.3f} ms (peer stores) vs {tb:.3f} ms (all_gather)")
    assert torch.equal(a, b)
    assert (a.float() - c.float()).abs().max().item() <= 0.1
dist.destroy_process_group()
