#!/usr/bin/env python
"""How much of a step is launch gaps?  Times the ViT + fusion (T = 40) enqueued kernel by kernel against a
replay of the same launches captured in a CUDA graph (torch.cuda.CUDAGraph around vz_vit_forward)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_zephyr_b200  # noqa
from vision_zephyr_b200.runtime import VisionEmbeddingPath, random_init_

T = int(sys.argv[1]) if len(sys.argv) > 1 else 40
path = random_init_(VisionEmbeddingPath(device="cuda"), seed=0)
tower = path.get_vision_tower()
proj = path.get_model().mm_projector
patches = (torch.randn((T * 576, 592), device="cuda") * 0.5).to(torch.bfloat16)
pre = proj.pre_norm_params()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


eager = timeit(lambda: tower.encode_patches(patches, pre_norm=pre))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    tower.encode_patches(patches, pre_norm=pre)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    out = tower.encode_patches(patches, pre_norm=pre)
graphed = timeit(g.replay)
ref = tower.encode_patches(patches, pre_norm=pre)
g.replay()
torch.cuda.synchronize()
print(f"T={T}: ViT + fusion eager {eager:.3f} ms, CUDA-graph replay {graphed:.3f} ms "
      f"({100 * (eager - graphed) / eager:.1f} % of the eager time is launch gaps); outputs equal: {torch.equal(out, ref)}")
