#!/usr/bin/env python
"""Where the single-image call (BASELINE config 1) spends its time: every stage enqueued kernel by kernel vs
replayed from a CUDA graph, plus the whole public-API call.  Run with VZ_GEMM_SK=0|1, VZ_TMAP_CACHE=0|1."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_zephyr_b200 as vz
from vision_zephyr_b200.projector import TextPack
from vision_zephyr_b200.runtime import VisionEmbeddingPath, random_init_

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1
path = random_init_(VisionEmbeddingPath(device="cuda"), seed=0)
tower, proj = path.get_vision_tower(), path.get_model().mm_projector
lut = vz.clip_lut()
img = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (336, 336, 3), dtype=np.uint8)).cuda()
ids = torch.randint(3, 32000, (1, 64), generator=torch.Generator().manual_seed(1))
ids[0, 10] = -200
ids = ids.cuda()
patches = (torch.randn((T * 576, 592), device="cuda") * 0.5).to(torch.bfloat16)
pre = proj.pre_norm_params()
L = 63
text = TextPack((torch.randn((L + 1, 4096), device="cuda") * 0.02).to(torch.bfloat16), torch.tensor([0, L], dtype=torch.int32, device="cuda"),
                L, 1, L, torch.zeros(T, dtype=torch.int32, device="cuda"))
text.text_emb[-1] = 0


def timeit(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def graphed(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


feats = tower.encode_patches(patches, pre_norm=pre)
out = torch.empty((T, 32, 4096), dtype=torch.bfloat16, device="cuda")
stages = {"tower": lambda: tower.encode_patches(patches, pre_norm=pre),
          "qformer": lambda: proj.forward_packed(feats, text, feats_normed=True, out=out)}
print(f"T={T} VZ_GEMM_SK={os.environ.get('VZ_GEMM_SK', '1')} VZ_TMAP_CACHE={os.environ.get('VZ_TMAP_CACHE', '1')}")
for name, fn in stages.items():
    e = timeit(fn)
    g = graphed(fn)
    r = timeit(g.replay)
    print(f"  {name:8s}: eager {e:.3f} ms, graph replay {r:.3f} ms")


def call():
    pb = vz.process_fixed_images([img], lut, out_mode="patches")
    return path.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, pb, [(336, 336)])[4]


if T == 1:
    print(f"  preprocess alone: {timeit(lambda: vz.process_fixed_images([img], lut, out_mode='patches')):.3f} ms")
    print(f"  whole call (public API): {timeit(call):.3f} ms")
