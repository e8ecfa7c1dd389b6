#!/usr/bin/env python
"""Single-image latency (BASELINE config 1: one 336x336 image, 64-token prompt) through the public API,
inputs resident on the GPU; CUDA events over 30 calls after warm-up."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_zephyr_b200 as vz
from vision_zephyr_b200.runtime import VisionEmbeddingPath, random_init_

path = random_init_(VisionEmbeddingPath(device="cuda"), seed=0)
lut = vz.clip_lut()
img = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (336, 336, 3), dtype=np.uint8)).cuda()
ids = torch.randint(3, 32000, (1, 64), generator=torch.Generator().manual_seed(1))
ids[0, 10] = -200
ids = ids.cuda()


def call():
    pb = vz.process_fixed_images([img], lut, out_mode="patches")
    return path.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, pb, [(336, 336)])[4]


for _ in range(5):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30):
    out = call()
e1.record()
torch.cuda.synchronize()
print(f"config 1 (1 tile) VZ_GEMM_SK={os.environ.get('VZ_GEMM_SK', '1')}: {e0.elapsed_time(e1) / 30:.3f} ms per call, output {tuple(out.shape)}")
