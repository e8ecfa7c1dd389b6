#!/usr/bin/env python
"""config-3 preprocess (8 anyres images -> 40 tiles) a few times: the target of ncu captures."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_zephyr_b200 as vz
from vision_zephyr_b200 import anyres
from vision_zephyr_b200.preprocess import build_plan, run_plan

sizes = [(1000, 900), (900, 1000), (1344, 1344), (700, 650), (1000, 900), (800, 760), (1200, 1100), (672, 672)]
pins = [[336, 672], [672, 336], [672, 672], [336, 1008], [1008, 336]]
imgs = [torch.from_numpy(np.random.default_rng(i).integers(0, 256, (h, w, 3), dtype=np.uint8)).cuda() for i, (w, h) in enumerate(sizes)]
plan = build_plan(imgs, [anyres.anyres_views((w, h), pins)[0] for (w, h) in sizes], vz.clip_lut())
out = run_plan(plan, "patches")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    flush.zero_()
    run_plan(plan, "patches", out)
torch.cuda.synchronize()
print("done", plan.n_tiles, plan.max_groups, plan.max_band_groups, plan.max_span128_px)
