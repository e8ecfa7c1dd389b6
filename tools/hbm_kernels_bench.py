#!/usr/bin/env python
"""Achieved bandwidth of the HBM-side kernels on BASELINE configs (CUDA events, L2 flushed between runs):
   splice (config 5 geometry), preprocess (config 2: 16 x 336^2 with 3 overlays; config 3: 8 anyres images)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_zephyr_b200 as vz
from vision_zephyr_b200 import anyres, arch
from vision_zephyr_b200.preprocess import build_plan, run_plan

PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=20):
    ts = []
    for i in range(reps + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def splice_c5():
    rng = np.random.default_rng(5)
    B, S, D, V, Q = 8, 2048, 4096, 32000, 32
    ids = np.full((B, S), 2, np.int64); mask = np.zeros((B, S), np.uint8)
    for b in range(B):
        n = int(rng.integers(256, 2048)); ids[b, :n] = rng.integers(3, V, n); ids[b, int(rng.integers(1, 33))] = -200; mask[b, :n] = 1
    ids_d, mask_d = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    labels_d = ids_d.clone()
    emb = torch.randn((V, D), device="cuda").to(torch.bfloat16)
    vis = torch.randn((B * 5 * Q, D), device="cuda").to(torch.bfloat16)
    descs = [anyres.slot_descriptor(b * 5 * Q, 5, Q, "flat") for b in range(B)]
    slots, prefix, total = arch._slots_to_device(descs, "cuda")
    plan = arch.splice_plan(ids_d, mask_d, slots, B, 0)
    info = plan.wait()
    Lmax, lens = info["Lmax"], info["lengths"]
    lib = vz._lib.load()
    out = arch.splice_scatter(ids_d, labels_d, emb, vis, None, slots, prefix, B, total, plan, Lmax, False)
    fn = lambda: arch.splice_scatter(ids_d, labels_d, emb, vis, None, slots, prefix, B, total, plan, Lmax, False)
    t = timeit(fn)
    bytes_ = sum(lens) * 8192 + B * Lmax * 8192 + 17 * B * S + 24 * B * Lmax
    print(f"splice_scatter config5: Lmax={Lmax} mean len={sum(lens)/B:.0f}  {bytes_/1e6:.1f} MB in {t*1e6:.1f} us "
          f"= {bytes_/t/1e9:.0f} GB/s ({bytes_/t/1e9/PEAK:.2f} of measured HBM peak)  [includes torch.empty of outputs]")
    tp = timeit(lambda: arch.splice_plan(ids_d, mask_d, slots, B, 0))
    print(f"splice_plan   config5: {tp*1e6:.1f} us (latency-bound single CTA; {17*B*S/1e3:.0f} KB)")


def preprocess(cfg, batch=16):
    lut = vz.clip_lut()
    if cfg == 2:
        imgs = [torch.from_numpy(np.random.default_rng(i).integers(0, 256, (336, 336, 3), dtype=np.uint8)).cuda() for i in range(batch)]
        prompts = []
        for i in range(batch):
            lay = np.zeros((336, 336, 4), np.uint8); lay[50:200, 60:220] = (0, 255, 0, 128)
            prompts.append([vz.VisualPrompt("rectangle", rgba=(255, 0, 0, 128), bbox=(30, 40, 200, 220), width=3),
                            vz.VisualPrompt("layer", layer=lay), vz.VisualPrompt("layer", layer=lay[::-1].copy())])
        views = [anyres.single_view((336, 336)) for _ in imgs]
        plan = build_plan(imgs, views, lut, prompts)
        bytes_ = batch * (338688 + 2 * 451584 + 677376)
    else:
        sizes = [(1000, 900), (900, 1000), (1344, 1344), (700, 650), (1000, 900), (800, 760), (1200, 1100), (672, 672)]
        pins = [[336, 672], [672, 336], [672, 672], [336, 1008], [1008, 336]]
        imgs = [torch.from_numpy(np.random.default_rng(i).integers(0, 256, (h, w, 3), dtype=np.uint8)).cuda() for i, (w, h) in enumerate(sizes)]
        views = [anyres.anyres_views((w, h), pins)[0] for (w, h) in sizes]
        plan = build_plan(imgs, views, lut)
        bytes_ = sum(3 * w * h for w, h in sizes) + plan.n_tiles * 677376
    out = run_plan(plan, "patches")
    t = timeit(lambda: run_plan(plan, "patches", out))
    print(f"preprocess config{cfg}{'' if batch == 16 else f' (batch {batch})'}: {plan.n_tiles} tiles, {bytes_/1e6:.1f} MB algorithmic in {t*1e6:.1f} us = {bytes_/t/1e9:.0f} GB/s "
          f"({bytes_/t/1e9/PEAK:.3f} of measured HBM peak)")


splice_c5()
preprocess(2)
preprocess(2, batch=256)
preprocess(3)
