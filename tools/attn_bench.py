#!/usr/bin/env python
"""Time the ViT attention kernels alone (T = 40 tiles) and check them against fp32 torch attention."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_zephyr_b200  # noqa
from vision_zephyr_b200 import _lib as L

lib = L.load()
T = 40
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn((T * 577, 3072), generator=g, device="cuda")).to(torch.bfloat16)
qkv[:, :2048] *= 1.7
out = torch.empty((T * 577, 1024), dtype=torch.bfloat16, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for impl in (1,):   # (the mma.sync kernel moved to libvz_b200_testonly.so)
    ts = []
    for i in range(13):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.vz_vit_attention(L.ptr(qkv), L.ptr(out), T, impl, L.stream_ptr()), "attn")
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    q, k, v = (qkv[:3 * 577].float().view(3, 577, 3, 16, 64)[:, :, i].transpose(1, 2) for i in range(3))
    ref = (torch.softmax((q @ k.transpose(-1, -2)) * 0.125, dim=-1) @ v).transpose(1, 2).reshape(3 * 577, 1024)
    err = (out[:3 * 577].float() - ref).abs().max().item()
    med = ts[len(ts) // 2]
    print(f"impl={impl} T={T}: {med * 1e3:7.1f} us  "
          f"{4 * 577 * 577 * 64 * 16 * T / med / 1e9:6.1f} TFLOP/s  max_abs_err {err:.4g}")
