#!/usr/bin/env python
"""The GEMM on the single-image shapes (BASELINE config 1: M = 577 in the ViT, M = 32 in the Q-Former), with and
without the stream-K schedule, cold (L2 flushed) and warm.  The floor of a weight-streaming GEMM is W bytes / HBM."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_zephyr_b200  # noqa
from vision_zephyr_b200 import _lib as L

lib = L.load()
HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6546.6
SHAPES = [("vit qkv", 577, 3072, 1024, 0, 0), ("vit o", 577, 1024, 1024, 0, 1), ("vit fc1", 577, 4096, 1024, 1, 0),
          ("vit fc2", 577, 1024, 4096, 0, 1), ("qf sa_in", 32, 12288, 4096, 0, 0), ("qf sa_out", 32, 4096, 4096, 0, 1),
          ("qf ffn1", 32, 8192, 4096, 2, 0), ("qf ffn2", 32, 4096, 8192, 0, 1), ("qf kv_text", 64, 8192, 4096, 0, 0)]
SK_WS = torch.zeros(lib.vz_gemm_sk_workspace_bytes(), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(name, M, N, K, act, res, sk, cold, reps=30):
    A = torch.randn((M, K), device="cuda").to(torch.bfloat16)
    W = (torch.randn((N, K), device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    R = torch.randn((M, N), device="cuda").to(torch.bfloat16) if res else None
    out = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    g = L.GemmArgs()
    g.A, g.W, g.out, g.bias = A.data_ptr(), W.data_ptr(), out.data_ptr(), bias.data_ptr()
    g.residual = R.data_ptr() if res else None
    g.M, g.N, g.K, g.lda, g.ldw, g.ldo, g.ldr = M, N, K, K, K, N, N
    g.act = act
    if sk:
        g.sk_ws, g.sk_ws_bytes = SK_WS.data_ptr(), SK_WS.numel()
    ts = []
    for i in range(reps + 3):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.vz_gemm_bf16(C.byref(g), L.stream_ptr()), name)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


if len(sys.argv) > 1:      # one shape, stream-K on, a few launches: the target of an ncu capture
    sel = [x for x in SHAPES if x[0] == sys.argv[1]][0]
    print(sel, run(*sel, True, True, reps=3))
    sys.exit(0)
print(f"{'shape':12s} {'M':>4s} {'N':>6s} {'K':>5s}  cold: no-SK / SK (us)   warm: no-SK / SK (us)   W-stream floor (us)")
for (name, M, N, K, act, res) in SHAPES:
    r = [run(name, M, N, K, act, res, sk, cold) for cold in (True, False) for sk in (False, True)]
    print(f"{name:12s} {M:4d} {N:6d} {K:5d}  {r[0]:8.1f} / {r[1]:6.1f}        {r[2]:8.1f} / {r[3]:6.1f}        {N * K * 2 / HBM / 1e3:6.1f}")
