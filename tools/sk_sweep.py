#!/usr/bin/env python
"""Stream-K GEMM against torch on the small-M shapes of a one-tile forward (debug sweep)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_zephyr_b200  # noqa
from vision_zephyr_b200 import _lib as L

lib = L.load()
ws = torch.zeros(lib.vz_gemm_sk_workspace_bytes(), dtype=torch.uint8, device="cuda")
torch.manual_seed(0)
for (M, N, K, res, act) in [(24, 8192, 4096, 0, 0), (32, 12288, 4096, 0, 0), (32, 4096, 4096, 1, 0), (32, 8192, 4096, 0, 2),
                            (32, 4096, 8192, 1, 0), (577, 3072, 1024, 0, 0), (577, 1024, 1024, 1, 0), (577, 4096, 1024, 0, 1),
                            (577, 1024, 4096, 1, 0), (576, 1024, 592, 0, 0), (256, 5120, 576, 0, 0), (32, 512, 5120, 0, 0)]:
    A = torch.randn((M, K), device="cuda").to(torch.bfloat16)
    W = (torch.randn((N, K), device="cuda") * K ** -0.5).to(torch.bfloat16)
    R = torch.randn((M, N), device="cuda").to(torch.bfloat16) if res else None
    ref = A.float() @ W.float().t()
    if act == 1:
        ref = ref * torch.sigmoid(1.702 * ref)
    elif act == 2:
        ref = torch.nn.functional.gelu(ref)
    if res:
        ref = ref + R.float()
    for sk in (0, 1):
        # uninitialised-looking output and scratch: NaN patterns must never leak into the result
        out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
        ws.view(torch.float32)[2048:].fill_(float("nan"))
        g = L.GemmArgs()
        g.A, g.W, g.out = A.data_ptr(), W.data_ptr(), out.data_ptr()
        g.residual = R.data_ptr() if res else None
        g.M, g.N, g.K, g.lda, g.ldw, g.ldo, g.ldr = M, N, K, K, K, N, N
        g.act = act
        if sk:
            g.sk_ws, g.sk_ws_bytes = ws.data_ptr(), ws.numel()
        L.check(lib.vz_gemm_bf16(C.byref(g), L.stream_ptr()), "gemm")
        torch.cuda.synchronize()
        err = (out.float() - ref).abs().max().item()
        print(f"M={M} N={N} K={K} res={res} act={act} sk={sk}: max err {err:.4g} nan={int(torch.isnan(out).sum())}")
