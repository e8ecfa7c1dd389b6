#!/usr/bin/env python
"""Time the tcgen05 GEMM on the path's shapes (CUDA events, L2 flushed between runs by rotating buffers)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_zephyr_b200  # noqa
from vision_zephyr_b200 import _lib as L

lib = L.load()
T = 40
M = T * 577
SHAPES = [("qkv", M, 3072, 1024, 0, 0, 1), ("o", M, 1024, 1024, 0, 1, 1), ("fc1", M, 4096, 1024, 1, 0, 1),
          ("fc2", M, 1024, 4096, 0, 1, 1), ("sa_in", 1280, 12288, 4096, 0, 0, 1), ("sa_out", 1280, 4096, 4096, 0, 1, 1),
          ("ffn1", 1280, 8192, 4096, 2, 0, 1), ("ffn2", 1280, 4096, 8192, 0, 1, 1), ("ca_q", 1280, 4096, 4096, 0, 0, 1)]


SK_WS = torch.zeros(lib.vz_gemm_sk_workspace_bytes(), dtype=torch.uint8, device="cuda")


def run(name, M, N, K, act, res, reps=20, ln=False, stats=False, sk=False):
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
    keep = []
    if ln:
        npp = 8
        st = torch.rand((M, npp, 2), device="cuda") + 1.0
        st[:, :, 1] += 200.0
        cs = torch.randn(N, device="cuda")
        keep += [st, cs]
        g.ln_stats, g.ln_colsum, g.ln_np, g.ln_eps = st.data_ptr(), cs.data_ptr(), npp, 1e-5
    if stats:
        npp = lib.vz_gemm_stats_partials(M, N)
        so = torch.empty((M, npp, 2), device="cuda")
        keep.append(so)
        g.stats_out, g.stats_np = so.data_ptr(), npp
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(reps + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.vz_gemm_bf16(C.byref(g), L.stream_ptr()), name)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    name = name + ("+ln" if ln else "") + ("+st" if stats else "") + ("+sk" if sk else "")
    print(f"{name:10s} M={M:6d} N={N:6d} K={K:5d} act={act} res={res}: {med * 1e3:8.1f} us  {2.0 * M * N * K / med / 1e9:7.1f} TFLOP/s")


if os.environ.get("VZ_BENCH_ONLY"):
    # one shape (an ncu target): VZ_BENCH_ONLY=o VZ_BENCH_STATS=1 VZ_BENCH_REPS=2
    sh = [x for x in SHAPES if x[0] == os.environ["VZ_BENCH_ONLY"]][0]
    run(*sh[:6], reps=int(os.environ.get("VZ_BENCH_REPS", "20")), stats=os.environ.get("VZ_BENCH_STATS") == "1",
        ln=os.environ.get("VZ_BENCH_LNC") == "1")
elif os.environ.get("VZ_BENCH_LN") == "1":
    for nm in ("qkv", "fc1"):
        sh = [x for x in SHAPES if x[0] == nm][0]
        run(*sh[:6]); run(*sh[:6], ln=True)
    for nm in ("o", "fc2"):
        sh = [x for x in SHAPES if x[0] == nm][0]
        run(*sh[:6]); run(*sh[:6], stats=True)
else:
    for s in SHAPES:
        run(*s[:6])
        run(*s[:6], sk=True)
