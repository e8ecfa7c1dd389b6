"""The cross-attention scores product of the Q-Former alone (T = 40: batch 40 of [256, 5120] x [576, 5120]^T -> f32
[256, 576]), CUDA events, L2 flushed: 95 us = 634 TFLOP/s as 128 x 128 tiles (the padded 256-wide 2-CTA form was 111.6 us)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_zephyr_b200 import _lib
from vision_zephyr_b200.gemm import gemm

T = 40
lib = _lib.load()
qk = torch.randn((T, 256, 5120), device="cuda").to(torch.bfloat16)
f = torch.randn((T, 576, 5120), device="cuda").to(torch.bfloat16)
S = torch.empty((T, 256, 576), dtype=torch.float32, device="cuda")
sk = torch.zeros(lib.vz_gemm_sk_workspace_bytes(), dtype=torch.uint8, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def run():
    gemm(qk, f, M=256, N=576, K=5120, lda=5120, ldw=5120, out=S, ldo=576, batch=T, a_bstride=256 * 5120,
         w_bstride=576 * 5120, o_bstride=256 * 576, out_f32=True, sk_ws=sk)


ts = []
for i in range(13):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    if i >= 3:
        ts.append(e0.elapsed_time(e1))
ts.sort()
ref = torch.bmm(qk[:2].float(), f[:2].float().transpose(1, 2))
err = (S[:2] - ref).abs().max().item() / ref.abs().max().item()
med = ts[len(ts) // 2]
print(f"scores T={T}: {med * 1e3:.1f} us  {2.0 * T * 256 * 576 * 5120 / med / 1e9:.0f} TFLOP/s  rel err {err:.2e}")
