#!/usr/bin/env python
"""The 'second bar' of BASELINE.md section 4: the reference's modules are plain PyTorch, so on a B200 the
reference IS PyTorch eager in bf16 (cuBLAS + ATen).  This times that formulation (oracle/model.py run on
the GPU in bf16, same shapes as bench.py: 40 tiles, 63 text tokens) next to nothing else.  Test/bench
infrastructure only."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_zephyr_b200  # noqa
from vision_zephyr_b200.runtime import VisionEmbeddingPath, random_init_
from oracle import model as M

dev = "cuda"
path = random_init_(VisionEmbeddingPath(device=dev), 0)
# HF-named weights for the eager formulation, taken from the same random init
g = torch.Generator(device=dev).manual_seed(0)
T, L = 40, 63
P = path.model.vision_tower._packed
clip = {"vision_model.embeddings.class_embedding": P["class_emb"],
        "vision_model.embeddings.patch_embedding.weight": P["patch_w"][:, :588].reshape(1024, 3, 14, 14).contiguous(),
        "vision_model.embeddings.position_embedding.weight": P["pos_emb"],
        "vision_model.pre_layrnorm.weight": P["pre_ln_g"].bfloat16(), "vision_model.pre_layrnorm.bias": P["pre_ln_b"].bfloat16()}
for l in range(24):
    q = f"vision_model.encoder.layers.{l}."
    wq, wk, wv = P[f"{l}.w_qkv"].split(1024, 0)
    bq, bk, bv = P[f"{l}.b_qkv"].bfloat16().split(1024, 0)
    clip.update({q + "self_attn.q_proj.weight": wq, q + "self_attn.k_proj.weight": wk, q + "self_attn.v_proj.weight": wv,
                 q + "self_attn.q_proj.bias": bq, q + "self_attn.k_proj.bias": bk, q + "self_attn.v_proj.bias": bv,
                 q + "self_attn.out_proj.weight": P[f"{l}.w_o"], q + "self_attn.out_proj.bias": P[f"{l}.b_o"].bfloat16(),
                 # (the packed weights have the LayerNorm affine folded in; timing only needs the shapes)
                 q + "layer_norm1.weight": torch.ones(1024, device=dev, dtype=torch.bfloat16),
                 q + "layer_norm1.bias": torch.zeros(1024, device=dev, dtype=torch.bfloat16),
                 q + "layer_norm2.weight": torch.ones(1024, device=dev, dtype=torch.bfloat16),
                 q + "layer_norm2.bias": torch.zeros(1024, device=dev, dtype=torch.bfloat16),
                 q + "mlp.fc1.weight": P[f"{l}.w_fc1"], q + "mlp.fc1.bias": P[f"{l}.b_fc1"].bfloat16(),
                 q + "mlp.fc2.weight": P[f"{l}.w_fc2"], q + "mlp.fc2.bias": P[f"{l}.b_fc2"].bfloat16()})
qf = {k: v.detach() for k, v in path.model.mm_projector.state_dict().items()}
px = torch.randn((T, 3, 336, 336), device=dev, dtype=torch.bfloat16)
text = (torch.randn((T, L, 4096), device=dev) * 0.02).to(torch.bfloat16)


def step():
    with torch.no_grad():
        return M.encode_images(clip, qf, px, text)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"PyTorch-eager bf16 (reference formulation, ViT + fusion + Q-Former, 40 tiles, L=63): {ms:.2f} ms/step "
      f"= {8 / ms * 1e3:.1f} images/s")
