#!/usr/bin/env python
"""Tiny instance of every kernel of the library, for `compute-sanitizer --tool memcheck` (SURVEY.md section 5:
the reference has no sanitizer runs; this is the B200 equivalent).  Random weights, T = 2 tiles."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_zephyr_b200 as vz
from vision_zephyr_b200 import anyres, arch
from vision_zephyr_b200.runtime import VisionEmbeddingPath, default_config, random_init_

PINS = [[336, 672], [672, 336], [672, 672], [336, 1008], [1008, 336]]
path = random_init_(VisionEmbeddingPath(default_config(mm_patch_merge_type="spatial_unpad"), device="cuda"), 0)
lut = vz.clip_lut()
rng = np.random.default_rng(0)
imgs = [torch.from_numpy(rng.integers(0, 256, (180, 250, 3), dtype=np.uint8)).cuda(),
        torch.from_numpy(rng.integers(0, 256, (336, 336, 3), dtype=np.uint8)).cuda()]
layer = np.zeros((180, 250, 4), np.uint8); layer[20:90, 30:200] = (1, 2, 3, 99)
prompts = [[vz.VisualPrompt("layer", layer=layer), vz.VisualPrompt("rectangle", rgba=(9, 8, 7, 128), bbox=(-5, 3, 300, 170), width=4)], []]
pb = vz.process_any_resolution_images(imgs, PINS, lut, prompts=prompts, out_mode="patches")
chw = vz.process_any_resolution_images(imgs, PINS, lut, prompts=prompts, out_mode="chw")
print("tiles", pb.tiles_per_image)
# single tile per image so the 'spatial_unpad' merge is the executable single-tile form (quirk Q2)
one = [c[:1] for c in chw]
ids = torch.randint(3, 32000, (3, 33), generator=torch.Generator().manual_seed(0)).cuda()
ids[0, 4] = -200; ids[1, 0] = -200
mask = torch.ones_like(ids); mask[1, 20:] = 0
out = path.prepare_inputs_labels_for_multimodal(ids[:2], None, mask[:2], None, ids[:2].clone(), one, [(250, 180), (336, 336)])
print("unpad single-tile", tuple(out[4].shape))
path.config.mm_patch_merge_type = "flat"
path.config.tokenizer_padding_side = "left"
path.config.tokenizer_model_max_length = 150
out = path.prepare_inputs_labels_for_multimodal(ids[:2], None, mask[:2], None, None, pb, [(250, 180), (336, 336)])
print("flat anyres left-pad truncated", tuple(out[4].shape))
feat = torch.randn((5 * 576, 4096), device="cuda").to(torch.bfloat16)
d = anyres.slot_descriptor(0, 5, 576, "spatial_unpad", "anyres", (1000, 900), str(PINS), 336, 24)
print("merge rows", arch.merge_rows(feat, torch.zeros(4096, device="cuda", dtype=torch.bfloat16), [d])[0].shape)
torch.cuda.synchronize()
print("sanitize smoke done")

# the prefill behind the splice: one decoder layer, ragged left-padded batch, KV cache filled
from transformers import DynamicCache, MistralConfig, MistralModel
from vision_zephyr_b200.mistral_prefill import MistralPrefillB200
cfg = MistralConfig(hidden_size=4096, intermediate_size=1024, num_hidden_layers=1, num_attention_heads=32,
                    num_key_value_heads=8, vocab_size=64, rms_norm_eps=1e-5, sliding_window=None)
_old = torch.get_default_dtype()
torch.set_default_dtype(torch.bfloat16)
with torch.device("cuda"):
    llm = MistralModel(cfg)
torch.set_default_dtype(_old)
llm.eval().requires_grad_(False)
x = torch.randn((3, 140, 4096), device="cuda").to(torch.bfloat16)
m = torch.zeros((3, 140), dtype=torch.long, device="cuda")
for b, n in enumerate((140, 1, 77)):
    m[b, 140 - n:] = 1
with torch.no_grad():
    cache = DynamicCache(config=cfg)
    h = MistralPrefillB200(llm).prefill(x, m, None, cache)
print("native prefill", tuple(h.shape), "cache", tuple(cache.layers[0].keys.shape), bool(torch.isfinite(h.float()).all()))
