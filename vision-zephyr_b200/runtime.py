"""A standalone host for the path (no LLM body): embedding table + tower + projector behind the
reference's mixin surface.  Used by bench.py, __graft_entry__.smoke() and the tests; a real
deployment mixes VisZephyrB200MetaForCausalLM into VisZephyrForCausalLM instead.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional

import torch
import torch.nn as nn

from .arch import VisZephyrB200MetaForCausalLM
from .projector import QFormerB200
from .vision_tower import CLIPVisionTowerB200

DEFAULT_PINPOINTS = "[[336, 672], [672, 336], [672, 672], [336, 1008], [1008, 336]]"


def default_config(**over):
    cfg = SimpleNamespace(hidden_size=4096, vocab_size=32000, mm_vision_tower="openai/clip-vit-large-patch14-336",
                          mm_vision_select_layer="-2,-5,-8,-11,6", mm_vision_select_feature="patch",
                          mm_patch_merge_type="flat", image_aspect_ratio="anyres",
                          mm_grid_pinpoints=DEFAULT_PINPOINTS, tokenizer_padding_side="right")
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


class _Inner(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.embed_tokens = nn.Embedding(config.vocab_size, config.hidden_size)
        self.vision_tower = CLIPVisionTowerB200(config.mm_vision_tower, config, delay_load=True)
        self.mm_projector = QFormerB200(config)
        if "unpad" in getattr(config, "mm_patch_merge_type", ""):
            self.image_newline = nn.Parameter(torch.zeros(config.hidden_size))

    def get_vision_tower(self):
        return self.vision_tower


class VisionEmbeddingPath(VisZephyrB200MetaForCausalLM):
    def __init__(self, config=None, device="cuda", dtype=torch.bfloat16):
        self.config = config or default_config()
        with torch.device("meta"):
            inner = _Inner(self.config)
        self.model = inner.to_empty(device=device)
        self.model.to(dtype)
        self.model.requires_grad_(False)      # an inference host: the projector's forward refuses to run with grads on
        self.device = torch.device(device)
        self.dtype = dtype

    def get_model(self):
        return self.model

    @torch.no_grad()
    def load_weights(self, clip_sd: Dict[str, torch.Tensor], qformer_sd: Dict[str, torch.Tensor],
                     embed: torch.Tensor, image_newline: Optional[torch.Tensor] = None):
        m = self.model
        m.vision_tower.load_model(state_dict=clip_sd, device=self.device)
        missing = m.mm_projector.load_state_dict({k: v.to(self.dtype) for k, v in qformer_sd.items()}, strict=True)
        m.embed_tokens.weight.copy_(embed.to(self.dtype))
        if image_newline is not None and hasattr(m, "image_newline"):
            m.image_newline.copy_(image_newline.to(self.dtype))
        return missing


@torch.no_grad()
def random_init_(path: VisionEmbeddingPath, seed: int = 0):
    """Random-init weights of the right architecture, generated ON THE DEVICE (benchmarks have no
    checkpoints; the tests use the CPU-seeded oracle weights instead).  Scales follow HF CLIP's
    initialiser and torch's defaults for the projector."""
    dev = path.device
    g = torch.Generator(device=dev).manual_seed(seed)
    W, L, MLP = 1024, 24, 4096

    def n(shape, std):
        return torch.randn(shape, generator=g, device=dev) * std

    p = "vision_model."
    sd = {p + "embeddings.class_embedding": n((W,), W ** -0.5),
          p + "embeddings.patch_embedding.weight": n((W, 3, 14, 14), 0.02),
          p + "embeddings.position_embedding.weight": n((577, W), 0.02),
          p + "pre_layrnorm.weight": 1 + n((W,), 0.1), p + "pre_layrnorm.bias": n((W,), 0.1)}
    in_std, out_std, fc_std = (W ** -0.5) * ((2 * L) ** -0.5), W ** -0.5, (2 * W) ** -0.5
    for l in range(L):
        q = f"{p}encoder.layers.{l}."
        for nm in ("q_proj", "k_proj", "v_proj"):
            sd[q + f"self_attn.{nm}.weight"] = n((W, W), 4 * in_std if nm != "v_proj" else in_std)
            sd[q + f"self_attn.{nm}.bias"] = n((W,), 0.02)
        sd[q + "self_attn.out_proj.weight"], sd[q + "self_attn.out_proj.bias"] = n((W, W), out_std * 0.5), n((W,), 0.02)
        for k in ("layer_norm1", "layer_norm2"):
            sd[q + k + ".weight"], sd[q + k + ".bias"] = 1 + n((W,), 0.1), n((W,), 0.05)
        sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"] = n((MLP, W), fc_std), n((MLP,), 0.02)
        sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"] = n((W, MLP), in_std), n((W,), 0.02)
    path.model.vision_tower.load_model(state_dict=sd, device=dev)
    del sd
    proj = path.model.mm_projector
    for name, prm in proj.named_parameters():
        if name == "learned_queries":
            prm.copy_(n(prm.shape, 1.0))
        elif name.endswith("norm.weight") or ".norm" in name and name.endswith("weight") or name.startswith("pre_norm.w"):
            prm.copy_(1 + n(prm.shape, 0.1))
        elif prm.dim() == 1:
            prm.copy_(n(prm.shape, 0.02))
        else:
            bound = (6.0 / (prm.shape[0] + prm.shape[1])) ** 0.5
            prm.copy_((torch.rand(prm.shape, generator=g, device=dev) * 2 - 1) * bound)
    path.model.embed_tokens.weight.copy_(n(path.model.embed_tokens.weight.shape, 0.02))
    if hasattr(path.model, "image_newline"):
        path.model.image_newline.copy_(n((path.config.hidden_size,), path.config.hidden_size ** -0.5))
    return path
