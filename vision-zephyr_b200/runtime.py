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
