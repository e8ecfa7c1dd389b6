"""The two B200 mixins applied to HF Mistral: the LLM-side callers of the path.

This is what INTEGRATION.md section 1 produces when applied to the reference's
  VisZephyrConfig / VisZephyrModel / VisZephyrForCausalLM   vis_zephyr/model/language_model/vis_zephyr.py:19-170
i.e. the SAME class statements with the B200 mixins as bases.  The LLM body is HF Mistral, except that a
sequence-starting, gradient-free forward can run its decoder stack natively on packed rows (mistral_prefill.py,
config.vz_native_prefill); the hand-over is here: forward / generate call
prepare_inputs_labels_for_multimodal and pass the 6-tuple on (vis_zephyr.py:76-98, :124-142), and
generation re-attaches `images` / `images_size` to the step inputs (:144-168).
Used by tests/test_gpu_llm.py, bench.py's c5_prefill_b8 workload and tools/; save_mm_projector is the
pre-training checkpoint writer of train/vis_zephyr_trainer.py:326-343.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn
from transformers import MistralConfig, MistralForCausalLM, MistralModel

from .arch import VisZephyrB200MetaForCausalLM, VisZephyrB200MetaModel

_TUPLE_KEYS = ("input_ids", "position_ids", "attention_mask", "past_key_values", "inputs_embeds", "labels")


class VisZephyrB200Config(MistralConfig):
    model_type = "vis_zephyr_b200"


class VisZephyrB200Model(VisZephyrB200MetaModel, MistralModel):
    """HF MistralModel + the B200 mixin.  With `config.vz_native_prefill = True` a gradient-free forward that starts
    a sequence (no cached tokens yet) runs the decoder stack natively on packed rows (mistral_prefill.py: RMSNorm and
    SwiGLU fused into tcgen05 GEMMs, no pad rows) and fills the KV cache HF's decode steps continue from; every other
    call -- training, decode steps, requests for attentions / hidden states -- is HF Mistral unchanged."""
    config_class = VisZephyrB200Config

    def __init__(self, config: MistralConfig):
        super().__init__(config)
        self._vz_prefill = None
        self._vz_prefill_key = None

    def native_prefill(self):
        """The packed-row prefill engine over this model's current weights (re-folded when they change)."""
        from .mistral_prefill import MistralPrefillB200
        key = tuple((p.data_ptr(), p._version) for p in self.layers.parameters())
        if self._vz_prefill is None or self._vz_prefill_key != key:
            self._vz_prefill = MistralPrefillB200(self)
            self._vz_prefill_key = key
        return self._vz_prefill

    def _native_prefill_applies(self, inputs_embeds, attention_mask, past_key_values, kwargs) -> bool:
        if not getattr(self.config, "vz_native_prefill", False) or torch.is_grad_enabled():
            return False
        if inputs_embeds is None or inputs_embeds.dim() != 3 or not inputs_embeds.is_cuda:
            return False
        if past_key_values is not None and past_key_values.get_seq_length() != 0:
            return False
        if attention_mask is not None and (not isinstance(attention_mask, torch.Tensor) or attention_mask.dim() != 2
                                           or attention_mask.shape[1] != inputs_embeds.shape[1]):
            return False              # pre-built 4-D masks / mask dicts: HF's own path
        return not (kwargs.get("output_attentions") or kwargs.get("output_hidden_states"))

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds=None, use_cache=None, **kwargs):
        if (input_ids is None) == (inputs_embeds is None):
            raise ValueError("You must specify exactly one of input_ids or inputs_embeds")
        if getattr(self.config, "vz_native_prefill", False) and not torch.is_grad_enabled():
            embeds = inputs_embeds if inputs_embeds is not None else self.embed_tokens(input_ids)
            if self._native_prefill_applies(embeds, attention_mask, past_key_values, kwargs):
                from transformers import DynamicCache
                from transformers.modeling_outputs import BaseModelOutputWithPast
                use_cache = self.config.use_cache if use_cache is None else use_cache
                if use_cache and past_key_values is None:
                    past_key_values = DynamicCache(config=self.config)
                hidden = self.native_prefill().prefill(embeds, attention_mask, position_ids,
                                                       past_key_values if use_cache else None,
                                                       row_sumsq=getattr(embeds, "vz_row_sumsq", None))
                return BaseModelOutputWithPast(last_hidden_state=hidden,
                                               past_key_values=past_key_values if use_cache else None)
        return super().forward(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                               past_key_values=past_key_values, inputs_embeds=inputs_embeds, use_cache=use_cache,
                               **kwargs)


class VisZephyrB200ForCausalLM(MistralForCausalLM, VisZephyrB200MetaForCausalLM):
    config_class = VisZephyrB200Config

    def __init__(self, config):
        super(MistralForCausalLM, self).__init__(config)       # skip MistralForCausalLM's own body
        self.model = VisZephyrB200Model(config)
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.post_init()

    def get_model(self):
        return self.model

    def _multimodal(self, input_ids, position_ids, attention_mask, past_key_values, labels, images, images_size):
        out = self.prepare_inputs_labels_for_multimodal(input_ids, position_ids, attention_mask, past_key_values,
                                                        labels, images, images_size)
        return dict(zip(_TUPLE_KEYS, out))

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds=None, labels=None, use_cache=None, output_attentions=None,
                output_hidden_states=None, images=None, images_size=None, return_dict=None, **kwargs):
        args = dict(input_ids=input_ids, position_ids=position_ids, attention_mask=attention_mask,
                    past_key_values=past_key_values, inputs_embeds=inputs_embeds, labels=labels)
        if inputs_embeds is None:
            args = self._multimodal(input_ids, position_ids, attention_mask, past_key_values, labels, images, images_size)
        # extra keyword arguments (e.g. logits_to_keep) go through to HF Mistral; the reference drops them
        return super().forward(use_cache=use_cache, output_attentions=output_attentions,
                               output_hidden_states=output_hidden_states, return_dict=return_dict, **args, **kwargs)

    @torch.no_grad()
    def generate(self, input_ids=None, images=None, images_size=None, **kwargs):
        position_ids = kwargs.pop("position_ids", None)
        attention_mask = kwargs.pop("attention_mask", None)
        if "inputs_embeds" in kwargs:
            raise NotImplementedError("`inputs_embeds` is not supported in this generate function.")
        if images is not None:
            mm = self._multimodal(input_ids, position_ids, attention_mask, None, None, images, images_size)
            position_ids, attention_mask, inputs_embeds = mm["position_ids"], mm["attention_mask"], mm["inputs_embeds"]
        else:
            inputs_embeds = self.get_model().embed_tokens(input_ids)
        return super().generate(position_ids=position_ids, attention_mask=attention_mask,
                                inputs_embeds=inputs_embeds, **kwargs)

    def prepare_inputs_for_generation(self, input_ids, past_key_values=None, inputs_embeds=None, **kwargs):
        extra = {k: kwargs.pop(k, None) for k in ("images", "images_size")}
        inputs = super().prepare_inputs_for_generation(input_ids=input_ids, past_key_values=past_key_values,
                                                       inputs_embeds=inputs_embeds, **kwargs)
        inputs.update({k: v for k, v in extra.items() if v is not None})
        return inputs


def mm_adapter_state(model: nn.Module, keys_to_match=("mm_projector", "vision_resampler")) -> Dict[str, torch.Tensor]:
    """The tensors train/vis_zephyr_trainer.py:326-337 selects for mm_projector.bin: every named parameter
    whose name contains one of the keys, under its full name (e.g. 'model.mm_projector.blocks.0.norm1.weight')."""
    return {k: p.detach().cpu().clone() for k, p in model.named_parameters() if any(m in k for m in keys_to_match)}


def save_mm_projector(model: nn.Module, path: str) -> None:
    """Write an mm_projector.bin that both the reference (vis_zephyr_arch.py:95-102, model/builder.py:118-120)
    and initialize_vision_modules here load by name."""
    torch.save(mm_adapter_state(model), path)


def random_mistral_config(num_hidden_layers: int = 32, **over) -> VisZephyrB200Config:
    """Zephyr-7B-beta / Mistral-7B geometry (the shipped config.json:10-35) with the mm_* keys of the path."""
    from .runtime import DEFAULT_PINPOINTS
    kw = dict(hidden_size=4096, intermediate_size=14336, num_hidden_layers=num_hidden_layers, num_attention_heads=32,
              num_key_value_heads=8, vocab_size=32000, max_position_embeddings=32768, rms_norm_eps=1e-5,
              rope_theta=10000.0, sliding_window=None, pad_token_id=2, bos_token_id=1, eos_token_id=2,
              mm_vision_tower="openai/clip-vit-large-patch14-336", mm_vision_select_layer="-2,-5,-8,-11,6",
              mm_vision_select_feature="patch", mm_patch_merge_type="flat", image_aspect_ratio="anyres",
              mm_grid_pinpoints=DEFAULT_PINPOINTS, tokenizer_padding_side="right")
    kw.update(over)
    return VisZephyrB200Config(**kw)


class FirstLayerQKV:
    """SURVEY.md 8(f) rank 3, first half: "the first layer's RMSNorm + QKV fused with the splice output".

    The scatter hands over every output row's sum of squares (config.vz_first_layer_stats = True ->
    inputs_embeds.vz_row_sumsq, see vz_splice_scatter_rms); with the RMSNorm gain folded into the stacked
    q / k / v weight, ONE tcgen05 GEMM then computes rmsnorm(x) [Wq; Wk; Wv]^T: the fused-LayerNorm epilogue of
    vz_gemm_bf16 with a zero row sum is exactly rstd * (x W'^T), rstd = rsqrt(sum x^2 / K + eps).  No pass over the
    spliced rows for the statistic, no normalised copy of them.  (What consumes q / k / v -- RoPE and the causal GQA
    attention of language_model/vis_zephyr.py:86-98 -- stays HF Mistral: DESIGN.md section 7.)"""

    def __init__(self, layer):
        """layer: the HF MistralDecoderLayer whose input_layernorm / self_attn.{q,k,v}_proj are folded."""
        att, norm = layer.self_attn, layer.input_layernorm
        W = torch.cat([att.q_proj.weight, att.k_proj.weight, att.v_proj.weight], 0).detach().float()
        self.weight = (W * norm.weight.detach().float()[None, :]).to(torch.bfloat16).contiguous()       # [6144, 4096]
        self.colsum = self.weight.float().sum(1).contiguous()       # multiplied by a zero mean; kept valid
        self.bias = torch.zeros(self.weight.shape[0], dtype=torch.float32, device=self.weight.device)
        self.eps = float(getattr(norm, "variance_epsilon", getattr(norm, "eps", 1e-5)))
        self.splits = [att.q_proj.weight.shape[0], att.k_proj.weight.shape[0], att.v_proj.weight.shape[0]]

    def __call__(self, inputs_embeds: torch.Tensor, row_sumsq: Optional[torch.Tensor] = None):
        """inputs_embeds bf16 [B, L, 4096] (+ its row statistics from the scatter) -> q, k, v as the layer's own
        projections of input_layernorm(inputs_embeds) would give them, [B, L, n] each."""
        from .gemm import gemm
        stats = row_sumsq if row_sumsq is not None else getattr(inputs_embeds, "vz_row_sumsq", None)
        if stats is None:
            raise ValueError("row statistics missing: set config.vz_first_layer_stats = True before the splice")
        B, L, K = inputs_embeds.shape
        x = inputs_embeds.reshape(B * L, K)
        N = self.weight.shape[0]
        out = torch.empty((B * L, N), dtype=torch.bfloat16, device=x.device)
        gemm(x, self.weight, M=B * L, N=N, K=K, lda=K, ldw=K, out=out, ldo=N, bias=self.bias,
             ln_stats=stats.reshape(B * L, 2), ln_colsum=self.colsum, ln_np=1, ln_eps=self.eps)
        q, k, v = out.view(B, L, N).split(self.splits, dim=-1)
        return q, k, v
