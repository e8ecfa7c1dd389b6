"""Deterministic random-init weights with the reference's state-dict names (test infrastructure).

CLIP ViT-L/14-336 follows HF CLIPVisionModel naming (what CLIPVisionTower.load_model loads,
vision_encoder/vision_encoder.py:44-56) and HF's init scales; the projector follows QFormer's keys
(multimodal_projector/builder.py:49-70) and torch's default inits.  LayerNorm gains/biases and
Linear biases are perturbed so that a dropped gamma/beta/bias is visible in a parity test.
All values are rounded to bf16-representable floats, so the fp32 oracle and the bf16 CUDA path see
identical weights and parity tests measure compute error only.
"""
import math

import torch

CLIP_W, CLIP_L, CLIP_MLP, CLIP_TOK = 1024, 24, 4096, 577
QF_W, QF_KV, QF_FFN, QF_BLOCKS, QF_Q = 4096, 5120, 8192, 8, 32


def _bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _n(gen, shape, std):
    return _bf16r(torch.randn(shape, generator=gen) * std)


def _u(gen, shape, bound):
    return _bf16r((torch.rand(shape, generator=gen) * 2 - 1) * bound)


def clip_state_dict(seed: int = 0, layers: int = CLIP_L):
    g = torch.Generator().manual_seed(seed)
    p = "vision_model."
    sd = {}
    sd[p + "embeddings.class_embedding"] = _n(g, (CLIP_W,), CLIP_W ** -0.5)
    sd[p + "embeddings.patch_embedding.weight"] = _n(g, (CLIP_W, 3, 14, 14), 0.02)
    sd[p + "embeddings.position_embedding.weight"] = _n(g, (CLIP_TOK, CLIP_W), 0.02)
    sd[p + "pre_layrnorm.weight"] = _bf16r(1 + 0.1 * torch.randn(CLIP_W, generator=g))
    sd[p + "pre_layrnorm.bias"] = _n(g, (CLIP_W,), 0.1)
    in_std = (CLIP_W ** -0.5) * ((2 * CLIP_L) ** -0.5)
    out_std = CLIP_W ** -0.5
    fc_std = (2 * CLIP_W) ** -0.5
    for l in range(layers):
        q = f"{p}encoder.layers.{l}."
        for nm in ("q_proj", "k_proj", "v_proj"):
            # larger than HF's init so that attention is not uniform (exercises the softmax)
            sd[q + f"self_attn.{nm}.weight"] = _n(g, (CLIP_W, CLIP_W), 4 * in_std if nm != "v_proj" else in_std)
            sd[q + f"self_attn.{nm}.bias"] = _n(g, (CLIP_W,), 0.02)
        sd[q + "self_attn.out_proj.weight"] = _n(g, (CLIP_W, CLIP_W), out_std * 0.5)
        sd[q + "self_attn.out_proj.bias"] = _n(g, (CLIP_W,), 0.02)
        sd[q + "layer_norm1.weight"] = _bf16r(1 + 0.1 * torch.randn(CLIP_W, generator=g))
        sd[q + "layer_norm1.bias"] = _n(g, (CLIP_W,), 0.05)
        sd[q + "layer_norm2.weight"] = _bf16r(1 + 0.1 * torch.randn(CLIP_W, generator=g))
        sd[q + "layer_norm2.bias"] = _n(g, (CLIP_W,), 0.05)
        sd[q + "mlp.fc1.weight"] = _n(g, (CLIP_MLP, CLIP_W), fc_std)
        sd[q + "mlp.fc1.bias"] = _n(g, (CLIP_MLP,), 0.02)
        sd[q + "mlp.fc2.weight"] = _n(g, (CLIP_W, CLIP_MLP), in_std)
        sd[q + "mlp.fc2.bias"] = _n(g, (CLIP_W,), 0.02)
    sd[p + "post_layernorm.weight"] = torch.ones(CLIP_W)
    sd[p + "post_layernorm.bias"] = torch.zeros(CLIP_W)
    return sd


def qformer_state_dict(seed: int = 1, blocks: int = QF_BLOCKS):
    g = torch.Generator().manual_seed(seed)
    sd = {"learned_queries": _n(g, (QF_Q, QF_W), 1.0)}

    def ln(prefix, dim):
        sd[prefix + ".weight"] = _bf16r(1 + 0.1 * torch.randn(dim, generator=g))
        sd[prefix + ".bias"] = _n(g, (dim,), 0.05)

    def xavier(shape):
        return _u(g, shape, math.sqrt(6.0 / (shape[0] + shape[1])))

    def linear(prefix, dout, din):
        sd[prefix + ".weight"] = _u(g, (dout, din), 1 / math.sqrt(din))
        sd[prefix + ".bias"] = _u(g, (dout,), 1 / math.sqrt(din))

    for i in range(blocks):
        b = f"blocks.{i}."
        ln(b + "norm1", QF_W)
        sd[b + "self_attn.in_proj_weight"] = xavier((3 * QF_W, QF_W))
        sd[b + "self_attn.in_proj_bias"] = _n(g, (3 * QF_W,), 0.02)
        linear(b + "self_attn.out_proj", QF_W, QF_W)
        ln(b + "norm2", QF_W)
        sd[b + "cross_attn.q_proj_weight"] = xavier((QF_W, QF_W))
        sd[b + "cross_attn.k_proj_weight"] = xavier((QF_W, QF_KV))
        sd[b + "cross_attn.v_proj_weight"] = xavier((QF_W, QF_KV))
        sd[b + "cross_attn.in_proj_bias"] = _n(g, (3 * QF_W,), 0.02)
        linear(b + "cross_attn.out_proj", QF_W, QF_W)
        ln(b + "norm3", QF_W)
        linear(b + "ffn.0", QF_FFN, QF_W)
        linear(b + "ffn.2", QF_W, QF_FFN)
    ln("pre_norm", QF_KV)
    ln("norm", QF_W)
    return sd


def embed_table(seed: int = 2, vocab: int = 32000, dim: int = QF_W, std: float = 0.02):
    g = torch.Generator().manual_seed(seed)
    return _n(g, (vocab, dim), std)


def image_newline(seed: int = 3, dim: int = QF_W):
    g = torch.Generator().manual_seed(seed)
    return _n(g, (dim,), 1 / math.sqrt(dim))
