"""fp32 torch restatement of the tensor part of the path (test infrastructure; see oracle/__init__).

clip_hidden_states   HF CLIPVisionModel(..., output_hidden_states=True) as called from
                     vision_encoder/vision_encoder.py:101-105 (embeddings, pre_layrnorm, 24 pre-LN
                     layers with quick-GELU, no post_layernorm on hidden states)
fuse_features        vision_encoder.py:58-78 + gating_fusion/gating_fusion.py:22-50
qformer_forward      multimodal_projector/builder.py:34-44 (block) and :72-92 (model), LITERALLY:
                     block 0 runs on all 32+L rows and the first 32 are kept afterwards
text_embeddings_for  vis_zephyr_arch.py:157-192 (ids != IMAGE_TOKEN_INDEX incl. pads, expand per
                     tile, zero-pad to the batch max, cat)
"""
import math

import torch
import torch.nn.functional as F

IMAGE_TOKEN_INDEX = -200


def _ln(x, w, b, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def clip_hidden_states(sd, pixel_values, layers=24, heads=16):
    """pixel_values f32 [T,3,336,336] -> list of 25 hidden states [T,577,1024]."""
    p = "vision_model."
    T = pixel_values.shape[0]
    x = F.conv2d(pixel_values, sd[p + "embeddings.patch_embedding.weight"], stride=14)   # [T,1024,24,24]
    x = x.flatten(2).transpose(1, 2)
    cls = sd[p + "embeddings.class_embedding"].expand(T, 1, -1)
    x = torch.cat([cls, x], dim=1) + sd[p + "embeddings.position_embedding.weight"][None]
    x = _ln(x, sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"])
    hs = [x]
    D = x.shape[-1]
    hd = D // heads
    for l in range(layers):
        q_ = f"{p}encoder.layers.{l}."
        r = x
        h = _ln(x, sd[q_ + "layer_norm1.weight"], sd[q_ + "layer_norm1.bias"])
        q = F.linear(h, sd[q_ + "self_attn.q_proj.weight"], sd[q_ + "self_attn.q_proj.bias"])
        k = F.linear(h, sd[q_ + "self_attn.k_proj.weight"], sd[q_ + "self_attn.k_proj.bias"])
        v = F.linear(h, sd[q_ + "self_attn.v_proj.weight"], sd[q_ + "self_attn.v_proj.bias"])
        sh = lambda t: t.view(T, -1, heads, hd).transpose(1, 2)
        a = torch.softmax((sh(q) @ sh(k).transpose(-1, -2)) * (hd ** -0.5), dim=-1) @ sh(v)
        a = a.transpose(1, 2).reshape(T, -1, D)
        x = r + F.linear(a, sd[q_ + "self_attn.out_proj.weight"], sd[q_ + "self_attn.out_proj.bias"])
        r = x
        h = _ln(x, sd[q_ + "layer_norm2.weight"], sd[q_ + "layer_norm2.bias"])
        h = F.linear(h, sd[q_ + "mlp.fc1.weight"], sd[q_ + "mlp.fc1.bias"])
        h = h * torch.sigmoid(1.702 * h)
        x = r + F.linear(h, sd[q_ + "mlp.fc2.weight"], sd[q_ + "mlp.fc2.bias"])
        hs.append(x)
    return hs


def fuse_features(hidden_states, num_groups=4):
    """hidden_states[-21:], CLS dropped, 4 x mean of 5 consecutive layers + last, cat on channels."""
    sel = [h[:, 1:] for h in hidden_states[-(4 * 5 + 1):]]
    last, inter = sel[-1], sel[:-1]
    per = len(inter) // num_groups
    groups = [torch.stack(inter[i * per:(i + 1) * per], 0).mean(0) for i in range(num_groups)]
    return torch.cat(groups + [last], dim=-1)


def _mha(x_q, x_kv, wq, wk, wv, bq, bk, bv, wo, bo, heads=8):
    B, Nq, D = x_q.shape
    hd = D // heads
    q = F.linear(x_q, wq, bq).view(B, Nq, heads, hd).transpose(1, 2)
    k = F.linear(x_kv, wk, bk).view(B, -1, heads, hd).transpose(1, 2)
    v = F.linear(x_kv, wv, bv).view(B, -1, heads, hd).transpose(1, 2)
    a = torch.softmax((q * (hd ** -0.5)) @ k.transpose(-1, -2), dim=-1) @ v
    return F.linear(a.transpose(1, 2).reshape(B, Nq, D), wo, bo)


def qformer_block(sd, i, queries, feats):
    b = f"blocks.{i}."
    D = queries.shape[-1]
    q = _ln(queries, sd[b + "norm1.weight"], sd[b + "norm1.bias"])
    w, bias = sd[b + "self_attn.in_proj_weight"], sd[b + "self_attn.in_proj_bias"]
    queries = queries + _mha(q, q, w[:D], w[D:2 * D], w[2 * D:], bias[:D], bias[D:2 * D], bias[2 * D:],
                             sd[b + "self_attn.out_proj.weight"], sd[b + "self_attn.out_proj.bias"])
    q = _ln(queries, sd[b + "norm2.weight"], sd[b + "norm2.bias"])
    bias = sd[b + "cross_attn.in_proj_bias"]
    queries = queries + _mha(q, feats, sd[b + "cross_attn.q_proj_weight"], sd[b + "cross_attn.k_proj_weight"],
                             sd[b + "cross_attn.v_proj_weight"], bias[:D], bias[D:2 * D], bias[2 * D:],
                             sd[b + "cross_attn.out_proj.weight"], sd[b + "cross_attn.out_proj.bias"])
    q = _ln(queries, sd[b + "norm3.weight"], sd[b + "norm3.bias"])
    h = F.gelu(F.linear(q, sd[b + "ffn.0.weight"], sd[b + "ffn.0.bias"]))
    return queries + F.linear(h, sd[b + "ffn.2.weight"], sd[b + "ffn.2.bias"])


def qformer_forward(sd, features, text_embeddings=None, blocks=8, num_queries=32):
    """features [T,576,5120], text_embeddings [T,L,4096] or None -> [T,32,4096]."""
    T = features.shape[0]
    feats = _ln(features, sd["pre_norm.weight"], sd["pre_norm.bias"])
    queries = sd["learned_queries"].unsqueeze(0).expand(T, -1, -1)
    x = torch.cat([queries, text_embeddings], dim=1) if text_embeddings is not None else queries
    x = qformer_block(sd, 0, x, feats)
    queries = x[:, :num_queries, :]
    for i in range(1, blocks):
        queries = qformer_block(sd, i, queries, feats)
    return _ln(queries, sd["norm.weight"], sd["norm.bias"])


def text_embeddings_for(input_ids, tiles_per_image, embed):
    """[sum T_i, L, D]: per image i, embed(ids[i][ids[i] != -200]) expanded to its T_i tiles and
    right-padded with zeros to the longest sample."""
    rows = []
    for i, t in enumerate(tiles_per_image):
        ids = input_ids[i]
        e = embed[ids[ids != IMAGE_TOKEN_INDEX]]
        rows.append(e.unsqueeze(0).expand(t, -1, -1))
    L = max(r.shape[1] for r in rows)
    out = []
    for r in rows:
        if r.shape[1] < L:
            r = torch.cat([r, torch.zeros((r.shape[0], L - r.shape[1], r.shape[2]), dtype=r.dtype)], dim=1)
        out.append(r)
    return torch.cat(out, dim=0)


@torch.no_grad()
def encode_images(clip_sd, qf_sd, pixel_values, text_embeddings):
    """vis_zephyr_arch.py:120-124."""
    feats = fuse_features(clip_hidden_states(clip_sd, pixel_values))
    return qformer_forward(qf_sd, feats, text_embeddings)
