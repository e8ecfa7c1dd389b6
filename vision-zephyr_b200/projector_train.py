"""Training path of the Q-Former projector: forward WITH autograd.

The reference trains `mm_projector` alone in stage 1 and together with the LLM in stage 2
(train/train.py:817-836); its projector is nn.MultiheadAttention / nn.Linear / nn.LayerNorm under PyTorch
autograd (multimodal_projector/builder.py:12-92).  The inference kernels (vz_qformer_forward) work on packed,
folded weights and keep no activations, so when gradients are required QFormerB200.forward comes here instead:

  * every nn.Linear of the projector is a LinearFn: forward, dX and dW all run on the tcgen05 GEMM
    (vision-zephyr_b200/gemm.py), which is > 99 % of the FLOPs;
  * cross-attention is CrossAttnFn: the same exact reassociation as the inference path (K and V are never
    materialised: scores = (q Wk) f^T, out = (P f) Wv^T + bv), and its BACKWARD is reassociated too -- every
    weight-gradient contraction runs over the 32 T query rows, never over the 576 T patch rows
    (dWk_h = q_h^T (dS f), dWv_h = da_h^T (P f)); the key bias gets its exact gradient, zero;
  * LayerNorm, GELU, the 32-query self-attention core and the residual adds are ordinary PyTorch ops on
    [32 T, 4096] tensors, recorded by autograd.
Block 0 follows the reference literally for what matters: the learned queries attend to themselves plus the
sample's text rows plus the zero-padded tail (quirk Q3); the rows that the reference pushes through block 0
and then drops (builder.py:84-87) are not computed, since nothing depends on them.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

from .gemm import LinearFn, gemm

HEADS, HD, KV, Q = 8, 512, 5120, 32
_SCALE = 1.0 / math.sqrt(HD)


def _bf(t):
    t = t.detach().to(torch.bfloat16)
    return t if t.is_contiguous() else t.contiguous()


class CrossAttnFn(torch.autograd.Function):
    """q [M,4096] (M = 32 T, bias included), f [T,576,5120] (normalised features), Wk / Wv [4096,5120], bv [4096]
    -> attention output before out_proj, [M,4096]."""

    @staticmethod
    def forward(ctx, q, f, Wk, Wv, bv):
        dev = q.device
        qb, fb, Wkb, Wvb = _bf(q), _bf(f), _bf(Wk), _bf(Wv)
        M, D = qb.shape
        T, NP = fb.shape[0], fb.shape[1]
        HQ = HEADS * Q
        # qk[(t,i), h, :] = q[(t,i), h-slice] . Wk_h           batch = heads, Wk_h = rows h*512.. of Wk as [K=512, N=5120]
        qk = torch.empty((M, HEADS, KV), dtype=torch.bfloat16, device=dev)
        gemm(qb, Wkb, M=M, N=KV, K=HD, lda=D, ldw=KV, out=qk, ldo=HEADS * KV, w_is_kn=True, batch=HEADS,
             a_bstride=HD, w_bstride=HD * KV, o_bstride=KV)
        # S[t] = qk[t] (256 x 5120) . f[t]^T                     batch = tiles, fp32 out
        S = torch.empty((T, HQ, NP), dtype=torch.float32, device=dev)
        gemm(qk, fb, M=HQ, N=NP, K=KV, lda=KV, ldw=KV, out=S, ldo=NP, batch=T, a_bstride=HQ * KV, w_bstride=NP * KV,
             o_bstride=HQ * NP, out_f32=True)
        P = torch.softmax(S * _SCALE, dim=-1).to(torch.bfloat16)
        # PF[t] = P[t] (256 x 576) . f[t]                        f read as [K=576, N=5120]
        PF = torch.empty((T, HQ, KV), dtype=torch.bfloat16, device=dev)
        gemm(P, fb, M=HQ, N=KV, K=NP, lda=NP, ldw=KV, out=PF, ldo=KV, w_is_kn=True, batch=T, a_bstride=HQ * NP,
             w_bstride=NP * KV, o_bstride=HQ * KV)
        # a[(t,i), h-slice] = PF[(t,i), h, :] . Wv_h^T + bv_h     batch = heads
        a = torch.empty((M, D), dtype=torch.bfloat16, device=dev)
        bvf = bv.detach().float().contiguous()
        gemm(PF, Wvb, M=M, N=HD, K=KV, lda=HEADS * KV, ldw=KV, out=a, ldo=D, bias=bvf, batch=HEADS, a_bstride=KV,
             w_bstride=HD * KV, o_bstride=HD, bias_bstride=HD)
        ctx.save_for_backward(qb, fb, Wkb, Wvb, qk, P, PF)
        ctx.dts = (q.dtype, f.dtype, Wk.dtype, Wv.dtype, bv.dtype)
        return a.to(q.dtype)

    @staticmethod
    def backward(ctx, da):
        qb, fb, Wkb, Wvb, qk, P, PF = ctx.saved_tensors
        qdt, fdt, wkdt, wvdt, bvdt = ctx.dts
        dev = qb.device
        M, D = qb.shape
        T, NP = fb.shape[0], fb.shape[1]
        HQ = HEADS * Q
        da = da.to(torch.bfloat16).contiguous()
        dbv = da.float().sum(0).to(bvdt) if ctx.needs_input_grad[4] else None
        # dPF[(t,i), h, :] = da[(t,i), h-slice] . Wv_h            Wv_h read as [K=512, N=5120]
        dPF = torch.empty((M, HEADS, KV), dtype=torch.bfloat16, device=dev)
        gemm(da, Wvb, M=M, N=KV, K=HD, lda=D, ldw=KV, out=dPF, ldo=HEADS * KV, w_is_kn=True, batch=HEADS,
             a_bstride=HD, w_bstride=HD * KV, o_bstride=KV)
        dWv = None
        if ctx.needs_input_grad[3]:
            # dWv_h = da_h^T (512 x M) . PF_h (M x 5120): the contraction runs over the query rows only
            daT = da.t().contiguous()                              # [4096, M]
            dWv = torch.empty((D, KV), dtype=torch.bfloat16, device=dev)
            gemm(daT, PF, M=HD, N=KV, K=M, lda=M, ldw=HEADS * KV, out=dWv, ldo=KV, w_is_kn=True, batch=HEADS,
                 a_bstride=HD * M, w_bstride=KV, o_bstride=HD * KV)
        # dP[t] = dPF[t] . f[t]^T, softmax backward in fp32
        dP = torch.empty((T, HQ, NP), dtype=torch.float32, device=dev)
        gemm(dPF, fb, M=HQ, N=NP, K=KV, lda=KV, ldw=KV, out=dP, ldo=NP, batch=T, a_bstride=HQ * KV, w_bstride=NP * KV,
             o_bstride=HQ * NP, out_f32=True)
        Pf = P.float()
        dS = (Pf * (dP - (dP * Pf).sum(-1, keepdim=True)) * _SCALE).to(torch.bfloat16)
        del dP, Pf
        # dqk[t] = dS[t] (256 x 576) . f[t]
        dqk = torch.empty((T, HQ, KV), dtype=torch.bfloat16, device=dev)
        gemm(dS, fb, M=HQ, N=KV, K=NP, lda=NP, ldw=KV, out=dqk, ldo=KV, w_is_kn=True, batch=T, a_bstride=HQ * NP,
             w_bstride=NP * KV, o_bstride=HQ * KV)
        dq = None
        if ctx.needs_input_grad[0]:
            # dq[(t,i), h-slice] = dqk[(t,i), h, :] . Wk_h^T
            dq = torch.empty((M, D), dtype=torch.bfloat16, device=dev)
            gemm(dqk, Wkb, M=M, N=HD, K=KV, lda=HEADS * KV, ldw=KV, out=dq, ldo=D, batch=HEADS, a_bstride=KV,
                 w_bstride=HD * KV, o_bstride=HD)
            dq = dq.to(qdt)
        dWk = None
        if ctx.needs_input_grad[2]:
            qT = qb.t().contiguous()                               # [4096, M]
            dWk = torch.empty((D, KV), dtype=torch.bfloat16, device=dev)
            gemm(qT, dqk, M=HD, N=KV, K=M, lda=M, ldw=HEADS * KV, out=dWk, ldo=KV, w_is_kn=True, batch=HEADS,
                 a_bstride=HD * M, w_bstride=KV, o_bstride=HD * KV)
            dWk = dWk.to(wkdt)
        df = None
        if ctx.needs_input_grad[1]:
            # df[t] = P[t]^T dPF[t] + dS[t]^T qk[t]  as ONE product over the concatenated contraction (2 x 256)
            At = torch.cat([P.transpose(1, 2), dS.transpose(1, 2)], dim=2).contiguous()       # [T,576,512]
            Bt = torch.cat([dPF.view(T, HQ, KV), qk.view(T, HQ, KV)], dim=1).contiguous()      # [T,512,5120]
            df = torch.empty((T, NP, KV), dtype=torch.bfloat16, device=dev)
            gemm(At, Bt, M=NP, N=KV, K=2 * HQ, lda=2 * HQ, ldw=KV, out=df, ldo=KV, w_is_kn=True, batch=T,
                 a_bstride=NP * 2 * HQ, w_bstride=2 * HQ * KV, o_bstride=NP * KV)
            df = df.to(fdt)
        return dq, df, dWk, (dWv.to(wvdt) if dWv is not None else None), dbv


def _ln(x, norm):
    return F.layer_norm(x, (x.shape[-1],), norm.weight, norm.bias, 1e-5)


def _heads(t, n):
    """[n*rows, 4096] -> [n, 8, rows, 512]"""
    return t.view(n, -1, HEADS, HD).transpose(1, 2)


def _self_attention(q, k, v):
    """softmax(q k^T / sqrt(512)) v on [n, 8, rows, 512] tensors, fp32 softmax (tiny: 32 query rows)"""
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * _SCALE
    return torch.matmul(torch.softmax(s, dim=-1), v.float()).to(q.dtype)


def qformer_train_forward(mod, features: torch.Tensor, text: Optional[torch.Tensor] = None,
                          tile_sample: Optional[torch.Tensor] = None) -> torch.Tensor:
    """features [T,576,5120]; text [n_samples, L, 4096] (zero padded to the batch-global L, quirk Q3) or None;
    tile_sample int64 [T] maps tiles to samples (default: every tile its own sample, the reference signature).
    Returns [T,32,4096] in the projector's dtype, connected to autograd."""
    T = features.shape[0]
    D = mod.hidden_size
    dt = mod.learned_queries.dtype
    f = _ln(features.to(dt), mod.pre_norm)                                     # pre_norm (builder.py:74)
    b0 = mod.blocks[0]
    sa = b0.self_attn
    lq = mod.learned_queries
    # ---- block 0 self-attention: the queries are the same for every tile, the keys differ per sample ----
    qkv0 = LinearFn.apply(_ln(lq, b0.norm1), sa.in_proj_weight, sa.in_proj_bias)           # [32, 12288]
    q0, k0, v0 = qkv0[:, :D], qkv0[:, D:2 * D], qkv0[:, 2 * D:]
    if text is not None:
        n_s, L = text.shape[0], text.shape[1]
        kv = LinearFn.apply(_ln(text.to(dt).reshape(n_s * L, D), b0.norm1), sa.in_proj_weight[D:], sa.in_proj_bias[D:])
        kt, vt = kv[:, :D].view(n_s, L, D), kv[:, D:].view(n_s, L, D)
        K = torch.cat([k0.unsqueeze(0).expand(n_s, -1, -1), kt], dim=1)
        V = torch.cat([v0.unsqueeze(0).expand(n_s, -1, -1), vt], dim=1)
        qh = _heads(q0.unsqueeze(0).expand(n_s, -1, -1).reshape(n_s * Q, D), n_s)
        att = _self_attention(qh, K.view(n_s, Q + L, HEADS, HD).transpose(1, 2), V.view(n_s, Q + L, HEADS, HD).transpose(1, 2))
        att = att.transpose(1, 2).reshape(n_s * Q, D)
        x1 = lq.repeat(n_s, 1) + LinearFn.apply(att, sa.out_proj.weight, sa.out_proj.bias)
        if tile_sample is None:
            if n_s != T:
                raise RuntimeError("Sizes of tensors must match except in dimension 1")
            x = x1
        else:
            x = x1.view(n_s, Q, D)[tile_sample.long()].reshape(T * Q, D)
    else:
        att = _self_attention(_heads(q0, 1), _heads(k0, 1), _heads(v0, 1)).transpose(1, 2).reshape(Q, D)
        x1 = lq + LinearFn.apply(att, sa.out_proj.weight, sa.out_proj.bias)
        x = x1.repeat(T, 1)
    for i, blk in enumerate(mod.blocks):
        if i > 0:
            sa = blk.self_attn
            qkv = LinearFn.apply(_ln(x, blk.norm1), sa.in_proj_weight, sa.in_proj_bias)
            att = _self_attention(_heads(qkv[:, :D], T), _heads(qkv[:, D:2 * D], T), _heads(qkv[:, 2 * D:], T))
            x = x + LinearFn.apply(att.transpose(1, 2).reshape(T * Q, D), sa.out_proj.weight, sa.out_proj.bias)
        ca = blk.cross_attn
        q = LinearFn.apply(_ln(x, blk.norm2), ca.q_proj_weight, ca.in_proj_bias[:D])
        a = CrossAttnFn.apply(q, f, ca.k_proj_weight, ca.v_proj_weight, ca.in_proj_bias[2 * D:])
        x = x + LinearFn.apply(a, ca.out_proj.weight, ca.out_proj.bias)
        h = F.gelu(LinearFn.apply(_ln(x, blk.norm3), blk.ffn["0"].weight, blk.ffn["0"].bias))
        x = x + LinearFn.apply(h, blk.ffn["2"].weight, blk.ffn["2"].bias)
    return _ln(x, mod.norm).view(T, Q, D)
