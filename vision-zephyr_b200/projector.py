"""Q-Former multimodal projector on the B200 kernels.

Drop-in for the reference's
  build_multimodal_projector / QFormer / QFormerBlock
  vis_zephyr/model/multimodal_projector/builder.py:12-101
The module owns parameters with EXACTLY the reference's state-dict keys (so `mm_projector.bin`
loads by name, vis_zephyr_arch.py:95-102) but its forward is `vz_qformer_forward`
(csrc/vz_model.cu): cross-attention reassociated so K and V are never materialised (batched tcgen05
GEMMs), fused 32-query self-attention kernels, and block-0 text conditioning without the dead rows.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib
from .projector_train import qformer_train_forward
from .vision_tower import PACK_GENERATION, GraphCache, Workspace

NUM_QUERIES, WIDTH, KV_WIDTH, HEADS, BLOCKS, FFN = 32, 4096, 5120, 8, 8, 8192


class _Norm(nn.Module):
    """parameter holder named like nn.LayerNorm (weight, bias)."""

    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))


class _Proj(nn.Module):
    """parameter holder named like nn.Linear (weight [out,in], bias), default nn.Linear init."""

    def __init__(self, din, dout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(dout, din))
        self.bias = nn.Parameter(torch.empty(dout))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1 / math.sqrt(din)
        nn.init.uniform_(self.bias, -bound, bound)


class _SelfAttnParams(nn.Module):
    """names of nn.MultiheadAttention with kdim == vdim == embed_dim."""

    def __init__(self, dim):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * dim, dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * dim))
        self.out_proj = _Proj(dim, dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)


class _CrossAttnParams(nn.Module):
    """names of nn.MultiheadAttention with kdim = vdim = 5120 != embed_dim."""

    def __init__(self, dim, kdim):
        super().__init__()
        self.q_proj_weight = nn.Parameter(torch.empty(dim, dim))
        self.k_proj_weight = nn.Parameter(torch.empty(dim, kdim))
        self.v_proj_weight = nn.Parameter(torch.empty(dim, kdim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * dim))
        self.out_proj = _Proj(dim, dim)
        for w in (self.q_proj_weight, self.k_proj_weight, self.v_proj_weight):
            nn.init.xavier_uniform_(w)
        nn.init.zeros_(self.out_proj.bias)


class _Block(nn.Module):
    def __init__(self, dim, kdim, ffn):
        super().__init__()
        self.norm1 = _Norm(dim)
        self.self_attn = _SelfAttnParams(dim)
        self.norm2 = _Norm(dim)
        self.cross_attn = _CrossAttnParams(dim, kdim)
        self.norm3 = _Norm(dim)
        # keys ffn.0.* and ffn.2.* (index 1 is the parameter-free GELU of the reference)
        self.ffn = nn.ModuleDict({"0": _Proj(dim, ffn), "2": _Proj(ffn, dim)})


@dataclass
class TextPack:
    """Text conditioning without duplication: packed non-image token embeddings of all samples
    (+ one trailing zero row), per-sample offsets, batch-global L and the tile->sample map."""
    text_emb: torch.Tensor       # bf16 [R+1, 4096]
    text_off: torch.Tensor       # int32 [B+1]
    text_rows: int               # R
    n_samples: int
    L: int
    tile_sample: torch.Tensor    # int32 [T]

    def dense(self) -> torch.Tensor:
        """[n_samples, L, 4096]: every sample's rows, zero padded to L (the reference's own layout,
        vis_zephyr_arch.py:178-189, before the per-tile expand).  Host round trip for the offsets."""
        off = self.text_off.tolist()
        out = torch.zeros((self.n_samples, self.L, self.text_emb.shape[1]), dtype=self.text_emb.dtype, device=self.text_emb.device)
        for s in range(self.n_samples):
            out[s, :off[s + 1] - off[s]] = self.text_emb[off[s]:off[s + 1]]
        return out


class QFormerB200(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.num_queries = NUM_QUERIES
        self.hidden_size = config.hidden_size
        if self.hidden_size != WIDTH:
            # the reference hard-codes 4096/5120 in the cross-attention (builder.py:19-25)
            raise ValueError("QFormer requires config.hidden_size == 4096")
        self.learned_queries = nn.Parameter(torch.randn(self.num_queries, self.hidden_size))
        self.blocks = nn.ModuleList([_Block(WIDTH, KV_WIDTH, FFN) for _ in range(BLOCKS)])
        self.pre_norm = _Norm(KV_WIDTH)
        self.norm = _Norm(WIDTH)
        self.force_simple_gemm = os.environ.get("VZ_FORCE_SIMPLE_GEMM") == "1"
        self._ws = Workspace()
        self._graphs = GraphCache()
        self._packed: Dict[str, torch.Tensor] = {}
        self._packed_key = None
        self._w: Optional[_lib.QfWeights] = None

    # -- weight packing -------------------------------------------------------------------------
    def _version_key(self):
        ps = list(self.parameters())
        return (tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps), ps[0].device, ps[0].dtype)

    def _ensure_packed(self):
        key = self._version_key()
        if key == self._packed_key:
            return
        dev = self.learned_queries.device
        bf = lambda t: t.detach().to(torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        P: Dict[str, torch.Tensor] = {}
        P["lq"] = bf(self.learned_queries)
        P["pre_g"], P["pre_b"] = f32(self.pre_norm.weight), f32(self.pre_norm.bias)
        P["norm_g"], P["norm_b"] = f32(self.norm.weight), f32(self.norm.bias)
        w = _lib.QfWeights()
        for i, blk in enumerate(self.blocks):
            ca, sa = blk.cross_attn, blk.self_attn
            # LayerNorm folding (csrc/vz_gemm.cu): LN(x) W^T + b = rstd (x W'^T - mu colsum(W')) + b'
            def fold(W, b, norm):
                W32, g32, beta32 = W.detach().float(), norm.weight.detach().float(), norm.bias.detach().float()
                Wf = (W32 * g32[None, :]).to(torch.bfloat16).contiguous()
                return Wf, (b.detach().float() + W32 @ beta32).contiguous(), Wf.float().sum(1).contiguous()
            if i == 0:   # block 0: the (few) query / text rows go through the LayerNorm kernel
                sa_w, sa_b, sa_s = bf(sa.in_proj_weight), f32(sa.in_proj_bias), None
            else:
                sa_w, sa_b, sa_s = fold(sa.in_proj_weight, sa.in_proj_bias, blk.norm1)
            q_w, q_b, q_s = fold(ca.q_proj_weight, ca.in_proj_bias[:WIDTH], blk.norm2)
            f1_w, f1_b, f1_s = fold(blk.ffn["0"].weight, blk.ffn["0"].bias, blk.norm3)
            ent = {
                "n1_g": f32(blk.norm1.weight), "n1_b": f32(blk.norm1.bias),
                "sa_in_w": sa_w, "sa_in_b": sa_b, "s_sa_in": sa_s,
                "sa_out_w": bf(sa.out_proj.weight), "sa_out_b": f32(sa.out_proj.bias),
                "ca_q_w": q_w, "ca_q_b": q_b, "s_ca_q": q_s, "ca_in_b": f32(ca.in_proj_bias),
                # K is never materialised: scores = (q Wk) f^T needs Wk^T with the head dim contiguous
                "ca_kT_w": bf(ca.k_proj_weight).t().contiguous(), "ca_v_w": bf(ca.v_proj_weight),
                "ca_out_w": bf(ca.out_proj.weight), "ca_out_b": f32(ca.out_proj.bias),
                "ffn1_w": f1_w, "ffn1_b": f1_b, "s_ffn1": f1_s,
                "ffn2_w": bf(blk.ffn["2"].weight), "ffn2_b": f32(blk.ffn["2"].bias),
            }
            for name, t in ent.items():
                P[f"{i}.{name}"] = t
                setattr(w.blocks[i], name, t.data_ptr() if t is not None else None)
        w.learned_queries = P["lq"].data_ptr()
        w.pre_g, w.pre_b = P["pre_g"].data_ptr(), P["pre_b"].data_ptr()
        w.norm_g, w.norm_b = P["norm_g"].data_ptr(), P["norm_b"].data_ptr()
        self._graphs.clear()                 # captured graphs hold the previous buffers' addresses
        self._pack_generation = next(PACK_GENERATION)
        P["pre_g"]._vz_generation = self._pack_generation   # read by the tower's graph key (pre_norm rides in its fusion kernel)
        self._packed, self._w, self._packed_key = P, w, key

    def pre_norm_params(self):
        """(gamma, beta) fp32 device tensors for fusing pre_norm into the tower's fusion kernel."""
        self._ensure_packed()
        return self._packed["pre_g"], self._packed["pre_b"]

    # -- compute --------------------------------------------------------------------------------
    def forward_packed(self, feats: torch.Tensor, text: Optional[TextPack], feats_normed: bool = False,
                       out: Optional[torch.Tensor] = None, graph: bool = False) -> torch.Tensor:
        """feats bf16 [T,576,5120] -> bf16 [T,32,4096].  graph=True (callers that consume the result at once):
        small batches replay a captured CUDA graph keyed by the text geometry (tiles, samples, rows, L)."""
        if (graph and out is None and feats.is_cuda and feats.dtype == torch.bfloat16 and not self.needs_autograd(feats)
                and GraphCache.usable(feats.shape[0])):
            self._ensure_packed()
            if text is None:
                return self._graphs.run(("qf", self._pack_generation, feats_normed), [feats.contiguous()],
                                        lambda f: self.forward_packed(f, None, feats_normed))
            key = ("qf", self._pack_generation, feats_normed, text.text_rows, text.n_samples, text.L)
            return self._graphs.run(
                key, [feats.contiguous(), text.text_emb, text.text_off, text.tile_sample],
                lambda f, e, o, ts: self.forward_packed(f, TextPack(e, o, text.text_rows, text.n_samples, text.L, ts), feats_normed))
        lib = _lib.load()
        if self.needs_autograd(feats):
            # the inference kernels run on detached, packed weights and keep no activations: with gradients on
            # (stage-1 / stage-2 training, train/train.py:817-836) take the autograd path instead
            if feats_normed:
                raise _lib.VzError("training path needs un-normalised features (pre_norm is a trained parameter): "
                                   "call the tower without pre_norm")
            dense, tile_sample = None, None
            if text is not None:
                dense, tile_sample = text.dense(), text.tile_sample
            res = qformer_train_forward(self, feats, dense, tile_sample)
            if out is not None:
                raise _lib.VzError("out= (peer-store transport) is an inference feature; training uses the autograd path")
            return res
        self._ensure_packed()
        if not feats.is_cuda:
            raise _lib.VzError("QFormerB200 runs on CUDA only (no CPU fallback)")
        if feats.dtype != torch.bfloat16:
            feats = feats.to(torch.bfloat16)
        feats = feats.contiguous()
        T = feats.shape[0]
        if feats.shape[1:] != (576, KV_WIDTH):
            raise ValueError(f"expected features [T,576,5120], got {tuple(feats.shape)}")
        dev = feats.device
        if out is None:
            out = torch.empty((T, NUM_QUERIES, WIDTH), dtype=torch.bfloat16, device=dev)
        ldo = out.stride(-2)
        n_s = text.n_samples if text is not None else 1
        rows = text.text_rows if text is not None else 0
        nbytes = lib.vz_qformer_workspace_bytes(T, n_s, rows)
        ws = self._ws.get(nbytes, dev)
        st = lib.vz_qformer_forward(
            C.byref(self._w), _lib.ptr(feats), 1 if feats_normed else 0, T,
            _lib.ptr(text.text_emb) if text is not None else None,
            _lib.ptr(text.text_off) if text is not None else None, rows, n_s,
            text.L if text is not None else 0,
            _lib.ptr(text.tile_sample) if text is not None else None,
            _lib.ptr(out), ldo, _lib.ptr(ws), ws.numel(), 1 if self.force_simple_gemm else 0,
            _lib.stream_ptr())
        _lib.check(st, "vz_qformer_forward")
        return out

    def needs_autograd(self, *inputs) -> bool:
        """True when a caller expects gradients through this forward (any parameter or input requires grad
        while grad mode is on): the training path (projector_train.py) is taken then."""
        if not torch.is_grad_enabled():
            return False
        return any(isinstance(t, torch.Tensor) and t.requires_grad for t in inputs) or \
            any(p.requires_grad for p in self.parameters())

    def forward(self, features, text_embeddings=None):
        """Reference signature (builder.py:72): features [T,576,5120], text_embeddings [T,L,4096]
        (already expanded per tile and zero padded) or None.  Every tile is treated as its own
        sample here; prepare_inputs_labels_for_multimodal uses forward_packed instead, which
        shares the text K/V between the tiles of a sample.  With gradients required it runs the autograd
        path (same math, tcgen05 GEMMs for every Linear and for the cross-attention, forward and backward)."""
        if self.needs_autograd(features, text_embeddings):
            if not features.is_cuda:
                raise _lib.VzError("QFormerB200 runs on CUDA only (no CPU fallback)")
            return qformer_train_forward(self, features, text_embeddings, None).to(features.dtype)
        in_dtype = features.dtype
        text = None
        if text_embeddings is not None:
            T, L = text_embeddings.shape[0], text_embeddings.shape[1]
            if T != features.shape[0]:
                raise RuntimeError("Sizes of tensors must match except in dimension 1")
            dev = features.device
            emb = torch.zeros((T * L + 1, WIDTH), dtype=torch.bfloat16, device=dev)
            emb[:T * L] = text_embeddings.reshape(T * L, WIDTH).to(torch.bfloat16)
            text = TextPack(text_emb=emb,
                            text_off=torch.arange(0, (T + 1) * L, max(L, 1), dtype=torch.int32, device=dev)[:T + 1]
                            if L > 0 else torch.zeros(T + 1, dtype=torch.int32, device=dev),
                            text_rows=T * L, n_samples=T, L=L,
                            tile_sample=torch.arange(T, dtype=torch.int32, device=dev))
        out = self.forward_packed(features, text, feats_normed=False)
        return out.to(in_dtype) if in_dtype != out.dtype and in_dtype.is_floating_point else out


def build_multimodal_projector(config, **kwargs):
    """multimodal_projector/builder.py:97-101: `mm_projector_type` is ignored, always the QFormer."""
    return QFormerB200(config)


# LLaVA-style alias named by BASELINE.json
build_vision_projector = build_multimodal_projector
