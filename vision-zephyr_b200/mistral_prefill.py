"""Native prefill of the Zephyr / Mistral decoder stack on the rows the splice produces.

The reference hands `inputs_embeds` to HF Mistral (language_model/vis_zephyr.py:86-98; with
train/zephyr_flash_attn_monkey_patch.py:86-136 the attention runs on unpadded rows).  Here the whole prefill runs
on PACKED rows -- only the real tokens of every sample, in order -- and a decoder layer is

    q|k|v   = rmsnorm(h) [Wq; Wk; Wv]^T        ONE tcgen05 GEMM: the RMSNorm gain is folded into the stacked weight
                                               and 1/rms is applied in the epilogue from per-row sums of squares
                                               (layer 0: handed over by the splice's scatter, vz_splice_scatter_rms;
                                               later layers: written by the previous down_proj epilogue)
    rope(q, k)                                 vz_rope_apply, in place on the packed q|k|v rows
    a       = causal GQA attention             vz_attn_causal: tcgen05 flash attention over the packed rows (S, P, O in
                                               tensor memory, 128-query x 64-key blocks up to the diagonal, work items
                                               sorted by length); VZ_LLM_ATTN=fa2 switches to flash-attn 2 varlen
                                               (library code) as a cross-check
    h       = h + a Wo^T          (+ stats)    GEMM, residual and next RMSNorm's row statistics in the epilogue
    m       = silu(g) * u                      ONE GEMM over gate / up rows interleaved in blocks of 64: the SwiGLU
                                               product is taken in the epilogue, [M, 2 I] never exists
    h       = h + m Wd^T          (+ stats)    GEMM, as above

i.e. four GEMM launches + one row kernel + one attention launch per layer, no normalised copy of the hidden state,
no pad rows.  The final RMSNorm, lm_head and the loss stay HF's (MistralForCausalLM.forward), fed with the padded
last_hidden_state this module returns; the KV cache is filled for HF's decode steps.
"""
from __future__ import annotations

import ctypes as C
import itertools
import os
from typing import List, Optional

import numpy as np

import torch

from . import _lib
from .gemm import gemm

ACT_SWIGLU = 3


def interleave_gate_up(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """[I, K], [I, K] -> [2 I, K]: blocks of 64 gate rows followed by the 64 matching up rows (VZ_ACT_SWIGLU)."""
    I, K = gate.shape
    if I % 64:
        raise ValueError(f"intermediate size {I} is not a multiple of 64")
    return torch.stack([gate.reshape(I // 64, 64, K), up.reshape(I // 64, 64, K)], dim=1).reshape(2 * I, K).contiguous()


class _Layer:
    __slots__ = ("w_qkv", "w_o", "w_gu", "w_d")


class MistralPrefillB200:
    """Packed-row prefill of an HF MistralModel's decoder stack (weights folded and re-laid-out once)."""

    def __init__(self, model):
        """model: HF MistralModel (or any object with .layers / .norm / .rotary_emb / .config of that shape)."""
        cfg = model.config
        self.hidden = cfg.hidden_size
        self.n_heads = cfg.num_attention_heads
        self.n_kv = cfg.num_key_value_heads
        self.head_dim = getattr(cfg, "head_dim", None) or self.hidden // self.n_heads
        self.inter = cfg.intermediate_size
        self.eps = float(cfg.rms_norm_eps)
        self.sliding_window = getattr(cfg, "sliding_window", None)
        if self.head_dim % 16 or self.hidden % 64 or self.inter % 64:
            raise ValueError("MistralPrefillB200: head_dim % 16, hidden % 64 and intermediate % 64 must be 0")
        self.q_cols, self.kv_cols = self.n_heads * self.head_dim, self.n_kv * self.head_dim
        self.qkv_cols = self.q_cols + 2 * self.kv_cols
        self.layers: List[_Layer] = []
        dev = None
        with torch.no_grad():
            for lyr in model.layers[: cfg.num_hidden_layers]:
                att, mlp = lyr.self_attn, lyr.mlp
                for lin in (att.q_proj, att.k_proj, att.v_proj, att.o_proj, mlp.gate_proj, mlp.up_proj, mlp.down_proj):
                    if lin.bias is not None:
                        raise ValueError("MistralPrefillB200: biased projections are not supported")
                g1 = lyr.input_layernorm.weight.detach().float()[None, :]
                g2 = lyr.post_attention_layernorm.weight.detach().float()[None, :]
                L = _Layer()
                wqkv = torch.cat([att.q_proj.weight, att.k_proj.weight, att.v_proj.weight], 0).detach().float()
                L.w_qkv = (wqkv * g1).to(torch.bfloat16).contiguous()
                L.w_o = att.o_proj.weight.detach().to(torch.bfloat16).contiguous()
                L.w_gu = interleave_gate_up((mlp.gate_proj.weight.detach().float() * g2).to(torch.bfloat16),
                                            (mlp.up_proj.weight.detach().float() * g2).to(torch.bfloat16))
                L.w_d = mlp.down_proj.weight.detach().to(torch.bfloat16).contiguous()
                self.layers.append(L)
                dev = L.w_qkv.device
        if dev is None or dev.type != "cuda":
            raise _lib.VzError("MistralPrefillB200 needs the model on a CUDA device (there is no CPU path)")
        self.device = dev
        self.inv_freq = model.rotary_emb.inv_freq.detach().to(device=dev, dtype=torch.float32).contiguous()
        if self.inv_freq.numel() != self.head_dim // 2:
            raise ValueError("rotary inv_freq does not match head_dim / 2")
        self.attention_scaling = float(getattr(model.rotary_emb, "attention_scaling", 1.0))
        if self.attention_scaling != 1.0:
            raise ValueError("scaled rotary embeddings are not supported")
        self.final_norm = model.norm
        self._ws = {}
        self.attn_impl = "fa2" if os.environ.get("VZ_LLM_ATTN", "") == "fa2" else "native"
        if self.attn_impl == "native" and self.head_dim != 128:
            raise ValueError("vz_attn_causal is built for head_dim 128 (set VZ_LLM_ATTN=fa2 for other geometries)")
        _lib.load()

    # ------------------------------------------------------------------------------------------
    def _buffers(self, M: int):
        key = torch.cuda.current_stream().cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws["cap"] < M:
            cap = max(M, 256)
            dev, bf = self.device, torch.bfloat16
            ws = dict(cap=cap,
                      qkv=torch.empty((cap, self.qkv_cols), dtype=bf, device=dev),
                      act=torch.empty((cap, self.inter), dtype=bf, device=dev),
                      attn=torch.empty((cap, self.q_cols), dtype=bf, device=dev),
                      h=[torch.empty((cap, self.hidden), dtype=bf, device=dev) for _ in range(2)],
                      stats=torch.empty((cap, self.hidden // 64, 2), dtype=torch.float32, device=dev),
                      cs=torch.empty((cap, self.head_dim // 2, 2), dtype=torch.float32, device=dev),
                      # stream-K scratch of the GEMMs: exclusive to one stream at a time (vz_b200.h), so per stream too
                      sk=torch.empty(_lib.load().vz_gemm_sk_workspace_bytes(), dtype=torch.uint8, device=dev))
            self._ws[key] = ws
        return ws

    def attention_items(self, lens):
        """device work list of vz_attn_causal for samples of `lens` rows: (items int32 [n, 4], n, algorithmic FLOPs)"""
        lib = _lib.load()
        lens_h = np.ascontiguousarray(lens, dtype=np.int32)
        n = lib.vz_attn_causal_items(lens_h.ctypes.data, len(lens_h), self.n_heads, None, 0, None)
        if n <= 0:
            raise _lib.VzError(f"vz_attn_causal_items: {n}")
        items_h = np.empty((n, 4), dtype=np.int32)
        flops = C.c_double(0.0)
        got = lib.vz_attn_causal_items(lens_h.ctypes.data, len(lens_h), self.n_heads, items_h.ctypes.data, n, C.byref(flops))
        if got != n:
            raise _lib.VzError(f"vz_attn_causal_items: {got} != {n}")
        return _lib.h2d(items_h, self.device), n, flops.value

    def forward_packed(self, x: torch.Tensor, positions: torch.Tensor, cu_seqlens: torch.Tensor, max_seqlen: int,
                       row_sumsq: Optional[torch.Tensor] = None, kv_sink=None, lens=None) -> torch.Tensor:
        """x bf16 [M, hidden] (packed real rows), positions int32 [M], cu_seqlens int32 [B + 1] (device);
        lens = the same lengths on the host (read back from cu_seqlens when not given).
        row_sumsq f32 [M, 2] = (anything, sum of squares) per row, or None (computed here).
        kv_sink(layer_idx, k [M, n_kv * hd] post-rope view, v view): called per layer before the buffers are reused.
        Returns the last decoder layer's output (BEFORE the final norm), bf16 [M, hidden] -- a workspace view."""
        lib = _lib.load()
        st = _lib.stream_ptr()
        M, H = x.shape
        if H != self.hidden or x.dtype != torch.bfloat16 or not x.is_contiguous():
            raise ValueError("forward_packed: x must be contiguous bf16 [M, hidden]")
        ws = self._buffers(M)
        qkv, act, (h_a, h_b), S, cs, sk = ws["qkv"], ws["act"], ws["h"], ws["stats"], ws["cs"], ws["sk"]
        # partial row statistics per output row of a [M, hidden] GEMM; every launch is ordered on one stream, so ONE
        # buffer serves all producers (a consumer has finished before the next producer starts)
        np_h = lib.vz_gemm_stats_partials(M, H)
        if row_sumsq is None:
            _lib.check(lib.vz_row_stats(x.data_ptr(), H, M, H, S.data_ptr(), st), "vz_row_stats")
            stats, np_in = S, 1
        else:
            if row_sumsq.shape != (M, 2) or row_sumsq.dtype != torch.float32 or not row_sumsq.is_contiguous():
                raise ValueError("row_sumsq must be contiguous f32 [M, 2]")
            stats, np_in = row_sumsq, 1
        _lib.check(lib.vz_rope_table(positions.data_ptr(), M, self.inv_freq.data_ptr(), self.head_dim // 2,
                                     cs.data_ptr(), st), "vz_rope_table")
        window = (-1, -1)
        if self.sliding_window is not None and max_seqlen > self.sliding_window:
            window = (int(self.sliding_window) - 1, 0)
        native_attn = self.attn_impl == "native" and window == (-1, -1)
        if native_attn:
            if lens is None:
                cu_h = cu_seqlens.cpu().tolist()
                lens = [b - a for a, b in zip(cu_h[:-1], cu_h[1:])]
            if sum(lens) != M:
                raise ValueError("forward_packed: lens do not add up to the packed row count")
            items, n_items, attn_flops = self.attention_items(lens)
            attn_out = ws["attn"]
            scale = float(self.head_dim) ** -0.5
        else:
            from flash_attn import flash_attn_varlen_func      # library attention (cross-check / sliding window)
        h = x
        q_v = qkv[:M, : self.q_cols].view(M, self.n_heads, self.head_dim)
        k_v = qkv[:M, self.q_cols: self.q_cols + self.kv_cols].view(M, self.n_kv, self.head_dim)
        v_v = qkv[:M, self.q_cols + self.kv_cols:].view(M, self.n_kv, self.head_dim)
        for li, L in enumerate(self.layers):
            # q | k | v = rmsnorm(h) W'^T
            gemm(h, L.w_qkv, M=M, N=self.qkv_cols, K=H, lda=H, ldw=H, out=qkv, ldo=self.qkv_cols,
                 ln_stats=stats, ln_np=np_in, ln_eps=self.eps, ln_rms=True, sk_ws=sk)
            _lib.check(lib.vz_rope_apply(qkv.data_ptr(), self.qkv_cols, M, self.n_heads + self.n_kv, self.head_dim,
                                         cs.data_ptr(), st), "vz_rope_apply")
            if kv_sink is not None:
                kv_sink(li, qkv[:M, self.q_cols: self.q_cols + self.kv_cols], qkv[:M, self.q_cols + self.kv_cols:])
            if native_attn:
                _lib.check(lib.vz_attn_causal(qkv.data_ptr(), self.qkv_cols, M, attn_out.data_ptr(), self.q_cols,
                                              items.data_ptr(), n_items, self.n_heads, self.n_kv, self.head_dim,
                                              scale, attn_flops, st), "vz_attn_causal")
                a = attn_out
            else:
                a = flash_attn_varlen_func(q_v, k_v, v_v, cu_seqlens, cu_seqlens, max_seqlen, max_seqlen, causal=True,
                                           window_size=window).view(M, self.q_cols)
            # h_a = h + a Wo^T, statistics of h_a for the post-attention RMSNorm
            gemm(a, L.w_o, M=M, N=H, K=self.q_cols, lda=self.q_cols, ldw=self.q_cols, out=h_a, ldo=H,
                 residual=h, ldr=H, stats_out=S, stats_np=np_h, sk_ws=sk)
            # m = silu(rmsnorm(h_a) Wg'^T) * (rmsnorm(h_a) Wu'^T)
            gemm(h_a, L.w_gu, M=M, N=2 * self.inter, K=H, lda=H, ldw=H, out=act, ldo=self.inter, act=ACT_SWIGLU,
                 ln_stats=S, ln_np=np_h, ln_eps=self.eps, ln_rms=True, sk_ws=sk)
            # h_b = h_a + m Wd^T, statistics of h_b for the next layer's input RMSNorm
            gemm(act, L.w_d, M=M, N=H, K=self.inter, lda=self.inter, ldw=self.inter, out=h_b, ldo=H,
                 residual=h_a, ldr=H, stats_out=S, stats_np=np_h, sk_ws=sk)
            h, stats, np_in = h_b, S, np_h
        return h[:M]

    # ------------------------------------------------------------------------------------------
    def prefill(self, inputs_embeds: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.Tensor] = None, past_key_values=None,
                row_sumsq: Optional[torch.Tensor] = None) -> torch.Tensor:
        """inputs_embeds bf16 [B, L, hidden] + 0/1 attention_mask [B, L] -> last_hidden_state [B, L, hidden] after the
        final norm (zeros at masked positions).  past_key_values (an HF Cache) receives every layer's post-rope keys
        and values in the padded [B, n_kv, L, head_dim] layout HF's decode steps expect."""
        lib = _lib.load()
        st = _lib.stream_ptr()
        B, L, H = inputs_embeds.shape
        dev = inputs_embeds.device
        if attention_mask is None:
            lens = [L] * B
            idx = torch.arange(B * L, dtype=torch.int32, device=dev)
        else:
            keep = attention_mask.reshape(B, L) != 0
            lens = keep.sum(1).tolist()                       # the one host read of the prefill
            idx = torch.nonzero_static(keep.reshape(-1), size=int(sum(lens))).reshape(-1).to(torch.int32)
        M = int(sum(lens))
        if M == 0:
            raise ValueError("prefill: the attention mask keeps no row")
        cu = _lib.h2d(torch.tensor([0] + list(itertools.accumulate(int(n) for n in lens)), dtype=torch.int32), dev)
        if position_ids is None:
            pos_full = torch.arange(L, device=dev, dtype=torch.int32)[None, :].expand(B, L)
        else:
            pos_full = position_ids.to(torch.int32).expand(B, L)
        positions = pos_full.reshape(-1)[idx.long()].contiguous()
        x_in = inputs_embeds.to(torch.bfloat16).contiguous()
        x = torch.empty((M, H), dtype=torch.bfloat16, device=dev)
        _lib.check(lib.vz_rows_move(x_in.data_ptr(), H * 2, x.data_ptr(), H * 2, idx.data_ptr(), M, H * 2, 1, st),
                   "vz_rows_move(gather)")
        rs = None
        if row_sumsq is not None:
            # 8-byte rows: below the 16-byte granule of vz_rows_move
            rs = row_sumsq.reshape(B * L, 2)[idx.long()].float().contiguous()
        kv_sink = None
        if past_key_values is not None:
            kvb = self.kv_cols * 2

            def kv_sink(li, k_rows, v_rows):
                k_pad = torch.zeros((B, L, self.n_kv, self.head_dim), dtype=torch.bfloat16, device=dev)
                v_pad = torch.zeros((B, L, self.n_kv, self.head_dim), dtype=torch.bfloat16, device=dev)
                for src, dst in ((k_rows, k_pad), (v_rows, v_pad)):
                    _lib.check(lib.vz_rows_move(src.data_ptr(), self.qkv_cols * 2, dst.data_ptr(), kvb, idx.data_ptr(),
                                                M, kvb, 0, st), "vz_rows_move(kv)")
                past_key_values.update(k_pad.transpose(1, 2), v_pad.transpose(1, 2), li)

        h = self.forward_packed(x, positions, cu, int(max(lens)), row_sumsq=rs, kv_sink=kv_sink,
                                lens=[int(n) for n in lens if n > 0])
        hn = self.final_norm(h).to(torch.bfloat16).contiguous()
        out = torch.zeros((B, L, H), dtype=torch.bfloat16, device=dev)
        _lib.check(lib.vz_rows_move(hn.data_ptr(), H * 2, out.data_ptr(), H * 2, idx.data_ptr(), M, H * 2, 0, st),
                   "vz_rows_move(scatter)")
        return out.to(inputs_embeds.dtype)
