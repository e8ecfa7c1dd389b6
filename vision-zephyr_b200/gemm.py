"""Python face of the tcgen05 GEMM (vz_gemm_bf16): the building block the training path composes.

out[M,N] = act(A[M,K] . W^T + bias) (+ residual), bf16 operands, fp32 accumulation in TMEM, optional batch.
W is [N,K] (K contiguous, an nn.Linear weight) or, with w_is_kn, [K,N] (N contiguous): the second form is what
makes every backward product a plain call -- dX = dY . W reads the weight as [K,N], dW = dY^T . X reads the
activation as [K,N] -- with no transposed copies of the big operands.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib


def gemm(A: torch.Tensor, W: torch.Tensor, *, M: int, N: int, K: int, lda: int, ldw: int, out: torch.Tensor, ldo: int,
         bias: Optional[torch.Tensor] = None, act: int = 0, w_is_kn: bool = False, batch: int = 1,
         a_bstride: int = 0, w_bstride: int = 0, o_bstride: int = 0, bias_bstride: int = 0,
         out_f32: bool = False, ln_stats: Optional[torch.Tensor] = None, ln_colsum: Optional[torch.Tensor] = None,
         ln_np: int = 0, ln_eps: float = 0.0, ln_rms: bool = False, residual: Optional[torch.Tensor] = None,
         ldr: int = 0, stats_out: Optional[torch.Tensor] = None, stats_np: int = 0,
         sk_ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Raw call: pointers are the tensors' data_ptr() (views welcome), sizes / strides in ELEMENTS."""
    lib = _lib.load()
    g = _lib.GemmArgs()
    g.A, g.W, g.out = A.data_ptr(), W.data_ptr(), out.data_ptr()
    g.bias = bias.data_ptr() if bias is not None else None
    g.residual = residual.data_ptr() if residual is not None else None
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldw, g.ldo, g.ldr = lda, ldw, ldo, ldr
    g.act, g.row_mode, g.rows_per, g.force_simple = act, 0, 0, 0
    g.batch, g.out_f32 = batch, 1 if out_f32 else 0
    g.a_bstride, g.w_bstride, g.o_bstride, g.r_bstride, g.bias_bstride = a_bstride, w_bstride, o_bstride, 0, bias_bstride
    g.ln_stats = ln_stats.data_ptr() if ln_stats is not None else None
    g.ln_colsum = ln_colsum.data_ptr() if ln_colsum is not None else None
    g.stats_out = stats_out.data_ptr() if stats_out is not None else None
    g.ln_np, g.ln_eps, g.stats_np = ln_np, ln_eps, stats_np
    g.sk_ws, g.sk_ws_bytes = (sk_ws.data_ptr(), sk_ws.numel() * sk_ws.element_size()) if sk_ws is not None else (None, 0)
    g.w_is_kn = 1 if w_is_kn else 0
    g.ln_rms = 1 if ln_rms else 0
    _lib.check(lib.vz_gemm_bf16(C.byref(g), _lib.stream_ptr()), f"vz_gemm_bf16 M={M} N={N} K={K} batch={batch} kn={w_is_kn}")
    return out


def _rows8(t: torch.Tensor) -> torch.Tensor:
    """[R, C] -> contiguous copy whose row count is a multiple of 8 (zero rows appended)"""
    R = t.shape[0]
    if R % 8 == 0 and t.is_contiguous():
        return t
    out = torch.zeros(((R + 7) // 8 * 8, t.shape[1]), dtype=t.dtype, device=t.device)
    out[:R] = t
    return out


def linear(x: torch.Tensor, W: torch.Tensor, bias_f32: Optional[torch.Tensor] = None, act: int = 0) -> torch.Tensor:
    """x [M,K] . W[N,K]^T (+ bias) -> [M,N] bf16."""
    x = x.contiguous()
    M, K = x.shape
    N = W.shape[0]
    out = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
    return gemm(x, W, M=M, N=N, K=K, lda=K, ldw=W.stride(0), out=out, ldo=N, bias=bias_f32, act=act)


def matmul_kn(x: torch.Tensor, Wkn: torch.Tensor) -> torch.Tensor:
    """x [M,K] . Wkn[K,N] -> [M,N] bf16 (N % 64 == 0).  The contraction length K = rows of Wkn must be a
    multiple of 8: shorter operands are zero padded."""
    x = x.contiguous()
    M, K = x.shape
    N = Wkn.shape[1]
    if K % 8:
        Kp = (K + 7) // 8 * 8
        xp = torch.zeros((M, Kp), dtype=x.dtype, device=x.device)
        xp[:, :K] = x
        x, Wkn, K = xp, _rows8(Wkn), Kp
    out = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
    return gemm(x, Wkn, M=M, N=N, K=K, lda=K, ldw=Wkn.stride(0), out=out, ldo=N, w_is_kn=True)


class LinearFn(torch.autograd.Function):
    """y = x W^T + b with all three products (y, dX = dY W, dW = dY^T X) on the tcgen05 GEMM."""

    @staticmethod
    def forward(ctx, x, W, b):
        x2 = x.reshape(-1, x.shape[-1]).to(torch.bfloat16).contiguous()
        Wb = W.detach().to(torch.bfloat16)
        if not Wb.is_contiguous():
            Wb = Wb.contiguous()
        y = linear(x2, Wb, b.detach().float().contiguous() if b is not None else None)
        ctx.save_for_backward(x2, Wb)
        ctx.meta = (x.shape, W.dtype, b.dtype if b is not None else None, x.dtype)
        return y.reshape(*x.shape[:-1], W.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, Wb = ctx.saved_tensors
        xshape, wdt, bdt, xdt = ctx.meta
        dy2 = dy.reshape(-1, dy.shape[-1]).to(torch.bfloat16).contiguous()
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = matmul_kn(dy2, Wb).reshape(xshape).to(xdt)          # [M,N] . W[N,K] (the weight read as [K', N'])
        if ctx.needs_input_grad[1]:
            dW = matmul_kn(dy2.t().contiguous(), x2).to(wdt)          # [N,M] . X[M,K]
        if bdt is not None and ctx.needs_input_grad[2]:
            db = dy2.float().sum(0).to(bdt)
        return dx, dW, db
