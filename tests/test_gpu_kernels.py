"""GPU parity of the individual kernels through the C ABI (ctypes), against plain torch fp32 math
on the same bf16 inputs (floating-point kernels) -- run with `-m gpu` on a B200."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    import vision_zephyr_b200  # noqa: F401
    from vision_zephyr_b200 import _lib
    return _lib, _lib.load()


def run_gemm(A, W, bias=None, act=0, residual=None, row_mode=0, rows_per=0, out=None, out_rows=None, simple=0):
    L, lib = _lib()
    M, K = A.shape
    N = W.shape[0]
    if out is None:
        out = torch.zeros((out_rows or M, N), dtype=torch.bfloat16, device=A.device)
    g = L.GemmArgs()
    g.A, g.W, g.out = A.data_ptr(), W.data_ptr(), out.data_ptr()
    g.bias = bias.data_ptr() if bias is not None else None
    g.residual = residual.data_ptr() if residual is not None else None
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldw, g.ldo = A.stride(0), W.stride(0), out.stride(0)
    g.ldr = residual.stride(0) if residual is not None else 0
    g.act, g.row_mode, g.rows_per, g.force_simple = act, row_mode, rows_per, simple
    L.check(lib.vz_gemm_bf16(C.byref(g), L.stream_ptr()), "vz_gemm_bf16")
    torch.cuda.synchronize()
    return out


def ref_gemm(A, W, bias=None, act=0, residual=None):
    y = A.float() @ W.float().t()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = y * torch.sigmoid(1.702 * y)
    elif act == 2:
        y = torch.nn.functional.gelu(y)
    if residual is not None:
        y = y + residual.float()
    return y


def _rand(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


def _check(out, ref, what, tol=2.0 ** -7):
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"{what}: max_abs_err={err:.4g} ref_max={scale:.4g}")
    assert err <= tol * max(scale, 1.0), what


@pytest.mark.parametrize("simple", [0, 1])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 128, 64), (300, 512, 1024), (1154, 3072, 1024),
                                   (32, 12288, 4096), (200, 384, 256), (1000, 4096, 8192), (130, 65536, 5120)])
def test_gemm_plain(M, N, K, simple):
    if simple and M * N * K > 3e10:
        pytest.skip("debug kernel is slow")
    A, W = _rand((M, K), 1.0, 1), _rand((N, K), K ** -0.5, 2)
    out = run_gemm(A, W, simple=simple)
    _check(out, ref_gemm(A, W), f"gemm {M}x{N}x{K} simple={simple}")


@pytest.mark.parametrize("act", [0, 1, 2])
def test_gemm_epilogues(act):
    M, N, K = 1154, 1024, 4096
    A, W = _rand((M, K), 1.0, 3), _rand((N, K), K ** -0.5, 4)
    bias = torch.randn(N, device="cuda") * 0.5
    res = _rand((M, N), 1.0, 5)
    out = run_gemm(A, W, bias=bias, act=act, residual=res)
    _check(out, ref_gemm(A, W, bias, act, res), f"gemm bias+act{act}+residual")
    # in-place residual (out aliases residual), as the Q-Former blocks use it
    x = res.clone()
    run_gemm(A, W, bias=bias, act=act, residual=x, out=x)
    _check(x, ref_gemm(A, W, bias, act, res), f"gemm in-place residual act{act}")


def test_gemm_patch_embed_rows():
    T = 3
    A, W = _rand((T * 576, 592), 1.0, 6), _rand((1024, 592), 0.05, 7)
    A[:, 588:] = 0
    pos = _rand((577, 1024), 1.0, 8)
    out = torch.full((T * 577, 1024), 7.0, dtype=torch.bfloat16, device="cuda")
    run_gemm(A, W, residual=pos, row_mode=1, rows_per=576, out=out)
    ref = (A.float() @ W.float().t()).view(T, 576, 1024) + pos[1:].float()[None]
    _check(out.view(T, 577, 1024)[:, 1:], ref, "patch-embed rows")
    assert (out.view(T, 577, 1024)[:, 0] == 7.0).all(), "CLS rows must be left to the CLS kernel"


def test_gemm_residual_mod_rows():
    A, W = _rand((96, 4096), 1.0, 9), _rand((4096, 4096), 4096 ** -0.5, 10)
    lq = _rand((32, 4096), 1.0, 11)
    out = run_gemm(A, W, residual=lq, row_mode=2, rows_per=32)
    ref = ref_gemm(A, W) + lq.float().repeat(3, 1)
    _check(out, ref, "residual row % 32")


def test_gemm_rejects_bad_arguments():
    L, lib = _lib()
    A, W = _rand((64, 72), 1, 1), _rand((64, 72), 1, 2)
    with pytest.raises(L.VzError):
        run_gemm(A[:, :70], W[:, :70])          # K % 8 != 0
    with pytest.raises(L.VzError):
        run_gemm(_rand((64, 64)), _rand((40, 64)))  # N % 32 != 0


@pytest.mark.parametrize("D,M", [(1024, 1154), (4096, 160), (5120, 576)])
def test_layernorm(D, M):
    L, lib = _lib()
    x = _rand((M, D), 3.0, 12) + 1.5
    g = (1 + 0.1 * torch.randn(D, device="cuda")).float()
    b = (0.1 * torch.randn(D, device="cuda")).float()
    out = torch.empty_like(x)
    L.check(lib.vz_layernorm_bf16(L.ptr(x), D, L.ptr(g), L.ptr(b), L.ptr(out), D, M, D, 1e-5, L.stream_ptr()), "ln")
    ref = torch.nn.functional.layer_norm(x.float(), (D,), g, b, 1e-5)
    _check(out, ref, f"layernorm D={D}", tol=2.0 ** -8)


def test_layernorm_of_zero_row_is_beta():
    L, lib = _lib()
    x = torch.zeros((2, 4096), dtype=torch.bfloat16, device="cuda")
    g = torch.ones(4096, device="cuda")
    b = torch.randn(4096, device="cuda").to(torch.bfloat16).float()
    out = torch.empty_like(x)
    L.check(lib.vz_layernorm_bf16(L.ptr(x), 4096, L.ptr(g), L.ptr(b), L.ptr(out), 4096, 2, 4096, 1e-5, L.stream_ptr()), "ln")
    assert torch.equal(out.float(), b[None].expand(2, -1))


@pytest.mark.parametrize("f32", [True, False])
def test_patchify(f32):
    L, lib = _lib()
    from oracle import pil_ops as P
    T = 2
    px = torch.randn((T, 3, 336, 336), device="cuda")
    src = px if f32 else px.to(torch.bfloat16)
    out = torch.empty((T * 576, 592), dtype=torch.bfloat16, device="cuda")
    L.check(lib.vz_patchify(L.ptr(src), 1 if f32 else 0, T, L.ptr(out), L.stream_ptr()), "patchify")
    ref = torch.from_numpy(P.patchify(src.float().cpu().numpy())).to(torch.bfloat16)
    assert torch.equal(out[:, :588].cpu(), ref)
    assert (out[:, 588:] == 0).all()


def test_gemm_batched_heads_and_f32_output():
    """batched form used by the reassociated cross-attention: strided A/W/out per batch, fp32 scores."""
    L, lib = _lib()
    # (1) batch over heads: A columns h*512.., W columns h*512.., out [(rows), h, N]
    M, H, K, N = 96, 8, 512, 5120
    q = _rand((M, H * K), 1.0, 21)
    wT = _rand((N, H * K), K ** -0.5, 22)
    out = torch.zeros((M, H, N), dtype=torch.bfloat16, device="cuda")
    g = L.GemmArgs()
    g.A, g.W, g.out = q.data_ptr(), wT.data_ptr(), out.data_ptr()
    g.M, g.N, g.K, g.lda, g.ldw, g.ldo = M, N, K, H * K, H * K, H * N
    g.batch, g.a_bstride, g.w_bstride, g.o_bstride = H, K, K, N
    L.check(lib.vz_gemm_bf16(C.byref(g), L.stream_ptr()), "batched heads")
    torch.cuda.synchronize()
    ref = torch.einsum("mhk,nhk->mhn", q.float().view(M, H, K), wT.float().view(N, H, K))
    _check(out, ref, "batched-heads gemm")
    # (2) batch over tiles, N = 576 (partial last n-tile), fp32 output, per-batch bias
    T, Mq, Kf, Np = 3, 256, 5120, 576
    a = _rand((T, Mq, Kf), 1.0, 23)
    f = _rand((T, Np, Kf), Kf ** -0.5, 24)
    bias = torch.randn((T, Np), device="cuda")
    s = torch.full((T, Mq, Np), float("nan"), dtype=torch.float32, device="cuda")
    g = L.GemmArgs()
    g.A, g.W, g.out, g.bias = a.data_ptr(), f.data_ptr(), s.data_ptr(), bias.data_ptr()
    g.M, g.N, g.K, g.lda, g.ldw, g.ldo = Mq, Np, Kf, Kf, Kf, Np
    g.batch, g.out_f32 = T, 1
    g.a_bstride, g.w_bstride, g.o_bstride, g.bias_bstride = Mq * Kf, Np * Kf, Mq * Np, Np
    L.check(lib.vz_gemm_bf16(C.byref(g), L.stream_ptr()), "batched f32")
    torch.cuda.synchronize()
    ref = torch.einsum("tmk,tnk->tmn", a.float(), f.float()) + bias[:, None, :]
    err = (s - ref).abs().max().item()
    print("batched f32 scores max_abs_err", err)
    assert err < 2e-3
    # (3) K = 576 (9 k-blocks), M not a multiple of 128 inside each batch
    P_ = _rand((T, 200, 576), 0.1, 25)
    fT = _rand((T, 640, 576), 1.0, 26)
    o = torch.zeros((T, 200, 640), dtype=torch.bfloat16, device="cuda")
    g = L.GemmArgs()
    g.A, g.W, g.out = P_.data_ptr(), fT.data_ptr(), o.data_ptr()
    g.M, g.N, g.K, g.lda, g.ldw, g.ldo = 200, 640, 576, 576, 576, 640
    g.batch, g.a_bstride, g.w_bstride, g.o_bstride = T, 200 * 576, 640 * 576, 200 * 640
    L.check(lib.vz_gemm_bf16(C.byref(g), L.stream_ptr()), "batched ragged")
    torch.cuda.synchronize()
    _check(o, torch.einsum("tmk,tnk->tmn", P_.float(), fT.float()), "batched ragged-M gemm")


def test_gemm_fused_layernorm_producer_and_consumer():
    """LayerNorm folded around the GEMM: the producer epilogue emits partial row statistics, the consumer
    (gamma/beta folded into its weights) finishes the normalisation in its epilogue."""
    L, lib = _lib()
    M, D, N = 1154, 1024, 3072
    A0, W0 = _rand((M, 512), 1.0, 41), _rand((D, 512), 512 ** -0.5, 42)
    res = _rand((M, D), 1.0, 43) + 0.7                       # non-zero row means
    np_ = lib.vz_gemm_stats_partials(M, D)
    stats = torch.full((M, np_, 2), float("nan"), dtype=torch.float32, device="cuda")
    x = torch.zeros((M, D), dtype=torch.bfloat16, device="cuda")
    g = L.GemmArgs()
    g.A, g.W, g.out, g.residual = A0.data_ptr(), W0.data_ptr(), x.data_ptr(), res.data_ptr()
    g.M, g.N, g.K, g.lda, g.ldw, g.ldo, g.ldr = M, D, 512, 512, 512, D, D
    g.stats_out, g.stats_np = stats.data_ptr(), np_
    L.check(lib.vz_gemm_bf16(C.byref(g), L.stream_ptr()), "producer")
    torch.cuda.synchronize()
    xf = ref_gemm(A0, W0, residual=res)                      # fp32 values the epilogue summed
    s = stats.sum(1)
    assert torch.allclose(s[:, 0], xf.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[:, 1], (xf * xf).sum(1), rtol=1e-4, atol=1e-2)
    # consumer
    gamma = (1 + 0.1 * torch.randn(D, device="cuda")).float()
    beta = (0.1 * torch.randn(D, device="cuda")).float()
    W, b = _rand((N, D), D ** -0.5, 44), torch.randn(N, device="cuda") * 0.1
    Wf = (W.float() * gamma[None]).to(torch.bfloat16).contiguous()
    bf_ = (b + W.float() @ beta).contiguous()
    cs = Wf.float().sum(1).contiguous()
    out = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
    g = L.GemmArgs()
    g.A, g.W, g.out, g.bias = x.data_ptr(), Wf.data_ptr(), out.data_ptr(), bf_.data_ptr()
    g.M, g.N, g.K, g.lda, g.ldw, g.ldo = M, N, D, D, D, N
    g.ln_stats, g.ln_colsum, g.ln_np, g.ln_eps = stats.data_ptr(), cs.data_ptr(), np_, 1e-5
    L.check(lib.vz_gemm_bf16(C.byref(g), L.stream_ptr()), "consumer")
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.float(), (D,), gamma, beta, 1e-5) @ W.float().t() + b
    _check(out, ref, "fused LayerNorm + GEMM", tol=2.0 ** -6)


@pytest.mark.parametrize("M,N,K,act,res", [(9000, 3072, 1024, 0, False), (10050, 1024, 4096, 0, True),
                                           (9000, 4096, 1024, 1, False), (9000, 1024, 1024, 2, True)])
def test_gemm_two_cta_form(M, N, K, act, res):
    """large problems take the cta_group::2 form (256-row pair tiles); ragged M tail included"""
    A, W = _rand((M, K), 1.0, 51), _rand((N, K), K ** -0.5, 52)
    bias = torch.randn(N, device="cuda") * 0.3
    R = _rand((M, N), 1.0, 53) if res else None
    out = run_gemm(A, W, bias=bias, act=act, residual=R)
    _check(out, ref_gemm(A, W, bias, act, R), f"2-CTA gemm {M}x{N}x{K} act={act} res={res}")


_CANARY = 1 << 20


def _sk_ws():
    """scratch of exactly the advertised size, followed by a canary that must never be written"""
    L, lib = _lib()
    n = lib.vz_gemm_sk_workspace_bytes()
    buf = torch.zeros(n + _CANARY, dtype=torch.uint8, device="cuda")
    buf[n:] = 0xA5
    return buf[:n], buf


def _canary_intact(buf):
    return bool((buf[-_CANARY:] == 0xA5).all().item())


@pytest.mark.parametrize("M,N,K,act,res", [
    (1280, 4096, 4096, 0, True),     # Q-Former out-proj: 80 pair tiles on 74 CTA pairs (2-CTA form)
    (1280, 8192, 4096, 2, False),    # FFN1: 160 pair tiles
    (1280, 4096, 8192, 0, True),     # FFN2: 128 k-blocks
    (1280, 512, 5120, 0, False),     # fewer tiles than SMs: every tile is shared by several CTAs (1-CTA form)
    (1154, 3072, 1024, 1, False),    # ragged M, quick-GELU
    (23080, 3072, 1024, 0, False),   # ViT qkv at the bench size: 1092 pair tiles, 14.76 rounds
    (577, 1024, 4096, 0, True),      # one tile: 40 tiles of 128x128 shared by all 148 SMs
    (24, 8192, 4096, 0, False),      # text K/V rows of a short prompt
])
def test_gemm_stream_k_tail(M, N, K, act, res):
    """With a stream-K scratch buffer the last, partially filled round of tiles is split along K over all
    SMs; partial sums are added in a fixed order, so two runs agree bit for bit and the result matches
    the whole-tile schedule within bf16 rounding."""
    L, lib = _lib()
    A, W = _rand((M, K), 1.0, 61), _rand((N, K), K ** -0.5, 62)
    bias = torch.randn(N, device="cuda") * 0.3
    R = _rand((M, N), 1.0, 63) if res else None
    ws, ws_all = _sk_ws()

    def run(use_sk):
        out = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
        g = L.GemmArgs()
        g.A, g.W, g.out, g.bias = A.data_ptr(), W.data_ptr(), out.data_ptr(), bias.data_ptr()
        g.residual = R.data_ptr() if R is not None else None
        g.M, g.N, g.K, g.lda, g.ldw, g.ldo, g.ldr = M, N, K, K, K, N, (N if R is not None else 0)
        g.act = act
        if use_sk:
            g.sk_ws, g.sk_ws_bytes = ws.data_ptr(), ws.numel()
        L.check(lib.vz_gemm_bf16(C.byref(g), L.stream_ptr()), "vz_gemm_bf16")
        torch.cuda.synchronize()
        return out

    ref = ref_gemm(A, W, bias, act, R)
    o_sk, o_sk2, o_dp = run(True), run(True), run(False)
    _check(o_sk, ref, f"stream-K gemm {M}x{N}x{K}")
    assert torch.equal(o_sk, o_sk2), "stream-K result must be deterministic"
    assert _canary_intact(ws_all), "stream-K scratch overrun"
    d = (o_sk.float() - o_dp.float()).abs().max().item()
    print("stream-K vs whole-tile max diff", d)
    assert d <= 2.0 ** -6 * max(ref.abs().max().item(), 1.0)


def test_gemm_stream_k_batched_f32():
    """batched fp32-output form (cross-attention scores): 120 tiles of 128x128... split along K = 5120"""
    L, lib = _lib()
    T, Mq, Kf, Np = 40, 256, 5120, 576
    a = _rand((T, Mq, Kf), 1.0, 71)
    f = _rand((T, Np, Kf), Kf ** -0.5, 72)
    ws, ws_all = _sk_ws()
    outs = []
    for use_sk in (True, False):
        s = torch.full((T, Mq, Np), float("nan"), dtype=torch.float32, device="cuda")
        g = L.GemmArgs()
        g.A, g.W, g.out = a.data_ptr(), f.data_ptr(), s.data_ptr()
        g.M, g.N, g.K, g.lda, g.ldw, g.ldo = Mq, Np, Kf, Kf, Kf, Np
        g.batch, g.out_f32 = T, 1
        g.a_bstride, g.w_bstride, g.o_bstride = Mq * Kf, Np * Kf, Mq * Np
        if use_sk:
            g.sk_ws, g.sk_ws_bytes = ws.data_ptr(), ws.numel()
        L.check(lib.vz_gemm_bf16(C.byref(g), L.stream_ptr()), "batched f32 stream-K")
        torch.cuda.synchronize()
        outs.append(s)
    ref = torch.einsum("tmk,tnk->tmn", a.float(), f.float())
    for s in outs:
        assert (s - ref).abs().max().item() < 2e-3
    assert _canary_intact(ws_all), "stream-K scratch overrun"


@pytest.mark.parametrize("T,M,N,K", [(40, 256, 5120, 576), (1, 256, 5120, 576), (3, 200, 640, 192), (1, 1280, 512, 4096)])
def test_gemm_weights_as_k_by_n(T, M, N, K):
    """w_is_kn: the second operand is [K, N] row-major (out = A W), consumed as an MN-major tcgen05 operand --
    the cross-attention's P f product reads the features in place instead of a transposed copy"""
    L, lib = _lib()
    A = _rand((T, M, K), 0.1, 81)
    Wkn = _rand((T, K, N), 1.0, 82)
    bias = torch.randn(N, device="cuda") * 0.2
    out = torch.full((T, M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    g = L.GemmArgs()
    g.A, g.W, g.out, g.bias = A.data_ptr(), Wkn.data_ptr(), out.data_ptr(), bias.data_ptr()
    g.M, g.N, g.K, g.lda, g.ldw, g.ldo = M, N, K, K, N, N
    g.batch, g.a_bstride, g.w_bstride, g.o_bstride = T, M * K, K * N, M * N
    g.w_is_kn = 1
    L.check(lib.vz_gemm_bf16(C.byref(g), L.stream_ptr()), "w_is_kn gemm")
    torch.cuda.synchronize()
    ref = torch.einsum("tmk,tkn->tmn", A.float(), Wkn.float()) + bias
    _check(out, ref, f"[K,N]-operand gemm T={T} {M}x{N}x{K}")


@pytest.mark.parametrize("M,N,K", [(32, 4096, 4096), (32, 12288, 4096), (64, 8192, 4096), (128, 4096, 8192), (37, 4096, 4096)])
@pytest.mark.parametrize("variant", ["bias", "ln_gelu", "res_stats", "res_mod"])
def test_gemm_single_m_tile_epilogues(M, N, K, variant):
    """single-m-tile, weight-streaming problems (the Q-Former at 1-4 tiles: a short A box, three quarters of the
    tile's lanes empty, every tile cut along K over all SMs by the stream-K schedule).  Every epilogue the
    orchestration uses there, against fp32 math, run to run, and against the whole-tile schedule."""
    L, lib = _lib()
    A, W = _rand((M, K), 1.0, 71), _rand((N, K), K ** -0.5, 72)
    bias = torch.randn(N, device="cuda") * 0.3
    ws, ws_all = _sk_ws()
    R = stats_in = cs = None
    act, row_mode, rows_per, np_in = 0, 0, 0, 0
    if variant == "ln_gelu":
        act, np_in = 2, 3
        xf = A.float()
        stats_in = torch.zeros((M, np_in, 2), dtype=torch.float32, device="cuda")
        stats_in[:, 0, 0], stats_in[:, 0, 1] = xf.sum(1) * 0.25, (xf * xf).sum(1) * 0.5      # partials that add up
        stats_in[:, 1, 0], stats_in[:, 1, 1] = xf.sum(1) * 0.75, (xf * xf).sum(1) * 0.25
        stats_in[:, 2, 1] = (xf * xf).sum(1) * 0.25
        cs = W.float().sum(1).contiguous()
    elif variant == "res_stats":
        R = _rand((M, N), 1.0, 73)
    elif variant == "res_mod":
        R, row_mode, rows_per = _rand((32, N), 1.0, 74), 2, 32
    np_out = lib.vz_gemm_stats_partials(M, N) if variant == "res_stats" else 0
    stats_out = torch.full((M, max(np_out, 1), 2), float("nan"), dtype=torch.float32, device="cuda")

    def run(use_ws):
        out = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
        g = L.GemmArgs()
        g.A, g.W, g.out, g.bias = A.data_ptr(), W.data_ptr(), out.data_ptr(), bias.data_ptr()
        g.residual = R.data_ptr() if R is not None else None
        g.M, g.N, g.K, g.lda, g.ldw, g.ldo, g.ldr = M, N, K, K, K, N, (N if R is not None else 0)
        g.act, g.row_mode, g.rows_per = act, row_mode, rows_per
        if stats_in is not None:
            g.ln_stats, g.ln_colsum, g.ln_np, g.ln_eps = stats_in.data_ptr(), cs.data_ptr(), np_in, 1e-5
        if np_out:
            g.stats_out, g.stats_np = stats_out.data_ptr(), np_out
        if use_ws:
            g.sk_ws, g.sk_ws_bytes = ws.data_ptr(), ws.numel()
        L.check(lib.vz_gemm_bf16(C.byref(g), L.stream_ptr()), "vz_gemm_bf16")
        torch.cuda.synchronize()
        return out

    y = A.float() @ W.float().t()
    if variant == "ln_gelu":
        xf = A.float()
        mu, var = xf.mean(1, keepdim=True), xf.var(1, unbiased=False, keepdim=True)
        y = torch.nn.functional.gelu(torch.rsqrt(var + 1e-5) * (y - mu * cs[None]) + bias)
    else:
        y = y + bias
        if R is not None:
            y = y + (R.float()[torch.arange(M, device="cuda") % 32] if variant == "res_mod" else R.float())
    o1, o2 = run(True), run(True)
    _check(o1, y, f"single-m-tile gemm {M}x{N}x{K} {variant}")
    assert torch.equal(o1, o2) and _canary_intact(ws_all)
    if np_out:
        s = stats_out.sum(1)
        assert torch.allclose(s[:, 0], y.sum(1), rtol=1e-4, atol=2e-2) and torch.allclose(s[:, 1], (y * y).sum(1), rtol=1e-4, atol=2e-2)
    o0 = run(False)                                           # no scratch: whole tiles only
    d = (o1.float() - o0.float()).abs().max().item()
    assert d <= 2.0 ** -6 * max(y.abs().max().item(), 1.0), d
