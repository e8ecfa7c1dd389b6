"""Scope row f-3: the Mistral prefill behind the splice, native on packed rows (mistral_prefill.py).

Checks, against PyTorch / HF Mistral on the same bf16 weights:
  * the two GEMM epilogue forms the decoder layer adds (RMSNorm consumer, SwiGLU over interleaved gate / up rows),
  * vz_rope_apply against HF's apply_rotary_pos_emb,
  * vz_rows_move (padded <-> packed),
  * the whole stack (2 layers, Zephyr geometry) against HF MistralModel for ragged right- and left-padded batches,
    including the KV cache it leaves for HF's decode steps.
Tolerance: bf16 path, cosine >= 0.999 per row and max-abs within 3 % of the tensor's range (same bar as the tokens).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import cos_rows

pytestmark = pytest.mark.gpu


def _rms(x, eps):
    x = x.float()
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)


@pytest.mark.parametrize("M,K,I", [(100, 512, 256), (700, 1024, 1024), (3000, 512, 4096), (2500, 4096, 2048),
                                   (8970, 4096, 14336)])        # the last one: config 5's gate | up GEMM at full size
def test_gemm_rmsnorm_swiglu_matches_torch(M, K, I):
    from vision_zephyr_b200 import _lib
    from vision_zephyr_b200.gemm import gemm
    from vision_zephyr_b200.mistral_prefill import ACT_SWIGLU, interleave_gate_up
    lib = _lib.load()
    gen = torch.Generator(device="cuda").manual_seed(M + K)
    x = (torch.randn((M, K), device="cuda", generator=gen) * 1.7).to(torch.bfloat16)
    gamma = 1.0 + 0.2 * torch.randn(K, device="cuda", generator=gen)
    wg = (torch.randn((I, K), device="cuda", generator=gen) / K ** 0.5 * gamma).to(torch.bfloat16)
    wu = (torch.randn((I, K), device="cuda", generator=gen) / K ** 0.5 * gamma).to(torch.bfloat16)
    eps = 1e-5
    stats = torch.empty((M, 2), dtype=torch.float32, device="cuda")
    _lib.check(lib.vz_row_stats(x.data_ptr(), K, M, K, stats.data_ptr(), _lib.stream_ptr()), "row_stats")
    assert torch.allclose(stats[:, 1], x.float().pow(2).sum(1), rtol=1e-5)
    assert torch.allclose(stats[:, 0], x.float().sum(1), rtol=1e-4, atol=1e-2)
    stats[:, 0] = 1e9            # RMS mode must not read the sum
    sk = torch.empty(lib.vz_gemm_sk_workspace_bytes(), dtype=torch.uint8, device="cuda")
    out = torch.full((M, I), float("nan"), dtype=torch.bfloat16, device="cuda")
    gemm(x, interleave_gate_up(wg, wu), M=M, N=2 * I, K=K, lda=K, ldw=K, out=out, ldo=I, act=ACT_SWIGLU,
         ln_stats=stats, ln_np=1, ln_eps=eps, ln_rms=True, sk_ws=sk)
    rstd = torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + eps)
    g = (x.float() @ wg.float().t()) * rstd
    u = (x.float() @ wu.float().t()) * rstd
    ref = torch.nn.functional.silu(g) * u
    got = out.float()
    assert torch.isfinite(got).all()
    assert cos_rows(got.cpu().numpy(), ref.cpu().numpy()).min() >= 0.9999
    assert (got - ref).abs().max().item() <= 0.02 * ref.abs().max().item() + 1e-3
    # plain RMS-fused linear (the q|k|v form), bias-free
    N = 2 * I
    w = interleave_gate_up(wg, wu)
    out2 = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    gemm(x, w, M=M, N=N, K=K, lda=K, ldw=K, out=out2, ldo=N, ln_stats=stats, ln_np=1, ln_eps=eps, ln_rms=True, sk_ws=sk)
    ref2 = (x.float() @ w.float().t()) * rstd
    assert cos_rows(out2.float().cpu().numpy(), ref2.cpu().numpy()).min() >= 0.9999
    assert (out2.float() - ref2).abs().max().item() <= 0.01 * ref2.abs().max().item() + 1e-3


def test_gemm_swiglu_without_norm_and_residual_stats_chain():
    """o_proj / down_proj form: residual + row statistics out, consumed by an RMS-fused GEMM (partials summed)."""
    from vision_zephyr_b200 import _lib
    from vision_zephyr_b200.gemm import gemm
    lib = _lib.load()
    M, K, H = 1500, 1024, 4096
    gen = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn((M, K), device="cuda", generator=gen).to(torch.bfloat16)
    h = torch.randn((M, H), device="cuda", generator=gen).to(torch.bfloat16)
    wo = (torch.randn((H, K), device="cuda", generator=gen) / K ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn((512, H), device="cuda", generator=gen) / H ** 0.5).to(torch.bfloat16)
    np_h = lib.vz_gemm_stats_partials(M, H)
    S = torch.zeros((M, np_h, 2), dtype=torch.float32, device="cuda")
    h1 = torch.empty_like(h)
    gemm(a, wo, M=M, N=H, K=K, lda=K, ldw=K, out=h1, ldo=H, residual=h, ldr=H, stats_out=S, stats_np=np_h)
    ref_h1 = h.float() + a.float() @ wo.float().t()
    assert cos_rows(h1.float().cpu().numpy(), ref_h1.cpu().numpy()).min() >= 0.9999
    assert torch.allclose(S[:, :, 1].sum(1), ref_h1.pow(2).sum(1), rtol=2e-3)
    out = torch.empty((M, 512), dtype=torch.bfloat16, device="cuda")
    gemm(h1, w2, M=M, N=512, K=H, lda=H, ldw=H, out=out, ldo=512, ln_stats=S, ln_np=np_h, ln_eps=1e-5, ln_rms=True)
    ref = _rms(h1, 1e-5) @ w2.float().t()
    assert cos_rows(out.float().cpu().numpy(), ref.cpu().numpy()).min() >= 0.9999


def test_rope_matches_hf_apply_rotary_pos_emb():
    from transformers import MistralConfig
    from transformers.models.mistral.modeling_mistral import MistralRotaryEmbedding, apply_rotary_pos_emb
    from vision_zephyr_b200 import _lib
    lib = _lib.load()
    cfg = MistralConfig(hidden_size=4096, num_attention_heads=32, num_key_value_heads=8, rope_theta=10000.0)
    rot = MistralRotaryEmbedding(cfg).cuda()
    M, nh, nkv, hd = 777, 32, 8, 128
    gen = torch.Generator(device="cuda").manual_seed(5)
    qkv = torch.randn((M, (nh + 2 * nkv) * hd), device="cuda", generator=gen).to(torch.bfloat16)
    pos = torch.randint(0, 2300, (M,), device="cuda", generator=gen).to(torch.int32)
    pos[:3] = torch.tensor([0, 1, 2047], dtype=torch.int32)
    q = qkv[:, : nh * hd].reshape(1, M, nh, hd).transpose(1, 2)
    k = qkv[:, nh * hd: (nh + nkv) * hd].reshape(1, M, nkv, hd).transpose(1, 2)
    cos, sin = rot(qkv, pos[None, :].long())
    q_ref, k_ref = apply_rotary_pos_emb(q, k, cos, sin)
    v_before = qkv[:, (nh + nkv) * hd:].clone()
    cs = torch.empty((M, hd // 2, 2), dtype=torch.float32, device="cuda")
    st = _lib.stream_ptr()
    _lib.check(lib.vz_rope_table(pos.data_ptr(), M, rot.inv_freq.float().contiguous().data_ptr(), hd // 2,
                                 cs.data_ptr(), st), "rope_table")
    _lib.check(lib.vz_rope_apply(qkv.data_ptr(), qkv.shape[1], M, nh + nkv, hd, cs.data_ptr(), st), "rope_apply")
    got_q = qkv[:, : nh * hd].reshape(M, nh, hd).float()
    got_k = qkv[:, nh * hd: (nh + nkv) * hd].reshape(M, nkv, hd).float()
    ref_q = q_ref[0].transpose(0, 1).float()
    ref_k = k_ref[0].transpose(0, 1).float()
    for got, ref in ((got_q, ref_q), (got_k, ref_k)):
        diff = (got - ref).abs()
        # same rounding points as the bf16 tensor expression; cos / sin may differ in the last fp32 bit before
        # their own bf16 rounding, which moves a handful of results by one bf16 step
        assert (diff > 0).float().mean().item() < 2e-3
        assert diff.max().item() <= 0.04
    assert torch.equal(qkv[:, (nh + nkv) * hd:], v_before)


def test_rows_move_gather_scatter_round_trip():
    from vision_zephyr_b200 import _lib
    lib = _lib.load()
    B, L, D = 3, 37, 256
    src = torch.randn((B * L, D), device="cuda").to(torch.bfloat16)
    keep = torch.rand(B * L, device="cuda") < 0.6
    idx = torch.nonzero(keep).reshape(-1).to(torch.int32)
    M = idx.numel()
    packed = torch.empty((M, D), dtype=torch.bfloat16, device="cuda")
    st = _lib.stream_ptr()
    _lib.check(lib.vz_rows_move(src.data_ptr(), D * 2, packed.data_ptr(), D * 2, idx.data_ptr(), M, D * 2, 1, st), "g")
    assert torch.equal(packed, src[idx.long()])
    back = torch.zeros_like(src)
    _lib.check(lib.vz_rows_move(packed.data_ptr(), D * 2, back.data_ptr(), D * 2, idx.data_ptr(), M, D * 2, 0, st), "s")
    assert torch.equal(back, src * keep[:, None])
    # strided source (the k columns of a packed q|k|v row)
    wide = torch.randn((M, 3 * D), device="cuda").to(torch.bfloat16)
    dst = torch.zeros((B * L, D), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.vz_rows_move(wide[:, D:].data_ptr(), 3 * D * 2, dst.data_ptr(), D * 2, idx.data_ptr(), M, D * 2, 0, st), "k")
    assert torch.equal(dst[idx.long()], wide[:, D: 2 * D])


@pytest.fixture(scope="module")
def mistral2():
    """2 decoder layers with Zephyr-7B geometry (hidden 4096, 32 / 8 heads of 128, intermediate 14336), random init"""
    from transformers import MistralConfig, MistralModel
    cfg = MistralConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=2, num_attention_heads=32,
                        num_key_value_heads=8, vocab_size=1000, max_position_embeddings=32768, rms_norm_eps=1e-5,
                        rope_theta=10000.0, sliding_window=None, attn_implementation="sdpa")
    torch.manual_seed(11)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device("cuda"):
            m = MistralModel(cfg)
    finally:
        torch.set_default_dtype(old)
    m.eval().requires_grad_(False)
    with torch.no_grad():       # non-trivial norm gains, so that the folding is exercised
        for lyr in m.layers:
            lyr.input_layernorm.weight.mul_(1.0 + 0.3 * torch.randn_like(lyr.input_layernorm.weight))
            lyr.post_attention_layernorm.weight.mul_(1.0 + 0.3 * torch.randn_like(lyr.post_attention_layernorm.weight))
        m.norm.weight.mul_(1.0 + 0.3 * torch.randn_like(m.norm.weight))
    return m


@pytest.mark.parametrize("side", ["right", "left"])
def test_native_prefill_matches_hf_mistral(mistral2, side):
    from transformers import DynamicCache
    from vision_zephyr_b200.mistral_prefill import MistralPrefillB200
    eng = MistralPrefillB200(mistral2)
    B, L, H = 3, 300, 4096
    lens = [300, 37, 181]
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((B, L, H), device="cuda", generator=gen).to(torch.bfloat16)
    mask = torch.zeros((B, L), dtype=torch.long, device="cuda")
    pos = torch.zeros((B, L), dtype=torch.long, device="cuda")
    for b, n in enumerate(lens):
        sl = slice(0, n) if side == "right" else slice(L - n, L)
        mask[b, sl] = 1
        pos[b, sl] = torch.arange(n, device="cuda")
    with torch.no_grad():
        ref_cache = DynamicCache(config=mistral2.config)
        ref = mistral2(inputs_embeds=x, attention_mask=mask, position_ids=pos, past_key_values=ref_cache, use_cache=True)
        cache = DynamicCache(config=mistral2.config)
        got = eng.prefill(x, mask, pos, cache)
    keep = mask.bool()
    r = ref.last_hidden_state[keep].float().cpu().numpy()
    g = got[keep].float().cpu().numpy()
    assert np.isfinite(g).all()
    assert cos_rows(g, r).min() >= 0.999
    assert np.abs(g - r).max() <= 0.03 * np.abs(r).max()
    assert got[~keep].abs().max().item() == 0.0
    for li in range(2):
        rk, rv = ref_cache.layers[li].keys, ref_cache.layers[li].values          # [B, 8, L, 128]
        gk, gv = cache.layers[li].keys, cache.layers[li].values
        assert gk.shape == rk.shape and gv.shape == rv.shape
        for a, b_ in ((gk, rk), (gv, rv)):
            a2 = a.transpose(1, 2)[keep].reshape(-1, 128).float().cpu().numpy()
            b2 = b_.transpose(1, 2)[keep].reshape(-1, 128).float().cpu().numpy()
            assert cos_rows(a2, b2).min() >= 0.998
            assert np.abs(a2 - b2).max() <= 0.03 * np.abs(b2).max()


def test_native_prefill_row_statistics_from_the_caller(mistral2):
    """row_sumsq handed in (the scatter's by-product) == computed inside"""
    from vision_zephyr_b200.mistral_prefill import MistralPrefillB200
    eng = MistralPrefillB200(mistral2)
    B, L, H = 2, 130, 4096
    x = torch.randn((B, L, H), device="cuda").to(torch.bfloat16)
    mask = torch.ones((B, L), dtype=torch.long, device="cuda")
    mask[1, 100:] = 0
    stats = torch.stack([torch.zeros((B, L), device="cuda"), x.float().pow(2).sum(-1)], -1)
    with torch.no_grad():
        a = eng.prefill(x, mask, None, None).clone()
        b = eng.prefill(x, mask, None, None, row_sumsq=stats)
    assert cos_rows(a[mask.bool()].float().cpu().numpy(), b[mask.bool()].float().cpu().numpy()).min() >= 0.9999


def _attn_reference(qkv, lens, nh, nkv, hd):
    """fp32 causal GQA attention per sample (q head h reads kv head h // (nh / nkv), like HF repeat_kv)"""
    out = torch.empty((qkv.shape[0], nh * hd), dtype=torch.float32, device=qkv.device)
    row = 0
    for n in lens:
        blk = qkv[row: row + n].float()
        q = blk[:, : nh * hd].reshape(n, nh, hd).transpose(0, 1)
        k = blk[:, nh * hd: (nh + nkv) * hd].reshape(n, nkv, hd).transpose(0, 1).repeat_interleave(nh // nkv, 0)
        v = blk[:, (nh + nkv) * hd:].reshape(n, nkv, hd).transpose(0, 1).repeat_interleave(nh // nkv, 0)
        o = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True)
        out[row: row + n] = o.transpose(0, 1).reshape(n, nh * hd)
        row += n
    return out


@pytest.mark.parametrize("lens", [[1], [64], [65, 128, 129], [300, 37, 181, 1, 127, 256], [2140, 514, 1645],
                                  [5, 40, 17] * 24, [4100]])
def test_causal_gqa_attention_matches_sdpa(lens):
    """vz_attn_causal on packed rows: tile / block boundary lengths, single-row samples, the longest config-5 sample"""
    from vision_zephyr_b200 import _lib
    lib = _lib.load()
    nh, nkv, hd = 32, 8, 128
    M = sum(lens)
    gen = torch.Generator(device="cuda").manual_seed(M)
    qkv = torch.randn((M, (nh + 2 * nkv) * hd), device="cuda", generator=gen).to(torch.bfloat16)
    qkv[:, : nh * hd] *= 2.0           # sharper softmax: the running-maximum rescale is exercised
    lens_h = np.asarray(lens, dtype=np.int32)
    n = lib.vz_attn_causal_items(lens_h.ctypes.data, len(lens), nh, None, 0, None)
    assert n == sum((s + 127) // 128 for s in lens) * nh
    items_h = np.empty((n, 4), dtype=np.int32)
    flops = C.c_double(0)
    assert lib.vz_attn_causal_items(lens_h.ctypes.data, len(lens), nh, items_h.ctypes.data, n, C.byref(flops)) == n
    assert flops.value == 4.0 * hd * sum(s * (s + 1) / 2 for s in lens) * nh
    nkb = (items_h[:, 2] + 63) // 64
    assert (np.diff(nkb) <= 0).all()                    # longest first
    items = torch.from_numpy(items_h).cuda()
    out = torch.full((M + 3, nh * hd), 7.0, dtype=torch.bfloat16, device="cuda")     # 3 guard rows behind the last sample
    _lib.check(lib.vz_attn_causal(qkv.data_ptr(), qkv.shape[1], M, out.data_ptr(), nh * hd, items.data_ptr(), n, nh, nkv,
                                  hd, hd ** -0.5, flops.value, _lib.stream_ptr()), "vz_attn_causal")
    torch.cuda.synchronize()
    ref = _attn_reference(qkv, lens, nh, nkv, hd)
    got = out[:M].float()
    assert torch.isfinite(got).all()
    assert (out[M:] == 7.0).all()                       # nothing written past the last valid row
    g2, r2 = got.reshape(M * nh, hd).cpu().numpy(), ref.reshape(M * nh, hd).cpu().numpy()
    assert cos_rows(g2, r2).min() >= 0.999
    assert np.abs(g2 - r2).max() <= 0.03 * np.abs(r2).max()


def test_native_attention_agrees_with_flash_attn_in_the_stack(mistral2):
    from vision_zephyr_b200.mistral_prefill import MistralPrefillB200
    eng = MistralPrefillB200(mistral2)
    assert eng.attn_impl == "native"
    B, L, H = 2, 400, 4096
    x = torch.randn((B, L, H), device="cuda").to(torch.bfloat16)
    mask = torch.ones((B, L), dtype=torch.long, device="cuda")
    mask[0, 333:] = 0
    with torch.no_grad():
        a = eng.prefill(x, mask, None, None).clone()
        eng.attn_impl = "fa2"
        b = eng.prefill(x, mask, None, None)
    keep = mask.bool()
    assert cos_rows(a[keep].float().cpu().numpy(), b[keep].float().cpu().numpy()).min() >= 0.9995


def test_sliding_window_shorter_than_the_prompt_takes_the_windowed_core():
    """Mistral's sliding window (config.sliding_window; Zephyr ships 4096, longer than any 2048-token prompt): when a
    sample is longer than the window the engine hands the attention to flash-attn's windowed varlen call, and the
    result follows HF's sliding-window mask."""
    from transformers import MistralConfig, MistralModel
    from vision_zephyr_b200.mistral_prefill import MistralPrefillB200
    cfg = MistralConfig(hidden_size=4096, intermediate_size=1024, num_hidden_layers=1, num_attention_heads=32,
                        num_key_value_heads=8, vocab_size=100, rms_norm_eps=1e-5, rope_theta=10000.0, sliding_window=48,
                        attn_implementation="sdpa")
    torch.manual_seed(5)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device("cuda"):
            m = MistralModel(cfg)
    finally:
        torch.set_default_dtype(old)
    m.eval().requires_grad_(False)
    B, L = 2, 160
    x = torch.randn((B, L, 4096), device="cuda").to(torch.bfloat16)
    mask = torch.ones((B, L), dtype=torch.long, device="cuda")
    mask[1, 100:] = 0
    with torch.no_grad():
        ref = m(inputs_embeds=x, attention_mask=mask, use_cache=False).last_hidden_state
        got = MistralPrefillB200(m).prefill(x, mask, None, None)
    keep = mask.bool()
    assert cos_rows(got[keep].float().cpu().numpy(), ref[keep].float().cpu().numpy()).min() >= 0.999


def test_engine_refolds_when_the_weights_change():
    """VisZephyrB200Model.native_prefill() keys its folded weights on (pointer, version) of every decoder parameter"""
    from vision_zephyr_b200.language_model import VisZephyrB200Model, random_mistral_config
    cfg = random_mistral_config(num_hidden_layers=1, intermediate_size=512, vocab_size=64)
    # the LLM part alone: no vision modules are built without mm_vision_tower
    cfg.mm_vision_tower = None
    torch.manual_seed(2)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device("cuda"):
            from transformers import MistralModel
            m = MistralModel(cfg)
    finally:
        torch.set_default_dtype(old)
    m.eval().requires_grad_(False)
    m.__class__ = type("Patched", (MistralModel,), {"native_prefill": VisZephyrB200Model.native_prefill})
    m._vz_prefill, m._vz_prefill_key = None, None
    e1 = m.native_prefill()
    assert m.native_prefill() is e1                      # unchanged weights: same engine
    x = torch.randn((1, 40, 4096), device="cuda").to(torch.bfloat16)
    with torch.no_grad():
        a = e1.prefill(x, None, None, None).clone()
        m.layers[0].mlp.down_proj.weight.mul_(0.5)       # in-place update bumps the version
        e2 = m.native_prefill()
        assert e2 is not e1
        b = e2.prefill(x, None, None, None)
        ref = m(inputs_embeds=x, use_cache=False).last_hidden_state
    assert not torch.equal(a, b)
    assert cos_rows(b[0].float().cpu().numpy(), ref[0].float().cpu().numpy()).min() >= 0.999
