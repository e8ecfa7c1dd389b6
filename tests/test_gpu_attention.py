"""GPU parity of the attention kernels (via the model-level C ABI) against torch fp32 attention."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


def test_vit_layer_stack_matches_oracle_layerwise(vision_path, seeded_weights):
    """Runs the whole tower and compares EVERY hidden state with the fp32 oracle, so a defect in the
    embedding GEMM, LayerNorm, attention or MLP shows up at the first layer it touches."""
    from oracle import model as M
    tower = vision_path.model.vision_tower
    px = torch.randn((1, 3, 336, 336), generator=torch.Generator().manual_seed(5)).to(torch.bfloat16).float()
    with torch.no_grad():
        ref = M.clip_hidden_states(seeded_weights["clip"], px)
    patches = tower._patches_of(px.cuda())
    fused, hidden = tower.encode_patches(patches, return_hidden=True)
    torch.cuda.synchronize()
    worst = 0.0
    for l in range(25):
        got = hidden[l].float().cpu()
        err = (got - ref[l]).abs().max().item()
        cos = torch.nn.functional.cosine_similarity(got.flatten(), ref[l].flatten(), dim=0).item()
        print(f"hidden[{l}] max_abs={err:.4g} ref_max={ref[l].abs().max().item():.4g} cos={cos:.6f}")
        worst = max(worst, 1 - cos)
        assert cos > 0.999, f"hidden state {l}"
    ref_f = M.fuse_features(ref)
    cosr = torch.nn.functional.cosine_similarity(fused.float().cpu(), ref_f, dim=-1)
    print("fused rows min cos", cosr.min().item())
    assert cosr.min() > 0.999


def test_vit_debug_gemm_path_agrees(vision_path):
    tower = vision_path.model.vision_tower
    px = torch.randn((1, 3, 336, 336), device="cuda")
    patches = tower._patches_of(px)
    a = tower.encode_patches(patches).float()
    tower.force_simple_gemm = True
    try:
        b = tower.encode_patches(patches).float()
    finally:
        tower.force_simple_gemm = False
    cos = torch.nn.functional.cosine_similarity(a, b, dim=-1).min().item()
    print("tcgen05 vs CUDA-core GEMM tower, min row cos", cos)
    assert cos > 0.999


def _attn_ref(q, k, v, scale, mult=None):
    s = (q.float() @ k.float().t()) * scale
    p = torch.exp(s - s.max(-1, keepdim=True).values)
    if mult is not None:
        p = p * mult[None]
    return (p / p.sum(-1, keepdim=True)) @ v.float()


def test_qformer_matches_oracle(vision_path, seeded_weights):
    """Q-Former alone on random features: no text, and the reference's dense text form."""
    from oracle import model as M
    proj = vision_path.model.mm_projector
    T, Ltxt = 2, 21
    feats = _rand((T, 576, 5120), 1.0, 3)
    text = _rand((T, Ltxt, 4096), 0.02, 4)
    text[1, 15:] = 0  # zero-padded rows as produced by vis_zephyr_arch.py:181-186
    with torch.no_grad():
        ref_nt = M.qformer_forward(seeded_weights["qf"], feats.float().cpu(), None)
        ref_t = M.qformer_forward(seeded_weights["qf"], feats.float().cpu(), text.float().cpu())
    got_nt = proj(feats, None).float().cpu()
    got_t = proj(feats, text).float().cpu()
    for name, got, ref in (("no-text", got_nt, ref_nt), ("text", got_t, ref_t)):
        cos = torch.nn.functional.cosine_similarity(got, ref, dim=-1)
        err = (got - ref).abs().max().item()
        print(f"qformer {name}: min cos {cos.min().item():.6f} max_abs {err:.4g}")
        assert cos.min() > 0.999 and err < 0.15, name   # stated tolerance, see tests/test_gpu_e2e.py
    # text conditioning must matter, otherwise the test above proves nothing
    assert (ref_t - ref_nt).abs().max() > 1e-3


@pytest.mark.parametrize("L", [31, 32, 33, 512, 2047])
def test_qformer_long_ragged_text_matches_oracle(vision_path, seeded_weights, L):
    """BASELINE config 5 regime: the packed text form (one row set per SAMPLE, zero padding collapsed to one
    key with multiplicity L - S') against the oracle run LITERALLY on all 32 + L rows of every tile.
    L = 31 / 32 / 33 straddle the 32-key chunk of the block-0 self-attention (own 32 rows + text + pad key),
    512 and 2047 walk 17 / 65 chunks across both segment boundaries; sample 1 keeps a third of the rows, so
    its pad key carries a multiplicity of up to 1 365; 3 tiles share the 2 samples' text K/V."""
    from oracle import model as M
    from vision_zephyr_b200.projector import TextPack
    proj = vision_path.model.mm_projector
    tiles = [1, 2]
    T = sum(tiles)
    s_len = [L, max(L // 3, 1)]
    feats = _rand((T, 576, 5120), 1.0, 30 + L)
    rows = [_rand((n, 4096), 0.02, 40 + L + i) for i, n in enumerate(s_len)]
    dense = torch.zeros((T, L, 4096), dtype=torch.bfloat16, device="cuda")
    t0 = 0
    for smp, nt in enumerate(tiles):
        dense[t0:t0 + nt, :s_len[smp]] = rows[smp][None]
        t0 += nt
    with torch.no_grad():
        ref = M.qformer_forward(seeded_weights["qf"], feats.float().cpu(), dense.float().cpu())
    R = sum(s_len)
    emb = torch.cat(rows + [torch.zeros((1, 4096), dtype=torch.bfloat16, device="cuda")])
    off = torch.tensor([0, s_len[0], R], dtype=torch.int32, device="cuda")
    tile_sample = torch.tensor([0, 1, 1], dtype=torch.int32, device="cuda")
    got = proj.forward_packed(feats, TextPack(emb, off, R, 2, L, tile_sample)).float().cpu()
    got_dense = proj(feats, dense).float().cpu()          # reference signature: every tile its own sample
    for name, g in (("packed", got), ("dense", got_dense)):
        cos = torch.nn.functional.cosine_similarity(g, ref, dim=-1)
        err = (g - ref).abs().max().item()
        print(f"qformer L={L} {name}: min cos {cos.min().item():.6f} max_abs {err:.4g}")
        assert cos.min() > 0.999 and err < 0.15, (L, name)
    # tiles 1 and 2 share a sample but see different features; the zero-pad tail must matter for sample 1
    if L >= 32:
        with torch.no_grad():
            short = M.qformer_forward(seeded_weights["qf"], feats[1:2].float().cpu(), dense[1:2, :s_len[1]].float().cpu())
        assert (short - ref[1:2]).abs().max() > 1e-3


@pytest.mark.parametrize("impl", [1, 0])
def test_vit_attention_kernels_match_torch(impl):
    """both attention kernels (tcgen05 = 1, legacy mma.sync = 0) against fp32 torch attention."""
    import vision_zephyr_b200  # noqa: F401
    from vision_zephyr_b200 import _lib as L
    lib = L.load()
    T = 3
    qkv = _rand((T * 577, 3072), 1.0, 31)
    qkv[:, :2048] *= 1.7          # sharper softmax
    out = torch.full((T * 577, 1024), float("nan"), dtype=torch.bfloat16, device="cuda")
    if impl == 1:
        L.check(lib.vz_vit_attention(L.ptr(qkv), L.ptr(out), T, impl, L.stream_ptr()), "vit attention")
    else:
        # the first (mma.sync) implementation: an independent cross-check kept OUT of the product library
        import ctypes
        import os
        assert lib.vz_vit_attention(L.ptr(qkv), L.ptr(out), T, 0, L.stream_ptr()) == -2      # VZ_ERR_UNSUPPORTED
        tl = ctypes.CDLL(os.path.join(os.path.dirname(L.LIB_PATH), "libvz_b200_testonly.so"))
        tl.vz_test_vit_attention_legacy.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int, ctypes.c_void_p]
        L.check(tl.vz_test_vit_attention_legacy(L.ptr(qkv), L.ptr(out), T, L.stream_ptr()), "legacy vit attention")
    torch.cuda.synchronize()
    q, k, v = (qkv.float().view(T, 577, 3, 16, 64)[:, :, i].transpose(1, 2) for i in range(3))
    ref = torch.softmax((q @ k.transpose(-1, -2)) * 0.125, dim=-1) @ v
    ref = ref.transpose(1, 2).reshape(T * 577, 1024)
    err = (out.float() - ref).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(out.float(), ref, dim=-1).min().item()
    print(f"vit attention impl={impl}: max_abs_err {err:.4g} (ref max {ref.abs().max().item():.3g}) min row cos {cos:.6f}")
    assert torch.isfinite(out.float()).all()
    assert err < 0.03 and cos > 0.9995


@pytest.mark.parametrize("T", [1, 2, 7, 40])
def test_vit_attention_persistent_schedule(T, impl=1):
    """the persistent tcgen05 kernel at tile counts where the 80 T work items are fewer than, not a multiple of, and
    far more than the 296 resident CTAs (item hand-over, deferred epilogues, Q double buffering), with score outliers
    that force accumulator rescales late in a row, and rows of the next tile behind every tile's 577th key"""
    import vision_zephyr_b200  # noqa: F401
    from vision_zephyr_b200 import _lib as L
    lib = L.load()
    qkv = _rand((T * 577, 3072), 1.0, 300 + T)
    qkv[:, :2048] *= 1.7
    qkv[5::97, 1024:2048] *= 6.0          # a few keys with very large scores, scattered over the blocks
    out = torch.full((T * 577, 1024), float("nan"), dtype=torch.bfloat16, device="cuda")
    L.check(lib.vz_vit_attention(L.ptr(qkv), L.ptr(out), T, impl, L.stream_ptr()), "vit attention")
    torch.cuda.synchronize()
    worst = 0.0
    for t0 in range(0, T, 8):                                  # reference in chunks of 8 tiles
        t1 = min(T, t0 + 8)
        x = qkv[t0 * 577:t1 * 577].float().view(t1 - t0, 577, 3, 16, 64)
        q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
        ref = (torch.softmax((q @ k.transpose(-1, -2)) * 0.125, dim=-1) @ v).transpose(1, 2).reshape(-1, 1024)
        got = out[t0 * 577:t1 * 577].float()
        assert torch.isfinite(got).all()
        worst = max(worst, (got - ref).abs().max().item())
    print(f"T={T}: max_abs_err {worst:.4g}")
    assert worst < 0.06
