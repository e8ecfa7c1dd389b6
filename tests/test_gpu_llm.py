"""The path behind its real callers: INTEGRATION.md section 1 applied literally to HF Mistral
(vision-zephyr_b200/language_model.py = the reference's language_model/vis_zephyr.py with the B200 mixins as
bases), a 2-layer random-init LLM with Zephyr's hidden size.  Checks forward, generate,
initialize_vision_modules (vis_zephyr_arch.py:49-102) and the mm_projector.bin save -> load round trip
(train/vis_zephyr_trainer.py:326-343)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from helpers import PINPOINTS_C3, cos_rows, synth_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def clip_dir(tmp_path_factory, seeded_weights):
    """a local HF checkpoint directory holding the seeded CLIP weights (no hub on the box)"""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    d = str(tmp_path_factory.mktemp("clip336"))
    cfg = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                           image_size=336, patch_size=14, projection_dim=768, hidden_act="quick_gelu",
                           layer_norm_eps=1e-5)
    with torch.device("meta"):
        hf = CLIPVisionModel(cfg)
    hf = hf.to_empty(device="cpu")
    hf.load_state_dict(seeded_weights["clip"], strict=False)
    hf.save_pretrained(d)
    return d


@pytest.fixture(scope="module")
def llm(clip_dir, seeded_weights):
    import vision_zephyr_b200  # noqa: F401
    from vision_zephyr_b200.language_model import VisZephyrB200ForCausalLM, random_mistral_config
    cfg = random_mistral_config(num_hidden_layers=2, intermediate_size=1024, mm_vision_tower=clip_dir,
                                mm_grid_pinpoints=str(PINPOINTS_C3))
    torch.manual_seed(0)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device("cuda"):
            model = VisZephyrB200ForCausalLM(cfg)
    finally:
        torch.set_default_dtype(old)
    model.eval().requires_grad_(False)
    inner = model.get_model()
    assert not inner.get_vision_tower().is_loaded                           # delay_load, like the reference
    inner.get_vision_tower().to("cuda")
    inner.get_vision_tower().load_model()                                   # reads the local checkpoint directory
    inner.mm_projector.load_state_dict({k: v.to(torch.bfloat16) for k, v in seeded_weights["qf"].items()})
    with torch.no_grad():
        inner.embed_tokens.weight.copy_(seeded_weights["embed"].to(torch.bfloat16))
    return model


def _inputs(golden_dir):
    import vision_zephyr_b200 as vz
    lut = np.load(f"{golden_dir}/golden_pixels.npz")["lut"]
    g = np.load(f"{golden_dir}/golden_model.npz")
    imgs = [torch.from_numpy(synth_image(0, 1000, 900)).cuda(), torch.from_numpy(synth_image(1, 637, 336)).cuda()]
    pb = vz.process_any_resolution_images(imgs, PINPOINTS_C3, lut, out_mode="patches")
    ids, mask, labels = (torch.from_numpy(g[k]).cuda() for k in ("c3_ids", "c3_mask", "c3_labels"))
    return g, pb, ids, mask, labels, [(1000, 900), (637, 336)]


def test_forward_through_the_mixin_matches_the_reference_and_plain_mistral(llm, golden_dir):
    from transformers import MistralForCausalLM
    g, pb, ids, mask, labels, sizes = _inputs(golden_dir)
    with torch.no_grad():
        out = llm(input_ids=ids, attention_mask=mask, labels=labels, images=pb, images_size=sizes)
        r = llm.prepare_inputs_labels_for_multimodal(ids, None, mask, None, labels, pb, sizes)
        plain = MistralForCausalLM.forward(llm, attention_mask=r[2], inputs_embeds=r[4], labels=r[5])
    assert list(r[4].shape) == g["c3_embeds_shape"].tolist()
    assert np.array_equal(r[5].cpu().numpy(), g["c3_out_labels"]) and np.array_equal(r[2].cpu().numpy(), g["c3_out_mask"])
    emb = r[4].float().cpu().numpy()
    got = np.concatenate([emb[0, 5:165].reshape(5, 32, 4096), emb[1, 20:116].reshape(3, 32, 4096)])
    ref = g["c3_vis"].astype(np.float32)
    assert cos_rows(got, ref).min() >= 0.999 and np.abs(got - ref).max() <= 0.15
    assert out.logits.shape == (2, r[4].shape[1], 32000) and torch.isfinite(out.loss)
    assert torch.equal(out.logits, plain.logits) and torch.equal(out.loss, plain.loss)


def test_generate_hands_the_spliced_prompt_to_hf_generate(llm, golden_dir):
    from transformers import MistralForCausalLM
    g, pb, ids, mask, labels, sizes = _inputs(golden_dir)
    gen = dict(max_new_tokens=4, do_sample=False, pad_token_id=2)
    toks = llm.generate(ids, images=pb, images_size=sizes, attention_mask=mask, **gen)
    assert toks.shape == (2, 4)
    with torch.no_grad():
        r = llm.prepare_inputs_labels_for_multimodal(ids, None, mask, None, None, pb, sizes)
    ref = MistralForCausalLM.generate(llm, inputs_embeds=r[4], attention_mask=r[2], **gen)
    assert torch.equal(toks, ref)
    # text-only generation embeds the ids itself (vis_zephyr.py:137-139)
    t2 = llm.generate(ids.clamp(min=3)[:, :16], attention_mask=torch.ones_like(ids[:, :16]), **gen)
    assert t2.shape == (2, 4)


def test_initialize_vision_modules_and_mm_projector_round_trip(llm, clip_dir, tmp_path, golden_dir):
    from vision_zephyr_b200.language_model import mm_adapter_state, save_mm_projector
    g, pb, ids, mask, labels, sizes = _inputs(golden_dir)
    inner = llm.get_model()
    with torch.no_grad():
        before = llm.prepare_inputs_labels_for_multimodal(ids, None, mask, None, labels, pb, sizes)[4].clone()
    path = str(tmp_path / "mm_projector.bin")
    save_mm_projector(llm, path)
    saved = mm_adapter_state(llm)
    assert all(k.startswith("model.mm_projector.") for k in saved) and len(saved) == 165
    with torch.no_grad():                                        # wreck the projector, then restore it by name
        for p in inner.mm_projector.parameters():
            p.mul_(0.5)
        wrecked = llm.prepare_inputs_labels_for_multimodal(ids, None, mask, None, labels, pb, sizes)[4]
    assert not torch.equal(before, wrecked)
    args = SimpleNamespace(mm_vision_tower=clip_dir, mm_vision_select_layer="-2,-5,-8,-11,6",
                           mm_vision_select_feature="patch", pretrain_mm_mlp_adapter=path, mm_patch_merge_type="flat",
                           mm_grid_pinpoints=str(PINPOINTS_C3), image_aspect_ratio="anyres")
    inner.initialize_vision_modules(args)                        # existing tower: load_model() again; projector: load by name
    assert inner.config.mm_hidden_size == 5120 and inner.config.use_mm_proj is True
    assert all(p.requires_grad for p in inner.mm_projector.parameters())   # un-frozen, as the reference does
    inner.mm_projector.requires_grad_(False)
    for k, v in inner.mm_projector.state_dict().items():
        assert torch.equal(v.cpu(), saved["model.mm_projector." + k]), k
    with torch.no_grad():
        after = llm.prepare_inputs_labels_for_multimodal(ids, None, mask, None, labels, pb, sizes)[4]
    assert torch.equal(before, after)


def test_projector_switches_to_the_autograd_path_when_gradients_are_on(llm):
    """the kernels see detached weights; with gradients required the module must not hand back a
    result without grad_fn (stage-1 training would silently stop updating mm_projector)."""
    proj = llm.get_model().mm_projector
    feats = torch.zeros((1, 576, 5120), dtype=torch.bfloat16, device="cuda")
    proj.requires_grad_(True)
    try:
        y = proj(feats, None)
        assert y.shape == (1, 32, 4096) and y.requires_grad and y.grad_fn is not None
        with torch.no_grad():
            y0 = proj(feats, None)
        assert not y0.requires_grad and (y.float() - y0.float()).abs().max() < 0.1
    finally:
        proj.requires_grad_(False)


def test_first_layer_rmsnorm_qkv_fused_with_the_splice(llm, golden_dir):
    """SURVEY.md 8(f) rank 3, first half: the scatter's row statistics + ONE tcgen05 GEMM == HF Mistral's
    input_layernorm followed by q_proj / k_proj / v_proj on the spliced rows (padding rows included)."""
    from vision_zephyr_b200.language_model import FirstLayerQKV
    g, pb, ids, mask, labels, sizes = _inputs(golden_dir)
    layer = llm.get_model().layers[0]
    with torch.no_grad():
        layer.input_layernorm.weight.copy_(1 + 0.1 * torch.randn_like(layer.input_layernorm.weight))
    llm.config.vz_first_layer_stats = True
    try:
        with torch.no_grad():
            r = llm.prepare_inputs_labels_for_multimodal(ids, None, mask, None, labels, pb, sizes)
    finally:
        llm.config.vz_first_layer_stats = False
    emb = r[4]
    stats = emb.vz_row_sumsq
    assert stats.shape == (emb.shape[0], emb.shape[1], 2) and float(stats[..., 0].abs().max()) == 0.0
    ss_ref = emb.float().pow(2).sum(-1)
    assert torch.allclose(stats[..., 1], ss_ref, rtol=1e-5, atol=1e-6)
    q, k, v = FirstLayerQKV(layer)(emb)
    with torch.no_grad():
        x32 = emb.float()
        eps = layer.input_layernorm.variance_epsilon
        h = x32 * torch.rsqrt(x32.pow(2).mean(-1, keepdim=True) + eps) * layer.input_layernorm.weight.float()
        att = layer.self_attn
        refs = [h @ p.weight.float().t() for p in (att.q_proj, att.k_proj, att.v_proj)]
    for name, got, ref in zip("qkv", (q, k, v), refs):
        assert got.shape == ref.shape
        err = (got.float() - ref).abs().max().item()
        cos = torch.nn.functional.cosine_similarity(got.float().flatten(), ref.flatten(), dim=0).item()
        print(f"fused first-layer {name}: max_abs {err:.4g} (ref max {ref.abs().max().item():.3g}) cos {cos:.6f}")
        assert cos >= 0.9999 and err <= 0.02 * max(1.0, ref.abs().max().item())
    # padding rows (all-zero embeddings) come out as exact zeros, like rmsnorm(0) W^T
    pad = (r[2] == 0)
    assert float(q[pad].abs().max()) == 0.0


def test_native_prefill_behind_the_splice_matches_hf_and_feeds_hf_decode(llm, golden_dir):
    """config.vz_native_prefill: the decoder stack of a sequence-starting, gradient-free forward runs on packed rows
    (mistral_prefill.py); logits, loss and the KV cache HF's decode steps continue from agree with HF Mistral."""
    g, pb, ids, mask, labels, sizes = _inputs(golden_dir)
    cfg = llm.config
    with torch.no_grad():
        ref = llm(input_ids=ids, attention_mask=mask, labels=labels, images=pb, images_size=sizes, use_cache=True)
        r = llm.prepare_inputs_labels_for_multimodal(ids, None, mask, None, labels, pb, sizes)
    keep = r[2].bool()
    launches0 = __import__("vision_zephyr_b200")._lib.load().vz_kernel_launches()
    cfg.vz_native_prefill = True
    cfg.vz_first_layer_stats = True
    try:
        with torch.no_grad():
            got = llm(input_ids=ids, attention_mask=mask, labels=labels, images=pb, images_size=sizes, use_cache=True)
            launches1 = __import__("vision_zephyr_b200")._lib.load().vz_kernel_launches()
            # a decode step on top of each cache: HF Mistral both times (cached tokens > 0)
            nxt = torch.full((2, 1), 5, dtype=torch.long, device="cuda")
            m2 = torch.cat([r[2], torch.ones((2, 1), dtype=r[2].dtype, device="cuda")], 1)
            pos = r[2].sum(1, keepdim=True)
            d_ref = llm(input_ids=nxt, attention_mask=m2, position_ids=pos, past_key_values=ref.past_key_values, use_cache=True)
            d_got = llm(input_ids=nxt, attention_mask=m2, position_ids=pos, past_key_values=got.past_key_values, use_cache=True)
            toks = llm.generate(ids, images=pb, images_size=sizes, attention_mask=mask, max_new_tokens=4,
                                do_sample=False, pad_token_id=2)
    finally:
        cfg.vz_native_prefill = False
        cfg.vz_first_layer_stats = False
    # 2 layers x (4 GEMMs + rope) + table / gather / scatter kernels were this library's launches (the tower and the
    # projector replay from CUDA graphs at this size and are not counted)
    assert launches1 - launches0 >= 2 * 5 + 3
    a, b = got.logits[keep].float().cpu().numpy(), ref.logits[keep].float().cpu().numpy()
    assert cos_rows(a, b).min() >= 0.999
    assert abs(got.loss.item() - ref.loss.item()) <= 0.02 * abs(ref.loss.item())
    a, b = d_got.logits[:, -1].float().cpu().numpy(), d_ref.logits[:, -1].float().cpu().numpy()
    assert cos_rows(a, b).min() >= 0.999
    assert toks.shape == (2, 4)
