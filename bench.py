#!/usr/bin/env python
"""bench.py -- anyres images/s through the image -> LLM-embedding path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

Main workload (config.workload = "c3_anyres_b8", BASELINE configs 3 / 4): per GPU, 8 synthetic RGB images
(sizes that all select the 672x672 pinpoint -> 1 global + 2x2 tiles = 5 tiles each, 40 tiles), 64-token
prompts with one <image> placeholder, random-init CLIP ViT-L/14-336 + Q-Former + 32000x4096 embedding table,
'flat' merge.  N GPUs = N such shards (weak scaling): every rank encodes its 8 images, ONE exchange step moves
the projected visual tokens to rank 0, which splices the global batch.

A step = preprocess kernels -> ViT -> fusion -> Q-Former -> (exchange) -> plan / gather / scatter.
`value`   : device-resident inputs (u8 images + ids already in HBM), CUDA-event timed, max over ranks.
`e2e`     : same metric through the public API with HOST buffers: pinned u8 images + ids copied H2D,
            descriptor tables rebuilt, outputs copied D2H, all inside the timed region.
`roofline`: the tcgen05 GEMM kernel (dominant), FLOPs = 2MNK per launch, CUDA events on its stream;
`roofline_extra`: the other kernel families of the step, measured the same way in the same pass.
`cpu_baseline`: the fp32 oracle port of the reference algorithm on the host cores, bounded sample.
`parity_ok` / `config.transport` (N > 1): every rank's projected rows, checksummed locally, are found
            bit-for-bit in rank 0's spliced output, on both transports (outside the timed region).
At N = 1 the same JSON line also carries `workloads` (BASELINE config 1 = single-image latency, config 5 =
S <= 2048 prompts spliced and prefilled through a random-init 32-layer Mistral-7B) and `eager_bf16` (the
reference's own formulation under PyTorch eager bf16 on this GPU: the real bar).  `--workload c5_prefill_b8`
or `--workload c1_latency` make one of those the main line instead.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PINPOINTS = [[336, 672], [672, 336], [672, 672], [336, 1008], [1008, 336]]
IMAGE_SIZES = [(1000, 900), (900, 1000), (1344, 1344), (700, 650), (1000, 900), (800, 760), (1200, 1100), (672, 672)]
IMAGES_PER_GPU, TILES_PER_IMAGE, SEQ = 8, 5, 64
METRIC = "anyres images/sec (ViT-L/14-336 + Q-Former)"
WEIGHT_BYTES = 2 * (303_507_456 - 1024 * 768 - 2 * 1024 + 1_678_428_160)   # bf16 CLIP (no projection head) + Q-Former

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel: ncu over the 178 GEMM launches of
# one step of this same command (profiles/r2_gemm_traffic.md: 22.77 GB read + 8.31 GB written; algorithmic ~27.6 GB)
TRAFFIC_PER_LAUNCH = 1.746e8


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def workload_config(workload):
    """the static description of a workload: identical in the b200 arm and the reference arm"""
    if workload == "c3_anyres_b8":
        return {"workload": "c3_anyres_b8", "images_per_gpu": IMAGES_PER_GPU, "tiles_per_image": TILES_PER_IMAGE,
                "seq_len": SEQ, "merge": "flat"}
    if workload == "c5_prefill_b8":
        return {"workload": "c5_prefill_b8", "images_per_gpu": IMAGES_PER_GPU, "tiles_per_image": TILES_PER_IMAGE,
                "seq_len": "U[256,2047] per sample, right-padded", "merge": "flat",
                "llm": "Mistral-7B geometry, 32 layers, random init (HF sdpa)"}
    return {"workload": "c1_latency", "images_per_gpu": 1, "tiles_per_image": 1, "seq_len": SEQ, "merge": "flat"}


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe), through
    NVML (the same counters nvidia-smi prints) from a background thread every 50 ms."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread = index, [], False, None
        self.t0 = self.t1 = None
        self.max_mhz = None
        self.power_limit_w = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            try:
                self.power_limit_w = pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0
            except Exception:
                self.power_limit_w = None
        except Exception:
            return

        def loop():
            while not self.stop_flag:
                try:
                    sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                    try:
                        rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                    self.rows.append((time.time(), float(sm), int(rs), pw))
                except Exception:
                    pass
                time.sleep(0.05)

        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(1.0)
        rows = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or r[0])]
        if not rows:
            rows = self.rows
        reasons = set()
        for r in rows:
            for name, bit in self.REASONS.items():
                if r[2] & bit:
                    reasons.add(name)
        sm = [r[1] for r in rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max((r[3] for r in rows), default=None), "power_limit_w": self.power_limit_w}


def make_inputs(rank, n_local=IMAGES_PER_GPU):
    imgs = []
    for i in range(n_local):
        W, H = IMAGE_SIZES[i % len(IMAGE_SIZES)]
        imgs.append(np.random.default_rng(1000 + rank * 64 + i).integers(0, 256, (H, W, 3), dtype=np.uint8))
    return imgs


def make_ids(n_samples):
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(3, 32000, (n_samples, SEQ), generator=g)
    ids[:, 10] = -200
    return ids


def make_c5_text(n_samples):
    """S_i ~ U[256,2047], tokens U[3,31999], one -200 at U[1,32], right-padded with pad id 2,
    attention_mask = ids != 2 (as the collator, train/train.py:692), labels = ids with the first third -100."""
    g = torch.Generator().manual_seed(5)
    lens = torch.randint(256, 2048, (n_samples,), generator=g)
    S = int(lens.max())
    ids = torch.full((n_samples, S), 2, dtype=torch.long)
    labels = torch.full((n_samples, S), -100, dtype=torch.long)
    for b in range(n_samples):
        n = int(lens[b])
        ids[b, :n] = torch.randint(3, 32000, (n,), generator=g)
        ids[b, int(torch.randint(1, 33, (1,), generator=g))] = -200
        labels[b, :n] = ids[b, :n]
        labels[b, :n // 3] = -100
    mask = (ids != 2).long()
    return ids, mask, labels, [int(x) for x in lens]


# ------------------------------------------------------------------------------------------------
def run_reference_arm(args):
    """The reference's algorithm on the box's host cores: the pinned fp32 oracle port (the reference
    is pure Python/PyTorch and cannot travel to the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import model as M, pil_ops as P, splice as S, weights
    torch.set_num_threads(os.cpu_count())
    cores = torch.get_num_threads()
    w_clip, w_qf, embed = weights.clip_state_dict(0), weights.qformer_state_dict(1), weights.embed_table(2)
    import vision_zephyr_b200 as vz
    lut = vz.clip_lut()
    ids = make_ids(1)

    def one_image(img):
        px = torch.from_numpy(P.process_any_resolution(img, PINPOINTS, lut))
        with torch.no_grad():
            text = M.text_embeddings_for(ids, [px.shape[0]], embed)
            vis = M.encode_images(w_clip, w_qf, px, text)
        feats = S.process_image_patches([vis.numpy()], [(img.shape[1], img.shape[0])], "flat", [(2, 2)])
        return S.splice(ids.numpy(), None, None, False, embed.numpy(), feats)[0]

    imgs = make_inputs(0)
    t0 = time.perf_counter()
    one_image(imgs[0])
    t_first = time.perf_counter() - t0
    steps, warm = args.steps, max(args.warmup - 1, 0)
    # one image per step keeps K+W steps within minutes (about 3 s per image on 16 cores); only a
    # pathological request (> 15 min) is cut short, and the line then reports the steps actually timed
    if t_first * (steps + warm) > 900.0:
        steps = max(1, int(900.0 / t_first) - warm)
    for i in range(warm):
        one_image(imgs[(i + 1) % len(imgs)])
    t0 = time.perf_counter()
    for i in range(steps):
        one_image(imgs[i % len(imgs)])
    dt = time.perf_counter() - t0
    value = steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value,
            "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": 1000 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config("c3_anyres_b8"),
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} timed steps x 1 anyres image of the workload's 8 (5 tiles, 63 text tokens), "
                                       "fp32 torch oracle; a rate, so comparable with the 8-image step"},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def event_time(fn, steps, warmup=3, flush=None):
    """mean ms per call over `steps` calls (CUDA events on the current stream, sync on both sides)"""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    ts = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def bench_c1_latency(path, lut, dev, steps):
    """BASELINE config 1: ONE 336x336 image, 64-token prompt, through the public API, inputs in HBM.
    Weight-streaming bound: one tile touches all 3.97 GB of bf16 weights once."""
    import vision_zephyr_b200 as vz
    pk, _ = peaks()
    img = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (336, 336, 3), dtype=np.uint8)).to(dev)
    ids = make_ids(1).to(dev)

    def call():
        pb = vz.process_fixed_images([img], lut, out_mode="patches")
        return path.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, pb, [(336, 336)])[4]

    ms = event_time(call, max(steps, 20), warmup=5)
    out = call()
    assert out.shape == (1, SEQ - 1 + 32, 4096)
    gbs = WEIGHT_BYTES / (ms * 1e-3) / 1e9
    return {"ms_per_image": ms, "images_per_s": 1000.0 / ms,
            "roofline": {"bound": "hbm", "what": "bf16 weights streamed once per call (ViT 0.61 GB + Q-Former 3.36 GB)",
                         "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                         "floor_ms": WEIGHT_BYTES / pk["hbm_gbs"] / 1e6},
            "config": workload_config("c1_latency")}


def bench_c5(dev, lut, steps, llm_layers=32):
    """BASELINE config 5: C3's 8 anyres images + prompts of S_i ~ U[256,2047] tokens, spliced and handed to a
    random-init Mistral-7B-geometry LLM THROUGH the reference-shaped caller (VisZephyrB200ForCausalLM.forward ->
    prepare_inputs_labels_for_multimodal -> MistralForCausalLM.forward(inputs_embeds=...))."""
    import vision_zephyr_b200 as vz
    from vision_zephyr_b200.language_model import VisZephyrB200ForCausalLM, random_mistral_config
    from vision_zephyr_b200.runtime import random_init_
    from vision_zephyr_b200 import _lib
    pk, _ = peaks()
    cfg = random_mistral_config(num_hidden_layers=llm_layers)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    t0 = time.time()
    try:
        with torch.device(dev):
            model = VisZephyrB200ForCausalLM(cfg)
    finally:
        torch.set_default_dtype(old)
    model.eval().requires_grad_(False)
    random_init_(model, seed=0)
    build_s = time.time() - t0
    imgs = [torch.from_numpy(x).to(dev) for x in make_inputs(0)]
    sizes = [IMAGE_SIZES[i % len(IMAGE_SIZES)] for i in range(IMAGES_PER_GPU)]
    ids, mask, labels, lens = make_c5_text(IMAGES_PER_GPU)
    ids, mask, labels = ids.to(dev), mask.to(dev), labels.to(dev)
    pb = vz.process_any_resolution_images(imgs, PINPOINTS, lut, out_mode="patches")

    def path_only():
        pbi = vz.process_any_resolution_images(imgs, PINPOINTS, lut, out_mode="patches")
        return model.prepare_inputs_labels_for_multimodal(ids, None, mask, None, labels, pbi, sizes)

    lib = _lib.load()
    with torch.no_grad():
        r = path_only()
        B, Lmax = r[4].shape[:2]
        real = int(r[2].sum())
        assert B == IMAGES_PER_GPU and Lmax == max(lens) - 1 + TILES_PER_IMAGE * 32 and torch.isfinite(r[4].float()).all()
        l0 = lib.vz_kernel_launches()
        ms_path = event_time(path_only, steps, warmup=3)
        launches = int(lib.vz_kernel_launches() - l0) // (steps + 3)

        def prefill():
            return model(input_ids=ids, attention_mask=mask, images=pb, images_size=sizes, use_cache=False, logits_to_keep=1)

        def llm_only():
            return model(inputs_embeds=r[4], attention_mask=r[2], use_cache=False, logits_to_keep=1)

        out = prefill()
        assert out.logits.shape[0] == B and torch.isfinite(out.logits.float()).all()
        ms_total = event_time(prefill, max(3, steps // 4), warmup=1)
        ms_llm = event_time(llm_only, max(3, steps // 4), warmup=1)
        # SURVEY 8(f) rank 3, the hand-over half of it: the splice's per-sample lengths let the LLM skip the padding
        # altogether -- the real rows are packed into ONE sequence with restarting position ids, which HF Mistral
        # runs as variable-length attention (flash-attn 2, library code).  Best effort: reported when the installed
        # flash-attn runs on this GPU.
        packed = {"error": None}
        try:
            keep = r[2].bool()
            emb_p = r[4][keep].unsqueeze(0)
            pos_p = (torch.cumsum(keep.long(), 1) - 1)[keep].unsqueeze(0)
            last = torch.cumsum(keep.sum(1), 0) - 1
            model.set_attn_implementation("flash_attention_2")

            def llm_packed():
                return model(inputs_embeds=emb_p, position_ids=pos_p, use_cache=False, logits_to_keep=last)

            o2 = llm_packed()
            assert o2.logits.shape[:2] == (1, B) and torch.isfinite(o2.logits.float()).all()
            packed["llm_only_ms"] = event_time(llm_packed, max(3, steps // 4), warmup=1)
            packed["tokens"] = int(emb_p.shape[1])
        except Exception as e:
            packed["error"] = f"{type(e).__name__}: {e}"[:300]
        finally:
            try:
                model.set_attn_implementation("sdpa")
            except Exception:
                pass
    # SURVEY 8(f) rank 3, the kernel half: the decoder stack itself on packed rows (mistral_prefill.py: RMSNorm + q|k|v,
    # o_proj + residual, SwiGLU gate|up, down_proj + residual as four tcgen05 GEMMs per layer, RoPE row kernel,
    # flash-attn 2 varlen as the attention core), behind the same caller (config.vz_native_prefill)
    native = {"error": None}
    try:
        with torch.no_grad():
            ref_h = model.model(inputs_embeds=r[4], attention_mask=r[2], use_cache=False).last_hidden_state
            model.config.vz_native_prefill = True
            model.config.vz_first_layer_stats = True
            t0 = time.time()
            model.get_model().native_prefill()
            native["fold_weights_s"] = time.time() - t0

            def prefill_native():
                return model(input_ids=ids, attention_mask=mask, images=pb, images_size=sizes, use_cache=False,
                             logits_to_keep=1)

            def llm_native():
                return model(inputs_embeds=r[4], attention_mask=r[2], use_cache=False, logits_to_keep=1)

            got_h = model.model(inputs_embeds=r[4], attention_mask=r[2], use_cache=False).last_hidden_state
            keep = r[2].bool()
            a, b = got_h[keep].float(), ref_h[keep].float()
            cosr = torch.nn.functional.cosine_similarity(a, b, dim=-1)
            native["parity_vs_hf"] = {"min_row_cosine": float(cosr.min()), "max_abs": float((a - b).abs().max()),
                                      "ref_abs_max": float(b.abs().max()), "rows": int(a.shape[0])}
            # 32 random-init layers in bf16: HF's own bf16 run sits at 0.995 against fp32 (profiles/r2_prefill.md)
            native["parity_ok"] = bool(cosr.min() >= 0.99)
            del ref_h, got_h, a, b
            l0 = lib.vz_kernel_launches()
            o3 = prefill_native()
            native["gpu_launches_per_step"] = int(lib.vz_kernel_launches() - l0)
            assert o3.logits.shape[0] == B and torch.isfinite(o3.logits.float()).all()
            native["prefill_ms_total"] = event_time(prefill_native, max(3, steps // 4), warmup=1)
            native["llm_only_ms"] = event_time(llm_native, max(3, steps // 4), warmup=1)
            native["prefill_tokens_per_s"] = real / native["prefill_ms_total"] * 1e3
            lib.vz_profile(1)
            llm_native()
            prof = _lib.profile_read()
            lib.vz_profile(0)
            n_g, ms_g, fl_g = prof["gemm_bf16_tcgen05"]
            native["gemm"] = {"launches": int(n_g), "ms": ms_g, "tflops": fl_g / (ms_g * 1e-3) / 1e12,
                              "frac_of_sustained_peak": fl_g / (ms_g * 1e-3) / 1e12 / pk.get("bf16_tflops_sustained", pk["bf16_tflops"])}
            if "llm_attn_causal" in prof and prof["llm_attn_causal"][0]:
                n_a, ms_a, fl_a = prof["llm_attn_causal"]
                native["attention"] = {"kernel": "vz_attn_causal (tcgen05, S / P / O in tensor memory)", "launches": int(n_a),
                                       "ms": ms_a, "us_per_layer": ms_a / n_a * 1e3,
                                       "tflops_algorithmic": fl_a / (ms_a * 1e-3) / 1e12}
            native["attention_core"] = model.get_model().native_prefill().attn_impl
            native["other_ms"] = native["llm_only_ms"] - ms_g - (prof.get("llm_attn_causal", (0, 0.0, 0))[1])
    except Exception as e:
        native["error"] = f"{type(e).__name__}: {e}"[:300]
    finally:
        model.config.vz_native_prefill = False
        model.config.vz_first_layer_stats = False
        model.get_model()._vz_prefill = None
        torch.cuda.empty_cache()
    # first LLM layer's RMSNorm + q/k/v projections: HF (RMSNorm kernel chain + three Linears) vs ONE tcgen05 GEMM fed by
    # the scatter's row statistics (language_model.FirstLayerQKV; parity in tests/test_gpu_llm.py)
    first_layer = {}
    try:
        from vision_zephyr_b200.language_model import FirstLayerQKV
        layer0 = model.get_model().layers[0]
        model.config.vz_first_layer_stats = True
        with torch.no_grad():
            r2 = path_only()
        model.config.vz_first_layer_stats = False
        fused = FirstLayerQKV(layer0)

        def hf_qkv():
            h = layer0.input_layernorm(r2[4])
            return layer0.self_attn.q_proj(h), layer0.self_attn.k_proj(h), layer0.self_attn.v_proj(h)

        with torch.no_grad():
            first_layer = {"hf_rmsnorm_plus_3_linears_ms": event_time(hf_qkv, 10, warmup=3),
                           "fused_gemm_ms": event_time(lambda: fused(r2[4]), 10, warmup=3),
                           "rows": int(r2[4].shape[0] * r2[4].shape[1])}
        del r2, fused
    except Exception as e:
        first_layer = {"error": f"{type(e).__name__}: {e}"[:300]}
    # the scatter alone at this geometry (HBM-bound): bytes = SURVEY 8(d)(4)
    from vision_zephyr_b200 import arch
    ctx = model._plan_splice(ids, mask, labels, [TILES_PER_IMAGE] * B, sizes)
    info = model._plan_info(ctx)
    vis = torch.randn((B * TILES_PER_IMAGE * 32, 4096), device=dev).to(torch.bfloat16)
    embed = model.get_model().embed_tokens.weight
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def scatter():
        return arch.splice_scatter(ctx["ids"], ctx["labels"], embed, vis, None, ctx["slots"], ctx["prefix"], B,
                                   ctx["total_vis_rows"], ctx["plan"], info["Lmax"], False)

    lib.vz_profile(1)
    for _ in range(10):
        flush.zero_()
        scatter()
    prof = _lib.profile_read()
    lib.vz_profile(0)
    n_sc, ms_sc, _ = prof["splice_scatter"]
    S = ids.shape[1]
    sc_bytes = sum(info["lengths"]) * 8192 + B * info["Lmax"] * 8192 + 17 * B * S + 24 * B * info["Lmax"]
    sc_gbs = sc_bytes / (ms_sc / n_sc * 1e-3) / 1e9
    del model, flush
    torch.cuda.empty_cache()
    return {"path_ms_per_step": ms_path, "path_images_per_s": B / ms_path * 1e3,
            "prefill_ms_total": ms_total, "llm_only_ms": ms_llm, "path_share_of_prefill": ms_path / ms_total,
            "spliced_tokens": real, "padded_tokens": B * Lmax, "prefill_tokens_per_s": real / ms_total * 1e3,
            "packed_varlen_prefill": packed, "native_prefill": native, "first_layer_qkv": first_layer,
            "Lmax": Lmax, "L_text": max(lens) - 1, "text_lens": lens, "gpu_launches_per_step": launches,
            "llm_build_s": build_s,
            "splice_scatter": {"bound": "hbm", "achieved": sc_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                               "frac": sc_gbs / pk["hbm_gbs"], "bytes": sc_bytes, "us": ms_sc / n_sc * 1e3,
                               "l2": "512 MB flush between launches"},
            "config": workload_config("c5_prefill_b8"),
            "how": "VisZephyrB200ForCausalLM.forward(input_ids, attention_mask, images, images_size, logits_to_keep=1): "
                   "B200 path + HF Mistral (sdpa, bf16); inputs resident in HBM"}


def bench_eager_bf16(path, dev, steps):
    """The reference's own formulation (oracle/model.py = its modules restated) under PyTorch eager bf16 on this
    GPU: cuBLAS + ATen, K/V projections and all 32 + L rows included, on the main workload's shapes (40 tiles,
    L = 63).  ViT + fusion + Q-Former only (no preprocess / splice), i.e. a lower bound on the reference's step."""
    from oracle import model as M
    T, L = IMAGES_PER_GPU * TILES_PER_IMAGE, SEQ - 1
    P = path.model.vision_tower._packed
    one = lambda n: torch.ones(n, device=dev, dtype=torch.bfloat16)
    zero = lambda n: torch.zeros(n, device=dev, dtype=torch.bfloat16)
    p = "vision_model."
    clip = {p + "embeddings.class_embedding": P["class_emb"],
            p + "embeddings.patch_embedding.weight": P["patch_w"][:, :588].reshape(1024, 3, 14, 14).contiguous(),
            p + "embeddings.position_embedding.weight": P["pos_emb"],
            p + "pre_layrnorm.weight": P["pre_ln_g"].bfloat16(), p + "pre_layrnorm.bias": P["pre_ln_b"].bfloat16()}
    for l in range(24):
        q = f"{p}encoder.layers.{l}."
        wq, wk, wv = P[f"{l}.w_qkv"].split(1024, 0)
        bq, bk, bv = P[f"{l}.b_qkv"].bfloat16().split(1024, 0)
        # (the packed weights have the LayerNorm affine folded in; timing only needs the shapes)
        clip.update({q + "self_attn.q_proj.weight": wq, q + "self_attn.k_proj.weight": wk, q + "self_attn.v_proj.weight": wv,
                     q + "self_attn.q_proj.bias": bq, q + "self_attn.k_proj.bias": bk, q + "self_attn.v_proj.bias": bv,
                     q + "self_attn.out_proj.weight": P[f"{l}.w_o"], q + "self_attn.out_proj.bias": P[f"{l}.b_o"].bfloat16(),
                     q + "layer_norm1.weight": one(1024), q + "layer_norm1.bias": zero(1024),
                     q + "layer_norm2.weight": one(1024), q + "layer_norm2.bias": zero(1024),
                     q + "mlp.fc1.weight": P[f"{l}.w_fc1"], q + "mlp.fc1.bias": P[f"{l}.b_fc1"].bfloat16(),
                     q + "mlp.fc2.weight": P[f"{l}.w_fc2"], q + "mlp.fc2.bias": P[f"{l}.b_fc2"].bfloat16()})
    qf = {k: v.detach() for k, v in path.model.mm_projector.state_dict().items()}
    px = torch.randn((T, 3, 336, 336), device=dev, dtype=torch.bfloat16)
    text = (torch.randn((T, L, 4096), device=dev) * 0.02).to(torch.bfloat16)

    def step():
        with torch.no_grad():
            return M.encode_images(clip, qf, px, text)

    ms = event_time(step, max(5, steps // 2), warmup=3)
    return {"value": IMAGES_PER_GPU / ms * 1e3, "unit": "images/s", "ms_per_step": ms,
            "what": "reference formulation (oracle/model.py) under PyTorch eager bf16 on this GPU, ViT + fusion + Q-Former "
                    "of the main workload (40 tiles, L = 63); preprocess and splice not included"}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3_anyres_b8", choices=["c3_anyres_b8", "c5_prefill_b8", "c1_latency"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the c1 / c5 / eager-bf16 side measurements at N = 1")
    ap.add_argument("--llm-layers", type=int, default=32)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    import ctypes as C
    import torch.distributed as dist
    import vision_zephyr_b200 as vz
    from vision_zephyr_b200 import _lib
    from vision_zephyr_b200.dist import shard_images
    from vision_zephyr_b200.runtime import VisionEmbeddingPath, random_init_

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    lut = vz.clip_lut()
    pk, pk_kind = peaks()

    # ---- side workloads as the main line ------------------------------------------------------------
    if args.workload != "c3_anyres_b8":
        if world > 1:
            raise SystemExit("--workload c5_prefill_b8 / c1_latency are single-GPU measurements")
        sampler = ClockSampler(local_rank)
        sampler.start()
        sampler.mark_begin()
        if args.workload == "c1_latency":
            path = random_init_(VisionEmbeddingPath(device=dev), seed=0)
            r = bench_c1_latency(path, lut, dev, args.steps)
            value, ms_step, roof = r["images_per_s"], r["ms_per_image"], r["roofline"]
        else:
            r = bench_c5(dev, lut, args.steps, args.llm_layers)
            value, ms_step = r["path_images_per_s"], r["path_ms_per_step"]
            roof = r["splice_scatter"]
        sampler.mark_end()
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": 1, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": r["config"], "detail": r,
                "roofline": roof, "clocks": sampler.stop()}
        print(json.dumps(line), flush=True)
        return

    path = random_init_(VisionEmbeddingPath(device=dev), seed=0)
    n_global = IMAGES_PER_GPU * world
    tiles_global = [TILES_PER_IMAGE] * n_global
    sizes_global = [IMAGE_SIZES[i % len(IMAGE_SIZES)] for r in range(world) for i in range(IMAGES_PER_GPU)]
    lo, hi = shard_images(tiles_global, world)[rank]
    assert (lo, hi) == (rank * IMAGES_PER_GPU, (rank + 1) * IMAGES_PER_GPU)
    host_imgs = [torch.from_numpy(x).pin_memory() for x in make_inputs(rank)]
    ids_host = make_ids(n_global).pin_memory()
    dev_imgs = [x.to(dev) for x in host_imgs]
    ids_dev = ids_host.to(dev)
    from vision_zephyr_b200.anyres import anyres_views
    from vision_zephyr_b200.preprocess import PatchBatch, build_plan, run_plan
    views = [anyres_views((int(im.shape[1]), int(im.shape[0])), PINPOINTS)[0] for im in dev_imgs]
    pre_plan = build_plan(dev_imgs, views, lut)
    assert pre_plan.tiles_per_image == [TILES_PER_IMAGE] * IMAGES_PER_GPU, pre_plan.tiles_per_image

    def step_device(keep_local=False):
        """inputs resident in HBM; preprocess descriptors (pure geometry) prebuilt"""
        patches = run_plan(pre_plan, "patches")
        pb = PatchBatch(patches, pre_plan.tiles_per_image, pre_plan.image_sizes)
        if world > 1:
            return path.prepare_inputs_labels_for_multimodal_sharded(ids_dev, None, None, None, None, pb, tiles_global,
                                                                     sizes_global, keep_local=keep_local)
        return path.prepare_inputs_labels_for_multimodal(ids_dev, None, None, None, None, pb, sizes_global)

    out_host = {}
    # e2e pipeline (user-level code around the public API): the H2D copy of step k + 1 and the D2H read of
    # step k - 1 run on their own streams while step k computes; every step still copies ITS inputs from
    # pinned host memory and reads ITS result back.  Two sets of device input buffers alternate.
    h2d_stream, d2h_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    staged = [None, None]          # per buffer set: (device images, device ids, ready event)
    set_free = [None, None]        # event: the compute that read this buffer set has finished
    e2e_state = {"k": 0, "d2h_done": None}

    def stage_inputs(k):
        """enqueue the host->device copies of step k on the copy stream"""
        s = k & 1
        with torch.cuda.stream(h2d_stream):
            if set_free[s] is not None:
                h2d_stream.wait_event(set_free[s])
            if staged[s] is None:
                imgs = [torch.empty_like(x, device=dev) for x in host_imgs]
                ids = torch.empty_like(ids_host, device=dev)
            else:
                imgs, ids, _ = staged[s]
            for d, h in zip(imgs, host_imgs):
                d.copy_(h, non_blocking=True)
            ids.copy_(ids_host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(h2d_stream)
        staged[s] = (imgs, ids, ev)

    def step_e2e():
        """public API from HOST buffers: H2D of images + ids, descriptor build, kernels, D2H of the result"""
        k = e2e_state["k"]
        cur = torch.cuda.current_stream()
        if staged[k & 1] is None or k == 0:
            stage_inputs(k)
        imgs, ids, ready = staged[k & 1]
        stage_inputs(k + 1)                       # next step's inputs travel while this step computes
        cur.wait_event(ready)
        pb = vz.process_any_resolution_images(imgs, PINPOINTS, lut, out_mode="patches")
        if world > 1:
            r = path.prepare_inputs_labels_for_multimodal_sharded(ids, None, None, None, None, pb, tiles_global, sizes_global)
        else:
            r = path.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, pb, sizes_global)
        done = torch.cuda.Event()
        done.record(cur)
        set_free[k & 1] = done
        if r[4] is not None:
            if "emb" not in out_host:
                out_host["emb"] = [torch.empty(r[4].shape, dtype=r[4].dtype).pin_memory() for _ in range(2)]
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                out_host["emb"][k & 1].copy_(r[4], non_blocking=True)
                r[4].record_stream(d2h_stream)
        e2e_state["k"] = k + 1
        return r

    def e2e_drain():
        h2d_stream.synchronize()
        d2h_stream.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up ------------------------------------------------------------------------------
    for _ in range(args.warmup):
        out = step_device()
    barrier()
    if rank == 0:
        assert out[4].shape == (n_global, SEQ - 1 + TILES_PER_IMAGE * 32, 4096), out[4].shape
        assert torch.isfinite(out[4].float()).all()

    # ---- multi-GPU correctness, outside the timed region ---------------------------------------------
    # Every rank checksums the rows ITS projector produced (fp32 sum + first and last row); rank 0 must find
    # exactly those rows at their place in the spliced global batch -- on the default transport (peer stores
    # when available) and on the NCCL all-gather, whose outputs must also be bit-identical to each other.
    transport, parity = "none", None
    if world > 1:
        D = 4096
        outs, parity = {}, {}
        env_before = os.environ.get("VZ_PEER_GATHER")
        for env in (env_before, "0"):
            if env is None:
                os.environ.pop("VZ_PEER_GATHER", None)
            else:
                os.environ["VZ_PEER_GATHER"] = env
            r = step_device(keep_local=True)
            name = path.last_transport
            loc = path.last_local_tokens
            chk = torch.cat([loc.float().sum().reshape(1), loc[0].float(), loc[-1].float()]).contiguous()
            allchk = torch.empty((world, chk.numel()), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(allchk, chk)
            torch.cuda.synchronize()
            if rank == 0:
                emb = r[4]
                ok = True
                for rr in range(world):
                    rows = emb[rr * IMAGES_PER_GPU:(rr + 1) * IMAGES_PER_GPU, 10:10 + TILES_PER_IMAGE * 32].reshape(-1, D)
                    got = torch.cat([rows.float().sum().reshape(1), rows[0].float(), rows[-1].float()])
                    ok = ok and bool(torch.equal(got[1:], allchk[rr, 1:]))
                    ok = ok and abs(float(got[0]) - float(allchk[rr, 0])) <= 1e-6 * max(1.0, abs(float(allchk[rr, 0])))
                parity[name] = ok
                outs[name] = emb.clone()
            if env == env_before:
                transport = name
        if env_before is None:
            os.environ.pop("VZ_PEER_GATHER", None)
        else:
            os.environ["VZ_PEER_GATHER"] = env_before
        if rank == 0 and len(outs) == 2:
            a, b = list(outs.values())
            parity["transports_bit_identical"] = bool(torch.equal(a, b))
        outs = None
        step_device()     # back on the default transport before timing
        barrier()

    # ---- timed region: device-resident inputs -------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.mark_begin()
    l0 = lib.vz_kernel_launches()
    ms_total = timed(step_device, args.steps)
    launches = int(lib.vz_kernel_launches() - l0)
    # ---- same steps again with CUDA events around every kernel launch (rooflines of the kernel families).
    # The two event records per launch (~470 per step) stretch a step by ~5 %, so `value` comes from the
    # clean pass above and the per-launch durations from this instrumented pass of the same K steps.
    lib.vz_profile(1)
    ms_instr = timed(step_device, args.steps)
    prof = _lib.profile_read()
    lib.vz_profile(0)
    if rank == 0:
        sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host buffers --------------------------------------------------------------------
    for _ in range(2):
        step_e2e()
    e2e_drain()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.current_stream().wait_stream(d2h_stream)      # the last result must have reached the host
    t1.record()
    e2e_drain()
    barrier()
    ms_e2e_t = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e_t, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms_e2e_t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms_total / args.steps
    value = n_global / (ms_step / 1000.0)
    e2e_value = n_global / (ms_e2e / args.steps / 1000.0)
    h2d = sum(x.numel() for x in host_imgs) + ids_host.numel() * 8 + pre_plan.h2d_bytes
    d2h = out_host["emb"][0].numel() * 2 + (2 * n_global + 4) * 4
    n_g, g_ms, g_fl = prof.get("gemm_bf16_tcgen05", (0, 0.0, 0.0))
    gemm_tflops = (g_fl / 1e12) / (g_ms / 1e3) if g_ms > 0 else 0.0
    peak_tf = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    hbm = pk["hbm_gbs"]
    # algorithmic bytes the launchers cannot know: preprocess = u8 sources + bf16 patch rows;
    # scatter = rows read + rows written + the integer side arrays
    pre_bytes = sum(x.numel() for x in host_imgs) + pre_plan.n_tiles * 677376
    Lout = SEQ - 1 + TILES_PER_IMAGE * 32
    sc_bytes = n_global * Lout * 8192 * 2 + 17 * n_global * SEQ + 24 * n_global * Lout
    extra = []
    for name, (n, ms, work) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        if name == "gemm_bf16_tcgen05":
            continue
        ent = {"kernel": name, "launches_per_step": n / args.steps, "ms_per_step": ms / args.steps,
               "share_of_step": ms / ms_instr}
        if name == "vit_attn_tc":
            tf = work / 1e12 / (ms / 1e3)
            ent.update(bound="tensor", achieved=tf, peak=peak_tf, unit="TFLOP/s", frac=tf / peak_tf,
                       note="exp2-bound before tensor-bound: 333 k exp2 per (tile, head)")
        elif name == "fuse":
            gb = work / 1e9 / (ms / 1e3)
            ent.update(bound="hbm", achieved=gb, peak=hbm, unit="GB/s", frac=gb / hbm)
        elif name == "splice_scatter":
            gb = sc_bytes * args.steps / 1e9 / (ms / 1e3)
            ent.update(bound="hbm", achieved=gb, peak=hbm, unit="GB/s", frac=gb / hbm,
                       note="L=63 geometry: 14.6 MB per launch, launch-latency sized; config 5 geometry under workloads")
        elif name in ("layernorm", "softmax_rows"):
            gb = work / 1e9 / (ms / 1e3)
            ent.update(bound="hbm", achieved=gb, peak=hbm, unit="GB/s", frac=gb / hbm)
        extra.append(ent)
    pre_ms = sum(prof.get(k, (0, 0.0, 0.0))[1] for k in ("preprocess_h", "preprocess_v", "preprocess_fused"))
    if pre_ms > 0:
        gb = pre_bytes * args.steps / 1e9 / (pre_ms / 1e3)
        extra.append({"kernel": "preprocess (h + v passes)", "ms_per_step": pre_ms / args.steps, "bound": "hbm",
                      "achieved": gb, "peak": hbm, "unit": "GB/s", "frac": gb / hbm, "bytes_per_step": pre_bytes,
                      "share_of_step": pre_ms / ms_instr})
    config = workload_config("c3_anyres_b8")
    config.update({"global_images": n_global, "parallelism": f"dp{world}", "transport": transport,
                   "l2_policy": "no flush needed: every step streams 4.0 GB of weights + 1.2 GB of hidden states + 0.9 GB of activations, far beyond the 126 MB L2",
                   "tiles_per_s": value * TILES_PER_IMAGE})
    line = {
        "metric": METRIC, "value": value, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": config,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e2e / args.steps,
                "how": "public API from pinned host buffers; the copies of step k+1 / k-1 overlap the kernels of step k (2 copy streams)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": "gemm_bf16_tcgen05_kernel", "bound": "tensor", "achieved": gemm_tflops, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": gemm_tflops / peak_tf if peak_tf else None,
                     # dram__bytes_read+write per launch, ncu, mean over the step's 178 GEMM launches
                     # (profiles/r2_gemm_traffic.md); algorithmic bytes of the same launches: ~155 MB per launch
                     "traffic": TRAFFIC_PER_LAUNCH, "traffic_source": "profiles/r2_gemm_traffic.md",
                     "peak_source": f"{pk_kind} bf16_tflops_sustained", "launches": int(n_g),
                     "gemm_ms_per_step": g_ms / args.steps,
                     "measured_in": "second pass of the same K steps with CUDA events around every kernel launch",
                     "instrumented_ms_per_step": ms_instr / args.steps,
                     "gemm_share_of_step": (g_ms / ms_instr) if ms_instr else None},
        "roofline_extra": extra,
    }
    if world > 1:
        line["parity_ok"] = bool(parity) and all(parity.values())
        line["parity"] = parity
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample()
    if world == 1 and not args.no_extras:
        wl = {}
        for name, fn in (("c1_latency", lambda: bench_c1_latency(path, lut, dev, args.steps)),
                         ("eager_bf16", lambda: bench_eager_bf16(path, dev, args.steps))):
            try:
                wl[name] = fn()
            except Exception as e:      # a side measurement must not lose the main line
                wl[name] = {"error": f"{type(e).__name__}: {e}"}
        line["eager_bf16"] = wl.pop("eager_bf16")
        del path
        torch.cuda.empty_cache()
        try:
            wl["c5_prefill_b8"] = bench_c5(dev, lut, args.steps, args.llm_layers)
        except Exception as e:
            wl["c5_prefill_b8"] = {"error": f"{type(e).__name__}: {e}"}
        line["workloads"] = wl
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_sample():
    """oracle port on the host cores, rank 0 at N=1: ONE anyres image (5 tiles, 63 text tokens)."""
    from oracle import model as M, pil_ops as P, splice as S, weights
    import vision_zephyr_b200 as vz
    torch.set_num_threads(os.cpu_count())
    w_clip, w_qf, embed = weights.clip_state_dict(0), weights.qformer_state_dict(1), weights.embed_table(2)
    lut = vz.clip_lut()
    ids = make_ids(1)
    img = make_inputs(0, 1)[0]
    t0 = time.perf_counter()
    px = torch.from_numpy(P.process_any_resolution(img, PINPOINTS, lut))
    with torch.no_grad():
        text = M.text_embeddings_for(ids, [px.shape[0]], embed)
        vis = M.encode_images(w_clip, w_qf, px, text)
    feats = S.process_image_patches([vis.numpy()], [(img.shape[1], img.shape[0])], "flat", [(2, 2)])
    S.splice(ids.numpy(), None, None, False, embed.numpy(), feats)
    dt = time.perf_counter() - t0
    return {"value": 1.0 / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "1 anyres image (1000x900 -> 5 tiles, 63 text tokens), fp32 torch oracle, cold run"}


if __name__ == "__main__":
    main()
