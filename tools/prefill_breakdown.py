"""Stage-by-stage CUDA-event timing of ONE decoder layer of the native prefill (mistral_prefill.py) on the
config-5 row count (8 970 packed rows of 8 samples), L2 flushed between stages, plus the whole 32-layer call.

    python tools/prefill_breakdown.py [--rows-from-c5]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

LENS = [514, 1645, 1166, 2140, 645, 1557, 815, 488]      # config 5 of bench.py: text length - 1 + 160 visual rows


def main():
    from flash_attn import flash_attn_varlen_func
    from transformers import MistralConfig, MistralModel
    from vision_zephyr_b200 import _lib
    from vision_zephyr_b200.gemm import gemm
    from vision_zephyr_b200.mistral_prefill import ACT_SWIGLU, MistralPrefillB200
    dev = "cuda"
    layers = int(os.environ.get("LAYERS", "32"))
    cfg = MistralConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=layers, num_attention_heads=32,
                        num_key_value_heads=8, vocab_size=32000, rms_norm_eps=1e-5, sliding_window=None)
    torch.manual_seed(0)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        m = MistralModel(cfg)
    torch.set_default_dtype(old)
    m.eval().requires_grad_(False)
    eng = MistralPrefillB200(m)
    lib = _lib.load()
    M, H, I = sum(LENS), 4096, 14336
    x = torch.randn((M, H), device=dev).to(torch.bfloat16)
    cu = torch.tensor([0] + list(__import__("itertools").accumulate(LENS)), dtype=torch.int32, device=dev)
    pos = torch.cat([torch.arange(n, dtype=torch.int32, device=dev) for n in LENS])
    ws = eng._buffers(M)
    qkv, act, (h_a, h_b), S, cs = ws["qkv"], ws["act"], ws["h"], ws["stats"], ws["cs"]
    st = _lib.stream_ptr()
    lib.vz_row_stats(x.data_ptr(), H, M, H, S.data_ptr(), st)
    lib.vz_rope_table(pos.data_ptr(), M, eng.inv_freq.data_ptr(), 64, cs.data_ptr(), st)
    L = eng.layers[0]
    np_h = lib.vz_gemm_stats_partials(M, H)
    stats0 = S[:, :1].clone().reshape(M, 2).contiguous()
    q_v = qkv[:M, :4096].view(M, 32, 128)
    k_v = qkv[:M, 4096:5120].view(M, 8, 128)
    v_v = qkv[:M, 5120:].view(M, 8, 128)
    a = torch.randn((M, 4096), device=dev).to(torch.bfloat16)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    items, n_items, attn_flops = eng.attention_items(LENS)
    attn_out = ws["attn"]
    stages = {
        "qkv_gemm (rmsnorm fused)": (lambda: gemm(x, L.w_qkv, M=M, N=6144, K=H, lda=H, ldw=H, out=qkv, ldo=6144, ln_stats=stats0,
                                                  ln_np=1, ln_eps=1e-5, ln_rms=True, sk_ws=ws["sk"]), 2.0 * M * 6144 * H),
        "rope": (lambda: lib.vz_rope_apply(qkv.data_ptr(), 6144, M, 40, 128, cs.data_ptr(), st), 0.0),
        "attention (vz_attn_causal, tcgen05)": (lambda: lib.vz_attn_causal(qkv.data_ptr(), 6144, M, attn_out.data_ptr(), 4096,
                                                                        items.data_ptr(), n_items, 32, 8, 128, 128 ** -0.5,
                                                                        attn_flops, st), sum(n * n for n in LENS) * 8192.0),
        "attention (flash-attn 2 varlen)": (lambda: flash_attn_varlen_func(q_v, k_v, v_v, cu, cu, max(LENS), max(LENS), causal=True),
                                            sum(n * n for n in LENS) * 8192.0 / 2 * 2),
        "o_proj (+residual +stats)": (lambda: gemm(a, L.w_o, M=M, N=H, K=H, lda=H, ldw=H, out=h_a, ldo=H, residual=x, ldr=H,
                                                   stats_out=S, stats_np=np_h, sk_ws=ws["sk"]), 2.0 * M * H * H),
        "gate_up (rmsnorm + swiglu fused)": (lambda: gemm(h_a, L.w_gu, M=M, N=2 * I, K=H, lda=H, ldw=H, out=act, ldo=I, act=ACT_SWIGLU,
                                                          ln_stats=S, ln_np=np_h, ln_eps=1e-5, ln_rms=True, sk_ws=ws["sk"]), 4.0 * M * I * H),
        "down_proj (+residual +stats)": (lambda: gemm(act, L.w_d, M=M, N=H, K=I, lda=I, ldw=I, out=h_b, ldo=H, residual=h_a, ldr=H,
                                                      stats_out=S, stats_np=np_h, sk_ws=ws["sk"]), 2.0 * M * I * H),
    }
    out = {"rows": M, "lens": LENS, "stages": {}}
    for name, (fn, flops) in stages.items():
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        out["stages"][name] = {"us": ms * 1e3, "tflops": flops / (ms * 1e-3) / 1e12 if flops else None}
    out["layer_sum_us"] = sum(v["us"] for k, v in out["stages"].items() if "flash-attn" not in k)
    with torch.no_grad():
        for _ in range(2):
            eng.forward_packed(x, pos, cu, max(LENS))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.forward_packed(x, pos, cu, max(LENS))
        e1.record()
        torch.cuda.synchronize()
        out["forward_packed_ms"] = e0.elapsed_time(e1) / 3
        out["layers"] = layers
    print(json.dumps(out))


if __name__ == "__main__":
    main()
