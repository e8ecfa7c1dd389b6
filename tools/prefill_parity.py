"""bf16 noise floor of the 32-layer prefill: HF Mistral (bf16, sdpa, padded) and the native packed prefill
(mistral_prefill.py), each against HF Mistral in fp32 on the same random-init weights and config-5 shaped inputs.

    python tools/prefill_parity.py [--layers 32] [--batch 4]

Prints one JSON line: row cosines (min / mean) and max-abs of the last hidden state at the real positions."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--max-len", type=int, default=2048)
    args = ap.parse_args()
    from transformers import MistralConfig, MistralModel
    from vision_zephyr_b200.mistral_prefill import MistralPrefillB200
    dev = "cuda"
    cfg = MistralConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=args.layers, num_attention_heads=32,
                        num_key_value_heads=8, vocab_size=32000, max_position_embeddings=32768, rms_norm_eps=1e-5,
                        rope_theta=10000.0, sliding_window=None, attn_implementation="sdpa")
    torch.manual_seed(0)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        m = MistralModel(cfg)
    torch.set_default_dtype(old)
    m.eval().requires_grad_(False)
    gen = torch.Generator(device=dev).manual_seed(0)
    B, L = args.batch, args.max_len
    lens = torch.randint(256, L + 1, (B,), generator=torch.Generator().manual_seed(1)).tolist()
    lens[0] = L
    x = m.embed_tokens(torch.randint(3, 32000, (B, L), device=dev, generator=gen))
    mask = torch.zeros((B, L), dtype=torch.long, device=dev)
    for b, n in enumerate(lens):
        mask[b, :n] = 1
    keep = mask.bool()
    with torch.no_grad():
        h_bf16 = m(inputs_embeds=x, attention_mask=mask, use_cache=False).last_hidden_state[keep].float()
        h_native = MistralPrefillB200(m).prefill(x, mask, None, None)[keep].float()
        m32 = m.float()
        h_fp32 = m32(inputs_embeds=x.float(), attention_mask=mask, use_cache=False).last_hidden_state[keep]

    def cmp(a, b):
        c = torch.nn.functional.cosine_similarity(a, b, dim=-1)
        return {"min_row_cosine": float(c.min()), "mean_row_cosine": float(c.mean()), "max_abs": float((a - b).abs().max())}

    print(json.dumps({"layers": args.layers, "rows": int(keep.sum()), "lens": lens, "ref_abs_max": float(h_fp32.abs().max()),
                      "hf_bf16_vs_fp32": cmp(h_bf16, h_fp32), "native_vs_fp32": cmp(h_native, h_fp32),
                      "native_vs_hf_bf16": cmp(h_native, h_bf16)}))


if __name__ == "__main__":
    main()
