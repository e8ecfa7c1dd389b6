"""ncu target: ONE forward_packed of the native prefill (LAYERS decoder layers, default 2) on the config-5 rows."""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LENS = [514, 1645, 1166, 2140, 645, 1557, 815, 488]


def main():
    from transformers import MistralConfig, MistralModel
    from vision_zephyr_b200.mistral_prefill import MistralPrefillB200
    layers = int(os.environ.get("LAYERS", "2"))
    cfg = MistralConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=layers, num_attention_heads=32,
                        num_key_value_heads=8, vocab_size=32000, rms_norm_eps=1e-5, sliding_window=None)
    torch.manual_seed(0)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    with torch.device("cuda"):
        m = MistralModel(cfg)
    torch.set_default_dtype(old)
    m.eval().requires_grad_(False)
    eng = MistralPrefillB200(m)
    M = sum(LENS)
    x = torch.randn((M, 4096), device="cuda").to(torch.bfloat16)
    cu = torch.tensor([0] + list(itertools.accumulate(LENS)), dtype=torch.int32, device="cuda")
    pos = torch.cat([torch.arange(n, dtype=torch.int32, device="cuda") for n in LENS])
    with torch.no_grad():
        for _ in range(int(os.environ.get("REPS", "2"))):
            eng.forward_packed(x, pos, cu, max(LENS), lens=LENS)
    torch.cuda.synchronize()
    print("ok", layers, "layers,", M, "rows")


if __name__ == "__main__":
    main()
