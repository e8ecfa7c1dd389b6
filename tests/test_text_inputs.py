"""'Next' row 2 of the scope table: prompt tokenisation with image placeholders (host logic, CPU test)
and the supervised collator (GPU kernel), both against goldens from the reference's own functions."""
import types

import numpy as np
import pytest
import torch

from oracle import splice as S

PROMPTS = ["<image>\nwhat is shown here ?", "describe <image> and also <image> please", "no picture at all",
           "<image>", "tail image <image>", "", "a <image><image> b"]


class StubTokenizer:
    def __init__(self, with_bos=True, pad_token_id=2, model_max_length=40):
        self.bos_token_id, self.pad_token_id, self.model_max_length, self.with_bos = 1, pad_token_id, model_max_length, with_bos

    def __call__(self, text):
        ids = ([self.bos_token_id] if self.with_bos else []) + [3 + (sum(map(ord, w)) % 997) for w in text.split()]
        return types.SimpleNamespace(input_ids=ids)


def test_tokenizer_image_token_matches_reference(golden_dir):
    import vision_zephyr_b200 as vz
    g = np.load(f"{golden_dir}/golden_text.npz")
    for bi, with_bos in enumerate([True, False]):
        tok = StubTokenizer(with_bos)
        for pi, prompt in enumerate(PROMPTS):
            ref = g[f"tok{bi}_{pi}"].tolist()
            assert S.tokenizer_image_token(prompt, tok) == ref, (bi, pi)       # oracle pinned
            assert vz.tokenizer_image_token(prompt, tok) == ref, (bi, pi)      # product (host logic)
    t = vz.tokenizer_image_token(PROMPTS[0], StubTokenizer(), return_tensors="pt")
    assert t.dtype == torch.long and t.tolist() == g["tok0_0"].tolist()
    with pytest.raises(ValueError):
        vz.tokenizer_image_token("x", StubTokenizer(), return_tensors="np")


def _collate_cases(g):
    for ci in range(4):
        B, pad, mml = (int(v) for v in g[f"col{ci}_cfg"])
        inst = [dict(input_ids=g[f"col{ci}_ids{b}"], labels=g[f"col{ci}_labels{b}"]) for b in range(B)]
        yield ci, inst, pad, mml


def test_collate_oracle_matches_reference(golden_dir):
    g = np.load(f"{golden_dir}/golden_text.npz")
    for ci, inst, pad, mml in _collate_cases(g):
        ids, labels, mask = S.collate(inst, pad, mml)
        assert np.array_equal(ids, g[f"col{ci}_out_ids"]) and np.array_equal(labels, g[f"col{ci}_out_labels"])
        assert np.array_equal(mask, g[f"col{ci}_out_mask"])


@pytest.mark.gpu
def test_collate_kernel_matches_reference(golden_dir):
    import vision_zephyr_b200 as vz
    g = np.load(f"{golden_dir}/golden_text.npz")
    for ci, inst, pad, mml in _collate_cases(g):
        batch = vz.collate_supervised(inst, pad, mml)
        torch.cuda.synchronize()
        assert np.array_equal(batch["input_ids"].cpu().numpy(), g[f"col{ci}_out_ids"]), ci
        assert np.array_equal(batch["labels"].cpu().numpy(), g[f"col{ci}_out_labels"]), ci
        assert batch["attention_mask"].dtype == torch.bool
        assert np.array_equal(batch["attention_mask"].cpu().numpy(), g[f"col{ci}_out_mask"]), ci
