"""Pin the merge + splice oracle against outputs of the reference's own mixin (golden_*.npz)."""
import numpy as np
import pytest

from helpers import PINPOINTS_C3
from oracle import splice as S

Q = 32


def _features_for(tiles):
    feats, base = [], 100000
    n = sum(tiles)
    allf = (base + np.arange(n * Q, dtype=np.float32)).reshape(n, Q, 1).repeat(4, axis=2)
    o = 0
    for t in tiles:
        feats.append(allf[o:o + t])
        o += t
    return feats


def load_splice_case(g, c):
    k = f"case{c}_"
    ids = g[k + "ids"]
    mask = g[k + "mask"] if g[k + "mask"].size else None
    labels = g[k + "labels"] if g[k + "labels"].size else None
    merge, left, max_len, has_pos = (int(v) for v in g[k + "cfg"])
    return dict(ids=ids, mask=mask, labels=labels, tiles=g[k + "tiles"].tolist(),
                merge=["flat", "spatial", "spatial_unpad"][merge], side="left" if left else "right",
                max_len=None if max_len < 0 else max_len, has_pos=bool(has_pos),
                emb_code=g[k + "emb_code"], out_labels=g[k + "out_labels"], out_mask=g[k + "out_mask"],
                out_mask_dtype=str(g[k + "out_mask_dtype"]), out_pos=g[k + "out_pos"])


def n_splice_cases(g):
    return len([k for k in g.files if k.endswith("_ids")])


def test_splice_oracle_matches_reference(golden_dir):
    g = np.load(f"{golden_dir}/golden_splice.npz")
    assert n_splice_cases(g) == 8
    embed = np.arange(500, dtype=np.float32)[:, None].repeat(4, axis=1)
    newline = np.full((4,), -7.0, np.float32)
    for c in range(n_splice_cases(g)):
        case = load_splice_case(g, c)
        feats = S.process_image_patches(_features_for(case["tiles"]), [(336, 336)] * len(case["tiles"]), case["merge"],
                                        [(1, 2)] * len(case["tiles"]), newline, side=24)
        e, l, m, p, lens = S.splice(case["ids"], case["mask"], case["labels"], case["has_pos"], embed, feats,
                                    case["max_len"], case["side"])
        assert np.array_equal(e[:, :, 0].astype(np.int64), case["emb_code"]), c
        if case["labels"] is not None:
            assert np.array_equal(l, case["out_labels"]), c
        if case["mask"] is not None:
            assert np.array_equal(m.astype(np.int64), case["out_mask"].astype(np.int64)), c
        if case["has_pos"]:
            assert np.array_equal(p, case["out_pos"]), c


def test_known_answer_lengths(golden_dir):
    """B=2, S=20, tiles (3,4), sample 1 masked from column 15 -> lengths (115, 142)."""
    g = np.load(f"{golden_dir}/golden_splice.npz")
    case = load_splice_case(g, 0)
    assert case["emb_code"].shape == (2, 142)
    assert (case["out_mask"] != 0).sum(1).tolist() == [115, 142]
    # supervised (non-ignored) positions the reference kept for this seed: visual rows and masked text are -100
    assert (case["out_labels"] != -100).sum(1).tolist() == [14, 8]
    assert case["out_mask_dtype"] == "torch.int64"


def test_merge_oracle_matches_reference(golden_dir):
    g = np.load(f"{golden_dir}/golden_merge.npz")
    n = len([k for k in g.files if k.endswith("_meta")])
    assert n == 10
    newline = np.full((2,), -7.0, np.float32)
    for c in range(n):
        W, H, n_w, n_h, T, unpad = (int(v) for v in g[f"case{c}_meta"])
        merge = "spatial_unpad" if unpad else "spatial"
        feat = np.arange(T * 576, dtype=np.float32).reshape(T, 576, 1).repeat(2, axis=2)
        single = np.arange(576, dtype=np.float32).reshape(1, 576, 1).repeat(2, axis=2)
        out = S.process_image_patches([feat, single], [(W, H), (336, 336)], merge, [(n_w, n_h), (1, 1)], newline, side=24)
        assert np.array_equal(out[0][:, 0].astype(np.int64), g[f"case{c}_rows"]), c
        assert np.array_equal(out[1][:, 0].astype(np.int64), g[f"case{c}_single_rows"]), c
    # row counts recorded in SURVEY.md 8(a) row A9
    assert g["case0_rows"].shape[0] == 2732 and g["case1_rows"].shape[0] == 870 and g["case2_rows"].shape[0] == 648
