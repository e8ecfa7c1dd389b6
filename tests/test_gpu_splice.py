"""GPU parity of plan + text gather + merge + scatter: bit-exact against the numpy oracle and the
reference's goldens (indices, labels, masks, position ids, lengths and the copied rows)."""
import numpy as np
import pytest
import torch

from helpers import PINPOINTS_C3
from test_oracle_splice import _features_for, load_splice_case, n_splice_cases

pytestmark = pytest.mark.gpu
Q = 32


def gpu_splice(ids, mask, labels, embed, feats_rows, descs, max_len, side, dtype=torch.float32, newline=None):
    import vision_zephyr_b200 as vz
    from vision_zephyr_b200 import arch
    dev = "cuda"
    ids_d = torch.from_numpy(ids).to(dev)
    mask_d = torch.from_numpy(mask.astype(np.uint8)).to(dev) if mask is not None else None
    labels_d = torch.from_numpy(labels).to(dev) if labels is not None else None
    emb_d = torch.from_numpy(embed).to(dev).to(dtype).contiguous()
    vis_d = torch.from_numpy(feats_rows).to(dev).to(dtype).contiguous()
    nl_d = torch.from_numpy(newline).to(dev).to(dtype).contiguous() if newline is not None else None
    slots_dev, prefix, total = arch._slots_to_device(descs, dev)
    plan = arch.splice_plan(ids_d, mask_d, slots_dev, len(descs), max_len or 0)
    info = plan.wait()
    out = arch.splice_scatter(ids_d, labels_d, emb_d, vis_d, nl_d, slots_dev, prefix, len(descs), total, plan,
                              info["Lmax"], side == "left")
    torch.cuda.synchronize()
    return [out[0].float().cpu().numpy()] + [o.cpu().numpy() for o in out[1:]], info, plan


def test_splice_matches_reference_goldens(golden_dir):
    from vision_zephyr_b200 import anyres
    g = np.load(f"{golden_dir}/golden_splice.npz")
    embed = np.arange(500, dtype=np.float32)[:, None].repeat(4, axis=1)
    newline = np.full((4,), -7.0, np.float32)
    for c in range(n_splice_cases(g)):
        case = load_splice_case(g, c)
        tiles = case["tiles"]
        feats = np.concatenate(_features_for(tiles)).reshape(-1, 4)
        descs, base = [], 0
        for t in tiles:
            descs.append(anyres.slot_descriptor(base, t, Q, case["merge"]))
            base += t * Q
        (e, l, m, p), info, plan = gpu_splice(case["ids"], case["mask"], case["labels"], embed, feats, descs,
                                              case["max_len"], case["side"], newline=newline)
        assert np.array_equal(e[:, :, 0].astype(np.int64), case["emb_code"]), c
        assert (e == e[:, :, :1]).all()
        if case["labels"] is not None:
            assert np.array_equal(l, case["out_labels"]), c
        else:
            assert (l == -100).all()
        if case["mask"] is not None:
            assert np.array_equal(m.astype(np.int64), case["out_mask"].astype(np.int64)), c
        if case["has_pos"]:
            assert np.array_equal(p, case["out_pos"]), c
        assert info["Lmax"] == case["emb_code"].shape[1]


@pytest.mark.parametrize("side", ["right", "left"])
@pytest.mark.parametrize("max_len", [None, 1500])
def test_splice_full_size_against_oracle(side, max_len):
    """config 5 geometry: B=8, S=2048, hidden 4096 bf16, one image token per sample, 160 visual rows."""
    from oracle import splice as S
    from vision_zephyr_b200 import anyres
    rng = np.random.default_rng(5)
    B, Smax, D, V = 8, 2048, 4096, 32000
    ids = np.full((B, Smax), 2, np.int64)
    mask = np.zeros((B, Smax), bool)
    for b in range(B):
        n = int(rng.integers(256, 2048))
        ids[b, :n] = rng.integers(3, V, n)
        ids[b, int(rng.integers(1, 33))] = -200
        mask[b, :n] = True
    labels = ids.copy()
    labels[:, :Smax // 3] = -100
    embed = (rng.standard_normal((V, D)) * 0.02).astype(np.float32)
    feats = [rng.standard_normal((5 * Q, D)).astype(np.float32) for _ in range(B)]
    bf = lambda a: torch.from_numpy(a).to(torch.bfloat16).float().numpy()
    embed, feats = bf(embed), [bf(f) for f in feats]
    ref = S.splice(ids, mask, labels, True, embed, feats, max_len, side)
    descs = [anyres.slot_descriptor(b * 5 * Q, 5, Q, "flat") for b in range(B)]
    (e, l, m, p), info, plan = gpu_splice(ids, mask, labels, embed, np.concatenate(feats), descs, max_len, side,
                                          dtype=torch.bfloat16)
    assert np.array_equal(torch.from_numpy(e).float().numpy() if e.dtype != np.float32 else e, ref[0])
    assert np.array_equal(l, ref[1]) and np.array_equal(m.astype(bool), ref[2]) and np.array_equal(p, ref[3])
    assert np.array_equal(plan.lengths.cpu().numpy(), ref[4])
    # size-independent properties: every row written once, mask rows == lengths, pads are zero
    assert (m.sum(1) == ref[4]).all()
    assert (np.abs(e.astype(np.float32))[~m.astype(bool)] == 0).all()


def test_merge_matches_reference_goldens(golden_dir):
    from vision_zephyr_b200 import anyres, arch
    g = np.load(f"{golden_dir}/golden_merge.npz")
    for c in range(10):
        W, H, n_w, n_h, T, unpad = (int(v) for v in g[f"case{c}_meta"])
        merge = "spatial_unpad" if unpad else "spatial"
        feat = torch.arange(T * 576, dtype=torch.float32).reshape(T * 576, 1).repeat(1, 8).cuda()
        newline = torch.full((8,), -7.0).cuda()
        d = anyres.slot_descriptor(0, T, 576, merge, "anyres", (W, H), str(PINPOINTS_C3), 336, 24)
        s = anyres.slot_descriptor(0, 1, 576, merge)
        out = arch.merge_rows(feat, newline, [d])[0].cpu().numpy()
        assert np.array_equal(out[:, 0].astype(np.int64), g[f"case{c}_rows"]), c
        assert (out == out[:, :1]).all()
        outs = arch.merge_rows(feat[:576], newline, [s])[0].cpu().numpy()
        assert np.array_equal(outs[:, 0].astype(np.int64), g[f"case{c}_single_rows"]), c


def test_text_gather_matches_oracle():
    from vision_zephyr_b200 import arch
    rng = np.random.default_rng(9)
    B, S, D, V = 5, 77, 64, 300
    ids = rng.integers(3, V, (B, S)).astype(np.int64)
    ids[0, 0] = -200
    ids[1, 40] = -200
    ids[1, 41] = -200
    ids[3, S - 1] = -200
    embed = rng.standard_normal((V, D)).astype(np.float32)
    ids_d = torch.from_numpy(ids).cuda()
    emb_d = torch.from_numpy(embed).cuda().to(torch.bfloat16)
    slots_dev, prefix, total = arch._slots_to_device([dict(row_base=0, n_rows=32, merge=0, hw=32, h=0, w=0, n_w=0,
                                                           n_h=0, y0=0, y1=0, x0=0, x1=0)] * 8, "cuda")
    plan = arch.splice_plan(ids_d, None, slots_dev, 8, 0)
    info = plan.wait()
    text_emb, text_off = arch.text_gather(ids_d, emb_d, plan, info["text_rows"])
    torch.cuda.synchronize()
    lens = [(ids[b] != -200).sum() for b in range(B)]
    assert info["text_rows"] == sum(lens) and info["L_text"] == max(lens)
    assert text_off.cpu().tolist() == np.concatenate([[0], np.cumsum(lens)]).tolist()
    ref = torch.cat([emb_d.cpu()[torch.from_numpy(ids[b][ids[b] != -200])] for b in range(B)])
    assert torch.equal(text_emb[:-1].cpu(), ref)
    assert (text_emb[-1] == 0).all()
