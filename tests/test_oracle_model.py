"""Pin the fp32 torch restatement of CLIP + fusion + Q-Former against tensors produced by the
reference's own modules (golden_model.npz, made by oracle/gen_golden.py with the same seeded weights)."""
import numpy as np
import torch

from helpers import PINPOINTS_C3, hf_processor, synth_image
from oracle import model as M
from oracle import pil_ops as P


def test_oracle_model_matches_reference_config1(golden_dir, seeded_weights):
    torch.set_num_threads(max(1, torch.get_num_threads()))
    g = np.load(f"{golden_dir}/golden_model.npz")
    lut = np.load(f"{golden_dir}/golden_pixels.npz")["lut"]
    img = synth_image(0, 336, 336)
    px = torch.from_numpy(P.normalize_lut(img[None], lut))
    ids = torch.from_numpy(g["c1_ids"])
    embed = seeded_weights["embed"]
    with torch.no_grad():
        hs = M.clip_hidden_states(seeded_weights["clip"], px)
        feats = M.fuse_features(hs)
        assert np.allclose(feats[0, ::48, ::40].numpy(), g["c1_tower_probe"], atol=2e-4, rtol=1e-4)
        text = M.text_embeddings_for(ids, [1], embed)
        assert text.shape == (1, 63, 4096)
        vis = M.qformer_forward(seeded_weights["qf"], feats, text)
        assert np.allclose(vis[0, :, ::64].numpy(), g["c1_vis_probe32"], atol=1e-3, rtol=1e-3)
        assert np.abs(vis[0].numpy() - g["c1_vis"].astype(np.float32)).max() < 5e-3  # fp16 storage
        nt = M.qformer_forward(seeded_weights["qf"], feats, None)
        assert np.allclose(nt[0, :, ::64].numpy(), g["c1_vis_notext_probe32"], atol=1e-3, rtol=1e-3)


def test_oracle_model_matches_reference_long_text(golden_dir, seeded_weights):
    """BASELINE config 5's text length (S = 2048, L = 2047 conditioning rows per tile, 1 + 3 tiles): the literal
    fp32 restatement == the reference's own modules (golden_model_long.npz).  Tile 0 only (the oracle pushes all
    32 + 2047 rows of a tile through block 0; the GPU test covers all four tiles against the same golden)."""
    g = np.load(f"{golden_dir}/golden_model_long.npz")
    lut = np.load(f"{golden_dir}/golden_pixels.npz")["lut"]
    px0 = P.normalize_lut(synth_image(7, 336, 336)[None], lut)
    ids = torch.from_numpy(g["ids"])
    embed = seeded_weights["embed"]
    with torch.no_grad():
        feats = M.fuse_features(M.clip_hidden_states(seeded_weights["clip"], torch.from_numpy(px0)))
        text = M.text_embeddings_for(ids, [1, 3], embed)
        assert text.shape == (4, 2047, 4096)
        vis = M.qformer_forward(seeded_weights["qf"], feats, text[:1])
    assert np.allclose(vis[0, :, ::64].numpy(), g["vis_probe32"][0], atol=1e-3, rtol=1e-3)
    assert np.abs(vis[0].numpy() - g["vis"][0].astype(np.float32)).max() < 5e-3  # fp16 storage
