"""Host-side integer logic of the product package (no GPU): resolution choice, grid shapes, LANCZOS
tables, unpad bounds / merged row counts, image sharding and the gloo exchange."""
import os
import sys

import numpy as np
import pytest
import torch

import vision_zephyr_b200 as vz
from vision_zephyr_b200 import anyres
from vision_zephyr_b200.dist import gather_visual_tokens, shard_images
from helpers import PINPOINTS_C3, PINPOINTS_SHIPPED
from oracle import pil_ops as P


def test_shipped_pinpoints_known_answers():
    """SURVEY.md 8(c) known-answer vectors."""
    pins = "'[[336, 672], [672, 336], [336, 1008], [1008, 336]]'"  # doubly quoted, as in config.json:16
    assert anyres.select_best_fit_resolution((637, 336), PINPOINTS_SHIPPED) == (672, 336)
    assert anyres.select_best_fit_resolution((681, 336), PINPOINTS_SHIPPED) == (1008, 336)
    assert anyres.select_best_fit_resolution((1920, 804), PINPOINTS_SHIPPED) == (1008, 336)
    assert anyres.select_best_fit_resolution((336, 336), PINPOINTS_SHIPPED) == (336, 672)
    assert anyres.calculate_grid_shape((637, 336), pins, 336) == (2, 1)
    assert anyres.calculate_grid_shape((1000, 900), str(PINPOINTS_C3), 336) == (2, 2)
    assert len(anyres.anyres_views((1000, 900), PINPOINTS_C3)[0]) == 5
    with pytest.raises(ValueError):
        anyres.calculate_grid_shape((10, 10), "7", 336)


@pytest.mark.parametrize("n_in,n_out", [(1000, 672), (900, 604), (336, 336), (200, 336), (1920, 336), (17, 336), (804, 140)])
def test_lanczos_table_equals_oracle_matrix(n_in, n_out):
    t = anyres.lanczos_table(n_in, n_out)
    ks, n = int(t[0]), int(t[1])
    assert n == n_out
    xmin, cnt, kk = t[2:2 + n], t[2 + n:2 + 2 * n], t[2 + 2 * n:].reshape(n, ks)
    K = np.zeros((n_out, n_in), np.int64)
    for x in range(n):
        K[x, xmin[x]:xmin[x] + cnt[x]] = kk[x, :cnt[x]]
        assert not kk[x, cnt[x]:].any()
    ref = P.coeff_matrix(n_in, n_out) if n_in != n_out else np.eye(n_in, dtype=np.int64) * (1 << 22)
    assert np.array_equal(K, ref)


@pytest.mark.parametrize("n_in,n_out", [(700, 470), (500, 336), (200, 448), (1300, 336), (336, 336)])
def test_bicubic_table_equals_oracle_matrix(n_in, n_out):
    t = anyres.resample_table(n_in, n_out, "bicubic")
    ks, n = int(t[0]), int(t[1])
    xmin, cnt, kk = t[2:2 + n], t[2 + n:2 + 2 * n], t[2 + 2 * n:].reshape(n, ks)
    K = np.zeros((n_out, n_in), np.int64)
    for x in range(n):
        K[x, xmin[x]:xmin[x] + cnt[x]] = kk[x, :cnt[x]]
    ref = P.coeff_matrix(n_in, n_out, "bicubic") if n_in != n_out else np.eye(n_in, dtype=np.int64) * (1 << 22)
    assert np.array_equal(K, ref)


def test_fixed_view_geometry():
    """canvas / view geometry of the process_images modes (mm_utils.py:16-87, train.py:570-590)"""
    v, c = anyres.fixed_view((700, 500), "pad")
    assert (c["W"], c["H"], c["pad_x"], c["pad_y"], c["bg"]) == (700, 700, 0, 100, (122, 116, 104))
    assert (v[0]["out_w"], v[0]["out_h"], v[0]["tile_x"], v[0]["tile_y"], v[0]["filt"]) == (336, 336, 0, 0, "bicubic")
    v, c = anyres.fixed_view((420, 901), "square")
    assert (c["W"], c["H"], c["pad_x"], c["pad_y"]) == (420, 420, 0, -240)
    v, c = anyres.fixed_view((700, 500), "plain")
    assert (v[0]["out_w"], v[0]["out_h"], v[0]["tile_x"], v[0]["tile_y"]) == (470, 336, 67, 0)
    assert anyres.single_view((336, 336)) == [dict(out_w=336, out_h=336, off_x=0, off_y=0, tile_x=0, tile_y=0)]
    with pytest.raises(ValueError):
        anyres.fixed_view((10, 10), "identity")


def test_max_span_covers_every_256_column_block():
    """the horizontal-pass kernel sizes its row buffers from preprocess._max_span: it must bound the source window
    of any 256 consecutive output columns, including the taps of every column inside the block"""
    from vision_zephyr_b200.preprocess import _max_span
    for n_in, n_out, filt in [(1000, 672, "lanczos"), (1344, 336, "lanczos"), (700, 470, "bicubic"), (200, 448, "bicubic"),
                              (336, 336, "lanczos"), (301, 1008, "lanczos")]:
        t = anyres.resample_table(n_in, n_out, filt)
        n = int(t[1])
        xmin, cnt = t[2:2 + n], t[2 + n:2 + 2 * n]
        span = _max_span(n_in, n_out, filt)
        for x0 in range(0, n, 256):
            xs = np.arange(x0, min(x0 + 256, n))
            assert int((xmin[xs] + cnt[xs]).max() - xmin[x0]) <= span
            assert (xmin[xs] >= xmin[x0]).all()


def test_merged_row_counts_match_reference(golden_dir):
    g = np.load(f"{golden_dir}/golden_merge.npz")
    for c in range(10):
        W, H, n_w, n_h, T, unpad = (int(v) for v in g[f"case{c}_meta"])
        d = anyres.slot_descriptor(0, T, 576, "spatial_unpad" if unpad else "spatial", "anyres", (W, H),
                                   str(PINPOINTS_C3), 336, 24)
        assert d["n_rows"] == g[f"case{c}_rows"].shape[0], c
        assert (d["n_w"], d["n_h"]) == (n_w, n_h)
        s = anyres.slot_descriptor(0, 1, 576, "spatial_unpad" if unpad else "spatial")
        assert s["n_rows"] == g[f"case{c}_single_rows"].shape[0]
    assert anyres.slot_descriptor(64, 5, 32, "flat")["n_rows"] == 160
    with pytest.raises(ValueError):
        anyres.slot_descriptor(0, 1, 32, "bogus")
    with pytest.raises(NotImplementedError):
        anyres.slot_descriptor(0, 5, 576, "spatial", "square", (10, 10), str(PINPOINTS_C3), 336, 24)


def test_builders_keep_reference_errors():
    from types import SimpleNamespace
    with pytest.raises(ValueError, match="Unknown vision tower path"):
        vz.build_vision_tower(SimpleNamespace(mm_vision_tower="/no/such/dir", mm_vision_select_layer="-2"))
    t = vz.build_vision_tower(SimpleNamespace(mm_vision_tower="openai/clip-vit-large-patch14-336",
                                              mm_vision_select_layer="-2,-5,-8,-11,6"), delay_load=True)
    assert (t.hidden_size, t.num_patches, t.is_loaded, t.select_layers) == (5120, 576, False, [-2, -5, -8, -11, 6])
    with pytest.raises(ValueError, match="Invalid format"):
        vz.build_vision_tower(SimpleNamespace(mm_vision_tower="openai/x", mm_vision_select_layer="a,b"), delay_load=True)
    assert vz.build_vision_projector is vz.build_multimodal_projector


def test_projector_state_dict_keys_match_reference():
    from types import SimpleNamespace
    with torch.device("meta"):
        p = vz.build_multimodal_projector(SimpleNamespace(hidden_size=4096))
    keys = set(p.state_dict().keys())
    assert len(keys) == 1 + 8 * 20 + 4
    for k in ["learned_queries", "blocks.0.self_attn.in_proj_weight", "blocks.7.cross_attn.k_proj_weight",
              "blocks.3.cross_attn.in_proj_bias", "blocks.3.cross_attn.out_proj.bias", "blocks.5.ffn.0.weight",
              "blocks.5.ffn.2.bias", "blocks.1.norm3.weight", "pre_norm.bias", "norm.weight"]:
        assert k in keys, k
    assert sum(v.numel() for v in p.state_dict().values()) == 1678428160
    with pytest.raises(ValueError):
        vz.build_multimodal_projector(SimpleNamespace(hidden_size=1024))


def test_shard_images_is_a_contiguous_partition():
    for tiles, world in [([5] * 8, 2), ([5] * 64, 8), ([3, 4, 5, 1, 2], 3), ([5] * 3, 8), ([1], 4)]:
        b = shard_images(tiles, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == len(tiles)
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
    assert shard_images([5] * 64, 8) == [(8 * r, 8 * r + 8) for r in range(8)]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tiles = [5, 3, 4, 5]
    bounds = shard_images(tiles, world)
    rows = [sum(tiles[a:b]) * 32 for a, b in bounds]
    start = sum(rows[:rank])
    local = (torch.arange(rows[rank] * 8, dtype=torch.float32).reshape(rows[rank], 8) + start * 8).to(torch.bfloat16)
    full = gather_visual_tokens(local, rows)
    ref = torch.arange(sum(rows) * 8, dtype=torch.float32).reshape(-1, 8).to(torch.bfloat16)
    q.put((rank, bool(torch.equal(full, ref)), tuple(full.shape)))
    dist.destroy_process_group()


def test_gather_visual_tokens_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert res == [(0, True, (17 * 32, 8)), (1, True, (17 * 32, 8))]


def test_shard_images_leaves_no_rank_empty_and_minimises_the_largest_shard():
    """the greedy cut handed out empty shards ([7,1,1,1] on 4 ranks); the partition must be the
    min-max contiguous one and every rank must get an image while n >= world."""
    import itertools
    import random
    assert shard_images([7, 1, 1, 1], 4) == [(0, 1), (1, 2), (2, 3), (3, 4)]
    rng = random.Random(0)
    for _ in range(1500):
        n, w = rng.randint(0, 9), rng.randint(1, 5)
        t = [rng.randint(1, 9) for _ in range(n)]
        b = shard_images(t, w)
        assert len(b) == w and b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        if n >= w:
            assert all(e > s for s, e in b), (t, w, b)
            best = min(max(sum(t[a:c]) for a, c in zip((0,) + cuts, cuts + (n,)))
                       for cuts in itertools.combinations(range(1, n), w - 1))
            assert max(sum(t[s:e]) for s, e in b) == best, (t, w, b, best)


def test_initialize_vision_modules_mirrors_the_reference(tmp_path):
    """vis_zephyr_arch.py:49-102 on a bare host (fresh-build branch, FSDP list form, 'unpad' newline, adapter
    load with the 'model.mm_projector.' prefix the trainer writes).  The 1.7 B-parameter projector is built on
    the meta device here; the value round trip runs on the GPU (tests/test_gpu_llm.py)."""
    from types import SimpleNamespace
    from transformers import CLIPVisionConfig, CLIPVisionModel
    from vision_zephyr_b200 import arch
    d = str(tmp_path / "clip")
    cfg = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                           image_size=336, patch_size=14, projection_dim=768, hidden_act="quick_gelu")
    with torch.device("meta"):
        hf = CLIPVisionModel(cfg)
    hf = hf.to_empty(device="cpu").to(torch.bfloat16)
    for p in hf.parameters():
        p.data.normal_(0, 0.02)
    hf.save_pretrained(d)
    del hf

    class Base(torch.nn.Module):
        def __init__(self, config):
            super().__init__()
            self.config = config
            self.dtype = torch.bfloat16

    class Host(arch.VisZephyrB200MetaModel, Base):
        pass

    built = {}

    def meta_projector(config, **kw):
        with torch.device("meta"):
            built["p"] = vz.build_multimodal_projector(config)
        return built["p"]

    host = Host(SimpleNamespace(hidden_size=4096))           # no mm_vision_tower in the config: nothing built yet
    assert host.get_vision_tower() is None and not hasattr(host, "mm_projector")
    args = SimpleNamespace(mm_vision_tower=d, mm_vision_select_layer="-2,-5,-8,-11,6", mm_vision_select_feature="patch",
                           pretrain_mm_mlp_adapter=None, mm_patch_merge_type="spatial_unpad",
                           mm_grid_pinpoints="[[336, 672]]", image_aspect_ratio="anyres")
    orig = arch.build_multimodal_projector
    arch.build_multimodal_projector = meta_projector
    try:
        host.initialize_vision_modules(args, fsdp=["full_shard"])
    finally:
        arch.build_multimodal_projector = orig
    assert type(host.vision_tower) is list and host.get_vision_tower().is_loaded
    c = host.config
    assert (c.mm_vision_tower, c.use_mm_proj, c.mm_projector_type, c.mm_hidden_size) == (d, True, "linear", 5120)
    assert (c.mm_vision_select_layer, c.mm_vision_select_feature, c.mm_patch_merge_type) == \
        ("-2,-5,-8,-11,6", "patch", "spatial_unpad")
    assert (c.mm_grid_pinpoints, c.image_aspect_ratio, c.mm_use_im_start_end) == ("[[336, 672]]", "anyres", False)
    assert host.mm_projector is built["p"]
    assert host.image_newline.shape == (4096,) and host.image_newline.dtype == torch.bfloat16
    assert 0.5 / 64 < host.image_newline.float().std().item() < 2.0 / 64          # N(0, 1/sqrt(4096))
    # second call: existing modules are kept, the tower reloads, a frozen projector is un-frozen
    host.mm_projector.requires_grad_(False)
    host.initialize_vision_modules(args, fsdp=["full_shard"])
    assert host.mm_projector is built["p"] and all(p.requires_grad for p in host.mm_projector.parameters())
    # key format of mm_projector.bin (train/vis_zephyr_trainer.py:326-343) -> get_w -> projector keys
    from vision_zephyr_b200.language_model import mm_adapter_state
    wrapper = torch.nn.Module()
    wrapper.model = host
    names = [k for k, _ in wrapper.named_parameters() if "mm_projector" in k]
    assert names[0].startswith("model.mm_projector.")
    stripped = {k.split("mm_projector.")[1] for k in names}
    assert stripped == set(host.mm_projector.state_dict().keys())


def test_gate_up_interleave_and_cpu_forward_stays_hf():
    """mistral_prefill: the SwiGLU weight layout (blocks of 64 gate rows + the 64 matching up rows), and the
    native-prefill switch of VisZephyrB200Model leaves CPU calls with HF Mistral (the engine has no CPU path)."""
    from vision_zephyr_b200.mistral_prefill import MistralPrefillB200, interleave_gate_up
    I, K = 192, 8
    gate = torch.arange(I * K, dtype=torch.float32).reshape(I, K)
    up = -gate
    w = interleave_gate_up(gate, up)
    assert w.shape == (2 * I, K)
    for j in range(I // 64):
        assert torch.equal(w[128 * j: 128 * j + 64], gate[64 * j: 64 * j + 64])
        assert torch.equal(w[128 * j + 64: 128 * j + 128], up[64 * j: 64 * j + 64])
    with pytest.raises(ValueError):
        interleave_gate_up(gate[:100], up[:100])
    from transformers import MistralConfig, MistralModel
    cfg = MistralConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2,
                        num_key_value_heads=1, vocab_size=50, sliding_window=None)
    m = MistralModel(cfg).eval()
    with pytest.raises(vz._lib.VzError):
        MistralPrefillB200(m)             # CPU weights: loud failure, no fallback inside the engine


def test_attn_causal_work_list_is_built_on_the_host():
    """vz_attn_causal_items is pure host code: (128-query tile, head) items of every sample, longest first."""
    import ctypes as C
    from vision_zephyr_b200 import _lib
    lib = _lib.load()
    lens = np.asarray([300, 0, 37, 129, 1], dtype=np.int32)
    nh = 4
    tiles = [(s + 127) // 128 for s in lens]
    n = lib.vz_attn_causal_items(lens.ctypes.data, len(lens), nh, None, 0, None)
    assert n == sum(tiles) * nh
    items = np.empty((n, 4), dtype=np.int32)
    flops = C.c_double(0)
    assert lib.vz_attn_causal_items(lens.ctypes.data, len(lens), nh, items.ctypes.data, n, C.byref(flops)) == n
    assert flops.value == 4.0 * 128 * sum(int(s) * (int(s) + 1) / 2 for s in lens) * nh
    assert lib.vz_attn_causal_items(lens.ctypes.data, len(lens), nh, items.ctypes.data, n - 1, None) == n   # too small: count only
    starts = np.concatenate([[0], np.cumsum(lens)])
    seen = set()
    for q_row0, kv_row0, n_keys, hq in items:
        h, qt = hq & 0xff, hq >> 8
        b = int(np.searchsorted(starts, kv_row0, side="right") - 1)
        while lens[b] == 0:
            b += 1                                        # an empty sample shares its start row with the next one
        assert starts[b] == kv_row0 and q_row0 == kv_row0 + qt * 128 and 0 <= h < nh
        assert n_keys == min((qt + 1) * 128, lens[b])
        seen.add((b, qt, h))
    assert len(seen) == n                                 # every (sample, tile, head) exactly once
    nkb = (items[:, 2] + 63) // 64
    assert (np.diff(nkb) <= 0).all()                      # longest first
    assert lib.vz_attn_causal_items(lens.ctypes.data, len(lens), 300, None, 0, None) < 0      # heads must fit 8 bits
