"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference)
on seeded inputs.  Run in the build container only (the reference does not exist on the GPU box):

    python oracle/gen_golden.py [--skip-model]

What is pinned
  golden_pixels.npz   process_any_resolution_image + CLIPImageProcessor on synthetic images:
                      tile counts, best-fit resolutions, SHA-256 of the f32 pixel tensors
  golden_vip.npz      image_blending (rectangle / mask / arrow) on 336x336 images: SHA-256 of the RGB result
  golden_vip_shapes.npz  image_blending for ellipse / triangle / scribble / mask contour and random widths / alphas
  golden_merge.npz    _process_image_patches row maps for flat / spatial / spatial_unpad
  golden_splice.npz   prepare_inputs_labels_for_multimodal index/label/mask/position outputs with a
                      stubbed encode_images (integer-coded features)
  golden_text.npz     tokenizer_image_token (stub word-hash tokenizer) and DataCollatorForSupervisedDataset
  golden_model.npz    CLIPVisionTower + QFormer outputs (fp32, seeded weights from oracle/weights.py)
                      for config 1 (one 336x336 tile, 63 text tokens) and a 5-tile anyres image
  golden_model_long.npz  the same at BASELINE config 5's text length: S = 2048, B = 2 (1 + 3 tiles), L = 2047
  golden_modes.npz    mm_utils.process_images in its four aspect modes through the Pillow-backed CLIP processor
"""
import argparse
import hashlib
import os
import random
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from oracle import weights  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
PINPOINTS_SHIPPED = [[336, 672], [672, 336], [336, 1008], [1008, 336]]
PINPOINTS_C3 = [[336, 672], [672, 336], [672, 672], [336, 1008], [1008, 336]]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def synth_image(i, W, H):
    return np.random.default_rng(1000 + i).integers(0, 256, (H, W, 3), dtype=np.uint8)


def make_processor():
    from transformers import CLIPImageProcessor
    return CLIPImageProcessor(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336},
                              image_mean=[0.48145466, 0.4578275, 0.40821073],
                              image_std=[0.26862954, 0.26130258, 0.27577711], resample=3)


def stub_shapely():
    """shapely is absent; rectangle/mask/arrow never use the shapely objects."""
    sh = types.ModuleType("shapely")
    ops = types.ModuleType("shapely.ops")
    geo = types.ModuleType("shapely.geometry")
    ops.unary_union = lambda polys: None
    geo.Polygon = lambda pts: None
    geo.Point = lambda *a: None
    sh.ops, sh.geometry = ops, geo
    sys.modules.update({"shapely": sh, "shapely.ops": ops, "shapely.geometry": geo})


# --------------------------------------------------------------------------------------------
def gen_pixels():
    from PIL import Image
    from vis_zephyr.model.multi_scale_process import process_any_resolution_image, calculate_grid_shape
    proc = make_processor()
    cases = [(0, 1000, 900, PINPOINTS_C3), (1, 637, 336, PINPOINTS_SHIPPED), (2, 336, 900, PINPOINTS_C3),
             (3, 1920, 804, PINPOINTS_SHIPPED), (4, 336, 336, PINPOINTS_SHIPPED), (5, 700, 650, PINPOINTS_C3),
             (6, 250, 180, PINPOINTS_C3)]
    out = {}
    for (i, W, H, pins) in cases:
        img = synth_image(i, W, H)
        px = process_any_resolution_image(Image.fromarray(img), proc, pins).numpy()
        out[f"case{i}_meta"] = np.array([W, H, px.shape[0], *calculate_grid_shape((W, H), str(pins), 336)], np.int64)
        out[f"case{i}_pins"] = np.array(pins, np.int64)
        out[f"case{i}_sha"] = np.array(sha(px.astype(np.float32)))
        out[f"case{i}_probe"] = px[:, :, ::67, ::59].astype(np.float32)
    # the 768-entry LUT of this installation's processor
    ramp = np.zeros((336, 336, 3), np.uint8)
    ramp[0, :256, :] = np.arange(256, dtype=np.uint8)[:, None]
    lut = proc.preprocess(Image.fromarray(ramp), return_tensors="pt")["pixel_values"][0][:, 0, :256].numpy()
    out["lut"] = lut.astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, "golden_pixels.npz"), **out)
    print("golden_pixels.npz", {k: v.tolist() for k, v in out.items() if k.endswith("meta")})


def gen_vip():
    from PIL import Image
    stub_shapely()
    from vis_zephyr.model.vip_processor.conversation_generator import image_blending
    out = {}
    colors = [(255, 0, 0), (0, 255, 0), (0, 0, 255), (255, 255, 0)]
    for i in range(4):
        img = synth_image(100 + i, 336, 336)
        rng = np.random.default_rng(2000 + i)
        pil = Image.fromarray(img)
        specs = []
        for inst, shape in enumerate(["rectangle", "mask", "arrow"]):
            x0, y0 = rng.uniform(20, 200, 2)
            x1, y1 = x0 + rng.uniform(30, 110), y0 + rng.uniform(30, 110)
            bbox = [float(x0), float(y0), float(x1), float(y1)]
            seed = 3000 + 10 * i + inst
            random.seed(seed)
            pil = image_blending(pil, shape=shape, bbox_coor=bbox, segmentation=None, image_size_anchor=336,
                                 rgb_color=colors[inst], vip_style="constant", alpha=128)
            specs.append(bbox + [seed])
            out[f"img{i}_after{inst}_sha"] = np.array(sha(np.asarray(pil)))
        out[f"img{i}_specs"] = np.array(specs, np.float64)
        out[f"img{i}_final_probe"] = np.asarray(pil)[::31, ::29].copy()
    np.savez_compressed(os.path.join(GOLD, "golden_vip.npz"), **out)
    print("golden_vip.npz written")


VIP_SHAPE_CASES = [  # (image W, H, shape, vip_style, alpha, width, python seed, numpy seed)
    (336, 336, "rectangle", None, None, None, 11, 0), (336, 336, "ellipse", None, None, None, 12, 0),
    (336, 336, "triangle", None, 128, None, 13, 101), (336, 336, "scribble", None, None, None, 14, 102),
    (336, 336, "mask contour", None, None, None, 15, 0), (336, 336, "mask", None, None, None, 16, 0),
    (336, 336, "arrow", None, None, None, 17, 0), (500, 400, "ellipse", None, 200, 2, 18, 0),
    (500, 400, "triangle", None, None, None, 19, 103), (640, 480, "scribble", None, 99, 1, 20, 104),
    (500, 400, "arrow", None, None, 3, 21, 0), (500, 400, "mask contour", None, 255, None, 22, 0),
    (500, 400, "rectangle", "constant", None, 4, 23, 0),
]


def gen_vip_shapes():
    """image_blending for the shapes golden_vip.npz does not hold (ellipse, triangle, scribble, mask contour) and
    for RANDOM widths / alphas, segmentation=None (no shapely query is made then): SHA-256 of the RGB result
    after every instance; three instances compound on each image."""
    from PIL import Image
    stub_shapely()
    from vis_zephyr.model.vip_processor.conversation_generator import image_blending
    out = {"n_cases": np.array(len(VIP_SHAPE_CASES))}
    colors = [(255, 0, 0), (0, 255, 0), (0, 0, 255), (255, 165, 0)]
    for ci, (W, H, shape, style, alpha, width, pseed, nseed) in enumerate(VIP_SHAPE_CASES):
        img = synth_image(300 + ci, W, H)
        rng = np.random.default_rng(4000 + ci)
        x0, y0 = rng.uniform(20, W * 0.5), rng.uniform(20, H * 0.5)
        bbox = [float(x0), float(y0), float(x0 + rng.uniform(40, W * 0.4)), float(y0 + rng.uniform(40, H * 0.4))]
        random.seed(pseed)
        np.random.seed(nseed)
        pil = image_blending(Image.fromarray(img), shape=shape, bbox_coor=bbox, segmentation=None, image_size_anchor=336,
                             rgb_color=colors[ci % 4], vip_style=style, alpha=alpha, width=width)
        out[f"c{ci}_bbox"] = np.array(bbox, np.float64)
        out[f"c{ci}_sha"] = np.array(sha(np.asarray(pil)))
        # a second and third instance on top (the reference's loop, processor.py:58-73)
        pil = image_blending(pil, shape="rectangle", bbox_coor=[5.5, 7.25, W - 9.0, H - 11.5], segmentation=None,
                             image_size_anchor=336, rgb_color=colors[(ci + 1) % 4], vip_style=None, alpha=None, width=None)
        pil = image_blending(pil, shape="mask", bbox_coor=bbox, segmentation=[[x0, y0, x0 + 30, y0 + 5, x0 + 12, y0 + 44]],
                             image_size_anchor=336, rgb_color=colors[(ci + 2) % 4], vip_style=None, alpha=None, width=None)
        out[f"c{ci}_sha3"] = np.array(sha(np.asarray(pil)))
        out[f"c{ci}_probe3"] = np.asarray(pil)[::37, ::41].copy()
    np.savez_compressed(os.path.join(GOLD, "golden_vip_shapes.npz"), **out)
    print("golden_vip_shapes.npz written", len(VIP_SHAPE_CASES))


# --------------------------------------------------------------------------------------------
class _StubTower:
    """what _process_image_patches reads from the tower (vis_zephyr_arch.py:423,430)."""
    num_patches_per_side = 24
    config = types.SimpleNamespace(image_size=336, patch_size=14)

    def __call__(self, *a, **k):
        raise RuntimeError("stub")


def make_ref_glue(hidden, vocab, merge_type, pins, padding_side="right", max_len=None, feature_fn=None):
    """An object running the reference's own mixin methods around a tiny embedding table."""
    from vis_zephyr.model.vis_zephyr_arch import VisZephyrMetaForCausalLM

    class Inner(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.embed_tokens = torch.nn.Embedding(vocab, hidden)
            with torch.no_grad():
                self.embed_tokens.weight.copy_(torch.arange(vocab, dtype=torch.float32)[:, None].expand(vocab, hidden))
            self.image_newline = torch.nn.Parameter(torch.full((hidden,), -7.0))
            self.vision_tower = _StubTower()

        def get_vision_tower(self):
            return self.vision_tower

    class Glue(VisZephyrMetaForCausalLM):
        def __init__(self):
            self.model = Inner()
            self.config = types.SimpleNamespace(hidden_size=hidden, mm_patch_merge_type=merge_type,
                                                image_aspect_ratio="anyres", mm_grid_pinpoints=str(pins),
                                                tokenizer_padding_side=padding_side)
            if max_len is not None:
                self.config.tokenizer_model_max_length = max_len
            self.device = torch.device("cpu")

        def get_model(self):
            return self.model

        def encode_images(self, images, text_embeddings):
            return feature_fn(images, text_embeddings)

    return Glue()


def gen_merge():
    out = {}
    cases = [("spatial_unpad", (1000, 900)), ("spatial_unpad", (637, 336)), ("spatial_unpad", (336, 900)),
             ("spatial_unpad", (900, 1000)), ("spatial", (1000, 900)), ("spatial", (637, 336)),
             ("spatial_unpad", (1344, 1344)), ("spatial_unpad", (700, 650)), ("spatial_unpad", (1920, 804)),
             ("spatial_unpad", (500, 1500))]
    from vis_zephyr.model.multi_scale_process import calculate_grid_shape
    D = 2
    for ci, (merge, size) in enumerate(cases):
        glue = make_ref_glue(D, 10, merge, PINPOINTS_C3)
        n_w, n_h = calculate_grid_shape(size, str(PINPOINTS_C3), 336)
        T = 1 + n_w * n_h
        feat = torch.arange(T * 576, dtype=torch.float32).reshape(T, 576, 1).expand(T, 576, D).contiguous()
        single = torch.arange(576, dtype=torch.float32).reshape(1, 576, 1).expand(1, 576, D).contiguous()
        merged = glue._process_image_patches([feat, single], [size, (336, 336)])
        out[f"case{ci}_rows"] = merged[0][:, 0].detach().numpy().astype(np.int64)      # -7 marks image_newline
        out[f"case{ci}_single_rows"] = merged[1][:, 0].detach().numpy().astype(np.int64)
        out[f"case{ci}_meta"] = np.array([size[0], size[1], n_w, n_h, T, 1 if "unpad" in merge else 0], np.int64)
    np.savez_compressed(os.path.join(GOLD, "golden_merge.npz"), **out)
    print("golden_merge.npz", {k: v.shape for k, v in out.items() if k.endswith("_rows")})


def gen_splice():
    out = {}
    rng = np.random.default_rng(7)
    D, vocab, Q = 4, 500, 32
    case = 0

    def run(ids, mask, labels, tiles, merge="flat", side="right", max_len=None, pos=None, sizes=None):
        nonlocal case

        def feats(images, text):
            n = images.shape[0]
            base = 100000
            return (base + torch.arange(n * Q, dtype=torch.float32)).reshape(n, Q, 1).expand(n, Q, D).contiguous()

        glue = make_ref_glue(D, vocab, merge, PINPOINTS_C3, side, max_len, feats)
        images = [torch.zeros(t, 3, 336, 336) for t in tiles]
        r = glue.prepare_inputs_labels_for_multimodal(
            torch.from_numpy(ids), None if pos is None else torch.from_numpy(pos),
            None if mask is None else torch.from_numpy(mask), None,
            None if labels is None else torch.from_numpy(labels), images, sizes)
        _, rpos, rmask, _, emb, rlab = r
        k = f"case{case}_"
        out[k + "ids"] = ids
        out[k + "mask"] = mask if mask is not None else np.zeros((0,), np.int64)
        out[k + "labels"] = labels if labels is not None else np.zeros((0,), np.int64)
        out[k + "tiles"] = np.array(tiles, np.int64)
        out[k + "cfg"] = np.array([{"flat": 0, "spatial": 1, "spatial_unpad": 2}[merge], 1 if side == "left" else 0,
                                   -1 if max_len is None else max_len, 0 if pos is None else 1], np.int64)
        out[k + "emb_code"] = emb[:, :, 0].detach().numpy().astype(np.int64)   # token id / 100000+row / 0 pad / -7 newline
        out[k + "out_labels"] = rlab.numpy() if rlab is not None else np.zeros((0,), np.int64)
        out[k + "out_mask"] = rmask.numpy() if rmask is not None else np.zeros((0,), np.int64)
        out[k + "out_mask_dtype"] = np.array(str(rmask.dtype) if rmask is not None else "none")
        out[k + "out_pos"] = rpos.numpy() if rpos is not None else np.zeros((0,), np.int64)
        case += 1

    def mk(B, S, n_img_tokens, pad_from=None, seed=0):
        r = np.random.default_rng(seed)
        ids = r.integers(3, vocab, (B, S)).astype(np.int64)
        for b in range(B):
            for p in sorted(r.choice(np.arange(1, S - 1), n_img_tokens[b], replace=False)):
                ids[b, p] = -200
        mask = np.ones((B, S), np.int64)
        if pad_from is not None:
            for b, p in enumerate(pad_from):
                if p is not None:
                    mask[b, p:] = 0
                    ids[b, p:] = 2
        labels = ids.copy()
        labels[:, : S // 3] = -100
        return ids, mask, labels

    # known-answer case of SURVEY.md 8(c): B=2, S=20, tiles (3,4), sample 1 masked from col 15
    ids, mask, labels = mk(2, 20, [1, 1], [None, 15], seed=1)
    run(ids, mask, labels, [3, 4])
    run(ids, None, None, [3, 4])                      # defaults: no mask / labels
    run(ids, mask.astype(bool), labels, [3, 4], side="left")
    ids, mask, labels = mk(4, 48, [1, 0, 1, 1], [None, 40, 30, None], seed=2)   # a text-only sample consumes a slot
    run(ids, mask, labels, [5, 1, 3, 2])
    run(ids, mask, labels, [5, 1, 3, 2], max_len=100)
    run(ids, mask, labels, [5, 1, 3, 2], side="left", max_len=120,
        pos=np.tile(np.arange(48, dtype=np.int64), (4, 1)))
    ids, mask, labels = mk(3, 33, [0, 1, 1], [None, None, 20], seed=3)          # image token at ragged places
    ids[0, 0] = -200                                                            # image token first
    labels[0, 0] = -200
    run(ids, mask, labels, [1, 1, 1], merge="spatial_unpad", sizes=[(336, 336)] * 3)  # single tile + newline row
    ids, mask, labels = mk(2, 64, [1, 1], None, seed=4)
    run(ids, mask, labels, [4, 4])
    np.savez_compressed(os.path.join(GOLD, "golden_splice.npz"), **out)
    print("golden_splice.npz cases:", case)


# --------------------------------------------------------------------------------------------
class StubTokenizer:
    """word-hash tokenizer: enough to drive tokenizer_image_token and the collator deterministically."""

    def __init__(self, with_bos=True, pad_token_id=2, model_max_length=40):
        self.bos_token_id, self.pad_token_id, self.model_max_length, self.with_bos = 1, pad_token_id, model_max_length, with_bos

    def __call__(self, text):
        ids = ([self.bos_token_id] if self.with_bos else []) + [3 + (sum(map(ord, w)) % 997) for w in text.split()]
        return types.SimpleNamespace(input_ids=ids)


TEXT_PROMPTS = ["<image>\nwhat is shown here ?", "describe <image> and also <image> please", "no picture at all",
                "<image>", "tail image <image>", "", "a <image><image> b"]


def gen_text():
    stub_shapely()
    from vis_zephyr.model.mm_utils import tokenizer_image_token
    # vis_zephyr/train/vis_zephyr_trainer.py imports a symbol that transformers 5.5 no longer has; the
    # collator does not use the trainer, so a placeholder module lets train.py import unmodified
    fake = types.ModuleType("vis_zephyr.train.vis_zephyr_trainer")
    fake.VisZephyrTrainer = object
    fake.maybe_zero = lambda *a, **k: None
    sys.modules["vis_zephyr.train.vis_zephyr_trainer"] = fake
    from vis_zephyr.train.train import DataCollatorForSupervisedDataset
    out = {}
    for bi, with_bos in enumerate([True, False]):
        tok = StubTokenizer(with_bos)
        for pi, prompt in enumerate(TEXT_PROMPTS):
            out[f"tok{bi}_{pi}"] = np.array(tokenizer_image_token(prompt, tok), np.int64)
    rng = np.random.default_rng(11)
    for ci, (B, max_len, mml) in enumerate([(4, 30, 40), (3, 70, 40), (1, 5, 40), (6, 40, 17)]):
        tok = StubTokenizer(True, 2, mml)
        inst = []
        for b in range(B):
            n = int(rng.integers(1, max_len + 1))
            ids = rng.integers(0, 50, n).astype(np.int64)          # includes real tokens equal to the pad id 2
            ids[rng.integers(0, n)] = -200
            labels = ids.copy()
            labels[: n // 2] = -100
            inst.append(dict(input_ids=torch.from_numpy(ids), labels=torch.from_numpy(labels)))
            out[f"col{ci}_ids{b}"], out[f"col{ci}_labels{b}"] = ids, labels
        batch = DataCollatorForSupervisedDataset(tokenizer=tok)(inst)
        out[f"col{ci}_cfg"] = np.array([B, 2, mml], np.int64)
        out[f"col{ci}_out_ids"] = batch["input_ids"].numpy()
        out[f"col{ci}_out_labels"] = batch["labels"].numpy()
        out[f"col{ci}_out_mask"] = batch["attention_mask"].numpy()
    np.savez_compressed(os.path.join(GOLD, "golden_text.npz"), **out)
    print("golden_text.npz written", len(out))


# --------------------------------------------------------------------------------------------
def _reference_glue():
    """the unmodified reference tower + QFormer + mixin around an embedding table, seeded weights"""
    import tempfile
    from PIL import Image
    from transformers import CLIPVisionConfig, CLIPVisionModel
    from vis_zephyr.model.vision_encoder.builder import build_vision_tower
    from vis_zephyr.model.multimodal_projector.builder import build_multimodal_projector
    from vis_zephyr.model.multi_scale_process import process_any_resolution_image
    from vis_zephyr.model.vis_zephyr_arch import VisZephyrMetaForCausalLM

    torch.set_num_threads(os.cpu_count())
    clip_sd = weights.clip_state_dict(0)
    qf_sd = weights.qformer_state_dict(1)
    embed = weights.embed_table(2)
    d = tempfile.mkdtemp(prefix="vz_clip336_")
    cfg = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                           image_size=336, patch_size=14, projection_dim=768, hidden_act="quick_gelu",
                           layer_norm_eps=1e-5)
    hf = CLIPVisionModel(cfg)
    missing = hf.load_state_dict(clip_sd, strict=False)
    print("clip load:", missing)
    hf.save_pretrained(d)
    proc = make_processor()
    proc.save_pretrained(d)

    mcfg = types.SimpleNamespace(mm_vision_tower=d, mm_vision_select_layer="-2,-5,-8,-11,6", mm_vision_select_feature="patch",
                                 hidden_size=4096, mm_patch_merge_type="flat", image_aspect_ratio="anyres",
                                 mm_grid_pinpoints=str(PINPOINTS_C3))
    tower = build_vision_tower(mcfg)
    projector = build_multimodal_projector(mcfg)
    print("qformer load:", projector.load_state_dict(qf_sd))
    projector.eval()

    class Inner(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.embed_tokens = torch.nn.Embedding(embed.shape[0], embed.shape[1])
            with torch.no_grad():
                self.embed_tokens.weight.copy_(embed)
            self.vision_tower = tower
            self.mm_projector = projector

        def get_vision_tower(self):
            return self.vision_tower

    class Glue(VisZephyrMetaForCausalLM):
        def __init__(self):
            self.model = Inner()
            self.config = mcfg
            self.device = torch.device("cpu")
            self.captured = {}

        def get_model(self):
            return self.model

        def encode_images(self, images, text_embeddings):
            feats = self.get_model().get_vision_tower()(images)
            self.captured["tower"] = feats
            out = self.get_model().mm_projector(feats, text_embeddings=text_embeddings)
            self.captured["vis"] = out
            return out

    return Glue(), proc, projector


def gen_model():
    from PIL import Image
    from vis_zephyr.model.multi_scale_process import process_any_resolution_image
    glue, proc, projector = _reference_glue()
    out = {}
    with torch.no_grad():
        # config 1: one 336x336 image, ids of length 64 with one -200 at position 10
        img = synth_image(0, 336, 336)
        px = proc.preprocess(Image.fromarray(img), return_tensors="pt")["pixel_values"]
        ids = torch.randint(3, 32000, (1, 64), generator=torch.Generator().manual_seed(11))
        ids[0, 10] = -200
        r = glue.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, [px], [(336, 336)])
        emb = r[4]
        out["c1_ids"] = ids.numpy()
        out["c1_tower_probe"] = glue.captured["tower"][0, ::48, ::40].numpy().astype(np.float32)
        out["c1_vis"] = glue.captured["vis"][0].numpy().astype(np.float16)
        out["c1_vis_probe32"] = glue.captured["vis"][0, :, ::64].numpy().astype(np.float32)
        out["c1_embeds_shape"] = np.array(emb.shape, np.int64)
        out["c1_embeds_probe"] = emb[0, :, ::512].numpy().astype(np.float32)
        print("c1 done", emb.shape)
        # no-text projector call on the same features
        nt = projector(glue.captured["tower"], text_embeddings=None)
        out["c1_vis_notext_probe32"] = nt[0, :, ::64].numpy().astype(np.float32)
        # anyres 1000x900 -> 5 tiles + a second sample with shorter text (exercises zero-padded text rows)
        img3 = synth_image(0, 1000, 900)
        px3 = process_any_resolution_image(Image.fromarray(img3), proc, PINPOINTS_C3)
        img4 = synth_image(1, 637, 336)
        px4 = process_any_resolution_image(Image.fromarray(img4), proc, PINPOINTS_C3)
        ids2 = torch.randint(3, 32000, (2, 48), generator=torch.Generator().manual_seed(12))
        ids2[0, 5] = -200
        ids2[1, 20] = -200
        ids2[1, 30:] = 2
        mask2 = (ids2 != 2).long()
        mask2[0, :] = 1
        labels2 = ids2.clone()
        labels2[:, :16] = -100
        r = glue.prepare_inputs_labels_for_multimodal(ids2, None, mask2, None, labels2, [px3, px4],
                                                      [(1000, 900), (637, 336)])
        out["c3_ids"] = ids2.numpy()
        out["c3_mask"] = mask2.numpy()
        out["c3_labels"] = labels2.numpy()
        out["c3_tiles"] = np.array([px3.shape[0], px4.shape[0]], np.int64)
        out["c3_vis"] = glue.captured["vis"].numpy().astype(np.float16)
        out["c3_vis_probe32"] = glue.captured["vis"][:, :, ::64].numpy().astype(np.float32)
        out["c3_out_labels"] = r[5].numpy()
        out["c3_out_mask"] = r[2].numpy()
        out["c3_embeds_shape"] = np.array(r[4].shape, np.int64)
        out["c3_embeds_probe"] = r[4][:, :, ::512].numpy().astype(np.float32)
        print("c3 done", r[4].shape)
    np.savez_compressed(os.path.join(GOLD, "golden_model.npz"), **out)
    print("golden_model.npz written")


def gen_model_long():
    """BASELINE config 5 regime (S = 2048): the reference's prepare_inputs_labels_for_multimodal on B = 2,
    sample 0 with ~2000 real tokens and ONE 336x336 tile, sample 1 with ~300 real tokens + pads and a
    3-tile anyres image -> text conditioning rows L = 2047 with a large zero-padded tail (quirk Q3)."""
    from PIL import Image
    from vis_zephyr.model.multi_scale_process import process_any_resolution_image
    glue, proc, projector = _reference_glue()
    out = {}
    S = 2048
    with torch.no_grad():
        g = torch.Generator().manual_seed(21)
        ids = torch.randint(3, 32000, (2, S), generator=g)
        ids[ids == 2] = 3
        ids[0, 17] = -200
        ids[0, 2001:] = 2                   # 2000 real tokens + the image slot, then pads
        ids[1, 9] = -200
        ids[1, 301:] = 2
        mask = (ids != 2).long()
        labels = ids.clone()
        labels[0, :700] = -100
        labels[1, :100] = -100
        labels[ids == 2] = -100
        img0 = synth_image(7, 336, 336)
        px0 = proc.preprocess(Image.fromarray(img0), return_tensors="pt")["pixel_values"]      # [1,3,336,336]
        img1 = synth_image(1, 637, 336)
        px1 = process_any_resolution_image(Image.fromarray(img1), proc, PINPOINTS_SHIPPED)     # [3,3,336,336]
        r = glue.prepare_inputs_labels_for_multimodal(ids, None, mask, None, labels, [px0, px1],
                                                      [(336, 336), (637, 336)])
        out["ids"], out["mask"], out["labels"] = ids.numpy(), mask.numpy(), labels.numpy()
        out["tiles"] = np.array([px0.shape[0], px1.shape[0]], np.int64)
        out["vis"] = glue.captured["vis"].numpy().astype(np.float16)               # [4,32,4096]
        out["vis_probe32"] = glue.captured["vis"][:, :, ::64].numpy().astype(np.float32)
        out["out_labels"] = r[5].numpy()
        out["out_mask"] = r[2].numpy()
        out["embeds_shape"] = np.array(r[4].shape, np.int64)
        out["embeds_probe"] = r[4][:, ::7, ::512].numpy().astype(np.float32)
        print("long done", r[4].shape)
    np.savez_compressed(os.path.join(GOLD, "golden_model_long.npz"), **out)
    print("golden_model_long.npz written")


MODE_SIZES = [(700, 500), (420, 901), (336, 336), (1000, 1000), (301, 640)]


def gen_modes():
    """mm_utils.process_images (expand2square / centre-square crop / LANCZOS squash / plain) through the
    Pillow-backed CLIP processor (what the reference's pinned transformers 4.52.4 runs): SHA-256 of the
    f32 pixel tensors, plus the 768-entry LUT of that processor."""
    from PIL import Image
    from transformers.models.clip import CLIPImageProcessorPil
    from vis_zephyr.model.mm_utils import process_images
    proc = CLIPImageProcessorPil(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336},
                                 image_mean=[0.48145466, 0.4578275, 0.40821073],
                                 image_std=[0.26862954, 0.26130258, 0.27577711], resample=3)
    ramp = np.zeros((336, 336, 3), np.uint8)
    ramp[:] = (np.arange(336) % 256)[None, :, None]
    px = proc.preprocess(Image.fromarray(ramp), return_tensors="pt")["pixel_values"][0].numpy()
    out = {"lut": np.stack([px[c, 0, :256] for c in range(3)]).astype(np.float32),
           "sizes": np.array(MODE_SIZES, np.int64)}
    rng = np.random.default_rng(12)
    for si, (W, H) in enumerate(MODE_SIZES):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        for mode in ("pad", "square", "resize", "plain"):
            cfg = types.SimpleNamespace(aspect_ratio_mode=mode)
            t = process_images(Image.fromarray(img), proc, cfg).numpy().astype(np.float32)
            assert t.shape == (3, 336, 336)
            out[f"s{si}_{mode}_sha"] = np.array(sha(t))
            out[f"s{si}_{mode}_probe"] = t[:, ::48, ::48].copy()
    np.savez_compressed(os.path.join(GOLD, "golden_modes.npz"), **out)
    print("golden_modes.npz written", len(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-model", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    todo = a.only.split(",") if a.only else ["pixels", "vip", "vip_shapes", "merge", "splice", "text", "modes", "model", "model_long"]
    if "pixels" in todo: gen_pixels()
    if "vip" in todo: gen_vip()
    if "vip_shapes" in todo: gen_vip_shapes()
    if "merge" in todo: gen_merge()
    if "splice" in todo: gen_splice()
    if "text" in todo: gen_text()
    if "modes" in todo: gen_modes()
    if "model" in todo and not a.skip_model: gen_model()
    if "model_long" in todo and not a.skip_model: gen_model_long()
