"""Pin the pixel oracle: bit-exact against Pillow (present on every box) and against the golden
SHA-256 digests produced by the reference's own process_any_resolution_image / image_blending."""
import numpy as np
import pytest
from PIL import Image, ImageDraw

from helpers import hf_processor, sha, synth_image, vip_overlays
from oracle import pil_ops as P


@pytest.mark.parametrize("W,H,w,h", [(1000, 900, 672, 604), (637, 336, 336, 336), (200, 150, 336, 252),
                                     (1920, 804, 336, 336), (681, 336, 1008, 497), (336, 336, 336, 336),
                                     (50, 40, 336, 336), (3000, 17, 336, 336)])
def test_lanczos_matches_pillow(W, H, w, h):
    img = np.random.default_rng(W + H).integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((w, h), Image.Resampling.LANCZOS))
    assert np.array_equal(ref, P.lanczos_resize(img, (w, h)))


def test_alpha_composite_matches_pillow_exhaustively():
    src, dst, a = np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing="ij")
    dst3 = np.stack([dst.reshape(4096, 4096)] * 3, -1).astype(np.uint8)
    ov = np.stack([src.reshape(4096, 4096)] * 3 + [a.reshape(4096, 4096)], -1).astype(np.uint8)
    ref = np.asarray(Image.alpha_composite(Image.fromarray(dst3).convert("RGBA"), Image.fromarray(ov, "RGBA")).convert("RGB"))
    assert np.array_equal(ref, P.alpha_composite_rgb(dst3, ov))


def test_rectangle_outline_matches_pillow():
    rng = np.random.default_rng(0)
    for _ in range(500):
        H, W = int(rng.integers(20, 120)), int(rng.integers(20, 120))
        x0, y0 = rng.uniform(-10, W - 5), rng.uniform(-10, H - 5)
        x1, y1 = x0 + rng.uniform(0, W), y0 + rng.uniform(0, H)
        wd = int(rng.integers(0, 12))
        im = Image.new("RGBA", (W, H), (0, 0, 0, 0))
        ImageDraw.Draw(im).rectangle([(x0, y0), (x1, y1)], outline=(255, 0, 0, 128), width=wd)
        assert np.array_equal(np.asarray(im)[..., 3] > 0, P.draw_rectangle_mask(H, W, (x0, y0, x1, y1), wd))


def test_anyres_pipeline_matches_reference_digests(golden_dir):
    g = np.load(f"{golden_dir}/golden_pixels.npz")
    lut = g["lut"]
    # the LUT of this installation's processor must be the one the goldens were made with
    from vision_zephyr_b200.preprocess import lut_from_processor
    assert np.array_equal(lut, lut_from_processor(hf_processor()))
    for i in range(7):
        W, H, T, n_w, n_h = (int(v) for v in g[f"case{i}_meta"])
        pins = g[f"case{i}_pins"].tolist()
        px = P.process_any_resolution(synth_image(i, W, H), pins, lut)
        assert px.shape == (T, 3, 336, 336)
        assert np.array_equal(px[:, :, ::67, ::59], g[f"case{i}_probe"])
        assert sha(px) == str(g[f"case{i}_sha"]), f"case {i}"


def test_visual_prompt_blend_matches_reference_digests(golden_dir):
    g = np.load(f"{golden_dir}/golden_vip.npz")
    for i in range(4):
        img = synth_image(100 + i, 336, 336)
        for inst, prim in enumerate(vip_overlays(i, g[f"img{i}_specs"])):
            if prim[0] == "rectangle":
                _, bbox, width, rgba = prim
                layer = np.zeros((336, 336, 4), np.uint8)
                layer[P.draw_rectangle_mask(336, 336, bbox, width)] = rgba
            else:
                layer = prim[1]
            img = P.alpha_composite_rgb(img, layer)
            assert sha(img) == str(g[f"img{i}_after{inst}_sha"]), (i, inst)


def test_patchify_is_conv_im2col():
    import torch
    rng = np.random.default_rng(3)
    px = rng.standard_normal((2, 3, 336, 336)).astype(np.float32)
    w = rng.standard_normal((8, 3, 14, 14)).astype(np.float32)
    ref = torch.nn.functional.conv2d(torch.from_numpy(px), torch.from_numpy(w), stride=14).flatten(2).transpose(1, 2)
    got = P.patchify(px) @ w.reshape(8, 588).T
    assert np.allclose(ref.reshape(-1, 8).numpy(), got, atol=1e-3)


def test_bicubic_resize_matches_pillow():
    """the oracle's bicubic pass (what CLIPImageProcessor.resize asks Pillow for) is bit-exact"""
    from PIL import Image
    rng = np.random.default_rng(11)
    for (W, H), (w, h) in [((700, 500), (470, 336)), ((336, 900), (336, 900)), ((500, 333), (504, 336)),
                           ((1300, 1300), (336, 336)), ((200, 150), (448, 336)), ((337, 336), (337, 336))]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((w, h), Image.BICUBIC))
        assert np.array_equal(P.pil_resize(img, (w, h), "bicubic"), ref), ((W, H), (w, h))


def test_process_images_modes_match_pillow_flow():
    """mm_utils.process_images / train.py:570-590 geometry (expand2square, centre crop, LANCZOS squash, then the
    CLIP processor's BICUBIC short-side resize + centre crop) restated on arrays == the same steps in Pillow"""
    from PIL import Image
    rng = np.random.default_rng(12)
    mean = (0.48145466, 0.4578275, 0.40821073)
    for (W, H) in [(700, 500), (420, 901), (336, 336), (1000, 1000), (301, 640)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        pil = Image.fromarray(img)
        for mode in ("pad", "square", "resize", "plain"):
            im = pil
            if mode == "pad":
                bg = tuple(int(x * 255) for x in mean)
                m = max(W, H)
                if W != H:
                    im = Image.new("RGB", (m, m), bg)
                    im.paste(pil, (0, (W - H) // 2) if W > H else ((H - W) // 2, 0))
            elif mode == "resize":
                im = pil.resize((336, 336), Image.Resampling.LANCZOS)
            elif mode == "square":
                m = min(W, H)
                left, top = int((W - m) / 2), int((H - m) / 2)
                im = pil.crop((left, top, left + m, top + m))
            w, h = im.size
            nw, nh = (336, int(336 * h / w)) if w <= h else (int(336 * w / h), 336)
            r = im.resize((nw, nh), Image.BICUBIC)
            left, top = (nw - 336) // 2, (nh - 336) // 2
            ref = np.asarray(r.crop((left, top, left + 336, top + 336)))
            assert np.array_equal(P.process_images_u8(img, mode, mean), ref), (W, H, mode)


def test_process_images_modes_match_reference_golden(golden_dir):
    """oracle process_images_u8 + LUT == mm_utils.process_images itself (golden_modes.npz: the reference
    function run through the Pillow-backed CLIP processor, the backend its pinned transformers uses)"""
    g = np.load(f"{golden_dir}/golden_modes.npz")
    lut = g["lut"]
    rng = np.random.default_rng(12)
    for si, (W, H) in enumerate(g["sizes"].tolist()):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        for mode in ("pad", "square", "resize", "plain"):
            got = P.normalize_lut(P.process_images_u8(img, mode)[None], lut)[0]
            assert np.array_equal(got[:, ::48, ::48], g[f"s{si}_{mode}_probe"]), (W, H, mode)
            assert sha(got.astype(np.float32)) == str(g[f"s{si}_{mode}_sha"]), (W, H, mode)
