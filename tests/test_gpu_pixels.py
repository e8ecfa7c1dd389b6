"""GPU parity of the fused preprocess kernel: bit-exact against the numpy oracle (which is pinned to
Pillow and to the reference's digests), for anyres tiling, identity 336 input and visual prompts."""
import numpy as np
import pytest
import torch

from helpers import PINPOINTS_C3, PINPOINTS_SHIPPED, sha, synth_image, vip_overlays

pytestmark = pytest.mark.gpu


def _lut(golden_dir):
    return np.load(f"{golden_dir}/golden_pixels.npz")["lut"]


@pytest.mark.parametrize("case", [(0, 1000, 900, PINPOINTS_C3), (1, 637, 336, PINPOINTS_SHIPPED),
                                  (2, 336, 900, PINPOINTS_C3), (3, 1920, 804, PINPOINTS_SHIPPED),
                                  (4, 336, 336, PINPOINTS_SHIPPED), (6, 250, 180, PINPOINTS_C3)])
def test_anyres_chw_bit_exact(case, golden_dir):
    import vision_zephyr_b200 as vz
    from oracle import pil_ops as P
    i, W, H, pins = case
    lut = _lut(golden_dir)
    img = synth_image(i, W, H)
    ref = P.process_any_resolution(img, pins, lut)
    got = vz.process_any_resolution_images([torch.from_numpy(img).cuda()], pins, lut, out_mode="chw")[0]
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    assert got.shape == ref.shape
    bad = int((got != ref).sum())
    print(f"case {i} {W}x{H}: tiles={ref.shape[0]} mismatching elements={bad}")
    assert bad == 0
    # the reference's own digest travels as a golden
    from helpers import sha
    assert sha(got) == str(np.load(f"{golden_dir}/golden_pixels.npz")[f"case{i}_sha"])


def test_anyres_patches_are_bf16_im2col_of_chw(golden_dir):
    import vision_zephyr_b200 as vz
    from oracle import pil_ops as P
    lut = _lut(golden_dir)
    imgs = [synth_image(0, 1000, 900), synth_image(5, 700, 650)]
    dev = [torch.from_numpy(x).cuda() for x in imgs]
    pb = vz.process_any_resolution_images(dev, PINPOINTS_C3, lut, out_mode="patches")
    assert pb.tiles_per_image == [5, 5] and pb.patches.shape == (10 * 576, 592)
    ref = np.concatenate([P.patchify(P.process_any_resolution(x, PINPOINTS_C3, lut)) for x in imgs])
    ref = torch.from_numpy(ref).to(torch.bfloat16)
    got = pb.patches.cpu()
    assert torch.equal(got[:, :588], ref)
    assert (got[:, 588:] == 0).all()


def test_visual_prompts_fixed336_bit_exact(golden_dir):
    """config 2: rectangle rasterised in the kernel, mask/arrow as host-rasterised RGBA layers."""
    import vision_zephyr_b200 as vz
    from oracle import pil_ops as P
    g = np.load(f"{golden_dir}/golden_vip.npz")
    lut = _lut(golden_dir)
    imgs, prompts, refs = [], [], []
    for i in range(4):
        img = synth_image(100 + i, 336, 336)
        ref = img
        plist = []
        for prim in vip_overlays(i, g[f"img{i}_specs"]):
            if prim[0] == "rectangle":
                _, bbox, width, rgba = prim
                plist.append(vz.VisualPrompt("rectangle", rgba=rgba, bbox=tuple(bbox), width=width))
                layer = np.zeros((336, 336, 4), np.uint8)
                layer[P.draw_rectangle_mask(336, 336, bbox, width)] = rgba
            else:
                layer = prim[1]
                plist.append(vz.VisualPrompt("layer", layer=layer))
            ref = P.alpha_composite_rgb(ref, layer)
        assert np.array_equal(ref[::31, ::29], g[f"img{i}_final_probe"])
        imgs.append(torch.from_numpy(img).cuda())
        prompts.append(plist)
        refs.append(P.normalize_lut(ref[None], lut))
    got = vz.process_fixed_images(imgs, lut, prompts, out_mode="chw")
    for i in range(4):
        assert np.array_equal(got[i].cpu().numpy(), refs[i]), i


def test_rectangle_kernel_raster_matches_pillow_semantics(golden_dir):
    import vision_zephyr_b200 as vz
    from oracle import pil_ops as P
    lut = _lut(golden_dir)
    rng = np.random.default_rng(5)
    imgs, prompts, refs = [], [], []
    for i in range(16):
        img = synth_image(200 + i, 336, 336)
        x0, y0 = rng.uniform(-10, 300), rng.uniform(-10, 300)
        bbox = (x0, y0, x0 + rng.uniform(0, 120), y0 + rng.uniform(0, 120))
        width = int(rng.integers(0, 14))
        rgba = tuple(int(v) for v in rng.integers(0, 256, 4))
        layer = np.zeros((336, 336, 4), np.uint8)
        layer[P.draw_rectangle_mask(336, 336, bbox, width)] = rgba
        refs.append(P.normalize_lut(P.alpha_composite_rgb(img, layer)[None], lut))
        imgs.append(torch.from_numpy(img).cuda())
        prompts.append([vz.VisualPrompt("rectangle", rgba=rgba, bbox=bbox, width=width)])
    got = vz.process_fixed_images(imgs, lut, prompts, out_mode="chw")
    for i in range(16):
        assert np.array_equal(got[i].cpu().numpy(), refs[i]), i


def test_blend_then_resize_on_large_image(golden_dir):
    """visual prompt on a non-square source, then the anyres path (blend feeds the resampler)."""
    import vision_zephyr_b200 as vz
    from oracle import pil_ops as P
    lut = _lut(golden_dir)
    img = synth_image(9, 900, 500)
    layer = np.zeros((500, 900, 4), np.uint8)
    layer[100:300, 200:700] = (10, 200, 30, 77)
    ref = P.process_any_resolution(P.alpha_composite_rgb(img, layer), PINPOINTS_C3, lut)
    got = vz.process_any_resolution_images([torch.from_numpy(img).cuda()], PINPOINTS_C3, lut,
                                           prompts=[[vz.VisualPrompt("layer", layer=layer)]], out_mode="chw")[0]
    assert np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.parametrize("mode", ["pad", "square", "resize", "plain"])
def test_process_images_modes_bit_exact(mode, golden_dir):
    """mm_utils.process_images / the training loader's 'pad' and plain branches (expand2square, centre crop,
    LANCZOS squash, CLIP processor BICUBIC resize + centre crop): kernel == oracle, with a visual prompt on
    one image (prompts stay in image coordinates under padding / cropping)"""
    import vision_zephyr_b200 as vz
    from oracle import pil_ops as P
    lut = _lut(golden_dir)
    sizes = [(700, 500), (420, 901), (336, 336), (1000, 1000), (301, 640)]
    imgs = [synth_image(20 + i, w, h) for i, (w, h) in enumerate(sizes)]
    layer = np.zeros((500, 700, 4), np.uint8)
    layer[50:450, 100:650] = (250, 20, 30, 99)
    prompts = [[vz.VisualPrompt("layer", layer=layer),
                vz.VisualPrompt("rectangle", rgba=(0, 0, 255, 200), bbox=(5, 5, 690, 480), width=4)]] + [[]] * 4
    got = vz.process_fixed_images([torch.from_numpy(x).cuda() for x in imgs], lut, prompts, out_mode="chw", mode=mode)
    for i, img in enumerate(imgs):
        src = img
        if i == 0:
            src = P.alpha_composite_rgb(img, layer)
            m = P.draw_rectangle_mask(500, 700, (5, 5, 690, 480), 4)
            ov = np.zeros((500, 700, 4), np.uint8)
            ov[m] = (0, 0, 255, 200)
            src = P.alpha_composite_rgb(src, ov)
        ref = P.normalize_lut(P.process_images_u8(src, mode)[None], lut)[0]
        assert np.array_equal(got[i][0].cpu().numpy(), ref), (mode, i)


def test_process_images_modes_match_reference_golden(golden_dir):
    """the four aspect modes through the kernel == SHA-256 of mm_utils.process_images' own output
    (golden_modes.npz, generated by the reference with the Pillow-backed CLIP processor)"""
    import vision_zephyr_b200 as vz
    g = np.load(f"{golden_dir}/golden_modes.npz")
    lut = g["lut"]
    rng = np.random.default_rng(12)
    imgs = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for (W, H) in g["sizes"].tolist()]
    # the generator drew image s, then ran its 4 modes: same draw order here
    for mode in ("pad", "square", "resize", "plain"):
        got = vz.process_fixed_images([torch.from_numpy(x).cuda() for x in imgs], lut, out_mode="chw", mode=mode)
        for si in range(len(imgs)):
            assert sha(got[si][0].cpu().numpy().astype(np.float32)) == str(g[f"s{si}_{mode}_sha"]), (si, mode)


def test_dp4a_two_kernel_and_fused_forms_agree(golden_dir, monkeypatch):
    """vz_preprocess3 (dp4a: limb-split coefficients, planar intermediate, windows aligned to four) ==
    vz_preprocess2 (one multiply per tap, RGBX intermediate) == vz_preprocess (fused, one CTA per band) bit for bit,
    on anyres + fixed-mode views with visual prompts and canvas padding, in both output layouts"""
    import vision_zephyr_b200 as vz
    from vision_zephyr_b200 import anyres
    from vision_zephyr_b200.preprocess import build_plan, run_plan
    lut = _lut(golden_dir)
    sizes = [(1000, 900), (637, 336), (336, 900), (1344, 1344), (301, 640), (1920, 804), (250, 180), (333, 1000)]
    imgs = [torch.from_numpy(synth_image(60 + i, w, h)).cuda() for i, (w, h) in enumerate(sizes)]
    layer = np.zeros((900, 1000, 4), np.uint8)
    layer[100:700, 50:900] = (20, 250, 30, 140)
    prompts = [[vz.VisualPrompt("layer", layer=layer), vz.VisualPrompt("rectangle", rgba=(255, 0, 0, 255), bbox=(10, 20, 950, 800), width=5)]]
    prompts += [[] for _ in sizes[1:]]
    views, canvases = [], []
    for i, (w, h) in enumerate(sizes):
        if i < 4 or i in (5, 6):
            views.append(anyres.anyres_views((w, h), PINPOINTS_C3)[0]); canvases.append(None)
        else:
            v, c = anyres.fixed_view((w, h), "pad" if i == 4 else "square"); views.append(v); canvases.append(c)
    plan = build_plan(imgs, views, lut, prompts, canvases)
    assert plan.max_ksize > 1 and plan.n_hviews >= len(sizes)
    monkeypatch.delenv("VZ_PRE_FUSED", raising=False)
    for mode in ("patches", "chw"):
        monkeypatch.setenv("VZ_PRE_FORM", "dp")
        dp = run_plan(plan, mode).clone()
        monkeypatch.setenv("VZ_PRE_FORM", "two")
        two = run_plan(plan, mode).clone()
        monkeypatch.setenv("VZ_PRE_FORM", "fused")
        fused = run_plan(plan, mode)
        torch.cuda.synchronize()
        assert torch.equal(two, fused), mode
        bad = (dp != two)
        assert not bad.any(), (mode, int(bad.sum()), bad.nonzero()[:8].tolist())


def test_dp4a_form_large_downscale_and_upscale(golden_dir, monkeypatch):
    """the dp4a form at the ends of its range: 37-tap windows (1920 -> 336, 10 groups per row) and an upscale
    (7-tap windows, 3 groups), against the oracle (== Pillow)"""
    import vision_zephyr_b200 as vz
    from oracle import pil_ops as P
    lut = _lut(golden_dir)
    monkeypatch.setenv("VZ_PRE_FORM", "dp")
    # (400, 3000): a 14-row band spans 45 four-row groups, more than the dp4a pass stages -> vz_preprocess2 takes over
    for i, (w, h) in enumerate([(1920, 1500), (150, 130), (2600, 400), (400, 3000)]):
        img = synth_image(80 + i, w, h)
        got = vz.process_fixed_images([torch.from_numpy(img).cuda()], lut, out_mode="chw", mode="resize")[0][0]
        ref = P.normalize_lut(P.process_images_u8(img, "resize")[None], lut)[0]
        assert np.array_equal(got.cpu().numpy(), ref), (w, h)


def test_identity_form_matches_fused_kernel(golden_dir, monkeypatch):
    """config 2 shape (336 x 336 images + visual prompts): the vectorised identity kernel (four pixels per thread)
    == the fused kernel's per-pixel identity path, bit for bit, in both output layouts"""
    import vision_zephyr_b200 as vz
    from vision_zephyr_b200.preprocess import build_plan, run_plan
    from vision_zephyr_b200 import anyres
    lut = _lut(golden_dir)
    rng = np.random.default_rng(3)
    imgs, prompts = [], []
    for i in range(5):
        imgs.append(torch.from_numpy(synth_image(120 + i, 336, 336)).cuda())
        lay = np.zeros((336, 336, 4), np.uint8)
        y0, x0 = int(rng.integers(0, 200)), int(rng.integers(0, 200))
        lay[y0:y0 + 120, x0:x0 + 97] = (int(rng.integers(0, 256)), 255, 7, int(rng.integers(1, 256)))
        lay2 = rng.integers(0, 256, (336, 336, 4), dtype=np.uint8)          # every alpha value, every pixel
        plist = [vz.VisualPrompt("rectangle", rgba=(255, 0, 0, 128), bbox=(30.7, 40.2, 200.9, 220.1), width=3),
                 vz.VisualPrompt("layer", layer=lay), vz.VisualPrompt("layer", layer=lay2),
                 vz.VisualPrompt("rectangle", rgba=(1, 2, 3, 255), bbox=(-5, -7, 340, 335), width=9)]
        prompts.append(plist[: i])                                           # 0 .. 4 instances
    views = [anyres.single_view((336, 336)) for _ in imgs]
    plan = build_plan(imgs, views, lut, prompts)
    assert plan.all_identity and plan.max_ksize == 1
    monkeypatch.delenv("VZ_PRE_FUSED", raising=False)
    for mode in ("patches", "chw"):
        monkeypatch.setenv("VZ_PRE_FORM", "dp")
        new = run_plan(plan, mode).clone()
        monkeypatch.setenv("VZ_PRE_FORM", "fused")
        old = run_plan(plan, mode)
        torch.cuda.synchronize()
        assert torch.equal(new, old), mode
