"""numpy restatement of the Pillow / HF-processor arithmetic on the pixel half of the path
(test infrastructure; pinned bit-exact against Pillow itself in tests/test_oracle_pixels.py).

alpha_composite_rgb   Image.alpha_composite(image.convert("RGBA"), overlay).convert("RGB")
                      vip_processor/conversation_generator.py:143-146 (Pillow AlphaComposite.c)
draw_rectangle_mask   ImageDraw.rectangle(outline, width)  vip_processor/shape_draw.py:68-71
                      (Pillow Draw.c ImagingDrawRectangle, fill=0)
lanczos_resize        image.resize(size, LANCZOS)  multi_scale_process.py:86-89,171-174
                      (Pillow Resample.c: precompute_coeffs, normalize_coeffs_8bpc, horizontal pass,
                      u8 intermediate, vertical pass)
process_any_resolution  multi_scale_process.py:136-183 on a u8 array + a 768-entry LUT for
                      CLIPImageProcessor.preprocess
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
TILE = 336


def alpha_composite_rgb(dst_rgb, overlay_rgba):
    """dst u8 [H,W,3] opaque, overlay u8 [H,W,4] -> u8 [H,W,3]."""
    dst = dst_rgb.astype(np.uint32)
    src = overlay_rgba[..., :3].astype(np.uint32)
    a = overlay_rgba[..., 3:4].astype(np.uint32)
    t = src * (a * 128) + dst * ((255 - a) * 128) + (0x80 << 7)
    out = ((((t >> 8) + t) >> 8) >> 7).astype(np.uint8)
    return np.where(a == 0, dst_rgb, out)


def draw_rectangle_mask(H, W, bbox, width):
    """bool [H,W]: pixels ImageDraw.rectangle(bbox, outline=..., width=width) overwrites.
    Literal union of the 4*width lines ImagingDrawRectangle draws (they leave the box when the
    outline is wider than the box); ImageDraw skips the call entirely for width == 0."""
    m = np.zeros((H, W), bool)
    if width == 0:
        return m
    x0, y0, x1, y1 = (int(v) for v in bbox)
    ys, xs = np.mgrid[0:H, 0:W]
    in_x = (xs >= x0) & (xs <= x1)
    # Draw.c line32 with dx == 0 paints |dy| points starting at the first end point (the last
    # end point is NOT painted)
    va, vb = y0 + width, y1 - width + 1
    in_v = ((ys >= va) & (ys < vb)) if vb >= va else ((ys <= va) & (ys > vb))
    for i in range(width):
        m |= in_x & ((ys == y0 + i) | (ys == y1 - i))
        m |= in_v & ((xs == x1 - i) | (xs == x0 + i))
    return m


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x):
    return _sinc(x) * _sinc(x / 3) if -3.0 <= x < 3.0 else 0.0


def _bicubic(x):
    """Pillow Resample.c bicubic_filter (a = -0.5)."""
    a = -0.5
    x = -x if x < 0.0 else x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


_FILTERS = {"lanczos": (_lanczos, 3.0), "bicubic": (_bicubic, 2.0)}


def coeff_matrix(in_size, out_size, filt="lanczos"):
    """Dense int64 [out,in] fixed-point matrix of one axis for a Pillow filter (zeros outside the taps)."""
    fn, fsupport = _FILTERS[filt]
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = fsupport * fs
    K = np.zeros((out_size, in_size), np.int64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size)
        w = [fn((x + xmin - center + 0.5) / fs) for x in range(xmax - xmin)]
        ww = 0.0
        for v in w:
            ww += v
        for i, v in enumerate(w):
            if ww != 0.0:
                v = v / ww
            K[xx, xmin + i] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
    return K


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def lanczos_resize(img, size):
    """img u8 [H,W,C] -> u8 [h,w,C] for size=(w,h), like PIL.Image.resize(size, LANCZOS)."""
    return pil_resize(img, size, "lanczos")


def pil_resize(img, size, filt="lanczos"):
    """img u8 [H,W,C] -> u8 [h,w,C] for size=(w,h), like PIL.Image.resize(size, LANCZOS | BICUBIC)."""
    H, W = img.shape[:2]
    w, h = size
    out = img
    if w != W:  # horizontal pass first
        Kh = coeff_matrix(W, w, filt)
        # float64 BLAS is exact here: |sum| < 2**53
        acc = np.einsum("ox,yxc->yoc", Kh.astype(np.float64), out.astype(np.float64), optimize=True).astype(np.int64)
        acc += 1 << (PRECISION_BITS - 1)
        out = _clip8(acc)
    if h != H:
        Kv = coeff_matrix(H, h, filt)
        acc = np.einsum("oy,yxc->oxc", Kv.astype(np.float64), out.astype(np.float64), optimize=True).astype(np.int64)
        acc += 1 << (PRECISION_BITS - 1)
        out = _clip8(acc)
    return out.copy() if out is img else out


def select_best_fit_resolution(original_resolution, possible_resolutions):
    ow, oh = original_resolution
    best, max_eff, min_waste = None, 0, float("inf")
    for w, h in possible_resolutions:
        s = min(w / ow, h / oh)
        dw, dh = int(ow * s), int(oh * s)
        eff = min(dw * dh, ow * oh)
        waste = w * h - eff
        if eff > max_eff or (eff == max_eff and waste < min_waste):
            max_eff, min_waste, best = eff, waste, (w, h)
    return best


def anyres_tiles_u8(img, pinpoints):
    """u8 [T,336,336,3]: [global squashed view] + row-major tiles of the resized, centre-padded image."""
    H, W = img.shape[:2]
    bw, bh = select_best_fit_resolution((W, H), pinpoints)
    s = min(bw / W, bh / H)
    nw, nh = int(W * s), int(H * s)
    resized = lanczos_resize(img, (nw, nh))
    canvas = np.zeros((bh, bw, 3), np.uint8)
    px, py = (bw - nw) // 2, (bh - nh) // 2
    canvas[py:py + nh, px:px + nw] = resized
    tiles = [lanczos_resize(img, (TILE, TILE))]
    for i in range(0, bh, TILE):
        for j in range(0, bw, TILE):
            tiles.append(canvas[i:i + TILE, j:j + TILE])
    return np.stack(tiles)


def expand2square(img, bg):
    """mm_utils.expand2square (mm_utils.py:16-35) on a u8 [H,W,3] array."""
    H, W = img.shape[:2]
    if W == H:
        return img
    m = max(W, H)
    out = np.empty((m, m, 3), np.uint8)
    out[:] = np.asarray(bg, np.uint8)
    if W > H:
        y = (W - H) // 2
        out[y:y + H] = img
    else:
        x = (H - W) // 2
        out[:, x:x + W] = img
    return out


def clip_processor_u8(img):
    """The geometric part of CLIPImageProcessor.preprocess with size={'shortest_edge': 336}, crop 336
    (transformers 4.52.4, the version the reference pins: PIL BICUBIC resize of the short side to 336,
    long side int(336 * long / short), then centre crop with floor-divided offsets)."""
    H, W = img.shape[:2]
    if W <= H:
        nw, nh = TILE, int(TILE * H / W)
    else:
        nw, nh = int(TILE * W / H), TILE
    r = pil_resize(img, (nw, nh), "bicubic")
    top, left = (nh - TILE) // 2, (nw - TILE) // 2
    return r[top:top + TILE, left:left + TILE]


def process_images_u8(img, mode, image_mean=(0.48145466, 0.4578275, 0.40821073)):
    """mm_utils.process_images (mm_utils.py:38-87) / train.py:570-590 up to the u8 336x336 tile:
    'pad' | 'resize' | 'square' | anything else ('plain')."""
    if mode == "pad":
        img = expand2square(img, tuple(int(x * 255) for x in image_mean))
    elif mode == "resize":
        img = pil_resize(img, (TILE, TILE), "lanczos")
    elif mode == "square":
        H, W = img.shape[:2]
        m = min(W, H)
        left, top = int((W - m) / 2), int((H - m) / 2)
        img = img[top:top + m, left:left + m]
    return clip_processor_u8(img)


def normalize_lut(tiles_u8, lut):
    """u8 [T,336,336,3] + f32 lut [3,256] -> f32 [T,3,336,336] (CLIPImageProcessor.preprocess)."""
    out = np.empty((tiles_u8.shape[0], 3) + tiles_u8.shape[1:3], np.float32)
    for c in range(3):
        out[:, c] = lut[c][tiles_u8[..., c]]
    return out


def process_any_resolution(img, pinpoints, lut):
    return normalize_lut(anyres_tiles_u8(img, pinpoints), lut)


def patchify(pixel_values):
    """f32 [T,3,336,336] -> f32 [T*576, 588] in (c,ky,kx) order (Conv2d(3,1024,14,14) as im2col)."""
    T = pixel_values.shape[0]
    x = pixel_values.reshape(T, 3, 24, 14, 24, 14).transpose(0, 2, 4, 1, 3, 5)
    return np.ascontiguousarray(x).reshape(T * 576, 588)
