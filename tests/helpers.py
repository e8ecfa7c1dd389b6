"""Shared test helpers: synthetic inputs identical to the ones oracle/gen_golden.py used."""
import hashlib
import random

import numpy as np

PINPOINTS_SHIPPED = [[336, 672], [672, 336], [336, 1008], [1008, 336]]
PINPOINTS_C3 = [[336, 672], [672, 336], [672, 672], [336, 1008], [1008, 336]]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def synth_image(i, W, H):
    return np.random.default_rng(1000 + i).integers(0, 256, (H, W, 3), dtype=np.uint8)


def hf_processor():
    from transformers import CLIPImageProcessor
    return CLIPImageProcessor(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336},
                              image_mean=[0.48145466, 0.4578275, 0.40821073],
                              image_std=[0.26862954, 0.26130258, 0.27577711], resample=3)


def vip_overlays(i, specs):
    """Re-create, with plain PIL calls, the RGBA canvases image_blending draws for the golden_vip cases
    (rectangle with vip_style='constant', mask without segmentation, arrow), alpha=128, 336x336.
    Follows vip_processor/conversation_generator.py:23-34,53-62,75-88,133-141 and
    shape_draw.py:14-71,188-198.  Returns a list of ('rectangle', bbox, width, rgba) / ('layer', array)."""
    import math
    from PIL import Image, ImageDraw
    colors = [(255, 0, 0), (0, 255, 0), (0, 0, 255)]
    out = []
    for inst, shape in enumerate(["rectangle", "mask", "arrow"]):
        bbox = [float(v) for v in specs[inst][:4]]
        seed = int(specs[inst][4])
        random.seed(seed)
        rgba = colors[inst] + (128,)
        canvas = Image.new("RGBA", (336, 336), (0, 0, 0, 0))
        d = ImageDraw.Draw(canvas)
        if shape == "rectangle":
            out.append(("rectangle", bbox, max(int(3 * 336 / 336), 1), rgba))
            continue
        if shape == "mask":
            lw = random.randint(int(0 * 336 / 336), int(2 * 336 / 336))
            seg = [bbox[0], bbox[1], bbox[0], bbox[3], bbox[2], bbox[3], bbox[2], bbox[1]]
            d.polygon([(seg[k], seg[k + 1]) for k in range(0, 8, 2)], outline=None, fill=rgba, width=lw)
        else:  # arrow
            lw = max(random.randint(int(1 * 336 / 336), int(6 * 336 / 336)), 1)
            max_len = max(int(50 * 336 / 336), 1)
            left, top, right, bottom = bbox
            cx, cy = (left + right) / 2, (top + bottom) / 2
            side = min(right - left, bottom - top)
            length = random.uniform(0.8 * side, max_len)
            ang = random.uniform(0, 2 * math.pi)
            cx += random.uniform(-0.25, 0.25) * (right - left)
            cy += random.uniform(-0.25, 0.25) * (bottom - top)
            head = max(random.uniform(0.2, 0.5) * length, int(6 * 336 / 336))
            ex = cx + (length - head) * math.cos(ang)
            ey = cy + (length - head) * math.sin(ang)
            if random.random() < 0.5:
                mx = (cx + ex) / 2 + random.uniform(-5, 5) * int(336 / 336)
                my = (cy + ey) / 2 + random.uniform(-5, 5) * int(336 / 336)
                d.line([(cx, cy), (mx, my), (ex, ey)], fill=rgba, width=lw)
            else:
                d.line([(cx, cy), (ex, ey)], fill=rgba, width=lw)
            pts = [(cx + head * math.cos(ang + math.pi / 3), cy + head * math.sin(ang + math.pi / 3)), (cx, cy),
                   (cx + head * math.cos(ang - math.pi / 3), cy + head * math.sin(ang - math.pi / 3))]
            if random.random() < 0.5:
                d.polygon(pts, fill=rgba)
            else:
                d.line(pts, fill=rgba, width=lw)
        out.append(("layer", np.asarray(canvas).copy()))
    return out


def cos_rows(a, b):
    """row-wise cosine similarity of two [..., D] float arrays."""
    a = a.reshape(-1, a.shape[-1]).astype(np.float64)
    b = b.reshape(-1, b.shape[-1]).astype(np.float64)
    return (a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1) + 1e-30)
