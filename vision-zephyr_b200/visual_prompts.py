"""Visual prompts: resolve a (shape, bbox, segmentation) instance into what the preprocess kernel blends.

Host-facing mirror of
  image_blending        vis_zephyr/model/vip_processor/conversation_generator.py:13-148
  draw_*                vis_zephyr/model/vip_processor/shape_draw.py:14-215
The reference draws every instance with PIL ImageDraw on a transparent RGBA canvas and alpha-composites it
onto the image, one instance after the other.  Here the GEOMETRY stays on the host (it consumes Python's
`random` / `numpy.random` state in the reference's order, so a seeded data loader draws the same shapes) and
the PIXELS go to the GPU: a rectangle becomes a kernel-rasterised primitive, every other shape a host-drawn
RGBA layer (drawn by Pillow itself, so its raster is Pillow's by construction); `vz_preprocess*` composites
them in drawing order with Pillow's exact integer blend, fused with resize / normalise / patchify.

    img = image_blending(img_u8_cuda, shape="ellipse", bbox_coor=b, rgb_color=(255, 0, 0))     # PromptedImage
    img = image_blending(img, shape="arrow", bbox_coor=b2, rgb_color=(0, 255, 0))              # compounds
    patches = process_fixed_images([img], lut)                                                 # one kernel

Without `shapely` (absent from this image, as from the survey's container) polygon queries use the small
even-odd / bounding-box helpers below; with segmentation=None no polygon query is ever made, which is the
regime the golden vectors pin (tests/golden/golden_vip_shapes.npz).
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .preprocess import VisualPrompt

# (lo, hi) multipliers of max(w, h) / anchor for the random line width of each outlined shape
_WIDTH_RANGE = {"rectangle": (2, 8), "ellipse": (2, 8), "arrow": (1, 6), "triangle": (2, 8), "scribble": (2, 12),
                "mask contour": (1, 2), "mask": (0, 2)}
SHAPES = ("rectangle", "ellipse", "arrow", "triangle", "point", "scribble", "mask contour", "mask")


@dataclass
class PromptedImage:
    """A u8 image [H,W,3] (CUDA tensor or array) plus the instances blended onto it so far, in order."""
    image: object
    prompts: List[VisualPrompt] = field(default_factory=list)

    @property
    def size(self) -> Tuple[int, int]:
        return int(self.image.shape[1]), int(self.image.shape[0])


class _Region:
    """What the reference asks of shapely's unary_union(polygons): bounds and contains (even-odd)."""

    def __init__(self, rings: Sequence[Sequence[Tuple[float, float]]]):
        self.rings = [list(r) for r in rings]
        xs = [x for r in self.rings for x, _ in r]
        ys = [y for r in self.rings for _, y in r]
        self.bounds = (min(xs), min(ys), max(xs), max(ys))

    def contains(self, x: float, y: float) -> bool:
        for ring in self.rings:
            inside, n = False, len(ring)
            for i in range(n):
                (x0, y0), (x1, y1) = ring[i], ring[(i + 1) % n]
                if (y0 > y) != (y1 > y) and x < (x1 - x0) * (y - y0) / (y1 - y0) + x0:
                    inside = not inside
            if inside:
                return True
        return False


def _scaled(v: float, size: int, anchor: int) -> int:
    return int(v * size / anchor)


def _random_width(shape: str, size: int, anchor: int, floor: bool = True) -> int:
    lo, hi = _WIDTH_RANGE[shape]
    w = random.randint(_scaled(lo, size, anchor), _scaled(hi, size, anchor))
    return max(w, 1) if floor else w


def _point_in(region: Optional[_Region], bbox):
    """shape_draw.py:224-247: uniform in the bbox, or rejection-sampled inside the polygon (50 tries)."""
    if region is None:
        left, top, right, bottom = bbox
        return np.random.uniform(left, right), np.random.uniform(top, bottom)
    x0, y0, x1, y1 = region.bounds
    for _ in range(50):
        x, y = np.random.uniform(x0, x1), np.random.uniform(y0, y1)
        if region.contains(x, y):
            return x, y
    return np.random.uniform(x0, x1), np.random.uniform(y0, y1)


def _widest_angle_ok(pts) -> bool:
    """shape_draw.py:249-266: every corner angle (cosine rule, degrees) at most 150."""
    for i in range(3):
        p1, p2, p3 = (np.array(pts[(i + k) % 3]) for k in range(3))
        a, b, c = np.linalg.norm(p3 - p2), np.linalg.norm(p1 - p3), np.linalg.norm(p1 - p2)
        if np.degrees(np.arccos((a ** 2 + c ** 2 - b ** 2) / (2 * a * c))) > 150:
            return False
    return True


def resolve_visual_prompt(image_size: Tuple[int, int], shape: str = "rectangle", bbox_coor=None, segmentation=None,
                          image_size_anchor: int = 336, rgb_color=None, vip_style=None, alpha=None,
                          width=None) -> VisualPrompt:
    """One instance of image_blending (conversation_generator.py:13-148) as a VisualPrompt for the kernel.
    Random numbers are drawn in the reference's order: alpha, the polygon choice, the line width, then the
    shape's own geometry."""
    from PIL import Image, ImageDraw
    img_w, img_h = image_size
    size = max(img_w, img_h)
    if alpha is None:
        alpha = random.randint(96, 255) if shape != "mask" else random.randint(48, 128)
    rgba = tuple(rgb_color) + (alpha,)
    region = None
    if segmentation is not None:
        rings = [[(s[i], s[i + 1]) for i in range(0, len(s), 2)] for s in segmentation]
        random.choice(rings)                       # the reference picks one polygon here and never uses it
        region = _Region(rings)

    def line_width(kind, floor=True):
        # the random width is drawn even when `width` overrides it (the reference evaluates it first)
        w = _random_width(kind, size, image_size_anchor, floor)
        return max(_scaled(width, size, image_size_anchor), 1) if width is not None else w

    if shape == "rectangle":
        lw = max(_scaled(3, size, image_size_anchor), 1) if vip_style == "constant" else _random_width("rectangle", size, image_size_anchor)
        if width is not None:
            lw = max(_scaled(width, size, image_size_anchor), 1)
        return VisualPrompt("rectangle", rgba=rgba, bbox=tuple(float(v) for v in bbox_coor), width=lw)

    canvas = Image.new("RGBA", (img_w, img_h), (0, 0, 0, 0))
    draw = ImageDraw.Draw(canvas)
    if shape == "ellipse":
        lw = line_width("ellipse")
        ratio = random.uniform(1, 1.5)
        x0, y0, x1, y1 = region.bounds if region is not None else bbox_coor
        cx, cy, ew, eh = (x0 + x1) / 2, (y0 + y1) / 2, (x1 - x0) * ratio, (y1 - y0) * ratio
        draw.ellipse([cx - ew / 2, cy - eh / 2, cx + ew / 2, cy + eh / 2], outline=rgba, width=lw)
    elif shape == "arrow":
        lw = line_width("arrow")
        _arrow(draw, bbox_coor, rgba, lw, max(_scaled(50, size, image_size_anchor), 1), size, image_size_anchor)
    elif shape == "triangle":
        lw = line_width("triangle")
        while True:
            pts = [_point_in(region, bbox_coor) for _ in range(3)]
            if _widest_angle_ok(pts):
                break
        draw.line([pts[0], pts[1], pts[2], pts[0]], fill=rgba, width=lw, joint="curve")
    elif shape == "point":
        radius = (max(_scaled(8, size, image_size_anchor), 1) if vip_style == "constant"
                  else max(random.randint(_scaled(5, size, image_size_anchor), _scaled(20, size, image_size_anchor)), 1))
        aspect = 1 if random.random() < 0.5 or vip_style == "constant" else random.uniform(0.5, 2.0)
        _point(draw, bbox_coor, region, rgba, radius, aspect)
    elif shape == "scribble":
        lw = line_width("scribble")
        p = [_point_in(region, bbox_coor) for _ in range(4)]
        prev = None
        for t in np.linspace(0, 1, int(1000 * size / image_size_anchor)):     # cubic Bezier, one segment per step
            u = 1 - t
            cur = tuple(u ** 3 * p[0][k] + 3 * u ** 2 * t * p[1][k] + 3 * u * t ** 2 * p[2][k] + t ** 3 * p[3][k]
                        for k in (0, 1))
            if prev:
                draw.line([prev, cur], fill=rgba, width=lw)
            prev = cur
    elif shape in ("mask contour", "mask"):
        lw = line_width(shape, floor=(shape == "mask contour"))
        rings = segmentation
        if rings is None:                          # the bbox as a 4-gon: (x0,y0) (x0,y1) (x1,y1) (x1,y0)
            x0, y0, x1, y1 = bbox_coor
            rings = [[x0, y0, x0, y1, x1, y1, x1, y0]]
        for s in rings:
            pts = [(s[i], s[i + 1]) for i in range(0, len(s), 2)]
            if shape == "mask":
                draw.polygon(pts, outline=None, fill=rgba, width=lw)
            else:                                  # the outline repeated on a (2 lw + 1)^2 grid of offsets
                for dx in range(-lw, lw + 1):
                    for dy in range(-lw, lw + 1):
                        draw.polygon([(x + dx, y + dy) for x, y in pts], outline=rgba)
    else:
        raise ValueError(f"unknown visual prompt shape {shape!r}")
    return VisualPrompt("layer", layer=np.asarray(canvas).copy())


def _arrow(draw, bbox, rgba, lw, max_len, size, anchor):
    """shape_draw.py:14-65: shaft from a jittered centre along a random angle (optionally through a wobbling
    midpoint), head at the centre as a filled or an open 'V' of +-60 degrees."""
    left, top, right, bottom = bbox
    cx, cy = (left + right) / 2, (top + bottom) / 2
    length = random.uniform(0.8 * min(right - left, bottom - top), max_len)
    ang = random.uniform(0, 2 * math.pi)
    cx += random.uniform(-0.25, 0.25) * (right - left)
    cy += random.uniform(-0.25, 0.25) * (bottom - top)
    head = max(random.uniform(0.2, 0.5) * length, int(6 * size / anchor))
    ex, ey = cx + (length - head) * math.cos(ang), cy + (length - head) * math.sin(ang)
    shaft = [(cx, cy), (ex, ey)]
    if random.random() < 0.5:
        k = int(size / anchor)
        shaft.insert(1, ((cx + ex) / 2 + random.uniform(-5, 5) * k, (cy + ey) / 2 + random.uniform(-5, 5) * k))
    draw.line(shaft, fill=rgba, width=lw)
    wings = [(cx + head * math.cos(ang + s * math.pi / 3), cy + head * math.sin(ang + s * math.pi / 3)) for s in (1, -1)]
    vee = [wings[0], (cx, cy), wings[1]]
    if random.random() < 0.5:
        draw.polygon(vee, fill=rgba)
    else:
        draw.line(vee, fill=rgba, width=lw)


def _point(draw, bbox, region, rgba, radius, aspect):
    """shape_draw.py:100-138: a filled ellipse at a Gaussian sample around the box centre (up to 10 tries to
    land inside the polygon).  Like the reference, it needs a segmentation: with none, `mask_polygon.contains`
    fails on None there, and so does this."""
    from scipy.stats import multivariate_normal
    x0, y0, x1, y1 = region.bounds if region is not None else bbox
    mean, cov = [(x1 + x0) / 2, (y1 + y0) / 2], [[(x1 - x0) / 8, 0], [0, (y1 - y0) / 8]]
    tries = 0
    while True:
        px, py = multivariate_normal.rvs(mean=mean, cov=cov)
        if region is None:
            raise AttributeError("'NoneType' object has no attribute 'contains'")
        if region.contains(px, py):
            break
        tries += 1
        if tries >= 10:
            px, py = multivariate_normal.rvs(mean=mean, cov=cov)
            break
    rx, ry = radius * aspect, radius / aspect
    draw.ellipse([px - rx, py - ry, px + rx, py + ry], fill=rgba, outline=rgba)


def image_blending(image, shape="rectangle", bbox_coor=None, segmentation=None, image_size_anchor=336, rgb_color=None,
                   vip_style=None, alpha=None, width=None) -> PromptedImage:
    """Same arguments as the reference's image_blending; instead of a composited PIL image it returns the image
    together with the (now one longer) list of instances, which process_fixed_images /
    process_any_resolution_images composite on the GPU.  Feeding the result back in compounds instances in
    order, like the reference's loop (vip_processor/processor.py:58-73)."""
    pi = image if isinstance(image, PromptedImage) else PromptedImage(image)
    vp = resolve_visual_prompt(pi.size, shape, bbox_coor, segmentation, image_size_anchor, rgb_color, vip_style, alpha, width)
    return PromptedImage(pi.image, pi.prompts + [vp])


def split_prompted(images, prompts=None):
    """[PromptedImage | tensor] (+ optional explicit prompt lists) -> (tensors, prompt lists or None)."""
    if not any(isinstance(x, PromptedImage) for x in images):
        return list(images), prompts
    out_i, out_p = [], []
    for k, x in enumerate(images):
        extra = list(prompts[k]) if prompts is not None else []
        if isinstance(x, PromptedImage):
            out_i.append(x.image)
            out_p.append(list(x.prompts) + extra)
        else:
            out_i.append(x)
            out_p.append(extra)
    return out_i, out_p
