"""Host-side integer/float bookkeeping of the anyres path: resolution selection, grid shape,
Pillow-exact LANCZOS coefficient tables, tile descriptors and unpad crop bounds.

Mirrors (names and argument meaning) vis_zephyr/model/multi_scale_process.py of the reference:
select_best_fit_resolution :29-68, calculate_grid_shape :117-133, unpad_image :188-211.
Only small scalar work lives here; every pixel and feature row is touched on the GPU.
"""
from __future__ import annotations

import ast
import math
from functools import lru_cache
from typing import List, Sequence, Tuple

import numpy as np

TILE = 336
PRECISION_BITS = 32 - 8 - 2  # Pillow Resample.c


def _robust_literal_eval(value_str):
    """multi_scale_process.py:12-26 -- the shipped config stores pinpoints as a quoted string."""
    if not isinstance(value_str, str):
        return value_str
    res = value_str
    while isinstance(res, str):
        try:
            res = ast.literal_eval(res)
        except (ValueError, SyntaxError):
            return res
    return res


def select_best_fit_resolution(original_resolution, possible_resolutions):
    """multi_scale_process.py:29-68: maximise effective area, then minimise waste; first wins."""
    ori_w, ori_h = original_resolution
    best, max_eff, min_waste = None, 0, float("inf")
    for w, h in possible_resolutions:
        scale = min(w / ori_w, h / ori_h)
        dw, dh = int(ori_w * scale), int(ori_h * scale)
        eff = min(dw * dh, ori_w * ori_h)
        waste = (w * h) - eff
        if eff > max_eff or (eff == max_eff and waste < min_waste):
            max_eff, min_waste, best = eff, waste, (w, h)
    return best


def calculate_grid_shape(image_size, grid_pinpoints, patch_size):
    """multi_scale_process.py:117-133 -> (n_w, n_h)."""
    possible = _robust_literal_eval(grid_pinpoints)
    if not isinstance(possible, list):
        raise ValueError(f"grid_pinpoints did not evaluate to a list: {grid_pinpoints}")
    w, h = select_best_fit_resolution(image_size, possible)
    return (w // patch_size, h // patch_size)


def resize_target(original_size, target_res):
    """multi_scale_process.py:81-93: aspect-preserving size and centre paste offset."""
    ow, oh = original_size
    tw, th = target_res
    s = min(tw / ow, th / oh)
    nw, nh = int(ow * s), int(oh * s)
    return (nw, nh), ((tw - nw) // 2, (th - nh) // 2)


def unpad_bounds(current_hw: Tuple[int, int], original_size: Tuple[int, int]):
    """Crop of unpad_image (multi_scale_process.py:188-211) on a [D, H, W] map, AS WRITTEN:
    the reference unpacks `current_w, current_h = image_tensor.shape[1:]`, i.e. it calls the
    row count "w" and the column count "h" (quirk Q4).  Returns (y0, y1, x0, x1) half-open."""
    H, W = current_hw
    original_w, original_h = original_size
    current_w, current_h = H, W  # the swap
    if original_w / original_h > current_w / current_h:
        factor = current_w / original_w
        new_h = int(original_h * factor)
        padding = (current_h - new_h) // 2
        y0, y1 = padding, current_h - padding  # slices dim 1 (rows) with "h" numbers
        return _clip_slice(y0, y1, H) + (0, W)
    factor = current_h / original_h
    new_w = int(original_w * factor)
    padding = (current_w - new_w) // 2
    x0, x1 = padding, current_w - padding      # slices dim 2 (cols) with "w" numbers
    return (0, H) + _clip_slice(x0, x1, W)


def _clip_slice(a: int, b: int, n: int):
    """python slice semantics a:b on a length-n axis."""
    lo, hi, _ = slice(a, b).indices(n)
    return (lo, max(hi, lo))


# ---------------------------------------------------------------------------------------------
# Pillow LANCZOS coefficients (Resample.c precompute_coeffs + normalize_coeffs_8bpc)
# ---------------------------------------------------------------------------------------------
def _sinc(x: float) -> float:
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x: float) -> float:
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def _bicubic(x: float) -> float:
    """Pillow Resample.c bicubic_filter (a = -0.5, support 2)."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


_FILTERS = {"lanczos": (_lanczos, 3.0), "bicubic": (_bicubic, 2.0)}


def lanczos_table(in_size: int, out_size: int) -> np.ndarray:
    return resample_table(in_size, out_size, "lanczos")


@lru_cache(maxsize=512)
def resample_table(in_size: int, out_size: int, filt: str = "lanczos") -> np.ndarray:
    """int32 table [ksize, out_size, xmin[out], count[out], kk[out*ksize]] for one axis and one Pillow
    filter ('lanczos': Image.resize(LANCZOS); 'bicubic': what CLIPImageProcessor.resize asks for).
    in_size == out_size yields the identity table (Image.resize returns a copy in that case,
    and Resample.c skips the pass)."""
    fn, fsupport = _FILTERS[filt]
    if in_size == out_size:
        ksize = 1
        xmin = np.arange(out_size, dtype=np.int32)
        cnt = np.ones(out_size, dtype=np.int32)
        kk = np.full(out_size, 1 << PRECISION_BITS, dtype=np.int32)
        return np.concatenate([np.array([ksize, out_size], np.int32), xmin, cnt, kk])
    scale = float(np.float32(in_size) - np.float32(0)) / out_size
    filterscale = max(scale, 1.0)
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        if lo < 0:
            lo = 0
        hi = int(center + support + 0.5)
        if hi > in_size:
            hi = in_size
        n = hi - lo
        w = [fn((x + lo - center + 0.5) * ss) for x in range(n)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        xmin[xx] = lo
        cnt[xx] = n
    return np.concatenate([np.array([ksize, out_size], np.int32), xmin, cnt, kk.reshape(-1)])


@lru_cache(maxsize=512)
def resample_table_dp(in_size: int, out_size: int, filt: str = "lanczos") -> np.ndarray:
    """The same axis table in the form the dp4a kernels read (vz_preprocess3, layout in include/vz_b200.h):
    every window aligned DOWN to a multiple of four source pixels, the coefficients shifted to match and split
    into byte limbs c = c0 + 256 c1 + 65536 c2 (c0, c1 unsigned, c2 signed), four taps per 32-bit word."""
    t = resample_table(in_size, out_size, filt)
    ks, n = int(t[0]), int(t[1])
    xmin, cnt = t[2:2 + n].astype(np.int64), t[2 + n:2 + 2 * n].astype(np.int64)
    kk = t[2 + 2 * n:].reshape(n, ks).astype(np.int64)
    G = (ks + 6) >> 2
    sh = xmin & 3
    slots = np.zeros((n, 4 * G), np.int64)
    cols = sh[:, None] + np.arange(ks)[None, :]
    live = np.arange(ks)[None, :] < cnt[:, None]
    rows = np.broadcast_to(np.arange(n)[:, None], cols.shape)
    slots[rows[live], cols[live]] = kk[live]
    limbs = np.stack([slots & 255, (slots >> 8) & 255, (slots >> 16) & 255, np.zeros_like(slots)], axis=1)   # [n,4,4G]
    b = limbs.reshape(n, 4, G, 4).astype(np.uint32)
    words = b[..., 0] | (b[..., 1] << 8) | (b[..., 2] << 16) | (b[..., 3] << 24)                             # [n,4,G]
    coef = np.ascontiguousarray(words.transpose(0, 2, 1)).reshape(-1)                                        # [n][G][4]
    n4 = (n + 3) & ~3
    abase = np.zeros(n4, np.int64)
    abase[:n] = xmin & ~3
    ngrp = np.zeros(n4, np.int64)
    ngrp[:n] = (cnt + sh + 3) >> 2
    head = np.array([G, n, 0, 0], np.int64)
    out = np.concatenate([head, abase, ngrp]).astype(np.uint32)
    return np.concatenate([out, coef.astype(np.uint32)]).view(np.int32)


def band_window_groups(in_size: int, out_size: int, filt: str, origin: int, band: int = 14, n_bands: int = 24) -> int:
    """most 4-row groups of the intermediate that the taps of one `band`-row output band touch, for a tile whose
    first row is output row `origin` of the axis (rows outside [0, out_size) read nothing)"""
    t = resample_table(in_size, out_size, filt)
    n = int(t[1])
    xmin, cnt = t[2:2 + n].astype(np.int64), t[2 + n:2 + 2 * n].astype(np.int64)
    lo, hi = xmin >> 2, (xmin + cnt + 3) >> 2
    worst = 1
    for b in range(n_bands):
        a, e = max(origin + b * band, 0), min(origin + (b + 1) * band, n)
        if e > a:
            worst = max(worst, int(hi[a:e].max() - lo[a:e].min()))
    return worst


class TablePool:
    """Packs the axis tables of a batch into one int32 buffer and hands out word offsets (every chunk starts on
    a 16-byte boundary: the dp4a tables are read with 128-bit loads)."""

    def __init__(self):
        self._off = {}
        self._chunks: List[np.ndarray] = []
        self._words = 0
        self.max_ksize = 1
        self.max_groups = 1

    def _add(self, key, t: np.ndarray) -> int:
        pad = (-t.size) % 4
        if pad:
            t = np.concatenate([t, np.zeros(pad, np.int32)])
        self._off[key] = self._words
        self._chunks.append(t)
        self._words += t.size
        return self._off[key]

    def offset_dp(self, in_size: int, out_size: int, filt: str = "lanczos") -> int:
        key = (in_size, out_size, filt, "dp")
        if key not in self._off:
            t = resample_table_dp(in_size, out_size, filt)
            self.max_groups = max(self.max_groups, int(t[0]))
            return self._add(key, t)
        return self._off[key]

    def offset(self, in_size: int, out_size: int, filt: str = "lanczos") -> int:
        key = (in_size, out_size, filt)
        if key not in self._off:
            t = resample_table(in_size, out_size, filt)
            self.max_ksize = max(self.max_ksize, int(t[0]))
            return self._add(key, t)
        return self._off[key]

    def pack(self) -> np.ndarray:
        return np.concatenate(self._chunks) if self._chunks else np.zeros(1, np.int32)


def anyres_views(image_size: Tuple[int, int], grid_pinpoints) -> Tuple[List[dict], Tuple[int, int]]:
    """Views (global + row-major grid tiles) produced by process_any_resolution_image
    (multi_scale_process.py:136-183) for an image of size (W, H).  Each view is a dict with the
    fields of vz_tile_desc minus the table offsets.  Returns (views, best_fit_resolution)."""
    possible = grid_pinpoints if isinstance(grid_pinpoints, list) else _robust_literal_eval(grid_pinpoints)
    best = select_best_fit_resolution(image_size, possible)
    (nw, nh), (px, py) = resize_target(image_size, best)
    views = [dict(out_w=TILE, out_h=TILE, off_x=0, off_y=0, tile_x=0, tile_y=0)]  # squashed global view
    W, H = best
    for i in range(0, H, TILE):
        for j in range(0, W, TILE):
            views.append(dict(out_w=nw, out_h=nh, off_x=px, off_y=py, tile_x=j, tile_y=i))
    return views, best


def single_view(image_size: Tuple[int, int], mode: str = "identity") -> List[dict]:
    """Fixed-336 path for an image that is already 336x336 ('identity'), or 'resize' = LANCZOS squash
    to 336x336 (mm_utils.py:59-63).  See fixed_view for the other modes of mm_utils.process_images."""
    return fixed_view(image_size, mode)[0]


def hf_resize_size(image_size: Tuple[int, int], short: int = TILE) -> Tuple[int, int]:
    """CLIPImageProcessor.resize with size={'shortest_edge': 336}: the short side becomes 336, the long
    side int(336 * long / short) (transformers get_resize_output_image_size, default_to_square=False)."""
    w, h = image_size
    if w <= h:
        return short, int(short * h / w)
    return int(short * w / h), short


def fixed_view(image_size: Tuple[int, int], mode: str = "identity", image_mean=None):
    """One 336x336 view per image, as mm_utils.process_images (mm_utils.py:38-87) and the 'pad' / plain
    branches of the training loader (train/train.py:570-590) produce it.  Returns (views, canvas) where
    canvas = dict(W, H, pad_x, pad_y, bg) is the virtual source the tables address:
      'identity'  336x336 input, nothing to do
      'resize'    LANCZOS squash to 336x336
      'plain'     CLIPImageProcessor.preprocess: BICUBIC resize of the short side to 336, centre crop
      'square'    centre crop to min(w, h) (mm_utils.py:65-74), then 'plain'
      'pad'       expand2square with the mean colour (mm_utils.py:16-35), then 'plain'"""
    w, h = image_size
    canvas = dict(W=w, H=h, pad_x=0, pad_y=0, bg=(0, 0, 0))
    if mode == "identity":
        if (w, h) != (TILE, TILE):
            raise ValueError("identity view needs a 336x336 image")
        return [dict(out_w=TILE, out_h=TILE, off_x=0, off_y=0, tile_x=0, tile_y=0)], canvas
    if mode == "resize":
        return [dict(out_w=TILE, out_h=TILE, off_x=0, off_y=0, tile_x=0, tile_y=0)], canvas
    if mode == "square":
        m = min(w, h)
        left, top = int((w - m) / 2), int((h - m) / 2)
        canvas = dict(W=m, H=m, pad_x=-left, pad_y=-top, bg=(0, 0, 0))
    elif mode == "pad":
        m = max(w, h)
        mean = image_mean if image_mean is not None else (0.48145466, 0.4578275, 0.40821073)
        bg = tuple(int(x * 255) for x in mean)
        canvas = dict(W=m, H=m, pad_x=(m - w) // 2 if h > w else 0, pad_y=(m - h) // 2 if w > h else 0, bg=bg)
    elif mode != "plain":
        raise ValueError(f"unknown mode {mode}")
    nw, nh = hf_resize_size((canvas["W"], canvas["H"]))
    # transformers center_crop: top = (h - 336) // 2, left = (w - 336) // 2
    view = dict(out_w=nw, out_h=nh, off_x=0, off_y=0, tile_x=(nw - TILE) // 2, tile_y=(nh - TILE) // 2, filt="bicubic")
    return [view], canvas


# ---------------------------------------------------------------------------------------------
# merge bookkeeping (vis_zephyr_arch.py:396-473)
# ---------------------------------------------------------------------------------------------
MERGE_FLAT, MERGE_SPATIAL, MERGE_SPATIAL_UNPAD, MERGE_SINGLE_NEWLINE = 0, 1, 2, 3


def slot_descriptor(row_base: int, n_tiles: int, rows_per_tile: int, merge_type: str,
                    image_aspect_ratio: str = "anyres", image_size=None, grid_pinpoints=None,
                    tile_size: int = TILE, side: int | None = None) -> dict:
    """Describe how one image's projector rows [n_tiles, rows_per_tile, D] are merged.
    `side` plays the role of vision_tower.num_patches_per_side (h = w = side)."""
    d = dict(row_base=row_base, n_rows=0, merge=MERGE_FLAT, hw=rows_per_tile, h=0, w=0, n_w=0, n_h=0,
             y0=0, y1=0, x0=0, x1=0)
    if merge_type == "flat":
        d["n_rows"] = n_tiles * rows_per_tile
        return d
    if not merge_type.startswith("spatial"):
        raise ValueError(f"Unknown mm_patch_merge_type: {merge_type}")
    if n_tiles == 1:
        if "unpad" in merge_type:
            d.update(merge=MERGE_SINGLE_NEWLINE, n_rows=rows_per_tile + 1)
        else:
            d["n_rows"] = rows_per_tile
        return d
    if side is None:
        side = int(round(math.sqrt(rows_per_tile)))
    h = w = side
    assert h * w == rows_per_tile  # vis_zephyr_arch.py:424
    if image_aspect_ratio != "anyres":
        raise NotImplementedError  # vis_zephyr_arch.py:434
    n_w, n_h = calculate_grid_shape(image_size, grid_pinpoints, tile_size)
    if n_w * n_h != n_tiles - 1:
        raise RuntimeError(f"grid {n_w}x{n_h} does not match {n_tiles - 1} high-resolution tiles")
    d.update(h=h, w=w, n_w=n_w, n_h=n_h)
    if "unpad" in merge_type:
        y0, y1, x0, x1 = unpad_bounds((n_h * h, n_w * w), image_size)
        d.update(merge=MERGE_SPATIAL_UNPAD, y0=y0, y1=y1, x0=x0, x1=x1,
                 n_rows=rows_per_tile + (y1 - y0) * (x1 - x0 + 1))
    else:
        d.update(merge=MERGE_SPATIAL, n_rows=rows_per_tile + n_h * h * n_w * w)
    return d
