"""GPU image preprocessing: visual-prompt blend + anyres tiling + normalise (+ patchify).

Host-facing mirror of
  process_any_resolution_image   vis_zephyr/model/multi_scale_process.py:136-183
  process_images (fixed 336)     vis_zephyr/model/mm_utils.py:38-87
  image_blending (pixel part)    vis_zephyr/model/vip_processor/conversation_generator.py:13-148
All pixel arithmetic runs in the CUDA kernel `vz_preprocess`; this module only builds the small
descriptor tables (tile geometry, LANCZOS coefficients) and moves them to the device.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from functools import lru_cache
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .anyres import TILE, TablePool, anyres_views, band_window_groups, fixed_view, resample_table, single_view

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def clip_lut(mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD, style: str = "hf-numpy") -> np.ndarray:
    """f32 [3,256] table of CLIPImageProcessor's rescale+normalise on every u8 value.
    'hf-numpy'  : transformers 4.52.4 slow processor (the reference's pin, environment.yaml:108):
                  f32(f64(u8) * (1/255)) then (x - f32 mean) / f32 std in float32.
    'hf-fused'  : torchvision fast path of newer transformers: (f32(u8) - 255 m) / (255 s)."""
    v = np.arange(256, dtype=np.float64)
    out = np.empty((3, 256), np.float32)
    for c in range(3):
        m, s = np.float32(mean[c]), np.float32(std[c])
        if style == "hf-numpy":
            x = (v * (1 / 255)).astype(np.float32)
            out[c] = (x - m) / s
        elif style == "hf-fused":
            mm = np.float32(mean[c]) * np.float32(255.0)
            sd = np.float32(std[c]) * np.float32(255.0)
            out[c] = (v.astype(np.float32) - mm) / sd
        else:
            raise ValueError(style)
    return out


def lut_from_processor(processor) -> np.ndarray:
    """Run ANY HF-style image processor on a 0..255 ramp and read the 768 values back: the LUT is
    then bit-identical to that installation's processor."""
    from PIL import Image
    size = processor.crop_size["height"]
    ramp = np.zeros((size, size, 3), np.uint8)
    ramp[0, :256, :] = np.arange(256, dtype=np.uint8)[:, None]
    px = processor.preprocess(Image.fromarray(ramp), return_tensors="pt")["pixel_values"][0]
    return px[:, 0, :256].contiguous().numpy().astype(np.float32)


@dataclass
class VisualPrompt:
    """One visual-prompt instance applied to an image, in drawing order.
    kind 'layer': a host-rasterised RGBA overlay [H,W,4] (any PIL ImageDraw shape);
    kind 'rectangle': ImageDraw.rectangle(outline=rgba, width) rasterised in the kernel
    (vip_processor/shape_draw.py:68-71)."""
    kind: str
    rgba: Tuple[int, int, int, int] = (0, 0, 0, 0)
    bbox: Optional[Tuple[float, float, float, float]] = None
    width: int = 1
    layer: Optional[np.ndarray] = None


@dataclass
class PreprocessPlan:
    """Device-resident descriptors for one batch (built once per batch on the host)."""
    n_tiles: int
    tiles_per_image: List[int]
    image_sizes: List[Tuple[int, int]]
    images_dev: torch.Tensor
    prims_dev: Optional[torch.Tensor]
    tiles_dev: torch.Tensor
    tables_dev: torch.Tensor
    lut_dev: torch.Tensor
    n_images: int
    n_prims: int
    max_src_w: int
    max_ksize: int
    keep: list = field(default_factory=list)  # tensors that must outlive the launch
    h2d_bytes: int = 0
    algorithmic_bytes: int = 0
    # two-kernel form (vz_preprocess2): one horizontal view per distinct (image, horizontal table)
    hviews_dev: Optional[torch.Tensor] = None
    n_hviews: int = 0
    scratch_pixels: int = 0
    max_span_px: int = 1
    max_span128_px: int = 1
    max_groups: int = 1          # largest G of the dp4a tables
    max_band_groups: int = 1     # most 4-row groups a 14-row output band's window spans
    all_identity: bool = False   # every tile = the whole of a 336 x 336 image (vz_preprocess_identity)
    max_rows: int = 1
    max_out_w: int = 1
    scratch: Optional[torch.Tensor] = None


def _struct_array_to_dev(arr, device) -> torch.Tensor:
    raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
    return _lib.h2d(raw, device)


def build_plan(images: Sequence[torch.Tensor], views: Sequence[List[dict]], lut: np.ndarray,
               prompts: Optional[Sequence[Sequence[VisualPrompt]]] = None,
               canvases: Optional[Sequence[Optional[dict]]] = None) -> PreprocessPlan:
    """images: u8 CUDA tensors [H,W,3]; views[i]: list of view dicts (see anyres.anyres_views);
    canvases[i] (optional): dict(W, H, pad_x, pad_y, bg) = the virtual source the views' tables address
    (anyres.fixed_view: expand2square padding / centre crop); default = the image itself."""
    device = images[0].device
    pool = TablePool()
    n_img = len(images)
    img_arr = (_lib.ImageDesc * n_img)()
    prim_list, tile_list, keep = [], [], []
    tiles_per_image, sizes = [], []
    max_w, algo = 1, 0
    hview_index, hview_list, scratch_px, max_span, max_rows, max_out_w, max_span128 = {}, [], 0, 1, 1, 1, 1
    all_identity = True
    max_band_groups = 1
    for i, im in enumerate(images):
        if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3 or not im.is_cuda:
            raise ValueError("images must be uint8 CUDA tensors of shape [H, W, 3]")
        im = im.contiguous()
        keep.append(im)
        H, W = int(im.shape[0]), int(im.shape[1])
        sizes.append((W, H))
        max_w = max(max_w, W)
        algo += 3 * W * H
        layers = []
        begin = len(prim_list)
        for p in (prompts[i] if prompts is not None else ()):
            pr = _lib.Prim()
            if p.kind == "layer":
                lay = np.ascontiguousarray(p.layer, dtype=np.uint8)
                if lay.shape != (H, W, 4):
                    raise ValueError("overlay layer must be [H, W, 4] uint8")
                pr.type, pr.layer = _lib.PRIM_LAYER, len(layers)
                layers.append(lay)
                algo += 4 * W * H
            elif p.kind == "rectangle":
                if int(p.width) == 0:
                    continue  # ImageDraw.rectangle draws nothing for width == 0
                x0, y0, x1, y1 = (int(v) for v in p.bbox)  # PIL truncates the float corners
                if x1 < x0 or y1 < y0:
                    raise ValueError("x1 must be greater than or equal to x0 (PIL semantics)")
                pr.type = _lib.PRIM_RECT
                pr.x0, pr.y0, pr.x1, pr.y1, pr.width = x0, y0, x1, y1, int(p.width)
                r, g, b, a = p.rgba
                pr.rgba = (r & 255) | ((g & 255) << 8) | ((b & 255) << 16) | ((a & 255) << 24)
            else:
                raise ValueError(f"unsupported visual prompt kind {p.kind!r}")
            prim_list.append(pr)
        lay_dev = None
        if layers:
            lay_dev = _lib.h2d(np.stack(layers), device)
            keep.append(lay_dev)
        d = img_arr[i]
        d.src = im.data_ptr()
        d.layers = lay_dev.data_ptr() if lay_dev is not None else None
        d.W, d.H = W, H
        cv = canvases[i] if canvases is not None and canvases[i] is not None else None
        cW, cH = (cv["W"], cv["H"]) if cv else (W, H)
        if cv:
            d.pad_x, d.pad_y = int(cv["pad_x"]), int(cv["pad_y"])
            r_, g_, b_ = cv.get("bg", (0, 0, 0))
            d.bg = (r_ & 255) | ((g_ & 255) << 8) | ((b_ & 255) << 16)
        max_w = max(max_w, cW)
        d.prim_begin, d.prim_count = begin, len(prim_list) - begin
        if d.prim_count > 32:
            raise ValueError("at most 32 visual-prompt instances per image")
        tiles_per_image.append(len(views[i]))
        plain_canvas = cv is None or (cv["W"], cv["H"], int(cv["pad_x"]), int(cv["pad_y"])) == (W, H, 0, 0)
        if (W, H) != (TILE, TILE) or not plain_canvas or im.data_ptr() % 4 or (lay_dev is not None and lay_dev.data_ptr() % 16):
            all_identity = False
        for v in views[i]:
            if (v["out_w"], v["out_h"], v["off_x"], v["off_y"], v["tile_x"], v["tile_y"]) != (TILE, TILE, 0, 0, 0, 0):
                all_identity = False
            t = _lib.TileDesc()
            t.image = i
            t.out_w, t.out_h = v["out_w"], v["out_h"]
            t.off_x, t.off_y = v["off_x"], v["off_y"]
            t.tile_x, t.tile_y = v["tile_x"], v["tile_y"]
            filt = v.get("filt", "lanczos")
            t.tab_h = pool.offset(cW, v["out_w"], filt)
            t.tab_v = pool.offset(cH, v["out_h"], filt)
            t.tab_v_dp = pool.offset_dp(cH, v["out_h"], filt)
            max_band_groups = max(max_band_groups, _band_groups(cH, v["out_h"], filt, v["tile_y"] - v["off_y"]))
            hk = (i, t.tab_h)
            if hk not in hview_index:
                hview_index[hk] = len(hview_list)
                hvd = _lib.HViewDesc()
                hvd.image, hvd.tab_h, hvd.out_w, hvd.rows, hvd.offset = i, t.tab_h, v["out_w"], cH, scratch_px
                hvd.tab_h_dp = pool.offset_dp(cW, v["out_w"], filt)
                hview_list.append(hvd)
                # room for both intermediate layouts: RGBX [rows][out_w] words (vz_preprocess2) and
                # [ceil(rows / 4)][out_w] uint4 (vz_preprocess3); every view starts on a 16-byte boundary
                scratch_px += (4 * ((cH + 3) // 4) * v["out_w"] + 3) // 4 * 4
                max_span = max(max_span, _max_span(cW, v["out_w"], filt))
                max_span128 = max(max_span128, _max_span(cW, v["out_w"], filt, 128))
                max_rows, max_out_w = max(max_rows, cH), max(max_out_w, v["out_w"])
            t.hview = hview_index[hk]
            tile_list.append(t)
    n_tiles = len(tile_list)
    tile_arr = (_lib.TileDesc * n_tiles)(*tile_list)
    prim_arr = (_lib.Prim * max(1, len(prim_list)))(*prim_list)
    tables = pool.pack()
    plan = PreprocessPlan(
        n_tiles=n_tiles, tiles_per_image=tiles_per_image, image_sizes=sizes,
        images_dev=_struct_array_to_dev(img_arr, device),
        prims_dev=_struct_array_to_dev(prim_arr, device) if prim_list else None,
        tiles_dev=_struct_array_to_dev(tile_arr, device),
        tables_dev=_lib.h2d(tables, device),
        lut_dev=_lib.h2d(np.ascontiguousarray(lut, np.float32), device),
        n_images=n_img, n_prims=len(prim_list), max_src_w=max_w, max_ksize=pool.max_ksize, keep=keep)
    hview_arr = (_lib.HViewDesc * len(hview_list))(*hview_list)
    plan.hviews_dev = _struct_array_to_dev(hview_arr, device)
    plan.n_hviews, plan.scratch_pixels = len(hview_list), scratch_px
    plan.max_span_px, plan.max_rows, plan.max_out_w = max_span, max_rows, max_out_w
    plan.max_span128_px = max_span128
    plan.all_identity = all_identity
    plan.max_groups, plan.max_band_groups = pool.max_groups, max_band_groups
    plan.h2d_bytes = (plan.images_dev.numel() + plan.tiles_dev.numel() + plan.tables_dev.numel() * 4 + plan.hviews_dev.numel() +
                      (plan.prims_dev.numel() if plan.prims_dev is not None else 0) + 768 * 4)
    plan.algorithmic_bytes = algo
    return plan


DP_MAX_BAND_GROUPS = 36      # 36 * 336 * 16 B = 194 KB: what the vertical dp4a pass can stage per band


@lru_cache(maxsize=4096)
def _band_groups(in_size: int, out_size: int, filt: str, origin: int) -> int:
    return band_window_groups(in_size, out_size, filt, origin)


@lru_cache(maxsize=512)
def _max_span(in_size: int, out_size: int, filt: str, block: int = 256) -> int:
    """widest source window (in pixels) that `block` consecutive output columns of an axis table read"""
    t = resample_table(in_size, out_size, filt)
    n = int(t[1])
    xmin, cnt = t[2:2 + n], t[2 + n:2 + 2 * n]
    span = 1
    for x0 in range(0, n, block):
        xl = min(x0 + block - 1, n - 1)
        span = max(span, int(xmin[xl] + cnt[xl] - xmin[x0]))
    return span


def run_plan(plan: PreprocessPlan, out_mode: str = "patches", out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Launch the preprocess kernels.  out_mode 'patches' -> bf16 [T*576, 592]; 'chw' -> f32 [T,3,336,336].
    Plans that resample anything take the dp4a form (vz_preprocess3: horizontal pass once per source row into a
    planar u8 intermediate, then vertical pass + normalise + patchify, four taps per instruction); VZ_PRE_FORM=two
    selects the older one-multiply-per-tap pair (vz_preprocess2), kept as a cross-check and for ksize > 58; plans
    made of identity views only (fixed-336 inputs) and VZ_PRE_FORM=fused / VZ_PRE_FUSED=1 take the fused kernel."""
    lib = _lib.load()
    dev = plan.tiles_dev.device
    T = plan.n_tiles
    if out_mode == "patches":
        if out is None:
            out = torch.empty((T * 576, 592), dtype=torch.bfloat16, device=dev)
        mode = _lib.OUT_PATCHES_BF16
    elif out_mode == "chw":
        if out is None:
            out = torch.empty((T, 3, TILE, TILE), dtype=torch.float32, device=dev)
        mode = _lib.OUT_CHW_F32
    else:
        raise ValueError(out_mode)
    form = os.environ.get("VZ_PRE_FORM", "fused" if os.environ.get("VZ_PRE_FUSED") == "1" else "dp")
    if plan.max_ksize > 1 and form != "fused":
        if plan.scratch is None or plan.scratch.numel() < plan.scratch_pixels:
            plan.scratch = torch.empty(plan.scratch_pixels, dtype=torch.int32, device=dev)
        if form == "dp" and plan.max_groups <= 16 and plan.max_band_groups <= DP_MAX_BAND_GROUPS:
            st = lib.vz_preprocess3(_lib.ptr(plan.images_dev), plan.n_images, _lib.ptr(plan.prims_dev), plan.n_prims,
                                    _lib.ptr(plan.hviews_dev), plan.n_hviews, _lib.ptr(plan.tiles_dev), T,
                                    _lib.ptr(plan.tables_dev), _lib.ptr(plan.lut_dev), mode, _lib.ptr(out),
                                    _lib.ptr(plan.scratch), plan.scratch_pixels, plan.max_span128_px, plan.max_rows,
                                    plan.max_out_w, plan.max_groups, plan.max_band_groups, _lib.stream_ptr())
            _lib.check(st, "vz_preprocess3")
            return out
        st = lib.vz_preprocess2(_lib.ptr(plan.images_dev), plan.n_images, _lib.ptr(plan.prims_dev), plan.n_prims,
                                _lib.ptr(plan.hviews_dev), plan.n_hviews, _lib.ptr(plan.tiles_dev), T,
                                _lib.ptr(plan.tables_dev), _lib.ptr(plan.lut_dev), mode, _lib.ptr(out),
                                _lib.ptr(plan.scratch), plan.scratch_pixels, plan.max_span_px, plan.max_rows,
                                plan.max_out_w, plan.max_ksize, _lib.stream_ptr())
        _lib.check(st, "vz_preprocess2")
        return out
    if plan.all_identity and form != "fused":
        st = lib.vz_preprocess_identity(_lib.ptr(plan.images_dev), plan.n_images, _lib.ptr(plan.prims_dev), plan.n_prims,
                                        _lib.ptr(plan.tiles_dev), T, _lib.ptr(plan.lut_dev), mode, _lib.ptr(out),
                                        _lib.stream_ptr())
        _lib.check(st, "vz_preprocess_identity")
        return out
    st = lib.vz_preprocess(_lib.ptr(plan.images_dev), plan.n_images, _lib.ptr(plan.prims_dev), plan.n_prims,
                           _lib.ptr(plan.tiles_dev), T, _lib.ptr(plan.tables_dev), _lib.ptr(plan.lut_dev),
                           mode, _lib.ptr(out), plan.max_src_w, plan.max_ksize, _lib.stream_ptr())
    _lib.check(st, "vz_preprocess")
    return out


class PatchBatch:
    """bf16 im2col patch rows [T*576, 592] produced by the fused preprocess kernel; accepted by
    CLIPVisionTowerB200.forward in place of reference-style pixel tensors."""

    def __init__(self, patches: torch.Tensor, tiles_per_image: List[int], image_sizes: List[Tuple[int, int]]):
        self.patches = patches
        self.tiles_per_image = list(tiles_per_image)
        self.image_sizes = list(image_sizes)

    @property
    def n_tiles(self) -> int:
        return self.patches.shape[0] // 576

    @property
    def device(self):
        return self.patches.device

    @property
    def dtype(self):
        return self.patches.dtype


def process_any_resolution_images(images: Sequence[torch.Tensor], grid_pinpoints, lut: np.ndarray,
                                  prompts=None, out_mode: str = "patches"):
    """Batch version of process_any_resolution_image (multi_scale_process.py:136-183) on u8 CUDA
    images [H,W,3].  Returns PatchBatch (out_mode='patches') or a list of f32 [T_i,3,336,336]
    tensors in the reference layout (out_mode='chw').  Items of `images` may be PromptedImage (the result of
    visual_prompts.image_blending): their instances are composited first, in order."""
    from .visual_prompts import split_prompted
    images, prompts = split_prompted(images, prompts)
    views = []
    for im in images:
        v, _ = anyres_views((int(im.shape[1]), int(im.shape[0])), grid_pinpoints)
        views.append(v)
    plan = build_plan(images, views, lut, prompts)
    out = run_plan(plan, out_mode)
    if out_mode == "patches":
        return PatchBatch(out, plan.tiles_per_image, plan.image_sizes)
    return list(torch.split(out, plan.tiles_per_image, dim=0))


def process_fixed_images(images: Sequence[torch.Tensor], lut: np.ndarray, prompts=None,
                         out_mode: str = "patches", mode: str = "identity", image_mean=None):
    """Fixed-336 path: blend visual prompts onto the images, bring them to 336x336 the way
    mm_utils.process_images / the training loader do (mode: 'identity' | 'resize' | 'plain' | 'square' |
    'pad', see anyres.fixed_view), normalise, patchify.  Config 2 = 'identity' with prompts.  Items of `images`
    may be PromptedImage (visual_prompts.image_blending)."""
    from .visual_prompts import split_prompted
    images, prompts = split_prompted(images, prompts)
    vc = [fixed_view((int(im.shape[1]), int(im.shape[0])), mode, image_mean) for im in images]
    views, canvases = [v for v, _ in vc], [c for _, c in vc]
    plan = build_plan(images, views, lut, prompts, canvases)
    out = run_plan(plan, out_mode)
    if out_mode == "patches":
        return PatchBatch(out, plan.tiles_per_image, plan.image_sizes)
    return list(torch.split(out, plan.tiles_per_image, dim=0))
