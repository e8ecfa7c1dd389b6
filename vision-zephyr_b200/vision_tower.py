"""CLIP ViT-L/14-336 vision tower with multi-layer feature fusion, running on the B200 kernels.

Drop-in for the reference's
  build_vision_tower            vis_zephyr/model/vision_encoder/builder.py:8-24
  CLIPVisionTower               vis_zephyr/model/vision_encoder/vision_encoder.py:13-151
  DenseChannelIntegrationFusion vis_zephyr/model/gating_fusion/gating_fusion.py:22-50
Same constructor arguments, same attributes/properties, same forward() input forms (list, 4-D,
3-D tensors) plus PatchBatch from the fused preprocess kernel.  The arithmetic is
`vz_vit_forward` (csrc/vz_model.cu); there is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import itertools
import os
import threading
from collections import OrderedDict
from types import SimpleNamespace
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib
from .preprocess import PatchBatch

WIDTH, LAYERS, TOKENS, PATCHES, MLP, PATCH_K = 1024, 24, 577, 576, 4096, 592


def default_clip_config():
    """CLIP ViT-L/14-336 geometry (what CLIPVisionConfig.from_pretrained would return)."""
    return SimpleNamespace(hidden_size=WIDTH, intermediate_size=MLP, num_hidden_layers=LAYERS,
                           num_attention_heads=16, image_size=336, patch_size=14, projection_dim=768,
                           hidden_act="quick_gelu", layer_norm_eps=1e-5)


class Workspace:
    """Grow-only device scratch of one module, ONE BUFFER PER (device, CUDA stream).

    The C ABI is re-entrant (every call works inside the workspace it is handed, on the stream it is
    handed); what the Python layer has to guarantee is that two concurrent calls never share a workspace.
    Activations, LayerNorm statistics and the stream-K hand-over slots all live in it, so it is keyed by the
    caller's current stream: two threads driving one model on two streams (the reference's
    serve/api.py:161-177 generate-in-a-thread case) get disjoint buffers.  A buffer is allocated while its
    stream is current, so PyTorch's caching allocator orders its reuse after the kernels of that stream when
    it is replaced by a larger one."""

    capture_sink: Optional[list] = None     # set by GraphCache while a CUDA graph is being captured

    def __init__(self):
        self._bufs: Dict[tuple, torch.Tensor] = {}
        self._lock = threading.Lock()

    def get(self, nbytes: int, device) -> torch.Tensor:
        device = torch.device(device)
        if torch.cuda.is_current_stream_capturing():
            # a captured graph owns its scratch: a fresh buffer from the graph's private pool, kept alive by the
            # graph's cache entry, never shared with eager calls or with other graphs
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            if Workspace.capture_sink is not None:
                Workspace.capture_sink.append(buf)
            return buf
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device(),
               torch.cuda.current_stream(device).cuda_stream)
        with self._lock:
            buf = self._bufs.get(key)
            if buf is None or buf.numel() < nbytes:
                self._bufs.pop(key, None)
                buf = None
                buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
                self._bufs[key] = buf
        return buf

    @property
    def buf(self):
        """the current stream's buffer (tests / tools)"""
        for k, v in self._bufs.items():
            if k[2] == torch.cuda.current_stream().cuda_stream:
                return v
        return None


GRAPH_MAX_TILES = int(os.environ.get("VZ_GRAPH_MAX_TILES", "8"))
_capture_lock = threading.Lock()
PACK_GENERATION = itertools.count(1)   # process-wide: identifies one packing of one module's weights


class GraphCache:
    """CUDA graphs of a module's kernel sequence for SMALL batches (the serving shape: a single image is ~230
    launches whose kernels last 5-25 us each, so launch gaps are ~10 % of the call).  One entry per (shape key,
    device, stream): static input / output buffers + the captured graph; least recently used entries are dropped.
    Replays copy the inputs into the static buffers and return the static output, which stays valid until the
    next replay of the same entry (callers inside this package consume it at once, in stream order)."""

    def __init__(self, capacity: int = 8):
        self.capacity = capacity
        self._entries: "OrderedDict[tuple, dict]" = OrderedDict()
        self._lock = threading.Lock()

    def clear(self) -> None:
        """drop every captured graph (the owner's weights were re-packed: the graphs hold their old addresses)"""
        with self._lock:
            self._entries.clear()

    @staticmethod
    def usable(n_tiles: int) -> bool:
        return (0 < n_tiles <= GRAPH_MAX_TILES and os.environ.get("VZ_GRAPHS", "1") != "0"
                and not torch.cuda.is_current_stream_capturing())

    def run(self, key: tuple, inputs: List[torch.Tensor], fn):
        """fn(*static_inputs) -> output tensor, made of stream-ordered launches only (no sync, no host reads)."""
        dev = inputs[0].device
        key = key + (dev.index, torch.cuda.current_stream(dev).cuda_stream) + tuple((tuple(t.shape), t.dtype) for t in inputs)
        with self._lock:
            ent = self._entries.get(key)
            if ent is not None:
                self._entries.move_to_end(key)
        if ent is None:
            static = [torch.empty_like(t) for t in inputs]
            for s_, t in zip(static, inputs):
                s_.copy_(t)
            fn(*static)                                   # warm-up: first-use initialisation happens outside the capture
            graph = torch.cuda.CUDAGraph()
            with _capture_lock:
                Workspace.capture_sink = keep = []
                try:
                    with torch.cuda.graph(graph):
                        out = fn(*static)
                finally:
                    Workspace.capture_sink = None
            ent = dict(static=static, graph=graph, out=out, keep=keep)
            with self._lock:
                self._entries[key] = ent
                while len(self._entries) > self.capacity:
                    self._entries.popitem(last=False)
        for s_, t in zip(ent["static"], inputs):
            if s_.data_ptr() != t.data_ptr():
                s_.copy_(t)
        ent["graph"].replay()
        return ent["out"]


class CLIPVisionTowerB200(nn.Module):
    def __init__(self, vision_tower_path, args, delay_load: bool = False):
        super().__init__()
        self.is_loaded = False
        self.vision_tower_path = ("openai/clip-vit-large-patch14-336" if vision_tower_path is None
                                  else vision_tower_path)
        self.select_feature = getattr(args, "mm_vision_select_feature", "patch")
        raw = getattr(args, "mm_vision_select_layer", None)
        if isinstance(raw, str):
            try:
                self.select_layers = [int(x.strip()) for x in raw.split(",")]
            except ValueError:
                raise ValueError("Invalid format for mm_vision_select_layer. Expected a comma-separated "
                                 f"string of integers, but got: {raw}")
        else:
            self.select_layers = [-2]  # parsed but unused by feature_select, like the reference (quirk Q5)
        self.cfg_only = default_clip_config()
        self.image_processor = None
        # bring-up switch: route every GEMM through the plain CUDA-core kernel instead of tcgen05
        self.force_simple_gemm = os.environ.get("VZ_FORCE_SIMPLE_GEMM") == "1"
        self._w: Optional[_lib.VitWeights] = None
        self._packed: Dict[str, torch.Tensor] = {}
        self._ws = Workspace()
        self._graphs = GraphCache()
        self._dtype = torch.bfloat16
        # a zero-size parameter-free anchor so .to(device) / .device work before loading
        self.register_buffer("_anchor", torch.zeros(1), persistent=False)
        if not delay_load:
            self.load_model()

    # -- loading ---------------------------------------------------------------------------
    def load_model(self, state_dict: Optional[Dict[str, torch.Tensor]] = None, device=None):
        """Load CLIP weights.  `state_dict` uses HF CLIPVisionModel names ('vision_model.*');
        if omitted they are read from `vision_tower_path` with transformers."""
        from transformers import CLIPImageProcessor
        if state_dict is None:
            from transformers import CLIPVisionModel
            hf = CLIPVisionModel.from_pretrained(self.vision_tower_path)
            state_dict = hf.state_dict()
            self.cfg_only = hf.config
        try:
            self.image_processor = CLIPImageProcessor.from_pretrained(self.vision_tower_path)
        except Exception:
            from .preprocess import OPENAI_CLIP_MEAN, OPENAI_CLIP_STD
            self.image_processor = CLIPImageProcessor(
                size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336},
                image_mean=list(OPENAI_CLIP_MEAN), image_std=list(OPENAI_CLIP_STD), resample=3)
        dev = torch.device(device) if device is not None else self._anchor.device
        self._pack(state_dict, dev)
        self.is_loaded = True
        return self

    def _pack(self, sd: Dict[str, torch.Tensor], dev):
        """HF layout -> kernel layout: bf16 weights ([N,K] row-major, q/k/v stacked, conv weight
        flattened (c,ky,kx) and K-padded 588->592), fp32 biases; layer_norm1/2 folded into the
        following Linear (gamma into the weight, beta into the bias, plus the column sums the fused
        epilogue needs)."""
        sd = {k[len("vision_tower."):] if k.startswith("vision_tower.") else k: v for k, v in sd.items()}
        p = "vision_model."
        bf = lambda t: t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        P: Dict[str, torch.Tensor] = {}
        conv = sd[p + "embeddings.patch_embedding.weight"].detach().reshape(WIDTH, 588)
        pw = torch.zeros((WIDTH, PATCH_K), dtype=torch.bfloat16, device=dev)
        pw[:, :588] = conv.to(device=dev, dtype=torch.bfloat16)
        P["patch_w"] = pw
        P["class_emb"] = bf(sd[p + "embeddings.class_embedding"])
        P["pos_emb"] = bf(sd[p + "embeddings.position_embedding.weight"])
        P["pre_ln_g"] = f32(sd[p + "pre_layrnorm.weight"])
        P["pre_ln_b"] = f32(sd[p + "pre_layrnorm.bias"])
        for l in range(LAYERS):
            q = f"{p}encoder.layers.{l}."
            # LayerNorm folding (csrc/vz_gemm.cu): LN(x) W^T + b = rstd (x W'^T - mu colsum(W')) + b'
            def fold(W, b, g, beta):
                W32, g32, beta32 = W.detach().float().to(dev), g.detach().float().to(dev), beta.detach().float().to(dev)
                Wf = (W32 * g32[None, :]).to(torch.bfloat16).contiguous()
                return Wf, (b.detach().float().to(dev) + W32 @ beta32).contiguous(), Wf.float().sum(1).contiguous()
            Wqkv = torch.cat([sd[q + "self_attn.q_proj.weight"], sd[q + "self_attn.k_proj.weight"],
                              sd[q + "self_attn.v_proj.weight"]], 0)
            bqkv = torch.cat([sd[q + "self_attn.q_proj.bias"], sd[q + "self_attn.k_proj.bias"],
                              sd[q + "self_attn.v_proj.bias"]], 0)
            P[f"{l}.w_qkv"], P[f"{l}.b_qkv"], P[f"{l}.s_qkv"] = fold(Wqkv, bqkv, sd[q + "layer_norm1.weight"],
                                                                    sd[q + "layer_norm1.bias"])
            P[f"{l}.w_o"], P[f"{l}.b_o"] = bf(sd[q + "self_attn.out_proj.weight"]), f32(sd[q + "self_attn.out_proj.bias"])
            P[f"{l}.w_fc1"], P[f"{l}.b_fc1"], P[f"{l}.s_fc1"] = fold(sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"],
                                                                    sd[q + "layer_norm2.weight"], sd[q + "layer_norm2.bias"])
            P[f"{l}.w_fc2"], P[f"{l}.b_fc2"] = bf(sd[q + "mlp.fc2.weight"]), f32(sd[q + "mlp.fc2.bias"])
        self._packed = P
        self._rebuild_pointers()

    def _apply(self, fn, *a, **k):
        """nn.Module.to()/cuda(): follow the device move (dtypes of the packed buffers are fixed:
        bf16 weights, f32 biases / LayerNorm parameters) and refresh the pointer table."""
        out = super()._apply(fn, *a, **k)
        dev = self._anchor.device
        if self._packed and self._packed["patch_w"].device != dev:
            self._packed = {key: t.to(dev) for key, t in self._packed.items()}
            self._rebuild_pointers()
        return out

    def _rebuild_pointers(self):
        self._graphs.clear()                 # captured graphs hold the previous buffers' addresses
        self._pack_generation = next(PACK_GENERATION)
        P, w = self._packed, _lib.VitWeights()
        w.patch_w, w.class_emb, w.pos_emb = P["patch_w"].data_ptr(), P["class_emb"].data_ptr(), P["pos_emb"].data_ptr()
        w.pre_ln_g, w.pre_ln_b = P["pre_ln_g"].data_ptr(), P["pre_ln_b"].data_ptr()
        for l in range(LAYERS):
            for name in ("w_qkv", "b_qkv", "s_qkv", "w_o", "b_o", "w_fc1", "b_fc1", "s_fc1", "w_fc2", "b_fc2"):
                setattr(w.layers[l], name, P[f"{l}.{name}"].data_ptr())
        self._w = w

    # -- compute ----------------------------------------------------------------------------
    def _patches_of(self, images) -> torch.Tensor:
        """pixel tensor [T,3,336,336] (f32/bf16) -> bf16 patch rows via vz_patchify."""
        lib = _lib.load()
        if images.shape[-3:] != (3, 336, 336):
            raise ValueError(f"Input image size ({images.shape[-2]}*{images.shape[-1]}) doesn't match model (336*336).")
        x = images.to(self.device)
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.to(torch.float32)
        x = x.contiguous()
        T = x.shape[0]
        patches = torch.empty((T * PATCHES, PATCH_K), dtype=torch.bfloat16, device=x.device)
        _lib.check(lib.vz_patchify(_lib.ptr(x), 1 if x.dtype == torch.float32 else 0, T, _lib.ptr(patches),
                                   _lib.stream_ptr()), "vz_patchify")
        return patches

    def encode_patches(self, patches: torch.Tensor, pre_norm=None, return_hidden: bool = False, graph: bool = False):
        """patches bf16 [T*576,592] -> fused features bf16 [T,576,5120] (QFormer.pre_norm applied in
        the fusion kernel when pre_norm=(gamma_f32, beta_f32)).  graph=True (callers that consume the result at
        once): batches of up to VZ_GRAPH_MAX_TILES tiles replay a captured CUDA graph of the ~125 launches."""
        if graph and not return_hidden and GraphCache.usable(patches.shape[0] // PATCHES):
            # (generation counters, not addresses: the allocator may hand a re-packed weight the old address)
            key = ("vit", self._pack_generation, getattr(pre_norm[0], "_vz_generation", 0) if pre_norm is not None else -1)
            return self._graphs.run(key, [patches], lambda p: self.encode_patches(p, pre_norm))
        if not self.is_loaded:
            raise RuntimeError("vision tower weights are not loaded (call load_model())")
        if self.select_feature != "patch":
            if self.select_feature == "cls_patch":
                raise NotImplementedError("select_feature='cls_patch' is not on the B200 path")
            raise ValueError(f"Unknown feature selection strategy: {self.select_feature}")
        lib = _lib.load()
        T = patches.shape[0] // PATCHES
        dev = patches.device
        fused = torch.empty((T, PATCHES, 5 * WIDTH), dtype=torch.bfloat16, device=dev)
        nbytes = lib.vz_vit_workspace_bytes_ex(T, 1 if return_hidden else 0)
        ws = self._ws.get(nbytes, dev)
        hidden = torch.empty((LAYERS + 1, T, TOKENS, WIDTH), dtype=torch.bfloat16, device=dev) if return_hidden else None
        g = b = None
        if pre_norm is not None:
            g, b = pre_norm
        _lib.check(lib.vz_vit_forward(C.byref(self._w), _lib.ptr(patches), T, _lib.ptr(fused), _lib.ptr(g),
                                      _lib.ptr(b), _lib.ptr(hidden), _lib.ptr(ws), ws.numel(),
                                      1 if self.force_simple_gemm else 0, _lib.stream_ptr()), "vz_vit_forward")
        return (fused, hidden) if return_hidden else fused

    @torch.no_grad()
    def forward(self, images):
        """vision_encoder.py:80-117: list -> list of features; 4-D / 3-D tensor -> features cast back
        to the input dtype (quirk Q7)."""
        if isinstance(images, PatchBatch):
            return self.encode_patches(images.patches)
        if isinstance(images, list):
            feats = []
            for image in images:
                x = image if image.ndim == 4 else image.unsqueeze(0)
                feats.append(self.encode_patches(self._patches_of(x)).to(image.dtype))
            return feats
        if images.ndim == 3:
            images = images.unsqueeze(0)
        return self.encode_patches(self._patches_of(images)).to(images.dtype)

    # -- reference properties -------------------------------------------------------------------
    @property
    def dummy_feature(self):
        return torch.zeros(1, self.hidden_size, device=self.device, dtype=self.dtype)

    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        if self._packed:
            return self._packed["patch_w"].device
        return self._anchor.device

    @property
    def config(self):
        return self.cfg_only

    @property
    def hidden_size(self):
        return self.config.hidden_size * 5

    @property
    def num_patches(self):
        return (self.config.image_size // self.config.patch_size) ** 2

    @property
    def num_patches_per_side(self):
        return self.config.image_size // self.config.patch_size


def build_vision_tower(vision_tower_cfg, **kwargs):
    """vision_encoder/builder.py:8-24 (same acceptance rule and error)."""
    path = getattr(vision_tower_cfg, "mm_vision_tower", getattr(vision_tower_cfg, "vision_tower", None))
    exists = os.path.exists(path) if path is not None else False
    if path is not None and (exists or path.startswith("openai") or path.startswith("laion")):
        return CLIPVisionTowerB200(vision_tower_path=path, args=vision_tower_cfg, **kwargs)
    raise ValueError(f"Unknown vision tower path: {path}")
