"""Multimodal glue: encode_images and prepare_inputs_labels_for_multimodal on the B200 kernels.

Drop-in for the reference's mixins
  VisZephyrMetaModel        vis_zephyr/model/vis_zephyr_arch.py:22-102
  VisZephyrMetaForCausalLM  vis_zephyr/model/vis_zephyr_arch.py:107-530
`VisZephyrB200MetaForCausalLM` keeps the method names, argument order and the 6-tuple that
`VisZephyrForCausalLM.forward/generate` (language_model/vis_zephyr.py:76-84,124-132) expect, so a
maintainer swaps the base class.  The Python loops and host syncs of the
reference (:236-305, :512-528) are replaced by vz_splice_plan / vz_text_gather /
vz_splice_scatter; the only host round-trip left is one small read of the planned lengths,
which the output tensor shapes depend on.
"""
from __future__ import annotations

import ctypes as C
from abc import ABC, abstractmethod
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .anyres import slot_descriptor
from .constants import IGNORE_INDEX, IMAGE_TOKEN_INDEX
from .preprocess import PatchBatch
from .projector import TextPack, build_multimodal_projector
from .vision_tower import build_vision_tower


class VisZephyrB200MetaModel:
    """Owns vision_tower + mm_projector (+ image_newline), like vis_zephyr_arch.py:22-47."""

    def __init__(self, config):
        super().__init__(config)
        if hasattr(config, "mm_vision_tower"):
            self.vision_tower = build_vision_tower(config, delay_load=True)
            self.mm_projector = build_multimodal_projector(config)
            if "unpad" in getattr(config, "mm_patch_merge_type", ""):
                self.image_newline = nn.Parameter(torch.empty(config.hidden_size, dtype=self.dtype))

    def get_vision_tower(self):
        vt = getattr(self, "vision_tower", None)
        if type(vt) is list:
            vt = vt[0]
        return vt

    def initialize_vision_modules(self, model_args, fsdp=None):
        """vis_zephyr_arch.py:49-102, statement for statement (called from train/train.py:803 after the LLM is
        built): build or load the tower (kept in a one-element list under FSDP so it is not wrapped), write the
        nine config.mm_* attributes, build the projector (+ image_newline ~ N(0, 1/hidden) for 'unpad' merge
        types) or un-freeze an existing one, and load `pretrain_mm_mlp_adapter` (an mm_projector.bin written by
        train/vis_zephyr_trainer.py:304-348) by stripping the 'mm_projector.' prefix."""
        vision_tower = model_args.mm_vision_tower
        mm_vision_select_layer = model_args.mm_vision_select_layer
        mm_vision_select_feature = model_args.mm_vision_select_feature
        mm_projector_train = model_args.pretrain_mm_mlp_adapter
        mm_patch_merge_type = model_args.mm_patch_merge_type

        self.config.mm_vision_tower = vision_tower
        if self.get_vision_tower() is None:
            vision_tower = build_vision_tower(model_args)
            self.vision_tower = [vision_tower] if fsdp and len(fsdp) > 0 else vision_tower
        else:
            vision_tower = self.vision_tower[0] if fsdp and len(fsdp) > 0 else self.vision_tower
            vision_tower.load_model()

        self.config.use_mm_proj = True
        self.config.mm_projector_type = getattr(model_args, "mm_projector_type", "linear")
        # the reference reads self.vision_tower.hidden_size (:73), which breaks on the FSDP list form; same value
        self.config.mm_hidden_size = vision_tower.hidden_size
        self.config.mm_vision_select_layer = mm_vision_select_layer
        self.config.mm_vision_select_feature = mm_vision_select_feature
        self.config.mm_patch_merge_type = mm_patch_merge_type
        self.config.mm_grid_pinpoints = getattr(model_args, "mm_grid_pinpoints", None)
        self.config.image_aspect_ratio = getattr(model_args, "image_aspect_ratio", "square")
        self.config.mm_use_im_start_end = getattr(model_args, "mm_use_im_start_end", False)

        if getattr(self, "mm_projector", None) is None:
            self.mm_projector = build_multimodal_projector(self.config)
            if "unpad" in mm_patch_merge_type:
                embed_std = 1 / torch.sqrt(torch.tensor(self.config.hidden_size, dtype=self.dtype))
                self.image_newline = nn.Parameter(torch.randn(self.config.hidden_size, dtype=self.dtype) * embed_std)
        else:
            for p in self.mm_projector.parameters():   # un-freeze a projector LoRA froze
                p.requires_grad = True

        if mm_projector_train is not None:
            projector_weights = torch.load(mm_projector_train, map_location="cpu")

            def get_w(weights, keyword):
                return {k.split(keyword + ".")[1]: v for k, v in weights.items() if keyword in k}

            self.mm_projector.load_state_dict(get_w(projector_weights, "mm_projector"))


def _slots_to_device(descs: List[dict], device):
    n = len(descs)
    arr = (_lib.SlotDesc * max(n, 1))()
    prefix = np.zeros(n + 1, np.int32)
    for i, d in enumerate(descs):
        for k, v in d.items():
            setattr(arr[i], k, int(v))
        prefix[i + 1] = prefix[i] + d["n_rows"]
    raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
    return (_lib.h2d(raw, device), _lib.h2d(prefix, device), int(prefix[-1]))


class SplicePlan:
    """Device outputs of vz_splice_plan + the host copy of the totals."""

    def __init__(self, tok_dest, slot_dest, lengths, text_len, totals_dev, host_meta, B, event):
        self.tok_dest, self.slot_dest, self.lengths, self.text_len = tok_dest, slot_dest, lengths, text_len
        self.totals_dev, self._host, self._B, self._event = totals_dev, host_meta, B, event

    def wait(self):
        """Block until the planned lengths are on the host (the only host round-trip of the path)."""
        self._event.synchronize()
        B, h = self._B, self._host
        t = h[2 * B:]
        return dict(Lmax=int(t[0]), L_text=int(t[1]), slots_used=int(t[2]), text_rows=int(t[3]),
                    lengths=h[:B].tolist(), text_len=h[B:2 * B].tolist())


def splice_plan(input_ids: torch.Tensor, mask_u8: Optional[torch.Tensor], slots_dev, n_slots: int,
                max_len: int = 0) -> SplicePlan:
    lib = _lib.load()
    B, S = input_ids.shape
    dev = input_ids.device
    tok_dest = torch.empty((B, S), dtype=torch.int32, device=dev)
    slot_dest = torch.empty((max(n_slots, 1), 2), dtype=torch.int32, device=dev)
    meta = torch.empty((2 * B + 4,), dtype=torch.int32, device=dev)  # lengths | text_len | totals
    lengths, text_len, totals = meta[:B], meta[B:2 * B], meta[2 * B:]
    _lib.check(lib.vz_splice_plan(_lib.ptr(input_ids), _lib.ptr(mask_u8), B, S, _lib.ptr(slots_dev), n_slots,
                                  int(max_len), _lib.ptr(tok_dest), _lib.ptr(slot_dest), _lib.ptr(lengths),
                                  _lib.ptr(text_len), _lib.ptr(totals), _lib.stream_ptr()), "vz_splice_plan")
    host = torch.empty((2 * B + 4,), dtype=torch.int32, pin_memory=True)
    host.copy_(meta, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    return SplicePlan(tok_dest, slot_dest, lengths, text_len, totals, host, B, ev)


def text_gather(input_ids: torch.Tensor, embed_weight: torch.Tensor, plan: SplicePlan, text_rows: int,
                lo: int = 0, hi: Optional[int] = None):
    """Packed non-image token embeddings of samples [lo, hi) (+ one zero row) and their offsets."""
    lib = _lib.load()
    B, S = input_ids.shape
    hi = B if hi is None else hi
    n = hi - lo
    D = embed_weight.shape[1]
    dev = input_ids.device
    text_emb = torch.empty((text_rows + 1, D), dtype=embed_weight.dtype, device=dev)
    text_off = torch.zeros((n + 1,), dtype=torch.int32, device=dev)
    if n > 0:
        _lib.check(lib.vz_text_gather(_lib.ptr(input_ids[lo:hi]), n, S, _lib.ptr(embed_weight), D,
                                      embed_weight.element_size(), _lib.ptr(plan.text_len[lo:hi]),
                                      _lib.ptr(text_emb), _lib.ptr(text_off), _lib.stream_ptr()), "vz_text_gather")
    else:
        text_emb.zero_()
    return text_emb, text_off


def splice_scatter(input_ids, labels, embed_weight, vis, image_newline, slots_dev, slot_prefix, n_slots,
                   total_vis_rows, plan: SplicePlan, Lout: int, pad_left: bool, row_stats: bool = False):
    """row_stats=True (bf16 tables only): also returns, as a 5th tensor, f32 [B, Lout, 2] = (0, sum of squares) of
    every output row -- the RMSNorm statistic of the first LLM layer (language_model.first_layer_qkv)."""
    lib = _lib.load()
    B, S = input_ids.shape
    D = embed_weight.shape[1]
    dev = input_ids.device
    out_embeds = torch.empty((B, Lout, D), dtype=embed_weight.dtype, device=dev)
    out_labels = torch.empty((B, Lout), dtype=torch.int64, device=dev)
    out_mask = torch.empty((B, Lout), dtype=torch.uint8, device=dev)
    out_pos = torch.empty((B, Lout), dtype=torch.int64, device=dev)
    ldv = vis.stride(0) if vis is not None else D
    stats = torch.empty((B, Lout, 2), dtype=torch.float32, device=dev) if row_stats else None
    _lib.check(lib.vz_splice_scatter_rms(
        _lib.ptr(input_ids), _lib.ptr(labels), B, S, _lib.ptr(embed_weight), _lib.ptr(vis), int(ldv),
        _lib.ptr(image_newline), D, embed_weight.element_size(), _lib.ptr(slots_dev), n_slots,
        _lib.ptr(slot_prefix), int(total_vis_rows), _lib.ptr(plan.tok_dest), _lib.ptr(plan.slot_dest),
        _lib.ptr(plan.lengths), int(Lout), 1 if pad_left else 0, _lib.ptr(out_embeds), _lib.ptr(out_labels),
        _lib.ptr(out_mask), _lib.ptr(out_pos), _lib.ptr(stats), _lib.stream_ptr()), "vz_splice_scatter")
    if row_stats:
        return out_embeds, out_labels, out_mask, out_pos, stats
    return out_embeds, out_labels, out_mask, out_pos


def merge_rows(vis: torch.Tensor, image_newline, descs: List[dict]) -> List[torch.Tensor]:
    """_process_image_patches (vis_zephyr_arch.py:396-473) alone: list of merged [n_i, D]."""
    lib = _lib.load()
    dev = vis.device
    D = vis.shape[-1]
    vis2 = vis.reshape(-1, D)
    slots_dev, prefix, total = _slots_to_device(descs, dev)
    out = torch.empty((total, D), dtype=vis.dtype, device=dev)
    _lib.check(lib.vz_merge_rows(_lib.ptr(vis2), vis2.stride(0), _lib.ptr(image_newline), D, vis.element_size(),
                                 _lib.ptr(slots_dev), len(descs), _lib.ptr(prefix), _lib.ptr(out),
                                 _lib.stream_ptr()), "vz_merge_rows")
    return list(torch.split(out, [d["n_rows"] for d in descs], dim=0))


class _SpliceFn(torch.autograd.Function):
    """vz_splice_scatter under autograd (training: the LLM loss must reach the projector through the spliced
    rows).  Forward = the scatter kernel.  Backward needs, for every output row, where it came from; instead of
    a second kernel the SAME scatter is run once more on index-coded 16-byte rows (token id + 1 for embedding
    rows, -(row + 1) for visual rows, -(R + 1) for image_newline, 0 for padding), which yields the exact
    provenance map of this very batch; the gradients are then index_add's of the incoming rows."""

    @staticmethod
    def forward(ctx, vis, embed, newline, sctx, Lmax, pad_left):
        nl = newline.detach().to(device=embed.device, dtype=embed.dtype).contiguous() if newline is not None else None
        out = splice_scatter(sctx["ids"], sctx["labels"], embed.detach(), vis.detach().to(embed.dtype).contiguous(), nl,
                             sctx["slots"], sctx["prefix"], sctx["n_images"], sctx["total_vis_rows"], sctx["plan"], Lmax, pad_left)
        ctx.sctx, ctx.Lmax, ctx.pad_left = sctx, Lmax, pad_left
        ctx.shapes = (vis.shape, vis.dtype, embed.shape, embed.dtype, newline is not None)
        ctx.mark_non_differentiable(out[1], out[2], out[3])
        return out

    @staticmethod
    def backward(ctx, d_emb, *unused):
        sctx = ctx.sctx
        vshape, vdt, eshape, edt, has_nl = ctx.shapes
        dev = d_emb.device
        R, V = vshape[0], eshape[0]
        code_table = (torch.arange(V, device=dev, dtype=torch.float32) + 1)[:, None].expand(V, 4).contiguous()
        code_vis = (-(torch.arange(max(R, 1), device=dev, dtype=torch.float32) + 1))[:, None].expand(max(R, 1), 4).contiguous()
        code_nl = torch.full((4,), -float(R + 1), device=dev, dtype=torch.float32) if has_nl else None
        codes = splice_scatter(sctx["ids"], sctx["labels"], code_table, code_vis, code_nl, sctx["slots"], sctx["prefix"],
                               sctx["n_images"], sctx["total_vis_rows"], sctx["plan"], ctx.Lmax, ctx.pad_left)[0]
        src = codes[..., 0].long().reshape(-1)
        g = d_emb.reshape(-1, d_emb.shape[-1])
        d_vis = d_embed = d_nl = None
        if ctx.needs_input_grad[0]:
            m = (src < 0) & (src >= -R)
            d_vis = torch.zeros(vshape, dtype=torch.float32, device=dev)
            d_vis.index_add_(0, -src[m] - 1, g[m].float())
            d_vis = d_vis.to(vdt)
        if ctx.needs_input_grad[1]:
            m = src > 0
            d_embed = torch.zeros(eshape, dtype=torch.float32, device=dev)
            d_embed.index_add_(0, src[m] - 1, g[m].float())
            d_embed = d_embed.to(edt)
        if has_nl and ctx.needs_input_grad[2]:
            d_nl = g[src == -(R + 1)].float().sum(0).to(edt)
        return d_vis, d_embed, d_nl, None, None, None


class VisZephyrB200MetaForCausalLM(ABC):
    """Same surface as VisZephyrMetaForCausalLM (vis_zephyr_arch.py:107)."""

    @abstractmethod
    def get_model(self):
        pass

    def get_vision_tower(self):
        return self.get_model().get_vision_tower()

    # ------------------------------------------------------------------------------------------
    def encode_images(self, images, text_embeddings):
        """vis_zephyr_arch.py:120-124.  `text_embeddings` may be the reference's dense
        [T,L,4096] tensor, None, or a TextPack (the de-duplicated form used internally)."""
        tower = self.get_model().get_vision_tower()
        proj = self.get_model().mm_projector
        if isinstance(images, PatchBatch):
            patches = images.patches
        else:
            if isinstance(images, (list, tuple)):
                images = torch.cat([x if x.ndim == 4 else x.unsqueeze(0) for x in images], dim=0)
            patches = tower._patches_of(images if images.ndim == 4 else images.unsqueeze(0))
        if text_embeddings is None or isinstance(text_embeddings, TextPack):
            # QFormer.pre_norm rides in the tower's fusion kernel
            feats = tower.encode_patches(patches, pre_norm=proj.pre_norm_params())
            return proj.forward_packed(feats, text_embeddings, feats_normed=True)
        return proj(tower.encode_patches(patches), text_embeddings=text_embeddings)

    # ------------------------------------------------------------------------------------------
    def _slot_descs(self, tiles_per_image: Sequence[int], images_size, rows_per_tile: int):
        cfg = self.config
        merge = getattr(cfg, "mm_patch_merge_type", "flat")
        aspect = getattr(cfg, "image_aspect_ratio", "square")
        tower = self.get_vision_tower()
        descs, base = [], 0
        for i, t in enumerate(tiles_per_image):
            size = images_size[i] if images_size is not None else None
            side = None
            if merge.startswith("spatial") and t > 1:
                # the reference reads vision_tower.num_patches_per_side and asserts h*w == rows
                # (vis_zephyr_arch.py:423-424); with the 32-row Q-Former output this cannot hold (quirk Q2)
                side = getattr(self, "merge_side_override", None) or tower.num_patches_per_side
                assert side * side == rows_per_tile
            descs.append(slot_descriptor(base, t, rows_per_tile, merge, aspect, size,
                                         getattr(cfg, "mm_grid_pinpoints", None),
                                         tower.config.image_size, side))
            base += t * rows_per_tile
        return descs

    def prepare_inputs_labels_for_multimodal(self, input_ids, position_ids, attention_mask, past_key_values,
                                             labels, images, images_size=None):
        """vis_zephyr_arch.py:129-333, same arguments and return tuple."""
        return self._prepare(input_ids, position_ids, attention_mask, past_key_values, labels, images,
                             images_size, None, None, 0)

    def prepare_inputs_labels_for_multimodal_sharded(self, input_ids, position_ids, attention_mask,
                                                     past_key_values, labels, local_images, tiles_per_image,
                                                     images_size=None, group=None, dst=0, keep_local=False):
        """Data-parallel form (north_star): every rank holds the (small) global text batch and encodes
        only ITS contiguous block of images (`local_images`, a PatchBatch or list of tile tensors for the
        images dist.shard_images assigns to this rank); the projected visual tokens travel to rank `dst`
        (peer stores or one all-gather, `self.last_transport` says which), which splices the global batch.
        Other ranks return None for the tensors.  keep_local=True (checks only) also leaves this rank's
        projected rows in `self.last_local_tokens`."""
        return self._prepare(input_ids, position_ids, attention_mask, past_key_values, labels, local_images,
                             images_size, list(tiles_per_image), group, dst, keep_local)

    def _prepare(self, input_ids, position_ids, attention_mask, past_key_values, labels, images, images_size,
                 global_tiles, group, dst, keep_local=False):
        vision_tower = self.get_vision_tower()
        if vision_tower is None or images is None or input_ids.shape[1] == 1:
            return input_ids, position_ids, attention_mask, past_key_values, None, labels
        model = self.get_model()
        embed = model.embed_tokens.weight
        dev = embed.device
        if not embed.is_cuda:
            raise _lib.VzError("the B200 path needs the model on a CUDA device (no CPU fallback)")

        # ---- images -> per-image tile counts ----------------------------------------------------
        if isinstance(images, PatchBatch):
            local_tiles = images.tiles_per_image
            patches = images.patches
        elif type(images) is list or images.ndim == 5:
            per_image = [x.unsqueeze(0) if x.ndim == 3 else x for x in images]
            local_tiles = [int(x.shape[0]) for x in per_image]
            patches = vision_tower._patches_of(torch.cat(per_image, dim=0)) if per_image else None
        else:
            # the reference's 4-D / 3-D branch fails inside QFormer.forward (quirk Q1)
            raise RuntimeError("Tensors must have same number of dimensions: got 3 and 2 "
                               "(pass images as a list of [T_i,3,336,336] tensors or a 5-D tensor)")
        proj = model.mm_projector
        sharded = global_tiles is not None
        if sharded:
            import torch.distributed as tdist
            from .dist import (TRANSPORT_NCCL, TRANSPORT_PEER, gather_visual_tokens, peer_gather_for, shard_images)
            world, rank = tdist.get_world_size(group), tdist.get_rank(group)
            tiles_per_image = global_tiles
            bounds = shard_images(tiles_per_image, world)
            lo, hi = bounds[rank]
            if list(local_tiles) != list(tiles_per_image[lo:hi]):
                raise ValueError(f"rank {rank} was given tiles {local_tiles}, expected {tiles_per_image[lo:hi]}")
            rows_per_rank = [sum(tiles_per_image[a:b]) * proj.num_queries for a, b in bounds]
            peer = peer_gather_for(sum(rows_per_rank), embed.shape[1], torch.bfloat16, dev, group)
            self.last_transport = TRANSPORT_PEER if peer is not None else TRANSPORT_NCCL
        else:
            tiles_per_image, lo, hi = list(local_tiles), 0, len(local_tiles)
            peer = None

        # ---- plan on the device while the host queues the tower ----------------------------------
        ctx = self._plan_splice(input_ids, attention_mask, labels, tiles_per_image, images_size)

        # ---- tower (independent of the text) ---------------------------------------------------
        # (training: pre_norm is a trained parameter, so it stays out of the tower's fusion kernel)
        train = proj.needs_autograd(embed)
        feats = None
        if hi > lo:
            # (small batches replay a CUDA graph; the features are consumed by the projector right below)
            feats = vision_tower.encode_patches(patches, pre_norm=None if train else proj.pre_norm_params(),
                                                graph=not train and peer is None)

        out_view = None
        if peer is not None and hi > lo and not keep_local:
            # the exchange step IS the projector's last kernel: its final LayerNorm stores go straight
            # into the destination rank's receive buffer (peer-mapped over NVLink)
            out_view = peer.slot(dst, sum(rows_per_rank[:rank]), rows_per_rank[rank])
        vis_local = self._project_shard(ctx, feats, tiles_per_image, lo, hi, out_view)
        if keep_local:
            self.last_local_tokens = vis_local

        # ---- the one exchange step ------------------------------------------------------------
        if sharded and peer is not None:
            if keep_local and hi > lo:
                peer.slot(dst, sum(rows_per_rank[:rank]), rows_per_rank[rank]).copy_(vis_local)
            peer.finish()                                   # one device-side barrier; no data collective
            if rank != dst:
                peer.skip_local()
                return None, None, None, past_key_values, None, None
            vis = peer.local(sum(rows_per_rank))
        elif sharded:
            vis = gather_visual_tokens(vis_local, rows_per_rank, group)
            if rank != dst:
                return None, None, None, past_key_values, None, None
        else:
            vis = vis_local
        return self._splice(ctx, vis, position_ids, attention_mask, past_key_values, labels)

    # -- the three stages of the path (also the units bench.py times and the multi-GPU checks re-run) ------
    def _plan_splice(self, input_ids, attention_mask, labels, tiles_per_image, images_size):
        """Upload ids / mask / labels + the slot table and launch the plan kernel (no host wait)."""
        model = self.get_model()
        embed = model.embed_tokens.weight
        dev = embed.device
        n_images = len(tiles_per_image)
        B, S = input_ids.shape
        if n_images > B:
            raise IndexError("index out of range: more images than samples")  # input_ids[i], :167
        ids_dev = input_ids.to(dev).contiguous()
        mask_u8 = None
        if attention_mask is not None:
            mask_u8 = attention_mask.to(dev).bool().to(torch.uint8).contiguous()
        labels_dev = labels.to(dev).contiguous() if labels is not None else None
        descs = self._slot_descs(tiles_per_image, images_size, model.mm_projector.num_queries)
        slots_dev, slot_prefix, total_vis_rows = _slots_to_device(descs, dev)
        max_len = getattr(self.config, "tokenizer_model_max_length", None) or 0
        plan = splice_plan(ids_dev, mask_u8, slots_dev, n_images, max_len)
        return dict(ids=ids_dev, mask=mask_u8, labels=labels_dev, slots=slots_dev, prefix=slot_prefix,
                    total_vis_rows=total_vis_rows, plan=plan, n_images=n_images, info=None)

    @staticmethod
    def _plan_info(ctx):
        """The one host round-trip of the path: the planned lengths (output shapes depend on them)."""
        if ctx["info"] is None:
            info = ctx["plan"].wait()
            if info["slots_used"] > ctx["n_images"]:
                raise IndexError("list index out of range: more image slots consumed than image features")
            ctx["info"] = info
        return ctx["info"]

    def _project_shard(self, ctx, feats, tiles_per_image, lo, hi, out_view=None):
        """Text conditioning + Q-Former for images [lo, hi): one text row set per SAMPLE, shared by its tiles;
        the reference conditions image i on ids[i] (:163-176) and pads to the BATCH-GLOBAL max (quirk Q3), so L
        comes from the global plan.  Returns bf16 [sum T_i * 32, 4096] (written into out_view when given)."""
        model = self.get_model()
        embed, proj = model.embed_tokens.weight, model.mm_projector
        dev = embed.device
        info = self._plan_info(ctx)
        if hi <= lo:
            return torch.empty((0, embed.shape[1]), dtype=torch.bfloat16, device=dev)
        L_text = max(info["text_len"][:ctx["n_images"]])
        if proj.needs_autograd(embed):
            # training: the text rows come from embed_tokens under autograd (the reference's own expression,
            # :163-189), so a trainable embedding table receives its gradient through the projector too
            rows = []
            for b in range(lo, hi):
                ids_b = ctx["ids"][b]
                e = torch.nn.functional.embedding(ids_b[ids_b != IMAGE_TOKEN_INDEX], embed)
                rows.append(torch.nn.functional.pad(e, (0, 0, 0, L_text - e.shape[0])))
            tile_sample = _lib.h2d(torch.repeat_interleave(torch.arange(hi - lo), torch.tensor(tiles_per_image[lo:hi])), dev)
            from .projector_train import qformer_train_forward
            vis_local = qformer_train_forward(proj, feats, torch.stack(rows), tile_sample)
            return vis_local.reshape(-1, vis_local.shape[-1])
        text_rows = sum(info["text_len"][lo:hi])
        text_emb, text_off = text_gather(ctx["ids"], embed, ctx["plan"], text_rows, lo, hi)
        if text_emb.dtype != torch.bfloat16:
            text_emb = text_emb.to(torch.bfloat16)
        tile_sample = _lib.h2d(torch.repeat_interleave(torch.arange(hi - lo, dtype=torch.int32),
                                                       torch.tensor(tiles_per_image[lo:hi])), dev)
        text = TextPack(text_emb, text_off, text_rows, hi - lo, L_text, tile_sample)
        if out_view is not None:
            out_view = out_view.view(sum(tiles_per_image[lo:hi]), proj.num_queries, -1)
        # (graph replay for small batches; its static output is consumed by the scatter before the next call)
        vis_local = proj.forward_packed(feats, text, feats_normed=True, out=out_view, graph=out_view is None)      # [T,32,4096] bf16
        return vis_local.reshape(-1, vis_local.shape[-1])

    def _splice(self, ctx, vis, position_ids, attention_mask, past_key_values, labels):
        """merge + splice + pad/collate of the global batch (vis_zephyr_arch.py:214-333, :396-530)."""
        model = self.get_model()
        embed = model.embed_tokens.weight
        dev = embed.device
        info = self._plan_info(ctx)
        if vis.dtype != embed.dtype:
            vis = vis.to(embed.dtype)
        newline = getattr(model, "image_newline", None)
        if newline is not None:
            newline = newline.detach().to(device=dev, dtype=embed.dtype).contiguous()
        pad_left = getattr(self.config, "tokenizer_padding_side", "right") == "left"
        newline_p = getattr(model, "image_newline", None)
        if torch.is_grad_enabled() and (vis.requires_grad or embed.requires_grad or
                                        (newline_p is not None and newline_p.requires_grad)):
            out_embeds, out_labels, out_mask, out_pos = _SpliceFn.apply(
                vis, embed, newline_p if newline_p is not None else None, ctx, info["Lmax"], pad_left)
        else:
            want_stats = bool(getattr(self.config, "vz_first_layer_stats", False)) and embed.dtype == torch.bfloat16
            res = splice_scatter(ctx["ids"], ctx["labels"], embed, vis, newline, ctx["slots"], ctx["prefix"],
                                 ctx["n_images"], ctx["total_vis_rows"], ctx["plan"], info["Lmax"], pad_left, want_stats)
            out_embeds, out_labels, out_mask, out_pos = res[:4]
            if want_stats:
                # rides on the tensor object: language_model.first_layer_qkv picks it up (config.vz_first_layer_stats)
                out_embeds.vz_row_sumsq = res[4]
        new_mask = None
        if attention_mask is not None:
            new_mask = out_mask.to(dtype=attention_mask.dtype)
        return (None,
                out_pos.to(position_ids.dtype) if position_ids is not None else None,
                new_mask,
                past_key_values,
                out_embeds,
                out_labels if labels is not None else None)
