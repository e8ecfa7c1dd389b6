"""The step in front of the splice ("next" row 2 of SURVEY.md 8(f)): prompt -> token ids with the image
placeholder, and the supervised collator, mirroring
  tokenizer_image_token                vis_zephyr/model/mm_utils.py:91-128
  DataCollatorForSupervisedDataset     vis_zephyr/train/train.py:657-707
The tokenizer itself stays the caller's (HF) object; padding / truncation / mask run on the GPU."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .constants import IGNORE_INDEX, IMAGE_TOKEN_INDEX


def tokenizer_image_token(prompt: str, tokenizer, image_token_index: int = IMAGE_TOKEN_INDEX,
                          return_tensors: Optional[str] = None):
    """Tokenise the text around every '<image>' and put `image_token_index` in between; a BOS token
    produced for the first chunk is kept once, the BOS of later chunks is dropped."""
    chunks = [tokenizer(piece).input_ids for piece in prompt.split("<image>")]
    ids: List[int] = []
    skip = 0
    if chunks and chunks[0] and chunks[0][0] == tokenizer.bos_token_id:
        skip = 1
        ids.append(chunks[0][0])
    for n, chunk in enumerate(chunks):
        if n > 0:
            ids.append(image_token_index)
        ids.extend(chunk[skip:])
    if return_tensors is not None:
        if return_tensors == "pt":
            return torch.tensor(ids, dtype=torch.long)
        raise ValueError(f"Unknown return_tensor type: {return_tensors}")
    return ids


def collate_supervised(instances: Sequence[Dict], pad_token_id: int, model_max_length: int, device="cuda"):
    """Batch dict of the reference collator (input_ids, labels, attention_mask[, images, images_size]);
    the ragged rows are packed on the host, copied once and padded / truncated / masked by vz_collate."""
    lib = _lib.load()
    ids = [torch.as_tensor(x["input_ids"], dtype=torch.long).reshape(-1) for x in instances]
    labs = [torch.as_tensor(x["labels"], dtype=torch.long).reshape(-1) for x in instances]
    B = len(ids)
    lens = [int(t.numel()) for t in ids]
    if any(int(l.numel()) != n for l, n in zip(labs, lens)):
        raise ValueError("input_ids and labels of an instance must have the same length")
    S_out = min(max(lens), int(model_max_length))
    offs = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int32))
    flat_ids = torch.cat(ids).to(device, non_blocking=True)
    flat_labs = torch.cat(labs).to(device, non_blocking=True)
    offs_d = offs.to(device, non_blocking=True)
    out_ids = torch.empty((B, S_out), dtype=torch.long, device=device)
    out_labels = torch.empty((B, S_out), dtype=torch.long, device=device)
    out_mask = torch.empty((B, S_out), dtype=torch.uint8, device=device)
    _lib.check(lib.vz_collate(_lib.ptr(flat_ids), _lib.ptr(flat_labs), _lib.ptr(offs_d), B, S_out, int(pad_token_id),
                              _lib.ptr(out_ids), _lib.ptr(out_labels), _lib.ptr(out_mask), _lib.stream_ptr()),
               "vz_collate")
    batch = dict(input_ids=out_ids, labels=out_labels, attention_mask=out_mask.bool())
    if "image" in instances[0]:
        batch["images"] = [x["image"] for x in instances]   # the B200 path takes per-image lists / PatchBatch
        if "images_size" in instances[0]:
            batch["images_size"] = [x["images_size"] for x in instances]
    return batch
