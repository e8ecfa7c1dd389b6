"""numpy restatement of the merge + splice half of the path (test infrastructure).

unpad_image            vis_zephyr/model/multi_scale_process.py:188-211 (AS WRITTEN, quirk Q4)
process_image_patches  vis_zephyr/model/vis_zephyr_arch.py:396-473
splice                 vis_zephyr/model/vis_zephyr_arch.py:214-333 + :476-530
"""
import numpy as np

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200


def unpad_image(t, original_size):
    """t [D,H,W]; note the reference's `current_w, current_h = t.shape[1:]`."""
    original_w, original_h = original_size
    current_w, current_h = t.shape[1:]
    if original_w / original_h > current_w / current_h:
        factor = current_w / original_w
        new_h = int(original_h * factor)
        padding = (current_h - new_h) // 2
        return t[:, padding:current_h - padding, :]
    factor = current_h / original_h
    new_w = int(original_w * factor)
    padding = (current_w - new_w) // 2
    return t[:, :, padding:current_w - padding]


def process_image_patches(features, images_size, merge_type, grid_shapes, image_newline=None, side=None):
    """features: list of [T_i, hw, D] arrays; grid_shapes[i] = (n_w, n_h) from calculate_grid_shape."""
    if merge_type == "flat":
        return [f.reshape(-1, f.shape[-1]) for f in features]
    if not merge_type.startswith("spatial"):
        raise ValueError(f"Unknown mm_patch_merge_type: {merge_type}")
    out = []
    for i, f in enumerate(features):
        if f.shape[0] > 1:
            base, tiles = f[0], f[1:]
            h = w = side
            assert h * w == base.shape[0]
            n_w, n_h = grid_shapes[i]
            tiles = tiles.reshape(n_h, n_w, h, w, -1)
            if "unpad" in merge_type:
                tiles = np.ascontiguousarray(tiles.transpose(4, 0, 2, 1, 3))
                tiles = tiles.reshape(tiles.shape[0], n_h * h, n_w * w)
                tiles = unpad_image(tiles, images_size[i])
                nl = np.broadcast_to(image_newline[:, None, None], tiles.shape[:-1] + (1,))
                tiles = np.concatenate([tiles, nl], axis=-1)
                tiles = tiles.reshape(tiles.shape[0], -1).T
            else:
                tiles = np.ascontiguousarray(tiles.transpose(0, 2, 1, 3, 4)).reshape(-1, tiles.shape[-1])
            f = np.concatenate([base, tiles], axis=0)
        else:
            f = f[0]
            if "unpad" in merge_type:
                f = np.concatenate([f, image_newline[None]], axis=0)
        out.append(f)
    return out


def splice(input_ids, attention_mask, labels, position_ids_given, embed, image_features,
           max_len=None, padding_side="right"):
    """Returns (embeds [B,Lmax,D], labels [B,Lmax], mask bool [B,Lmax], pos [B,Lmax], lengths).
    input_ids int64 [B,S]; attention_mask bool [B,S] or None; labels int64 [B,S] or None;
    image_features: list of [n_i, D]."""
    B, S = input_ids.shape
    D = embed.shape[1]
    if attention_mask is None:
        attention_mask = np.ones((B, S), bool)
    else:
        attention_mask = attention_mask.astype(bool)
    if labels is None:
        labels = np.full((B, S), IGNORE_INDEX, np.int64)
    new_embeds, new_labels = [], []
    cur = 0
    for b in range(B):
        ids = input_ids[b][attention_mask[b]]
        lab = labels[b][attention_mask[b]]
        pos = np.where(ids == IMAGE_TOKEN_INDEX)[0].tolist()
        if len(pos) == 0:
            _ = image_features[cur]  # the slot is consumed (IndexError if there is none)
            new_embeds.append(embed[ids])
            new_labels.append(lab)
            cur += 1
            continue
        bounds = [-1] + pos + [ids.shape[0]]
        e_parts, l_parts = [], []
        for i in range(len(bounds) - 1):
            chunk = ids[bounds[i] + 1:bounds[i + 1]]
            e_parts.append(embed[chunk])
            l_parts.append(lab[bounds[i] + 1:bounds[i + 1]])
            if i < len(pos):
                f = image_features[cur]
                e_parts.append(f)
                l_parts.append(np.full((f.shape[0],), IGNORE_INDEX, np.int64))
                cur += 1
        new_embeds.append(np.concatenate(e_parts, axis=0))
        new_labels.append(np.concatenate(l_parts, axis=0))
    if max_len is not None:
        new_embeds = [x[:max_len] for x in new_embeds]
        new_labels = [x[:max_len] for x in new_labels]
    Lmax = max(x.shape[0] for x in new_embeds)
    out_e = np.zeros((B, Lmax, D), embed.dtype)
    out_l = np.full((B, Lmax), IGNORE_INDEX, np.int64)
    out_m = np.zeros((B, Lmax), bool)
    out_p = np.zeros((B, Lmax), np.int64)
    lengths = []
    for b, (e, l) in enumerate(zip(new_embeds, new_labels)):
        n = e.shape[0]
        lengths.append(n)
        if n == 0:
            continue
        sl = slice(Lmax - n, Lmax) if padding_side == "left" else slice(0, n)
        out_e[b, sl] = e
        out_l[b, sl] = l
        out_m[b, sl] = True
        out_p[b, sl] = np.arange(n)
    return out_e, out_l, out_m, out_p, np.array(lengths)


def collate(instances, pad_token_id, model_max_length):
    """DataCollatorForSupervisedDataset (vis_zephyr/train/train.py:657-707): returns ids, labels, mask."""
    B = len(instances)
    L = max(len(x["input_ids"]) for x in instances)
    ids = np.full((B, L), pad_token_id, np.int64)
    labels = np.full((B, L), IGNORE_INDEX, np.int64)
    for b, x in enumerate(instances):
        n = len(x["input_ids"])
        ids[b, :n] = x["input_ids"]
        labels[b, :n] = x["labels"]
    ids, labels = ids[:, :model_max_length], labels[:, :model_max_length]
    return ids, labels, ids != pad_token_id


def tokenizer_image_token(prompt, tokenizer, image_token_index=IMAGE_TOKEN_INDEX):
    """vis_zephyr/model/mm_utils.py:91-128, followed step by step (separator insertion, offset trick)."""
    chunks = [tokenizer(c).input_ids for c in prompt.split("<image>")]
    out, offset = [], 0
    if len(chunks) > 0 and len(chunks[0]) > 0 and chunks[0][0] == tokenizer.bos_token_id:
        offset = 1
        out.append(chunks[0][0])
    seps = [[image_token_index] * (offset + 1)] * len(chunks)
    inter = [e for pair in zip(chunks, seps) for e in pair][:-1]
    for x in inter:
        out.extend(x[offset:])
    return out
