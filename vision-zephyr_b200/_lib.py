"""ctypes binding of libvz_b200.so (the C ABI declared in include/vz_b200.h).

There is no CPU fallback: if the CUDA library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvz_b200.so")

VIT_LAYERS = 24
QF_BLOCKS = 8


class VzError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("W", C.c_void_p), ("out", C.c_void_p), ("bias", C.c_void_p),
        ("residual", C.c_void_p),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("lda", C.c_int), ("ldw", C.c_int), ("ldo", C.c_int), ("ldr", C.c_int),
        ("act", C.c_int), ("row_mode", C.c_int), ("rows_per", C.c_int), ("force_simple", C.c_int),
        ("batch", C.c_int), ("out_f32", C.c_int),
        ("a_bstride", C.c_longlong), ("w_bstride", C.c_longlong), ("o_bstride", C.c_longlong),
        ("r_bstride", C.c_longlong), ("bias_bstride", C.c_longlong),
        ("ln_stats", C.c_void_p), ("ln_colsum", C.c_void_p), ("ln_np", C.c_int), ("ln_eps", C.c_float),
        ("stats_out", C.c_void_p), ("stats_np", C.c_int),
        ("sk_ws", C.c_void_p), ("sk_ws_bytes", C.c_size_t), ("w_is_kn", C.c_int), ("ln_rms", C.c_int),
    ]


class ImageDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("layers", C.c_void_p), ("W", C.c_int), ("H", C.c_int),
                ("prim_begin", C.c_int), ("prim_count", C.c_int),
                ("pad_x", C.c_int), ("pad_y", C.c_int), ("bg", C.c_uint32), ("reserved", C.c_int)]


class Prim(C.Structure):
    _fields_ = [("type", C.c_int), ("layer", C.c_int), ("x0", C.c_int), ("y0", C.c_int),
                ("x1", C.c_int), ("y1", C.c_int), ("width", C.c_int), ("rgba", C.c_uint32)]


class TileDesc(C.Structure):
    _fields_ = [("image", C.c_int), ("out_w", C.c_int), ("out_h", C.c_int), ("off_x", C.c_int),
                ("off_y", C.c_int), ("tile_x", C.c_int), ("tile_y", C.c_int), ("tab_h", C.c_int),
                ("tab_v", C.c_int), ("hview", C.c_int), ("tab_v_dp", C.c_int), ("reserved", C.c_int)]


class HViewDesc(C.Structure):
    _fields_ = [("image", C.c_int), ("tab_h", C.c_int), ("out_w", C.c_int), ("rows", C.c_int),
                ("offset", C.c_longlong), ("tab_h_dp", C.c_int), ("reserved", C.c_int)]


class VitLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "w_qkv", "b_qkv", "s_qkv", "w_o", "b_o", "w_fc1", "b_fc1", "s_fc1", "w_fc2", "b_fc2")]


class VitWeights(C.Structure):
    _fields_ = [("patch_w", C.c_void_p), ("class_emb", C.c_void_p), ("pos_emb", C.c_void_p),
                ("pre_ln_g", C.c_void_p), ("pre_ln_b", C.c_void_p),
                ("layers", VitLayer * VIT_LAYERS)]


class QfBlock(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "n1_g", "n1_b", "sa_in_w", "sa_in_b", "s_sa_in", "sa_out_w", "sa_out_b",
        "ca_q_w", "ca_q_b", "s_ca_q", "ca_kT_w", "ca_v_w", "ca_in_b", "ca_out_w", "ca_out_b",
        "ffn1_w", "ffn1_b", "s_ffn1", "ffn2_w", "ffn2_b")]


class QfWeights(C.Structure):
    _fields_ = [("learned_queries", C.c_void_p), ("pre_g", C.c_void_p), ("pre_b", C.c_void_p),
                ("norm_g", C.c_void_p), ("norm_b", C.c_void_p), ("blocks", QfBlock * QF_BLOCKS)]


class SlotDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "row_base", "n_rows", "merge", "hw", "h", "w", "n_w", "n_h", "y0", "y1", "x0", "x1")]


# enums of include/vz_b200.h
ACT_NONE, ACT_QUICK_GELU, ACT_GELU_ERF = 0, 1, 2
ROWS_PLAIN, ROWS_PATCH_EMBED, ROWS_RES_MOD = 0, 1, 2
PRIM_LAYER, PRIM_RECT = 0, 1
OUT_PATCHES_BF16, OUT_CHW_F32 = 0, 1
MERGE_FLAT, MERGE_SPATIAL, MERGE_SPATIAL_UNPAD, MERGE_SINGLE_NEWLINE = 0, 1, 2, 3
PROF_COUNT = 13

# every symbol include/vz_b200.h declares: name -> (restype, argtypes)
_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t
SYMBOLS = {
    "vz_status_string": (C.c_char_p, [_i]),
    "vz_last_cuda_error": (_i, []),
    "vz_version": (_i, []),
    "vz_gemm_bf16": (_i, [C.POINTER(GemmArgs), _vp]),
    "vz_gemm_stats_partials": (_i, [_i, _i]),
    "vz_gemm_sk_workspace_bytes": (_sz, []),
    "vz_kernel_launches": (C.c_longlong, []),
    "vz_gemm_profile": (_i, [_i]),
    "vz_gemm_profile_read": (_i, [C.POINTER(C.c_longlong), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "vz_profile": (_i, [_i]),
    "vz_profile_read": (_i, [_i, C.POINTER(C.c_longlong), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "vz_profile_tag_name": (C.c_char_p, [_i]),
    "vz_layernorm_bf16": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, C.c_float, _vp]),
    "vz_preprocess": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _vp]),
    "vz_preprocess2": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, C.c_longlong, _i, _i, _i, _i, _vp]),
    "vz_preprocess3": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, C.c_longlong, _i, _i, _i, _i, _i, _vp]),
    "vz_preprocess_identity": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp]),
    "vz_patchify": (_i, [_vp, _i, _i, _vp, _vp]),
    "vz_vit_workspace_bytes": (_sz, [_i]),
    "vz_vit_workspace_bytes_ex": (_sz, [_i, _i]),
    "vz_vit_attention": (_i, [_vp, _vp, _i, _i, _vp]),
    "vz_vit_forward": (_i, [C.POINTER(VitWeights), _vp, _i, _vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "vz_qformer_workspace_bytes": (_sz, [_i, _i, _i]),
    "vz_qformer_forward": (_i, [C.POINTER(QfWeights), _vp, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _i,
                                _vp, _sz, _i, _vp]),
    "vz_splice_plan": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vz_text_gather": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "vz_splice_scatter": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _vp, _i, _i, _vp, _i, _vp, _i, _vp, _vp,
                               _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "vz_splice_scatter_rms": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _vp, _i, _i, _vp, _i, _vp, _i, _vp, _vp,
                                   _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vz_collate": (_i, [_vp, _vp, _vp, _i, _i, C.c_int64, _vp, _vp, _vp, _vp]),
    "vz_merge_rows": (_i, [_vp, _i, _vp, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "vz_rope_table": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "vz_rope_apply": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "vz_row_stats": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "vz_rows_move": (_i, [_vp, C.c_longlong, _vp, C.c_longlong, _vp, _i, _i, _i, _vp]),
    "vz_attn_causal_items": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "vz_attn_causal": (_i, [_vp, _i, _i, _vp, _i, _vp, _i, _i, _i, _i, C.c_float, C.c_double, _vp]),
}

_lock = threading.Lock()
_lib = None


def load():
    """Load the shared library once; raise VzError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise VzError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C vision-zephyr_b200/csrc`). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int, what: str):
    if status != 0:
        lib = load()
        msg = lib.vz_status_string(status).decode()
        extra = ""
        if status == -3:
            extra = f" (cuda error {lib.vz_last_cuda_error()})"
        raise VzError(f"{what}: {msg}{extra}")


def h2d(host, device, dtype=None):
    """Small host array / CPU tensor -> device WITHOUT a stream sync.  A copy from pageable memory makes the driver
    synchronise the stream first (the host then cannot run ahead of the GPU, and the GPU idles while the next
    launches are being prepared); from a pinned staging buffer the copy is just another stream-ordered operation.
    PyTorch's caching host allocator keeps the staging buffer alive until the copy has run."""
    import numpy as np
    import torch
    t = torch.from_numpy(np.ascontiguousarray(host)) if isinstance(host, np.ndarray) else host
    if dtype is not None:
        t = t.to(dtype)
    dev = torch.device(device)
    if dev.type != "cuda":
        return t.to(dev)
    return t.contiguous().pin_memory().to(dev, non_blocking=True)


def ptr(t):
    """device pointer of a torch tensor (or None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def profile_read():
    """{kernel family: (launches, total ms, total algorithmic work)} of everything recorded since vz_profile(1)."""
    lib = load()
    out = {}
    for tag in range(PROF_COUNT):
        n, ms, wk = C.c_longlong(0), C.c_double(0), C.c_double(0)
        check(lib.vz_profile_read(tag, C.byref(n), C.byref(ms), C.byref(wk)), "vz_profile_read")
        if n.value:
            out[lib.vz_profile_tag_name(tag).decode()] = (int(n.value), float(ms.value), float(wk.value))
    return out
