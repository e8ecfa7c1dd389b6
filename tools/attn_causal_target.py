"""ncu target: vz_attn_causal alone on the config-5 lengths (8 samples, 8 970 packed rows, 32 / 8 heads of 128)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LENS = [514, 1645, 1166, 2140, 645, 1557, 815, 488]


def main():
    import ctypes as C
    from vision_zephyr_b200 import _lib
    lib = _lib.load()
    nh, nkv, hd = 32, 8, 128
    M = sum(LENS)
    qkv = torch.randn((M, (nh + 2 * nkv) * hd), device="cuda").to(torch.bfloat16)
    lens_h = np.asarray(LENS, dtype=np.int32)
    n = lib.vz_attn_causal_items(lens_h.ctypes.data, len(LENS), nh, None, 0, None)
    items_h = np.empty((n, 4), dtype=np.int32)
    flops = C.c_double(0)
    lib.vz_attn_causal_items(lens_h.ctypes.data, len(LENS), nh, items_h.ctypes.data, n, C.byref(flops))
    items = torch.from_numpy(items_h).cuda()
    out = torch.empty((M, nh * hd), dtype=torch.bfloat16, device="cuda")
    reps = int(os.environ.get("REPS", "3"))
    for _ in range(reps):
        _lib.check(lib.vz_attn_causal(qkv.data_ptr(), qkv.shape[1], M, out.data_ptr(), nh * hd, items.data_ptr(), n, nh,
                                      nkv, hd, hd ** -0.5, flops.value, _lib.stream_ptr()), "vz_attn_causal")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.vz_attn_causal(qkv.data_ptr(), qkv.shape[1], M, out.data_ptr(), nh * hd, items.data_ptr(), n, nh, nkv, hd,
                           hd ** -0.5, flops.value, _lib.stream_ptr())
    e1.record()
    torch.cuda.synchronize()
    print(f"vz_attn_causal: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch, {n} items")


if __name__ == "__main__":
    main()
