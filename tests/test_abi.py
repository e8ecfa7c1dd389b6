"""The C-ABI library builds, loads and exports every symbol include/vz_b200.h declares (no GPU)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vz_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vz_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    import vision_zephyr_b200  # noqa: F401
    from vision_zephyr_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_lib.SYMBOLS) == names, "python binding table and header disagree"


def test_status_strings_and_version():
    from vision_zephyr_b200 import _lib
    lib = _lib.load()
    assert lib.vz_version() == 100
    assert lib.vz_status_string(0) == b"ok"
    assert b"workspace" in lib.vz_status_string(-5)
    assert lib.vz_vit_workspace_bytes(1) > 577 * 1024 * 2 * 25
    assert lib.vz_qformer_workspace_bytes(1, 1, 63) > 3 * 576 * 5120 * 2


def test_bad_arguments_are_rejected_without_a_gpu():
    from vision_zephyr_b200 import _lib
    lib = _lib.load()
    assert lib.vz_gemm_bf16(None, None) == -1
    assert lib.vz_layernorm_bf16(None, 0, None, None, None, 0, 0, 0, 1e-5, None) == -1
    assert lib.vz_splice_plan(None, None, 0, 0, None, 0, 0, None, None, None, None, None, None) == -1


def test_struct_layouts_match_the_header():
    from vision_zephyr_b200 import _lib
    assert ctypes.sizeof(_lib.GemmArgs) == 5 * 8 + 13 * 4 + 4 + 5 * 8 + 2 * 8 + 2 * 4 + 8 + 4 + 4 + 2 * 8 + 8
    assert ctypes.sizeof(_lib.ImageDesc) == 48
    assert ctypes.sizeof(_lib.Prim) == 32
    assert ctypes.sizeof(_lib.TileDesc) == 48
    assert ctypes.sizeof(_lib.HViewDesc) == 32
    assert ctypes.sizeof(_lib.SlotDesc) == 48
    assert ctypes.sizeof(_lib.VitWeights) == 5 * 8 + 24 * 10 * 8
    assert ctypes.sizeof(_lib.QfWeights) == 5 * 8 + 8 * 20 * 8
