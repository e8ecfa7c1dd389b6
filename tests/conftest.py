import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200) device; run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def seeded_weights():
    """fp32, bf16-representable weights shared by the oracle and the CUDA path (oracle/weights.py)."""
    from oracle import weights
    return dict(clip=weights.clip_state_dict(0), qf=weights.qformer_state_dict(1), embed=weights.embed_table(2),
                newline=weights.image_newline(3))


@pytest.fixture(scope="session")
def vision_path(seeded_weights):
    """The product path on cuda:0 loaded with the seeded weights."""
    import torch
    import vision_zephyr_b200  # noqa: F401
    from vision_zephyr_b200.runtime import VisionEmbeddingPath
    p = VisionEmbeddingPath(device="cuda")
    p.load_weights(seeded_weights["clip"], seeded_weights["qf"], seeded_weights["embed"], seeded_weights["newline"])
    torch.cuda.synchronize()
    return p
