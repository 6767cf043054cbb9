import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def built_lib():
    """libembclip_b200.so, (re)built when sources are newer.  nvcc cross-compiles without a GPU."""
    from embclip_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def rn50_visual():
    import torch
    from oracle.clip_model import build_rn50, freeze_model, init_synthetic_rn50_visual
    torch.manual_seed(0)
    m = build_rn50().visual
    init_synthetic_rn50_visual(m, seed=1234)
    return freeze_model(m)


def synthetic_frames(batch: int, res: int = 224, seed: int = 0):
    """uint8 frames randint(0,256) -> /255 -> CLIP mean/std, fp32 NHWC."""
    import torch
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, (batch, res, res, 3), generator=g, dtype=torch.uint8)
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073])
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711])
    return (u8.float() / 255.0 - mean) / std
