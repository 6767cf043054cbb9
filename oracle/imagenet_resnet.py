"""ORACLE (test infrastructure, not product code): the reference's ImageNet baseline encoder, built from the very
library the reference calls -- ``torchvision.models.resnet50`` cut after layer4
(/root/reference/primitive_probing/generate_data/thor_image_features.py:46-49; forward :101-105; the same lines in
reachable_image_features.py:49-52,81-85) and frozen as ``freeze_model`` does (:26-33).

PARITY PINNED TO THE LIBRARY: unlike the CLIP towers (whose module lives in an un-vendored git pin), this path's
arithmetic is torchvision's own ``ResNet`` / ``Bottleneck`` (torchvision is importable in this image), so the GPU
kernels are checked against the module the reference itself instantiates, with weights mapped by key name.  The
checkpoint (`pretrained=True`) cannot be downloaded offline: weights are the seeded synthetic set of
``embclip_b200.synthetic.synthetic_torchvision_rn50_state_dict`` (weight *generation* is shared; no arithmetic is).

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this file.
"""
from __future__ import annotations

import torch
import torch.nn as nn

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # resnet_preprocess, thor_image_features.py:36-44
IMAGENET_STD = (0.229, 0.224, 0.225)


def freeze_model(model: nn.Module) -> nn.Module:
    """thor_image_features.py:26-33."""
    for p in model.parameters():
        p.requires_grad = False
    for m in model.modules():
        if "BatchNorm" in type(m).__name__:
            m.momentum = 0.0
    return model.eval()


def build_imagenet_rn50(seed: int = 4321):
    """-> (trunk nn.Sequential [B,3,224,224] -> [B,2048,7,7], pool -> [B,2048], full torchvision state dict)."""
    from torchvision import models
    from embclip_b200.synthetic import synthetic_torchvision_rn50_state_dict
    resnet = models.resnet50(weights=None)
    sd = synthetic_torchvision_rn50_state_dict(seed)
    missing, unexpected = resnet.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("fc.") for k in missing), (missing, unexpected)
    trunk = freeze_model(nn.Sequential(*list(resnet.children())[:-2]))          # :47
    pool = nn.Sequential(nn.AdaptiveAvgPool2d(output_size=(1, 1)), nn.Flatten())  # :51-54
    return trunk, pool, sd


def normalize_imagenet(u8_nhwc: torch.Tensor) -> torch.Tensor:
    """uint8 NHWC -> fp32 NHWC, T.ToTensor() + T.Normalize of resnet_preprocess (:39-43)."""
    mean, std = torch.tensor(IMAGENET_MEAN), torch.tensor(IMAGENET_STD)
    return (u8_nhwc.float() / 255.0 - mean) / std
