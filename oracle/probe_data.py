"""ORACLE (test infrastructure, not product code): restatement of the reference's offline feature cacher, one frame at a
time, exactly as /root/reference/primitive_probing/generate_data/thor_image_features.py:69-140 processes a scene:

  * labels (:69-87, :115-127): per target object a mask `all(semantic_frame == colour, axis=-1)` (all-False when the object has
    no colour in this frame), presence = mask.sum() > 0, localisation = presence inside each cell of a 3 x 3 grid whose
    bounds are int(i * H / 3) (row-major cells -> tensor [9, 52], int64);
  * CLIP features (:57-67, :109-113): clip_preprocess(frame) = Resize(224, bicubic) -> CenterCrop(224) -> ToTensor ->
    Normalize(CLIP mean / std) [UPSTREAM clip/clip.py::_transform]; trunk with attnpool replaced by Identity ->
    'clip_conv' [2048,7,7]; attnpool on the trunk output -> 'clip_attnpool' [1024]; AdaptiveAvgPool2d(1) -> 'clip_avgpool'
    [2048]; everything `.float()[0].cpu()`.
  * `free_space` is copied from the point (:137).  The 'imagenet_*' entries (:101-105) come from torchvision's own ResNet-50
    cut after layer4 (oracle/imagenet_resnet.py) on `resnet_preprocess(frame)` (:36-44: same geometry, ImageNet mean / std).
  * `reachable_features` restates reachable_image_features.py:77-98 (three pooled embeddings per PNG).

Parity unpinned: the reference has no test or fixture for this file.  Only tests/ may import this module.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch
from PIL import Image

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def class_mask(semantic_frame: np.ndarray, class_color) -> np.ndarray:           # thor_image_features.py:69-73
    if class_color is None:
        return np.zeros(semantic_frame.shape[:2], dtype=bool)
    return np.all(semantic_frame == class_color, axis=-1)


def obj_presence(class_masks: np.ndarray) -> np.ndarray:                         # :75-76
    return class_masks.sum(axis=(1, 2)) > 0


def grid_bboxes(image_shape, grid_sizes):                                        # :78-87
    for i in range(grid_sizes[0]):
        for j in range(grid_sizes[1]):
            yield (int(i * image_shape[0] / grid_sizes[0]), int((i + 1) * image_shape[0] / grid_sizes[0]),
                   int(j * image_shape[1] / grid_sizes[1]), int((j + 1) * image_shape[1] / grid_sizes[1]))


def clip_preprocess(frame: np.ndarray) -> torch.Tensor:
    """uint8 HWC -> normalised float32 CHW 224 x 224 (clip.load's transform on a PIL image)."""
    img = Image.fromarray(frame)
    w, h = img.size
    s = 224 / min(w, h)
    if (w, h) != (224, 224):
        img = img.resize((max(224, round(w * s)), max(224, round(h * s))), Image.BICUBIC)
        w, h = img.size
        l, t = int(round((w - 224) / 2.0)), int(round((h - 224) / 2.0))
        img = img.crop((l, t, l + 224, t + 224))
    x = torch.from_numpy(np.asarray(img.convert("RGB"), dtype=np.uint8).copy()).permute(2, 0, 1).float() / 255.0
    return (x - torch.tensor(CLIP_MEAN).view(3, 1, 1)) / torch.tensor(CLIP_STD).view(3, 1, 1)


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def resnet_preprocess(frame: np.ndarray) -> torch.Tensor:
    """thor_image_features.py:36-44: the same Resize / CenterCrop as clip_preprocess, ImageNet normalisation."""
    x = clip_preprocess(frame) * torch.tensor(CLIP_STD).view(3, 1, 1) + torch.tensor(CLIP_MEAN).view(3, 1, 1)     # back to [0, 1]
    return (x - torch.tensor(IMAGENET_MEAN).view(3, 1, 1)) / torch.tensor(IMAGENET_STD).view(3, 1, 1)


@torch.no_grad()
def reachable_features(images: Dict[str, np.ndarray], visual, resnet_trunk=None) -> Dict[str, Dict[str, torch.Tensor]]:
    """reachable_image_features.py:77-98, one image at a time."""
    out = {}
    for name, frame in images.items():
        d = {}
        if resnet_trunk is not None:
            d["imagenet_avgpool"] = resnet_trunk(resnet_preprocess(frame).unsqueeze(0)).mean(dim=(2, 3))[0].cpu()
        t = visual.trunk(clip_preprocess(frame).unsqueeze(0))
        d["clip_avgpool"] = t.float().mean(dim=(2, 3))[0].cpu()
        d["clip_attnpool"] = visual.attnpool(t).float()[0].cpu()
        out[name] = d
    return out


@torch.no_grad()
def scene_features(points: Sequence[dict], visual, target_objects: Sequence[str], resnet_trunk=None) -> List[Dict[str, torch.Tensor]]:
    """One dict per point, batch 1 like the reference loop (:99-138).  `visual` = oracle.clip_model ModifiedResNet;
    `resnet_trunk` = oracle.imagenet_resnet trunk (adds the imagenet_* keys)."""
    out = []
    for point in points:
        x = clip_preprocess(point["frame"]).unsqueeze(0)
        t = visual.trunk(x)
        masks = np.array([class_mask(point["semantic_frame"], point["object_id_to_color"].get(o, None)) for o in target_objects])
        inet = {}
        if resnet_trunk is not None:
            r = resnet_trunk(resnet_preprocess(point["frame"]).unsqueeze(0))
            inet = {"imagenet_conv": r[0].cpu(), "imagenet_avgpool": r.mean(dim=(2, 3))[0].cpu()}
        out.append({**inet, **{
            "clip_conv": t.float()[0].cpu(),
            "clip_attnpool": visual.attnpool(t).float()[0].cpu(),
            "clip_avgpool": t.float().mean(dim=(2, 3))[0].cpu(),
            "object_presence": torch.tensor(obj_presence(masks), dtype=int),
            "object_localization": torch.tensor(np.array([obj_presence(masks[:, y1:y2, x1:x2])
                                                          for (y1, y2, x1, x2) in grid_bboxes(masks.shape[1:3], (3, 3))]), dtype=int),
            "free_space": point["valid_moves_forward"],
        }})
    return out
