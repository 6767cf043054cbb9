"""Batched offline feature cacher: the drop-in for the per-frame loop of
primitive_probing/generate_data/thor_image_features.py:91-140.

The reference pushes ONE frame at a time through CLIP and pulls five tensors back with `.cpu()` per frame.  Here a scene's
frames are resized on the host (PIL bicubic + centre crop, the same `clip_preprocess` geometry), stacked as raw uint8 NHWC,
and encoded in batches by `ClipRN50Encoder` (normalisation happens in the stem kernel); all three heads come back in one
device->host copy per batch.  The label tensors are computed with vectorised numpy instead of a Python loop per object.

Output layout is the reference's (`thor_{split}.pt` = {scene_name: [ {key: tensor} per point ]}, :129-140) with the keys
'clip_conv' [2048,7,7], 'clip_attnpool' [1024], 'clip_avgpool' [2048], 'object_presence' int64 [52],
'object_localization' int64 [9,52], 'free_space'.  The 'imagenet_conv' / 'imagenet_avgpool' keys are NOT produced: the
torchvision ImageNet ResNet-50 is outside this library's scope (section 8f item 4); `primitive_probing/train.py` only reads
the key named by its --embedding_type, so the CLIP probes run unchanged.
"""
from __future__ import annotations

import os
from glob import glob
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .encoder import ClipRN50Encoder

# the 52 iTHOR object types probed by the reference (primitive_probing/constants.py:1) -- label order is part of the file format
TARGET_OBJECTS = (
    "AlarmClock Apple ArmChair Bathtub Bed Bowl Box Bread Cabinet Chair CoffeeMachine CoffeeTable Cup DeskLamp DiningTable Egg "
    "Faucet FloorLamp Fridge GarbageCan HandTowel HousePlant Laptop Lettuce Microwave Mug Painting Pan Pillow Plate Plunger Pot "
    "Potato RemoteControl ScrubBrush SideTable Sink SinkBasin SoapBar SoapBottle Sofa Spatula Spoon SprayBottle Statue "
    "StoveBurner Television Toaster Toilet ToiletPaper Tomato Towel").split()


def resize_center_crop(frame: np.ndarray, size: int = 224) -> np.ndarray:
    """uint8 HWC -> uint8 [size,size,3]: Resize(size, bicubic) on the shorter side, then CenterCrop(size)."""
    if frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != 3:
        raise ValueError(f"frame must be uint8 [H,W,3], got {frame.dtype} {frame.shape}")
    h, w = frame.shape[:2]
    if (h, w) == (size, size):
        return frame
    from PIL import Image
    s = size / min(w, h)
    img = Image.fromarray(frame).resize((max(size, round(w * s)), max(size, round(h * s))), Image.BICUBIC)
    w, h = img.size
    l, t = int(round((w - size) / 2.0)), int(round((h - size) / 2.0))
    return np.asarray(img.crop((l, t, l + size, t + size)), dtype=np.uint8)


def presence_labels(semantic_frame: np.ndarray, object_id_to_color: Dict[str, Sequence[int]],
                    target_objects: Sequence[str] = TARGET_OBJECTS, grid=(3, 3)):
    """(object_presence int64 [n_obj], object_localization int64 [grid cells, n_obj]) of one frame.
    An object is present where every channel of the semantic frame equals its colour (thor_image_features.py:69-76); cell
    bounds are int(i * H / 3) (:78-87).  One pass: the frame is reduced to a per-pixel colour key, then compared per object."""
    sem = np.asarray(semantic_frame)
    H, W = sem.shape[:2]
    key = np.zeros((H, W), dtype=np.int64)
    for c in range(sem.shape[2]):
        key = key * 256 + sem[..., c].astype(np.int64)
    ys = [int(i * H / grid[0]) for i in range(grid[0] + 1)]
    xs = [int(j * W / grid[1]) for j in range(grid[1] + 1)]
    pres = np.zeros(len(target_objects), dtype=np.int64)
    loc = np.zeros((grid[0] * grid[1], len(target_objects)), dtype=np.int64)
    for k, o in enumerate(target_objects):
        col = object_id_to_color.get(o, None)
        if col is None:
            continue
        col = np.asarray(col).reshape(-1)
        if col.shape[0] != sem.shape[2] or np.any(col < 0) or np.any(col > 255):
            continue                                                    # cannot equal any uint8 pixel
        ck = 0
        for c in range(sem.shape[2]):
            ck = ck * 256 + int(col[c])
        m = key == ck
        if not m.any():
            continue
        pres[k] = 1
        for i in range(grid[0]):
            for j in range(grid[1]):
                loc[i * grid[1] + j, k] = int(m[ys[i]:ys[i + 1], xs[j]:xs[j + 1]].any())
    return torch.from_numpy(pres), torch.from_numpy(loc)


class FeatureCacher:
    def __init__(self, encoder: ClipRN50Encoder, batch: int = 256, target_objects: Sequence[str] = TARGET_OBJECTS):
        self.enc, self.batch, self.target_objects = encoder, int(batch), tuple(target_objects)
        if self.batch <= 0:
            raise ValueError("batch must be positive")
        self._pin: Optional[torch.Tensor] = None

    def encode_frames(self, frames: Sequence[np.ndarray]) -> Dict[str, torch.Tensor]:
        """Raw uint8 frames (any resolution) -> {'clip_conv' [N,2048,7,7], 'clip_attnpool' [N,1024], 'clip_avgpool' [N,2048]} on CPU."""
        n = len(frames)
        R = self.enc.cfg["input_resolution"]
        outs = {"clip_conv": torch.empty(n, self.enc.embed, self.enc.fres, self.enc.fres),
                "clip_attnpool": torch.empty(n, self.enc.cfg["output_dim"]), "clip_avgpool": torch.empty(n, self.enc.embed)}
        if self._pin is None:
            self._pin = torch.empty(self.batch, R, R, 3, dtype=torch.uint8).pin_memory()
        for i0 in range(0, n, self.batch):
            b = min(self.batch, n - i0)
            for j in range(b):
                self._pin[j] = torch.from_numpy(np.ascontiguousarray(resize_center_crop(frames[i0 + j], R)))
            dev = self._pin[:b].to(self.enc.device, non_blocking=True)
            o = self.enc.forward(dev, ("trunk", "avgpool", "attnpool"))
            outs["clip_conv"][i0:i0 + b] = o["trunk"].cpu()
            outs["clip_attnpool"][i0:i0 + b] = o["attnpool"].cpu()
            outs["clip_avgpool"][i0:i0 + b] = o["avgpool"].cpu()
        return outs

    def scene_features(self, points: Sequence[dict]) -> List[Dict[str, torch.Tensor]]:
        """The list the reference appends to `features[scene_name]` (thor_image_features.py:99-138), batched."""
        feats = self.encode_frames([p["frame"] for p in points])
        out = []
        for i, p in enumerate(points):
            pres, loc = presence_labels(p["semantic_frame"], p["object_id_to_color"], self.target_objects)
            out.append({"clip_conv": feats["clip_conv"][i].clone(), "clip_attnpool": feats["clip_attnpool"][i].clone(),
                        "clip_avgpool": feats["clip_avgpool"][i].clone(), "object_presence": pres,
                        "object_localization": loc, "free_space": p["valid_moves_forward"]})
        return out

    def cache_split(self, data_dir: str, output_dir: str, split: str) -> str:
        """{data_dir}/{split}/*.npy (written by thor_frames.py) -> {output_dir}/thor_{split}.pt"""
        features = {}
        for scene in sorted(glob(os.path.join(data_dir, split, "*.npy"))):
            name = os.path.splitext(os.path.basename(scene))[0]
            features[name] = self.scene_features(list(np.load(scene, allow_pickle=True)))
        os.makedirs(output_dir, exist_ok=True)
        path = os.path.join(output_dir, f"thor_{split}.pt")
        torch.save(features, path)
        return path


def main(argv=None) -> None:
    import argparse
    ap = argparse.ArgumentParser(description="Batched CLIP-RN50 feature cacher (same arguments as thor_image_features.py:16-23)")
    ap.add_argument("--data_dir", type=str, default="data/ithor_scenes")
    ap.add_argument("--output_dir", type=str, default="data")
    ap.add_argument("--weights", type=str, default=os.environ.get("EMBCLIP_CLIP_WEIGHTS"),
                    help="CLIP RN50 state dict (.pt); required -- there is no download path offline")
    ap.add_argument("--batch", type=int, default=256)
    args = ap.parse_args(argv)
    if not args.weights:
        raise SystemExit("feature_cacher: pass --weights (or set $EMBCLIP_CLIP_WEIGHTS) to a CLIP RN50 state dict")
    sd = torch.load(args.weights, map_location="cpu")
    sd = sd.state_dict() if hasattr(sd, "state_dict") else sd
    fc = FeatureCacher(ClipRN50Encoder(sd, "cuda:0"), args.batch)
    for split in ("train", "val", "test"):
        print(fc.cache_split(args.data_dir, args.output_dir, split))


if __name__ == "__main__":
    main()
