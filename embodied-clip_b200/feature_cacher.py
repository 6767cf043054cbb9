"""Batched offline feature cacher: the drop-in for the per-frame loop of
primitive_probing/generate_data/thor_image_features.py:91-140.

The reference pushes ONE frame at a time through CLIP and pulls five tensors back with `.cpu()` per frame.  Here a scene's
frames are resized on the host (PIL bicubic + centre crop, the same `clip_preprocess` geometry), stacked as raw uint8 NHWC,
and encoded in batches by `ClipRN50Encoder` (normalisation happens in the stem kernel); all three heads come back in one
device->host copy per batch.  The label tensors are computed with vectorised numpy instead of a Python loop per object.

Output layout is the reference's (`thor_{split}.pt` = {scene_name: [ {key: tensor} per point ]}, :129-140) with the keys
'imagenet_conv' [2048,7,7], 'imagenet_avgpool' [2048] (when an ImageNet encoder is given: `TorchvisionResNet50Encoder`, the
reference's `models.resnet50` cut after layer4, :46-49,101-105), 'clip_conv' [2048,7,7], 'clip_attnpool' [1024],
'clip_avgpool' [2048], 'object_presence' int64 [52], 'object_localization' int64 [9,52], 'free_space'.

`cache_reachable` is the twin of primitive_probing/generate_data/reachable_image_features.py:24-25,77-100: a directory of
PNG frames -> `reachable_image_features.pt` = {image_name: {'imagenet_avgpool', 'clip_avgpool', 'clip_attnpool'}}.
"""
from __future__ import annotations

import os
from glob import glob
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .encoder import ClipRN50Encoder, TorchvisionResNet50Encoder

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
    def __init__(self, encoder: ClipRN50Encoder, batch: int = 256, target_objects: Sequence[str] = TARGET_OBJECTS,
                 imagenet_encoder: Optional[TorchvisionResNet50Encoder] = None):
        self.enc, self.batch, self.target_objects = encoder, int(batch), tuple(target_objects)
        self.imagenet = imagenet_encoder
        if self.batch <= 0:
            raise ValueError("batch must be positive")
        if self.imagenet is not None and self.imagenet.device != self.enc.device:
            raise ValueError("the CLIP and ImageNet encoders must live on the same device")
        self._pin: Optional[torch.Tensor] = None

    def encode_frames(self, frames: Sequence[np.ndarray]) -> Dict[str, torch.Tensor]:
        """Raw uint8 frames (any resolution) -> {'clip_conv' [N,2048,7,7], 'clip_attnpool' [N,1024], 'clip_avgpool' [N,2048]
        (+ 'imagenet_conv' [N,2048,7,7], 'imagenet_avgpool' [N,2048] with an ImageNet encoder)} on CPU.  Both encoders read the
        SAME resized uint8 batch: `resnet_preprocess` and `clip_preprocess` share the geometry (Resize 224 bicubic, CenterCrop
        224) and differ only in the mean / std, which each stem applies on the device."""
        n = len(frames)
        R = self.enc.cfg["input_resolution"]
        outs = {"clip_conv": torch.empty(n, self.enc.embed, self.enc.fres, self.enc.fres),
                "clip_attnpool": torch.empty(n, self.enc.cfg["output_dim"]), "clip_avgpool": torch.empty(n, self.enc.embed)}
        if self.imagenet is not None:
            outs["imagenet_conv"] = torch.empty(n, self.imagenet.embed, self.imagenet.fres, self.imagenet.fres)
            outs["imagenet_avgpool"] = torch.empty(n, self.imagenet.embed)
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
            if self.imagenet is not None:
                o = self.imagenet.forward(dev, ("trunk", "avgpool"))
                outs["imagenet_conv"][i0:i0 + b] = o["trunk"].cpu()
                outs["imagenet_avgpool"][i0:i0 + b] = o["avgpool"].cpu()
        return outs

    def scene_features(self, points: Sequence[dict]) -> List[Dict[str, torch.Tensor]]:
        """The list the reference appends to `features[scene_name]` (thor_image_features.py:99-138), batched."""
        feats = self.encode_frames([p["frame"] for p in points])
        out = []
        for i, p in enumerate(points):
            pres, loc = presence_labels(p["semantic_frame"], p["object_id_to_color"], self.target_objects)
            d = {}
            if self.imagenet is not None:                       # key order of thor_image_features.py:129-138
                d["imagenet_conv"] = feats["imagenet_conv"][i].clone()
                d["imagenet_avgpool"] = feats["imagenet_avgpool"][i].clone()
            d.update({"clip_conv": feats["clip_conv"][i].clone(), "clip_attnpool": feats["clip_attnpool"][i].clone(),
                      "clip_avgpool": feats["clip_avgpool"][i].clone(), "object_presence": pres,
                      "object_localization": loc, "free_space": p["valid_moves_forward"]})
            out.append(d)
        return out

    def reachable_features(self, images: Dict[str, np.ndarray]) -> Dict[str, Dict[str, torch.Tensor]]:
        """{image_name: uint8 HWC frame} -> {image_name: {'imagenet_avgpool', 'clip_avgpool', 'clip_attnpool'}}
        (reachable_image_features.py:77-98), batched."""
        names = list(images)
        feats = self.encode_frames([images[k] for k in names])
        out = {}
        for i, k in enumerate(names):
            d = {}
            if self.imagenet is not None:
                d["imagenet_avgpool"] = feats["imagenet_avgpool"][i].clone()
            d["clip_avgpool"] = feats["clip_avgpool"][i].clone()
            d["clip_attnpool"] = feats["clip_attnpool"][i].clone()
            out[k] = d
        return out

    def cache_reachable(self, data_dir: str, output_dir: str) -> str:
        """{data_dir}/*.png -> {output_dir}/reachable_image_features.pt (reachable_image_features.py:24-25,100)."""
        from PIL import Image
        images = {}
        for path in sorted(glob(os.path.join(data_dir, "*.png"))):
            images[os.path.splitext(os.path.basename(path))[0]] = np.asarray(Image.open(path).convert("RGB"), dtype=np.uint8)
        os.makedirs(output_dir, exist_ok=True)
        out = os.path.join(output_dir, "reachable_image_features.pt")
        torch.save(self.reachable_features(images), out)
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
    ap.add_argument("--imagenet_weights", type=str, default=os.environ.get("EMBCLIP_IMAGENET_WEIGHTS"),
                    help="torchvision resnet50 state dict (.pth): adds the imagenet_conv / imagenet_avgpool keys")
    ap.add_argument("--reachable", action="store_true",
                    help="reachable_image_features.py mode: --data_dir holds *.png, writes reachable_image_features.pt")
    ap.add_argument("--batch", type=int, default=256)
    args = ap.parse_args(argv)
    if not args.weights:
        raise SystemExit("feature_cacher: pass --weights (or set $EMBCLIP_CLIP_WEIGHTS) to a CLIP RN50 state dict")
    load = lambda path: (lambda sd: sd.state_dict() if hasattr(sd, "state_dict") else sd)(torch.load(path, map_location="cpu"))
    inet = TorchvisionResNet50Encoder(load(args.imagenet_weights), "cuda:0") if args.imagenet_weights else None
    fc = FeatureCacher(ClipRN50Encoder(load(args.weights), "cuda:0"), args.batch, imagenet_encoder=inet)
    if args.reachable:
        print(fc.cache_reachable(args.data_dir, args.output_dir))
        return
    for split in ("train", "val", "test"):
        print(fc.cache_split(args.data_dir, args.output_dir, split))


if __name__ == "__main__":
    main()
