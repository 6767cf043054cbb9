"""Batched feature cacher (embclip_b200/feature_cacher.py) against the restated reference loop (oracle/probe_data.py,
thor_image_features.py:69-140).  Labels are integer work: bit-exact.  Features: rel-L2 <= 1e-3 vs the fp32 oracle (the
north-star tolerance), on 300 x 300 frames like thor_frames.py:33-34 renders, so the bicubic resize is on the path."""
import numpy as np
import pytest
import torch


def _scene(n, seed, res=300, n_colors=70):
    """Synthetic thor_frames.py points: smooth-ish RGB frame, blocky semantic frame, colour table for a subset of the objects."""
    from embclip_b200.feature_cacher import TARGET_OBJECTS
    rng = np.random.RandomState(seed)
    palette = rng.randint(0, 256, size=(n_colors, 3)).astype(np.uint8)
    points = []
    for _ in range(n):
        low = rng.randint(0, 256, size=(res // 20, res // 20, 3)).astype(np.uint8)
        frame = np.kron(low, np.ones((20, 20, 1), dtype=np.uint8))
        frame = np.clip(frame.astype(np.int32) + rng.randint(-8, 9, size=frame.shape), 0, 255).astype(np.uint8)
        ids = rng.randint(0, n_colors, size=(res // 25, res // 25))
        sem = palette[np.kron(ids, np.ones((25, 25), dtype=np.int64))]
        objs = rng.choice(len(TARGET_OBJECTS), size=30, replace=False)
        table = {TARGET_OBJECTS[o]: tuple(int(v) for v in palette[rng.randint(0, n_colors)]) for o in objs}
        table["NotATarget|1"] = (1, 2, 3)
        points.append({"frame": frame, "semantic_frame": sem, "object_id_to_color": table, "valid_moves_forward": int(rng.randint(0, 11))})
    return points


def test_target_objects_match_reference_order():
    from embclip_b200.feature_cacher import TARGET_OBJECTS
    assert len(TARGET_OBJECTS) == 52 and TARGET_OBJECTS[0] == "AlarmClock" and TARGET_OBJECTS[-1] == "Towel"
    assert list(TARGET_OBJECTS) == sorted(TARGET_OBJECTS)          # the reference list is alphabetical (constants.py:1)


@pytest.mark.parametrize("res", [300, 224, 97])
def test_labels_bit_exact_vs_oracle(res):
    from embclip_b200.feature_cacher import TARGET_OBJECTS, presence_labels
    from oracle.probe_data import class_mask, grid_bboxes, obj_presence
    for p in _scene(6, seed=res, res=res if res % 25 == 0 else 300):
        sem = p["semantic_frame"][:res, :res]
        pres, loc = presence_labels(sem, p["object_id_to_color"])
        masks = np.array([class_mask(sem, p["object_id_to_color"].get(o, None)) for o in TARGET_OBJECTS])
        ref_p = torch.tensor(obj_presence(masks), dtype=int)
        ref_l = torch.tensor(np.array([obj_presence(masks[:, a:b, c:d]) for (a, b, c, d) in grid_bboxes(masks.shape[1:3], (3, 3))]), dtype=int)
        assert pres.dtype == ref_p.dtype and loc.dtype == ref_l.dtype and loc.shape == (9, 52)
        assert torch.equal(pres, ref_p) and torch.equal(loc, ref_l)
        assert pres.sum() > 0                                        # the fixture is not vacuous


def test_labels_edge_cases():
    from embclip_b200.feature_cacher import presence_labels
    sem = np.zeros((30, 30, 3), dtype=np.uint8)
    pres, loc = presence_labels(sem, {})                             # no colours at all
    assert pres.sum() == 0 and loc.sum() == 0
    pres, loc = presence_labels(sem, {"Apple": (0, 0, 0), "Bed": (300, 0, 0), "Bowl": None})
    assert pres[1] == 1 and loc[:, 1].all() and pres.sum() == 1      # colour outside uint8 can never match


def test_resize_matches_oracle_preprocess():
    from embclip_b200.feature_cacher import resize_center_crop
    from oracle.probe_data import CLIP_MEAN, CLIP_STD, clip_preprocess
    for p in _scene(2, seed=3) + [{"frame": np.random.RandomState(0).randint(0, 256, (240, 320, 3)).astype(np.uint8)}]:
        u8 = resize_center_crop(p["frame"])
        assert u8.shape == (224, 224, 3) and u8.dtype == np.uint8
        x = (torch.from_numpy(u8.copy()).permute(2, 0, 1).float() / 255 - torch.tensor(CLIP_MEAN).view(3, 1, 1)) / torch.tensor(CLIP_STD).view(3, 1, 1)
        assert torch.equal(x, clip_preprocess(p["frame"]))
    with pytest.raises(ValueError):
        resize_center_crop(np.zeros((10, 10), dtype=np.uint8))


@pytest.mark.gpu
def test_scene_features_vs_oracle(built_lib, rn50_visual, tmp_path):
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from embclip_b200.encoder import ClipRN50Encoder
    from embclip_b200.feature_cacher import FeatureCacher
    from oracle.probe_data import scene_features
    from embclip_b200.feature_cacher import TARGET_OBJECTS
    points = _scene(5, seed=11)
    fc = FeatureCacher(ClipRN50Encoder(rn50_visual.state_dict(), "cuda:0"), batch=3)     # ragged last batch
    ours = fc.scene_features(points)
    ref = scene_features(points, rn50_visual, TARGET_OBJECTS)
    assert len(ours) == len(ref) == 5
    for o, r in zip(ours, ref):
        assert set(o) == set(r)
        for k in ("clip_conv", "clip_attnpool", "clip_avgpool"):
            assert o[k].shape == r[k].shape and o[k].dtype == torch.float32 and o[k].device.type == "cpu"
            err = ((o[k] - r[k]).norm() / r[k].norm()).item()
            assert err <= 1e-3, (k, err)
        assert torch.equal(o["object_presence"], r["object_presence"]) and torch.equal(o["object_localization"], r["object_localization"])
        assert o["free_space"] == r["free_space"]
    # file layout: {scene: [dict per point]} in thor_{split}.pt (thor_image_features.py:129-140)
    d = tmp_path / "scenes" / "val"
    d.mkdir(parents=True)
    np.save(d / "FloorPlan21.npy", np.array(points, dtype=object), allow_pickle=True)
    path = fc.cache_split(str(tmp_path / "scenes"), str(tmp_path / "out"), "val")
    saved = torch.load(path)
    assert path.endswith("thor_val.pt") and list(saved) == ["FloorPlan21"] and len(saved["FloorPlan21"]) == 5
    assert torch.equal(saved["FloorPlan21"][2]["clip_attnpool"], ours[2]["clip_attnpool"])


@pytest.mark.gpu
def test_imagenet_keys_and_reachable_twin(built_lib, rn50_visual, tmp_path):
    """thor_{split}.pt with the imagenet_conv / imagenet_avgpool keys (thor_image_features.py:101-105,129-131) and the
    reachable_image_features.py twin (PNG directory -> {image: 3 pooled embeddings}), against the restated loops driven by
    torchvision's own ResNet-50."""
    from PIL import Image
    from embclip_b200.encoder import ClipRN50Encoder, TorchvisionResNet50Encoder
    from embclip_b200.feature_cacher import FeatureCacher, TARGET_OBJECTS
    from oracle.imagenet_resnet import build_imagenet_rn50
    from oracle.probe_data import reachable_features, scene_features
    trunk, _, sd = build_imagenet_rn50()
    fc = FeatureCacher(ClipRN50Encoder(rn50_visual.state_dict(), "cuda:0"), batch=4, imagenet_encoder=TorchvisionResNet50Encoder(sd, "cuda:0"))
    points = _scene(3, seed=21)
    ours = fc.scene_features(points)
    ref = scene_features(points, rn50_visual, TARGET_OBJECTS, resnet_trunk=trunk)
    assert list(ours[0]) == list(ref[0]) == ["imagenet_conv", "imagenet_avgpool", "clip_conv", "clip_attnpool", "clip_avgpool",
                                             "object_presence", "object_localization", "free_space"]
    rl = lambda a, b: ((a - b).norm() / b.norm()).item()
    for o, r in zip(ours, ref):
        for k in ("imagenet_conv", "imagenet_avgpool", "clip_conv", "clip_attnpool", "clip_avgpool"):
            assert o[k].shape == r[k].shape and rl(o[k], r[k]) <= 1e-3, (k, rl(o[k], r[k]))
    d = tmp_path / "edge_full"
    d.mkdir()
    images = {}
    for i, p in enumerate(points):
        Image.fromarray(p["frame"]).save(d / f"img_{i:03d}.png")
        images[f"img_{i:03d}"] = p["frame"]
    path = fc.cache_reachable(str(d), str(tmp_path / "out"))
    saved = torch.load(path)
    ref_r = reachable_features(images, rn50_visual, resnet_trunk=trunk)
    assert path.endswith("reachable_image_features.pt") and list(saved) == list(ref_r)
    for name in saved:
        assert list(saved[name]) == ["imagenet_avgpool", "clip_avgpool", "clip_attnpool"]
        for k in saved[name]:
            assert rl(saved[name][k], ref_r[name][k]) <= 1e-3, (name, k)
