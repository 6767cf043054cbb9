"""AllenAct plugin surface for the B200 encoder: drop-in mirrors of
``allenact_plugins/clip_plugin/clip_preprocessors.py`` (`ClipResNetPreprocessor`, `ClipResNetEmbedder`) on top
of ``allenact/base_abstractions/preprocessor.py`` (`Preprocessor`) -- allenai/allenact v0.5.0, the dependency
pinned at /root/reference/readme_files/baselines_robothor_objectnav.md:6 and instantiated by the experiment
config named at :51.  Same class names, constructor arguments, attributes
(``input_uuids``, ``uuid``, ``observation_space``), ``process`` / ``to`` semantics and error behaviour; the
arithmetic runs in libembclip_b200.so instead of ``clip.load(...).visual``.

When ``allenact`` is importable the classes subclass its real ``Preprocessor`` so ``SensorPreprocessorGraph``
accepts them unchanged; offline (this image has neither allenact nor gym) they subclass a local protocol stub
with the same contract.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Sequence

import torch

try:                                                    # pragma: no cover - not installable offline
    from allenact.base_abstractions.preprocessor import Preprocessor as _AllenActPreprocessor
except Exception:                                       # noqa: BLE001
    _AllenActPreprocessor = None

try:                                                    # pragma: no cover
    import gym as _gym
except Exception:                                       # noqa: BLE001
    _gym = None


class Box:
    """Stand-in for ``gym.spaces.Box`` when gym is absent (only low / high / shape / dtype are read)."""

    def __init__(self, low, high, shape, dtype="float32"):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


def _make_box(shape):
    if _gym is not None:                                # pragma: no cover
        import numpy as np
        return _gym.spaces.Box(low=-np.inf, high=np.inf, shape=shape)
    return Box(float("-inf"), float("inf"), shape)


class Preprocessor:
    """allenact/base_abstractions/preprocessor.py `Preprocessor`: not an nn.Module; identified by `uuid`,
    consumes `input_uuids`, advertises `observation_space`; `process(obs)` returns a tensor; `to(device)`."""

    input_uuids: List[str]
    uuid: str
    observation_space: Any

    def __init__(self, input_uuids: List[str], output_uuid: str, observation_space: Any, **kwargs: Any) -> None:
        self.uuid = output_uuid
        self.input_uuids = input_uuids
        self.observation_space = observation_space

    def process(self, obs: Dict[str, Any], *args: Any, **kwargs: Any) -> Any:
        raise NotImplementedError()

    def to(self, device: torch.device) -> "Preprocessor":
        raise NotImplementedError()


_Base = _AllenActPreprocessor if _AllenActPreprocessor is not None else Preprocessor


def load_clip_visual_state_dict(clip_model_type: str, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                                weights_path: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """Where the frozen weights come from, in order: an explicit state dict; a file (argument or
    $EMBCLIP_CLIP_WEIGHTS: a torch-saved state dict, or the official TorchScript archive RN50.pt);
    ``clip.load`` when the openai/CLIP package is installed; seeded synthetic weights ONLY when
    $EMBCLIP_SYNTHETIC_WEIGHTS=1 (benchmarks / tests -- there is no checkpoint offline)."""
    if state_dict is not None:
        return state_dict
    path = weights_path or os.environ.get("EMBCLIP_CLIP_WEIGHTS")
    if path:
        try:
            obj = torch.jit.load(path, map_location="cpu").state_dict()
        except RuntimeError:
            obj = torch.load(path, map_location="cpu")
            obj = obj.get("state_dict", obj) if isinstance(obj, dict) else obj.state_dict()
        return obj
    try:                                                # pragma: no cover - package absent offline
        import clip
        return clip.load(clip_model_type, device="cpu")[0].state_dict()
    except ImportError:
        pass
    if os.environ.get("EMBCLIP_SYNTHETIC_WEIGHTS") == "1":
        from .synthetic import synthetic_rn50_state_dict
        if clip_model_type == "RN50x16":
            return synthetic_rn50_state_dict(seed=1234, layers=(6, 8, 18, 8), width=96, output_dim=768, input_resolution=384)
        return synthetic_rn50_state_dict(seed=1234)
    raise RuntimeError(
        f"ClipResNetPreprocessor: no weights for '{clip_model_type}': pass clip_state_dict=..., set "
        "$EMBCLIP_CLIP_WEIGHTS to RN50.pt / a saved state dict, install openai/CLIP, or set "
        "$EMBCLIP_SYNTHETIC_WEIGHTS=1 for seeded synthetic weights.")


class ClipResNetEmbedder:
    """Mirror of clip_preprocessors.py `ClipResNetEmbedder(resnet, pool)`: callable on NCHW frames
    [B,3,224,224] -> [B,2048,7,7] (pool=False) or [B,2048] (pool=True, adaptive average pool).
    Always frozen / eval (BatchNorm statistics are folded into the conv weights at construction)."""

    def __init__(self, clip_visual_state_dict: Dict[str, torch.Tensor], pool: bool = True, device: Any = "cuda:0",
                 input_resolution: int = 224):
        from .encoder import ClipRN50Encoder
        self.pool = pool
        # AllenAct's RGB sensor renders 224 x 224 for every CLIP ResNet (RN50x16's native 384 only matters to its attention
        # pool, which this embedder never calls): the trunk plan is built for the frames it will actually see
        self.encoder = ClipRN50Encoder(clip_visual_state_dict, device, input_resolution=input_resolution)

    def eval(self) -> "ClipResNetEmbedder":
        return self

    def forward_nhwc(self, x_nhwc: torch.Tensor) -> torch.Tensor:
        head = "avgpool" if self.pool else "trunk"
        return self.encoder(x_nhwc, want=(head,))[head]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.forward_nhwc(x.permute(0, 2, 3, 1).contiguous().float())

    __call__ = forward


class ClipResNetPreprocessor(_Base):
    """Preprocess RGB (or 1-channel depth) images with the frozen CLIP ResNet trunk.

    process(obs): ``obs[rgb_input_uuid]`` float32 NHWC [B,224,224,3] already mean/std normalised by the sensor
    -> float32 [B,2048,7,7] (pool=False) or [B,2048] (pool=True)."""

    CLIP_RGB_MEANS = (0.48145466, 0.4578275, 0.40821073)
    CLIP_RGB_STDS = (0.26862954, 0.26130258, 0.27577711)
    SUPPORTED = {"RN50": (2048, 7, 7), "RN50x16": (3072, 7, 7)}     # output shapes AllenAct declares at 224 x 224 input

    def __init__(self, rgb_input_uuid: str, clip_model_type: str, pool: bool,
                 device: Optional[torch.device] = None, device_ids: Optional[Sequence[Any]] = None,
                 output_uuid: str = "rgb_clip_resnet", clip_state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 weights_path: Optional[str] = None, **kwargs: Any):
        if clip_model_type not in self.SUPPORTED:
            raise AssertionError(f"clip_model_type '{clip_model_type}' not built for B200 (available: {sorted(self.SUPPORTED)})")
        output_shape = self.SUPPORTED[clip_model_type]
        if pool:
            output_shape = output_shape[:1]
        self.clip_model_type = clip_model_type
        self.pool = pool
        self.device = torch.device("cpu") if device is None else torch.device(device)
        self.device_ids = list(device_ids) if device_ids is not None else list(range(torch.cuda.device_count()))
        self._clip_state_dict = clip_state_dict
        self._weights_path = weights_path
        self._resnet: Optional[ClipResNetEmbedder] = None
        super().__init__(input_uuids=[rgb_input_uuid], output_uuid=output_uuid,
                         observation_space=_make_box(output_shape), **kwargs)

    @property
    def resnet(self) -> ClipResNetEmbedder:
        """Built lazily on first use, as upstream (the trainer calls .to(device) first)."""
        if self._resnet is None:
            if self.device.type != "cuda":
                raise RuntimeError("ClipResNetPreprocessor (embclip_b200): no CPU path -- call .to(cuda device) first")
            sd = load_clip_visual_state_dict(self.clip_model_type, self._clip_state_dict, self._weights_path)
            self._resnet = ClipResNetEmbedder(sd, pool=self.pool, device=self.device)
        return self._resnet

    def to(self, device: torch.device) -> "ClipResNetPreprocessor":
        device = torch.device(device)
        if self._resnet is not None and device != self.device:
            self._resnet = None                          # re-pack on the new device at next use
        self.device = device
        return self

    def process(self, obs: Dict[str, Any], *args: Any, **kwargs: Any) -> Any:
        x = obs[self.input_uuids[0]].to(self.device)     # bhwc, kept channels-last: the kernels are NHWC
        if x.shape[-1] == 1:                             # depth: repeat across the 3 channels
            x = x.repeat(1, 1, 1, 3)
        if x.dtype == torch.uint8:                       # raw RGB from an un-normalised sensor: normalised in the stem kernel
            return self.resnet.forward_nhwc(x.contiguous())
        return self.resnet.forward_nhwc(x.float().contiguous())


class ClipViTEmbedder:
    """Mirror of clip_preprocessors.py `ClipViTEmbedder`: NCHW frames [B,3,224,224] -> CLIP ViT image features [B,512]."""

    def __init__(self, clip_state_dict: Dict[str, torch.Tensor], device: Any = "cuda:0"):
        from .vit import ClipViTEncoder
        self.encoder = ClipViTEncoder(clip_state_dict, device)

    def eval(self) -> "ClipViTEmbedder":
        return self

    def forward_nhwc(self, x_nhwc: torch.Tensor) -> torch.Tensor:
        return self.encoder(x_nhwc)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.forward_nhwc(x.permute(0, 2, 3, 1).contiguous().float())

    __call__ = forward


class ClipViTPreprocessor(_Base):
    """Mirror of allenact_plugins/clip_plugin `ClipViTPreprocessor` (the zero-shot ObjectNav configs,
    /root/reference/readme_files/zeroshot_objectnav.md:17,27): process(obs) -> float32 [B,512] image features."""

    CLIP_RGB_MEANS = ClipResNetPreprocessor.CLIP_RGB_MEANS
    CLIP_RGB_STDS = ClipResNetPreprocessor.CLIP_RGB_STDS
    SUPPORTED = {"ViT-B/32": (512,)}

    def __init__(self, rgb_input_uuid: str, clip_model_type: str = "ViT-B/32", device: Optional[torch.device] = None,
                 device_ids: Optional[Sequence[Any]] = None, output_uuid: str = "rgb_clip_vit",
                 clip_state_dict: Optional[Dict[str, torch.Tensor]] = None, weights_path: Optional[str] = None, **kwargs: Any):
        if clip_model_type not in self.SUPPORTED:
            raise AssertionError(f"clip_model_type '{clip_model_type}' not built for B200 (available: {sorted(self.SUPPORTED)})")
        self.clip_model_type = clip_model_type
        self.device = torch.device("cpu") if device is None else torch.device(device)
        self.device_ids = list(device_ids) if device_ids is not None else list(range(torch.cuda.device_count()))
        self._clip_state_dict, self._weights_path = clip_state_dict, weights_path
        self._vit: Optional[ClipViTEmbedder] = None
        super().__init__(input_uuids=[rgb_input_uuid], output_uuid=output_uuid,
                         observation_space=_make_box(self.SUPPORTED[clip_model_type]), **kwargs)

    @property
    def vit(self) -> ClipViTEmbedder:
        if self._vit is None:
            if self.device.type != "cuda":
                raise RuntimeError("ClipViTPreprocessor (embclip_b200): no CPU path -- call .to(cuda device) first")
            if self._clip_state_dict is None and not (self._weights_path or os.environ.get("EMBCLIP_CLIP_WEIGHTS")):
                try:                                    # pragma: no cover - package absent offline
                    import clip
                    self._clip_state_dict = clip.load(self.clip_model_type, device="cpu")[0].state_dict()
                except ImportError:
                    raise RuntimeError("ClipViTPreprocessor: pass clip_state_dict=..., set $EMBCLIP_CLIP_WEIGHTS or install openai/CLIP")
            sd = load_clip_visual_state_dict(self.clip_model_type, self._clip_state_dict, self._weights_path)
            self._vit = ClipViTEmbedder(sd, device=self.device)
        return self._vit

    def to(self, device: torch.device) -> "ClipViTPreprocessor":
        device = torch.device(device)
        if self._vit is not None and device != self.device:
            self._vit = None
        self.device = device
        return self

    def process(self, obs: Dict[str, Any], *args: Any, **kwargs: Any) -> Any:
        x = obs[self.input_uuids[0]].to(self.device)
        if x.shape[-1] == 1:
            x = x.repeat(1, 1, 1, 3)
        return self.vit.forward_nhwc(x.float().contiguous())
