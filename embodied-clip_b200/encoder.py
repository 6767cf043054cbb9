"""Host-side driver of the CLIP-RN50 encoder plan in libembclip_b200.so.

``ClipRN50Encoder`` is what the plugin surface (plugin.py) and bench.py call: it owns the library
handle, the packed weight blob and the workspace (torch CUDA allocations -- torch is plumbing for
device memory and streams only; every FLOP runs in the library's kernels).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Tuple

import torch

from . import _lib
from .packing import infer_rn_cfg, infer_torchvision_cfg, pack_blob


class ClipRN50Encoder:
    """Frozen CLIP ModifiedResNet forward: frames fp32 NHWC [B,R,R,3] (already mean/std normalised) or uint8 NHWC (raw
    RGB, normalised in the stem kernel) ->
    'trunk' fp32 [B,2048,7,7] | 'avgpool' fp32 [B,2048] | 'attnpool' fp32 [B,1024].

    Replaces ``clip_model.visual`` as used at
    primitive_probing/generate_data/thor_image_features.py:57-67,109-113."""

    HEADS = ("trunk", "avgpool", "attnpool")
    CLIP_RGB_MEANS = (0.48145466, 0.4578275, 0.40821073)
    CLIP_RGB_STDS = (0.26862954, 0.26130258, 0.27577711)

    ARCH = 0                     # embclip_rn50_cfg.arch

    def _infer_cfg(self, state_dict):
        return infer_rn_cfg(state_dict)

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: torch.device | str = "cuda:0",
                 input_resolution: Optional[int] = None):
        """``input_resolution``: run the trunk at another resolution than the checkpoint's attention pool was trained for
        (AllenAct feeds RN50x16 -- native 384 x 384 -- with 224 x 224 frames and uses the trunk / avg-pool outputs only).
        The 'attnpool' head exists only when the positional embedding matches the resolution and has at most 64 tokens."""
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("embclip_b200 has no CPU path: ClipRN50Encoder needs a CUDA (sm_100a) device")
        cfg = self._infer_cfg(state_dict)
        native = cfg["input_resolution"]
        if input_resolution is not None:
            cfg["input_resolution"] = int(input_resolution)
        tokens = (cfg["input_resolution"] // 32) ** 2 + 1
        self.has_attnpool = cfg["input_resolution"] == native and tokens <= 64 and cfg["output_dim"] > 0
        if not self.has_attnpool:
            cfg["input_resolution_native"] = native
        self.cfg = cfg
        c = _lib.RN50Cfg()
        c.layers[:] = cfg["layers"]
        c.width, c.heads, c.input_resolution = cfg["width"], cfg["heads"], cfg["input_resolution"]
        c.output_dim = cfg["output_dim"] if self.has_attnpool else 0          # 0: plan without the attention-pool head
        c.arch = self.ARCH
        self._h = C.c_void_p()
        _lib.check(self.lib.embclip_rn50_create(C.byref(c), C.byref(self._h)))
        self.param_infos = self._param_infos()
        blob = pack_blob(state_dict, self.param_infos, arch=self.ARCH)
        with torch.cuda.device(self.device):
            self._blob = blob.to(self.device)
            _lib.check(self.lib.embclip_rn50_bind_weights(self._h, self._blob.data_ptr(), self._blob.numel()))
        self._ws: Optional[torch.Tensor] = None
        self.embed = cfg["width"] * 32
        self.fres = cfg["input_resolution"] // 32

    # ------------------------------------------------------------------ introspection
    def _param_infos(self) -> List[Tuple[str, str, tuple, int, int]]:
        n = _lib.check(self.lib.embclip_rn50_num_params(self._h))
        out = []
        for i in range(n):
            pi = _lib.ParamInfo()
            _lib.check(self.lib.embclip_rn50_param_info(self._h, i, C.byref(pi)))
            out.append((pi.name.decode(), "f16" if pi.dtype == _lib.DTYPE_F16 else "f32",
                        tuple(pi.shape[:pi.ndim]), int(pi.offset), int(pi.nbytes)))
        return out

    def workspace_bytes(self, batch: int) -> int:
        return int(self.lib.embclip_rn50_workspace_bytes(self._h, batch))

    def _workspace(self, batch: int) -> torch.Tensor:
        need = self.workspace_bytes(batch)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            raw = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
            off = (-raw.data_ptr()) % 1024
            self._ws = raw[off:off + need]
        return self._ws

    def launches_per_forward(self, want: Iterable[str]) -> int:
        w = set(want)
        return _lib.check(self.lib.embclip_rn50_launches_per_forward(self._h, "trunk" in w, "avgpool" in w, "attnpool" in w))

    def activations(self, batch: int) -> Dict[str, torch.Tensor]:
        """Views of the intermediate NHWC activations of the LAST forward at this batch size (test hook)."""
        ws = self._workspace(batch)
        n = _lib.check(self.lib.embclip_rn50_num_acts(self._h))
        out = {}
        for i in range(n):
            ai = _lib.ActInfo()
            _lib.check(self.lib.embclip_rn50_act_info(self._h, batch, i, C.byref(ai)))
            if ai.n == 0:       # not materialised: handed to its fused consumers on chip (bneck_tail's pooled-output variant)
                continue
            dt = torch.float16 if ai.dtype == _lib.DTYPE_F16 else torch.float32
            numel = ai.n * ai.h * ai.w * ai.c
            nbytes = numel * (2 if dt == torch.float16 else 4)
            out[ai.name.decode()] = ws[ai.offset:ai.offset + nbytes].view(dt).view(ai.n, ai.h, ai.w, ai.c)
        return out

    # ------------------------------------------------------------------ compute
    def _outputs(self, batch: int, want: Iterable[str]) -> Dict[str, torch.Tensor]:
        outs = {}
        for k in want:
            if k == "attnpool" and not self.has_attnpool:
                raise ValueError("the 'attnpool' head is not available: the positional embedding is for "
                                 f"{self.cfg.get('input_resolution_native', self.cfg['input_resolution'])} x "
                                 f"{self.cfg.get('input_resolution_native', self.cfg['input_resolution'])} frames and the library builds it for at most 64 tokens")
            if k == "trunk":
                outs[k] = torch.empty(batch, self.embed, self.fres, self.fres, dtype=torch.float32, device=self.device)
            elif k == "avgpool":
                outs[k] = torch.empty(batch, self.embed, dtype=torch.float32, device=self.device)
            elif k == "attnpool":
                outs[k] = torch.empty(batch, self.cfg["output_dim"], dtype=torch.float32, device=self.device)
            else:
                raise ValueError(f"unknown head '{k}' (choose from {self.HEADS})")
        return outs

    def _check_frames(self, frames: torch.Tensor) -> torch.Tensor:
        R = self.cfg["input_resolution"]
        if frames.device != self.device:
            raise ValueError(f"frames on {frames.device}, encoder on {self.device}")
        if frames.dtype not in (torch.float32, torch.uint8) or frames.dim() != 4 or tuple(frames.shape[1:]) != (R, R, 3):
            raise ValueError(f"frames must be float32 (normalised) or uint8 (raw) NHWC [B,{R},{R},3], got {frames.dtype} {tuple(frames.shape)}")
        return frames.contiguous()

    def forward(self, frames: torch.Tensor, want: Iterable[str] = ("trunk",),
                out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        want = tuple(want)
        frames = self._check_frames(frames)
        B = frames.shape[0]
        if B == 0:
            return self._outputs(0, want)
        outs = out if out is not None else self._outputs(B, want)
        ws = self._workspace(B)
        ptr = lambda k: outs[k].data_ptr() if k in outs else None
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            if frames.dtype == torch.uint8:       # raw RGB: mean/std normalisation happens inside the stem kernel
                mean, std = (C.c_float * 3)(*self.CLIP_RGB_MEANS), (C.c_float * 3)(*self.CLIP_RGB_STDS)
                _lib.check(self.lib.embclip_rn50_forward_u8(self._h, frames.data_ptr(), mean, std, B, ptr("trunk"), ptr("avgpool"),
                                                            ptr("attnpool"), ws.data_ptr(), ws.numel(), stream))
            else:
                _lib.check(self.lib.embclip_rn50_forward(self._h, frames.data_ptr(), B, ptr("trunk"), ptr("avgpool"),
                                                         ptr("attnpool"), ws.data_ptr(), ws.numel(), stream))
        return outs

    __call__ = forward

    def encode_rows(self, frames: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Trunk forward, returned as fp16 NHWC pixel rows [B * fres * fres, embed]: the layout (and rounding) of
        ``ResnetTensorNavActorCritic.pack_features(trunk)``, without the fp32 NCHW round trip.  For rollout loops that keep
        their feature storage on the device."""
        frames = self._check_frames(frames)
        B = frames.shape[0]
        rows = B * self.fres * self.fres
        if out is None:
            out = torch.empty(rows, self.embed, dtype=torch.float16, device=self.device)
        elif out.dtype != torch.float16 or out.numel() != rows * self.embed or not out.is_contiguous() or out.device != self.device:
            raise ValueError(f"out must be a contiguous fp16 tensor of {rows} x {self.embed} elements on {self.device}")
        if B == 0:
            return out
        ws = self._workspace(B)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            u8 = frames.dtype == torch.uint8
            mean = (C.c_float * 3)(*self.CLIP_RGB_MEANS) if u8 else None
            std = (C.c_float * 3)(*self.CLIP_RGB_STDS) if u8 else None
            _lib.check(self.lib.embclip_rn50_encode_rows_f16(self._h, frames.data_ptr(), int(u8), mean, std, B, out.data_ptr(),
                                                             ws.data_ptr(), ws.numel(), stream))
        return out

    def profile(self, frames: torch.Tensor, want: Iterable[str] = ("trunk",)) -> List[Tuple[str, float]]:
        """Per-op device time (ms) of one forward, CUDA events on the current stream."""
        want = tuple(want)
        frames = self._check_frames(frames)
        B = frames.shape[0]
        outs = self._outputs(B, want)
        ws = self._workspace(B)
        max_ops = 256
        ms = (C.c_float * max_ops)()
        names = C.create_string_buffer(64 * max_ops)
        ptr = lambda k: outs[k].data_ptr() if k in outs else None
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            tail = (ptr("trunk"), ptr("avgpool"), ptr("attnpool"), ws.data_ptr(), ws.numel(), stream,
                    C.cast(ms, C.c_void_p), C.cast(names, C.c_void_p), max_ops)
            if frames.dtype == torch.uint8:
                mean, std = (C.c_float * 3)(*self.CLIP_RGB_MEANS), (C.c_float * 3)(*self.CLIP_RGB_STDS)
                n = _lib.check(self.lib.embclip_rn50_profile_u8(self._h, frames.data_ptr(), mean, std, B, *tail))
            else:
                n = _lib.check(self.lib.embclip_rn50_profile(self._h, frames.data_ptr(), B, *tail))
        return [(names.raw[i * 64:(i + 1) * 64].split(b"\0")[0].decode(), float(ms[i])) for i in range(n)]

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.embclip_rn50_destroy(self._h)
                self._h = None
        except Exception:
            pass


class TorchvisionResNet50Encoder(ClipRN50Encoder):
    """The reference's ImageNet baseline encoder: ``models.resnet50(pretrained=True)`` cut after layer4
    (``nn.Sequential(*list(resnet.children())[:-2])``, primitive_probing/generate_data/thor_image_features.py:46-49; forward at
    :101-105, reachable_image_features.py:49-52,81-85), frozen with ``freeze_model`` (:26-33).

    frames fp32 NHWC [B,224,224,3] (normalised with the ImageNet mean / std of ``resnet_preprocess``, :36-44) or uint8 NHWC
    (raw RGB, normalised in the im2col kernel) -> 'trunk' fp32 [B,2048,7,7] ('imagenet_conv') | 'avgpool' fp32 [B,2048]
    ('imagenet_avgpool').  Same kernels as the CLIP plan: the 7x7/2 stem is an im2col + GEMM, the stride-2 3x3 conv is the
    halo kernel keeping every second position, the stride-2 1x1 downsample is K-concatenated to conv3 after a ::2 subsample.
    The state dict is torchvision's (``resnet50().state_dict()``) or the cut ``nn.Sequential``'s."""

    ARCH = 1
    HEADS = ("trunk", "avgpool")
    CLIP_RGB_MEANS = (0.485, 0.456, 0.406)        # name kept from the base class: the mean / std applied to uint8 frames
    CLIP_RGB_STDS = (0.229, 0.224, 0.225)

    def _infer_cfg(self, state_dict):
        return infer_torchvision_cfg(state_dict)
