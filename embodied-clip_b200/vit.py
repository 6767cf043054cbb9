"""Host-side drivers of the CLIP transformer towers in libembclip_b200.so (zero-shot ObjectNav path, BASELINE.json
config 5): ``ClipViTEncoder`` = ``CLIP.encode_image`` for ViT-B/32,
``ClipTextEncoder`` = ``CLIP.encode_text``, ``ClipZeroShot`` = ``CLIP.forward``'s cosine-similarity logits with the
text side computed once per prompt set and cached.

Weights come from a CLIP state dict with the official key names (openai/CLIP clip/model.py:
``visual.conv1.weight``, ``visual.transformer.resblocks.0.attn.in_proj_weight``, ``token_embedding.weight``,
``text_projection``, ``logit_scale`` ...).  Packing (CPU, no GPU needed; tests/test_packing.py):
  * patch conv [768,3,32,32] -> [768, (kh,kw,c)] fp16, the K order of the patch rows;
  * in_proj weight/bias: the q third is pre-scaled by head_dim**-0.5 = 1/8 (exact in fp16);
  * proj / text_projection are stored transposed ([out, width]) as GEMM weights;
  * LayerNorm parameters, biases, positional / class / token embeddings stay fp32.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

HEAD_DIM = 64


def _blocks(sd: Dict[str, torch.Tensor], prefix: str) -> int:
    import re
    return len({int(m.group(1)) for k in sd for m in [re.match(re.escape(prefix) + r"transformer\.resblocks\.(\d+)\.", k)] if m})


def infer_vit_cfg(sd: Dict[str, torch.Tensor]) -> dict:
    """Mirror of clip/model.py build_model's shape inference for the ViT branch."""
    w = sd["visual.conv1.weight"]
    width, patch = int(w.shape[0]), int(w.shape[-1])
    grid = round((sd["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    return dict(kind=_lib.TF_VISION, width=width, layers=_blocks(sd, "visual."), heads=width // HEAD_DIM,
                output_dim=int(sd["visual.proj"].shape[1]), patch_size=patch, input_resolution=patch * grid,
                context_length=0, vocab_size=0)


def infer_text_cfg(sd: Dict[str, torch.Tensor]) -> dict:
    width = int(sd["ln_final.weight"].shape[0])
    return dict(kind=_lib.TF_TEXT, width=width, layers=_blocks(sd, ""), heads=width // HEAD_DIM,
                output_dim=int(sd["text_projection"].shape[1]), patch_size=0, input_resolution=0,
                context_length=int(sd["positional_embedding"].shape[0]), vocab_size=int(sd["token_embedding.weight"].shape[0]))


def packed_tower_tensors(sd: Dict[str, torch.Tensor], cfg: dict) -> Dict[str, torch.Tensor]:
    """name (as in embclip_tf_param_info) -> CPU tensor in its final dtype / layout."""
    vision = cfg["kind"] == _lib.TF_VISION
    pre = "visual." if vision else ""
    D = cfg["width"]
    out: Dict[str, torch.Tensor] = {}
    f32 = lambda k: sd[k].float().contiguous()
    if vision:
        out["patch.w"] = sd["visual.conv1.weight"].float().permute(0, 2, 3, 1).reshape(D, -1).contiguous().half()
        out["cls"] = f32("visual.class_embedding")
        out["pos"] = f32("visual.positional_embedding")
        out["ln_pre.w"], out["ln_pre.b"] = f32("visual.ln_pre.weight"), f32("visual.ln_pre.bias")
        out["ln_post.w"], out["ln_post.b"] = f32("visual.ln_post.weight"), f32("visual.ln_post.bias")
        out["head.w"] = sd["visual.proj"].float().t().contiguous().half()
    else:
        out["tok_emb"] = f32("token_embedding.weight")
        out["pos"] = f32("positional_embedding")
        out["ln_post.w"], out["ln_post.b"] = f32("ln_final.weight"), f32("ln_final.bias")
        out["head.w"] = sd["text_projection"].float().t().contiguous().half()
    s = HEAD_DIM ** -0.5
    for i in range(cfg["layers"]):
        p, q = f"{pre}transformer.resblocks.{i}.", f"blk{i}."
        w, b = sd[p + "attn.in_proj_weight"].float().clone(), sd[p + "attn.in_proj_bias"].float().clone()
        w[:D] *= s
        b[:D] *= s
        out[q + "qkv.w"], out[q + "qkv.b"] = w.half(), b
        out[q + "out.w"], out[q + "out.b"] = sd[p + "attn.out_proj.weight"].float().half(), f32(p + "attn.out_proj.bias")
        out[q + "ln1.w"], out[q + "ln1.b"] = f32(p + "ln_1.weight"), f32(p + "ln_1.bias")
        out[q + "ln2.w"], out[q + "ln2.b"] = f32(p + "ln_2.weight"), f32(p + "ln_2.bias")
        out[q + "fc.w"], out[q + "fc.b"] = sd[p + "mlp.c_fc.weight"].float().half(), f32(p + "mlp.c_fc.bias")
        out[q + "proj.w"], out[q + "proj.b"] = sd[p + "mlp.c_proj.weight"].float().half(), f32(p + "mlp.c_proj.bias")
    return out


class _Tower:
    def __init__(self, sd: Dict[str, torch.Tensor], cfg: dict, device):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("embclip_b200 has no CPU path: the CLIP transformer towers need a CUDA (sm_100a) device")
        self.cfg = cfg
        c = _lib.TFCfg(**cfg)
        self._h = C.c_void_p()
        _lib.check(self.lib.embclip_tf_create(C.byref(c), C.byref(self._h)))
        tensors = packed_tower_tensors(sd, cfg)
        infos: List[Tuple[str, str, tuple, int, int]] = []
        for i in range(_lib.check(self.lib.embclip_tf_num_params(self._h))):
            pi = _lib.ParamInfo()
            _lib.check(self.lib.embclip_tf_param_info(self._h, i, C.byref(pi)))
            infos.append((pi.name.decode(), "f16" if pi.dtype == _lib.DTYPE_F16 else "f32", tuple(pi.shape[:pi.ndim]), int(pi.offset), int(pi.nbytes)))
        blob = torch.zeros(int(self.lib.embclip_tf_blob_bytes(self._h)), dtype=torch.uint8)
        for name, dtype, shape, off, nbytes in infos:
            if name not in tensors:
                raise KeyError(f"packing: library asks for '{name}', which the state dict does not provide")
            t = tensors[name]
            want = torch.float16 if dtype == "f16" else torch.float32
            if t.dtype != want or tuple(t.shape) != tuple(shape):
                raise ValueError(f"packing: '{name}' is {t.dtype}{tuple(t.shape)}, library expects {want}{tuple(shape)}")
            blob[off:off + nbytes] = t.contiguous().view(torch.uint8).reshape(-1)
        self.param_infos = infos
        with torch.cuda.device(self.device):
            self._blob = blob.to(self.device)
            _lib.check(self.lib.embclip_tf_bind_weights(self._h, self._blob.data_ptr(), self._blob.numel()))
        self._ws: Optional[torch.Tensor] = None

    def _workspace(self, batch: int) -> torch.Tensor:
        need = int(self.lib.embclip_tf_workspace_bytes(self._h, batch))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            raw = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
            off = (-raw.data_ptr()) % 1024
            self._ws = raw[off:off + need]
        return self._ws

    def launches_per_forward(self) -> int:
        return _lib.check(self.lib.embclip_tf_launches_per_forward(self._h))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.embclip_tf_destroy(self._h)
                self._h = None
        except Exception:
            pass


class ClipViTEncoder(_Tower):
    """frames fp32 NHWC [B,224,224,3] (mean/std normalised) -> image features fp32 [B,512] (``CLIP.encode_image``)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda:0"):
        super().__init__(state_dict, infer_vit_cfg(state_dict), device)

    def forward(self, frames: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        R = self.cfg["input_resolution"]
        if frames.device != self.device:
            raise ValueError(f"frames on {frames.device}, encoder on {self.device}")
        if frames.dtype != torch.float32 or frames.dim() != 4 or tuple(frames.shape[1:]) != (R, R, 3):
            raise ValueError(f"frames must be float32 NHWC [B,{R},{R},3], got {frames.dtype} {tuple(frames.shape)}")
        frames = frames.contiguous()
        B = frames.shape[0]
        if out is None:
            out = torch.empty(B, self.cfg["output_dim"], dtype=torch.float32, device=self.device)
        if B == 0:
            return out
        ws = self._workspace(B)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.embclip_vit_forward(self._h, frames.data_ptr(), B, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                                    torch.cuda.current_stream(self.device).cuda_stream))
        return out

    __call__ = forward


class ClipTextEncoder(_Tower):
    """token ids int64 [K,77] -> text features fp32 [K,512] (``CLIP.encode_text``; EOT = the highest token id)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda:0"):
        super().__init__(state_dict, infer_text_cfg(state_dict), device)

    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        L = self.cfg["context_length"]
        if tokens.dim() != 2 or tokens.shape[1] != L:
            raise ValueError(f"tokens must be [prompts, {L}], got {tuple(tokens.shape)}")
        t = tokens.to(self.device, torch.int64).contiguous()
        K = t.shape[0]
        out = torch.empty(K, self.cfg["output_dim"], dtype=torch.float32, device=self.device)
        if K == 0:
            return out
        ws = self._workspace(K)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.embclip_text_forward(self._h, t.data_ptr(), K, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                                     torch.cuda.current_stream(self.device).cuda_stream))
        return out

    __call__ = forward


class ClipZeroShot:
    """``CLIP.forward``: logits_per_image = exp(logit_scale) * normalize(encode_image(x)) @ normalize(encode_text(t))^T.
    The prompt set is fixed per task (12 RoboTHOR object classes), so the text features are computed once and cached."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda:0"):
        self.image = ClipViTEncoder(state_dict, device)
        self.text = ClipTextEncoder(state_dict, device)
        self.logit_scale = float(state_dict["logit_scale"])
        self._text_features: Optional[torch.Tensor] = None

    def set_prompts(self, tokens: torch.Tensor) -> torch.Tensor:
        self._text_features = self.text(tokens)
        return self._text_features

    def logits(self, image_features: torch.Tensor, text_features: Optional[torch.Tensor] = None) -> torch.Tensor:
        tf = text_features if text_features is not None else self._text_features
        if tf is None:
            raise RuntimeError("ClipZeroShot: call set_prompts(tokens) first")
        B, E = image_features.shape
        K = tf.shape[0]
        out = torch.empty(B, K, dtype=torch.float32, device=image_features.device)
        with torch.cuda.device(image_features.device):
            _lib.check(self.image.lib.embclip_clip_logits(image_features.data_ptr(), tf.data_ptr(), B, K, E, self.logit_scale, out.data_ptr(),
                                                          torch.cuda.current_stream(image_features.device).cuda_stream))
        return out

    def forward(self, frames: torch.Tensor, tokens: Optional[torch.Tensor] = None) -> torch.Tensor:
        if tokens is not None:
            self.set_prompts(tokens)
        return self.logits(self.image(frames))

    __call__ = forward
