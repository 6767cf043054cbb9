"""Host-side weight preparation for the ModifiedResNet plan in libembclip_b200.so.

Takes a CLIP state dict with the official key names (openai/CLIP ``clip/model.py``:
``visual.conv1.weight``, ``visual.bn1.running_mean``, ``visual.layer1.0.downsample.0.weight``,
``visual.attnpool.q_proj.weight`` ...; the ``visual.`` prefix is optional) and produces the packed
blob the library asks for through ``embclip_rn50_param_info``:

* BatchNorm (eval mode, eps 1e-5 -- the state the reference freezes it in,
  primitive_probing/generate_data/thor_image_features.py:26-33) is folded into the preceding conv in
  fp32 *before* the cast to fp16:  w' = w * gamma / sqrt(var + eps),  b' = beta - mean * gamma / sqrt(var + eps).
* 3x3 weights go tap-major: [Cout, (kh, kw, Cin)] -- the K order the implicit-GEMM producer walks.
* A bottleneck's downsample conv is concatenated to its conv3 along K ([W3 | Wd], bias b3 + bd): one
  GEMM over [avgpool(relu(bn2(conv2))) ; avgpool(x)] replaces conv3 + downsample + add.
* AttentionPool2d: q is pre-scaled by head_dim**-0.5 (= 1/8, exact in fp16); Wk is stored transposed
  ([c, (head, d)]) for the folded-query contraction; bk is dropped (it shifts all keys of a head by the
  same amount, which softmax ignores).

Pure torch-on-CPU code: no GPU needed, covered by tests/test_packing.py.
"""
from __future__ import annotations

import re
from typing import Dict, Tuple

import torch

BN_EPS = 1e-5


def strip_visual_prefix(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    if any(k.startswith("visual.") for k in sd):
        return {k[len("visual."):]: v for k, v in sd.items() if k.startswith("visual.")}
    return dict(sd)


def infer_rn_cfg(sd: Dict[str, torch.Tensor]) -> dict:
    """Mirror of clip/model.py build_model's shape inference for the ModifiedResNet branch."""
    sd = strip_visual_prefix(sd)
    layers = []
    for b in (1, 2, 3, 4):
        idx = {int(m.group(1)) for k in sd for m in [re.match(rf"layer{b}\.(\d+)\.", k)] if m}
        layers.append(len(idx))
    width = sd["layer1.0.conv1.weight"].shape[0]
    grid = round((sd["attnpool.positional_embedding"].shape[0] - 1) ** 0.5)
    return dict(layers=tuple(layers), width=int(width), heads=int(width * 32 // 64),
                output_dim=int(sd["attnpool.c_proj.weight"].shape[0]), input_resolution=int(grid * 32))


def fold_bn(sd: Dict[str, torch.Tensor], conv: str, bn: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (weight [Cout,Cin,kh,kw] fp32 with BN scale folded in, bias [Cout] fp32)."""
    w = sd[conv + ".weight"].float()
    scale = sd[bn + ".weight"].float() / torch.sqrt(sd[bn + ".running_var"].float() + BN_EPS)
    bias = sd[bn + ".bias"].float() - sd[bn + ".running_mean"].float() * scale
    return w * scale[:, None, None, None], bias


def _kmajor(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, kh, kw] -> [Cout, kh*kw*Cin] (tap-major, channel fastest)."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def packed_tensors(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """name (as in embclip_param_info) -> CPU tensor in its final dtype / layout."""
    sd = strip_visual_prefix(sd)
    out: Dict[str, torch.Tensor] = {}

    # The library carries the stem's width/2 channels in a multiple of 32 (RN50x16: 48 -> 64).  The extra channels get zero
    # weights and zero bias, so they are exactly 0 after ReLU and contribute nothing downstream.
    c_half = sd["conv1.weight"].shape[0]
    c_pad = (c_half + 31) // 32 * 32
    pad_out = lambda w_, b_: (torch.nn.functional.pad(w_, (0, 0, 0, 0, 0, 0, 0, c_pad - w_.shape[0])), torch.nn.functional.pad(b_, (0, c_pad - b_.shape[0])))
    pad_in = lambda w_: torch.nn.functional.pad(w_, (0, 0, 0, 0, 0, c_pad - w_.shape[1]))
    w, b = fold_bn(sd, "conv1", "bn1")
    w, b = pad_out(w, b)
    out["stem.conv1.w"] = w.permute(2, 3, 1, 0).reshape(27, -1).contiguous()         # [(kh,kw,ci), co] fp32
    out["stem.conv1.b"] = b
    # tensor-core stem: weight rows [w_hi | w_hi | w_lo | 0] against im2col rows [v_hi | v_lo | v_hi | 0] (32-wide slots, 27 used)
    w27 = out["stem.conv1.w"].t().contiguous()                                        # [co, 27]
    w_hi = w27.half()
    w_lo = (w27 - w_hi.float()).half()
    wtc = torch.zeros(w27.shape[0], 128, dtype=torch.float16)
    wtc[:, 0:27], wtc[:, 32:59], wtc[:, 64:91] = w_hi, w_hi, w_lo
    out["stem.conv1.wtc"] = wtc
    for i in (2, 3):
        w, b = fold_bn(sd, f"conv{i}", f"bn{i}")
        w = pad_in(w)
        if i == 2:
            w, b = pad_out(w, b)
        out[f"stem.conv{i}.w"] = _kmajor(w).half()
        out[f"stem.conv{i}.b"] = b

    blocks = sorted({(int(m.group(1)), int(m.group(2))) for k in sd
                     for m in [re.match(r"layer(\d)\.(\d+)\.conv1\.weight", k)] if m})
    for (li, bi) in blocks:
        p = f"layer{li}.{bi}"
        w, b = fold_bn(sd, p + ".conv1", p + ".bn1")
        out[p + ".conv1.w"] = _kmajor(w).half()
        out[p + ".conv1.b"] = b
        w, b = fold_bn(sd, p + ".conv2", p + ".bn2")
        out[p + ".conv2.w"] = _kmajor(w).half()
        out[p + ".conv2.b"] = b
        w3, b3 = fold_bn(sd, p + ".conv3", p + ".bn3")
        w3 = _kmajor(w3)
        if p + ".downsample.0.weight" in sd:
            wd, bd = fold_bn(sd, p + ".downsample.0", p + ".downsample.1")
            w3 = torch.cat([w3, _kmajor(wd)], dim=1)
            b3 = b3 + bd
        out[p + ".conv3.w"] = w3.half()
        out[p + ".conv3.b"] = b3

    head_dim = 64
    s = head_dim ** -0.5
    out["attnpool.pos"] = sd["attnpool.positional_embedding"].float().contiguous()
    out["attnpool.q.w"] = (sd["attnpool.q_proj.weight"].float() * s).half()
    out["attnpool.q.b"] = sd["attnpool.q_proj.bias"].float() * s
    out["attnpool.kT.w"] = sd["attnpool.k_proj.weight"].float().t().contiguous().half()
    out["attnpool.v.w"] = sd["attnpool.v_proj.weight"].float().half()
    out["attnpool.v.b"] = sd["attnpool.v_proj.bias"].float()
    out["attnpool.c.w"] = sd["attnpool.c_proj.weight"].float().half()
    out["attnpool.c.b"] = sd["attnpool.c_proj.bias"].float()
    return out


# ---------------------------------------------------------------------------------------------------
# torchvision ResNet-50 (the reference's ImageNet baseline encoder, thor_image_features.py:46-49)
# ---------------------------------------------------------------------------------------------------
_TV_SEQ = {"0": "conv1", "1": "bn1", "4": "layer1", "5": "layer2", "6": "layer3", "7": "layer4"}
STEM7_K = 160          # 7 * 7 * 3 = 147 taps, zero-padded to a multiple of the 32-element k-block (csrc/tv_kernels.cuh)


def torchvision_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Accepts `models.resnet50().state_dict()` (conv1.weight, layer1.0.conv1.weight, ..., fc.* ignored) or the state dict of
    the reference's `nn.Sequential(*list(resnet.children())[:-2])` (0.weight, 1.running_mean, 4.0.conv1.weight, ...)."""
    if "conv1.weight" in sd:
        return {k: v for k, v in sd.items() if not k.startswith("fc.")}
    out = {}
    for k, v in sd.items():
        head, _, rest = k.partition(".")
        if head in _TV_SEQ:
            out[_TV_SEQ[head] + "." + rest] = v
    if "conv1.weight" not in out:
        raise KeyError("not a torchvision ResNet state dict (no conv1.weight / 0.weight)")
    return out


def infer_torchvision_cfg(sd: Dict[str, torch.Tensor]) -> dict:
    sd = torchvision_keys(sd)
    layers = []
    for b in (1, 2, 3, 4):
        idx = {int(m.group(1)) for k in sd for m in [re.match(rf"layer{b}\.(\d+)\.", k)] if m}
        layers.append(len(idx))
    width = int(sd["layer1.0.conv1.weight"].shape[0])
    if tuple(sd["conv1.weight"].shape) != (64, 3, 7, 7) or "layer1.0.conv3.weight" not in sd:
        raise ValueError("torchvision plan: expected a Bottleneck ResNet with a 7x7 stem of 64 channels")
    return dict(layers=tuple(layers), width=width, heads=0, output_dim=0, input_resolution=224, arch=1)


def packed_tensors_torchvision(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Same layouts as `packed_tensors` for the bottlenecks; the 7x7 stem becomes a [64, 160] fp16 GEMM weight whose K order
    (kh, kw, c) is the one im2col7x7s2_kernel writes."""
    sd = torchvision_keys(sd)
    out: Dict[str, torch.Tensor] = {}
    w, b = fold_bn(sd, "conv1", "bn1")
    w = _kmajor(w)                                                      # [64, 147]
    out["stem.conv1.w"] = torch.nn.functional.pad(w, (0, STEM7_K - w.shape[1])).half()
    out["stem.conv1.b"] = b
    blocks = sorted({(int(m.group(1)), int(m.group(2))) for k in sd
                     for m in [re.match(r"layer(\d)\.(\d+)\.conv1\.weight", k)] if m})
    for (li, bi) in blocks:
        p = f"layer{li}.{bi}"
        for i in (1, 2):
            w, b = fold_bn(sd, f"{p}.conv{i}", f"{p}.bn{i}")
            out[f"{p}.conv{i}.w"] = _kmajor(w).half()
            out[f"{p}.conv{i}.b"] = b
        w3, b3 = fold_bn(sd, p + ".conv3", p + ".bn3")
        w3 = _kmajor(w3)
        if p + ".downsample.0.weight" in sd:
            wd, bd = fold_bn(sd, p + ".downsample.0", p + ".downsample.1")
            w3 = torch.cat([w3, _kmajor(wd)], dim=1)
            b3 = b3 + bd
        out[p + ".conv3.w"] = w3.half()
        out[p + ".conv3.b"] = b3
    return out


def pack_blob(sd: Dict[str, torch.Tensor], param_infos, arch: int = 0) -> torch.Tensor:
    """param_infos: iterable of (name, dtype('f16'|'f32'), shape tuple, offset, nbytes) as reported by the
    library.  Returns a uint8 CPU tensor holding every packed tensor at its offset."""
    tensors = packed_tensors_torchvision(sd) if arch == 1 else packed_tensors(sd)
    infos = list(param_infos)
    total = max(off + ((nb + 255) // 256) * 256 for _, _, _, off, nb in infos)
    blob = torch.zeros(total, dtype=torch.uint8)
    for name, dtype, shape, off, nbytes in infos:
        if name not in tensors:
            raise KeyError(f"packing: library asks for '{name}', which the state dict does not provide")
        t = tensors[name]
        want = torch.float16 if dtype == "f16" else torch.float32
        if t.dtype != want or tuple(t.shape) != tuple(shape):
            raise ValueError(f"packing: '{name}' is {t.dtype}{tuple(t.shape)}, library expects {want}{tuple(shape)}")
        raw = t.contiguous().view(torch.uint8).reshape(-1)
        if raw.numel() != nbytes:
            raise ValueError(f"packing: '{name}' has {raw.numel()} bytes, library expects {nbytes}")
        blob[off:off + nbytes] = raw
    return blob
