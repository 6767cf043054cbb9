"""ORACLE (test infrastructure, not product code): the fp32 ModifiedResNet restatement of
oracle/clip_model.py re-evaluated with the kernel path's *rounding points* -- BatchNorm folded in fp32,
weights rounded to fp16 once, every fused op's output rounded to fp16, all sums in fp32.  This is also the
numeric regime of the reference's own in-tree call (clip.load on CUDA converts the model to fp16,
primitive_probing/generate_data/thor_image_features.py:57, with ``.float()`` only on the results :111-113),
except that the reference rounds after every conv / BN / ReLU / add separately while this path rounds once
per fused op.

Used by tests to separate two questions: (1) do the CUDA kernels compute exactly what the design says
(compare against this module, tight tolerance), and (2) how far is the design from the fp32 reference
(compare against oracle/clip_model.py, the north-star 1e-3 bar).

Returns every intermediate activation NHWC under the names the library reports
(embclip_rn50_act_info), so parity can be checked layer by layer.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .clip_model import ModifiedResNet

BN_EPS = 1e-5


def _h(t: torch.Tensor) -> torch.Tensor:
    return t.half().float()


def _fold(conv, bn):
    s = bn.weight / torch.sqrt(bn.running_var + BN_EPS)
    return conv.weight * s[:, None, None, None], bn.bias - bn.running_mean * s


@torch.no_grad()
def rn50_fp16_path(m: ModifiedResNet, frames_nchw: torch.Tensor, quantize: bool = True,
                   feed: Dict[str, torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """frames_nchw fp32 [B,3,R,R] -> dict of NHWC activations (+ 'trunk_nchw', 'avgpool', 'attnpool').
    quantize=False evaluates the same fused graph in pure fp32 (checks the folding algebra).

    feed: NHWC activations produced by the implementation under test.  When given, every op is evaluated on
    the *fed* inputs instead of this function's own chain, so each returned tensor is the expected output of
    ONE op given the inputs the kernels actually saw -- per-op isolation.  (Chained end to end, two fp16
    pipelines that differ only in fp32 summation order still drift apart by ~1e-3: a different summation
    order flips a few percent of the fp16 roundings per layer and the flips cascade.)"""
    q = _h if quantize else (lambda t: t)
    acts: Dict[str, torch.Tensor] = {}

    def put(name, t_nchw):
        acts[name] = t_nchw.permute(0, 2, 3, 1).contiguous()
        if feed is not None and name in feed:
            return feed[name].float().reshape(acts[name].shape).permute(0, 3, 1, 2).contiguous()
        return t_nchw

    def cbr(x, conv, bn, relu=True, extra=None, round_out=True, round_w=True):
        w, b = _fold(conv, bn)
        y = F.conv2d(x, q(w) if round_w else w, None, conv.stride, conv.padding) + b[None, :, None, None]
        if extra is not None:
            y = y + extra
        if relu:
            y = F.relu(y)
        return q(y) if round_out else y

    # stem conv1: hi/lo-split fp16 operands on the tensor cores = fp32 inputs x fp32 weights to 2^-22; only its output is rounded
    x = put("stem.conv1", cbr(frames_nchw, m.conv1, m.bn1, round_w=False))
    x = put("stem.conv2", cbr(x, m.conv2, m.bn2))
    # AvgPool2d(2) is fused into conv3's epilogue: the full-resolution map is rounded to fp16, averaged in fp32
    # and rounded once more (same rounding points as the separate pool kernel it replaced)
    x = put("stem.conv3", q(F.avg_pool2d(cbr(x, m.conv3, m.bn3), 2)))
    stages = [m.layer1, m.layer2, m.layer3, m.layer4]
    for li, layer in enumerate(stages):
        for bi, blk in enumerate(layer):
            p = f"layer{li + 1}.{bi}"
            last = li == 3 and bi == len(layer) - 1
            a = put(p + ".conv1", cbr(x, blk.conv1, blk.bn1))
            b = cbr(a, blk.conv2, blk.bn2)
            xp = x
            if blk.stride > 1:                      # anti-aliased stride: avgpool fused into conv2's epilogue
                b = q(F.avg_pool2d(b, blk.stride))
                xp = put(p + ".xpool", q(F.avg_pool2d(x, blk.stride)))
            b = put(p + ".conv2", b)
            if blk.downsample is not None:
                wd, bd = _fold(blk.downsample[1], blk.downsample[2])
                idn = F.conv2d(xp, q(wd)) + bd[None, :, None, None]      # fused along K: never rounded on its own
            else:
                idn = x
            x = put(p + ".conv3", cbr(b, blk.conv3, blk.bn3, extra=idn, round_out=not last))
    acts["trunk_nchw"] = x                     # (fed value of the last conv3 when feed is given)
    acts["avgpool"] = x.mean(dim=(2, 3))

    # attention pool with the kernel path's algebra and rounding points
    ap = m.attnpool
    B, E, Hf, Wf = x.shape
    heads, hd = ap.num_heads, E // ap.num_heads
    t = x.reshape(B, E, Hf * Wf).permute(0, 2, 1)                         # [B, P, E]
    def put2(name, val, shape4):
        acts[name] = val.reshape(shape4)
        if feed is not None and name in feed:
            return feed[name].float().reshape(val.shape)
        return val

    L = Hf * Wf + 1
    tok = put2("attnpool.tokens", q(torch.cat([t.mean(dim=1, keepdim=True), t], dim=1) + ap.positional_embedding[None]), (B, 1, L, E))
    s = hd ** -0.5
    qv = put2("attnpool.q", q(tok[:, 0] @ q(ap.q_proj.weight * s).t() + ap.q_proj.bias * s), (B, 1, 1, E))
    wk = q(ap.k_proj.weight).view(heads, hd, E)
    qt = put2("attnpool.qk", q(torch.einsum("bhd,hdc->bhc", qv.view(B, heads, hd), wk)), (B, 1, heads, E))
    # tensor-core core kernel: unnormalised exp(s - max) is rounded to fp16 (operand of the P.T contraction); the row sum
    # that normalises the result is taken over the UNROUNDED values
    sc = torch.einsum("bhc,bjc->bhj", qt, tok)
    e = torch.exp(sc - sc.amax(dim=-1, keepdim=True))
    xbar = put2("attnpool.xbar", q(torch.einsum("bhj,bjc->bhc", q(e), tok) / e.sum(dim=-1, keepdim=True)), (B, 1, heads, E))
    wv = q(ap.v_proj.weight).view(heads, hd, E)
    o = put2("attnpool.v", q(torch.einsum("bhc,hdc->bhd", xbar, wv).reshape(B, E) + ap.v_proj.bias), (B, 1, 1, E))
    acts["attnpool"] = o @ q(ap.c_proj.weight).t() + ap.c_proj.bias
    return acts


@torch.no_grad()
def torchvision_fp16_path(trunk: torch.nn.Sequential, frames_nchw: torch.Tensor, quantize: bool = True) -> Dict[str, torch.Tensor]:
    """The same "ideal fp16 path" for the torchvision ResNet-50 trunk (oracle/imagenet_resnet.py): BatchNorm folded in fp32,
    weights and the im2col'd frames rounded to fp16 once, one rounding per fused op (7x7 stem conv, conv1, strided conv2,
    conv3 + K-concatenated strided downsample + add), max-pool exact, fp32 sums.  NCHW in, dict of NHWC activations out."""
    q = _h if quantize else (lambda t: t)
    acts: Dict[str, torch.Tensor] = {}

    def put(name, t):
        acts[name] = t.permute(0, 2, 3, 1).contiguous()
        return t

    def cbr(x, conv, bn, relu=True, extra=None, round_out=True):
        w, b = _fold(conv, bn)
        y = F.conv2d(x, q(w), None, conv.stride, conv.padding) + b[None, :, None, None]
        if extra is not None:
            y = y + extra
        if relu:
            y = F.relu(y)
        return q(y) if round_out else y

    x = put("stem.conv1", cbr(q(frames_nchw), trunk[0], trunk[1]))
    x = put("stem.maxpool", trunk[3](x))
    for li in range(4):
        layer = trunk[4 + li]
        for bi, blk in enumerate(layer):
            p = f"layer{li + 1}.{bi}"
            last = li == 3 and bi == len(layer) - 1
            a = put(p + ".conv1", cbr(x, blk.conv1, blk.bn1))
            b = put(p + ".conv2", cbr(a, blk.conv2, blk.bn2))
            if blk.downsample is not None:
                wd, bd = _fold(blk.downsample[0], blk.downsample[1])
                idn = F.conv2d(x, q(wd), None, blk.downsample[0].stride) + bd[None, :, None, None]
            else:
                idn = x
            x = put(p + ".conv3", cbr(b, blk.conv3, blk.bn3, extra=idn, round_out=not last))
    acts["trunk_nchw"] = x
    acts["avgpool"] = x.mean(dim=(2, 3))
    return acts
