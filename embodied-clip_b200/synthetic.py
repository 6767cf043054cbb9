"""Seeded synthetic CLIP-RN50 weights with the official state-dict key names (no checkpoint exists offline;
SURVEY.md section 8d config 2).  Well conditioned: activations stay O(1) through all 16 residual blocks in
fp16, and the attention-pool logits are O(1) like a trained model's.

conv ~ N(0, 2/fan_in); BN gamma in U(.5,1.5) (x0.5 on every block's last BN and downsample BN), beta ~ N(0,.1),
running_mean ~ N(0,.1), running_var in U(.5,1.5); linear ~ N(0, 1/in) (x0.25 on q/k projections), bias ~ N(0,.1).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import torch


def synthetic_rn50_state_dict(seed: int = 1234, layers=(3, 4, 6, 3), width: int = 64, output_dim: int = 1024,
                              input_resolution: int = 224, prefix: str = "") -> "OrderedDict[str, torch.Tensor]":
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    rn = lambda *s: torch.randn(*s, generator=g)
    ru = lambda *s: torch.rand(*s, generator=g)

    def conv(name, cout, cin, k):
        sd[prefix + name + ".weight"] = rn(cout, cin, k, k) * (2.0 / (cin * k * k)) ** 0.5

    def bn(name, c, gain=1.0):
        sd[prefix + name + ".weight"] = (0.5 + ru(c)) * gain
        sd[prefix + name + ".bias"] = 0.1 * rn(c)
        sd[prefix + name + ".running_mean"] = 0.1 * rn(c)
        sd[prefix + name + ".running_var"] = 0.5 + ru(c)
        sd[prefix + name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    def linear(name, cout, cin, gain=1.0):
        sd[prefix + name + ".weight"] = rn(cout, cin) * cin ** -0.5 * gain
        sd[prefix + name + ".bias"] = 0.1 * rn(cout)

    conv("conv1", width // 2, 3, 3); bn("bn1", width // 2)
    conv("conv2", width // 2, width // 2, 3); bn("bn2", width // 2)
    conv("conv3", width, width // 2, 3); bn("bn3", width)
    inplanes = width
    for li, nblocks in enumerate(layers):
        planes = width << li
        for bi in range(nblocks):
            p = f"layer{li + 1}.{bi}"
            stride = 2 if (bi == 0 and li > 0) else 1
            conv(p + ".conv1", planes, inplanes, 1); bn(p + ".bn1", planes)
            conv(p + ".conv2", planes, planes, 3); bn(p + ".bn2", planes)
            conv(p + ".conv3", planes * 4, planes, 1); bn(p + ".bn3", planes * 4, gain=0.5)
            if stride > 1 or inplanes != planes * 4:
                conv(p + ".downsample.0", planes * 4, inplanes, 1); bn(p + ".downsample.1", planes * 4, gain=0.5)
            inplanes = planes * 4
    embed = width * 32
    tokens = (input_resolution // 32) ** 2 + 1
    sd[prefix + "attnpool.positional_embedding"] = rn(tokens, embed) * embed ** -0.5
    linear("attnpool.k_proj", embed, embed, gain=0.25)
    linear("attnpool.q_proj", embed, embed, gain=0.25)
    linear("attnpool.v_proj", embed, embed)
    linear("attnpool.c_proj", output_dim, embed)
    return sd


def synthetic_clip_vit_b32_state_dict(seed: int = 1234, width: int = 768, layers: int = 12, patch: int = 32, resolution: int = 224,
                                      embed_dim: int = 512, text_width: int = 512, text_layers: int = 12, context: int = 77,
                                      vocab: int = 49408) -> "OrderedDict[str, torch.Tensor]":
    """Seeded synthetic CLIP ViT-B/32 weights (image tower + text tower + logit_scale) with the official state-dict key
    names.  Matrices ~ N(0, 1/fan_in) (embeddings N(0, 0.02^2)), LayerNorm gains
    1 + 0.1 N(0,1), biases 0.02 N(0,1), logit_scale = ln(100)."""
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    rn = lambda *s: torch.randn(*s, generator=g)

    def mat(name, *shape, emb=False):
        sd[name] = rn(*shape) * (0.02 if emb else shape[-1] ** -0.5)

    def ln(name, d):
        sd[name + ".weight"] = 1.0 + 0.1 * rn(d)
        sd[name + ".bias"] = 0.02 * rn(d)

    def blocks(prefix, d, n):
        for i in range(n):
            p = f"{prefix}transformer.resblocks.{i}."
            ln(p + "ln_1", d)
            mat(p + "attn.in_proj_weight", 3 * d, d); sd[p + "attn.in_proj_bias"] = 0.02 * rn(3 * d)
            mat(p + "attn.out_proj.weight", d, d); sd[p + "attn.out_proj.bias"] = 0.02 * rn(d)
            ln(p + "ln_2", d)
            mat(p + "mlp.c_fc.weight", 4 * d, d); sd[p + "mlp.c_fc.bias"] = 0.02 * rn(4 * d)
            mat(p + "mlp.c_proj.weight", d, 4 * d); sd[p + "mlp.c_proj.bias"] = 0.02 * rn(d)

    sd["visual.conv1.weight"] = rn(width, 3, patch, patch) * (3 * patch * patch) ** -0.5
    sd["visual.class_embedding"] = 0.02 * rn(width)
    mat("visual.positional_embedding", (resolution // patch) ** 2 + 1, width, emb=True)
    ln("visual.ln_pre", width)
    blocks("visual.", width, layers)
    ln("visual.ln_post", width)
    mat("visual.proj", width, embed_dim)
    sd["visual.proj"] = rn(width, embed_dim) * width ** -0.5
    mat("token_embedding.weight", vocab, text_width, emb=True)
    mat("positional_embedding", context, text_width, emb=True)
    blocks("", text_width, text_layers)
    ln("ln_final", text_width)
    sd["text_projection"] = rn(text_width, embed_dim) * text_width ** -0.5
    sd["logit_scale"] = torch.tensor(4.605170185988092)
    return sd


def synthetic_torchvision_rn50_state_dict(seed: int = 4321, layers=(3, 4, 6, 3)) -> "OrderedDict[str, torch.Tensor]":
    """Seeded synthetic weights with torchvision's ``resnet50().state_dict()`` key names (conv1 / bn1 / layer{1-4}.{i}.{conv,bn}{1-3} /
    downsample.{0,1}; no fc) for the reference's ImageNet baseline encoder (thor_image_features.py:46-49).  Same
    conditioning as the CLIP generator above: conv ~ N(0, 2/fan_in), BN gamma in U(.5,1.5) (x0.25 on bn3 / downsample BN: this
    architecture adds the un-pooled identity, so the same residual gain as the CLIP set would grow the stream 12x over the 16
    blocks), beta, running_mean ~ N(0,.1), running_var in U(.5,1.5)."""
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    rn = lambda *s: torch.randn(*s, generator=g)
    ru = lambda *s: torch.rand(*s, generator=g)

    def conv(name, cout, cin, k):
        sd[name + ".weight"] = rn(cout, cin, k, k) * (2.0 / (cin * k * k)) ** 0.5

    def bn(name, c, gain=1.0):
        sd[name + ".weight"] = (0.5 + ru(c)) * gain
        sd[name + ".bias"] = 0.1 * rn(c)
        sd[name + ".running_mean"] = 0.1 * rn(c)
        sd[name + ".running_var"] = 0.5 + ru(c)
        sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    conv("conv1", 64, 3, 7); bn("bn1", 64)
    inplanes = 64
    for li, nblocks in enumerate(layers):
        planes = 64 << li
        for bi in range(nblocks):
            p = f"layer{li + 1}.{bi}"
            stride = 2 if (bi == 0 and li > 0) else 1
            conv(p + ".conv1", planes, inplanes, 1); bn(p + ".bn1", planes)
            conv(p + ".conv2", planes, planes, 3); bn(p + ".bn2", planes)
            conv(p + ".conv3", planes * 4, planes, 1); bn(p + ".bn3", planes * 4, gain=0.25)
            if stride > 1 or inplanes != planes * 4:
                conv(p + ".downsample.0", planes * 4, inplanes, 1); bn(p + ".downsample.1", planes * 4, gain=0.25)
            inplanes = planes * 4
    return sd
