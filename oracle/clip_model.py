"""ORACLE (test infrastructure, not product code): plain-PyTorch fp32 restatement of the
CLIP towers that EmbCLIP's hot path runs.

PARITY UNPINNED against the pinned third-party source: the arithmetic lives in
``openai/CLIP @ 40f5484c1c74edd83cb9cf687c6ab92b28d8b656`` (pin:
/root/reference/primitive_probing/environment.yml:22), which is neither vendored in
/root/reference nor installable here (no network, no weights), and the reference holds no
tests / golden vectors for it.  What IS pinned, by tests/test_oracle_*.py:
  * AttentionPool2d calls ``torch.nn.functional.multi_head_attention_forward`` -- the very
    function clip/model.py calls -- and is cross-checked against a hand-expanded softmax form;
  * the ViT / text towers are cross-checked against the independent implementation in
    ``transformers`` 5.5 (``CLIPVisionModelWithProjection`` / ``CLIPTextModelWithProjection``)
    with weights mapped one to one;
  * the per-layer MAC count reproduces the published 12.22 / 8.82 / 5.96 GFLOP tower figures.

In-tree call sites this module stands in for (the only places /root/reference touches the
encoder):  ``clip.load('RN50')`` primitive_probing/generate_data/thor_image_features.py:57,
``.visual`` :59, trunk forward :109, ``attnpool`` :62/:112, avg-pool head :63-66/:113,
``freeze_model`` :26-33 (BN momentum 0 + eval()); same in reachable_image_features.py:29-36,60-67,88-93.

State-dict key names follow the upstream module so an official ``RN50.pt`` / ``ViT-B-32.pt``
state dict loads unchanged (``visual.conv1.weight``, ``visual.layer1.0.downsample.0.weight``,
``visual.attnpool.q_proj.weight``, ``visual.transformer.resblocks.0.attn.in_proj_weight`` ...).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# ModifiedResNet (clip/model.py: Bottleneck, AttentionPool2d, ModifiedResNet)
# --------------------------------------------------------------------------------------
class Bottleneck(nn.Module):
    """clip/model.py `Bottleneck`.  Anti-aliased: every conv is stride 1,
    a stride>1 block average-pools after conv2 and in front of the 1x1 downsample conv."""

    expansion = 4

    def __init__(self, inplanes: int, planes: int, stride: int = 1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.avgpool = nn.AvgPool2d(stride) if stride > 1 else nn.Identity()
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.stride = stride
        self.downsample = None
        if stride > 1 or inplanes != planes * self.expansion:
            # key "-1" is the parameter-free pool, so the conv / bn keep keys "0" / "1"
            self.downsample = nn.Sequential(OrderedDict([
                ("-1", nn.AvgPool2d(stride)),
                ("0", nn.Conv2d(inplanes, planes * self.expansion, 1, stride=1, bias=False)),
                ("1", nn.BatchNorm2d(planes * self.expansion)),
            ]))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        identity = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.avgpool(out)
        out = self.bn3(self.conv3(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        return self.relu(out + identity)


class AttentionPool2d(nn.Module):
    """clip/model.py `AttentionPool2d` at the pinned commit: all 50 tokens
    are used as queries and row 0 is returned."""

    def __init__(self, spacial_dim: int, embed_dim: int, num_heads: int, output_dim: int = None):
        super().__init__()
        self.positional_embedding = nn.Parameter(
            torch.randn(spacial_dim ** 2 + 1, embed_dim) / embed_dim ** 0.5)
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.c_proj = nn.Linear(embed_dim, output_dim or embed_dim)
        self.num_heads = num_heads

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x.reshape(x.shape[0], x.shape[1], x.shape[2] * x.shape[3]).permute(2, 0, 1)  # (HW)NC
        x = torch.cat([x.mean(dim=0, keepdim=True), x], dim=0)                          # (HW+1)NC
        x = x + self.positional_embedding[:, None, :].to(x.dtype)
        x, _ = F.multi_head_attention_forward(
            query=x, key=x, value=x,
            embed_dim_to_check=x.shape[-1],
            num_heads=self.num_heads,
            q_proj_weight=self.q_proj.weight,
            k_proj_weight=self.k_proj.weight,
            v_proj_weight=self.v_proj.weight,
            in_proj_weight=None,
            in_proj_bias=torch.cat([self.q_proj.bias, self.k_proj.bias, self.v_proj.bias]),
            bias_k=None, bias_v=None, add_zero_attn=False, dropout_p=0.0,
            out_proj_weight=self.c_proj.weight, out_proj_bias=self.c_proj.bias,
            use_separate_proj_weight=True, training=self.training, need_weights=False)
        return x[0]


class ModifiedResNet(nn.Module):
    """clip/model.py `ModifiedResNet`: 3-conv stem + avg-pool, four
    bottleneck stages, attention pool.  RN50 = layers (3,4,6,3), width 64, heads 32, out 1024."""

    def __init__(self, layers=(3, 4, 6, 3), output_dim=1024, heads=32, input_resolution=224, width=64):
        super().__init__()
        self.output_dim = output_dim
        self.input_resolution = input_resolution
        self.conv1 = nn.Conv2d(3, width // 2, kernel_size=3, stride=2, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(width // 2)
        self.conv2 = nn.Conv2d(width // 2, width // 2, kernel_size=3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(width // 2)
        self.conv3 = nn.Conv2d(width // 2, width, kernel_size=3, padding=1, bias=False)
        self.bn3 = nn.BatchNorm2d(width)
        self.avgpool = nn.AvgPool2d(2)
        self.relu = nn.ReLU(inplace=True)
        self._inplanes = width
        self.layer1 = self._make_layer(width, layers[0])
        self.layer2 = self._make_layer(width * 2, layers[1], stride=2)
        self.layer3 = self._make_layer(width * 4, layers[2], stride=2)
        self.layer4 = self._make_layer(width * 8, layers[3], stride=2)
        embed_dim = width * 32
        self.attnpool = AttentionPool2d(input_resolution // 32, embed_dim, heads, output_dim)

    def _make_layer(self, planes: int, blocks: int, stride: int = 1) -> nn.Sequential:
        seq = [Bottleneck(self._inplanes, planes, stride)]
        self._inplanes = planes * Bottleneck.expansion
        for _ in range(1, blocks):
            seq.append(Bottleneck(self._inplanes, planes))
        return nn.Sequential(*seq)

    def stem(self, x: torch.Tensor) -> torch.Tensor:
        for conv, bn in ((self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3)):
            x = self.relu(bn(conv(x)))
        return self.avgpool(x)

    def trunk(self, x: torch.Tensor) -> torch.Tensor:
        """[B,3,R,R] -> [B, 32*width, R/32, R/32]; what the in-tree code gets after it swaps
        attnpool for Identity (thor_image_features.py:67,109)."""
        x = x.type(self.conv1.weight.dtype)
        x = self.stem(x)
        x = self.layer1(x)
        x = self.layer2(x)
        x = self.layer3(x)
        return self.layer4(x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.attnpool(self.trunk(x))


# --------------------------------------------------------------------------------------
# Transformer towers (clip/model.py: LayerNorm, QuickGELU, ResidualAttentionBlock,
# Transformer, VisionTransformer, CLIP)
# --------------------------------------------------------------------------------------
class LayerNorm(nn.LayerNorm):
    """Computes in fp32 whatever the input dtype, casts back (clip/model.py `LayerNorm`)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return super().forward(x.type(torch.float32)).type(x.dtype)


class QuickGELU(nn.Module):
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    """SURVEY.md §8a A6."""

    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask

    def attention(self, x: torch.Tensor) -> torch.Tensor:
        mask = self.attn_mask.to(dtype=x.dtype, device=x.device) if self.attn_mask is not None else None
        return self.attn(x, x, x, need_weights=False, attn_mask=mask)[0]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x + self.attention(self.ln_1(x))
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.resblocks(x)


class VisionTransformer(nn.Module):
    """SURVEY.md §8a A5.  ViT-B/32 = res 224, patch 32, width 768, 12 layers, 12 heads, out 512."""

    def __init__(self, input_resolution=224, patch_size=32, width=768, layers=12, heads=12, output_dim=512):
        super().__init__()
        self.input_resolution, self.output_dim = input_resolution, output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(
            scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.conv1(x)                                   # [B, width, g, g]
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
        cls = self.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device)
        x = torch.cat([cls, x], dim=1) + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = self.transformer(x.permute(1, 0, 2)).permute(1, 0, 2)   # NLD -> LND -> NLD
        x = self.ln_post(x[:, 0, :])
        return x @ self.proj


class CLIP(nn.Module):
    """clip/model.py `CLIP`: visual tower (RN or ViT by `vision_layers`),
    causal text transformer, cosine-similarity logits."""

    def __init__(self, embed_dim: int, image_resolution: int, vision_layers, vision_width: int,
                 vision_patch_size, context_length: int, vocab_size: int, transformer_width: int,
                 transformer_heads: int, transformer_layers: int):
        super().__init__()
        self.context_length = context_length
        if isinstance(vision_layers, (tuple, list)):
            self.visual = ModifiedResNet(layers=tuple(vision_layers), output_dim=embed_dim,
                                         heads=vision_width * 32 // 64,
                                         input_resolution=image_resolution, width=vision_width)
        else:
            self.visual = VisionTransformer(input_resolution=image_resolution, patch_size=vision_patch_size,
                                            width=vision_width, layers=vision_layers,
                                            heads=vision_width // 64, output_dim=embed_dim)
        self.transformer = Transformer(transformer_width, transformer_layers, transformer_heads,
                                       attn_mask=self.build_attention_mask())
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * 2.6592600369327779)   # ln(1/0.07)
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        nn.init.normal_(self.text_projection, std=transformer_width ** -0.5)

    def build_attention_mask(self) -> torch.Tensor:
        mask = torch.full((self.context_length, self.context_length), float("-inf"))
        return mask.triu_(1)

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image: torch.Tensor) -> torch.Tensor:
        return self.visual(image.type(self.dtype))

    def encode_text(self, text: torch.Tensor) -> torch.Tensor:
        x = self.token_embedding(text).type(self.dtype) + self.positional_embedding.type(self.dtype)
        x = self.transformer(x.permute(1, 0, 2)).permute(1, 0, 2)
        x = self.ln_final(x).type(self.dtype)
        # features of the end-of-text token = the highest token id in every prompt
        return x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ self.text_projection

    def forward(self, image: torch.Tensor, text: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        img = self.encode_image(image)
        txt = self.encode_text(text)
        img = img / img.norm(dim=1, keepdim=True)
        txt = txt / txt.norm(dim=1, keepdim=True)
        logits_per_image = self.logit_scale.exp() * img @ txt.t()
        return logits_per_image, logits_per_image.t()


def build_rn50(embed_dim=1024) -> CLIP:
    """CLIP-RN50 hyper-parameters (what `clip.load('RN50')` builds from its state dict)."""
    return CLIP(embed_dim=embed_dim, image_resolution=224, vision_layers=(3, 4, 6, 3), vision_width=64,
                vision_patch_size=None, context_length=77, vocab_size=49408,
                transformer_width=512, transformer_heads=8, transformer_layers=12)


def build_rn50x16(embed_dim=768, image_resolution=384) -> CLIP:
    """CLIP-RN50x16 (`clip.load('RN50x16')`): layers (6,8,18,8), width 96, 48 heads, 384 x 384, embed 768 -- the second encoder
    AllenAct's ClipResNetPreprocessor accepts.  `image_resolution=224` builds the attention pool
    for the 7 x 7 grid AllenAct's 224 x 224 frames produce (a synthetic-weights convenience: the real checkpoint's positional
    embedding is 12 x 12 + 1, and AllenAct never calls its attention pool)."""
    return CLIP(embed_dim=embed_dim, image_resolution=image_resolution, vision_layers=(6, 8, 18, 8), vision_width=96,
                vision_patch_size=None, context_length=77, vocab_size=49408,
                transformer_width=768, transformer_heads=12, transformer_layers=12)


def build_vit_b32(embed_dim=512) -> CLIP:
    return CLIP(embed_dim=embed_dim, image_resolution=224, vision_layers=12, vision_width=768,
                vision_patch_size=32, context_length=77, vocab_size=49408,
                transformer_width=512, transformer_heads=8, transformer_layers=12)


def freeze_model(model: nn.Module) -> nn.Module:
    """Restates thor_image_features.py:26-33 / reachable_image_features.py:29-36."""
    for p in model.parameters():
        p.requires_grad = False
    for m in model.modules():
        if "BatchNorm" in type(m).__name__:
            m.momentum = 0.0
    return model.eval()


# --------------------------------------------------------------------------------------
# Seeded synthetic weights (no checkpoints exist offline)
# --------------------------------------------------------------------------------------
@torch.no_grad()
def init_synthetic_rn50_visual(visual: ModifiedResNet, seed: int = 1234) -> ModifiedResNet:
    """Loads the seeded, well-conditioned synthetic weights the product package generates
    (embclip_b200/synthetic.py -- weight *generation* is shared so both sides see the same tensors; no
    arithmetic of the path is).  strict=True doubles as a check that the restated module has exactly the
    official state-dict keys."""
    from embclip_b200.synthetic import synthetic_rn50_state_dict
    layers = tuple(len(getattr(visual, f"layer{i}")) for i in (1, 2, 3, 4))
    visual.load_state_dict(synthetic_rn50_state_dict(seed, layers=layers, width=visual.conv3.out_channels,
                                                     output_dim=visual.output_dim,
                                                     input_resolution=visual.input_resolution), strict=True)
    return visual


@torch.no_grad()
def init_synthetic_transformer(model: nn.Module, seed: int = 1234) -> nn.Module:
    """Seeded init for the ViT / text towers (std chosen like clip/model.py initialize_parameters)."""
    g = torch.Generator().manual_seed(seed)
    for name, p in model.named_parameters():
        if p.dim() >= 2:
            std = 0.02 if "embedding" in name else p.shape[-1] ** -0.5
            p.copy_(torch.randn(p.shape, generator=g) * std)
        elif "ln_" in name and name.endswith("weight"):
            p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
        elif name.endswith("logit_scale"):
            p.fill_(4.605170185988092)   # ln(100), SURVEY.md §8d config 5
        else:
            p.copy_(0.02 * torch.randn(p.shape, generator=g))
    return model


def count_macs_rn50(visual: ModifiedResNet, res: int = 224) -> dict:
    """Per-stage MACs of one frame, by forward hooks (FLOP identity check)."""
    macs = {}

    def conv_hook(name):
        def fn(m, inp, out):
            k = m.kernel_size[0] * m.kernel_size[1] * m.in_channels // m.groups
            macs[name] = out.numel() * k
        return fn

    hs = [m.register_forward_hook(conv_hook(n)) for n, m in visual.named_modules() if isinstance(m, nn.Conv2d)]
    with torch.no_grad():
        visual.trunk(torch.zeros(1, 3, res, res))
    for h in hs:
        h.remove()
    return macs
