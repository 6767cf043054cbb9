"""Generates the committed golden fixtures from the fp32 oracle (run on CPU, in this container):
    python tests/golden/make_golden.py
The reference itself cannot produce them (its CLIP / AllenAct dependencies and weights are absent offline,
SURVEY.md section 8c), so the vectors pin the ORACLE restatement with seeded synthetic weights."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from conftest import synthetic_frames  # noqa: E402
from oracle.clip_model import build_rn50, freeze_model, init_synthetic_rn50_visual  # noqa: E402


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    m = freeze_model(init_synthetic_rn50_visual(build_rn50().visual, seed=1234))
    frames = synthetic_frames(2, seed=0)
    with torch.no_grad():
        trunk = m.trunk(frames.permute(0, 3, 1, 2).contiguous())
        attn = m.attnpool(trunk)
    out = {
        "frames_probe": frames[:, ::37, ::41].contiguous(),
        "trunk_every16": trunk[:, ::16].contiguous(),
        "avgpool": trunk.mean(dim=(2, 3)),
        "attnpool": attn.clone(),
        "weights_seed": 1234, "frames_seed": 0,
        "weight_probe": m.layer3[2].conv2.weight[:4, :4, 1, 1].clone(),
    }
    torch.save(out, os.path.join(HERE, "rn50_b2_seed0.pt"))
    print({k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in out.items()})


def main_vit():
    from oracle.clip_model import build_vit_b32, init_synthetic_transformer
    from test_vit_gpu import synthetic_prompts
    torch.manual_seed(0)
    m = freeze_model(init_synthetic_transformer(build_vit_b32(), seed=1234))
    frames, tokens = synthetic_frames(2, seed=0), synthetic_prompts(12, seed=0)
    with torch.no_grad():
        img = m.encode_image(frames.permute(0, 3, 1, 2).contiguous())
        txt = m.encode_text(tokens)
        logits, _ = m(frames.permute(0, 3, 1, 2).contiguous(), tokens)
    out = {"image_features": img, "text_features": txt, "logits_per_image": logits, "tokens": tokens, "weights_seed": 1234, "frames_seed": 0}
    torch.save(out, os.path.join(HERE, "vit_b32_b2_seed0.pt"))
    print({k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
    main_vit()
