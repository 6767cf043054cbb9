"""A/B timing helper: back-to-back encoder forwards (all heads unless $HEADS), median over groups of 20 forwards.
Environment toggles (EMBCLIP_NO_SIDE, EMBCLIP_NO_TAILFUSE, ...) are read once per process, so run one process per arm."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embclip_b200.encoder import ClipRN50Encoder  # noqa: E402
from embclip_b200.synthetic import synthetic_rn50_state_dict  # noqa: E402

HEADS = tuple(os.environ.get("HEADS", "trunk,avgpool,attnpool").split(","))
enc = ClipRN50Encoder(synthetic_rn50_state_dict(), "cuda:0")
for B in [int(a) for a in sys.argv[1:]] or [256]:
    frames = torch.randn(B, 224, 224, 3, device="cuda")
    out = enc._outputs(B, HEADS)
    for _ in range(20):
        enc(frames, HEADS, out=out)
    torch.cuda.synchronize()
    groups = []
    for _ in range(15):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            enc(frames, HEADS, out=out)
        e1.record()
        torch.cuda.synchronize()
        groups.append(e0.elapsed_time(e1) / 20)
    groups.sort()
    print(f"B={B}: median {groups[len(groups) // 2]:.3f} ms  min {groups[0]:.3f}  max {groups[-1]:.3f}  -> {B / groups[len(groups) // 2] * 1e3:.0f} frames/s")
