"""A/B timing of the ViT-B/32 image tower (B frames, default 512): median over groups of 10 forwards."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embclip_b200.vit import ClipViTEncoder
from embclip_b200.synthetic import synthetic_clip_vit_b32_state_dict
enc = ClipViTEncoder(synthetic_clip_vit_b32_state_dict(seed=1234), "cuda:0")
for B in [int(a) for a in sys.argv[1:]] or [512]:
    x = torch.randn(B, 224, 224, 3, device="cuda")
    for _ in range(5):
        enc(x)
    torch.cuda.synchronize()
    g = []
    for _ in range(9):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            enc(x)
        e1.record(); torch.cuda.synchronize()
        g.append(e0.elapsed_time(e1) / 10)
    g.sort()
    print(f"ViT-B/32 B={B}: median {g[4]:.3f} ms  min {g[0]:.3f} -> {B / g[4] * 1e3:.0f} frames/s")
