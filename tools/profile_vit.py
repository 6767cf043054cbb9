"""One ViT-B/32 forward at B=512 for ncu launch lists (and a CUDA-event total)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embclip_b200.vit import ClipViTEncoder
from embclip_b200.synthetic import synthetic_clip_vit_b32_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
enc = ClipViTEncoder(synthetic_clip_vit_b32_state_dict(seed=1234), "cuda:0")
x = torch.randn(B, 224, 224, 3, device="cuda")
for _ in range(2):
    enc(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); enc(x); e1.record(); torch.cuda.synchronize()
print(f"B={B} forward {e0.elapsed_time(e1):.3f} ms")
