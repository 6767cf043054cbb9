"""Three encoder forwards at batch B (default 256), all heads: the workload ncu captures are taken on."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embclip_b200.encoder import ClipRN50Encoder  # noqa: E402
from embclip_b200.synthetic import synthetic_rn50_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
enc = ClipRN50Encoder(synthetic_rn50_state_dict(), "cuda:0")
frames = torch.randn(B, 224, 224, 3, device="cuda")
for _ in range(3):
    enc(frames, ("trunk", "avgpool", "attnpool"))
torch.cuda.synchronize()
