"""Rollout-sized forwards: ms per encode_rows() call (trunk + row export, PDL on, back to back) at a few batch sizes, and the
packed collect step (encoder + act).  A/B switches are environment variables read by the library (EMBCLIP_NO_NARROW, ...)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embclip_b200.actor_critic import PPOTrainer, ResnetTensorNavActorCritic  # noqa: E402
from embclip_b200.encoder import ClipRN50Encoder  # noqa: E402
from embclip_b200.harness import SyntheticPPOStep  # noqa: E402
from embclip_b200.synthetic import synthetic_rn50_state_dict  # noqa: E402

enc = ClipRN50Encoder(synthetic_rn50_state_dict(), "cuda:0")
res = {}
for B in [int(a) for a in sys.argv[1:]] or [7, 8, 15, 30, 60]:
    frames = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty(B * 49, 2048, dtype=torch.float16, device="cuda")
    for _ in range(5):
        enc.encode_rows(frames, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        enc.encode_rows(frames, out=out)
    e1.record()
    torch.cuda.synchronize()
    res[f"encode_rows_ms_B{B}"] = e0.elapsed_time(e1) / n
    model = ResnetTensorNavActorCritic(device="cuda:0", seed=1)
    st = SyntheticPPOStep(enc, model, PPOTrainer(model), T=32, N=B)
    st.collect(lambda t: frames)
    e0.record()
    st.collect(lambda t: frames)
    e1.record()
    torch.cuda.synchronize()
    res[f"collect_step_ms_B{B}"] = e0.elapsed_time(e1) / 32
print(json.dumps(res, indent=1))
