"""Experiment: where does one rollout-collection step (B = N samplers) go?  encoder / act() / sampling, CUDA events."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embclip_b200.actor_critic import PPOTrainer, ResnetTensorNavActorCritic  # noqa: E402
from embclip_b200.encoder import ClipRN50Encoder  # noqa: E402
from embclip_b200.harness import SyntheticPPOStep  # noqa: E402
from embclip_b200.synthetic import synthetic_rn50_state_dict  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 60
T = 32
enc = ClipRN50Encoder(synthetic_rn50_state_dict(), "cuda:0")
model = ResnetTensorNavActorCritic(device="cuda:0")
st = SyntheticPPOStep(enc, model, PPOTrainer(model), T=T, N=N, packed_rollout=False)
stp = SyntheticPPOStep(enc, model, PPOTrainer(model), T=T, N=N, packed_rollout=True)
frames = torch.randint(0, 256, (N, 224, 224, 3), dtype=torch.uint8, device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)
tot = {"encode": 0.0, "act": 0.0, "sample": 0.0}
for rep in range(3):
    h = st.memory0[0]
    acc = {"encode": 0.0, "act": 0.0, "sample": 0.0}
    marks = []
    for t in range(T):
        e = [ev() for _ in range(4)]
        e[0].record()
        st.enc.forward(frames, ("trunk",), out={"trunk": st.features[t]})
        e[1].record()
        logits, values, h = st._act(t, h)
        e[2].record()
        probs = torch.softmax(logits, -1)
        a = torch.multinomial(probs, 1, generator=st.gen)[:, 0]
        st.actions[t] = a
        st.log_probs[t] = torch.log_softmax(logits, -1).gather(-1, a[:, None])[:, 0]
        st.values[t, :, 0] = values
        e[3].record()
        marks.append(e)
    torch.cuda.synchronize()
    for e in marks:
        acc["encode"] += e[0].elapsed_time(e[1]); acc["act"] += e[1].elapsed_time(e[2]); acc["sample"] += e[2].elapsed_time(e[3])
    tot = acc
print(f"N={N}: per step  encode {tot['encode'] / T:.3f} ms  act {tot['act'] / T:.3f} ms  sample {tot['sample'] / T:.3f} ms")
e0, e1 = ev(), ev()
e0.record()
st.collect(lambda t: frames)
e1.record()
torch.cuda.synchronize()
print(f"collect(T={T}) {e0.elapsed_time(e1) / T:.3f} ms per step (AllenAct data flow)")
stp.collect(lambda t: frames)
e0.record()
stp.collect(lambda t: frames)
e1.record()
torch.cuda.synchronize()
print(f"collect(T={T}) {e0.elapsed_time(e1) / T:.3f} ms per step (packed rollout)")
