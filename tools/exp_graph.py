"""Experiment: is the small-batch encoder forward host-bound?  Times B-frame forwards launched eagerly vs replayed from a
CUDA graph (torch.cuda.CUDAGraph around the C-ABI call)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embclip_b200.encoder import ClipRN50Encoder  # noqa: E402
from embclip_b200.synthetic import synthetic_rn50_state_dict  # noqa: E402

HEADS = tuple(os.environ.get("HEADS", "trunk,avgpool,attnpool").split(","))
enc = ClipRN50Encoder(synthetic_rn50_state_dict(), "cuda:0")
for B in [int(a) for a in sys.argv[1:]] or [8, 60, 256]:
    frames = torch.randn(B, 224, 224, 3, device="cuda")
    out = enc._outputs(B, HEADS)
    for _ in range(5):
        enc(frames, HEADS, out=out)
    torch.cuda.synchronize()
    n = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        enc(frames, HEADS, out=out)
    t_host = (time.perf_counter() - t0) / n * 1e3
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / n
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        enc(frames, HEADS, out=out)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            enc(frames, HEADS, out=out)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    graph = e0.elapsed_time(e1) / n
    print(f"B={B}: eager {eager:.3f} ms/forward (host enqueue {t_host:.3f} ms), graph replay {graph:.3f} ms  -> {B / graph * 1e3:.0f} frames/s")
