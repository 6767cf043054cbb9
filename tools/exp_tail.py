"""Experiment: per-tile time of bneck_tail when its working set is L2-resident (k tiles per CTA, k = 1..3) vs HBM-resident."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embclip_b200 import _lib
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
def run(M, n1=64, reps=200):
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    y2, res = rn(M, 64).relu().half(), rn(M, 256).relu().half()
    w3, b3, w1, b1 = (rn(256, 64) / 8).half(), rn(256), (rn(n1, 256) / 16).half(), rn(n1)
    out, y1 = torch.empty(M, 256, device="cuda", dtype=torch.float16), torch.empty(M, n1, device="cuda", dtype=torch.float16)
    call = lambda: _lib.check(lib.embclip_bneck_tail_f16(y2.data_ptr(), None, w3.data_ptr(), b3.data_ptr(), res.data_ptr(), out.data_ptr(),
                                                       w1.data_ptr(), b1.data_ptr(), y1.data_ptr(), M, n1, st))
    for _ in range(10): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for k in (1, 2, 3, 4, 8, 16, 42):
    M = 128 * 148 * k
    us = run(M, reps=200 if k < 16 else 50)
    print(f"k={k:2d} tiles/CTA  M={M:7d}  bytes {M * 1280 / 1e6:6.1f} MB  {us:7.1f} us  -> {us / k:5.2f} us/tile  {M * 1280 / us / 1e6:5.2f} TB/s")
