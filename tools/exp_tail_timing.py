"""Experiment: bneck_tail variants timed alone at the layer-1 shape (M = 256 x 56 x 56)."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from embclip_b200 import _lib
lib = _lib.load()

M, n1 = 256 * 56 * 56, int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
y2 = rn(M, 64).relu().half(); w3 = (rn(256, 64) / 8).half(); b3 = rn(256); res = rn(M, 256).relu().half()
w1 = (rn(n1, 256) / 16).half(); b1 = rn(n1)
out = torch.empty(M, 256, device="cuda", dtype=torch.float16); y1 = torch.empty(M, n1, device="cuda", dtype=torch.float16)
st = torch.cuda.current_stream().cuda_stream
p = lambda t: C.c_void_p(t.data_ptr())
POOL = len(sys.argv) > 2
pool = torch.empty(M // 4, 256, device="cuda", dtype=torch.float16)
def run():
    if POOL:
        rc = lib.embclip_bneck_tail_pool_f16(p(y2), p(w3), p(b3), p(res), p(pool), 1, 56, p(w1), p(b1), p(y1), M, n1, C.c_void_p(st))
    else:
        rc = lib.embclip_bneck_tail_f16(p(y2), None, p(w3), p(b3), p(res), p(out), p(w1), p(b1), p(y1), M, n1, C.c_void_p(st))
    assert rc == 0, lib.embclip_last_error()
for _ in range(3): run()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("pool" if POOL else "plain", "n1", n1, "kernel ms min", min(ts), "median", sorted(ts)[5])
