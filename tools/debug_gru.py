"""Localise GRU kernel mismatches: per-step / per-sampler / per-unit-block error maps vs torch.nn.GRU."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embclip_b200 import _lib
from oracle.allenact_models import RNNStateEncoder

lib = _lib.load()
st = lambda: torch.cuda.current_stream().cuda_stream

def case(T, N, H, seed=0):
    torch.manual_seed(seed)
    I = 40
    enc = RNNStateEncoder(I, H)
    with torch.no_grad():
        enc.rnn.bias_ih_l0.normal_(0, 0.1); enc.rnn.bias_hh_l0.normal_(0, 0.1)
    x = torch.randn(T, N, I, requires_grad=True)
    h0 = torch.randn(1, N, H) * 0.5
    masks = (torch.rand(T, N, 1) > 0.15).float(); masks[0, :max(1, N // 3)] = 0
    out_ref, hT = enc(x, h0, masks)
    gi_ref = (x @ enc.rnn.weight_ih_l0.t() + enc.rnn.bias_ih_l0).detach()
    gi_ref.requires_grad_(True)
    # manual recurrence with autograd on gi for dgi reference
    W_hh, b_hh = enc.rnn.weight_hh_l0.detach(), enc.rnn.bias_hh_l0.detach()
    h = h0[0]; outs = []; ghs = []
    for t in range(T):
        hm = masks[t] * h
        gh = hm @ W_hh.t() + b_hh; gh.retain_grad(); ghs.append(gh)
        g = gi_ref[t]
        r = torch.sigmoid(g[:, :H] + gh[:, :H]); z = torch.sigmoid(g[:, H:2*H] + gh[:, H:2*H]); n = torch.tanh(g[:, 2*H:] + r * gh[:, 2*H:])
        h = (1 - z) * n + z * hm; outs.append(h)
    o2 = torch.stack(outs)
    dout = torch.randn(T, N, H) / (T * N); dhT = torch.randn(N, H) / N
    ((o2 * dout).sum() + (o2[-1] * dhT).sum()).backward()
    dgi_ref = gi_ref.grad; dgh_ref = torch.stack([g.grad for g in ghs])
    dev = "cuda"
    gi = gi_ref.detach().to(dev).contiguous()
    w, b = W_hh.to(dev).contiguous(), b_hh.to(dev).contiguous()
    h0d, md = h0[0].to(dev).contiguous(), masks[..., 0].to(dev).contiguous()
    out = torch.zeros(T, N, H, device=dev); sv = [torch.zeros(T, N, H, device=dev) for _ in range(4)]
    scratch = torch.zeros(64, dtype=torch.int32, device=dev)
    rc = lib.embclip_gru_forward(gi.data_ptr(), w.data_ptr(), b.data_ptr(), h0d.data_ptr(), md.data_ptr(), None, T, N, H, out.data_ptr(),
                                 *[s.data_ptr() for s in sv], scratch.data_ptr(), st())
    torch.cuda.synchronize()
    print(f"--- T={T} N={N} H={H} fwd rc={rc} {lib.embclip_last_error().decode() if rc else ''}")
    e = (out.cpu() - o2.detach()).abs()
    print(" fwd max err per step:", [f"{v:.1e}" for v in e.amax(dim=(1, 2)).tolist()][:8])
    print(" fwd max err per sampler:", [f"{v:.1e}" for v in e.amax(dim=(0, 2)).tolist()][:40])
    print(" fwd max err per unit-block(8):", [f"{v:.1e}" for v in e.amax(dim=(0, 1)).reshape(-1, 8).amax(1).tolist()][:16])
    doutd, dhTd = dout.to(dev).contiguous(), dhT.to(dev).contiguous()
    dgi = torch.zeros(T, N, 3 * H, device=dev); dgh = torch.zeros(T, N, 3 * H, device=dev)
    hm = torch.zeros(T, N, H, device=dev, dtype=torch.float16); dh0 = torch.zeros(N, H, device=dev)
    o2d = o2.detach().to(dev).contiguous()
    rc = lib.embclip_gru_backward(w.data_ptr(), h0d.data_ptr(), md.data_ptr(), out.data_ptr(), *[s.data_ptr() for s in sv], doutd.data_ptr(),
                                  dhTd.data_ptr(), None, T, N, H, dgi.data_ptr(), dgh.data_ptr(), hm.data_ptr(), dh0.data_ptr(), None, scratch.data_ptr(), st())
    torch.cuda.synchronize()
    print(f" bwd rc={rc} {lib.embclip_last_error().decode() if rc else ''}")
    for name, got, ref in (("dgi", dgi, dgi_ref), ("dgh", dgh, dgh_ref)):
        e = (got.cpu() - ref).abs(); sc = ref.abs().max().item()
        print(f" {name} scale {sc:.2e} max err per step:", [f"{v:.1e}" for v in e.amax(dim=(1, 2)).tolist()][:8])
        print(f" {name} max err per sampler:", [f"{v:.1e}" for v in e.amax(dim=(0, 2)).tolist()][:40])
        print(f" {name} max err per gate:", [f"{v:.1e}" for v in e.amax(dim=(0, 1)).reshape(3, -1).amax(1).tolist()])

for c in [(1, 1, 64), (2, 1, 64), (5, 7, 128), (3, 33, 512), (3, 32, 512), (3, 60, 512), (16, 60, 512)]:
    case(*c)
