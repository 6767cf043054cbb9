"""Layer-by-layer forward / backward comparison of the actor-critic plan against the oracle (intermediates via
embclip_ac_act_info)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from embclip_b200 import _lib
from embclip_b200.actor_critic import ResnetTensorNavActorCritic
from oracle.allenact_models import ResnetTensorNavActorCritic as RefAC, ppo_loss
import test_actor_critic_gpu as tt

def rel(a, b):
    a, b = a.detach().float().cpu().flatten(), b.detach().float().cpu().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

def main(T, N):
    torch.manual_seed(11)
    ref = RefAC()
    with torch.no_grad():
        for n_, p_ in ref.named_parameters():
            if "bias" in n_: p_.normal_(0, 0.05)
        ref.actor.linear.weight.mul_(30.0)
    ours = ResnetTensorNavActorCritic(device="cuda:0"); ours.load_state_dict(ref.state_dict())
    ro = tt._rollout(T, N, seed=100 + T); batch = tt._loss_batch(ref, ro, seed=7)
    keep = {}
    enc = ref.goal_visual_encoder
    def hook(name):
        def f(mod, inp, out):
            out.retain_grad(); keep[name] = out
        return f
    enc.resnet_compressor[1].register_forward_hook(hook("compress0"))
    enc.resnet_compressor[3].register_forward_hook(hook("compress2"))
    enc.target_obs_combiner[1].register_forward_hook(hook("combine0"))
    enc.target_obs_combiner[2].register_forward_hook(hook("x_nchw"))
    def rnn_hook(m, i, o):
        o[0].retain_grad(); keep.setdefault("h_segs", []).append(o[0])
    ref.state_encoder.rnn.register_forward_hook(rnn_hook)
    ref.actor.linear.register_forward_hook(hook("raw_logits"))
    ref.zero_grad()
    distr, v, _ = tt._ref_forward(ref, ro)
    total, parts = ppo_loss(distr, v, batch); total.backward()
    lib = _lib.load(); plan = ours._plan; dev = "cuda"
    st = torch.cuda.current_stream().cuda_stream
    pf = ours.pack_features(ro["features"].to(dev)); ws = ours._workspace(T, N); P = ours.flat_params.data
    goals = ro["goals"].to(dev).contiguous(); masks = ro["masks"][..., 0].to(dev).contiguous(); h0 = ro["memory"][0].to(dev).contiguous()
    logits = torch.empty(T, N, 6, device=dev); values = torch.empty(T, N, device=dev); sums = torch.zeros(3, device=dev); grads = torch.zeros_like(P)
    d = lambda k: batch[k].reshape(T, N).to(dev).contiguous()
    a_, olp, adv, ov, rt = d("actions"), d("old_action_log_probs"), d("norm_adv_targ"), d("values"), d("returns")
    assert lib.embclip_ac_forward(plan._h, P.data_ptr(), pf.data.data_ptr(), goals.data_ptr(), masks.data_ptr(), h0.data_ptr(), T, N, logits.data_ptr(), values.data_ptr(), None, ws.data_ptr(), ws.numel(), 1, st) == 0
    assert lib.embclip_ac_ppo_loss(plan._h, P.data_ptr(), T, N, a_.data_ptr(), olp.data_ptr(), adv.data_ptr(), ov.data_ptr(), rt.data_ptr(), 0.1, 0.5, 0.01, 1.0 / (T * N), logits.data_ptr(), values.data_ptr(), sums.data_ptr(), ws.data_ptr(), ws.numel(), st) == 0
    assert lib.embclip_ac_backward(plan._h, P.data_ptr(), pf.data.data_ptr(), goals.data_ptr(), masks.data_ptr(), h0.data_ptr(), T, N, None, None, None, grads.data_ptr(), ws.data_ptr(), ws.numel(), st) == 0
    torch.cuda.synchronize()
    acts = {k: v_.cpu().float() for k, v_ in ours.activations(T, N).items()}
    F = T * N
    S = acts["loss_scale"][0, 0].item()
    print(f"=== T={T} N={N}  loss scale {S:g} (1/S {acts['loss_scale'][0,1].item():g})")
    nhwc = lambda t: t.reshape(F, t.shape[-3], 49).permute(0, 2, 1).reshape(F * 49, -1)     # [F,C,7,7] -> [F*49, C]
    for name in ("compress0", "compress2", "combine0"):
        print(f" fwd {name:12s} {rel(acts[name], nhwc(keep[name])):.2e}   bwd d_{name:12s} {rel(acts['d_' + name] / S, nhwc(keep[name].grad)):.2e}")
    xr = keep["x_nchw"]
    print(f" fwd x            {rel(acts['x'].reshape(F * 49, 32), nhwc(xr)):.2e}   bwd d_x {rel(acts['d_x'].reshape(F * 49, 32) / S, nhwc(xr.grad)):.2e}")
    h_ref = torch.cat(keep["h_segs"], 0).reshape(F, 512); dh_ref = torch.cat([h.grad for h in keep["h_segs"]], 0).reshape(F, 512)
    print(f" fwd h            {rel(acts['h'], h_ref):.2e}   bwd d_h {rel(acts['d_h'], dh_ref):.2e}")
    lr = torch.log_softmax(logits, -1).cpu()
    print(f" fwd logp {rel(lr, distr.logits):.2e} values {rel(values, v[..., 0]):.2e}")
    print(f" d_gi_f16 vs d_gi: {rel(acts['d_gi_f16'] / S, acts['d_gi']):.2e};  d_gi absmax {acts['d_gi'].abs().max():.3e} median {acts['d_gi'].abs().median():.3e}")
    dl_ref = keep["raw_logits"].grad.reshape(F, 6); dl = acts["d_logits"]
    row_err = (dl - dl_ref).norm(dim=1) / dl_ref.norm(dim=1).clamp_min(1e-30)
    badrows = (row_err > 1e-3).nonzero().flatten().tolist()
    print(f" d_logits rel {rel(dl, dl_ref):.2e}; rows off by > 1e-3: {badrows[:10]} of {F}")
    with torch.no_grad():
        lp = distr.log_prob(batch["actions"]).reshape(F); ratio = torch.exp(lp - batch["old_action_log_probs"].reshape(F))
    for r_ in badrows[:5]:
        print(f"   row {r_}: ratio {ratio[r_].item():.7f} adv {batch['norm_adv_targ'].reshape(F)[r_].item():.4f} ours {dl[r_].tolist()} ref {dl_ref[r_].tolist()}")
    # ReLU-mask-aligned comparison: gradient w.r.t. the pre-activation = grad(post) * (post > 0)
    for name in ("compress0", "compress2", "combine0"):
        gref = nhwc(keep[name].grad) * (nhwc(keep[name]) > 0)
        ours_ = acts['d_' + name] / S
        flips = ((acts[name] > 0) != (nhwc(keep[name]) > 0)).float().mean().item()
        agree = (acts[name] > 0) == (nhwc(keep[name]) > 0)
        print(f" bwd d_{name} (pre-ReLU) rel {rel(ours_, gref):.2e}; mask flips {flips:.2e}; rel on agreeing elements {rel(ours_ * agree, gref * agree):.2e}")
    refp = dict(ref.named_parameters())
    for name, shape, off, n in plan.params:
        print(f" grad {name:55s} {rel(grads[off:off + n].view(shape), refp[name].grad):.2e}   |ref| {refp[name].grad.norm():.2e}")

main(6, 5)
main(16, 60)
