"""Device timing (CUDA events) of the PPO-update pieces at BASELINE config 3 shape (T=128, N=60 by default)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embclip_b200 import _lib
from embclip_b200.actor_critic import ResnetTensorNavActorCritic, PPOTrainer, PackedFeatures


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    dev = "cuda"
    lib = _lib.load()
    m = ResnetTensorNavActorCritic(device=dev, seed=0)
    plan = m._plan
    F = T * N
    feats = torch.randn(F, 2048, 49, device=dev).relu_()
    pf16 = torch.empty(F * 49, 2048, dtype=torch.float16, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    res["pack_features_ms"] = timed(lambda: lib.embclip_ac_pack_features(plan._h, feats.data_ptr(), F, pf16.data_ptr(), st))
    del feats
    goals = torch.randint(0, 12, (T, N), device=dev)
    masks = (torch.rand(T, N, device=dev) > 0.01).float()
    h0 = torch.zeros(N, 512, device=dev)
    ws = m._workspace(T, N)
    P = m.flat_params.data
    logits = torch.empty(T, N, 6, device=dev); values = torch.empty(T, N, device=dev)
    fwd = lambda sv: lib.embclip_ac_forward(plan._h, P.data_ptr(), pf16.data_ptr(), goals.data_ptr(), masks.data_ptr(), h0.data_ptr(), T, N,
                                            logits.data_ptr(), values.data_ptr(), None, ws.data_ptr(), ws.numel(), sv, st)
    assert fwd(1) == 0, lib.embclip_last_error()
    res["forward_train_ms"] = timed(lambda: fwd(1))
    res["forward_infer_ms"] = timed(lambda: fwd(0))
    actions = torch.randint(0, 6, (T, N), device=dev)
    olp = torch.log_softmax(logits, -1).gather(-1, actions[..., None])[..., 0].contiguous()
    adv = torch.randn(T, N, device=dev); rets = values + torch.randn(T, N, device=dev); ov = values.clone()
    sums = torch.zeros(3, device=dev); grads = torch.zeros_like(P)
    loss = lambda: lib.embclip_ac_ppo_loss(plan._h, P.data_ptr(), T, N, actions.data_ptr(), olp.data_ptr(), adv.data_ptr(), ov.data_ptr(),
                                           rets.data_ptr(), 0.1, 0.5, 0.01, 1.0 / F, logits.data_ptr(), values.data_ptr(), sums.data_ptr(),
                                           ws.data_ptr(), ws.numel(), st)
    bwd = lambda: lib.embclip_ac_backward(plan._h, P.data_ptr(), pf16.data_ptr(), goals.data_ptr(), masks.data_ptr(), h0.data_ptr(), T, N,
                                          None, None, None, grads.data_ptr(), ws.data_ptr(), ws.numel(), st)
    fwd(1); assert loss() == 0 and bwd() == 0, lib.embclip_last_error()
    res["ppo_loss_ms"] = timed(loss)
    res["backward_ms"] = timed(lambda: (grads.zero_(), bwd()))
    tr = PPOTrainer(m, update_repeats=4)
    ro = dict(features=PackedFeatures(pf16, T, N), goals=goals, masks=masks, memory=h0, actions=actions, old_action_log_probs=olp,
              values=ov, returns=rets, norm_adv_targ=adv)
    res["update_4_passes_ms"] = timed(lambda: tr.update(ro), iters=3, warm=1)
    # GRU alone
    gi = torch.randn(T, N, 1536, device=dev)
    out = torch.empty(T, N, 512, device=dev); sv = [torch.empty(T, N, 512, device=dev) for _ in range(4)]
    scratch = torch.zeros(64, dtype=torch.int32, device=dev)
    w_hh = torch.randn(1536, 512, device=dev) * 0.04; b_hh = torch.zeros(1536, device=dev)
    res["gru_forward_ms"] = timed(lambda: lib.embclip_gru_forward(gi.data_ptr(), w_hh.data_ptr(), b_hh.data_ptr(), h0.data_ptr(), masks.data_ptr(), None, T, N, 512,
                                                                 out.data_ptr(), *[s.data_ptr() for s in sv], scratch.data_ptr(), st))
    dout = torch.randn(T, N, 512, device=dev) * 1e-4
    dgi = torch.empty(T, N, 1536, device=dev); dgh = torch.empty(T, N, 1536, device=dev); hm = torch.empty(T, N, 512, device=dev, dtype=torch.float16)
    res["gru_backward_ms"] = timed(lambda: lib.embclip_gru_backward(w_hh.data_ptr(), h0.data_ptr(), masks.data_ptr(), out.data_ptr(), *[s.data_ptr() for s in sv],
                                                                   dout.data_ptr(), None, None, T, N, 512, dgi.data_ptr(), dgh.data_ptr(), hm.data_ptr(), None,
                                                                   None, scratch.data_ptr(), st))
    res["frames"] = F
    res["hbm_floor_update_pass_ms"] = 2 * F * 49 * 2048 * 2 / 6.5329e12 * 1e3     # features read by forward GEMM and by dW1
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"profile_ac_T{T}_N{N}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
