"""Pins the AllenAct-side oracle (oracle/allenact_models.py) and the in-tree probe restatement (oracle/probe.py)
with independent checks available offline: torch.nn.GRU step semantics, a float64 closed form of GAE, a hand
computation of the PPO loss, and train.py's documented quirks."""
import pytest
import torch
import torch.nn.functional as F

from oracle import allenact_models as am
from oracle import probe


def test_seq_forward_equals_stepwise_masked_gru():
    torch.manual_seed(0)
    enc = am.RNNStateEncoder(24, 16)
    T, N = 9, 5
    x, h0 = torch.randn(T, N, 24), torch.randn(1, N, 16)
    masks = (torch.rand(T, N, 1) > 0.3).float()
    masks[0, :2] = 0
    out, hT = enc(x, h0, masks)
    # reference semantics: h_t = GRUCell(x_t, masks[t] * h_{t-1}), gate order r, z, n
    W_ih, W_hh, b_ih, b_hh = enc.rnn.weight_ih_l0, enc.rnn.weight_hh_l0, enc.rnn.bias_ih_l0, enc.rnn.bias_hh_l0
    h, outs = h0[0], []
    for t in range(T):
        h = masks[t] * h
        gi, gh = x[t] @ W_ih.t() + b_ih, h @ W_hh.t() + b_hh
        r = torch.sigmoid(gi[:, :16] + gh[:, :16])
        z = torch.sigmoid(gi[:, 16:32] + gh[:, 16:32])
        n = torch.tanh(gi[:, 32:] + r * gh[:, 32:])
        h = (1 - z) * n + z * h
        outs.append(h)
    assert torch.allclose(out, torch.stack(outs), atol=1e-5)
    assert torch.allclose(hT[0], h, atol=1e-5)
    # T == 1 path
    o1, h1 = enc(x[:1], h0, masks[:1])
    assert torch.allclose(o1[0], outs[0], atol=1e-5)


def test_actor_critic_shapes_and_param_count():
    torch.manual_seed(0)
    m = am.ResnetTensorNavActorCritic()
    assert sum(p.numel() for p in m.parameters()) == 3_480_775
    assert tuple(m.state_encoder.rnn.weight_ih_l0.shape) == (1536, 1568)
    T, N = 3, 4
    obs = {m.rgb_uuid: torch.randn(T, N, 2048, 7, 7).relu(), m.goal_uuid: torch.randint(0, 12, (T, N))}
    distr, values, h = m(obs, torch.zeros(1, N, 512), None, torch.ones(T, N, 1))
    assert distr.logits.shape == (T, N, 6) and values.shape == (T, N, 1) and h.shape == (1, N, 512)
    assert torch.equal(distr.mode(), distr.logits.argmax(-1))
    a = torch.randint(0, 6, (T, N))
    assert torch.allclose(distr.log_prob(a), F.log_softmax(distr.logits, -1).gather(-1, a[..., None])[..., 0])


def test_gae_closed_form():
    torch.manual_seed(1)
    T, N, gamma, tau = 12, 3, 0.99, 0.95
    r, v = torch.randn(T, N, 1), torch.randn(T + 1, N, 1)
    masks = (torch.rand(T + 1, N, 1) > 0.2).float()
    nv = torch.randn(N, 1)
    ret = am.compute_returns_gae(r, v, masks, nv, gamma, tau)
    vd, rd, md = v.double().clone(), r.double(), masks.double()
    vd[-1] = nv.double()
    for t in range(T):
        acc, w = torch.zeros(N, 1, dtype=torch.float64), torch.ones(N, 1, dtype=torch.float64)
        for k in range(t, T):                                            # A_t = sum_k (gamma tau)^(k-t) prod(m) delta_k
            delta = rd[k] + gamma * vd[k + 1] * md[k + 1] - vd[k]
            acc += w * delta
            w = w * gamma * tau * md[k + 1]
        assert torch.allclose(ret[t].double(), acc + vd[t], atol=1e-5)
    adv = am.normalized_advantages(ret, v)
    assert abs(adv.mean().item()) < 1e-6 and abs(adv.std().item() - 1) < 1e-3


def test_ppo_loss_by_hand():
    torch.manual_seed(2)
    T, N = 4, 3
    logits = torch.randn(T, N, 6)
    values = torch.randn(T, N, 1)
    batch = dict(actions=torch.randint(0, 6, (T, N)), old_action_log_probs=torch.randn(T, N) * 0.1 - 1.8,
                 values=values + 0.3 * torch.randn(T, N, 1), returns=torch.randn(T, N, 1), norm_adv_targ=torch.randn(T, N, 1))
    total, parts = am.ppo_loss(am.CategoricalDistr(logits=logits), values, batch)
    lp = F.log_softmax(logits, -1)
    logp = lp.gather(-1, batch["actions"][..., None])
    ratio = (logp - batch["old_action_log_probs"][..., None]).exp()
    act = -torch.min(ratio * batch["norm_adv_targ"], ratio.clamp(0.9, 1.1) * batch["norm_adv_targ"]).mean()
    vc = batch["values"] + (values - batch["values"]).clamp(-0.1, 0.1)
    val = 0.5 * torch.max((values - batch["returns"]) ** 2, (vc - batch["returns"]) ** 2).mean()
    ent = (lp.exp() * lp).sum(-1).mean()                                # = -entropy
    assert torch.allclose(parts["action"], act, atol=1e-6) and torch.allclose(parts["value"], val, atol=1e-6)
    assert torch.allclose(parts["entropy"], ent, atol=1e-6)
    assert torch.allclose(total, act + 0.5 * val + 0.01 * ent, atol=1e-6)


def test_ppo_update_runs_and_clips():
    torch.manual_seed(3)
    m = am.ResnetTensorNavActorCritic()
    opt = torch.optim.Adam(m.parameters(), lr=3e-4)
    T, N = 4, 3
    feats = torch.randn(T, N, 2048, 7, 7).relu()
    goals = torch.randint(0, 12, (T, N))
    masks = torch.ones(T, N, 1)
    masks[0, 0] = 0
    with torch.no_grad():
        d, v, _ = m({m.rgb_uuid: feats, m.goal_uuid: goals}, torch.zeros(1, N, 512), None, masks)
        actions = d.sample()
        old_lp = d.log_prob(actions)
    ret = am.compute_returns_gae(0.1 * torch.randn(T, N, 1), torch.cat([v, v[-1:]]), torch.ones(T + 1, N, 1), v[-1])
    rollout = dict(features=feats, goals=goals, masks=masks, actions=actions, old_action_log_probs=old_lp, values=v,
                   returns=ret[:-1], norm_adv_targ=am.normalized_advantages(ret, torch.cat([v, v[-1:]])), memory=torch.zeros(1, N, 512))
    seen = []
    before = [p.clone() for p in m.parameters()]
    info = am.ppo_update(m, opt, rollout, update_repeats=2, grad_hook=lambda ps: seen.append(len(ps)))
    assert len(seen) == 2 and all(torch.isfinite(torch.tensor(list(info.values()))))
    assert any(not torch.equal(a, b) for a, b in zip(before, m.parameters()))


def test_probe_matches_train_py_quirks():
    torch.manual_seed(1)                                               # pl.seed_everything(1), train.py:117
    B = 32                                                             # BASELINE.json config 1
    for emb, dim in (("clip_avgpool", 2048), ("clip_attnpool", 1024)):
        enc = probe.LinearEncoder(emb, "object_presence")
        x, y = torch.randn(B, dim), (torch.rand(B, 52) < 0.1).long()
        out = enc(x)
        assert out.shape == (B, 52) and out.min() >= 0 and out.max() <= 1
        assert torch.allclose(enc.compute_loss((x, y)), F.binary_cross_entropy(torch.sigmoid(enc.model[0](x)), y.float()))
    fs = probe.LinearEncoder("clip_avgpool", "free_space")
    x, y = torch.randn(B, 2048), torch.randint(0, 15, (B,))
    loss = fs.compute_loss((x, y))
    assert y.max() <= 10                                               # clamped in place (train.py:65)
    sm = torch.softmax(fs.model[0](x), 1)
    assert torch.allclose(loss, F.cross_entropy(sm, y))                # cross-entropy of a softmax (train.py:35,78)
    loc = probe.LinearEncoder("clip_avgpool", "object_localization")
    x, y = torch.randn(B, 2048, 7, 7), (torch.rand(B, 9, 52) < 0.1).long()
    assert loc(x).shape == (B, 52, 9)
    l, met = loc.compute_loss((x, y), eval=True)
    assert torch.isfinite(l) and 0 <= met["accuracy"] <= 1
    re = probe.LinearEncoder("clip_attnpool", "reachability")
    x, idx, y = torch.randn(B, 1024), torch.randint(0, 110, (B,)), (torch.rand(B) < 0.5)
    l = re.compute_loss((x, (idx, y)))
    assert torch.allclose(l, F.binary_cross_entropy(re(x)[range(B), idx.tolist()], y.float()))
    opt = torch.optim.Adam(fs.parameters(), lr=1e-3)                   # train.py:111-113, lr :137
    l0 = probe.probe_train_step(fs, opt, (torch.randn(B, 2048), torch.randint(0, 11, (B,))))
    assert l0 > 0


# --------------------------------------------------------------------------------------------------------------------
# RolloutStorage bookkeeping on the CPU (the product class is pure torch indexing apart from compute_returns / feature
# packing, which need the CUDA library): insert / recurrent_generator / after_update against oracle/allenact_storage.py with
# a stub in place of the actor-critic (fp32 features, packed_features=False).
# --------------------------------------------------------------------------------------------------------------------
class _StubAC:
    resnet_uuid, goal_uuid, hidden_size = "rgb_clip_resnet", "goal_object_type_ind", 16
    resnet_tensor_shape = (8, 2, 2)

    def __init__(self):
        self.flat_params = torch.zeros(1)

    def _recurrent_memory_specification(self):
        return dict(rnn=((("layer", 1), ("sampler", None), ("hidden", self.hidden_size)), torch.float32))


def test_rollout_storage_bookkeeping_cpu():
    from embclip_b200.actor_critic import Memory
    from embclip_b200.storage import RolloutStorage
    from oracle.allenact_storage import RefRolloutStorage
    T, N = 4, 5
    ours = RolloutStorage(T, N, _StubAC(), packed_features=False, seed=7)
    ref = RefRolloutStorage(T, N, hidden=16, seed=7)
    g = torch.Generator().manual_seed(0)
    for rollout in range(2):
        for t in range(T):
            obs = {"rgb_clip_resnet": torch.randn(N, 8, 2, 2, generator=g), "goal_object_type_ind": torch.randint(0, 12, (N,), generator=g)}
            mem = torch.randn(1, N, 16, generator=g)
            a, lp, v = torch.randint(0, 6, (N, 1), generator=g), torch.randn(N, 1, generator=g), torch.randn(N, 1, generator=g)
            r, m = torch.randn(N, 1, generator=g), (torch.rand(N, 1, generator=g) > 0.2).float()
            ref.insert(obs, mem, a, lp, v, r, m)
            ours.insert(obs, Memory(rnn=(mem, 1)), a, lp, v, r, m)
        assert ours.step == ref.step == 0
        for k in ("actions", "prev_actions", "masks", "action_log_probs", "value_preds", "rewards"):
            assert torch.equal(getattr(ours, k), getattr(ref, k)), k
        assert torch.equal(ours.memory.tensor("rnn"), ref.memory["rnn"])
        ref.compute_returns(torch.zeros(N, 1), True, 0.99, 0.95)
        ours.returns.copy_(ref.returns)                      # (compute_returns is the CUDA kernel: covered by tests/test_storage.py)
        adv = ref.returns[:-1] - ref.value_preds[:-1]
        for nmb in (1, 2, 5):
            ob = list(ours.recurrent_generator(adv, adv.mean(), adv.std(), nmb))
            rb = list(ref.recurrent_generator(adv, adv.mean(), adv.std(), nmb))
            assert [b["samplers"] for b in ob] == [b["samplers"] for b in rb]
            for o, r_ in zip(ob, rb):
                for k in ("actions", "prev_actions", "values", "returns", "masks", "old_action_log_probs", "adv_targ", "norm_adv_targ"):
                    assert torch.equal(o[k], r_[k]), k
                assert torch.equal(o["memory"].tensor("rnn"), r_["memory"]["rnn"])
                for k in o["observations"]:
                    assert torch.equal(o["observations"][k], r_["observations"][k]), k
        ours.after_update()
        ref.after_update()
        assert torch.equal(ours.masks[0], ref.masks[0]) and torch.equal(ours.prev_actions[0], ref.prev_actions[0])
        assert torch.equal(ours.observations["rgb_clip_resnet"][0], ref.observations["rgb_clip_resnet"][0])
    with pytest.raises(RuntimeError):
        ours.to("cpu").compute_returns(torch.zeros(N, 1))    # no CPU path for the arithmetic
    with pytest.raises(AssertionError):
        list(ours.recurrent_generator(adv, adv.mean(), adv.std(), N + 1))


def test_linear_decay_matches_upstream_definition():
    from embclip_b200.actor_critic import LinearDecay
    ld = LinearDecay(steps=1000, startp=1.0, endp=0.0)
    assert ld(0) == 1.0 and ld(250) == 0.75 and ld(1000) == 0.0 and ld(5000) == 0.0 and ld(-3) == 1.0
    ld2 = LinearDecay(steps=10, startp=0.5, endp=0.1)
    assert abs(ld2(5) - 0.3) < 1e-12
