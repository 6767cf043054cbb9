"""Synthetic-rollout harness for BASELINE.json configs 3 / 4 ("RoboTHOR ObjectNav: CLIP-RN50 + GRU actor-critic PPO update
on synthetic rollouts, 128 steps x 60 samplers"): the rollout-collection and update halves of allenact's
``OnPolicyTrainer`` train loop with the simulator replaced by synthetic frames / rewards.

Per PPO step (one rollout of T steps x N samplers on this rank):
  collect   for t in 0..T-1:  ClipResNetPreprocessor.process(frames_t)  -> features[t]  (fp32 [N,2048,7,7], RolloutStorage)
                              actor_critic(features[t], memory, masks[t]) (T = 1 path)  -> action sample, log-prob, value
  returns   RolloutStorage.compute_returns (GAE) + advantage normalisation               (one kernel)
  update    update_repeats x (forward, PPO loss, backward, flat-bucket all-reduce, clip, Adam)

Everything on the device runs in libembclip_b200.so; torch supplies memory, streams, the categorical sampler and
torch.distributed.  Used by bench.py (--workload ppo and the "ppo_step" block) and __graft_entry__.smoke().
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from .actor_critic import PPOTrainer, ResnetTensorNavActorCritic, compute_returns_gae
from .encoder import ClipRN50Encoder


class SyntheticPPOStep:
    def __init__(self, encoder: ClipRN50Encoder, model: ResnetTensorNavActorCritic, trainer: PPOTrainer, T: int = 128, N: int = 60,
                 seed: int = 0, gamma: float = 0.99, tau: float = 0.95, packed_rollout: bool = True):
        """packed_rollout=True (default): the rollout keeps CLIP features as the fp16 pixel rows the update reads
        (``encode_rows`` -> ``act`` -> ``PackedFeatures``), one library call per actor step.  False: the AllenAct data flow
        verbatim -- ``ClipResNetPreprocessor``-style fp32 [N,2048,7,7] features into RolloutStorage, ``forward`` + torch sampling,
        ``pack_features`` before the update.  Both produce bit-identical logits / values (tests/test_actor_critic_gpu.py)."""
        self.packed = bool(packed_rollout)
        self.enc, self.model, self.trainer, self.T, self.N = encoder, model, trainer, T, N
        self.gamma, self.tau = gamma, tau
        dev = model.flat_params.device
        self.device = dev
        g = torch.Generator(device=dev).manual_seed(seed)
        H = model.hidden_size
        C_, Hh, Ww = model.resnet_tensor_shape
        # RolloutStorage tensors ([steps, samplers, ...])
        if self.packed:
            self.features16 = torch.empty(T, N * Hh * Ww, C_, dtype=torch.float16, device=dev)
            self.h_buf = torch.zeros(2, N, H, device=dev)
        else:
            self.features = torch.empty(T, N, C_, Hh, Ww, dtype=torch.float32, device=dev)
        self.goals = torch.randint(0, model._plan.cfg["num_goals"], (T, N), device=dev, generator=g)
        self.masks = (torch.rand(T + 1, N, 1, device=dev, generator=g) > 0.01).float()      # ~1 % episode boundaries
        self.masks[0, : N // 4] = 0
        self.rewards = 0.1 * torch.randn(T, N, 1, device=dev, generator=g)
        self.memory0 = torch.zeros(1, N, H, device=dev)
        self.actions = torch.zeros(T, N, dtype=torch.int64, device=dev)
        self.log_probs = torch.zeros(T, N, device=dev)
        self.values = torch.zeros(T + 1, N, 1, device=dev)
        self.gen = g
        self.kernel_launches = 0

    # ---------------------------------------------------------------- rollout collection
    def _act(self, t: int, h: torch.Tensor):
        with torch.no_grad():
            logits, values, h_new = self.model.forward_tensors(self.features[t:t + 1], self.goals[t:t + 1], h, self.masks[t:t + 1])
        return logits[0], values[0], h_new

    def _collect_packed(self, frames_at: Callable[[int], torch.Tensor]) -> None:
        T, N = self.T, self.N
        u = torch.rand(T, N, device=self.device, generator=self.gen)
        h = self.memory0[0]
        for t in range(T):
            self.enc.encode_rows(frames_at(t), out=self.features16[t])
            _, _, _, h, _ = self.model.act(self.features16[t], self.goals[t], self.masks[t], h, u[t], actions=self.actions[t],
                                           action_log_probs=self.log_probs[t], values=self.values[t, :, 0],
                                           memory_out=self.h_buf[t & 1])
        with torch.no_grad():
            from .actor_critic import PackedFeatures
            _, v_next, _ = self.model.forward_tensors(PackedFeatures(self.features16[T - 1], 1, N), self.goals[T - 1:T], h, self.masks[T:T + 1])
        self.values[T, :, 0] = v_next[0]

    def collect(self, frames_at: Callable[[int], torch.Tensor]) -> None:
        if self.packed:
            return self._collect_packed(frames_at)
        T = self.T
        h = self.memory0[0]
        for t in range(T):
            self.enc.forward(frames_at(t), ("trunk",), out={"trunk": self.features[t]})
            logits, values, h = self._act(t, h)
            probs = torch.softmax(logits, -1)
            a = torch.multinomial(probs, 1, generator=self.gen)[:, 0]
            self.actions[t] = a
            self.log_probs[t] = torch.log_softmax(logits, -1).gather(-1, a[:, None])[:, 0]
            self.values[t, :, 0] = values
        # bootstrap value of the observation after the last step (synthetic: the last frames again)
        with torch.no_grad():
            _, v_next, _ = self.model.forward_tensors(self.features[T - 1:T], self.goals[T - 1:T], h, self.masks[T:T + 1])
        self.values[T, :, 0] = v_next[0]

    # ---------------------------------------------------------------- returns + update
    def update(self, global_rows: Optional[int] = None) -> Dict[str, torch.Tensor]:
        T = self.T
        returns, _, nadv = compute_returns_gae(self.rewards, self.values, self.masks, self.values[T], self.gamma, self.tau)
        if self.packed:
            from .actor_critic import PackedFeatures
            feats = PackedFeatures(self.features16.view(-1, self.features16.shape[-1]), T, self.N)
        else:
            feats = self.model.pack_features(self.features)
        rollout = dict(features=feats, goals=self.goals, masks=self.masks[:T], memory=self.memory0,
                       actions=self.actions, old_action_log_probs=self.log_probs, values=self.values[:T], returns=returns,
                       norm_adv_targ=nadv)
        return self.trainer.update(rollout, global_rows=global_rows)

    def step(self, frames_at: Callable[[int], torch.Tensor], global_rows: Optional[int] = None) -> Dict[str, torch.Tensor]:
        self.collect(frames_at)
        return self.update(global_rows)

    def launches_per_step(self) -> int:
        """Kernels of OURS launched by one step() (count of library launches; torch's sampler kernels excluded)."""
        enc = self.enc.launches_per_forward(("trunk",))
        act = 1 + 5 + 1 + 5 + 1 + 1                      # pack features, weight layouts, goal rows, 5 GEMMs, GRU, heads
        if self.packed:
            # per step: trunk without the NCHW head + row export; goal rows, 5 GEMMs, GRU, heads, sampler (weight layouts once)
            step = self.enc.launches_per_forward(()) + 1 + (1 + 5 + 1 + 1 + 1)
            fwd = 9 + 1 + 5 + 1 + 1
            bwd = 1 + 1 + 1 + 2 + 9 + 5 + 5 + 1
            upd = self.trainer.update_repeats * (fwd + 1 + bwd + 2)
            return self.T * step + 5 + (5 + 1 + 5 + 1 + 1) + 1 + upd
        fwd = 9 + 1 + 5 + 1 + 1                          # training forward: + transposed weight layouts
        bwd = 1 + 1 + 1 + 2 + 9 + 5 + 5 + 1              # heads bwd, BPTT, scale, casts, wgrads, dgrads, col-sums, goal grad
        upd = self.trainer.update_repeats * (fwd + 1 + bwd + 2)
        return self.T * (enc + act) + act + 2 + 1 + upd
