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

from .actor_critic import PPOTrainer, ResnetTensorNavActorCritic
from .encoder import ClipRN50Encoder


class SyntheticPPOStep:
    """A thin user of ``embclip_b200.storage.RolloutStorage``: the storage owns every rollout tensor; this class only
    produces synthetic frames / goals / masks / rewards in place of the simulator and drives encoder -> actor -> storage
    -> returns -> update in AllenAct's order.  The per-step producers (``encode_rows`` / ``act``) write straight into the
    storage's slots (``RolloutStorage.slots``), so no step tensor is copied twice."""

    def __init__(self, encoder: ClipRN50Encoder, model: ResnetTensorNavActorCritic, trainer: PPOTrainer, T: int = 128, N: int = 60,
                 seed: int = 0, gamma: float = 0.99, tau: float = 0.95, packed_rollout: bool = True):
        """packed_rollout=True (default): the rollout keeps CLIP features as the fp16 pixel rows the update reads
        (``encode_rows`` -> ``act`` -> ``PackedFeatures``), one library call per actor step.  False: the AllenAct data flow
        verbatim -- ``ClipResNetPreprocessor``-style fp32 [N,2048,7,7] features into RolloutStorage, ``forward`` + torch sampling,
        ``pack_features`` before the update.  Both produce bit-identical logits / values (tests/test_actor_critic_gpu.py)."""
        from .storage import RolloutStorage
        self.packed = bool(packed_rollout)
        self.enc, self.model, self.trainer, self.T, self.N = encoder, model, trainer, T, N
        self.gamma, self.tau = gamma, tau
        dev = model.flat_params.device
        self.device = dev
        g = torch.Generator(device=dev).manual_seed(seed)
        self.storage = RolloutStorage(T, N, model, packed_features=self.packed, seed=seed)
        st = self.storage
        st._feature_slot()
        # the simulator's side of the rollout, synthetic: goals, episode boundaries (~1 %), rewards
        st.observations[model.goal_uuid] = torch.zeros(T + 1, N, dtype=torch.int64, device=dev)
        st.observations[model.goal_uuid][:T] = torch.randint(0, model._plan.cfg["num_goals"], (T, N), device=dev, generator=g)
        st.observations[model.goal_uuid][T] = st.observations[model.goal_uuid][T - 1]
        st.masks.copy_((torch.rand(T + 1, N, 1, device=dev, generator=g) > 0.01).float())
        st.masks[0, : N // 4] = 0
        st.rewards.copy_(0.1 * torch.randn(T, N, 1, device=dev, generator=g))
        self.gen = g
        self.kernel_launches = 0

    # views under the names the rest of the repo (tests, bench) reads
    @property
    def features16(self):
        return self.storage.observations[self.model.resnet_uuid][:self.T]

    @property
    def features(self):
        return self.storage.observations[self.model.resnet_uuid][:self.T]

    @property
    def goals(self):
        return self.storage.observations[self.model.goal_uuid][:self.T]

    @property
    def masks(self):
        return self.storage.masks

    @property
    def values(self):
        return self.storage.value_preds

    @property
    def actions(self):
        return self.storage.actions[..., 0]

    @property
    def log_probs(self):
        return self.storage.action_log_probs[..., 0]

    # ---------------------------------------------------------------- rollout collection
    def collect(self, frames_at: Callable[[int], torch.Tensor]) -> None:
        """T rollout steps.  Step t: the preprocessor encodes the observation of step t into the storage's slot t, the actor
        reads slot t (features, goal, mask, memory) and its action / log-prob / value land in row t, the new memory in row t + 1
        -- RolloutStorage.insert's placement.  The observation after the last step is synthetic (the last frames again)."""
        T, N, st, mdl = self.T, self.N, self.storage, self.model
        feat = st.observations[mdl.resnet_uuid]
        goals = st.observations[mdl.goal_uuid]
        mem = st.memory.tensor("rnn")                          # [T + 1, 1, N, H]
        if self.packed:
            u = torch.rand(T, N, device=self.device, generator=self.gen)
            for t in range(T):
                self.enc.encode_rows(frames_at(t), out=feat[t])
                mdl.act(feat[t], goals[t], st.masks[t], mem[t, 0], u[t], actions=st.actions[t, :, 0],
                        action_log_probs=st.action_log_probs[t, :, 0], values=st.value_preds[t, :, 0], memory_out=mem[t + 1, 0])
        else:
            for t in range(T):
                self.enc.forward(frames_at(t), ("trunk",), out={"trunk": feat[t]})
                with torch.no_grad():
                    logits, values, h_new = mdl.forward_tensors(feat[t:t + 1], goals[t:t + 1], mem[t], st.masks[t:t + 1])
                probs = torch.softmax(logits[0], -1)
                a = torch.multinomial(probs, 1, generator=self.gen)[:, 0]
                st.actions[t, :, 0] = a
                st.action_log_probs[t, :, 0] = torch.log_softmax(logits[0], -1).gather(-1, a[:, None])[:, 0]
                st.value_preds[t, :, 0] = values[0]
                mem[t + 1, 0] = h_new
        st.prev_actions[1:].copy_(st.actions)
        feat[T].copy_(feat[T - 1])                              # synthetic: the observation after the last step = the last frames
        st.step = 0
        # bootstrap value of that observation (OnPolicyTrainer: actor_critic(rollouts.pick_observation_step(-1), ...).values)
        with torch.no_grad():
            out, _ = mdl(st.pick_observation_step(T), st.pick_memory_step(T), st.pick_prev_actions_step(T), st.masks[T:T + 1])
        self._next_value = out.values[0]

    # ---------------------------------------------------------------- returns + update
    def update(self, global_rows: Optional[int] = None) -> Dict[str, torch.Tensor]:
        st = self.storage
        st.compute_returns(self._next_value, True, self.gamma, self.tau)
        info = self.trainer.update_from_storage(st, global_rows=global_rows)
        st.after_update()
        return info

    def step(self, frames_at: Callable[[int], torch.Tensor], global_rows: Optional[int] = None) -> Dict[str, torch.Tensor]:
        self.collect(frames_at)
        return self.update(global_rows)

    def launches_per_step(self) -> int:
        """Kernels of OURS launched by one step() (count of library launches; torch's sampler kernels excluded)."""
        enc = self.enc.launches_per_forward(("trunk",))
        act = 1 + 5 + 1 + 5 + 1 + 1                      # pack features, weight layouts, goal rows, 5 GEMMs, GRU, heads
        reps = self.trainer.update_repeats * self.trainer.num_mini_batch
        if self.packed:
            # per step: trunk without the NCHW head + row export; goal rows, 5 GEMMs, GRU, heads, sampler (weight layouts once)
            step = self.enc.launches_per_forward(()) + 1 + (1 + 5 + 1 + 1 + 1)
            fwd = 9 + 1 + 5 + 1 + 1
            bwd = 1 + 1 + 1 + 2 + 9 + 5 + 5 + 1
            upd = reps * (fwd + 1 + bwd + 2)
            return self.T * step + 5 + (5 + 1 + 5 + 1 + 1) + 1 + upd
        fwd = 9 + 1 + 5 + 1 + 1                          # training forward: + transposed weight layouts
        bwd = 1 + 1 + 1 + 2 + 9 + 5 + 5 + 1              # heads bwd, BPTT, scale, casts, wgrads, dgrads, col-sums, goal grad
        upd = reps * (fwd + 1 + bwd + 2)
        return self.T * (enc + act) + act + 2 + 1 + upd
