"""ORACLE (test infrastructure, not product code): plain-PyTorch CPU restatement of
allenact/algorithms/onpolicy_sync/storage.py::RolloutStorage and of the update loop of
allenact/algorithms/onpolicy_sync/engine.py::OnPolicyTrainer.update [UPSTREAM allenact v0.5.0, pin at
/root/reference/readme_files/baselines_robothor_objectnav.md:6], restricted to what the ObjectNav experiment config uses
(one discrete action, tensor observations, one "rnn" memory).  PARITY UNPINNED: allenact cannot be installed here and the
reference holds no fixture for it; written from the upstream semantics listed in SURVEY.md 8a A12-A13.

Tensor layout as upstream: observations / memory / masks / prev_actions / value_preds / returns have num_steps + 1 rows
(row 0 = state before the first step of this rollout), rewards / actions / action_log_probs have num_steps rows; everything
is [step, sampler, ...]; memory "rnn" is [step, layer, sampler, hidden].  Only tests/ may import this module."""
from __future__ import annotations

import random
from typing import Dict, Iterator, Optional

import numpy as np
import torch


class RefRolloutStorage:
    def __init__(self, num_steps: int, num_samplers: int, hidden: int = 512, seed: Optional[int] = None):
        T, N = num_steps, num_samplers
        self.num_steps, self.num_samplers = T, N
        self.observations: Dict[str, torch.Tensor] = {}
        self.memory = {"rnn": torch.zeros(T + 1, 1, N, hidden)}
        self.value_preds = torch.zeros(T + 1, N, 1)
        self.returns = torch.zeros(T + 1, N, 1)
        self.rewards = torch.zeros(T, N, 1)
        self.action_log_probs = torch.zeros(T, N, 1)
        self.actions = torch.zeros(T, N, 1, dtype=torch.int64)
        self.prev_actions = torch.zeros(T + 1, N, 1, dtype=torch.int64)
        self.masks = torch.ones(T + 1, N, 1)
        self.step = 0
        self._rng = random.Random(seed) if seed is not None else random

    def insert_observations(self, obs: Dict[str, torch.Tensor], time_step: int) -> None:
        for k, v in obs.items():
            if k not in self.observations:
                self.observations[k] = torch.zeros(self.num_steps + 1, *v.shape, dtype=v.dtype)
            self.observations[k][time_step].copy_(v)

    def insert(self, observations, memory, actions, action_log_probs, value_preds, rewards, masks) -> None:
        self.insert_observations(observations, self.step + 1)
        self.memory["rnn"][self.step + 1].copy_(memory)
        self.actions[self.step].copy_(actions)
        self.prev_actions[self.step + 1].copy_(actions)
        self.masks[self.step + 1].copy_(masks)
        self.action_log_probs[self.step].copy_(action_log_probs)
        self.value_preds[self.step].copy_(value_preds)
        self.rewards[self.step].copy_(rewards)
        self.step = (self.step + 1) % self.num_steps

    def compute_returns(self, next_value: torch.Tensor, use_gae: bool, gamma: float, tau: float) -> None:
        if use_gae:
            self.value_preds[-1] = next_value
            gae = 0
            for step in reversed(range(self.rewards.shape[0])):
                delta = self.rewards[step] + gamma * self.value_preds[step + 1] * self.masks[step + 1] - self.value_preds[step]
                gae = delta + gamma * tau * self.masks[step + 1] * gae
                self.returns[step] = gae + self.value_preds[step]
        else:
            self.returns[-1] = next_value
            for step in reversed(range(self.rewards.shape[0])):
                self.returns[step] = self.returns[step + 1] * gamma * self.masks[step + 1] + self.rewards[step]

    def recurrent_generator(self, advantages, adv_mean, adv_std, num_mini_batch: int) -> Iterator[Dict[str, object]]:
        normalized = (advantages - adv_mean) / (adv_std + 1e-5)
        N = self.rewards.shape[1]
        assert N >= num_mini_batch
        inds = np.round(np.linspace(0, N, num_mini_batch + 1, endpoint=True)).astype(np.int32)
        pairs = list(zip(inds[:-1], inds[1:]))
        self._rng.shuffle(pairs)
        for a, b in pairs:
            cur = list(range(a, b))
            yield {"observations": {k: v[:-1][:, cur] for k, v in self.observations.items()},
                   "memory": {"rnn": self.memory["rnn"][0][:, cur]},
                   "actions": self.actions[:, cur], "prev_actions": self.prev_actions[:-1][:, cur],
                   "values": self.value_preds[:-1][:, cur], "returns": self.returns[:-1][:, cur], "masks": self.masks[:-1][:, cur],
                   "old_action_log_probs": self.action_log_probs[:, cur], "adv_targ": advantages[:, cur],
                   "norm_adv_targ": normalized[:, cur], "samplers": (int(a), int(b))}

    def after_update(self) -> None:
        for v in self.observations.values():
            v[0].copy_(v[-1])
        self.memory["rnn"][0].copy_(self.memory["rnn"][-1])
        self.masks[0].copy_(self.masks[-1])
        self.prev_actions[0].copy_(self.prev_actions[-1])


def ref_update_from_storage(model, optimizer, storage: RefRolloutStorage, update_repeats: int, num_mini_batch: int,
                            max_grad_norm: float = 0.5, lr_lambda=None, base_lr: float = 3e-4, total_steps: int = 0):
    """engine.py update(): advantages over the whole block, then repeats x mini-batches x (forward, PPO loss, backward,
    clip_grad_norm_, optimizer.step); LambdaLR(LinearDecay) evaluated at the steps collected BEFORE this update."""
    from .allenact_models import PPOConfig, ppo_loss
    if lr_lambda is not None:
        for g in optimizer.param_groups:
            g["lr"] = base_lr * lr_lambda(total_steps)
    advantages = storage.returns[:-1] - storage.value_preds[:-1]
    adv_mean, adv_std = advantages.mean(), advantages.std()
    info = {}
    for _ in range(update_repeats):
        for batch in storage.recurrent_generator(advantages, adv_mean, adv_std, num_mini_batch):
            obs = batch["observations"]
            distr, values, _ = model({model.rgb_uuid: obs[model.rgb_uuid], model.goal_uuid: obs[model.goal_uuid]},
                                     batch["memory"]["rnn"], batch["prev_actions"], batch["masks"])
            b = dict(actions=batch["actions"][..., 0], old_action_log_probs=batch["old_action_log_probs"][..., 0],
                     values=batch["values"], returns=batch["returns"], norm_adv_targ=batch["norm_adv_targ"])
            total, parts = ppo_loss(distr, values, b, **PPOConfig)
            optimizer.zero_grad()
            total.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)
            optimizer.step()
            info = {k: float(v) for k, v in parts.items()}
            info["total"] = float(total)
    return info
