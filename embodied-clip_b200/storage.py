"""On-device ``RolloutStorage`` for the PPO path: a drop-in for
allenact/algorithms/onpolicy_sync/storage.py::RolloutStorage [UPSTREAM, allenact v0.5.0 -- pin at
/root/reference/readme_files/baselines_robothor_objectnav.md:6; the launch line that builds it is :48-51] restricted to what
the ObjectNav experiment config uses: one discrete action per step, tensor observations, one recurrent memory ("rnn").

Same surface and semantics: ``insert`` (observations / memory land at ``step + 1``, everything else at ``step``),
``compute_returns`` (GAE reverse scan, one kernel: ``embclip_gae``), ``recurrent_generator`` (mini-batches = contiguous
chunks of samplers with bounds ``round(linspace(0, N, num_mini_batch + 1))``, chunk order shuffled), ``after_update`` (last
step becomes step 0), ``pick_observation_step`` / ``pick_memory_step``, ``to``.  Tensors are ``[steps (+1), samplers, ...]``.

What differs, by design: the CLIP feature observation (``[N, 2048, 7, 7]`` fp32 per step upstream = 3.08 GB per 128 x 60
rollout) can be kept as the fp16 pixel rows the update kernels read (``packed_features=True``: ``[T + 1, N * 49, 2048]``,
half the bytes, no re-pack per update); ``insert`` then accepts either the fp32 tensor (packed by ``embclip_ac_pack_features``)
or the rows ``ClipRN50Encoder.encode_rows`` produced, and the generator yields ``PackedFeatures`` under the same observation
key, which ``ResnetTensorNavActorCritic.forward`` consumes directly.  Everything stays on the device; nothing syncs the host.
"""
from __future__ import annotations

import random
from typing import Any, Dict, Iterator, Optional, Sequence

import torch

from .actor_critic import Memory, PackedFeatures, ResnetTensorNavActorCritic, compute_returns_gae


class RolloutStorage:
    def __init__(self, num_steps: int, num_samplers: int, actor_critic: ResnetTensorNavActorCritic, packed_features: bool = True,
                 seed: Optional[int] = None, *args, **kwargs):
        self.num_steps, self.num_samplers = int(num_steps), int(num_samplers)
        self.actor_critic = actor_critic
        self.feature_uuid = actor_critic.resnet_uuid
        self.packed_features = bool(packed_features)
        self.device = actor_critic.flat_params.device
        T, N = self.num_steps, self.num_samplers
        spec = actor_critic._recurrent_memory_specification()
        self.memory = Memory()
        for key, (dims, dtype) in spec.items():
            shape = [T + 1] + [N if size is None else int(size) for _, size in dims]
            sampler_dim = 1 + [name for name, _ in dims].index("sampler")
            self.memory[key] = (torch.zeros(*shape, dtype=dtype, device=self.device), sampler_dim)
        self.observations: Dict[str, torch.Tensor] = {}
        self.value_preds = torch.zeros(T + 1, N, 1, device=self.device)
        self.returns = torch.zeros(T + 1, N, 1, device=self.device)
        self.rewards = torch.zeros(T, N, 1, device=self.device)
        self.action_log_probs = torch.zeros(T, N, 1, device=self.device)
        self.actions = torch.zeros(T, N, 1, dtype=torch.int64, device=self.device)
        self.prev_actions = torch.zeros(T + 1, N, 1, dtype=torch.int64, device=self.device)
        self.masks = torch.ones(T + 1, N, 1, device=self.device)
        self.advantages: Optional[torch.Tensor] = None           # filled by compute_returns (same kernel)
        self.norm_advantages: Optional[torch.Tensor] = None
        self.step = 0
        self._rng = random.Random(seed) if seed is not None else random   # upstream shuffles with the global `random`

    # ------------------------------------------------------------------ placement
    def to(self, device) -> "RolloutStorage":
        device = torch.device(device)
        if device != self.device:
            if device.type != "cuda":
                raise RuntimeError("embclip_b200 RolloutStorage lives on the trainer's CUDA device (no CPU path)")
            for k in ("value_preds", "returns", "rewards", "action_log_probs", "actions", "prev_actions", "masks"):
                setattr(self, k, getattr(self, k).to(device))
            self.observations = {k: v.to(device) for k, v in self.observations.items()}
            for k in list(self.memory):
                self.memory[k] = (self.memory[k][0].to(device), self.memory[k][1])
            self.device = device
        return self

    # ------------------------------------------------------------------ insertion
    def _feature_slot(self) -> torch.Tensor:
        C_, Hh, Ww = self.actor_critic.resnet_tensor_shape
        T, N = self.num_steps, self.num_samplers
        if self.feature_uuid not in self.observations:
            if self.packed_features:
                self.observations[self.feature_uuid] = torch.zeros(T + 1, N * Hh * Ww, C_, dtype=torch.float16, device=self.device)
            else:
                self.observations[self.feature_uuid] = torch.zeros(T + 1, N, C_, Hh, Ww, dtype=torch.float32, device=self.device)
        return self.observations[self.feature_uuid]

    def insert_observations(self, processed_observations: Dict[str, torch.Tensor], time_step: int) -> None:
        N = self.num_samplers
        for uuid, value in processed_observations.items():
            if uuid == self.feature_uuid:
                slot = self._feature_slot()
                C_, Hh, Ww = self.actor_critic.resnet_tensor_shape
                if not self.packed_features:
                    slot[time_step].copy_(value.reshape(N, C_, Hh, Ww))
                elif isinstance(value, PackedFeatures):
                    slot[time_step].copy_(value.data.reshape(N * Hh * Ww, C_))
                elif value.dtype == torch.float16:                        # rows from ClipRN50Encoder.encode_rows
                    slot[time_step].copy_(value.reshape(N * Hh * Ww, C_))
                else:                                                    # fp32 [N, 2048, 7, 7] from the preprocessor
                    slot[time_step].copy_(self.actor_critic.pack_features(value.reshape(1, N, C_, Hh, Ww)).data)
                continue
            value = torch.as_tensor(value, device=self.device)
            if uuid not in self.observations:
                self.observations[uuid] = torch.zeros(self.num_steps + 1, *value.shape, dtype=value.dtype, device=self.device)
            self.observations[uuid][time_step].copy_(value)

    def insert_memory(self, memory: Optional[Memory], time_step: int) -> None:
        if memory is None:
            return
        for key in self.memory:
            src = memory.tensor(key) if hasattr(memory, "tensor") else memory[key]
            self.memory[key][0][time_step].copy_(src.reshape(self.memory[key][0][time_step].shape))

    def insert(self, observations: Dict[str, torch.Tensor], memory: Optional[Memory], actions: torch.Tensor,
               action_log_probs: torch.Tensor, value_preds: torch.Tensor, rewards: torch.Tensor, masks: torch.Tensor) -> None:
        """All arguments without the step dimension ([samplers, ...]), as OnPolicyRLEngine.collect_rollout_step passes them."""
        N = self.num_samplers
        self.insert_observations(observations, time_step=self.step + 1)
        self.insert_memory(memory, time_step=self.step + 1)
        self.actions[self.step].copy_(actions.reshape(N, 1))
        self.prev_actions[self.step + 1].copy_(actions.reshape(N, 1))
        self.masks[self.step + 1].copy_(masks.reshape(N, 1))
        self.action_log_probs[self.step].copy_(action_log_probs.reshape(N, 1))
        self.value_preds[self.step].copy_(value_preds.reshape(N, 1))
        self.rewards[self.step].copy_(rewards.reshape(N, 1))
        self.step = (self.step + 1) % self.num_steps

    # ------------------------------------------------------------------ reads used by the rollout loop
    def pick_observation_step(self, step: int) -> Dict[str, Any]:
        """Observations of one step with a leading steps dimension of 1 (what actor_critic.forward takes)."""
        out: Dict[str, Any] = {}
        for uuid, t in self.observations.items():
            if uuid == self.feature_uuid and self.packed_features:
                out[uuid] = PackedFeatures(t[step], 1, self.num_samplers)
            else:
                out[uuid] = t[step:step + 1]
        return out

    def pick_memory_step(self, step: int) -> Memory:
        m = Memory()
        for key, (t, sampler_dim) in self.memory.items():
            m[key] = (t[step], sampler_dim - 1)
        return m

    def pick_prev_actions_step(self, step: int) -> torch.Tensor:
        return self.prev_actions[step:step + 1]

    # ------------------------------------------------------------------ returns
    def compute_returns(self, next_value: torch.Tensor, use_gae: bool = True, gamma: float = 0.99, tau: float = 0.95) -> None:
        """returns[t] = gae_t + value_preds[t], gae_t = delta_t + gamma tau masks[t+1] gae_{t+1},
        delta_t = rewards[t] + gamma value_preds[t+1] masks[t+1] - value_preds[t]; value_preds[-1] = next_value.
        use_gae=False is the discounted-return recursion returns[t] = returns[t+1] gamma masks[t+1] + rewards[t], which is the
        same scan with tau = 1.  Also leaves ``advantages`` = returns[:-1] - value_preds[:-1] and its normalised form
        (mean / (std + 1e-5) over this rank's [T, N] block, as OnPolicyTrainer.update takes them) from the same kernel."""
        T = self.num_steps
        self.value_preds[T].copy_(next_value.reshape(self.num_samplers, 1))
        ret, adv, nadv = compute_returns_gae(self.rewards, self.value_preds, self.masks, self.value_preds[T], gamma,
                                             tau if use_gae else 1.0)
        self.returns[:T].copy_(ret)
        if not use_gae:
            self.returns[T].copy_(self.value_preds[T])
        self.advantages, self.norm_advantages = adv, nadv

    # ------------------------------------------------------------------ mini-batches
    def minibatch_bounds(self, num_mini_batch: int):
        N = self.num_samplers
        if N < num_mini_batch:
            raise AssertionError(f"number of samplers ({N}) must be at least the number of mini-batches ({num_mini_batch})")
        inds = [int(round(i * N / num_mini_batch)) for i in range(num_mini_batch + 1)]
        return list(zip(inds[:-1], inds[1:]))

    def recurrent_generator(self, advantages: Optional[torch.Tensor] = None, adv_mean: Optional[torch.Tensor] = None,
                            adv_std: Optional[torch.Tensor] = None, num_mini_batch: int = 1) -> Iterator[Dict[str, Any]]:
        """Yields one dict per mini-batch with the upstream keys.  advantages / adv_mean / adv_std as upstream; all None uses
        the advantages compute_returns left behind."""
        T, N = self.num_steps, self.num_samplers
        if advantages is None:
            if self.advantages is None:
                raise RuntimeError("recurrent_generator: call compute_returns first (or pass advantages)")
            advantages, normalized = self.advantages, self.norm_advantages
        else:
            normalized = (advantages - adv_mean) / (adv_std + 1e-5)
        pairs = self.minibatch_bounds(num_mini_batch)
        self._rng.shuffle(pairs)
        C_, Hh, Ww = self.actor_critic.resnet_tensor_shape
        for a, b in pairs:
            whole = (a, b) == (0, N)
            cut = (lambda t: t[:T]) if whole else (lambda t: t[:T, a:b].contiguous())
            obs: Dict[str, Any] = {}
            for uuid, t in self.observations.items():
                if uuid == self.feature_uuid and self.packed_features:
                    rows = t[:T] if whole else t[:T].view(T, N, Hh * Ww * C_)[:, a:b].contiguous()
                    obs[uuid] = PackedFeatures(rows.reshape(-1, C_), T, b - a)
                else:
                    obs[uuid] = cut(t)
            mem = Memory()
            for key, (t, sampler_dim) in self.memory.items():
                first = t[0]
                mem[key] = (first if whole else first.narrow(sampler_dim - 1, a, b - a).contiguous(), sampler_dim - 1)
            yield {"observations": obs, "memory": mem, "actions": cut(self.actions), "prev_actions": cut(self.prev_actions),
                   "values": cut(self.value_preds), "returns": cut(self.returns), "masks": cut(self.masks),
                   "old_action_log_probs": cut(self.action_log_probs), "adv_targ": cut(advantages),
                   "norm_adv_targ": cut(normalized), "samplers": (a, b)}

    # ------------------------------------------------------------------ roll over
    def after_update(self) -> None:
        for t in self.observations.values():
            t[0].copy_(t[-1])
        for key, (t, _) in self.memory.items():
            t[0].copy_(t[-1])
        self.masks[0].copy_(self.masks[-1])
        self.prev_actions[0].copy_(self.prev_actions[-1])
