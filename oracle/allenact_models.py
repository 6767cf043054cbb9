"""ORACLE (test infrastructure, not product code): plain-PyTorch fp32 restatement of the
AllenAct pieces on EmbCLIP's PPO hot path.

PARITY UNPINNED against the pinned third-party source ``allenai/allenact @ v0.5.0`` (pin:
/root/reference/readme_files/baselines_robothor_objectnav.md:6; launch lines :48-51), which is
not vendored in /root/reference, not installable here (``gym`` / ``allenact`` absent, no
network) and ships no golden vectors for this path.  What IS pinned by
tests/test_oracle_allenact.py: ``torch.nn.GRU`` -- the very module RNNStateEncoder wraps -- is
the recurrence used here, the masked seq_forward is cross-checked against a step-by-step
single_forward loop, and GAE is cross-checked against a float64 closed form.

Restated (module :: class):
  A8  allenact/embodiedai/models/basic_models.py :: RNNStateEncoder (GRU, 1 layer)
  A9  allenact/algorithms/onpolicy_sync/policy.py :: LinearActorHead, LinearCriticHead
  A10 projects/objectnav_baselines/models/object_nav_models.py :: ResnetTensorGoalEncoder,
      ResnetTensorNavActorCritic (what objectnav_robothor_rgb_clipresnet50gru_ddppo builds)
  A11 allenact/algorithms/onpolicy_sync/losses/ppo.py :: PPO.loss_per_step / PPOConfig
  A12 allenact/algorithms/onpolicy_sync/storage.py :: RolloutStorage.compute_returns (GAE),
      advantage normalisation of recurrent_generator
  A13 allenact/algorithms/onpolicy_sync/engine.py :: OnPolicyTrainer.update / backprop_step
  A14 allenact/base_abstractions/distributions.py :: CategoricalDistr
Parameter names are upstream's, so an AllenAct checkpoint's ``model_state_dict`` loads.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn


class CategoricalDistr(torch.distributions.Categorical):
    """A14: log_prob / entropy keep a trailing-dim-free [T,N] shape; mode() = argmax."""

    def mode(self) -> torch.Tensor:
        return self._param.argmax(dim=-1, keepdim=False)

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        if value.shape == self.logits.shape[:-1]:
            return super().log_prob(value)
        if value.shape == self.logits.shape[:-1] + (1,):
            return super().log_prob(value.squeeze(-1)).unsqueeze(-1)
        raise ValueError(f"bad action shape {tuple(value.shape)} for logits {tuple(self.logits.shape)}")


class LinearActorHead(nn.Module):
    def __init__(self, num_inputs: int, num_outputs: int):
        super().__init__()
        self.linear = nn.Linear(num_inputs, num_outputs)
        nn.init.orthogonal_(self.linear.weight, gain=0.01)
        nn.init.constant_(self.linear.bias, 0)

    def forward(self, x: torch.Tensor) -> CategoricalDistr:
        return CategoricalDistr(logits=self.linear(x))


class LinearCriticHead(nn.Module):
    def __init__(self, input_size: int):
        super().__init__()
        self.fc = nn.Linear(input_size, 1)
        nn.init.orthogonal_(self.fc.weight)
        nn.init.constant_(self.fc.bias, 0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fc(x).view(*x.shape[:2], -1)


class RNNStateEncoder(nn.Module):
    """A8.  forward(x [T,N,I], hidden [1,N,H], masks [T,N,1]) -> (out [T,N,H], hidden [1,N,H]).
    masks[t] == 0 marks an episode start: the hidden state entering step t is zeroed (or
    replaced by the learned init state)."""

    def __init__(self, input_size: int, hidden_size: int, num_layers: int = 1,
                 trainable_masked_hidden_state: bool = False):
        super().__init__()
        self.rnn = nn.GRU(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers)
        self.trainable_masked_hidden_state = trainable_masked_hidden_state
        if trainable_masked_hidden_state:
            self.init_hidden_state = nn.Parameter(0.1 * torch.randn((num_layers, 1, hidden_size)))
        for name, p in self.rnn.named_parameters():
            if "weight" in name:
                nn.init.orthogonal_(p)
            elif "bias" in name:
                nn.init.constant_(p, 0)

    def _mask_hidden(self, h: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        if not self.trainable_masked_hidden_state:
            return masks * h
        return masks * h + (1.0 - masks) * self.init_hidden_state.repeat(1, h.shape[1], 1)

    def single_forward(self, x, h, masks):
        out, h = self.rnn(x, self._mask_hidden(h, masks[0].view(1, -1, 1)))
        return out, h

    def seq_forward(self, x, h, masks):
        """Upstream splits T at the steps where any sampler restarts (host sync via
        ``.nonzero().cpu()``) and runs cuDNN GRU per segment."""
        T = x.shape[0]
        cut = (masks[1:] == 0.0).any(dim=-1).any(dim=-1).nonzero().flatten().cpu().tolist()
        bounds = [0] + [c + 1 for c in cut] + [T]
        outs = []
        for s, e in zip(bounds[:-1], bounds[1:]):
            o, h = self.rnn(x[s:e], self._mask_hidden(h, masks[s].view(1, -1, 1)))
            outs.append(o)
        return torch.cat(outs, dim=0), h

    def forward(self, x, h, masks):
        if x.shape[0] == 1:
            return self.single_forward(x, h, masks)
        return self.seq_forward(x, h, masks)


class ResnetTensorGoalEncoder(nn.Module):
    """A10 front half: 1x1-conv compressor of the [2048,7,7] CLIP feature, goal embedding
    broadcast over 7x7, 1x1-conv combiner, flatten in (C,H,W) order -> 32*49 = 1568."""

    def __init__(self, resnet_tensor_shape=(2048, 7, 7), num_goals: int = 12, class_dims: int = 32,
                 resnet_compressor_hidden_out_dims=(128, 32), combiner_hidden_out_dims=(128, 32)):
        super().__init__()
        self.resnet_tensor_shape = tuple(resnet_tensor_shape)
        self.class_dims = class_dims
        r, c = resnet_compressor_hidden_out_dims, combiner_hidden_out_dims
        self.embed_class = nn.Embedding(num_goals, class_dims)
        self.resnet_compressor = nn.Sequential(
            nn.Conv2d(resnet_tensor_shape[0], r[0], 1), nn.ReLU(),
            nn.Conv2d(r[0], r[1], 1), nn.ReLU())
        self.target_obs_combiner = nn.Sequential(
            nn.Conv2d(r[1] + class_dims, c[0], 1), nn.ReLU(),
            nn.Conv2d(c[0], c[1], 1))
        self.output_dims = c[-1] * resnet_tensor_shape[1] * resnet_tensor_shape[2]

    def forward(self, feats: torch.Tensor, goals: torch.Tensor) -> torch.Tensor:
        """feats [T,N,C,H,W] fp32, goals [T,N] int64 -> [T,N,1568]."""
        T, N = feats.shape[:2]
        f = feats.reshape(T * N, *feats.shape[2:])
        g = self.embed_class(goals.reshape(T * N))
        g = g.view(-1, self.class_dims, 1, 1).expand(-1, -1, f.shape[-2], f.shape[-1])
        x = self.target_obs_combiner(torch.cat([self.resnet_compressor(f), g], dim=1))
        return x.reshape(T, N, -1)


class ResnetTensorNavActorCritic(nn.Module):
    """A10.  forward(observations, memory, prev_actions, masks) with memory = the "rnn" tensor
    [1,N,512] (the AllenAct `Memory` wrapper lives in the plugin mirror, not the oracle)."""

    def __init__(self, num_actions: int = 6, num_goals: int = 12, hidden_size: int = 512,
                 goal_dims: int = 32, resnet_tensor_shape=(2048, 7, 7),
                 rgb_uuid: str = "rgb_clip_resnet", goal_uuid: str = "goal_object_type_ind",
                 trainable_masked_hidden_state: bool = False):
        super().__init__()
        self.rgb_uuid, self.goal_uuid = rgb_uuid, goal_uuid
        self.goal_visual_encoder = ResnetTensorGoalEncoder(resnet_tensor_shape, num_goals, goal_dims)
        self.state_encoder = RNNStateEncoder(self.goal_visual_encoder.output_dims, hidden_size,
                                             trainable_masked_hidden_state=trainable_masked_hidden_state)
        self.actor = LinearActorHead(hidden_size, num_actions)
        self.critic = LinearCriticHead(hidden_size)

    def forward(self, observations: Dict[str, torch.Tensor], memory: torch.Tensor,
                prev_actions: Optional[torch.Tensor], masks: torch.Tensor
                ) -> Tuple[CategoricalDistr, torch.Tensor, torch.Tensor]:
        x = self.goal_visual_encoder(observations[self.rgb_uuid], observations[self.goal_uuid])
        x, h = self.state_encoder(x, memory, masks)
        return self.actor(x), self.critic(x), h


# ---------------------------------------------------------------------------------------------
# A11: PPO loss.  PPOConfig = clip 0.1, value coef 0.5, entropy coef 0.01, clipped value loss.
# ---------------------------------------------------------------------------------------------
PPOConfig = dict(clip_param=0.1, value_loss_coef=0.5, entropy_coef=0.01)


def ppo_loss(distr: CategoricalDistr, values: torch.Tensor, batch: Dict[str, torch.Tensor],
             clip_param: float = 0.1, value_loss_coef: float = 0.5, entropy_coef: float = 0.01,
             use_clipped_value_loss: bool = True) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """batch: actions [T,N] int64, old_action_log_probs [T,N], values / returns /
    norm_adv_targ [T,N,1].  Returns (total_loss, {value, action, entropy} means)."""
    logp = distr.log_prob(batch["actions"]).unsqueeze(-1)
    entropy = distr.entropy().unsqueeze(-1)
    ratio = torch.exp(logp - batch["old_action_log_probs"].unsqueeze(-1))
    clamped = torch.clamp(ratio, 1.0 - clip_param, 1.0 + clip_param)
    adv = batch["norm_adv_targ"]
    surr1, surr2 = ratio * adv, clamped * adv
    action_loss = -torch.where(surr2 < surr1, surr2, surr1)
    if use_clipped_value_loss:
        v_clip = batch["values"] + (values - batch["values"]).clamp(-clip_param, clip_param)
        value_loss = 0.5 * torch.max((values - batch["returns"]).pow(2), (v_clip - batch["returns"]).pow(2))
    else:
        value_loss = 0.5 * (batch["returns"] - values).pow(2)
    parts = {"value": value_loss.mean(), "action": action_loss.mean(), "entropy": (-entropy).mean()}
    total = parts["action"] + value_loss_coef * parts["value"] + entropy_coef * parts["entropy"]
    return total, parts


# ---------------------------------------------------------------------------------------------
# A12: GAE returns + advantage normalisation.
# ---------------------------------------------------------------------------------------------
def compute_returns_gae(rewards: torch.Tensor, value_preds: torch.Tensor, masks: torch.Tensor,
                        next_value: torch.Tensor, gamma: float = 0.99, tau: float = 0.95) -> torch.Tensor:
    """rewards [T,N,1]; value_preds [T+1,N,1] (last row overwritten by next_value);
    masks [T+1,N,1].  Returns `returns` [T+1,N,1] (row T left zero, as upstream)."""
    T = rewards.shape[0]
    vp = value_preds.clone()
    vp[-1] = next_value
    returns = torch.zeros_like(vp)
    gae = torch.zeros_like(vp[0])
    for t in reversed(range(T)):
        delta = rewards[t] + gamma * vp[t + 1] * masks[t + 1] - vp[t]
        gae = delta + gamma * tau * masks[t + 1] * gae
        returns[t] = gae + vp[t]
    return returns


def normalized_advantages(returns: torch.Tensor, value_preds: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    adv = returns[:-1] - value_preds[:-1]
    return (adv - adv.mean()) / (adv.std() + eps)


# ---------------------------------------------------------------------------------------------
# A13: one PPO update (update_repeats x 1 mini-batch), single process.
# ---------------------------------------------------------------------------------------------
def ppo_update(model: ResnetTensorNavActorCritic, optimizer: torch.optim.Optimizer,
               rollout: Dict[str, torch.Tensor], update_repeats: int = 4, max_grad_norm: float = 0.5,
               grad_hook=None) -> Dict[str, float]:
    """rollout: features [T,N,2048,7,7], goals [T,N], masks [T,N,1], actions [T,N],
    old_action_log_probs [T,N], values [T,N,1], returns [T,N,1], norm_adv_targ [T,N,1],
    memory [1,N,512].  `grad_hook(params)` is where the engine all-reduces gradients."""
    info = {}
    for _ in range(update_repeats):
        distr, values, _ = model({model.rgb_uuid: rollout["features"], model.goal_uuid: rollout["goals"]},
                                 rollout["memory"], None, rollout["masks"])
        total, parts = ppo_loss(distr, values, rollout, **PPOConfig)
        optimizer.zero_grad()
        total.backward()
        if grad_hook is not None:
            grad_hook([p for p in model.parameters() if p.grad is not None])
        nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)
        optimizer.step()
        info = {k: float(v) for k, v in parts.items()}
        info["total"] = float(total)
    return info
