"""AllenAct actor-critic surface for the B200 PPO-update path.

Mirrors of allenai/allenact v0.5.0 (pin: /root/reference/readme_files/baselines_robothor_objectnav.md:6; the
experiment config that instantiates them -- objectnav_robothor_rgb_clipresnet50gru_ddppo -- is named at :51):

* ``ResnetTensorNavActorCritic``  projects/objectnav_baselines/models/object_nav_models.py -- an ``nn.Module`` whose
  ``forward(observations, memory, prev_actions, masks) -> (ActorCriticOutput, Memory)`` keeps the upstream
  signature, tensor shapes ([steps, samplers, ...]) and ``state_dict`` key names, and is autograd-compatible
  (``total_loss.backward()`` fills ``p.grad`` of ``actor_critic.parameters()``), but whose arithmetic runs in
  libembclip_b200.so (``embclip_ac_forward`` / ``embclip_ac_backward``).
* ``CategoricalDistr``            allenact/base_abstractions/distributions.py (log_prob / entropy / mode = argmax)
* ``PPO`` / ``PPOConfig``         allenact/algorithms/onpolicy_sync/losses/ppo.py
* ``compute_returns_gae``         allenact/algorithms/onpolicy_sync/storage.py RolloutStorage.compute_returns
* ``PPOTrainer``                  the update()/backprop_step() half of allenact/algorithms/onpolicy_sync/engine.py:
  update_repeats x (forward, PPO loss, backward, gradient all-reduce, clip_grad_norm_, Adam), all on flat buffers.

No CPU path: every class raises unless its tensors are on a CUDA device and the library is built.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, NamedTuple, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib


# ---------------------------------------------------------------------------------------------------
# small upstream value types
# ---------------------------------------------------------------------------------------------------
class CategoricalDistr(torch.distributions.Categorical):
    """allenact CategoricalDistr: ``mode()`` is the argmax action; log_prob accepts [..] or [.., 1] actions."""

    def mode(self) -> torch.Tensor:
        return self._param.argmax(dim=-1, keepdim=False)

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        if value.shape == self.logits.shape[:-1]:
            return super().log_prob(value)
        if value.shape == self.logits.shape[:-1] + (1,):
            return super().log_prob(value.squeeze(-1)).unsqueeze(-1)
        raise ValueError(f"bad action shape {tuple(value.shape)} for logits {tuple(self.logits.shape)}")


class ActorCriticOutput(NamedTuple):
    distributions: CategoricalDistr
    values: torch.Tensor                # [steps, samplers, 1]
    extras: Dict[str, Any]


class Memory(dict):
    """Minimal stand-in for allenact.base_abstractions.misc.Memory: key -> (tensor, sampler_dim)."""

    def tensor(self, key: str) -> torch.Tensor:
        return self[key][0]

    def set_tensor(self, key: str, tensor: torch.Tensor) -> "Memory":
        self[key] = (tensor, self[key][1] if key in self else 1)
        return self


PPOConfig = dict(clip_param=0.1, value_loss_coef=0.5, entropy_coef=0.01)


# ---------------------------------------------------------------------------------------------------
# library handle + flat parameter storage
# ---------------------------------------------------------------------------------------------------
class _ACPlan:
    def __init__(self, cfg: Dict[str, int]):
        self.lib = _lib.load()
        c = _lib.ACCfg(**cfg)
        self.cfg = dict(cfg)
        self._h = C.c_void_p()
        _lib.check(self.lib.embclip_ac_create(C.byref(c), C.byref(self._h)))
        self.n_floats = int(self.lib.embclip_ac_param_floats(self._h))
        self.params = []                 # (name, shape, float offset, numel)
        for i in range(_lib.check(self.lib.embclip_ac_num_params(self._h))):
            pi = _lib.ParamInfo()
            _lib.check(self.lib.embclip_ac_param_info(self._h, i, C.byref(pi)))
            shape = tuple(pi.shape[:pi.ndim])
            self.params.append((pi.name.decode(), shape, int(pi.offset) // 4, int(pi.nbytes) // 4))

    def workspace_bytes(self, T: int, N: int) -> int:
        return int(self.lib.embclip_ac_workspace_bytes(self._h, T, N))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.embclip_ac_destroy(self._h)
                self._h = None
        except Exception:
            pass


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


class PackedFeatures:
    """fp16 [T*N*49, 2048] rows of a rollout's CLIP features (``embclip_ac_pack_features``): built once per rollout and
    shared by all update passes.  ``ResnetTensorNavActorCritic.forward`` accepts it in place of the fp32 tensor."""

    def __init__(self, data: torch.Tensor, T: int, N: int):
        self.data, self.T, self.N = data, T, N


class _ACFunction(torch.autograd.Function):
    """forward = embclip_ac_forward, backward = embclip_ac_backward; gradients arrive as one flat tensor."""

    @staticmethod
    def forward(ctx, flat_params, model, feats16, goals, masks, h0, T, N, need_grad):
        plan, dev = model._plan, flat_params.device
        A, H = plan.cfg["num_actions"], plan.cfg["hidden"]
        ws = model._workspace(T, N)
        logits = torch.empty(T, N, A, dtype=torch.float32, device=dev)
        values = torch.empty(T, N, dtype=torch.float32, device=dev)
        h_last = torch.empty(N, H, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(plan.lib.embclip_ac_forward(plan._h, flat_params.data_ptr(), feats16.data_ptr(), goals.data_ptr(),
                                                   masks.data_ptr(), h0.data_ptr(), T, N, logits.data_ptr(), values.data_ptr(),
                                                   h_last.data_ptr(), ws.data_ptr(), ws.numel(), int(need_grad), _stream(dev)))
        ctx.model, ctx.T, ctx.N = model, T, N
        ctx.ws_gen = model._touch_workspace()          # the forward's intermediates live in the model's shared workspace
        ctx.save_for_backward(flat_params, feats16, goals, masks, h0)
        return logits, values, h_last

    @staticmethod
    def backward(ctx, dlogits, dvalues, dh_last):
        flat_params, feats16, goals, masks, h0 = ctx.saved_tensors
        model, T, N = ctx.model, ctx.T, ctx.N
        plan, dev = model._plan, flat_params.device
        A = plan.cfg["num_actions"]
        dl = _f32c(dlogits) if dlogits is not None else torch.zeros(T, N, A, dtype=torch.float32, device=dev)
        dv = _f32c(dvalues) if dvalues is not None else torch.zeros(T, N, dtype=torch.float32, device=dev)
        dh = _f32c(dh_last) if dh_last is not None else None
        grads = torch.zeros_like(flat_params)
        ws = model._workspace(T, N)
        with torch.cuda.device(dev):
            if model._ws_gen != ctx.ws_gen:
                # another forward()/act()/update() ran on this module since (a second minibatch, a bootstrap-value forward, a
                # larger block that re-allocated the workspace): the saved intermediates are gone -- recompute them from the
                # saved inputs instead of differentiating someone else's activations
                H = plan.cfg["hidden"]
                scratch = torch.empty(T * N * (A + 1) + N * H, dtype=torch.float32, device=dev)
                _lib.check(plan.lib.embclip_ac_forward(plan._h, flat_params.data_ptr(), feats16.data_ptr(), goals.data_ptr(),
                                                       masks.data_ptr(), h0.data_ptr(), T, N, scratch.data_ptr(),
                                                       scratch[T * N * A:].data_ptr(), scratch[T * N * (A + 1):].data_ptr(),
                                                       ws.data_ptr(), ws.numel(), 1, _stream(dev)))
                model.recomputed_backwards += 1
            model._touch_workspace()
            _lib.check(plan.lib.embclip_ac_backward(plan._h, flat_params.data_ptr(), feats16.data_ptr(), goals.data_ptr(),
                                                    masks.data_ptr(), h0.data_ptr(), T, N, dl.data_ptr(), dv.data_ptr(),
                                                    dh.data_ptr() if dh is not None else None, grads.data_ptr(), ws.data_ptr(),
                                                    ws.numel(), _stream(dev)))
        return grads, None, None, None, None, None, None, None, None


class ResnetTensorNavActorCritic(nn.Module):
    """Drop-in for allenact's ``ResnetTensorNavActorCritic`` (RGB-only CLIP-ResNet tensor + goal object type):

        compressor conv1x1 2048->128, ReLU, 128->32, ReLU  ||  goal Embedding(12, 32) broadcast over 7x7
        -> combiner conv1x1 64->128, ReLU, 128->32 -> flatten 1568 -> RNNStateEncoder(GRU 512) -> actor / critic

    Parameters are views into ONE flat fp32 tensor (``self.flat_params``) laid out by the library, exposed under
    the upstream names (``goal_visual_encoder.resnet_compressor.0.weight``, ``state_encoder.rnn.weight_ih_l0``,
    ``actor.linear.weight`` ...) so AllenAct checkpoints load through ``load_state_dict`` unchanged."""

    def __init__(self, action_space: Any = None, observation_space: Any = None, goal_sensor_uuid: str = "goal_object_type_ind",
                 rgb_resnet_preprocessor_uuid: str = "rgb_clip_resnet", hidden_size: int = 512, goal_dims: int = 32,
                 resnet_compressor_hidden_out_dims: Tuple[int, int] = (128, 32),
                 combiner_hidden_out_dims: Tuple[int, int] = (128, 32), num_actions: Optional[int] = None,
                 num_goals: Optional[int] = None, resnet_tensor_shape: Tuple[int, int, int] = (2048, 7, 7),
                 device: Any = "cuda:0", seed: Optional[int] = None, trainable_masked_hidden_state: bool = False):
        super().__init__()
        if num_actions is None:
            num_actions = int(getattr(action_space, "n", 6))
        if num_goals is None:
            try:
                num_goals = int(observation_space.spaces[goal_sensor_uuid].n)
            except Exception:
                num_goals = 12
        self.action_space, self.observation_space = action_space, observation_space
        self.goal_uuid, self.resnet_uuid = goal_sensor_uuid, rgb_resnet_preprocessor_uuid
        self.hidden_size = hidden_size
        self.resnet_tensor_shape = tuple(resnet_tensor_shape)
        cfg = dict(feat_channels=resnet_tensor_shape[0], feat_pixels=resnet_tensor_shape[1] * resnet_tensor_shape[2],
                   compress_hidden=resnet_compressor_hidden_out_dims[0], compress_out=resnet_compressor_hidden_out_dims[1],
                   goal_dims=goal_dims, combine_hidden=combiner_hidden_out_dims[0], combine_out=combiner_hidden_out_dims[1],
                   hidden=hidden_size, num_actions=num_actions, num_goals=num_goals,
                   trainable_masked_hidden_state=int(bool(trainable_masked_hidden_state)))
        self._plan = _ACPlan(cfg)
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("embclip_b200 ResnetTensorNavActorCritic has no CPU path: construct it on a CUDA device")
        self.flat_params = nn.Parameter(torch.zeros(self._plan.n_floats, dtype=torch.float32, device=dev))
        self._ws: Optional[torch.Tensor] = None
        self._ws_gen = 0
        self.recomputed_backwards = 0          # backward passes that had to re-run their forward (workspace overwritten in between)
        self._init_parameters(seed)

    # ------------------------------------------------------------------ parameters under upstream names
    def named_views(self) -> Dict[str, torch.Tensor]:
        """Upstream-named views of the flat parameter tensor.  They are views of the Parameter itself (taken under
        ``no_grad``), so an in-place write through one bumps ``flat_params._version`` and ``params_version()`` sees it
        (views of ``.data`` would not share the counter)."""
        with torch.no_grad():
            return {name: self.flat_params[off:off + n].view(shape) for name, shape, off, n in self._plan.params}

    def _touch_workspace(self) -> int:
        """Every call that overwrites the shared workspace takes a new generation number (see ``_ACFunction.backward``)."""
        self._ws_gen += 1
        return self._ws_gen

    def _init_parameters(self, seed: Optional[int]) -> None:
        """Upstream initialisers: conv / embedding = torch defaults; GRU weights orthogonal, biases 0
        (RNNStateEncoder.layer_init); actor orthogonal gain 0.01, critic orthogonal, biases 0."""
        g = torch.Generator().manual_seed(seed) if seed is not None else None
        v = self.named_views()
        with torch.no_grad():
            for name, t in v.items():
                cpu = torch.empty(t.shape, dtype=torch.float32)
                if name.endswith("embed_class.weight"):
                    cpu.normal_(0, 1, generator=g)
                elif name.endswith("init_hidden_state"):                # RNNStateEncoder: 0.1 * randn
                    cpu.normal_(0, 0.1, generator=g)
                elif "rnn.weight" in name:
                    nn.init.orthogonal_(cpu, generator=g)
                elif name == "actor.linear.weight":
                    nn.init.orthogonal_(cpu, gain=0.01, generator=g)
                elif name == "critic.fc.weight":
                    nn.init.orthogonal_(cpu, generator=g)
                elif name.endswith(".weight"):                      # conv1x1: kaiming_uniform(a=sqrt(5)) = U(+-1/sqrt(fan_in))
                    bound = (1.0 / cpu[0].numel()) ** 0.5
                    cpu.uniform_(-bound, bound, generator=g)
                elif "rnn.bias" in name or name in ("actor.linear.bias", "critic.fc.bias"):
                    cpu.zero_()
                else:                                               # conv bias: U(+-1/sqrt(fan_in)) of its weight
                    w = v[name[:-len("bias")] + "weight"]
                    bound = (1.0 / w[0].numel()) ** 0.5
                    cpu.uniform_(-bound, bound, generator=g)
                t.copy_(cpu)
        self.mark_params_changed()

    # nn.Module checkpoint protocol under the upstream key names.  Implemented at the _save_to / _load_from level so the
    # module also round-trips as a SUBMODULE (a parent's state_dict() passes destination / prefix and ignores the return value).
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for k, t in self.named_views().items():
            destination[prefix + k] = t if keep_vars else t.detach().clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        v = self.named_views()
        with torch.no_grad():
            for k, t in v.items():
                key = prefix + k
                if key not in state_dict:
                    missing_keys.append(key)
                    continue
                src = state_dict[key]
                if tuple(src.shape) != tuple(t.shape):
                    error_msgs.append(f"size mismatch for {key}: copying a param with shape {tuple(src.shape)} from checkpoint, "
                                      f"the shape in current model is {tuple(t.shape)}.")
                    continue
                t.copy_(src.to(t.device, torch.float32))
        if strict:                           # this module has no children: every key under its prefix is its own
            unexpected_keys.extend(key for key in state_dict if key.startswith(prefix) and key[len(prefix):] not in v)
        self.mark_params_changed()           # the fp16 weight layouts cached by act() are stale now

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        if assign:
            raise RuntimeError("load_state_dict(assign=True): parameters are views of one flat buffer and cannot be re-assigned")
        return super().load_state_dict(state_dict, strict=strict)

    # ------------------------------------------------------------------ AllenAct ActorCriticModel surface
    @property
    def recurrent_hidden_state_size(self) -> int:
        return self.hidden_size

    @property
    def num_recurrent_layers(self) -> int:
        return 1

    def _recurrent_memory_specification(self):
        return dict(rnn=((("layer", 1), ("sampler", None), ("hidden", self.hidden_size)), torch.float32))

    def _workspace(self, T: int, N: int) -> torch.Tensor:
        need = self._plan.workspace_bytes(T, N)
        if self._ws is None or self._ws.numel() < need or self._ws.device != self.flat_params.device:
            self._ws = None
            raw = torch.empty(need + 1024, dtype=torch.uint8, device=self.flat_params.device)   # small blocks are only 512-B aligned
            off = (-raw.data_ptr()) % 1024
            self._ws = raw[off:off + need]
        return self._ws

    def activations(self, T: int, N: int) -> Dict[str, torch.Tensor]:
        """Views of the workspace intermediates of the LAST forward / backward on a [T, N] block (test hook)."""
        ws = self._workspace(T, N)
        out = {}
        for i in range(_lib.check(self._plan.lib.embclip_ac_num_acts(self._plan._h))):
            ai = _lib.ActInfo()
            _lib.check(self._plan.lib.embclip_ac_act_info(self._plan._h, T, N, i, C.byref(ai)))
            dt = torch.float16 if ai.dtype == _lib.DTYPE_F16 else torch.float32
            nbytes = ai.w * ai.c * (2 if dt == torch.float16 else 4)
            out[ai.name.decode()] = ws[ai.offset:ai.offset + nbytes].view(dt).view(ai.w, ai.c)
        return out

    def pack_features(self, feats: torch.Tensor) -> PackedFeatures:
        """fp32 [T, N, C, H, W] -> PackedFeatures (fp16 pixel rows)."""
        C_, Hh, Ww = self.resnet_tensor_shape
        if feats.dim() != 5 or tuple(feats.shape[2:]) != (C_, Hh, Ww):
            raise ValueError(f"features must be [steps, samplers, {C_}, {Hh}, {Ww}], got {tuple(feats.shape)}")
        if feats.device != self.flat_params.device:
            raise ValueError(f"features on {feats.device}, model on {self.flat_params.device}")
        T, N = feats.shape[:2]
        x = _f32c(feats)
        out = torch.empty(T * N * Hh * Ww, C_, dtype=torch.float16, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(self._plan.lib.embclip_ac_pack_features(self._plan._h, x.data_ptr(), T * N, out.data_ptr(), _stream(x.device)))
        return PackedFeatures(out, T, N)

    # ------------------------------------------------------------------ rollout step
    def params_version(self) -> int:
        """A number that changes whenever the parameter VALUES may have changed: torch's in-place counter of the flat tensor
        (optimizers, load_state_dict, .data writes through views) plus ``_params_epoch``, which ``PPOTrainer`` bumps after its
        raw-pointer Adam step.  ``act`` passes it to the library so the fp16 weight layouts are rebuilt once per update, not
        once per rollout step."""
        key = (self.flat_params._version, self.flat_params.data_ptr(), getattr(self, "_params_epoch", 0))
        if key != getattr(self, "_ver_key", None):
            self._ver_key = key
            self._ver_id = getattr(self, "_ver_id", 0) + 1
        return self._ver_id

    def mark_params_changed(self) -> None:
        self._params_epoch = getattr(self, "_params_epoch", 0) + 1

    @torch.no_grad()
    def act(self, feats_rows: torch.Tensor, goals: torch.Tensor, masks: torch.Tensor, memory: torch.Tensor, uniforms: torch.Tensor,
            actions: Optional[torch.Tensor] = None, action_log_probs: Optional[torch.Tensor] = None,
            values: Optional[torch.Tensor] = None, memory_out: Optional[torch.Tensor] = None):
        """One rollout step in one library call (OnPolicyRLEngine.act [UPSTREAM]: forward with steps = 1, then
        ``distributions.sample()`` and ``log_prob``): feats_rows fp16 [N*P, C] (``ClipRN50Encoder.encode_rows`` /
        ``pack_features``), goals [N], masks [N] (0 = episode start), memory fp32 [N, H], uniforms fp32 [N] in [0, 1)
        -> (actions int64 [N], log-probs [N], values [N], new memory [N, H], logits [N, A]).  No autograd."""
        dev = self.flat_params.device
        N, H, A = memory.reshape(-1, self.hidden_size).shape[0], self.hidden_size, self._plan.cfg["num_actions"]
        C_, Hh, Ww = self.resnet_tensor_shape
        if feats_rows.dtype != torch.float16 or feats_rows.numel() != N * Hh * Ww * C_ or not feats_rows.is_contiguous():
            raise ValueError(f"feats_rows must be contiguous fp16 [{N * Hh * Ww}, {C_}], got {feats_rows.dtype} {tuple(feats_rows.shape)}")
        for name, t in (("feats_rows", feats_rows), ("goals", goals), ("masks", masks), ("memory", memory), ("uniforms", uniforms)):
            if t.device != dev:
                raise ValueError(f"{name} on {t.device}, model on {dev}")
        g = goals.reshape(N).to(torch.int64).contiguous()
        m = _f32c(masks.reshape(N))
        h0 = _f32c(memory.reshape(N, H))
        u = _f32c(uniforms.reshape(N))
        mk = lambda t, shape, dt: t if t is not None else torch.empty(*shape, dtype=dt, device=dev)
        actions, action_log_probs = mk(actions, (N,), torch.int64), mk(action_log_probs, (N,), torch.float32)
        values, memory_out = mk(values, (N,), torch.float32), mk(memory_out, (N, H), torch.float32)
        for name, t, dt, n in (("actions", actions, torch.int64, N), ("action_log_probs", action_log_probs, torch.float32, N),
                               ("values", values, torch.float32, N), ("memory_out", memory_out, torch.float32, N * H)):
            if t.dtype != dt or t.numel() != n or not t.is_contiguous() or t.device != dev:
                raise ValueError(f"{name} must be a contiguous {dt} tensor of {n} elements on {dev}")
        logits = torch.empty(N, A, dtype=torch.float32, device=dev)
        ws = self._workspace(1, N)
        self._touch_workspace()
        with torch.cuda.device(dev):
            _lib.check(self._plan.lib.embclip_ac_act(self._plan._h, self.flat_params.data_ptr(), self.params_version(), feats_rows.data_ptr(),
                                                     g.data_ptr(), m.data_ptr(), h0.data_ptr(), N, u.data_ptr(), actions.data_ptr(),
                                                     action_log_probs.data_ptr(), values.data_ptr(), memory_out.data_ptr(),
                                                     logits.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)))
        return actions, action_log_probs, values, memory_out, logits

    def forward_tensors(self, feats, goals: torch.Tensor, memory: torch.Tensor, masks: torch.Tensor):
        """-> (logits [T,N,A], values [T,N], h_last [N,H]); differentiable w.r.t. the parameters."""
        pf = feats if isinstance(feats, PackedFeatures) else self.pack_features(feats)
        T, N = pf.T, pf.N
        dev = self.flat_params.device
        g = goals.reshape(T, N).to(dev, torch.int64).contiguous()
        m = _f32c(masks.reshape(T, N).to(dev))
        h0 = _f32c(memory.reshape(N, self.hidden_size).to(dev))
        need_grad = torch.is_grad_enabled() and self.flat_params.requires_grad    # (grad mode is off inside Function.forward)
        return _ACFunction.apply(self.flat_params, self, pf.data, g, m, h0, T, N, need_grad)

    def forward(self, observations: Dict[str, Any], memory: Any, prev_actions: Optional[torch.Tensor],
                masks: torch.Tensor) -> Tuple[ActorCriticOutput, Any]:
        h0 = memory.tensor("rnn") if hasattr(memory, "tensor") else memory
        logits, values, h_last = self.forward_tensors(observations[self.resnet_uuid], observations[self.goal_uuid], h0, masks)
        out = ActorCriticOutput(distributions=CategoricalDistr(logits=logits), values=values.unsqueeze(-1), extras={})
        h_new = h_last.unsqueeze(0)
        if hasattr(memory, "set_tensor"):
            return out, memory.set_tensor("rnn", h_new)
        return out, h_new


# ---------------------------------------------------------------------------------------------------
# RolloutStorage.compute_returns / advantage normalisation
# ---------------------------------------------------------------------------------------------------
def compute_returns_gae(rewards: torch.Tensor, value_preds: torch.Tensor, masks: torch.Tensor, next_value: torch.Tensor,
                        gamma: float = 0.99, tau: float = 0.95, eps: float = 1e-5):
    """rewards [T,N,1]; value_preds [T+1,N,1] (row T ignored, replaced by next_value [N,1]); masks [T+1,N,1]
    -> (returns [T,N,1], advantages [T,N,1], normalised advantages [T,N,1]); one kernel, no host loop."""
    lib = _lib.load()
    dev = rewards.device
    if dev.type != "cuda":
        raise RuntimeError("embclip_b200.compute_returns_gae: CUDA tensors only (no CPU path)")
    T, N = rewards.shape[:2]
    r = _f32c(rewards.reshape(T, N))
    v = _f32c(value_preds.reshape(T + 1, N)).clone()
    v[T] = next_value.reshape(N)
    m = _f32c(masks.reshape(T + 1, N))
    ret = torch.empty(T, N, dtype=torch.float32, device=dev)
    adv = torch.empty_like(ret)
    nadv = torch.empty_like(ret)
    with torch.cuda.device(dev):
        _lib.check(lib.embclip_gae(r.data_ptr(), v.data_ptr(), m.data_ptr(), T, N, gamma, tau, ret.data_ptr(), adv.data_ptr(),
                                   nadv.data_ptr(), eps, _stream(dev)))
    return ret.unsqueeze(-1), adv.unsqueeze(-1), nadv.unsqueeze(-1)


# ---------------------------------------------------------------------------------------------------
# OnPolicyTrainer.update / backprop_step
# ---------------------------------------------------------------------------------------------------
class LinearDecay:
    """allenact.utils.experiment_utils.LinearDecay [UPSTREAM]: multiplier going linearly from `startp` to `endp` over
    `steps` environment steps, constant afterwards.  ``PPOTrainer(lr_schedule=LinearDecay(steps=ppo_steps))``."""

    def __init__(self, steps: int, startp: float = 1.0, endp: float = 0.0):
        self.steps, self.startp, self.endp = int(steps), float(startp), float(endp)

    def __call__(self, epoch: int) -> float:
        epoch = max(min(int(epoch), self.steps), 0)
        return self.startp + (self.endp - self.startp) * (epoch / float(self.steps))


class PPOTrainer:
    """update_repeats x (forward, PPO.loss, backward, gradient all-reduce, clip_grad_norm_(0.5), Adam(lr)).

    The gradient of the local batch is pre-scaled by local/global batch size and SUM all-reduced as one flat
    bucket (upstream: one async all_reduce per parameter tensor), then every rank applies the identical step."""

    def __init__(self, model: ResnetTensorNavActorCritic, lr: float = 3e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: float = 0.5, update_repeats: int = 4, clip_param: float = 0.1, value_loss_coef: float = 0.5,
                 entropy_coef: float = 0.01, process_group: Any = None, lr_schedule: Optional[Any] = None,
                 num_mini_batch: int = 1, distributed: bool = True, seed: int = 0):
        """lr_schedule: a multiplier ``f(total_steps) -> float`` on the base lr (``LinearDecay``), applied the way AllenAct's
        ``LambdaLR(optimizer, lr_lambda=LinearDecay(steps))`` + ``lr_scheduler.step(epoch=total_steps)`` does: the lr of an
        update is ``lr * f(environment steps collected before it)``.  num_mini_batch: samplers are split into that many
        contiguous chunks per update repeat, visited in shuffled order (RolloutStorage.recurrent_generator [UPSTREAM]).
        distributed=False keeps the gradient local even inside an initialised process group (single-rank reference runs)."""
        self.model, self.base_lr, self.betas, self.eps = model, lr, betas, eps
        self.lr_schedule = lr_schedule
        self.total_steps = 0                   # environment steps (global rows) consumed by update() so far
        self.max_grad_norm, self.update_repeats = max_grad_norm, update_repeats
        self.clip_param, self.value_loss_coef, self.entropy_coef = clip_param, value_loss_coef, entropy_coef
        self.process_group = process_group
        self.distributed = bool(distributed)
        self.num_mini_batch = int(num_mini_batch)
        if self.num_mini_batch < 1:
            raise ValueError("num_mini_batch must be >= 1")
        import random
        self._rng = random.Random(seed)         # mini-batch order (upstream: python's global `random.shuffle`)
        p = model.flat_params
        self.grads = torch.zeros_like(p.data)
        self.exp_avg = torch.zeros_like(p.data)
        self.exp_avg_sq = torch.zeros_like(p.data)
        self.sumsq = torch.zeros(1024, dtype=torch.float32, device=p.device)   # [0] = sum g^2; the rest is the kernel's scratch (EMBCLIP_SUMSQ_FLOATS)
        self.loss_sums = torch.zeros(3, dtype=torch.float32, device=p.device)
        self.step_count = 0
        self.kernel_launch_estimate = 0

    @property
    def lr(self) -> float:
        """Learning rate the NEXT update will use."""
        return self.base_lr * (float(self.lr_schedule(self.total_steps)) if self.lr_schedule is not None else 1.0)

    @lr.setter
    def lr(self, value: float) -> None:
        self.base_lr = float(value)

    def _world(self) -> int:
        import torch.distributed as dist
        if not self.distributed:
            return 1
        return dist.get_world_size(self.process_group) if dist.is_available() and dist.is_initialized() else 1

    def _step_block(self, bk: Dict[str, Any], T: int, mb_grows: float, lr: float) -> None:
        """One backprop_step on a contiguous [T, n] block: forward, PPO loss (+ dlogits / dvalues), backward, gradient
        all-reduce, global-norm clip + Adam.  bk: n, feats (fp16 rows), goals, masks, h0, actions, old_lp, old_v, rets, nadv."""
        from .distributed import allreduce_flat_
        mdl, plan = self.model, self.model._plan
        lib, dev = plan.lib, mdl.flat_params.device
        A = plan.cfg["num_actions"]
        n = bk["n"]
        P = mdl.flat_params.data
        st = _stream(dev)
        logits = torch.empty(T, n, A, dtype=torch.float32, device=dev)
        values = torch.empty(T, n, dtype=torch.float32, device=dev)
        ws = mdl._workspace(T, n)
        mdl._touch_workspace()
        with torch.cuda.device(dev):
            _lib.check(lib.embclip_ac_forward(plan._h, P.data_ptr(), bk["feats"].data_ptr(), bk["goals"].data_ptr(),
                                              bk["masks"].data_ptr(), bk["h0"].data_ptr(), T, n, logits.data_ptr(),
                                              values.data_ptr(), None, ws.data_ptr(), ws.numel(), 1, st))
            _lib.check(lib.embclip_ac_ppo_loss(plan._h, P.data_ptr(), T, n, bk["actions"].data_ptr(), bk["old_lp"].data_ptr(),
                                               bk["nadv"].data_ptr(), bk["old_v"].data_ptr(), bk["rets"].data_ptr(),
                                               self.clip_param, self.value_loss_coef, self.entropy_coef, 1.0 / mb_grows,
                                               logits.data_ptr(), values.data_ptr(), self.loss_sums.data_ptr(), ws.data_ptr(),
                                               ws.numel(), st))
            self.grads.zero_()
            _lib.check(lib.embclip_ac_backward(plan._h, P.data_ptr(), bk["feats"].data_ptr(), bk["goals"].data_ptr(),
                                               bk["masks"].data_ptr(), bk["h0"].data_ptr(), T, n, None, None, None,
                                               self.grads.data_ptr(), ws.data_ptr(), ws.numel(), st))
            if self.distributed:
                allreduce_flat_(self.grads, self.process_group)   # the path's one collective (no-op when world == 1)
            self.step_count += 1
            _lib.check(lib.embclip_sumsq_f32(self.grads.data_ptr(), self.grads.numel(), self.sumsq.data_ptr(), st))
            _lib.check(lib.embclip_adam_clip_step(P.data_ptr(), self.grads.data_ptr(), self.exp_avg.data_ptr(),
                                                  self.exp_avg_sq.data_ptr(), P.numel(), self.sumsq.data_ptr(),
                                                  self.max_grad_norm, lr, self.betas[0], self.betas[1], self.eps,
                                                  self.step_count, st))
        mdl.mark_params_changed()                              # raw-pointer update: torch's version counter does not see it

    def _loss_info(self, rows: int) -> Dict[str, torch.Tensor]:
        s = self.loss_sums / rows
        return {"action": s[0], "value": s[1], "entropy": -s[2],
                "total": s[0] + self.value_loss_coef * s[1] - self.entropy_coef * s[2], "grad_norm": self.sumsq.sqrt()[0]}

    def update_from_storage(self, storage: Any, global_rows: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """OnPolicyTrainer.update(rollouts) [UPSTREAM engine.py]: ``update_repeats`` x ``storage.recurrent_generator(...,
        num_mini_batch)`` x backprop_step, on an ``embclip_b200.storage.RolloutStorage`` whose ``compute_returns`` has run
        (advantages normalised over this rank's block, as upstream).  The caller runs ``storage.after_update()`` afterwards."""
        mdl = self.model
        dev = mdl.flat_params.device
        T, N, H = storage.num_steps, storage.num_samplers, mdl.hidden_size
        world = self._world()
        grows = global_rows if global_rows is not None else T * N * world
        lr = self.lr
        rows = T * N
        packed_once: Dict[Any, PackedFeatures] = {}        # fp32 storage (the AllenAct flow): pack each chunk once per update, not per repeat
        for _ in range(self.update_repeats):
            for batch in storage.recurrent_generator(None, None, None, self.num_mini_batch):
                a, b = batch["samplers"]
                n = b - a
                feats = batch["observations"][mdl.resnet_uuid]
                if isinstance(feats, PackedFeatures):
                    pf = feats
                else:
                    if (a, b) not in packed_once:
                        packed_once[(a, b)] = mdl.pack_features(feats)
                    pf = packed_once[(a, b)]
                bk = dict(n=n, feats=pf.data, goals=batch["observations"][mdl.goal_uuid].reshape(T, n).to(dev, torch.int64).contiguous(),
                          masks=_f32c(batch["masks"].reshape(T, n)), h0=_f32c(batch["memory"].tensor("rnn").reshape(n, H)),
                          actions=batch["actions"].reshape(T, n).to(dev, torch.int64).contiguous(),
                          old_lp=_f32c(batch["old_action_log_probs"].reshape(T, n)), old_v=_f32c(batch["values"].reshape(T, n)),
                          rets=_f32c(batch["returns"].reshape(T, n)), nadv=_f32c(batch["norm_adv_targ"].reshape(T, n)))
                self._step_block(bk, T, grows * n / N, lr)
                rows = T * n
        self.total_steps += int(grows)
        return self._loss_info(rows)

    def update(self, rollout: Dict[str, Any], global_rows: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """rollout: features (fp32 [T,N,2048,7,7] or PackedFeatures), goals [T,N], masks [T,N,1], memory [1,N,H],
        actions [T,N], old_action_log_probs [T,N], values [T,N,1], returns [T,N,1], norm_adv_targ [T,N,1].
        Returns the last pass's loss terms as 0-d device tensors (no host sync inside)."""
        from .distributed import allreduce_flat_
        mdl, plan = self.model, self.model._plan
        lib, dev = plan.lib, mdl.flat_params.device
        pf = rollout["features"] if isinstance(rollout["features"], PackedFeatures) else mdl.pack_features(rollout["features"])
        T, N = pf.T, pf.N
        A, H = plan.cfg["num_actions"], plan.cfg["hidden"]
        goals = rollout["goals"].reshape(T, N).to(dev, torch.int64).contiguous()
        masks = _f32c(rollout["masks"].reshape(T, N))
        h0 = _f32c(rollout["memory"].reshape(N, H))
        actions = rollout["actions"].reshape(T, N).to(dev, torch.int64).contiguous()
        old_lp = _f32c(rollout["old_action_log_probs"].reshape(T, N))
        old_v = _f32c(rollout["values"].reshape(T, N))
        rets = _f32c(rollout["returns"].reshape(T, N))
        nadv = _f32c(rollout["norm_adv_targ"].reshape(T, N))
        world = self._world()
        rows = T * N
        grows = global_rows if global_rows is not None else rows * world
        P = mdl.flat_params.data
        st = _stream(dev)
        # mini-batches = contiguous chunks of samplers (RolloutStorage.recurrent_generator [UPSTREAM]: chunk bounds
        # round(linspace(0, N, num_mini_batch + 1)), chunk ORDER shuffled per repeat); one chunk = the whole [T, N] block
        nmb = self.num_mini_batch
        if nmb > N:
            raise ValueError(f"num_mini_batch {nmb} exceeds the {N} samplers of this rank")
        bounds = [int(round(i * N / nmb)) for i in range(nmb + 1)]
        chunks = [(a, b) for a, b in zip(bounds[:-1], bounds[1:])]
        C_ = pf.data.shape[-1]

        def block(a: int, b: int):
            if (a, b) == (0, N):
                return dict(n=N, feats=pf.data, goals=goals, masks=masks, h0=h0, actions=actions, old_lp=old_lp, old_v=old_v,
                            rets=rets, nadv=nadv)
            cut = lambda t: t[:, a:b].contiguous()
            return dict(n=b - a, feats=pf.data.view(T, N, -1)[:, a:b].reshape(-1, C_).contiguous(), goals=cut(goals), masks=cut(masks),
                        h0=h0[a:b].contiguous(), actions=cut(actions), old_lp=cut(old_lp), old_v=cut(old_v), rets=cut(rets), nadv=cut(nadv))

        blocks = {c: block(*c) for c in chunks}            # sliced once, reused by every repeat
        lr = self.lr
        for _ in range(self.update_repeats):
            order = list(chunks)
            if nmb > 1:
                self._rng.shuffle(order)
            for c in order:
                bk = blocks[c]
                self._step_block(bk, T, grows * bk["n"] / N, lr)      # rows of this mini-batch over all ranks (equal splits on every rank)
                rows = T * bk["n"]
        self.total_steps += int(grows)                                  # lr_scheduler.step(epoch=total_steps) [UPSTREAM]
        return self._loss_info(rows)
