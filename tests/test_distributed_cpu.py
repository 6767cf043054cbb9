"""N > 1 host logic on CPU (gloo, world_size 2): sampler sharding, the flat gradient bucket and its SUM all-reduce
reproduce the single-process gradient of the global batch, and both ranks end with identical parameters."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_samplers_matches_allenact_split():
    from embclip_b200.distributed import shard_samplers
    assert [shard_samplers(60, 8, r)[1] for r in range(8)] == [8, 8, 8, 8, 7, 7, 7, 7]
    cover = []
    for r in range(8):
        s, c = shard_samplers(60, 8, r)
        cover += list(range(s, s + c))
    assert cover == list(range(60))
    assert shard_samplers(60, 1, 0) == (0, 60)
    assert shard_samplers(3, 4, 3) == (3, 0)
    with pytest.raises(ValueError):
        shard_samplers(60, 8, 8)


def _rollout(T, N, seed):
    g = torch.Generator().manual_seed(seed)
    return dict(features=torch.randn(T, N, 2048, 7, 7, generator=g).relu_(), goals=torch.randint(0, 12, (T, N), generator=g),
                masks=(torch.rand(T, N, 1, generator=g) > 0.1).float(), memory=torch.randn(1, N, 512, generator=g) * 0.3,
                actions=torch.randint(0, 6, (T, N), generator=g), old_action_log_probs=-1.8 + 0.1 * torch.randn(T, N, generator=g),
                values=torch.randn(T, N, 1, generator=g), returns=torch.randn(T, N, 1, generator=g),
                norm_adv_targ=torch.randn(T, N, 1, generator=g))


def _shard(ro, s, c):
    out = {}
    for k, v in ro.items():
        out[k] = v[:, s:s + c].contiguous()
    return out


def _loss_sum(model, ro):
    """Sum over rows of the per-row PPO objective (the mean is taken over the GLOBAL batch by the caller)."""
    from oracle.allenact_models import ppo_loss
    distr, v, _ = model({model.rgb_uuid: ro["features"], model.goal_uuid: ro["goals"]}, ro["memory"], None, ro["masks"])
    total, _ = ppo_loss(distr, v, ro)
    return total * ro["actions"].numel()


def _worker(rank, world, port, T, N, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from embclip_b200.distributed import allreduce_flat_, flatten_grads, global_rows, shard_samplers, unflatten_to_grads
        from oracle.allenact_models import ResnetTensorNavActorCritic
        torch.manual_seed(5)                                  # identical replicas
        model = ResnetTensorNavActorCritic()
        opt = torch.optim.Adam(model.parameters(), lr=3e-4)
        s, c = shard_samplers(N, world, rank)
        ro = _shard(_rollout(T, N, seed=1), s, c)
        rows = global_rows(T * c)
        assert rows == T * N
        for _ in range(2):
            opt.zero_grad()
            (_loss_sum(model, ro) / rows).backward()          # local gradient, pre-scaled by 1 / global rows
            flat = allreduce_flat_(flatten_grads(model.parameters()))
            unflatten_to_grads(flat, model.parameters())
            torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
            opt.step()
        ret[rank] = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    finally:
        dist.destroy_process_group()


def test_flat_bucket_allreduce_equals_global_batch_gradient():
    T, N, world = 3, 5, 2
    with socket.socket() as s_:
        s_.bind(("127.0.0.1", 0))
        port = s_.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, T, N, ret), nprocs=world, join=True)
    assert torch.equal(ret[0], ret[1]), "ranks diverged"
    # single process on the whole batch
    from oracle.allenact_models import ResnetTensorNavActorCritic
    torch.manual_seed(5)
    model = ResnetTensorNavActorCritic()
    opt = torch.optim.Adam(model.parameters(), lr=3e-4)
    ro = _rollout(T, N, seed=1)
    for _ in range(2):
        opt.zero_grad()
        (_loss_sum(model, ro) / (T * N)).backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
        opt.step()
    single = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert (single - ret[0]).abs().max().item() <= 2e-5       # Adam is sign-like on the first steps: lr-sized slack on near-zero grads
