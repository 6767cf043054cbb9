"""Multi-GPU host logic of the hot path: one process per GPU, samplers sharded across
ranks, frozen encoder replicated (no collective), and ONE collective per update pass -- a SUM all-reduce of the
flat gradient bucket (upstream allenact OnPolicyTrainer.backprop_step issues one async all_reduce per parameter
tensor instead).  Pure torch.distributed: NCCL on GPUs, gloo in the CPU tests (tests/test_distributed_cpu.py)."""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_samplers(total: int, world: int, rank: int) -> Tuple[int, int]:
    """AllenAct's even split of `total` samplers over `world` trainers: the first `total % world` ranks get one
    extra (60 over 8 -> 8,8,8,8,7,7,7,7).  Returns (first sampler index, count)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} for world size {world}")
    base, extra = divmod(total, world)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def global_rows(local_rows: int, group=None) -> int:
    """Sum of the per-rank batch rows (steps x local samplers): the denominator of the global loss mean."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_rows
    t = torch.tensor([local_rows], dtype=torch.int64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def allreduce_flat_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of one flat bucket.  The caller has already scaled its local gradient by
    1 / global_rows (the loss is a mean over the GLOBAL batch), so after the sum every rank holds the global
    gradient and applies an identical clip + Adam step."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def flatten_grads(params: Iterable[torch.nn.Parameter], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Pack p.grad of every parameter into one contiguous bucket (for models whose parameters are not already views
    of a flat buffer, e.g. the oracle in the CPU tests)."""
    ps: List[torch.nn.Parameter] = [p for p in params]
    n = sum(p.numel() for p in ps)
    if out is None:
        out = torch.zeros(n, dtype=ps[0].dtype, device=ps[0].device)
    off = 0
    for p in ps:
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        out[off:off + p.numel()].copy_(g.reshape(-1))
        off += p.numel()
    return out


def unflatten_to_grads(flat: torch.Tensor, params: Iterable[torch.nn.Parameter]) -> None:
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view_as(p).clone()
        off += p.numel()
