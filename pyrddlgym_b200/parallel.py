"""Multi-GPU plumbing: env sharding and the one collective of the path.

Environments never interact (each reference simulator is a closed object), so the env axis
is split into contiguous blocks, one process per GPU, with no data-path collective. Philox
counters use the global env id (``env_offset + local index``), hence results do not depend on
the number of ranks. The only exchange is the reduction of per-env returns into the
statistics ``BaseAgent.evaluate`` reports (/root/reference/pyRDDLGym/core/policy.py:120-126):
two all-reduces of a few scalars per rollout (NCCL over NVLink on GPUs, gloo in CPU tests).
"""
import math

import torch
import torch.distributed as dist


def shard_envs(global_envs, world_size, rank):
    """Contiguous block of env ids owned by ``rank``: returns (n_local, env_offset)."""
    base, extra = divmod(int(global_envs), int(world_size))
    n_local = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return n_local, offset


def reduce_return_stats(returns, group=None):
    """returns: f64[n_local] per-env (discounted) returns of this rank's shard. Gives the dict of
    BaseAgent.evaluate -- mean, std (population), min, max -- over ALL ranks, plus the count."""
    r = returns.to(torch.float64)
    s = torch.stack([r.sum(), (r * r).sum(), torch.tensor(float(r.numel()), dtype=torch.float64, device=r.device)])
    big = torch.tensor(float('inf'), dtype=torch.float64, device=r.device)
    mm = torch.stack([r.min() if r.numel() else big, -r.max() if r.numel() else big])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mm, op=dist.ReduceOp.MIN, group=group)
    total, sq, n = [float(x) for x in s.tolist()]
    mean = total / n
    var = max(0.0, sq / n - mean * mean)
    return {'mean': mean, 'std': math.sqrt(var), 'min': float(mm[0]), 'max': -float(mm[1]), 'count': int(n)}


def gather_median(returns, group=None):
    """Median of the per-env returns over all ranks (BaseAgent.evaluate reports it; unlike the
    other statistics it needs the values themselves: one all-gather of f64[n_local], 32 KB per
    rank at 4096 envs)."""
    r = returns.to(torch.float64).contiguous()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        sizes = [torch.zeros(1, dtype=torch.int64, device=r.device) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([r.numel()], dtype=torch.int64, device=r.device), group=group)
        m = int(max(int(s.item()) for s in sizes))
        pad = torch.full((m,), float('nan'), dtype=torch.float64, device=r.device)
        pad[:r.numel()] = r
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        r = torch.cat([q[:int(s.item())] for q, s in zip(parts, sizes)])
    return float(torch.quantile(r, 0.5).item()) if r.numel() else float('nan')
