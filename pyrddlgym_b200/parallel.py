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


def local_return_stats(returns):
    """f64[5] = {sum, sum of squares, count, min, max} of this rank's per-env returns, on the tensor's
    device. CUDA tensors: one launch of the library's block reduction (rddl_return_stats); CPU tensors
    (the gloo tests of this plumbing) use torch."""
    r = returns.to(torch.float64).contiguous()
    if r.is_cuda:
        from . import native
        out = torch.empty(5, dtype=torch.float64, device=r.device)
        native.return_stats(r, out, torch.cuda.current_stream(r.device).cuda_stream)
        return out
    inf = float('inf')
    return torch.tensor([float(r.sum()), float((r * r).sum()), float(r.numel()),
                         float(r.min()) if r.numel() else inf, float(r.max()) if r.numel() else -inf],
                        dtype=torch.float64)


def return_stats_async(returns, group=None):
    """Per-rank statistics of every rank, f64[world, 5], without synchronising the host: the one
    collective of the path (an all-gather of five numbers per rank)."""
    return gather_return_stats(local_return_stats(returns), group)


def gather_return_stats(local, group=None):
    """All-gather of the ranks' f64[5] local statistics -> f64[world, 5]."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        out = torch.empty(world * 5, dtype=torch.float64, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out.view(world, 5)
    return local[None]


def finish_return_stats(gathered):
    """The dict of BaseAgent.evaluate (mean, population std, min, max) + count from return_stats_async's
    tensor (this is where the host waits for the device)."""
    g = gathered.cpu().tolist()
    total = sum(x[0] for x in g)
    sq = sum(x[1] for x in g)
    n = sum(x[2] for x in g)
    mean = total / n
    var = max(0.0, sq / n - mean * mean)
    return {'mean': mean, 'std': math.sqrt(var), 'min': min(x[3] for x in g), 'max': max(x[4] for x in g),
            'count': int(n)}


def reduce_return_stats(returns, group=None):
    """returns: f64[n_local] per-env (discounted) returns of this rank's shard. Gives the dict of
    BaseAgent.evaluate -- mean, std (population), min, max -- over ALL ranks, plus the count."""
    return finish_return_stats(return_stats_async(returns, group))


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
