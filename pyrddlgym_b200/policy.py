"""Batched policies and ``evaluate`` -- the device-side counterpart of
/root/reference/pyRDDLGym/core/policy.py (BaseAgent.evaluate :45-126, RandomAgent :129-178,
NoOpAgent :181-195) for N environments stepped in lock-step.

One env = one episode: ``evaluate`` resets the batch, draws the whole action sequence
``[horizon, N, ...]`` on the device, runs it through ``BatchedSimulator.rollout`` (one fused
kernel for single-tile plane programs) and reduces the per-env discounted returns to the
statistics dict of the reference -- across all ranks when torch.distributed is initialised.
"""
import torch

from pyrddlgym_b200 import parallel


class NoOpPolicy:
    """NoOpAgent (policy.py:181-195): every action-fluent keeps its default value."""

    def sample(self, sim, n_steps, generator=None):
        return {}


class RandomPolicy:
    """RandomAgent (policy.py:129-178) for tensor actions: every action grounding is drawn
    uniformly from its range (bool: {0, 1}; int / real: ``bounds[name] = (lo, hi)``), then only
    ``max_allowed_actions`` randomly chosen groundings per env and step keep their draw, the
    rest fall back to the default (the reference keeps ``num_actions`` random keys of the
    sampled dict, policy.py:166-178). ``p_bool`` overrides the Bernoulli(1/2) of bool actions."""

    def __init__(self, bounds=None, p_bool=0.5, max_allowed_actions=None):
        self.bounds = dict(bounds or {})
        self.p_bool = float(p_bool)
        self.max_allowed_actions = max_allowed_actions

    def sample(self, sim, n_steps, generator=None):
        p = sim.program
        N, dev = sim.n_envs, sim.device
        out, sizes = {}, []
        for name in sim.action_names:
            shape = (n_steps, N) + tuple(p.tensor_shape(name))
            dt = p.tensors[name].dtype
            if dt == 'b':
                out[name] = torch.rand(shape, device=dev, generator=generator) < self.p_bool
            else:
                lo, hi = self.bounds.get(name, (0.0, 1.0))
                u = torch.rand(shape, device=dev, generator=generator, dtype=torch.float64)
                if dt == 'i':
                    out[name] = (torch.floor(u * (hi - lo + 1)) + lo).to(torch.int64)
                else:
                    out[name] = u * (hi - lo) + lo
            sizes.append(int(out[name][0, 0].numel()))
        total = sum(sizes)
        k = p.max_allowed_actions if self.max_allowed_actions is None else self.max_allowed_actions
        if k < total:
            # keep k random groundings per (step, env): rank of i.i.d. scores
            score = torch.rand((n_steps, N, total), device=dev, generator=generator)
            kth = torch.topk(score, k, dim=-1).values[..., -1:]
            keep = score >= kth
            off = 0
            for name, sz in zip(sim.action_names, sizes):
                m = keep[..., off:off + sz].reshape(out[name].shape)
                default = sim._init[name].to(out[name].dtype)
                out[name] = torch.where(m, out[name], default.expand_as(out[name]))
                off += sz
        return out


def evaluate(sim, policy, n_steps=None, gamma=None, seed=None, group=None, episodes=None):
    """BaseAgent.evaluate (policy.py:45-126) with one episode per env and rollout: returns
    {'mean', 'median', 'min', 'max', 'std'} of the discounted returns over all episodes (and ranks).
    ``episodes`` (per rank; default = n_envs, one rollout): more episodes than envs run as successive
    rollouts of the whole batch -- the reference's ``episodes`` loop, n_envs at a time."""
    p = sim.program
    n_steps = int(p.horizon if n_steps is None else n_steps)
    episodes = sim.n_envs if episodes is None else int(episodes)
    if episodes < 1:
        raise ValueError('episodes must be >= 1')
    gen = None
    if seed is not None:
        gen = torch.Generator(device=sim.device).manual_seed(int(seed) + sim.env_offset)
    chunks, seed0, k = [], sim.seed_value, 0
    while sum(int(c.numel()) for c in chunks) < episodes:
        # reset() restarts the Philox step index, so every further rollout of the same envs runs under a
        # seed of its own (restored afterwards): no rollout repeats another's random numbers
        sim.seed((seed0 + k * 0x9E3779B97F4A7C15) & 0x7FFFFFFFFFFFFFFF)
        sim.reset()
        actions = policy.sample(sim, n_steps, generator=gen)
        returns, _ = sim.rollout(actions, n_steps, gamma=gamma)
        chunks.append(returns.clone())
        k += 1
    sim.seed(seed0)
    returns = torch.cat(chunks)[:episodes]
    stats = parallel.reduce_return_stats(returns, group=group)
    stats['median'] = parallel.gather_median(returns, group=group)
    return {k: stats[k] for k in ('mean', 'median', 'min', 'max', 'std')}
