"""Batched policies and ``evaluate`` -- the device-side counterpart of
/root/reference/pyRDDLGym/core/policy.py (BaseAgent.evaluate :45-126, RandomAgent :129-178,
NoOpAgent :181-195) for N environments stepped in lock-step.

One env = one episode. ``RandomPolicy`` / ``NoOpPolicy`` are *device-side* policies: they describe
how every action grounding is drawn (``BatchedSimulator.policy_descriptor`` -> rddl_policy_t) and the
rollout kernels draw the actions themselves from Philox counters -- no action tensor is sampled by
PyTorch, stored in HBM or copied over PCIe. ``evaluate`` resets the batch, runs the horizon through
``BatchedSimulator.rollout_policy`` (one fused kernel for single-tile plane programs) and reduces
the per-env discounted returns to the statistics dict of the reference -- across all ranks when
torch.distributed is initialised. Policies that only provide ``sample(sim, n_steps)`` (explicit action
sequences ``[n_steps, N, ...]``) run through ``BatchedSimulator.rollout`` instead.
"""
import torch

from pyrddlgym_b200 import parallel


class NoOpPolicy:
    """NoOpAgent (policy.py:181-195): every action-fluent keeps its default value."""

    def spec(self, sim):
        return {name: ('default',) for name in sim.action_names}

    def descriptor(self, sim):
        return sim.policy_descriptor(self.spec(sim), max_nondef=-1)


class RandomPolicy:
    """RandomAgent (policy.py:129-178) on the device: every action grounding is drawn uniformly from
    its range (bool: Bernoulli(``p_bool``), 1/2 = the agent's Discrete(2); int / real: uniform on
    ``bounds[name] = (lo, hi)``), then only ``max_allowed_actions`` randomly chosen groundings per env
    and step keep their draw and the rest fall back to the default (the reference keeps ``num_actions``
    random keys of the sampled dict, policy.py:166-178; default: the instance's max-nondef-actions).
    ``p_bool`` may be a dict name -> p."""

    def __init__(self, bounds=None, p_bool=0.5, max_allowed_actions=None):
        self.bounds = dict(bounds or {})
        self.p_bool = p_bool
        self.max_allowed_actions = max_allowed_actions

    def spec(self, sim):
        out = {}
        for name in sim.action_names:
            if sim.program.tensors[name].dtype == 'b':
                p = self.p_bool.get(name, 0.5) if isinstance(self.p_bool, dict) else self.p_bool
                out[name] = ('bernoulli', float(p))
            else:
                lo, hi = self.bounds.get(name, (0.0, 1.0))
                out[name] = ('uniform', lo, hi)
        return out

    def descriptor(self, sim):
        return sim.policy_descriptor(self.spec(sim), max_nondef=self.max_allowed_actions)

    def sample(self, sim, n_steps, generator=None):
        """The explicit action sequence ``[n_steps, N, ...]`` this policy takes from the simulator's
        current Philox step on (materialised step by step by rddl_policy_sample; for inspection and
        tests -- ``evaluate`` never builds it)."""
        desc = self.descriptor(sim)
        steps = [{k: v.clone() for k, v in sim.sample_policy_actions(desc, sim.step_count + t).items()}
                 for t in range(n_steps)]
        return {k: torch.stack([s[k] for s in steps]) for k in steps[0]} if steps else {}


def evaluate(sim, policy, n_steps=None, gamma=None, seed=None, group=None, episodes=None, check_status=True):
    """BaseAgent.evaluate (policy.py:45-126) with one episode per env and rollout: returns
    {'mean', 'median', 'min', 'max', 'std'} of the discounted returns over all episodes (and ranks).
    ``episodes`` (per rank; default = n_envs, one rollout): more episodes than envs run as successive
    rollouts of the whole batch -- the reference's ``episodes`` loop, n_envs at a time; like the
    reference's generator stream, the Philox step index keeps advancing across the resets, so no
    rollout repeats another's random numbers. ``seed`` (optional) reseeds the simulator first
    (env.reset(seed=seed) on the first episode, policy.py:80,118). ``check_status``: raise if any env
    flagged an invalid distribution parameter during the rollouts (the reference raises
    RDDLValueOutOfRangeError from inside step)."""
    p = sim.program
    n_steps = int(p.horizon if n_steps is None else n_steps)
    episodes = sim.n_envs if episodes is None else int(episodes)
    if episodes < 1:
        raise ValueError('episodes must be >= 1')
    if seed is not None:
        sim.seed(int(seed))
    desc = policy.descriptor(sim) if hasattr(policy, 'descriptor') else None
    chunks, status = [], None
    while sum(int(c.numel()) for c in chunks) < episodes:
        sim.reset()
        if desc is not None:
            returns, _ = sim.rollout_policy(desc, n_steps, gamma=gamma)
        else:
            returns, _ = sim.rollout(policy.sample(sim, n_steps), n_steps, gamma=gamma)
        chunks.append(returns.clone())
        status = sim.status.clone() if status is None else status | sim.status     # (reset clears the live word)
    if check_status:
        sim.status.copy_(status)
        sim.raise_for_status()
    returns = torch.cat(chunks)[:episodes]
    stats = parallel.reduce_return_stats(returns, group=group)
    stats['median'] = parallel.gather_median(returns, group=group)
    return {k: stats[k] for k in ('mean', 'median', 'min', 'max', 'std')}
