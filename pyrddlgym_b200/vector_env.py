"""Vector-environment adapter over ``BatchedSimulator``: the batched counterpart of
/root/reference/pyRDDLGym/core/env.py (RDDLEnv.reset :282-310, RDDLEnv.step :226-280) with the
calling convention of ``gymnasium.vector.VectorEnv`` (duck-typed: gymnasium is not a dependency).

N environments advance in lock-step on the device. ``step`` returns ``(obs, reward, terminated,
truncated, info)`` with tensors of leading dimension N: ``terminated`` is the RDDL termination of
the new state (env.py:262-264), ``truncated`` the horizon (env.py:265-266: an episode ends after
``horizon`` steps, and env.py:260: it is also truncated when a state invariant fails). ``info['status']``
is the per-env error word of the step (invalid distribution parameters, clamped object indices: the
reference raises RDDLValueOutOfRangeError; ``raise_on_status=True`` does the same here at the cost of
one device synchronisation per step). With ``auto_reset`` (default) the environments whose episode ended are reset in
place after the step -- only their rows of the state tensors are re-initialised, the others keep
running -- and ``info['episode_return']`` / ``info['episode_length']`` carry the finished episodes'
discounted return and length (NaN / 0 elsewhere). Observations are the state fluents (MDP) or the
observ-fluents (POMDP, env.py:270-273).

``action_bounds`` / ``observation_bounds``: name -> (lower, upper) device tensors of the fluent's shape,
the box bounds RDDLEnv derives for its gym spaces from the preconditions and invariants
(RDDLConstraints, constraints.py:107-229; env.py:152-193) -- computed once on the host
(pyrddlgym_b200/constraints.py) and kept on the device; ``clip_actions=True`` clamps real / int actions
into them before every step (what a Box space does for an RL library), ``random_policy()`` is the
device-side RandomAgent over exactly these bounds.
"""
import numpy as np
import torch

from pyrddlgym_b200.constraints import BoxBounds
from pyrddlgym_b200.simulator import BatchedSimulator


class BatchedRDDLEnv:
    def __init__(self, program, n_envs, device='cuda:0', seed=0, horizon=None, auto_reset=True,
                 raise_on_status=False, clip_actions=False, max_bound=float('inf'), **sim_kwargs):
        self.sim = program if isinstance(program, BatchedSimulator) else BatchedSimulator(
            program, n_envs=n_envs, device=device, seed=seed, **sim_kwargs)
        self.program = self.sim.program
        self.num_envs = self.sim.n_envs
        self.device = self.sim.device
        self.horizon = int(self.program.horizon if horizon is None else horizon)
        self.discount = float(self.program.discount)
        self.auto_reset = bool(auto_reset)
        self.raise_on_status = bool(raise_on_status)
        self.clip_actions = bool(clip_actions)
        self.box = BoxBounds(self.program, max_bound=max_bound)
        dev_bounds = self.box.to_device(self.device)
        observ = self.program.names_of_kind('observ')
        self.observation_names = list(observ) if observ else list(self.sim.state_names)
        self.action_names = list(self.sim.action_names)
        self.action_bounds = {n: dev_bounds[n] for n in self.action_names}
        self.observation_bounds = {n: dev_bounds[n] for n in self.observation_names if n in dev_bounds}
        N, dev = self.num_envs, self.device
        self.episode_step = torch.zeros(N, dtype=torch.int64, device=dev)
        self.episode_return = torch.zeros(N, dtype=torch.float64, device=dev)
        self._disc = torch.ones(N, dtype=torch.float64, device=dev)

    # ---------------------------------------------------------------------------------
    def _obs(self):
        return {name: self.sim.fluent(name) for name in self.observation_names}

    def reset(self, seed=None):
        """All environments back to the initial state -> (obs, info)."""
        if seed is not None:
            self.sim.seed(seed)
        self.sim.reset()
        self.episode_step.zero_()
        self.episode_return.zero_()
        self._disc.fill_(1.0)
        return self._obs(), {}

    def step(self, actions):
        """actions: name -> tensor / array [N, *shape] (missing action-fluents keep their defaults,
        like RDDLEnv.step's partial action dicts)."""
        full = {name: self.sim._init[name].expand((self.num_envs,) + tuple(self.sim._init[name].shape))
                for name in self.action_names}
        full.update(actions or {})
        if self.clip_actions:
            for name, v in list(full.items()):
                if self.program.tensors[name].dtype != 'b':
                    lo, hi = self.action_bounds[name]
                    v = v if isinstance(v, torch.Tensor) else torch.as_tensor(v)
                    full[name] = torch.clamp(v.to(self.device), min=lo.to(v.dtype), max=hi.to(v.dtype))
        _, reward, terminated = self.sim.step(full)
        reward = reward.clone()
        terminated = terminated.to(torch.bool).clone()
        self.episode_return += reward * self._disc
        self._disc *= self.discount
        self.episode_step += 1
        truncated = (self.episode_step >= self.horizon) & ~terminated
        if self.sim.table.n_checks['invariant']:
            truncated |= ~self.sim.check_state_invariants() & ~terminated       # env.py:260
        if self.raise_on_status:
            self.sim.raise_for_status()
        ended = terminated | truncated
        info = {'status': self.sim.status.clone(), 'episode_return': torch.where(ended, self.episode_return, torch.full_like(self.episode_return, float('nan'))),
                'episode_length': torch.where(ended, self.episode_step, torch.zeros_like(self.episode_step))}
        if self.auto_reset:
            self.sim.reset_envs(ended)
            keep = (~ended).to(torch.float64)
            self.episode_return *= keep
            self._disc = torch.where(ended, torch.ones_like(self._disc), self._disc)
            self.episode_step *= (~ended).to(torch.int64)
        return self._obs(), reward, terminated, truncated, info

    def random_policy(self, p_bool=0.5, max_allowed_actions=None):
        """RandomAgent over the box bounds (policy.py:129-178 samples the gym action space RDDLEnv builds
        from them): a device-side ``RandomPolicy``. Needs finite bounds that are the same for every
        grounding of a fluent (the device policy takes one range per fluent)."""
        from pyrddlgym_b200.policy import RandomPolicy
        bounds = {}
        for name in self.action_names:
            if self.program.tensors[name].dtype == 'b':
                continue
            lo, hi = self.box.bounds[name]
            if not (np.isfinite(lo).all() and np.isfinite(hi).all()) or np.ptp(lo) != 0 or np.ptp(hi) != 0:
                raise ValueError(f'<{name}>: bounds are not finite and uniform over its groundings')
            bounds[name] = (float(np.asarray(lo).reshape(-1)[0]), float(np.asarray(hi).reshape(-1)[0]))
        return RandomPolicy(bounds=bounds, p_bool=p_bool, max_allowed_actions=max_allowed_actions)

    def close(self):
        self.sim.close()
