"""TEST INFRASTRUCTURE ONLY -- an engine with BatchedSimulator's interface that executes the op
table with the Python emulator, so that the host-side logic of the drop-in subclass
(pyrddlgym_b200/backend.py) can be exercised where the reference is importable but no GPU
exists. The product never constructs it (backend.py defaults to the CUDA BatchedSimulator)."""
import numpy as np

from optable_emulator import EmulatedSimulator
from pyrddlgym_b200 import lowering, optable as ot


class EmulatorEngine:

    def __init__(self, program, n_envs, device=None, seed=0):
        self.program = program
        self.table = lowering.lower(program)
        self.n_envs = n_envs
        self.sim = EmulatedSimulator(program, self.table, n_envs, seed=seed)
        self.state_names = list(program.next_state)

    def seed(self, seed):
        self.sim.seed = int(seed)
        self.sim.step_count = 0

    @property
    def states(self):
        return {s: self.sim.fluent(s) for s in self.state_names}

    def fluent(self, name):
        return self.sim.fluent(name)

    def reset(self):
        done = self.sim.reset()
        return self.states, done

    def step(self, actions=None, variates=None):
        r, d = self.sim.step(actions or {}, variates)
        return self.states, r, d

    def check_state_invariants(self):
        if not self.table.n_checks['invariant']:
            return np.ones(self.n_envs, dtype=bool)
        return self.sim.check(ot.PLAN_INVARIANT)

    def check_action_preconditions(self, actions=None):
        if not self.table.n_checks['precondition']:
            return np.ones(self.n_envs, dtype=bool)
        return self.sim.check(ot.PLAN_PRECONDITION, actions)

    def check_each(self, plan):
        key = 'invariant' if plan == ot.PLAN_INVARIANT else 'precondition'
        return self.sim.fluent('__chk')[:, :self.table.n_checks[key]].astype(bool)

    def check_terminal_states(self):
        self.sim.em.run(ot.PLAN_TERMINAL, self.sim.bufs, self.n_envs, self.sim.status)
        return self.sim.fluent('__done').reshape(-1).astype(bool)

    def raise_for_status(self):
        if self.sim.status.any():
            from pyrddlgym_b200.simulator import StatusError
            raise StatusError('distribution parameter out of range', self.sim.status)
