"""Differential fuzzing: random HIR programs (tests/fuzz_programs.py) through the lowering and
(a) the op-table emulator on CPU, (b) the CUDA interpreter on GPU, against the NumPy oracle, with
injected base variates. Bars: exact for bool/int fluents, 1e-5 relative for reals."""
import numpy as np
import pytest

import cases
from fuzz_programs import random_program
from oracle import harness
from oracle.oracle_sim import OracleSimulator
from pyrddlgym_b200 import lowering, optable as ot


def _drive(p, sim_step, sim_fluent, sim_checks, ora, N, steps, rng):
    for t in range(steps):
        acts = harness.random_actions(p, N, rng)
        var = harness.random_variates(p, N, rng)
        pre, inv_fn = sim_checks
        cases.compare('pre', pre(acts), ora.check_action_preconditions(acts))
        r, d = sim_step(acts, var)
        _, r0, d0 = ora.step(acts, var)
        cases.compare(f'reward/{t}', r, r0, atol=1e-7)
        cases.compare(f'done/{t}', d, d0)
        cases.compare('inv', inv_fn(), ora.check_state_invariants())
        for name, decl in p.tensors.items():
            if decl.kind not in ('nonfluent', 'next', 'action'):
                cases.compare(f'{name}/{t}', sim_fluent(name), ora.subs[name], atol=1e-7)


@pytest.mark.parametrize('seed', range(12))
def test_fuzz_emulated_optable_vs_oracle(seed):
    from optable_emulator import EmulatedSimulator
    p = random_program(seed, depth=3)
    N = 2
    em = EmulatedSimulator(p, lowering.lower(p), N, seed=1)
    ora = OracleSimulator(p, n_envs=N, seed=1)
    _drive(p, em.step, em.fluent,
           (lambda a: em.check(ot.PLAN_PRECONDITION, a), lambda: em.check(ot.PLAN_INVARIANT)),
           ora, N, 2, np.random.default_rng(seed))
    assert np.array_equal(em.status, ora.status)


@pytest.mark.gpu
@pytest.mark.parametrize('seed', range(100, 160))
def test_fuzz_cuda_vs_oracle(seed):
    import torch
    from pyrddlgym_b200.simulator import BatchedSimulator
    p = random_program(seed, depth=4)
    N = 6
    sim = BatchedSimulator(p, n_envs=N, device='cuda:0', seed=1)
    ora = OracleSimulator(p, n_envs=N, seed=1)
    to_np = lambda x: x.detach().cpu().numpy()

    def step(a, v):
        _, r, d = sim.step(a, v)
        return to_np(r), to_np(d)

    _drive(p, step, lambda n: to_np(sim.fluent(n)),
           (lambda a: to_np(sim.check_action_preconditions(a)), lambda: to_np(sim.check_state_invariants())),
           ora, N, 3, np.random.default_rng(seed))
    assert np.array_equal(to_np(sim.status).astype(np.uint32), ora.status)
