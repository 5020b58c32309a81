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
    # every third program runs through the run-time specialised scalar kernels from its first launch
    sim = BatchedSimulator(p, n_envs=N, device='cuda:0', seed=1, specialize=True if seed % 3 == 0 else None)
    ora = OracleSimulator(p, n_envs=N, seed=1)
    to_np = lambda x: x.detach().cpu().numpy()

    def step(a, v):
        _, r, d = sim.step(a, v)
        return to_np(r), to_np(d)

    _drive(p, step, lambda n: to_np(sim.fluent(n)),
           (lambda a: to_np(sim.check_action_preconditions(a)), lambda: to_np(sim.check_state_invariants())),
           ora, N, 3, np.random.default_rng(seed))
    assert np.array_equal(to_np(sim.status).astype(np.uint32), ora.status)
    if seed % 3 == 0:
        assert sim.specialised_kernels >= 1, sim.native.jit_error()


def _grid_drive(p, table, make, N, steps, seed, inject):
    ora = OracleSimulator(p, n_envs=N, seed=3, bitsliced_nodes=table.plane_nodes)
    sim = make(p, table, N)
    rng = np.random.default_rng(seed)
    for t in range(steps):
        acts = harness.random_actions(p, N, rng, p_bool=0.2)
        var = harness.random_variates(p, N, rng) if inject else None
        r, d = sim.step(acts, var)
        _, r0, d0 = ora.step(acts, var)
        for k in range(3):
            assert np.array_equal(np.asarray(sim.fluent(f's{k}')).astype(bool), ora.subs[f's{k}']), (f's{k}', t)
        np.testing.assert_allclose(r, r0, rtol=1e-9, atol=1e-9)
    return sim


@pytest.mark.parametrize('seed', range(16))
def test_fuzz_grid_programs_take_the_plane_path_and_match_the_oracle(seed):
    from fuzz_programs import random_grid_program
    from optable_emulator import EmulatedSimulator
    p = random_grid_program(seed)
    tab = lowering.lower(p)
    kinds = [l['kernel'] for l in tab.launches if l['plan'] == ot.PLAN_STEP]
    assert kinds[0] == 'rddl_plane_interp', (kinds, tab.plane_rejects)
    for inject in (True, False):
        _grid_drive(p, tab, lambda p_, t_, n: EmulatedSimulator(p_, t_, n, seed=3), 2, 2, seed, inject)


@pytest.mark.gpu
@pytest.mark.parametrize('seed', range(200, 240))
def test_fuzz_grid_cuda_plane_vs_scalar_vs_oracle(seed):
    import torch
    from fuzz_programs import random_grid_program
    from pyrddlgym_b200.simulator import BatchedSimulator
    H, W = [(4, 32), (3, 64), (33, 32), (5, 96)][seed % 4]
    p = random_grid_program(seed, H=H, W=W)
    N = 37

    class _S:
        def __init__(self, sim):
            self.sim = sim

        def step(self, a, v):
            _, r, d = self.sim.step(a, v)
            return r.cpu().numpy(), d.cpu().numpy()

        def fluent(self, n):
            return self.sim.fluent(n).cpu().numpy()

    tab = lowering.lower(p)
    assert tab.launches[0]['kernel'] == 'rddl_plane_interp', tab.plane_rejects
    for inject in (True, False):
        _grid_drive(p, tab, lambda p_, t_, n: _S(BatchedSimulator(p_, n_envs=n, seed=3, optable=t_)), N, 3, seed, inject)
    # scalar interpreter on the same program (injected uniforms): bit-identical planes
    a = BatchedSimulator(p, n_envs=N, seed=3)
    b = BatchedSimulator(p, n_envs=N, seed=3, use_planes=False)
    rng = np.random.default_rng(seed)
    for t in range(3):
        acts = harness.random_actions(p, N, rng, p_bool=0.2)
        var = harness.random_variates(p, N, rng)
        _, ra, _ = a.step(acts, var)
        _, rb, _ = b.step(acts, var)
        for k in range(3):
            assert torch.equal(a.fluent(f's{k}'), b.fluent(f's{k}'))
        np.testing.assert_allclose(ra.cpu().numpy(), rb.cpu().numpy(), rtol=1e-9)
    # specialised (NVRTC) vs interpreter build of the plane kernel, native Philox: bit-identical
    c = BatchedSimulator(p, n_envs=N, seed=3)
    d = BatchedSimulator(p, n_envs=N, seed=3, specialize=False)
    for t in range(3):
        acts = harness.random_actions(p, N, rng, p_bool=0.2)
        _, rc, dc = c.step(acts)
        _, rd, dd = d.step(acts)
        for k in range(3):
            assert torch.equal(c.fluent(f's{k}'), d.fluent(f's{k}'))
        assert torch.equal(rc, rd) and torch.equal(dc, dd)
    assert c.specialised_kernels >= 1 and d.specialised_kernels == 0, c.native.jit_error()
    # bit-packed action input (packed_actions=True) on the same program: identical transitions, whichever build
    if seed % 2 == 0:
        e = BatchedSimulator(p, n_envs=N, seed=3, packed_actions=True, specialize=(None if seed % 4 == 0 else False))
        f = BatchedSimulator(p, n_envs=N, seed=3)
        for t in range(3):
            acts = harness.random_actions(p, N, rng, p_bool=0.2)
            _, re_, de = e.step(acts)
            _, rf, df = f.step(acts)
            for k in range(3):
                assert torch.equal(e.fluent(f's{k}'), f.fluent(f's{k}'))
            assert torch.equal(re_, rf) and torch.equal(de, df)
            assert torch.equal(e.count_nondefault_actions(), f.count_nondefault_actions())
