"""Shared helpers for the parity tests: golden-case registry, golden replay driver."""
import os

import numpy as np

from pyrddlgym_b200 import workloads

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# must match oracle/gen_fixtures.py CASES (spec factories only)
SPECS = {
    'cartpole': lambda: workloads.cartpole(),
    'opzoo': lambda: workloads.opzoo(),
    'wildfire8': lambda: workloads.wildfire(8),
    'wildfire8_sep': lambda: workloads.wildfire(8, separable=True),
    'wildfire32': lambda: workloads.wildfire(32),
    'reservoir30': lambda: workloads.reservoir(30),
    'reservoir30_exp': lambda: workloads.reservoir(30, rain='exponential'),
    'pair16': lambda: workloads.pair_contraction(16),
    'reservoir1000': lambda: workloads.reservoir(1000),
}

# float tolerance of the north star: 1e-5 relative (bool / int are compared exactly)
RTOL = 1e-5
ATOL = 1e-9

FUZZSPEC_CASES = sorted(f[:-len('.program.npz')] for f in os.listdir(os.path.join(GOLDEN_DIR, 'fuzzspec'))
                        if f.endswith('.program.npz')) if os.path.isdir(os.path.join(GOLDEN_DIR, 'fuzzspec')) else []


def load_golden(case):
    with np.load(os.path.join(GOLDEN_DIR, case + '.npz')) as z:
        return {k: z[k] for k in z.files}


def golden_step_inputs(g, t):
    acts = {k.split('/', 2)[2]: v for k, v in g.items() if k.startswith(f'act/{t}/')}
    var = {int(k.split('/', 2)[2]): v for k, v in g.items() if k.startswith(f'var/{t}/')}
    return acts, var


def golden_fluents(g, t):
    return {k.split('/', 2)[2]: v for k, v in g.items() if k.startswith(f'flu/{t}/')}


def compare(name, got, want, rtol=RTOL, atol=ATOL):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f'{name}: shape {got.shape} vs {want.shape}'
    if want.dtype.kind in 'biu':
        # discrete fluents: bit exact (a real-typed buffer holding an int-valued CPF, see
        # SURVEY hard part 3, is compared by value)
        bad = got.astype(np.float64) != want.astype(np.float64)
        assert not bad.any(), f'{name}: {int(bad.sum())} exact mismatches'
    else:
        np.testing.assert_allclose(got.astype(np.float64), want.astype(np.float64), rtol=rtol, atol=atol,
                                   err_msg=name)


def replay(case, make_sim, to_np=np.asarray, steps=None):
    """Replays a golden trajectory through ``make_sim(program, n_envs)``; the simulator must
    offer reset() -> (.., done), step(actions, variates) -> (.., reward, done),
    check_state_invariants(), check_action_preconditions(actions), fluent(name)."""
    if case.startswith(('fuzzspec', 'gridspec')):
        # random domains (tests/fuzz_specs.py): program + trajectory generated from the reference by
        # oracle/gen_fuzz_fixtures.py
        from pyrddlgym_b200 import hir
        p = hir.Program.load(os.path.join(GOLDEN_DIR, 'fuzzspec', case + '.program.npz'))
        with np.load(os.path.join(GOLDEN_DIR, 'fuzzspec', case + '.npz')) as z:
            g = {k: z[k] for k in z.files}
    else:
        g = load_golden(case)
        p = workloads.load_program(SPECS[case]())
    N, T = int(g['n_envs']), int(g['steps'])
    sim = make_sim(p, N)
    _, done0 = sim.reset()
    compare('done0', to_np(done0), g['done0'])
    for t in range(T if steps is None else min(T, steps)):
        acts, var = golden_step_inputs(g, t)
        if f'pre/{t}' in g:
            compare(f'pre/{t}', to_np(sim.check_action_preconditions(acts)), g[f'pre/{t}'])
        _, reward, done = sim.step(acts, var)
        compare(f'reward/{t}', to_np(reward), g[f'reward/{t}'])
        compare(f'done/{t}', to_np(done), g[f'done/{t}'])
        if f'inv/{t}' in g:
            compare(f'inv/{t}', to_np(sim.check_state_invariants()), g[f'inv/{t}'])
        for name, want in golden_fluents(g, t).items():
            compare(f'{name}/{t}', to_np(sim.fluent(name)), want)
    return sim
