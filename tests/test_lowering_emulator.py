"""Lowering + op-table semantics checked on CPU: the scalar Python emulator of the bytecode
replays the reference golden trajectories (the CUDA interpreter is checked by the gpu tests)."""
import numpy as np
import pytest

import cases
from optable_emulator import EmulatedSimulator
from pyrddlgym_b200 import lowering, optable as ot


class _Adapter(EmulatedSimulator):
    def __init__(self, p, n):
        super().__init__(p, lowering.lower(p), n)

    def reset(self):
        return None, super().reset()

    def step(self, acts, var):
        r, d = super().step(acts, var)
        return None, r, d

    def check_state_invariants(self):
        return self.check(ot.PLAN_INVARIANT)

    def check_action_preconditions(self, acts):
        return self.check(ot.PLAN_PRECONDITION, acts)


@pytest.mark.parametrize('case,steps', [('cartpole', 12), ('opzoo', 5), ('wildfire8', 4), ('wildfire8_sep', 4),
                                        ('wildfire32', 1), ('reservoir30', 3), ('reservoir30_exp', 2),
                                        ('pair16', 3)])
def test_emulated_optable_matches_reference_golden(case, steps):
    sim = cases.replay(case, lambda p, n: _Adapter(p, n), steps=steps)
    assert not sim.status.any()


@pytest.mark.parametrize('case', cases.FUZZSPEC_CASES)
def test_emulated_optable_matches_reference_on_random_domains(case):
    cases.replay(case, lambda p, n: _Adapter(p, n))


def test_wildfire_neighbour_mask_becomes_csr():
    from pyrddlgym_b200 import workloads
    for sep in (False, True):
        p = workloads.load_program(workloads.wildfire(6, separable=sep))
        tab = lowering.lower(p)
        step = [l for l in tab.launches if l['plan'] == ot.PLAN_STEP]
        # burning', out-of-fuel' and the four reward sums fuse into one element launch
        assert step[0]['scope'] == ['x_pos', 'y_pos'] and step[0]['nred'] == 4
        assert len(step) == 2
        em = EmulatedSimulator(p, tab, 1)
        names = [ot._OPS[int(i['op'])] for i in em.em.insns]
        assert 'CSRLOOP' in names and 'LOOP' not in names


def test_native_samplers_lower_to_sample_ops_and_run_in_the_emulator():
    from pyrddlgym_b200 import workloads
    p = workloads.load_program(workloads.distzoo(3))
    tab = lowering.lower(p)
    em = EmulatedSimulator(p, tab, 40, seed=3)
    names = [ot._OPS[int(i['op'])] for i in em.em.insns]
    assert names.count('SAMPLE') == len(workloads.DISTZOO_PARAMS)
    em.step({})
    assert not em.status.any()
    x = em.fluent('poisson-large').astype(np.float64)
    assert abs(x.mean() - 47.0) < 4.0
    assert (em.fluent('geometric') >= 1).all() and (em.fluent('beta') > 0).all() and (em.fluent('beta') < 1).all()
