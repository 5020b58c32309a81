"""Box bounds from preconditions / invariants (pyrddlgym_b200/constraints.py) against the reference's
RDDLConstraints (/root/reference/pyRDDLGym/core/constraints.py:107-229) on the same models."""
import numpy as np
import pytest

from oracle.ref_import import reference_available
from pyrddlgym_b200 import workloads
from pyrddlgym_b200.constraints import BoxBounds


@pytest.mark.skipif(not reference_available(), reason='reference pyRDDLGym not present')
@pytest.mark.parametrize('name,make', [
    ('cartpole', lambda: workloads.cartpole()),
    ('reservoir12', lambda: workloads.reservoir(12)),
    ('pair8', lambda: workloads.pair_contraction(8)),
    ('opzoo', lambda: workloads.opzoo()),
    ('wildfire6', lambda: workloads.wildfire(6)),
])
def test_box_bounds_equal_the_reference_constraints(name, make):
    from oracle import ast_builders
    from oracle.ref_import import load_reference
    from pyrddlgym_b200 import frontend
    ref = load_reference()
    from pyRDDLGym.core.constraints import RDDLConstraints
    sim = ref.RDDLSimulator(ast_builders.build_model(make()), rng=np.random.default_rng(0), keep_tensors=True)
    want = RDDLConstraints(sim, vectorized=True)
    got = BoxBounds(frontend.program_from_simulator(sim))
    assert set(got.bounds) == set(want.bounds)
    for k, (lo, hi) in want.bounds.items():
        np.testing.assert_array_equal(np.asarray(got.bounds[k][0]), np.asarray(lo, dtype=np.float64), err_msg=k)
        np.testing.assert_array_equal(np.asarray(got.bounds[k][1]), np.asarray(hi, dtype=np.float64), err_msg=k)
    assert got.is_box_preconditions == list(want.is_box_preconditions)
    assert got.is_box_invariants == list(want.is_box_invariants)


def test_box_bounds_without_the_reference():
    p = workloads.load_program(workloads.cartpole())
    b = BoxBounds(p)
    assert b.bounds['force'] == (-10.0, 10.0) and b.is_box_preconditions == [True, True]
    assert b.bounds['pos'][0] == -np.inf and b.bounds['pos'][1] == np.inf
    r = BoxBounds(workloads.load_program(workloads.reservoir(9)), max_bound=1e6)
    assert r.bounds['release'][0].tolist() == [0.0] * 9 and r.bounds['release'][1].tolist() == [1e6] * 9
    assert r.bounds['rlevel'][0].tolist() == [0.0] * 9


@pytest.mark.gpu
def test_vector_env_bounds_on_device_clip_and_random_policy():
    import torch
    from pyrddlgym_b200.policy import evaluate
    from pyrddlgym_b200.vector_env import BatchedRDDLEnv
    p = workloads.load_program(workloads.cartpole())
    env = BatchedRDDLEnv(p, 64, clip_actions=True)
    lo, hi = env.action_bounds['force']
    assert lo.is_cuda and float(lo) == -10.0 and float(hi) == 10.0
    env.reset()
    force = torch.linspace(-30, 30, 64, dtype=torch.float64, device='cuda')
    env.step({'force': force})
    assert float(env.sim.t['force'].abs().max()) == 10.0            # clamped into the precondition box on the device
    assert bool(env.sim.check_action_preconditions().all())
    pol = env.random_policy()
    assert pol.bounds == {'force': (-10.0, 10.0)}
    st = evaluate(env.sim, pol, seed=3)
    assert 5.0 < st['mean'] < 80.0
