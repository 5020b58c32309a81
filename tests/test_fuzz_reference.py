"""Random RDDL domains through the UNMODIFIED reference simulator and through this repo's front end +
oracle + lowering (op-table emulator): pins the oracle -- the checker of every GPU parity test -- to the
reference on programs nobody wrote by hand. Needs the reference checkout (build container only)."""
import os

import numpy as np
import pytest

import cases
from fuzz_specs import random_spec
from oracle import harness
from oracle.oracle_sim import OracleSimulator
from optable_emulator import EmulatedSimulator
from pyrddlgym_b200 import lowering

pytestmark = pytest.mark.skipif(not os.path.isdir('/root/reference/pyRDDLGym'), reason='reference checkout not present')


@pytest.mark.parametrize('seed', range(80))
def test_random_domain_reference_vs_oracle_vs_emulated_optable(seed):
    from oracle.gen_fixtures import reference_program
    spec = random_spec(seed, depth=3 + seed % 3, n=3 + seed % 4, m=2 + (seed // 4) % 3)
    p = reference_program(spec)
    N, T = 3, 3
    ref = harness.ReferenceBatch(spec, N)
    ora = OracleSimulator(p, n_envs=N, seed=1)
    emu = EmulatedSimulator(p, lowering.lower(p), N, seed=1)
    d0 = ref.reset()
    assert np.array_equal(d0, ora.reset()[1]) and np.array_equal(d0, emu.reset())
    rng = np.random.default_rng(seed)
    fluents = [n for n in p.tensors if p.tensors[n].kind not in ('nonfluent', 'next', 'action')]
    for t in range(T):
        acts = harness.random_actions(p, N, rng, p_bool=0.4)
        var = harness.random_variates(p, N, rng)
        r0, dn0 = ref.step(acts, var)
        _, r1, dn1 = ora.step(acts, var)
        r2, dn2 = emu.step(acts, var)
        np.testing.assert_allclose(r1, r0, rtol=cases.RTOL, atol=1e-12)
        np.testing.assert_allclose(r2, r0, rtol=cases.RTOL, atol=1e-12)
        assert np.array_equal(dn0, dn1) and np.array_equal(dn0, dn2)
        for f in fluents:
            want = ref.fluent(f)
            for name, got in (('oracle', ora.subs[f]), ('emulator', emu.fluent(f))):
                got = np.asarray(got).reshape(want.shape)
                if want.dtype.kind == 'f':
                    np.testing.assert_allclose(got, want, rtol=cases.RTOL, atol=1e-12, err_msg=f'{name} {f} step {t}')
                else:
                    assert np.array_equal(got.astype(want.dtype), want), (name, f, t)
        assert np.array_equal(ref.invariants(), ora.check_state_invariants())


@pytest.mark.parametrize('seed', range(400, 424))
def test_random_grid_domain_reference_vs_oracle_vs_plane_optable(seed):
    """Wildfire-like random grid domains: the reference vs the oracle vs the op table lowered to the BIT-PLANE
    path (stencil counts, truth tables, Bernoulli thresholds folded on the host), under injected uniforms."""
    from fuzz_specs import random_grid_spec
    from oracle.gen_fixtures import reference_program
    H, W = [(3, 32), (2, 64), (5, 32), (1, 96)][seed % 4]
    spec = random_grid_spec(seed, H, W)
    p = reference_program(spec)
    tab = lowering.lower(p)
    assert tab.launches[0]['kernel'] == 'rddl_plane_interp', tab.plane_rejects
    N, T = 2, 3
    ref = harness.ReferenceBatch(spec, N)
    ora = OracleSimulator(p, n_envs=N, seed=1)
    emu = EmulatedSimulator(p, tab, N, seed=1)
    ref.reset(); ora.reset(); emu.reset()
    rng = np.random.default_rng(seed)
    for t in range(T):
        acts = harness.random_actions(p, N, rng, p_bool=0.2)
        var = harness.random_variates(p, N, rng)
        r0, d0 = ref.step(acts, var)
        _, r1, d1 = ora.step(acts, var)
        r2, d2 = emu.step(acts, var)
        np.testing.assert_allclose(r1, r0, rtol=1e-12)
        np.testing.assert_allclose(r2, r0, rtol=1e-12)
        assert np.array_equal(d0, d1) and np.array_equal(d0, d2)
        for k in range(3):
            want = ref.fluent(f's{k}')
            assert np.array_equal(np.asarray(ora.subs[f's{k}']).reshape(want.shape), want), (k, t)
            assert np.array_equal(np.asarray(emu.fluent(f's{k}')).reshape(want.shape).astype(bool), want), (k, t)


def test_parameter_errors_behind_scalar_guards_match_the_reference_exceptions():
    """`if (c) then Normal(m, v) else ...` with one value of c per env: the reference short-circuits the `if`
    (simulator.py:912-919), so an invalid v only raises when its branch is taken. Per env: reference raised
    RDDLValueOutOfRangeError <=> status word set (oracle and emulated op table), also under nested guards."""
    from pyrddlgym_b200.workloads import add, gt, ite, lt, mul, num, pv, rand
    from oracle.gen_fixtures import reference_program
    from oracle.ref_import import load_reference
    inner = ite(lt(pv('u'), num(1.0)), rand('Normal', num(0.0), pv('act')), rand('Exponential', pv('act')))
    spec = dict(
        name='guards', types=[('ta', 'object')],
        pvariables=[('f0', 'state-fluent', 'real', None, 0.5), ('act', 'action-fluent', 'real', None, 0.0),
                    ('u', 'action-fluent', 'real', None, 0.0)],
        cpfs=[(("f0'", None), ite(gt(pv('u'), num(0.0)), inner, mul(pv('f0'), num(0.5))))],
        reward=pv('f0'), preconds=[], invariants=[], terminals=[],
        objects=[('ta', ['a1'])], init_non_fluent=[], init_state=[],
        max_nondef_actions='pos-inf', horizon=5, discount=1.0)
    p = reference_program(spec)
    ref_err = load_reference().RDDLValueOutOfRangeError if hasattr(load_reference(), 'RDDLValueOutOfRangeError') else Exception
    #        u: outer guard / which inner branch     act: variance (Normal, >= 0 required) or scale (Exponential, > 0)
    u = np.array([-1.0, 0.5, 0.5, 2.0, 2.0, -3.0])
    act = np.array([-2.0, 1.0, -1.0, 1.0, -1.0, 0.0])
    N = len(u)
    want = np.array([False, False, True, False, True, False])      # raised by the reference
    ora = OracleSimulator(p, n_envs=N, seed=1)
    emu = EmulatedSimulator(p, lowering.lower(p), N, seed=1)
    ora.reset(); emu.reset()
    var = harness.random_variates(p, N, np.random.default_rng(0))
    acts = {'act': act, 'u': u}
    raised = []
    for e in range(N):
        rb = harness.ReferenceBatch(spec, 1)
        rb.reset()
        try:
            rb.step({k: v[e:e + 1] for k, v in acts.items()}, {k: v[e:e + 1] for k, v in var.items()})
            raised.append(False)
        except Exception as ex:
            assert 'OutOfRange' in type(ex).__name__, ex
            raised.append(True)
    assert np.array_equal(np.array(raised), want)
    ora.step(acts, var)
    emu.step(acts, var)
    assert np.array_equal(ora.status != 0, want), ora.status
    assert np.array_equal(emu.status != 0, want), emu.status


def test_parameter_errors_behind_a_scalar_switch_match_the_reference_exceptions():
    """switch on an enum-valued scalar fluent (one value per env): the reference evaluates only the selected
    case when it short-circuits (simulator.py:933-943); per env, raised <=> status word set."""
    from pyrddlgym_b200.workloads import mul, num, pv, rand, switch
    from oracle.gen_fixtures import reference_program
    modes = ['@normal', '@expo', '@idle', '@other']
    body = switch(pv('mode'), [('@normal', rand('Normal', num(0.0), num(1.0))),
                               ('@expo', rand('Exponential', pv('act')))], default=mul(pv('f0'), num(0.5)))
    spec = dict(
        name='switchguard', types=[('ta', 'object'), ('mode_t', modes)],
        pvariables=[('f0', 'state-fluent', 'real', None, 0.5), ('mode', 'state-fluent', 'mode_t', None, '@idle'),
                    ('act', 'action-fluent', 'real', None, 0.0)],
        cpfs=[(("f0'", None), body), (("mode'", None), pv('mode'))],
        reward=pv('f0'), preconds=[], invariants=[], terminals=[],
        objects=[('ta', ['a1'])], init_non_fluent=[], init_state=[],
        max_nondef_actions='pos-inf', horizon=5, discount=1.0)
    act_values = [-1.0, 2.0]
    for mi, mode in enumerate(modes):
        spec_m = dict(spec, init_state=[(('mode', None), mode)])
        p = reference_program(spec_m)
        N = len(act_values)
        acts = {'act': np.array(act_values)}
        var = harness.random_variates(p, N, np.random.default_rng(mi))
        raised = []
        for e in range(N):
            rb = harness.ReferenceBatch(spec_m, 1)
            rb.reset()
            try:
                rb.step({k: v[e:e + 1] for k, v in acts.items()}, {k: v[e:e + 1] for k, v in var.items()})
                raised.append(False)
            except Exception as ex:
                assert 'OutOfRange' in type(ex).__name__, ex
                raised.append(True)
        want = np.array(raised)
        # act = -1 is an invalid Exponential scale. @normal (enum value 0): short-circuit, only the Normal case
        # is evaluated -> no error. @expo (1): short-circuit into the Exponential -> error. @idle / @other (2, 3):
        # the reference short-circuits only for the first two enum values (`first_elem = bool(pred)`), so it
        # evaluates every case -> error.
        assert want.tolist() == [mode != '@normal', False], (mode, raised)
        ora = OracleSimulator(p, n_envs=N, seed=1)
        emu = EmulatedSimulator(p, lowering.lower(p), N, seed=1)
        ora.reset(); emu.reset()
        ora.step(acts, var)
        emu.step(acts, var)
        assert np.array_equal(ora.status != 0, want), (mode, ora.status)
        assert np.array_equal(emu.status != 0, want), (mode, emu.status)
