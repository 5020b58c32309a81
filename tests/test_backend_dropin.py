"""The drop-in subclass B200RDDLSimulator against the reference RDDLSimulator on the same
model and actions (needs the reference, i.e. runs in the build container; the kernels are
replaced by the op-table emulator through the engine seam -- the CUDA engine is covered by
the gpu tests)."""
import numpy as np
import pytest

from oracle.ref_import import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason='reference pyRDDLGym not present')


def _pair(spec_fn, **kw):
    from oracle import ast_builders
    from oracle.ref_import import load_reference
    from pyrddlgym_b200.backend import backend_class
    from fake_engine import EmulatorEngine
    ref = load_reference()
    a = ref.RDDLSimulator(ast_builders.build_model(spec_fn()), rng=np.random.default_rng(0), **kw)
    b = backend_class()(ast_builders.build_model(spec_fn()), n_envs=1, engine_factory=EmulatorEngine, **kw)
    return ref, a, b


def _same(x, y):
    assert type(x) is type(y) or (np.isscalar(x) and np.isscalar(y)), (type(x), type(y))
    np.testing.assert_allclose(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('keep_tensors', [False, True])
def test_cartpole_dropin_matches_reference(keep_tensors):
    from pyrddlgym_b200 import workloads
    ref, a, b = _pair(workloads.cartpole, keep_tensors=keep_tensors)
    assert isinstance(b, ref.RDDLSimulator)
    oa, da = a.reset()
    ob, db = b.reset()
    assert da == db and set(oa) == set(ob)
    rng = np.random.default_rng(3)
    for t in range(60):
        act = {'force': float(rng.uniform(-10, 10))}
        sa, sb = a.prepare_actions_for_sim(act), b.prepare_actions_for_sim(act)
        a.check_default_action_count(sa)
        b.check_default_action_count(sb)
        assert a.check_action_preconditions(sa, silent=True) == b.check_action_preconditions(sb, silent=True)
        oa, ra, da = a.step(sa)
        ob, rb, db = b.step(sb)
        assert isinstance(rb, float) and isinstance(db, bool)
        assert ra == rb and da == db
        assert set(oa) == set(ob)
        for k in oa:
            _same(oa[k], ob[k])
        assert a.check_state_invariants(silent=True) == b.check_state_invariants(silent=True)
        assert set(a.states) == set(b.states)
        if da:
            break
    assert da        # the pole falls within 60 random pushes


def test_precondition_violation_raises_reference_exception():
    from pyrddlgym_b200 import workloads
    ref, a, b = _pair(workloads.cartpole, keep_tensors=True)
    a.reset()
    b.reset()
    bad = b.prepare_actions_for_sim({'force': 11.0})
    assert b.check_action_preconditions(bad, silent=True) is False
    with pytest.raises(ref.exc.RDDLActionPreconditionNotSatisfiedError):
        b.check_action_preconditions(bad, silent=False)
    with pytest.raises(ref.exc.RDDLActionPreconditionNotSatisfiedError):
        a.check_action_preconditions(a.prepare_actions_for_sim({'force': 11.0}), silent=False)


def test_pair_contraction_invariants_and_grounded_state():
    from pyrddlgym_b200 import workloads
    ref, a, b = _pair(lambda: workloads.pair_contraction(12), keep_tensors=False)
    a.reset()
    b.reset()
    rng = np.random.default_rng(1)
    for t in range(5):
        u = rng.uniform(-3, 3, size=12)
        act = {f'u___n{i + 1}': float(u[i]) for i in range(12)}
        sa, sb = a.prepare_actions_for_sim(act), b.prepare_actions_for_sim(act)
        oa, ra, da = a.step(sa)
        ob, rb, db = b.step(sb)
        assert set(oa) == set(ob) and 'x___n1' in ob
        for k in oa:
            _same(oa[k], ob[k])
        np.testing.assert_allclose(ra, rb, rtol=1e-9)
        assert a.check_state_invariants(silent=True) == b.check_state_invariants(silent=True)


def test_invalid_distribution_parameter_maps_to_value_out_of_range_error():
    from pyrddlgym_b200 import workloads
    from oracle import ast_builders
    from oracle.ref_import import load_reference
    from pyrddlgym_b200.backend import backend_class
    from fake_engine import EmulatorEngine
    ref = load_reference()
    spec = workloads.reservoir(6)
    spec['init_non_fluent'] = [(k, (-1.0 if k[0] == 'RAIN_VAR' and k[1] == ['r3'] else v))
                               for (k, v) in spec['init_non_fluent']]
    b = backend_class()(ast_builders.build_model(spec), n_envs=1, engine_factory=EmulatorEngine, keep_tensors=True)
    b.reset()
    with pytest.raises(ref.exc.RDDLValueOutOfRangeError):
        b.step(b.prepare_actions_for_sim({}))


@pytest.mark.parametrize('seed', range(300, 312))
def test_random_deterministic_domains_dropin_matches_reference(seed):
    """Random deterministic domains (tests/fuzz_specs.py): the drop-in subclass returns what the reference
    returns -- observation dicts (grounded names / tensors, object values as strings or indices), float
    reward, bool done, check results -- step after step under the same grounded action dicts."""
    from fuzz_specs import random_spec
    kw = dict(keep_tensors=bool(seed % 2), objects_as_strings=bool((seed // 2) % 2))
    spec_fn = lambda: random_spec(seed, depth=3, n=3 + seed % 3, m=2 + seed % 2, deterministic=True)
    ref, a, b = _pair(spec_fn, **kw)
    oa, da = a.reset()
    ob, db = b.reset()
    assert da == db and set(oa) == set(ob)
    rng = np.random.default_rng(seed)
    n = 3 + seed % 3
    for t in range(4):
        act = {f'act___a{i + 1}': float(np.round(rng.uniform(-2, 2), 3)) for i in range(n) if rng.random() < 0.7}
        act.update({f'bact___a{i + 1}': True for i in range(n) if rng.random() < 0.4})
        if rng.random() < 0.5:
            act['iact'] = int(rng.integers(0, 4))
        sa, sb = a.prepare_actions_for_sim(act), b.prepare_actions_for_sim(act)
        assert a.check_action_preconditions(sa, silent=True) == b.check_action_preconditions(sb, silent=True)
        oa, ra, da = a.step(sa)
        ob, rb, db = b.step(sb)
        assert isinstance(rb, float) and isinstance(db, bool)
        np.testing.assert_allclose(rb, ra, rtol=1e-9, atol=1e-12)
        assert da == db and set(oa) == set(ob)
        for k in oa:
            if isinstance(oa[k], str) or (isinstance(oa[k], np.ndarray) and oa[k].dtype.kind in 'UO'):
                assert np.array_equal(np.asarray(oa[k]), np.asarray(ob[k])), k
            else:
                _same(oa[k], ob[k])
        assert a.check_state_invariants(silent=True) == b.check_state_invariants(silent=True)
