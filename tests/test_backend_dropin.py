"""The drop-in subclass B200RDDLSimulator against the live, unmodified reference RDDLSimulator on the
same model and actions. Needs the reference: /root/reference in the build container, the pip-installed
copy under baseline/_ref on the GPU box (oracle/install_reference.py). Every test runs twice:
``emulator`` -- the kernels replaced by the op-table emulator through the engine seam (host logic, CPU)
-- and ``cuda`` (-m gpu) -- the real CUDA engine through the C ABI."""
import numpy as np
import pytest

from oracle.ref_import import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason='reference pyRDDLGym not present')

ENGINES = ['emulator', pytest.param('cuda', marks=pytest.mark.gpu)]


def _factory(engine):
    if engine == 'cuda':
        return None               # backend default: BatchedSimulator on cuda:0
    from fake_engine import EmulatorEngine
    return EmulatorEngine


def _pair(spec_fn, engine='emulator', **kw):
    from oracle import ast_builders
    from oracle.ref_import import load_reference
    from pyrddlgym_b200.backend import backend_class
    ref = load_reference()
    a = ref.RDDLSimulator(ast_builders.build_model(spec_fn()), rng=np.random.default_rng(0), **kw)
    b = backend_class()(ast_builders.build_model(spec_fn()), n_envs=1, engine_factory=_factory(engine), **kw)
    if engine == 'cuda':
        from pyrddlgym_b200.simulator import BatchedSimulator
        assert isinstance(b.engine, BatchedSimulator)
    return ref, a, b


def _same(x, y):
    assert type(x) is type(y) or (np.isscalar(x) and np.isscalar(y)), (type(x), type(y))
    np.testing.assert_allclose(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('engine', ENGINES)
@pytest.mark.parametrize('keep_tensors', [False, True])
def test_cartpole_dropin_matches_reference(keep_tensors, engine):
    from pyrddlgym_b200 import workloads
    ref, a, b = _pair(workloads.cartpole, engine, keep_tensors=keep_tensors)
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


@pytest.mark.parametrize('engine', ENGINES)
def test_precondition_violation_raises_reference_exception(engine):
    from pyrddlgym_b200 import workloads
    ref, a, b = _pair(workloads.cartpole, engine, keep_tensors=True)
    a.reset()
    b.reset()
    bad = b.prepare_actions_for_sim({'force': 11.0})
    assert b.check_action_preconditions(bad, silent=True) is False
    with pytest.raises(ref.exc.RDDLActionPreconditionNotSatisfiedError):
        b.check_action_preconditions(bad, silent=False)
    with pytest.raises(ref.exc.RDDLActionPreconditionNotSatisfiedError):
        a.check_action_preconditions(a.prepare_actions_for_sim({'force': 11.0}), silent=False)


@pytest.mark.parametrize('engine', ENGINES)
def test_pair_contraction_invariants_and_grounded_state(engine):
    from pyrddlgym_b200 import workloads
    ref, a, b = _pair(lambda: workloads.pair_contraction(12), engine, keep_tensors=False)
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


@pytest.mark.parametrize('engine', ENGINES)
def test_invalid_distribution_parameter_maps_to_value_out_of_range_error(engine):
    from pyrddlgym_b200 import workloads
    from oracle import ast_builders
    from oracle.ref_import import load_reference
    from pyrddlgym_b200.backend import backend_class
    EmulatorEngine = _factory(engine)
    ref = load_reference()
    spec = workloads.reservoir(6)
    spec['init_non_fluent'] = [(k, (-1.0 if k[0] == 'RAIN_VAR' and k[1] == ['r3'] else v))
                               for (k, v) in spec['init_non_fluent']]
    b = backend_class()(ast_builders.build_model(spec), n_envs=1, engine_factory=EmulatorEngine, keep_tensors=True)
    b.reset()
    with pytest.raises(ref.exc.RDDLValueOutOfRangeError):
        b.step(b.prepare_actions_for_sim({}))


@pytest.mark.parametrize('engine', ENGINES)
@pytest.mark.parametrize('seed', range(300, 312))
def test_random_deterministic_domains_dropin_matches_reference(seed, engine):
    """Random deterministic domains (tests/fuzz_specs.py): the drop-in subclass returns what the reference
    returns -- observation dicts (grounded names / tensors, object values as strings or indices), float
    reward, bool done, check results -- step after step under the same grounded action dicts."""
    from fuzz_specs import random_spec
    kw = dict(keep_tensors=bool(seed % 2), objects_as_strings=bool((seed // 2) % 2))
    spec_fn = lambda: random_spec(seed, depth=3, n=3 + seed % 3, m=2 + seed % 2, deterministic=True)
    ref, a, b = _pair(spec_fn, engine, **kw)
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


@pytest.mark.gpu
@pytest.mark.parametrize('case,N,T', [('wildfire8', 5, 8), ('wildfire8_sep', 3, 6), ('reservoir30', 4, 6),
                                       ('reservoir30_exp', 2, 4), ('opzoo', 3, 5), ('cartpole', 6, 12)])
def test_cuda_engine_in_lockstep_with_live_reference_simulators(case, N, T):
    """The definition of batched parity: N independent *unmodified* reference
    simulators, stepped live on this machine, against the CUDA kernels on the same actions and -- for
    stochastic CPFs -- the same injected base variates. bool / int exact, reals to 1e-5."""
    import cases
    from oracle import harness
    from pyrddlgym_b200 import frontend
    from pyrddlgym_b200.simulator import BatchedSimulator
    spec = cases.SPECS[case]()
    rb = harness.ReferenceBatch(spec, N)
    p = frontend.program_from_simulator(rb.sims[0])
    sim = BatchedSimulator(p, n_envs=N, device='cuda:0', seed=3)
    cases.compare('done0', sim.reset()[1].cpu().numpy(), rb.reset())
    rng = np.random.default_rng(17)
    for t in range(T):
        acts = harness.random_actions(p, N, rng, p_bool=0.1)
        var = harness.random_variates(p, N, rng)
        if p.preconditions:
            cases.compare('pre', sim.check_action_preconditions(acts).cpu().numpy(), rb.preconditions(acts))
        _, r, d = sim.step(acts, var)
        r0, d0 = rb.step(acts, var)
        cases.compare(f'reward/{t}', r.cpu().numpy(), r0)
        cases.compare(f'done/{t}', d.cpu().numpy(), d0)
        if p.invariants:
            cases.compare('inv', sim.check_state_invariants().cpu().numpy(), rb.invariants())
        for name, decl in p.tensors.items():
            if decl.kind not in ('nonfluent', 'next', 'action'):
                cases.compare(f'{name}/{t}', sim.fluent(name).cpu().numpy(), rb.fluent(name))
    sim.raise_for_status()


@pytest.mark.gpu
def test_batched_dropin_subclass_on_cuda_vs_reference_simulators():
    """B200RDDLSimulator(n_envs = N) on the CUDA engine: reset / step / check_* with a leading env axis
    against N reference simulators (deterministic domain: no variates to share)."""
    from oracle import ast_builders, harness
    from pyrddlgym_b200 import workloads
    from pyrddlgym_b200.backend import backend_class
    N = 7
    spec = workloads.cartpole()
    rb = harness.ReferenceBatch(spec, N)
    b = backend_class()(ast_builders.build_model(spec), n_envs=N, keep_tensors=True)
    _, d0 = b.reset()
    assert np.array_equal(d0.cpu().numpy().astype(bool), rb.reset())
    rng = np.random.default_rng(2)
    for t in range(40):
        force = rng.uniform(-12, 12, size=N)            # some outside FORCE-MAX: preconditions fail per env
        acts = {'force': force}
        assert np.array_equal(b.check_action_preconditions(acts).cpu().numpy(), rb.preconditions(acts))
        st, r, d = b.step(acts)
        r0, d0 = rb.step(acts, {})
        np.testing.assert_allclose(r.cpu().numpy(), r0, rtol=1e-9)
        assert np.array_equal(d.cpu().numpy().astype(bool), d0)
        for name in ('pos', 'ang-pos', 'vel', 'ang-vel'):
            np.testing.assert_allclose(st[name].cpu().numpy(), rb.fluent(name), rtol=1e-9, atol=1e-12)
    assert set(b.states) == {'pos', 'ang-pos', 'vel', 'ang-vel'}      # RDDLSimulator.states works batched


_A19_DOMAIN = '''
domain deferred {
    types { node : object; };
    pvariables {
        M(node, node) : { non-fluent, real, default = 0.0 };
        MU(node)      : { non-fluent, real, default = 0.0 };
        x(node)       : { state-fluent, real, default = 1.0 };
        u             : { action-fluent, real, default = 0.0 };
    };
    cpfs { %s };
    reward = u;
}
non-fluents nf_deferred { domain = deferred; objects { node : {a, b}; };
    non-fluents { M(a, a) = 2.0; M(b, b) = 3.0; M(a, b) = 0.5; M(b, a) = 0.5; }; }
instance inst_deferred { domain = deferred; non-fluents = nf_deferred; max-nondef-actions = pos-inf; horizon = 5; discount = 1.0; }
'''


@pytest.mark.parametrize('cpf,python_functions', [
    ("x'(?i) = x(?i) + det_{?p : node, ?q : node} M(?p, ?q);", None),
    ("x'(?i) = sum_{?j : node} (inverse[row = ?i, col = ?j][M(?i, ?j)] * x(?j));", None),
    ("x'(?i) = MultivariateNormal[?i](MU(_), M(_, _));", None),
    ("x'(?i) = Dirichlet[?i](x(_));", None),
    ("x'(?i) = x(?i) + $ext[](u);", {'ext': lambda u: u}),
])
def test_deferred_constructs_raise_not_implemented(cpf, python_functions):
    """external Python functions, random vectors and matrix operators are on the
    reference's path (simulator.py:824-888,1274-1422) but cannot run on the device; the backend refuses
    them at compile time with the reference's own RDDLNotImplementedError -- no silent CPU fallback --
    while the reference itself builds and steps the same model."""
    from fake_engine import EmulatorEngine
    from oracle.ref_import import load_reference
    from pyrddlgym_b200 import rddl_reader
    from pyrddlgym_b200.backend import backend_class
    ref = load_reference()
    spec = rddl_reader.parse_rddl(_A19_DOMAIN % cpf)
    kw = {'python_functions': python_functions} if python_functions else {}
    a = ref.RDDLSimulator(ref.RDDLLiftedModel(rddl_reader.reference_ast(spec)), rng=np.random.default_rng(0),
                          keep_tensors=True, **kw)
    a.reset()
    a.step(a.prepare_actions_for_sim({'u': 0.5}))                  # the reference runs it
    with pytest.raises(ref.exc.RDDLNotImplementedError):
        backend_class()(ref.RDDLLiftedModel(rddl_reader.reference_ast(spec)), n_envs=1, engine_factory=EmulatorEngine,
                        keep_tensors=True, **kw)
