"""Device-side random policy (RandomAgent / NoOpAgent of /root/reference/pyRDDLGym/core/policy.py:129-195
drawn inside the kernels) and the max-nondef count rule (simulator.py:329-360).

CPU: properties of the NumPy restatement of the policy draw (oracle/philox_np.py) and the C structs.
GPU: the kernels against that restatement bit for bit, and rddl_rollout_policy against the same
rollout with the materialised action sequences."""
import ctypes

import numpy as np
import pytest

from oracle import philox_np
from pyrddlgym_b200 import hir, native, workloads


def _entries(sim, desc):
    """The descriptor as the plain dicts oracle/philox_np.policy_actions takes."""
    out = []
    for i in range(desc.n_entries):
        e = desc.entries[i]
        name = sim.names[e.tensor]
        out.append(dict(tensor=int(e.tensor), kind=int(e.kind), lo=e.lo, hi=e.hi, dflt=e.dflt,
                        shape=tuple(sim.program.tensor_shape(name)), dtype=sim.program.tensors[name].dtype))
    return out, (None if desc.max_nondef < 0 else int(desc.max_nondef))


# ------------------------------------------------------------------------------------------ CPU

def test_policy_structs_match_the_header_layout():
    assert ctypes.sizeof(native.PolicyEntry) == 32
    assert ctypes.sizeof(native.Policy) == 8 + 8 * 32


def test_policy_restatement_selection_is_a_uniform_k_subset():
    k, total = 5, 37
    sel = philox_np.policy_select(9, np.arange(4000), 3, k, total)
    assert sel.shape == (4000, k) and sel.min() >= 0 and sel.max() < total
    assert all(len(set(r)) == k for r in sel.tolist())              # without replacement
    freq = np.bincount(sel.reshape(-1), minlength=total) / (4000 * k)
    assert np.abs(freq - 1.0 / total).max() < 0.25 / total            # every grounding equally likely
    assert not np.array_equal(sel, philox_np.policy_select(9, np.arange(4000), 4, k, total))


def test_policy_restatement_draws():
    ents = [dict(tensor=3, kind=philox_np.POL_BERNOULLI, lo=0.05, hi=0, dflt=False, shape=(32, 32), dtype='b'),
            dict(tensor=4, kind=philox_np.POL_UNIFORM_REAL, lo=-1.0, hi=3.0, dflt=0.0, shape=(7,), dtype='f'),
            dict(tensor=5, kind=philox_np.POL_UNIFORM_INT, lo=2, hi=5, dflt=0, shape=(), dtype='i'),
            dict(tensor=6, kind=philox_np.POL_DEFAULT, lo=0, hi=0, dflt=True, shape=(3,), dtype='b')]
    env = np.arange(600)
    a = philox_np.policy_actions(ents, None, 42, env, 7)
    assert abs(a[3].mean() - 0.05) < 0.002 and a[3].dtype == np.bool_
    assert a[4].min() >= -1.0 and a[4].max() < 3.0 and abs(a[4].mean() - 1.0) < 0.1
    assert set(np.unique(a[5]).tolist()) == {2, 3, 4, 5} and a[6].all()
    # shard invariance: a draw depends on the global env id only
    b = philox_np.policy_actions(ents, None, 42, env[300:], 7)
    assert all(np.array_equal(a[t][300:], b[t]) for t in a)
    # max-nondef: at most k groundings per env differ from the default
    c = philox_np.policy_actions(ents, 3, 42, env, 7)
    nondef = (c[3] != False).reshape(600, -1).sum(1) + (c[4] != 0.0).sum(1) + (c[5] != 0) + (~c[6]).sum(1)  # noqa: E712
    assert nondef.max() <= 3
    z = philox_np.policy_actions(ents, 0, 42, env, 7)
    assert not z[3].any() and not z[4].any() and not z[5].any()


# ------------------------------------------------------------------------------------------ GPU

def _make(p, n, **kw):
    from pyrddlgym_b200.simulator import BatchedSimulator
    return BatchedSimulator(p, n_envs=n, device='cuda:0', **kw)


def _np(x):
    return x.detach().cpu().numpy()


def _with_scalar_action(p):
    """Wildfire whose reward also reads a scalar real action `u` (a fused scalar epilogue that LOADs
    an action fluent -- the case the fused rollout must not take)."""
    p.tensors['u'] = hir.TensorDecl('u', 'action', 'f', [], 'real')
    p.init['u'] = np.zeros((), dtype=np.float64)
    load = hir.Node('load', tensor='u', index=[], dtype='f', scope=[])
    p.reward = hir.Node('bin', op='sub', args=[p.reward, load], dtype='f', scope=[])
    p.max_allowed_actions += 1
    return p


def _programs():
    return {
        'wildfire32': (lambda: workloads.load_program(workloads.wildfire(32)), dict(p_bool=0.05), None),
        'wildfire32_k3': (lambda: workloads.load_program(workloads.wildfire(32, max_nondef=3)), dict(p_bool=0.5), None),
        'wildfire32_k1': (lambda: workloads.load_program(workloads.wildfire(32, max_nondef=1)), dict(p_bool=0.5), None),
        'wildfire256_sep': (lambda: workloads.load_program(workloads.wildfire(256, separable=True)), dict(p_bool=0.03), None),
        'wildfire32_u': (lambda: _with_scalar_action(workloads.load_program(workloads.wildfire(32))),
                         dict(p_bool=0.05, bounds={'u': (-2.0, 5.0)}), None),
        'cartpole': (lambda: workloads.load_program(workloads.cartpole()), dict(bounds={'force': (-10.0, 10.0)}), None),
        'reservoir64': (lambda: workloads.load_program(workloads.reservoir(64)), dict(bounds={'release': (0.0, 10.0)}), None),
        'reservoir64_k4': (lambda: workloads.load_program(workloads.reservoir(64)), dict(bounds={'release': (0.0, 10.0)}), 4),
        'opzoo': (lambda: workloads.load_program(workloads.opzoo()), dict(p_bool=0.3), 'auto'),
    }


def _policy(name, p):
    from pyrddlgym_b200.policy import RandomPolicy
    _, kw, k = _programs()[name]
    kw = dict(kw)
    if k == 'auto':      # every non-bool action fluent gets bounds
        kw['bounds'] = {n: ((0, 3) if p.tensors[n].dtype == 'i' else (-1.0, 2.0)) for n in p.names_of_kind('action')}
        k = None
    return RandomPolicy(max_allowed_actions=k, **kw)


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(_programs()))
def test_policy_sample_matches_the_numpy_restatement(name):
    p = _programs()[name][0]()
    N = 37
    sim = _make(p, N, seed=1234567891011, env_offset=5)
    pol = _policy(name, p)
    if len(sim.action_names) > native.MAX_POLICY_ENTRIES:
        pytest.skip('more action fluents than policy entries')
    desc = pol.descriptor(sim)
    ents, k = _entries(sim, desc)
    for step in (0, 3, 70000):
        got = sim.sample_policy_actions(desc, step)
        want = philox_np.policy_actions(ents, k, sim.seed_value, np.arange(N) + 5, step)
        for e in ents:
            nm = sim.names[e['tensor']]
            assert np.array_equal(_np(got[nm]), want[e['tensor']]), (name, nm, step)
    if k is not None and k > 0:
        counts = _np(sim.count_nondefault_actions())
        assert counts.max() <= k
        total = sum(want[e['tensor']].reshape(N, -1).shape[1] for e in ents)
        if all(e['kind'] != philox_np.POL_BERNOULLI for e in ents):
            assert counts.min() == min(k, total)           # continuous draws never hit the default


@pytest.mark.gpu
@pytest.mark.parametrize('specialize', [None, False])
@pytest.mark.parametrize('name', sorted(_programs()))
def test_rollout_policy_equals_rollout_with_materialised_actions(name, specialize):
    """rddl_rollout_policy (actions drawn in registers) == rddl_policy_sample per step + rddl_rollout,
    bit for bit: returns, per-step rewards, final state. Covers the fused single-launch rollout
    (wildfire32*), the multi-tile plane path (wildfire256), the scalar kernels and a fused scalar
    epilogue that reads an action (wildfire32_u)."""
    import torch
    p = _programs()[name][0]()
    N, T = 45, 6
    a = _make(p, N, seed=77, env_offset=3, specialize=specialize)
    b = _make(p, N, seed=77, env_offset=3, specialize=specialize)
    if len(a.action_names) > native.MAX_POLICY_ENTRIES:
        pytest.skip('more action fluents than policy entries')
    pol = _policy(name, p)
    ra, wa = a.rollout_policy(pol.descriptor(a), T, keep_rewards=True)
    seq = pol.sample(b, T)
    rb, wb = b.rollout(seq, T, keep_rewards=True)
    assert torch.equal(wa, wb), name
    assert torch.equal(ra, rb), name
    for s in a.state_names:
        assert torch.equal(a.fluent(s), b.fluent(s)), (name, s)
    assert torch.equal(a.status, b.status)
    assert a.step_count == b.step_count == T
    if name == 'wildfire32':
        assert a.launches_issued < b.launches_issued or a.launches_issued <= 3      # one fused launch
        assert bool(a.fluent('burning').any())
    # one policy-driven transition (n_steps = 1) continues both identically
    ra1, wa1 = a.rollout_policy(pol.descriptor(a), 1, keep_rewards=True)
    nxt = pol.sample(b, 1)
    _, rwd, _ = b.step({k: v[0] for k, v in nxt.items()})
    assert torch.equal(wa1[0], rwd), name
    assert torch.equal(ra1, wa1[0])                      # returns of a fresh 1-step rollout = its reward


@pytest.mark.gpu
def test_rollout_with_one_step_and_mixed_grid_scalar_programs_equals_steps():
    """rollout(n_steps=1) accumulates returns / rewards; a plane program whose fused scalar
    epilogue reads a scalar action takes the per-step path (rollout == repeated steps)."""
    import torch
    from oracle import harness
    for mk in (lambda: workloads.load_program(workloads.wildfire(32)),
               lambda: _with_scalar_action(workloads.load_program(workloads.wildfire(32)))):
        p = mk()
        N = 19
        rng = np.random.default_rng(5)
        for T in (1, 4):
            a, b = _make(p, N, seed=9), _make(p, N, seed=9)
            acts = [harness.random_actions(p, N, rng, p_bool=0.05) for _ in range(T)]
            seq = {k: torch.from_numpy(np.stack([x[k] for x in acts])).cuda() for k in acts[0]}
            ret, rew = a.rollout(seq, T, gamma=0.9, keep_rewards=True)
            want = torch.zeros(N, dtype=torch.float64, device='cuda')
            for t in range(T):
                _, r, _ = b.step(acts[t])
                assert torch.equal(rew[t], r), (T, t)
                want += r * (0.9 ** t)
            assert torch.allclose(ret, want, rtol=1e-12, atol=0) and bool((ret != 0).any())
            for s in a.state_names:
                assert torch.equal(a.fluent(s), b.fluent(s))


@pytest.mark.gpu
def test_check_default_action_count_is_a_device_reduction():
    import torch
    from pyrddlgym_b200.simulator import InvalidActionCount
    p = workloads.load_program(workloads.wildfire(32, max_nondef=2))
    N = 50
    sim = _make(p, N)
    put = torch.zeros((N, 32, 32), dtype=torch.bool, device='cuda')
    cut = torch.zeros((N, 32, 32), dtype=torch.bool, device='cuda')
    put[7, 3, 4] = True
    cut[7, 9, 9] = True
    sim.check_default_action_count({'put-out': put, 'cut-out': cut})
    assert _np(sim.count_nondefault_actions()).tolist() == [2 if e == 7 else 0 for e in range(N)]
    cut[33, 0, 0] = cut[33, 31, 31] = put[33, 5, 5] = True
    with pytest.raises(InvalidActionCount) as ei:
        sim.check_default_action_count({'put-out': put, 'cut-out': cut})
    assert 'got 3' in str(ei.value) and int(ei.value.counts[33]) == 3
    # real-valued actions: enforce_for_non_bool (simulator.py:353-354)
    pr = workloads.load_program(workloads.reservoir(30))
    sr = _make(pr, 4)
    rel = torch.zeros((4, 30), dtype=torch.float64, device='cuda')
    rel[2, :5] = 1.5
    assert _np(sr.count_nondefault_actions({'release': rel})).tolist() == [0, 0, 5, 0]
    assert _np(sr.count_nondefault_actions({'release': rel}, enforce_for_non_bool=False)).tolist() == [0, 0, 0, 0]


@pytest.mark.gpu
def test_evaluate_uses_the_device_policy_and_episodes_differ_after_reset():
    """the Philox step index is not rewound by reset(), so successive episodes of the same
    envs draw fresh noise; seed() restarts it (same seed -> same statistics)."""
    import torch
    from pyrddlgym_b200.policy import RandomPolicy, evaluate
    p = workloads.load_program(workloads.wildfire(32))
    sim = _make(p, 64, seed=5)
    pol = RandomPolicy(p_bool=0.05)
    st1 = evaluate(sim, pol, n_steps=20, seed=11)
    st1b = evaluate(sim, pol, n_steps=20, seed=11)
    assert st1 == st1b
    desc = pol.descriptor(sim)
    sim.reset()
    r1, _ = sim.rollout_policy(desc, 20)
    sim.reset()
    r2, _ = sim.rollout_policy(desc, 20)
    assert not torch.equal(r1, r2)
