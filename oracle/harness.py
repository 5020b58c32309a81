"""TEST INFRASTRUCTURE ONLY -- shared helpers for golden generation and parity tests:
seeded random actions, seeded base variates, and a lock-step driver for N reference
simulators (one per env, /root/reference/pyRDDLGym/core/simulator.py:442-502)."""
import numpy as np

from pyrddlgym_b200 import hir


def random_actions(program, n_envs, rng, p_bool=0.1):
    """name -> [N, *shape] array; bool ~ Bernoulli(p_bool), real ~ U(-1, 3), int ~ {0..3}."""
    out = {}
    for name in program.names_of_kind('action'):
        decl = program.tensors[name]
        shp = (n_envs,) + program.tensor_shape(name)
        if decl.dtype == 'b':
            out[name] = rng.random(shp) < p_bool
        elif decl.dtype == 'i':
            out[name] = rng.integers(0, 4, size=shp, dtype=np.int64)
        else:
            out[name] = rng.uniform(-1.0, 3.0, size=shp)
    return out


def random_variates(program, n_envs, rng):
    """node id -> base variates [N, *scope] of the kind each distribution consumes."""
    out = {}
    for n in program.random_nodes():
        shp = (n_envs,) + program.shape_of(n.scope)
        kind = hir.DIST_BASE[n.op]
        if kind == 'U':
            out[n.id] = rng.random(shp)
        elif kind == 'N':
            out[n.id] = rng.standard_normal(shp)
        elif kind == 'E':
            out[n.id] = rng.standard_exponential(shp)
        else:
            out[n.id] = rng.standard_cauchy(shp)
    return out


class ReferenceBatch:
    """N independent reference simulators stepped in lock-step (the definition of batched
    parity)."""

    def __init__(self, spec, n_envs):
        from oracle.ast_builders import build_model, ref_simulator_class
        cls = ref_simulator_class()
        self.sims = [cls(build_model(spec), rng=np.random.default_rng(0), keep_tensors=True,
                         objects_as_strings=False) for _ in range(n_envs)]
        self.N = n_envs

    def reset(self):
        dones = []
        for s in self.sims:
            _, d = s.reset()
            dones.append(d)
        return np.array(dones)

    def step(self, actions, variates):
        rewards, dones = [], []
        for e, s in enumerate(self.sims):
            s.variates = {k: v[e] for k, v in variates.items()}
            act = {}
            for k, v in actions.items():
                a = v[e]
                act[k] = a if a.ndim else a[()]
            _, r, d = s.step(act)
            rewards.append(r)
            dones.append(d)
        return np.array(rewards), np.array(dones)

    def fluent(self, name):
        return np.stack([np.asarray(s.subs[name]) for s in self.sims], axis=0)

    def invariants(self):
        return np.array([s.check_state_invariants(silent=True) for s in self.sims])

    def preconditions(self, actions):
        out = []
        for e, s in enumerate(self.sims):
            act = {k: (v[e] if v[e].ndim else v[e][()]) for k, v in actions.items()}
            out.append(s.check_action_preconditions(act, silent=True))
        return np.array(out)


def compare(name, got, want, rtol=1e-9, atol=1e-12):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f'{name}: shape {got.shape} vs {want.shape}'
    if want.dtype.kind in 'biu' and got.dtype.kind in 'biu':
        bad = got.astype(np.int64) != want.astype(np.int64)
        assert not bad.any(), f'{name}: {int(bad.sum())} integer/bool mismatches'
    else:
        np.testing.assert_allclose(got.astype(np.float64), want.astype(np.float64),
                                   rtol=rtol, atol=atol, err_msg=name)
