"""The NumPy oracle replays the reference-generated golden trajectories (pins the oracle)."""
import numpy as np
import pytest

import cases
from oracle.oracle_sim import OracleSimulator


class _Adapter(OracleSimulator):
    def fluent(self, name):
        return self.subs[name]


@pytest.mark.parametrize('case', sorted(cases.SPECS))
def test_oracle_matches_reference_golden(case):
    sim = cases.replay(case, lambda p, n: _Adapter(p, n))
    assert not sim.status.any()


@pytest.mark.parametrize('case', cases.FUZZSPEC_CASES)
def test_oracle_matches_reference_on_random_domains(case):
    """Random domains (tests/fuzz_specs.py), trajectories from the reference (oracle/gen_fuzz_fixtures.py)."""
    cases.replay(case, lambda p, n: _Adapter(p, n))


@pytest.mark.parametrize('case', ['wildfire8', 'wildfire8_sep', 'wildfire32'])
def test_stencil_restatement_matches_reference_golden(case):
    """Pins oracle/wildfire_stencil.py (used at grid sizes the reference cannot build)."""
    from oracle.wildfire_stencil import wildfire_step
    from pyrddlgym_b200 import workloads
    g = cases.load_golden(case)
    p = workloads.load_program(cases.SPECS[case]())
    (node,) = p.random_nodes()
    N = int(g['n_envs'])
    target = np.asarray(p.init['TARGET'])
    b = np.broadcast_to(np.asarray(p.init['burning']), (N,) + target.shape).copy()
    f = np.broadcast_to(np.asarray(p.init['out-of-fuel']), (N,) + target.shape).copy()
    for t in range(int(g['steps'])):
        acts, var = cases.golden_step_inputs(g, t)
        b, f, r = wildfire_step(b, f, acts['put-out'], acts['cut-out'], target, var[node.id])
        assert np.array_equal(b, g[f'flu/{t}/burning'])
        assert np.array_equal(f, g[f'flu/{t}/out-of-fuel'])
        np.testing.assert_allclose(r, g[f'reward/{t}'], rtol=1e-12)
