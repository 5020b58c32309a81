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
