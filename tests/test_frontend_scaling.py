"""Front end for huge instances: frontend.program_from_spec runs the unmodified reference
front end on a surrogate instance of a few objects per type and re-instantiates the lifted program at
full size. Needs the reference (build container / baseline/_ref)."""
import time

import numpy as np
import pytest

from oracle.ref_import import reference_available
from pyrddlgym_b200 import workloads

pytestmark = pytest.mark.skipif(not reference_available(), reason='reference pyRDDLGym not present')


def _full(spec):
    from oracle import ast_builders
    from oracle.ref_import import load_reference
    from pyrddlgym_b200 import frontend
    ref = load_reference()
    sim = ref.RDDLSimulator(ast_builders.build_model(spec), rng=np.random.default_rng(0), keep_tensors=True)
    return frontend.program_from_simulator(sim, name=spec['name'])


@pytest.mark.parametrize('name,make', [
    ('wildfire16', lambda: workloads.wildfire(16)),
    ('wildfire24_sep', lambda: workloads.wildfire(24, separable=True)),
    ('reservoir40', lambda: workloads.reservoir(40)),
    ('reservoir40_exp', lambda: workloads.reservoir(40, rain='exponential')),
    ('pair24', lambda: workloads.pair_contraction(24)),
    ('cartpole', lambda: workloads.cartpole()),
    ('opzoo', lambda: workloads.opzoo()),
    ('distzoo', lambda: workloads.distzoo()),
])
def test_scaled_front_end_equals_the_full_front_end(name, make):
    """Same program -- initial tensors, HIR, op table byte for byte -- as the reference front end run on
    the whole instance, for every workload family (object and enum literals, switch / Discrete case
    orders, nested indices, two interm levels)."""
    from oracle.ref_import import load_reference
    from pyrddlgym_b200 import frontend, lowering
    load_reference()
    a = frontend.program_from_spec(make())
    b = _full(make())
    assert a.types == b.types and a.objects == b.objects and set(a.init) == set(b.init)
    for k in a.init:
        # (scalars: () here, (1,) from np.ascontiguousarray in the full path)
        assert a.init[k].dtype == b.init[k].dtype and np.array_equal(a.init[k].reshape(-1), b.init[k].reshape(-1)), k
    assert (a.horizon, a.discount, a.max_allowed_actions) == (b.horizon, b.discount, b.max_allowed_actions)
    assert lowering.lower(a).blob == lowering.lower(b).blob


def test_wildfire_1024_builds_in_seconds_and_equals_the_shipped_program():
    """BASELINE config [3]: the reference cannot build this instance at all; the scaled
    front end does in seconds, and the result is the program the benchmark runs (workloads.load_program
    of the lifted JSON generated from the reference at small size)."""
    from oracle.ref_import import load_reference
    from pyrddlgym_b200 import frontend, lowering
    load_reference()
    spec = workloads.wildfire(1024, separable=True)
    t0 = time.perf_counter()
    p = frontend.program_from_spec(spec)
    dt = time.perf_counter() - t0
    assert dt < 60.0, dt
    q = workloads.load_program(spec)
    assert p.types == {'x_pos': 1024, 'y_pos': 1024}
    for k in q.init:
        assert np.array_equal(p.init[k], q.init[k]), k
    assert lowering.lower(p).blob == lowering.lower(q).blob


def test_scaled_front_end_from_rddl_text():
    """.rddl text -> no-PLY reader -> scaled front end -> program (the whole path a user without ply takes)."""
    from oracle.ref_import import load_reference
    from pyrddlgym_b200 import frontend, lowering, rddl_reader
    from test_rddl_reader import DOMAIN
    load_reference()
    spec = rddl_reader.parse_rddl(DOMAIN)
    p = frontend.program_from_spec(spec)
    assert p.types['res'] == 3 and p.init['CAP'].tolist() == [4.0, 10.5, 10.5] and p.init['K'] == 7
    assert p.init['LINK'][0, 1] and not p.init['LINK'][1, 2] and p.init['m'] == 1
    tab = lowering.lower(p)
    assert tab.n_insns > 0 and p.horizon == 20
