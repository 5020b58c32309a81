"""TEST INFRASTRUCTURE ONLY -- run in the build container (where /root/reference exists).

Generates, by running the UNMODIFIED reference front end and simulator:
  * pyrddlgym_b200/programs/<domain>.json : the lifted HIR of every synthetic workload
    (reference parser-AST -> RDDLLiftedModel -> RDDLLevelAnalysis -> RDDLObjectsTracer ->
    frontend.program_from_simulator), so that machines without pyRDDLGym can run them;
  * tests/golden/<case>.npz : trajectories of N independent reference RDDLSimulator
    instances (base variates injected per AST node, see oracle/ast_builders.py) -- the
    golden vectors that pin oracle/oracle_sim.py and the CUDA path.

    python -m oracle.gen_fixtures [case ...]      (default: every case)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ast_builders, harness  # noqa: E402
from pyrddlgym_b200 import frontend, workloads  # noqa: E402

# case name -> (spec factory, n_envs, steps, action sampler kwargs)
CASES = {
    'cartpole': (lambda: workloads.cartpole(), 4, 40, dict(real=(-10.0, 10.0))),
    'opzoo': (lambda: workloads.opzoo(), 3, 5, {}),
    'wildfire8': (lambda: workloads.wildfire(8), 3, 8, dict(p_bool=0.05)),
    'wildfire8_sep': (lambda: workloads.wildfire(8, separable=True), 3, 8, dict(p_bool=0.05)),
    'wildfire32': (lambda: workloads.wildfire(32), 1, 4, dict(p_bool=0.05)),
    'reservoir30': (lambda: workloads.reservoir(30), 3, 6, dict(real=(-1.0, 10.0))),
    'reservoir30_exp': (lambda: workloads.reservoir(30, rain='exponential'), 3, 6, dict(real=(0.0, 10.0))),
    'pair16': (lambda: workloads.pair_contraction(16), 3, 6, dict(real=(-12.0, 12.0))),
    # BASELINE config [2] at full size: the shape that takes the env-tiled CSR gather kernels
    'reservoir1000': (lambda: workloads.reservoir(1000), 2, 3, dict(real=(-1.0, 10.0))),
}

# programs without golden trajectories (natively sampled distributions: moment / KS tests only)
PROGRAM_ONLY = {'distzoo': lambda: workloads.distzoo()}


def reference_program(spec):
    model = ast_builders.build_model(spec)
    cls = ast_builders.ref_simulator_class()
    sim = cls(model, rng=np.random.default_rng(0), keep_tensors=True, objects_as_strings=False)
    return frontend.program_from_simulator(sim, name=spec['name'])


def actions_for(program, n_envs, rng, p_bool=0.1, real=(-1.0, 3.0)):
    out = harness.random_actions(program, n_envs, rng, p_bool=p_bool)
    for name in program.names_of_kind('action'):
        if program.tensors[name].dtype == 'f':
            out[name] = rng.uniform(real[0], real[1], size=out[name].shape)
    return out


def gen_case(name):
    factory, N, T, akw = CASES[name]
    spec = factory()
    p = reference_program(spec)
    rb = harness.ReferenceBatch(spec, N)
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    out = {'n_envs': N, 'steps': T}
    out['done0'] = rb.reset()
    fluents = [n for n in p.tensors if p.tensors[n].kind not in ('nonfluent', 'next', 'action')]
    for t in range(T):
        a = actions_for(p, N, rng, **akw)
        v = harness.random_variates(p, N, rng)
        for k, x in a.items():
            out[f'act/{t}/{k}'] = x
        for k, x in v.items():
            out[f'var/{t}/{k}'] = x
        if p.preconditions:
            out[f'pre/{t}'] = rb.preconditions(a)
        r, d = rb.step(a, v)
        out[f'reward/{t}'] = r
        out[f'done/{t}'] = d
        if p.invariants:
            out[f'inv/{t}'] = rb.invariants()
        for f in fluents:
            out[f'flu/{t}/{f}'] = rb.fluent(f)
    return p, out


def main():
    os.makedirs(os.path.join(ROOT, 'tests', 'golden'), exist_ok=True)
    os.makedirs(os.path.join(ROOT, 'pyrddlgym_b200', 'programs'), exist_ok=True)
    only = [a for a in sys.argv[1:] if a in CASES]
    for name in (only or CASES):
        p, out = gen_case(name)
        path = workloads.program_path(p.name)
        types, objects = p.types, p.objects
        p.types, p.objects = {}, {}          # lifted: object counts come from the instance
        with open(path, 'w') as f:
            f.write(p.to_json())
        p.types, p.objects = types, objects
        np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', name + '.npz'), **out)
        print(name, '->', os.path.relpath(path, ROOT), os.path.getsize(path), 'B;',
              'golden', os.path.getsize(os.path.join(ROOT, 'tests', 'golden', name + '.npz')), 'B')
    for name, factory in ({} if only else PROGRAM_ONLY).items():
        p = reference_program(factory())
        path = workloads.program_path(p.name)
        p.types, p.objects = {}, {}
        with open(path, 'w') as f:
            f.write(p.to_json())
        print(name, '->', os.path.relpath(path, ROOT))


if __name__ == '__main__':
    main()
