"""TEST INFRASTRUCTURE ONLY -- golden trajectories of RANDOM domains (tests/fuzz_specs.py) from the
unmodified reference simulator, for the machines that do not have the reference (the GPU box):

    python oracle/gen_fuzz_fixtures.py          # -> tests/golden/fuzzspec/<name>.program.npz, <name>.npz

<name>.program.npz is the front end's HIR program of the domain (hir.Program.save: JSON + initial tensors);
<name>.npz holds actions, injected base variates and the reference's rewards / terminations / fluents per step
(same keys as tests/golden/*.npz, oracle/gen_fixtures.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from fuzz_specs import random_grid_spec, random_spec          # noqa: E402
from oracle import harness                  # noqa: E402
from oracle.gen_fixtures import reference_program   # noqa: E402

SEEDS = list(range(100, 116))
GRID_SEEDS = list(range(500, 508))       # Wildfire-like grid domains (bit-plane path)
OUT = os.path.join(ROOT, 'tests', 'golden', 'fuzzspec')


def gen(seed, N=5, T=4, grid=False):
    if grid:
        H, W = [(3, 32), (2, 64), (5, 32), (1, 96)][seed % 4]
        spec = random_grid_spec(seed, H, W)
    else:
        spec = random_spec(seed, depth=3 + seed % 3)
    p = reference_program(spec)
    rb = harness.ReferenceBatch(spec, N)
    rng = np.random.default_rng(seed)
    out = {'n_envs': N, 'steps': T, 'done0': rb.reset()}
    fluents = [n for n in p.tensors if p.tensors[n].kind not in ('nonfluent', 'next', 'action')]
    for t in range(T):
        a = harness.random_actions(p, N, rng, p_bool=0.4)
        v = harness.random_variates(p, N, rng)
        for k, x in a.items():
            out[f'act/{t}/{k}'] = x
        for k, x in v.items():
            out[f'var/{t}/{k}'] = x
        out[f'pre/{t}'] = rb.preconditions(a)
        r, d = rb.step(a, v)
        out[f'reward/{t}'], out[f'done/{t}'] = r, d
        out[f'inv/{t}'] = rb.invariants()
        for f in fluents:
            out[f'flu/{t}/{f}'] = rb.fluent(f)
    return p, out


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    for seed, grid in [(s_, False) for s_ in SEEDS] + [(s_, True) for s_ in GRID_SEEDS]:
        p, out = gen(seed, grid=grid)
        p.save(os.path.join(OUT, p.name + '.program.npz'))
        np.savez_compressed(os.path.join(OUT, p.name + '.npz'), **out)
        print(p.name, os.path.getsize(os.path.join(OUT, p.name + '.program.npz')), os.path.getsize(os.path.join(OUT, p.name + '.npz')))
