"""Times the launches of one workload's step plan (per-launch CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyrddlgym_b200 import workloads
from pyrddlgym_b200.simulator import BatchedSimulator
name, size, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
spec = {'reservoir': lambda: workloads.reservoir(size), 'reservoir_exp': lambda: workloads.reservoir(size, rain='exponential'),
        'pair': lambda: workloads.pair_contraction(size), 'cartpole': lambda: workloads.cartpole()}[name]()
p = workloads.load_program(spec)
sim = BatchedSimulator(p, n_envs=N, seed=1, specialize=True)
print([(l['kernel'], l['scope'], l['n_insns'], l['nregs']) for l in sim.table.launches if l['plan'] == 0])
g = torch.Generator(device='cuda').manual_seed(0)
acts = {}
for an in p.names_of_kind('action'):
    shp = (N,) + tuple(p.tensor_shape(an))
    acts[an] = torch.rand(shp, device='cuda', generator=g, dtype=torch.float64) * 5.0
for _ in range(3):
    sim.step(acts)
torch.cuda.synchronize()
sim.native.profile(True)
for _ in range(K):
    sim.step(acts)
    if name != 'cartpole':
        sim.check_state_invariants()
torch.cuda.synchronize()
prof = sim.native.profile_read(sim.table.launches)
tot = sum(r['ms'] for r in prof)
for r in prof:
    print(f"  launch {r['launch']:2d} {r['name']:<30s} {1e3 * r['ms'] / r['count']:9.1f} us x{r['count']}")
print(f'{name} size={size} N={N}: {1e3 * tot / K:.1f} us per step(+checks) -> {N * K / (tot * 1e-3):.3e} env-steps/s')
sim.raise_for_status()
