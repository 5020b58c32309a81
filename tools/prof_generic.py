"""Short single-GPU run of a scalar workload for ncu (reservoir / pair / cartpole; no timing claims)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyrddlgym_b200 import workloads
from pyrddlgym_b200.simulator import BatchedSimulator
name, size, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
spec = {'reservoir': lambda: workloads.reservoir(size), 'pair': lambda: workloads.pair_contraction(size),
        'cartpole': lambda: workloads.cartpole()}[name]()
p = workloads.load_program(spec)
sim = BatchedSimulator(p, n_envs=N, seed=1, specialize=True, contract_precision=os.environ.get('CONTRACT', 'f64'))
g = torch.Generator(device='cuda').manual_seed(0)
acts = {an: torch.rand((N,) + tuple(p.tensor_shape(an)), device='cuda', generator=g, dtype=torch.float64) * 5.0
        for an in p.names_of_kind('action')}
for _ in range(K):
    sim.step(acts)
torch.cuda.synchronize()
print('done', name, size, N, K, float(sim.fluent('__reward').sum()))
