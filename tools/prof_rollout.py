"""Short fused-rollout run for ncu (wildfire n x n, N envs, T steps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyrddlgym_b200 import workloads
from pyrddlgym_b200.simulator import BatchedSimulator
n, N, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
p = workloads.load_program(workloads.wildfire(n, separable=(n > 48)))
sim = BatchedSimulator(p, n_envs=N, seed=1)
g = torch.Generator(device='cuda').manual_seed(0)
seq = {'put-out': torch.rand((T, N, n, n), device='cuda', generator=g) < 0.05,
       'cut-out': torch.rand((T, N, n, n), device='cuda', generator=g) < 0.05}
for _ in range(3):
    sim.reset(); sim.rollout(seq, T)
torch.cuda.synchronize()
print('done', float(sim.fluent('__reward').sum()))
