"""Short single-GPU run for ncu: a few steps of one workload (no timing claims)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyrddlgym_b200 import workloads
from pyrddlgym_b200.simulator import BatchedSimulator

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
p = workloads.load_program(workloads.wildfire(n, separable=(n > 48)))
sim = BatchedSimulator(p, n_envs=N, seed=1)
g = torch.Generator(device='cuda').manual_seed(0)
acts = [{'put-out': torch.rand((N, n, n), device='cuda', generator=g) < 0.05,
         'cut-out': torch.rand((N, n, n), device='cuda', generator=g) < 0.05} for _ in range(2)]
for t in range(steps):
    sim.step(acts[t & 1])
torch.cuda.synchronize()
print('done', n, N, steps, float(sim.fluent('__reward').sum()))
