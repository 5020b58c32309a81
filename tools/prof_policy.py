"""Short device-side-policy run for ncu (wildfire n x n, N envs, T-step policy rollouts; no timing claims)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyrddlgym_b200 import workloads
from pyrddlgym_b200.policy import RandomPolicy
from pyrddlgym_b200.simulator import BatchedSimulator
n, N, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
p = workloads.load_program(workloads.wildfire(n, separable=(n > 48)))
sim = BatchedSimulator(p, n_envs=N, seed=1, specialize=True)
desc = RandomPolicy(p_bool=0.05).descriptor(sim)
for _ in range(3):
    sim.reset()
    ret, _ = sim.rollout_policy(desc, T)
torch.cuda.synchronize()
print('done', float(ret.mean()))
