"""Times the dominant kernel of a wildfire config (per-launch CUDA events), for A/B runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyrddlgym_b200 import workloads
from pyrddlgym_b200.simulator import BatchedSimulator
n = int(sys.argv[1]); N = int(sys.argv[2]); K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
p = workloads.load_program(workloads.wildfire(n, separable=(n > 48)))
sim = BatchedSimulator(p, n_envs=N, seed=1)
g = torch.Generator(device='cuda').manual_seed(0)
seq = {'put-out': torch.rand((K, N, n, n), device='cuda', generator=g) < 0.05,
       'cut-out': torch.rand((K, N, n, n), device='cuda', generator=g) < 0.05}
for _ in range(3):
    sim.reset(); sim.rollout(seq, K)
torch.cuda.synchronize()
sim.reset()
sim.native.profile(True)
sim.rollout(seq, K)
torch.cuda.synchronize()
prof = sim.native.profile_read(sim.table.launches)
top = max(prof, key=lambda r: r['ms'])
us_per_step = 1e3 * top['ms'] / K
gbs = (6 * n * n + 9) * N / (us_per_step * 1e-6) / 1e9
print(f"jit built={sim.native.info(5)} failed={sim.native.info(6)} cache_hits={sim.native.info(7)} nvrtc={sim.native.info(8)}", end=' ')
print(f"{os.environ.get('RDDL_B200_LIB', 'default'):>28s} n={n} N={N}: {us_per_step:8.1f} us/step  {gbs:7.0f} GB/s  {gbs / 6541.1:.3f} of peak", flush=True)
