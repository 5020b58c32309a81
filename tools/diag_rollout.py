import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyrddlgym_b200 import workloads
from pyrddlgym_b200.simulator import BatchedSimulator
n, N, K = 32, 4096, 100
p = workloads.load_program(workloads.wildfire(n))
sim = BatchedSimulator(p, n_envs=N, seed=1)
g = torch.Generator(device='cuda').manual_seed(0)
seq = {'put-out': torch.rand((K, N, n, n), device='cuda', generator=g) < 0.05,
       'cut-out': torch.rand((K, N, n, n), device='cuda', generator=g) < 0.05}
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(4):
    sim.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    ret, _ = sim.rollout(seq, K)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'rep {rep}: events {e0.elapsed_time(e1):.3f} ms, host call {1e3*(t1-t0):.3f} ms, to sync {1e3*(t2-t0):.3f} ms, burning {int(sim.fluent("burning").sum())}', flush=True)
for T in (1, 2, 10, 50):
    sim.reset(); torch.cuda.synchronize()
    e0.record(); sim.rollout({k: v[:T] for k, v in seq.items()}, T); e1.record(); torch.cuda.synchronize()
    print(f'T={T}: {e0.elapsed_time(e1):.3f} ms -> {1e3*e0.elapsed_time(e1)/T:.1f} us/step')
