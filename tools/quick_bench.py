import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from pyrddlgym_b200 import workloads
from pyrddlgym_b200.simulator import BatchedSimulator
for (n, N, sep) in [(32, 4096, False), (32, 4096, True), (128, 256, True)]:
    p = workloads.load_program(workloads.wildfire(n, separable=sep))
    sim = BatchedSimulator(p, n_envs=N, seed=1)
    put = torch.rand((N, n, n), device='cuda') < 0.05
    cut = torch.rand((N, n, n), device='cuda') < 0.05
    acts = {'put-out': put, 'cut-out': cut}
    for _ in range(3): sim.step(acts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 20
    e0.record()
    for _ in range(K): sim.step(acts)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    byt = 6 * n * n * N
    print(f'wildfire n={n} N={N} sep={sep}: {ms:.3f} ms/step, {N/ms*1e3:.3e} env-steps/s, {byt/ms/1e6:.1f} GB/s algorithmic', flush=True)
