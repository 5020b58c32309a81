"""Times the contraction launch of the pair workload (f64 DMMA kernel vs the tcgen05 split-bf16 path).
usage: python tools/time_contract.py [X] [N] [K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyrddlgym_b200 import workloads
from pyrddlgym_b200.simulator import BatchedSimulator
X = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
K = int(sys.argv[3]) if len(sys.argv) > 3 else 50
p = workloads.load_program(workloads.pair_contraction(X))
g = torch.Generator(device='cuda').manual_seed(0)
u = torch.rand((N, X), device='cuda', generator=g, dtype=torch.float64) * 2 - 1
res = {}
for prec in ('f64', 'bf16x3'):
    sim = BatchedSimulator(p, n_envs=N, seed=1, specialize=True, contract_precision=prec)
    for _ in range(30):
        sim.step({'u': u})
    torch.cuda.synchronize()
    sim.native.profile(True)
    for _ in range(K):
        sim.step({'u': u})
    torch.cuda.synchronize()
    prof = sim.native.profile_read(sim.table.launches)
    for r in prof:
        print(f"{prec:7s} launch {r['launch']:2d} {r['name']:<30s} {1e3 * r['ms'] / r['count']:9.2f} us x{r['count']}", flush=True)
    res[prec] = sim.t['__ct0'].clone()
    sim.raise_for_status()
d = (res['f64'] - res['bf16x3']).abs().max().item()
print('max |f64 - bf16x3| =', d, ' max |out| =', res['f64'].abs().max().item())

# whole steps replayed from a CUDA graph (no host launch latency between the kernels of a step)
for prec in ('f64', 'bf16x3'):
    sim = BatchedSimulator(p, n_envs=N, seed=1, specialize=True, contract_precision=prec)
    for _ in range(30):
        sim.step({'u': u})
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(20):
            sim.step({'u': u})
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f'{prec:7s} graph replay: {1e3 * e0.elapsed_time(e1) / 200:.2f} us per step', flush=True)
