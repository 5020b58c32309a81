"""Specialised vs interpreter plane kernels over many batch sizes (every tile geometry plane_geometry can pick)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyrddlgym_b200 import workloads
from pyrddlgym_b200.simulator import BatchedSimulator

bad = 0
for n, sep in ((32, False), (8, True), (64, True), (96, True)):
    p = workloads.load_program(workloads.wildfire(n, separable=sep))
    for N in (1, 2, 13, 147, 148, 149, 293, 443, 1000, 2961, 4099):
        a = BatchedSimulator(p, n_envs=N, seed=7)
        b = BatchedSimulator(p, n_envs=N, seed=7, specialize=False)
        g = torch.Generator(device='cuda').manual_seed(N)
        T = 3
        seq = {k: torch.rand((T, N, n, n), device='cuda', generator=g) < 0.1 for k in ('put-out', 'cut-out')}
        a.step({k: v[0] for k, v in seq.items()}); b.step({k: v[0] for k, v in seq.items()})
        ra, wa = a.rollout(seq, T, keep_rewards=True); rb, wb = b.rollout(seq, T, keep_rewards=True)
        ok = torch.equal(ra, rb) and torch.equal(wa, wb) and all(torch.equal(a.fluent(s), b.fluent(s)) for s in p.next_state)
        has_plane = any(l['kind'] == 1 for l in a.table.launches)
        ok = ok and bool(a.fluent('burning').any()) and (a.specialised_kernels >= 1 or not has_plane) and int(a.status.max()) == 0
        bad += not ok
        print(n, sep, N, 'ok' if ok else 'MISMATCH', flush=True)
print('failures', bad)
sys.exit(1 if bad else 0)
