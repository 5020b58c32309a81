import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import cases
from pyrddlgym_b200 import workloads
from pyrddlgym_b200.simulator import BatchedSimulator
def mk(p, n, **kw): return BatchedSimulator(p, n_envs=n, device='cuda:0', **kw)
_np = lambda x: x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
which = sys.argv[1]
if which != 'none':
    cases.replay(which, mk, to_np=_np)
g = cases.load_golden('wildfire32')
p = workloads.load_program(cases.SPECS['wildfire32']())
sim = mk(p, 1)
sim.reset()
acts, var = cases.golden_step_inputs(g, 0)
sim.step(acts, var)
torch.cuda.synchronize()
got = _np(sim.fluent('burning'))[0]; want = g['flu/0/burning'][0]
bad = np.argwhere(got != want)
print('mismatches', len(bad), bad.tolist())
T = np.asarray(p.init['TARGET'])
print('TARGET at bad', [bool(T[x, y]) for x, y in bad], 'got', [bool(got[x,y]) for x,y in bad], 'want', [bool(want[x,y]) for x,y in bad])
print('put-out at bad', [bool(acts['put-out'][0][x,y]) for x,y in bad])
print('reward', _np(sim.t['__reward']), g['reward/0'])
