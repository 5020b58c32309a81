"""Fills the specialised-kernel cubin cache (pyrddlgym_b200/csrc/jit_cache) on a machine without a GPU:
NVRTC compiles for sm_100a anywhere. usage: python tools/prewarm_jit.py [n:N ...]  (wildfire n x n, N envs)"""
import os, sys, time
import torch  # noqa: F401  (first: the process then binds the same NVRTC as the simulator does -> same cache keys)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyrddlgym_b200 import workloads, lowering, native

cfgs = [tuple(int(v) for v in a.split(':')) for a in sys.argv[1:]] or [(32, 4096), (1024, 128), (256, 512)]
for n, N in cfgs:
    for sep in ({False, n > 48} if n <= 48 else {True}):
        t = lowering.lower(workloads.load_program(workloads.wildfire(n, separable=sep)))
        t0 = time.time()
        k = native.precompile(t.blob, N)
        print(f'wildfire {n}x{n} sep={sep} N={N}: {k} kernel(s) in {time.time() - t0:.1f} s', flush=True)
