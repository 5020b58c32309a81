"""Writes the specialised-kernel source of a wildfire plane launch (for offline nvcc / SASS inspection).
usage: python tools/jit_source.py <n> <n_envs> <outdir>"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyrddlgym_b200 import workloads, lowering, native

n, N, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
p = workloads.load_program(workloads.wildfire(n, separable=(n > 48)))
t = lowering.lower(p)
lib = ctypes.CDLL(native.LIB_PATH)
lib.rddl_plane_source.restype = ctypes.c_int64
lib.rddl_plane_source.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int64, ctypes.c_char_p, ctypes.c_size_t]
blob = bytes(t.blob)
li = [i for i, L in enumerate(t.launches) if L['kind'] == 1][0]
buf = ctypes.create_string_buffer(1 << 20)
r = lib.rddl_plane_source(blob, len(blob), li, N, buf, len(buf))
assert r > 0, r
os.makedirs(out, exist_ok=True)
open(os.path.join(out, 'plane_static_data.inc'), 'wb').write(buf.value)
open(os.path.join(out, 'kernel.cu'), 'w').write('#define PLANE_STATIC 1\n#include "rddl_device.cuh"\n#include "rddl_plane.cuh"\n')
print('wrote', out, r, 'bytes')
