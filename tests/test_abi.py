"""The C-ABI library builds, loads, and exports every symbol include/rddl_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'rddl_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(rddl_[a-z_0-9]+)\s*\(', text)))


def test_header_symbols_are_exported():
    import __graft_entry__ as g
    lib_path = g.build()
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert 'rddl_step' in names and 'rddl_program_create' in names
    for n in names:
        assert hasattr(lib, n), f'{n} declared in rddl_b200.h but not exported'
    from pyrddlgym_b200 import native
    assert sorted(native.EXPORTS) == [n for n in names if n in native.EXPORTS]
    assert set(native.EXPORTS) <= set(names)


def test_create_rejects_garbage_without_gpu():
    from pyrddlgym_b200 import native
    lib = native.load()
    h = ctypes.c_void_p()
    buf = ctypes.create_string_buffer(b'\0' * 256, 256)
    rc = lib.rddl_program_create(ctypes.cast(buf, ctypes.c_void_p), 256, 0, ctypes.byref(h))
    assert rc == -1
    assert b'magic' in lib.rddl_last_error(None)


def test_optable_structs_match_the_c_header():
    from pyrddlgym_b200 import optable as ot
    text = open(os.path.join(ROOT, 'pyrddlgym_b200', 'csrc', 'rddl_plan.h')).read()
    ops = re.search(r'enum RddlOp : uint8_t \{(.*?)OP_COUNT_', text, re.S).group(1)
    c_ops = [o.strip()[3:] for o in ops.replace('\n', ' ').split(',') if o.strip()]
    assert c_ops == ot._OPS
    assert int(re.search(r'RDDL_VERSION (\d+)', text).group(1)) == ot.VERSION


def test_backend_refuses_cpu_device():
    import pytest
    from pyrddlgym_b200 import native, workloads
    from pyrddlgym_b200.simulator import BatchedSimulator
    p = workloads.load_program(workloads.cartpole())
    with pytest.raises(native.NativeError):
        BatchedSimulator(p, n_envs=2, device='cpu')


def test_every_workload_optable_decodes_on_the_host():
    """rddl_program_create parses and pre-decodes the op table before touching CUDA: without a
    GPU it must get past the format checks and fail at cudaSetDevice (RDDL_ERR_CUDA = -2)."""
    import ctypes
    import torch
    from pyrddlgym_b200 import lowering, native, workloads
    if torch.cuda.is_available():
        import pytest
        pytest.skip('host-only check')
    lib = native.load()
    specs = [workloads.cartpole(), workloads.opzoo(), workloads.wildfire(32), workloads.wildfire(64, separable=True),
             workloads.reservoir(40), workloads.pair_contraction(24)]
    for spec in specs:
        blob = lowering.lower(workloads.load_program(spec)).blob
        h = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(blob, len(blob))
        rc = lib.rddl_program_create(ctypes.cast(buf, ctypes.c_void_p), len(blob), 0, ctypes.byref(h))
        assert rc == -2, (spec['name'], rc, lib.rddl_last_error(None))


def test_specialised_kernels_compile_without_a_gpu(tmp_path, monkeypatch):
    """rddl_plane_precompile: the generated constants + the kernel sources (csrc/rddl_plane.cuh,
    csrc/rddl_level_static.cuh) compile for sm_100a with NVRTC on a machine without a device, land in
    the cubin cache, and the second call is a cache hit."""
    import time
    from pyrddlgym_b200 import lowering, native, workloads
    monkeypatch.setenv('RDDL_B200_JIT_CACHE', str(tmp_path))
    tab = lowering.lower(workloads.load_program(workloads.wildfire(64, separable=True)))
    li = [i for i, L in enumerate(tab.launches) if L['kind'] == 1]
    n_scalar = 1 if any(L['kind'] == 0 for L in tab.launches) else 0      # one module for all scalar launches
    assert li
    src = native.plane_source(tab.blob, li[0], 4096)
    assert '#define PS_NOPS %d' % tab.launches[li[0]]['n_insns'] in src and 'constexpr PlaneProg kG' in src
    assert native.precompile(tab.blob, 4096) == len(li) + n_scalar
    planes = [f for f in os.listdir(tmp_path) if f.startswith('plane_') and f.endswith('.cubin')]
    levels = [f for f in os.listdir(tmp_path) if f.startswith('level_') and f.endswith('.cubin')]
    assert len(planes) == len(li) and len(levels) == n_scalar
    assert os.path.getsize(os.path.join(tmp_path, planes[0])) > 10000
    t0 = time.time()
    assert native.precompile(tab.blob, 4096) == len(li) + n_scalar
    assert time.time() - t0 < 0.5                      # served from the cache
    # a different batch geometry (smaller tiles for tiny batches) is a different plane kernel
    assert native.precompile(tab.blob, 3) == len(li) + n_scalar
    assert len([f for f in os.listdir(tmp_path) if f.startswith('plane_')]) == 2 * len(li)
    # a scalar-only program: one module holding a kernel per scalar launch
    cart = lowering.lower(workloads.load_program(workloads.cartpole()))
    assert native.precompile(cart.blob, 64) == 1


def test_tile_geometry_of_the_specialised_plane_kernel():
    """plane_geometry (csrc/rddl_vm.cu), read back through the generated constants: the tile always holds
    whole envs and fits the thread block; a batch that fits one wave of resident CTAs is spread so that
    no SM holds more than ceil(N / 148) + a tile of envs; big batches keep the 32768-cell tiles at 3 CTAs/SM."""
    import re
    from pyrddlgym_b200 import lowering, native, workloads
    tab = lowering.lower(workloads.load_program(workloads.wildfire(32)))
    wpe = 32 * 32 // 32
    seen = {}
    for n_envs in (1, 16, 37, 4096, 5000, 8192, 65536):
        src = native.plane_source(tab.blob, 0, n_envs)
        d = {k: int(v) for k, v in re.findall(r'#define (PS_\w+) (\d+)', src)}
        f = [x for x in re.split(r'[{},]+', src.split('kL = ')[1].split(';')[0]) if x.strip()]
        ept, ipt, chunks = int(f[14]), int(f[15]), int(f[16])
        assert chunks == 1 and ipt == ept * wpe and 1 <= ept <= 32
        assert d['PS_BLOCK'] % 32 == 0 and d['PS_BLOCK'] <= 256 and d['PS_BLOCK'] * d['PS_WPT'] >= ipt
        assert (d['PS_BLOCK'] - 32) * d['PS_WPT'] < ipt                      # no idle warp
        tiles = -(-n_envs // ept)
        if tiles <= 148 * d['PS_MINB']:                                       # one wave: balanced over the SMs
            per_sm = -(-tiles // 148)
            assert per_sm * ept <= -(-n_envs // 148) + ept
        seen[n_envs] = (d['PS_WPT'], ept, d['PS_BLOCK'], d['PS_MINB'])
    assert seen[4096] == (2, 14, 224, 2)          # BASELINE config 2: 293 CTAs of 14 envs, 2 per SM
    assert seen[65536] == (4, 32, 256, 3)         # many waves: full tiles, 3 resident CTAs
