"""Writes profiles/r02_sass_*.txt: opcode histograms (cuobjdump -sass) of the hot kernels as this tree builds them --
the ahead-of-time tcgen05 contraction from librddl_b200.so, and the NVRTC-specialised plane / fused scalar kernels
of the bench workloads (compiled here: NVRTC needs no device). usage: python tools/sass_excerpt.py"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: F401
from pyrddlgym_b200 import lowering, native, workloads

OUT = os.path.join(ROOT, 'profiles')
SO = os.path.join(ROOT, 'pyrddlgym_b200', 'csrc', 'librddl_b200.so')
KEY = re.compile(r'UTC|UTMA|LDTM|STTM|SYNCS|ACQBULK|LDGSTS|DEPBAR|LDG\.E\.(EF\.)?128|STG\.E\.(EF\.)?128|LOP3|POPC|DMMA|DADD|DFMA|DMUL|MUFU|LDS|REDUX|SHFL|IMAD\.WIDE')


def histogram(sass):
    ops = collections.Counter()
    for l in sass.splitlines():
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Za-z0-9_.]+)', l)
        if m:
            ops[m.group(2)] += 1
    return ops


def write(name, title, sass, note=''):
    ops = histogram(sass)
    with open(os.path.join(OUT, name), 'w') as f:
        f.write(f'{title}\n{note}\ninstructions: {sum(ops.values())}\n\nmnemonics that matter for this kernel:\n')
        for k, v in sorted(ops.items(), key=lambda kv: -kv[1]):
            if KEY.search(k):
                f.write(f'  {v:6d}  {k}\n')
        f.write('\nall opcodes (base mnemonic):\n')
        base = collections.Counter()
        for k, v in ops.items():
            base[k.split('.')[0]] += v
        for k, v in sorted(base.items(), key=lambda kv: -kv[1]):
            f.write(f'  {v:6d}  {k}\n')
    print('wrote', name, sum(ops.values()), 'instructions')


def sass_of(path, fun=None):
    cmd = ['cuobjdump', '-sass'] + (['-fun', fun] if fun else []) + [path]
    return subprocess.run(cmd, capture_output=True, text=True).stdout


# 1. tcgen05 contraction (ahead of time)
for bn in (32, 64):
    fun = f'_Z16rddl_contract_tcILi{bn}EEv14CUtensorMap_stS0_Pdlii'
    write(f'r02_sass_contract_tc_bn{bn}.txt', f'rddl_contract_tc<{bn}> (csrc/rddl_contract_tc.cuh), sm_100a, from librddl_b200.so',
          sass_of(SO, fun), 'UTCHMMA = tcgen05.mma kind::f16, UTMALDG = cp.async.bulk.tensor, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit')

# 1b. FP64 DMMA contraction, cp.async pipeline (ahead of time)
for bn in (32, 64):
    fun = f'_Z17rddl_contract_f64ILi{bn}ELb1EEvPKdS1_PdliillS2_Pji'
    write(f'r02_sass_contract_f64_async_bn{bn}.txt', f'rddl_contract_f64<{bn}, true> (csrc/rddl_contract.cuh), sm_100a, from librddl_b200.so',
          sass_of(SO, fun), 'DMMA = mma.sync.m8n8k4.f64, LDGSTS = cp.async (16-byte copies global -> shared), LDGDEPBAR / DEPBAR = commit / wait group')

# 2. specialised kernels: compile into a scratch cache and disassemble the cubins
with tempfile.TemporaryDirectory() as tmp:
    os.environ['RDDL_B200_JIT_CACHE'] = tmp
    jobs = [('plane_static_wildfire1024_4096env', workloads.wildfire(1024, separable=True), 4096, 'rddl_plane_static'),
            ('plane_static_wildfire32_4096env', workloads.wildfire(32), 4096, 'rddl_plane_static'),
            ('level_fused_reservoir1000', workloads.reservoir(1000), 4096, 'rddl_level_fused_0')]
    for tag, spec, envs, fun in jobs:
        before = set(os.listdir(tmp))
        native.precompile(lowering.lower(workloads.load_program(spec)).blob, envs)
        for cub in sorted(set(os.listdir(tmp)) - before):
            s = sass_of(os.path.join(tmp, cub), fun)
            if 'Function' in s:
                write(f'r02_sass_{tag}.txt', f'{fun} of {tag} (NVRTC-specialised build, sm_100a)', s,
                      'LDG.E.EF.128 / STG.E.EF.128 = 128-bit streaming accesses of the byte planes; LOP3 = one logical node for 32 cells')
