"""bench.py -- batched env-steps/sec of the B200 RDDL stepping backend (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload wildfire32]
    python bench.py --impl reference ...        # CPU port of the reference algorithm (oracle)

A "step" is one batched transition (RDDLSimulator.step for every env) of BASELINE.json
config[1]: Wildfire 32x32, 4096 envs per GPU, random-policy actions (i.i.d. Bernoulli(0.05)
put-out / cut-out per cell). `value` is timed with the action planes already resident in HBM
(a rotating pool larger than L2); `e2e` goes through the host-facing call with pinned host
action buffers copied H2D and the new state + reward + done copied D2H every step.
Multi-GPU (torchrun): env batch sharded, 4096 envs per rank (weak scaling), one NCCL
all-reduce of the return statistics at the end of the timed rollout.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = 'batched env-steps/sec'
UNIT = 'env-steps/s'

WORKLOADS = {
    # BASELINE.json configs: [1] wildfire32 (the default), [2] reservoir1000, [3] wildfire1024, [4] pair512
    'wildfire32': dict(kind='wildfire', n=32, separable=False, envs=4096),
    'wildfire32_sep': dict(kind='wildfire', n=32, separable=True, envs=4096),
    'wildfire1024': dict(kind='wildfire', n=1024, separable=True, envs=256),
    'reservoir1000': dict(kind='reservoir', n=1000, envs=4096),
    'pair512': dict(kind='pair', n=512, envs=1024),
}


def make_spec(name):
    from pyrddlgym_b200 import workloads
    w = WORKLOADS[name]
    if w['kind'] == 'wildfire':
        return workloads.wildfire(w['n'], separable=w['separable'])
    if w['kind'] == 'reservoir':
        return workloads.reservoir(w['n'])
    return workloads.pair_contraction(w['n'])


def algorithmic_bytes_per_env_step(name):
    """SURVEY 8(d). Wildfire: read burning, out-of-fuel, put-out, cut-out + write burning',
    out-of-fuel' (1 byte per bool cell) + reward f64 + done u8. Reservoir (f64): read rlevel,
    release + write rlevel' (interm fluents counted as on-chip). Pair contraction: read x, u +
    write x' (W is shared by all envs)."""
    w = WORKLOADS[name]
    if w['kind'] == 'wildfire':
        return 6 * w['n'] * w['n'] + 9
    return 24 * w['n'] + 9


# DRAM traffic of the dominant kernel per env-step, from `ncu --set full` captures
# (dram__bytes_read.sum + dram__bytes_write.sum; profiles/r01_ncu_full_*.csv). None = not captured.
MEASURED_TRAFFIC_BYTES_PER_ENV_STEP = {
    'wildfire32': (176.496128e6 + 7.141888e6) / (20 * 4096),      # fused 20-step rollout, plane_c2_v6
    'wildfire1024': (537.102592e6 + 223.248384e6) / 128,          # one step, plane_1024_v5
}


def workload_label(name):
    w = WORKLOADS[name]
    if w['kind'] == 'wildfire':
        return f"Wildfire {w['n']}x{w['n']} ({'separable' if w['separable'] else '4-D NEIGHBOR'})"
    if w['kind'] == 'reservoir':
        return f"Reservoir R={w['n']} (sum-over-upstream, Normal rain)"
    return f"Pair sum-product contraction {w['n']}x{w['n']} + invariants"


def draw_actions(name, p, k, N, dev, g):
    """Random-policy action sequence [k, N, ...] resident on the device."""
    import torch
    w = WORKLOADS[name]
    out = {}
    for an in p.names_of_kind('action'):
        shp = (k, N) + tuple(p.tensor_shape(an))
        if p.tensors[an].dtype == 'b':
            out[an] = torch.rand(shp, device=dev, generator=g) < 0.05
        elif w['kind'] == 'reservoir':
            out[an] = torch.rand(shp, device=dev, generator=g, dtype=torch.float64) * 10.0
        else:
            out[an] = torch.rand(shp, device=dev, generator=g, dtype=torch.float64) * 2.0 - 1.0
    return out


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML from a background thread while the GPU
    is under load (in-process: spawning nvidia-smi next to the timed region stalls launches)."""

    REASONS = [('hw_slowdown', 0x8), ('sw_power_cap', 0x4), ('sw_thermal_slowdown', 0x20),
               ('hw_thermal_slowdown', 0x40), ('hw_power_brake', 0x80)]

    def __init__(self, index=0, period=0.02):
        self.index, self.period = index, period
        self.sm, self.reasons = [], set()
        self.max_mhz = None
        self.timed = False
        self._stop = threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                if self.timed:
                    self.sm.append(mhz)
                    for name, bit in self.REASONS:
                        if bits & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        if self.nv is None or self.thread is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvml unavailable']}
        self._stop.set()
        self.thread.join(timeout=1)
        return {'sm_mhz': float(np.median(self.sm)) if self.sm else None, 'sm_max_mhz': self.max_mhz,
                'samples': len(self.sm), 'reasons': sorted(self.reasons)}


# ---------------------------------------------------------------------------------------
# CPU arm: the NumPy oracle port of the reference algorithm
# ---------------------------------------------------------------------------------------

def _cpu_worker(args):
    name, n_envs, steps, seed = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    from pyrddlgym_b200 import workloads
    from oracle.oracle_sim import OracleSimulator
    p = workloads.load_program(make_spec(name))
    sim = OracleSimulator(p, n_envs=n_envs, seed=seed)
    rng = np.random.default_rng(seed)
    acts = []
    for _ in range(2):
        a = {}
        for an in p.names_of_kind('action'):
            shp = (n_envs,) + tuple(p.tensor_shape(an))
            a[an] = (rng.random(shp) < 0.05) if p.tensors[an].dtype == 'b' else rng.random(shp) * 5.0
        acts.append(a)
    sim.step(acts[0])
    t0 = time.perf_counter()
    for t in range(steps):
        sim.step(acts[t & 1])
    return time.perf_counter() - t0


def cpu_baseline(name, cores=1, envs_per_core=4, steps=8):
    """Oracle port timed on `cores` host cores (one process each, envs_per_core envs, `steps`
    steps). Returns env-steps/s aggregated over cores."""
    if WORKLOADS[name]['kind'] == 'wildfire' and WORKLOADS[name]['n'] > 64:
        return None, 'reference formulation cannot run n>48 (O(n^4) temporaries)'
    jobs = [(name, envs_per_core, steps, 100 + i) for i in range(cores)]
    if cores == 1:
        dts = [_cpu_worker(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context('spawn').Pool(cores) as pool:
            dts = pool.map(_cpu_worker, jobs)
    total = cores * envs_per_core * steps
    return total / max(dts), f'{cores} process(es) x {envs_per_core} envs x {steps} steps of {name} (NumPy oracle port)'


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    use = max(1, min(cores, 32))
    per_step_envs = 4
    t0 = time.perf_counter()
    # warm-up + K timed "steps": each step = every core advancing per_step_envs envs once
    v, sample = cpu_baseline(args.workload, cores=use, envs_per_core=per_step_envs,
                             steps=max(1, args.steps))
    wall = time.perf_counter() - t0
    if v is None:
        print(json.dumps({'impl': 'reference', 'unavailable': sample}))
        return
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * use * per_step_envs / v, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'u8/f64', 'data': 'synthetic',
        'config': {'workload': args.workload, 'envs_per_step': use * per_step_envs, 'wall_s': round(wall, 2)},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': use, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------

def run_gpu(args):
    import torch
    import torch.distributed as dist
    from pyrddlgym_b200 import workloads
    from pyrddlgym_b200.simulator import BatchedSimulator

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    name = args.workload
    w = WORKLOADS[name]
    n, N = w['n'], args.envs or w['envs']
    p = workloads.load_program(make_spec(name))
    sim = BatchedSimulator(p, n_envs=N, device=dev, seed=42, env_offset=rank * N)
    K, W = args.steps, args.warmup
    gamma = float(p.discount)
    act_bytes = sum(N * int(np.prod(p.tensor_shape(a), dtype=np.int64)) * (1 if p.tensors[a].dtype == 'b' else 8)
                    for a in p.names_of_kind('action'))       # action bytes per step
    state_bytes = sum(N * int(np.prod(p.tensor_shape(a), dtype=np.int64)) * (1 if p.tensors[a].dtype == 'b' else 8)
                      for a in p.next_state)

    # random-policy action sequence for the timed rollout, resident in HBM: [K, N, n, n] x 2
    # planes (larger than the 126 MB L2 unless K is tiny); the warm-up uses its own sequence
    g = torch.Generator(device=dev).manual_seed(1234 + rank)

    def draw(k):
        return draw_actions(name, p, k, N, dev, g)

    seq = draw(K)
    warm = draw(max(W, 1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    returns = torch.zeros(N, dtype=torch.float64, device=dev)

    def reduce_stats():
        # statistics of BaseAgent.evaluate (pyRDDLGym/core/policy.py:120-126) over all ranks:
        # two NCCL all-reduces of a few scalars (pyrddlgym_b200/parallel.py)
        from pyrddlgym_b200 import parallel
        return parallel.reduce_return_stats(returns)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # bring the clocks up before the W warm-up steps (an idle GPU needs ~100 ms to boost)
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 0.3:
        sim.rollout(warm, max(W, 1), gamma=gamma)
        torch.cuda.synchronize()
    sim.reset()
    sim.rollout(warm, max(W, 1), gamma=gamma, returns=returns)     # W untimed warm-up steps
    reduce_stats()                                                 # (loads torch's reduction kernels)
    returns.zero_()
    barrier()
    l0 = sim.launches_issued
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.timed = True
    e0.record()
    sim.rollout(seq, K, gamma=gamma, returns=returns)              # exactly K timed steps
    stats = reduce_stats()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sim.launches_issued - l0
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = world * N * K / (ms * 1e-3)

    # the same K steps through the per-step call (rddl_step per transition), for comparison
    sim.reset()
    for t in range(min(W, K)):
        sim.step({k: v[t] for k, v in seq.items()})
    barrier()
    e0.record()
    for t in range(K):
        sim.step({k: v[t] for k, v in seq.items()})
    e1.record()
    barrier()
    step_api_value = N * K / (e0.elapsed_time(e1) * 1e-3)

    # dominant kernel: per-launch CUDA events on the launch stream (second pass, same inputs)
    sim.reset()
    sim.native.profile(True)
    sim.rollout(seq, K, gamma=gamma)
    torch.cuda.synchronize()
    prof = sim.native.profile_read(sim.table.launches)
    sim.native.profile(False)
    sampler.timed = False
    top = max(prof, key=lambda r: r['ms']) if prof else None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured' if 'hbm_gbs' in peaks else 'fallback'
    roofline = None
    if top:
        steps_per_launch = K / top['count']
        abytes = algorithmic_bytes_per_env_step(name) * N * steps_per_launch
        per_launch_ms = top['ms'] / top['count']
        ach = abytes / (per_launch_ms * 1e-3) / 1e9
        roofline = {'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                    'traffic': (MEASURED_TRAFFIC_BYTES_PER_ENV_STEP[name] * N * steps_per_launch
                                if name in MEASURED_TRAFFIC_BYTES_PER_ENV_STEP else None),
                    'peak_source': peak_src, 'kernel': top['name'],
                    'kernel_us': per_launch_ms * 1e3, 'steps_per_launch': steps_per_launch,
                    'kernel_share_of_step': top['ms'] / max(1e-9, sum(r['ms'] for r in prof)),
                    'algorithmic_bytes_per_launch': abytes,
                    'note': 'algorithmic bytes per env-step as in SURVEY 8(d) (state + action reads, next-state writes, reward, done); '
                            'a fused multi-step launch keeps the state planes on chip, so its DRAM '
                            'traffic is below this figure'}

    # end-to-end through the host-facing call: pinned host action sequence -> H2D, rollout,
    # per-step rewards + returns + final state -> D2H
    h_seq = {k: v.cpu().pin_memory() for k, v in seq.items()}
    d_seq = {k: torch.empty_like(v) for k, v in seq.items()}
    h_state = {s: torch.empty_like(sim.fluent(s), device='cpu').pin_memory() for s in sim.state_names}
    h_rew = torch.empty((K, N), dtype=torch.float64).pin_memory()
    h_ret = torch.empty(N, dtype=torch.float64).pin_memory()

    def e2e_rollout():
        for k in d_seq:
            d_seq[k].copy_(h_seq[k], non_blocking=True)
        ret, rew = sim.rollout(d_seq, K, gamma=gamma, keep_rewards=True)
        for s in sim.state_names:
            h_state[s].copy_(sim.fluent(s), non_blocking=True)
        h_rew.copy_(rew, non_blocking=True)
        h_ret.copy_(ret, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller owns the results after the call

    sim.reset()
    e2e_rollout()
    sim.reset()
    barrier()
    e0.record()
    e2e_rollout()
    e1.record()
    barrier()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * N * K / (float(ms_e2e.item()) * 1e-3)
    clocks = sampler.stop() if rank == 0 else None
    h2d = act_bytes
    d2h = (state_bytes + 8 * N) // K + 8 * N

    if rank == 0:
        cpu_v, cpu_sample = cpu_baseline(name, cores=1, envs_per_core=4, steps=16) if world == 1 else (None, None)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u8/f64', 'data': 'synthetic',
            'config': {'workload': f'{workload_label(name)}, '
                                   f'{N} envs/GPU, {K}-step random-policy rollout',
                       'envs_per_gpu': N, 'global_envs': N * world, 'grid': n,
                       'api': 'rddl_rollout (K steps per call); per-step rddl_step loop reported as step_api_value',
                       'step_api_value': step_api_value,
                       'l2': f'inputs larger than L2: the action sequence read by the timed rollout is '
                             f'{K * act_bytes / 1e6:.0f} MB (L2 = 126 MB); state tensors {state_bytes / 1e6:.1f} MB',
                       'parallelism': f'env-shard x{world}, return statistics all-reduced (NCCL) once per rollout',
                       'launches': [l['kernel'] for l in sim.table.launches if l['plan'] == 0],
                       'return_stats': stats},
            'roofline': roofline,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': float(ms_e2e.item()) / K},
            'gpu_launches': launches,
            'clocks': clocks,
        }
        if cpu_v is not None:
            line['cpu_baseline'] = {'value': cpu_v, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'sample': cpu_sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='wildfire32', choices=sorted(WORKLOADS))
    ap.add_argument('--envs', type=int, default=0)
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
