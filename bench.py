"""bench.py -- batched env-steps/sec of the B200 RDDL stepping backend (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload wildfire32]
    python bench.py --impl reference ...        # CPU port of the reference algorithm (oracle)

A bench "step" is one pass of the hot path over one batch of BASELINE.json config[1]: one 100-step
random-policy rollout (i.i.d. Bernoulli(0.05) put-out / cut-out per cell) of Wildfire 32x32 for 4096
envs per GPU -- BaseAgent.evaluate's episode for the whole batch: reset, 100 transitions with
discounted return accumulation, return statistics. `value` = env-steps/s with the action sequence
already resident in HBM (839 MB per rollout, larger than L2); `e2e` goes through the host-facing
call with the pinned host action sequence copied H2D and per-step rewards, returns and the final
state copied D2H every rollout. Multi-GPU (torchrun): env batch sharded, 4096 envs per rank (weak
scaling), one NCCL all-gather of five numbers per rank for the return statistics of each rollout.
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
    # rollout = env-steps per bench step (BASELINE config[1]: 100-step rollouts; shorter where the
    # action sequence of a rollout would not fit comfortably in HBM)
    'wildfire32': dict(kind='wildfire', n=32, separable=False, envs=4096, rollout=100),
    'wildfire32_sep': dict(kind='wildfire', n=32, separable=True, envs=4096, rollout=100),
    'wildfire1024': dict(kind='wildfire', n=1024, separable=True, envs=256, rollout=10),
    'reservoir1000': dict(kind='reservoir', n=1000, envs=4096, rollout=20),
    'pair512': dict(kind='pair', n=512, envs=1024, rollout=40),
}


_RESULT_FD = None


def emit(line):
    """The one JSON line of the run, on the real stdout."""
    data = (json.dumps(line) + '\n').encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


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
    # rddl_plane_static, fused 20-step rollout of 4096 envs (r01_ncu_full_plane_static_rollout20_32x32_4096env.csv)
    'wildfire32': (168.882688e6 + 4.095232e6) / (20 * 4096),
    # rddl_plane_static, one step of 128 envs (r01_ncu_full_plane_static_1024x1024_128env.csv)
    'wildfire1024': (302.261760e6 + 9.407744e6) / 128,
}


def layout_bytes_per_env_step(name, sim, steps_per_launch):
    """Bytes one env-step must move with the HBM layout this backend actually uses: action planes as
    the caller's bool bytes; bool state fluents the op table packs (OpTable.packed) as bits, and --
    when a launch fuses several steps -- read/written once per launch, not per step."""
    w = WORKLOADS[name]
    if w['kind'] != 'wildfire':
        return algorithmic_bytes_per_env_step(name)
    cells = w['n'] * w['n']
    state = 4 * cells * (0.125 if sim.packed else 1.0) / steps_per_launch     # 2 read + 2 written
    return 2 * cells + state + 9


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

    def sample(self):
        """One NVML query; recorded when inside the timed region."""
        if self.nv is None:
            return
        nv = self.nv
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

    def _run(self):
        while not self._stop.is_set():
            self.sample()
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
    name = args.workload
    w = WORKLOADS[name]
    cores = os.cpu_count() or 1
    use = max(1, min(cores, 32))
    per_step_envs, per_step_len = 4, 4
    t0 = time.perf_counter()
    # K timed "steps": each a bounded sample of the GPU arm's step (a rollout of the batch) -- every
    # core advances per_step_envs envs by per_step_len transitions of the same workload
    v, sample = cpu_baseline(name, cores=use, envs_per_core=per_step_envs, steps=max(1, args.steps) * per_step_len)
    wall = time.perf_counter() - t0
    if v is None:
        emit({'impl': 'reference', 'unavailable': sample})
        return
    N, R = args.envs or w['envs'], args.rollout or w['rollout']
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * use * per_step_envs * per_step_len / v, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'u8/f64', 'data': 'synthetic',
        'config': {'workload': f'{workload_label(name)}, {N} envs/GPU, {R}-step random-policy rollouts',
                   'step': f'bounded sample of one rollout of the batch: {use} cores x {per_step_envs} envs x '
                           f'{per_step_len} transitions = {use * per_step_envs * per_step_len} env-steps',
                   'wall_s': round(wall, 2)},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': use, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------

def run_gpu(args):
    import torch
    import torch.distributed as dist
    from pyrddlgym_b200 import parallel, workloads
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
    R = args.rollout or w['rollout']                 # env-steps per rollout = per bench step
    p = workloads.load_program(make_spec(name))
    # specialize=True: every kernel of the program is specialised on its first launch (during the warm-up)
    sim = BatchedSimulator(p, n_envs=N, device=dev, seed=42, env_offset=rank * N, specialize=True)
    K, W = args.steps, args.warmup
    gamma = float(p.discount)
    act_bytes = sum(N * int(np.prod(p.tensor_shape(a), dtype=np.int64)) * (1 if p.tensors[a].dtype == 'b' else 8)
                    for a in p.names_of_kind('action'))       # action bytes per env-step of the batch
    state_bytes = sum(N * int(np.prod(p.tensor_shape(a), dtype=np.int64)) * (1 if p.tensors[a].dtype == 'b' else 8)
                      for a in p.next_state)

    # random-policy action sequence of one rollout, resident in HBM: [R, N, ...] per action fluent
    # (larger than the 126 MB L2, so every rollout streams it from DRAM again)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    seq = draw_actions(name, p, R, N, dev, g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    returns = torch.zeros(N, dtype=torch.float64, device=dev)

    local_stats = torch.empty(5, dtype=torch.float64, device=dev)

    def episode_local():
        # one bench step = BaseAgent.evaluate's episode for every env of the batch
        # (pyRDDLGym/core/policy.py:85-126): reset, R transitions with discounted return
        # accumulation, this rank's part of the return statistics
        sim.reset()
        returns.zero_()
        sim.rollout(seq, R, gamma=gamma, returns=returns)
        local_stats.copy_(parallel.local_return_stats(returns))

    # The collective of episode k runs on a side stream and overlaps the rollout of episode k + 1
    # (two result slots; the main stream only waits before it reuses a slot); everything is
    # joined before the timed region ends.
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    slots = [torch.empty(5, dtype=torch.float64, device=dev) for _ in range(2)]
    slot_free = [None, None]
    counter = [0]

    def episode(graph=None):
        if graph is not None:
            graph.replay()
        else:
            episode_local()
        # ... and the statistics over all ranks: one NCCL all-gather of 5 numbers per rank
        if world == 1 or args.sync_collective:
            return parallel.gather_return_stats(local_stats)
        k = counter[0] & 1
        counter[0] += 1
        main = torch.cuda.current_stream()
        if slot_free[k] is not None:
            main.wait_event(slot_free[k])
        slots[k].copy_(local_stats)
        ready = torch.cuda.Event()
        ready.record(main)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            out = parallel.gather_return_stats(slots[k])
            slot_free[k] = torch.cuda.Event()
            slot_free[k].record(side)
        return out

    def join_collectives():
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)

    sampler = ClockSampler(local, period=0.001)
    if rank == 0:
        sampler.start()
    # bring the clocks up before the W warm-up steps (an idle GPU needs ~100 ms to boost)
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 0.3:
        episode_local()            # (local work only: a time-based loop must not issue collectives)
        torch.cuda.synchronize()
    launches_per_episode = 0
    for _ in range(max(W, 1)):                                     # W untimed warm-up steps
        la = sim.launches_issued
        g_ = episode()
        join_collectives()
        parallel.finish_return_stats(g_)
        launches_per_episode = sim.launches_issued - la
    # The local part of the episode is a fixed launch sequence (reset copies, one rollout kernel,
    # the statistics kernel): captured once into a CUDA graph and replayed per step, so the host's
    # per-launch latency is not part of the step.
    graph = None
    if not args.no_graph:
        try:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                episode_local()              # (the collective stays outside the graph)
            g_ = episode(graph)
            join_collectives()
            parallel.finish_return_stats(g_)
        except Exception as e:               # capture not possible here: issue the episodes eagerly
            sys.stderr.write('bench: CUDA graph capture failed (%s); running eagerly\n' % e)
            graph = None
            torch.cuda.synchronize()
            g_ = episode(None)               # (same number of collectives as the ranks whose capture worked)
            join_collectives()
            parallel.finish_return_stats(g_)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0 = sim.launches_issued
    sampler.timed = True
    e0.record()
    for _ in range(K):                                             # exactly K timed steps
        gathered = episode(graph)
    if rank == 0:
        sampler.sample()        # the queue is still draining: at least one sample under load even for tiny K
    join_collectives()
    stats = parallel.finish_return_stats(gathered)
    e1.record()
    barrier()
    sampler.timed = False
    ms = e0.elapsed_time(e1)
    # kernels of this library per episode (one graph replay issues what one eager episode does) + rddl_return_stats
    launches = K * launches_per_episode + K
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = world * N * R * K / (ms * 1e-3)

    # one episode through the per-transition call (rddl_step per env-step), for comparison
    sim.reset()
    for t in range(min(3, R)):
        sim.step({k: v[t] for k, v in seq.items()})
    sim.reset()
    barrier()
    e0.record()
    for t in range(R):
        sim.step({k: v[t] for k, v in seq.items()})
    e1.record()
    barrier()
    step_api_value = N * R / (e0.elapsed_time(e1) * 1e-3)

    # dominant kernel: per-launch CUDA events on the launch stream (one more episode, same inputs)
    sim.reset()
    sim.native.profile(True)
    sim.rollout(seq, R, gamma=gamma)
    torch.cuda.synchronize()
    prof = sim.native.profile_read(sim.table.launches)
    sim.native.profile(False)
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
        steps_per_launch = R / top['count']
        abytes = algorithmic_bytes_per_env_step(name) * N * steps_per_launch
        lbytes = layout_bytes_per_env_step(name, sim, steps_per_launch) * N * steps_per_launch
        per_launch_ms = top['ms'] / top['count']
        ach = abytes / (per_launch_ms * 1e-3) / 1e9
        roofline = {'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                    'traffic': (MEASURED_TRAFFIC_BYTES_PER_ENV_STEP[name] * N * steps_per_launch
                                if name in MEASURED_TRAFFIC_BYTES_PER_ENV_STEP else None),
                    'peak_source': peak_src,
                    'kernel': ({'rddl_plane_interp': 'rddl_plane_static (NVRTC-specialised build of rddl_plane_interp)',
                                'rddl_level_interp': 'rddl_level_static_%d (NVRTC-specialised build of rddl_level_interp)' % top['launch']
                                }.get(top['name'], top['name']) if sim.specialised_kernels else top['name']),
                    'kernel_us': per_launch_ms * 1e3, 'env_steps_per_launch': steps_per_launch,
                    'kernel_share_of_rollout': top['ms'] / max(1e-9, sum(r['ms'] for r in prof)),
                    'algorithmic_bytes_per_launch': abytes,
                    'layout_bytes_per_launch': lbytes,
                    'frac_layout': lbytes / (per_launch_ms * 1e-3) / 1e9 / peak,
                    'note': 'achieved/frac use the algorithmic bytes per env-step of SURVEY 8(d) (byte-per-bool state + action '
                            'reads, next-state writes, reward, done). This backend keeps bool state bit-packed in HBM and, '
                            'in a fused multi-step launch, on chip, so it has to move fewer bytes than that: '
                            'layout_bytes_per_launch / frac_layout are the same figures for the bytes its own layout needs'}

    # end-to-end through the host-facing call, per episode: pinned host action sequence -> H2D,
    # reset + rollout, per-step rewards + returns + final state -> D2H
    h_seq = {k: v.cpu().pin_memory() for k, v in seq.items()}
    d_seq = {k: torch.empty_like(v) for k, v in seq.items()}
    h_state = {s: torch.empty_like(sim.fluent(s), device='cpu').pin_memory() for s in sim.state_names}
    h_rew = torch.empty((R, N), dtype=torch.float64).pin_memory()
    h_ret = torch.empty(N, dtype=torch.float64).pin_memory()

    def e2e_episode():
        for k in d_seq:
            d_seq[k].copy_(h_seq[k], non_blocking=True)
        sim.reset()
        ret, rew = sim.rollout(d_seq, R, gamma=gamma, keep_rewards=True)
        for s in sim.state_names:
            h_state[s].copy_(sim.fluent(s), non_blocking=True)
        h_rew.copy_(rew, non_blocking=True)
        h_ret.copy_(ret, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller owns the results after the call

    Ke = max(1, min(K, args.e2e_steps))
    e2e_episode()
    barrier()
    e0.record()
    for _ in range(Ke):
        e2e_episode()
    e1.record()
    barrier()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * N * R * Ke / (float(ms_e2e.item()) * 1e-3)
    clocks = sampler.stop() if rank == 0 else None
    h2d = act_bytes * R
    d2h = state_bytes + 8 * N * R + 8 * N

    if rank == 0:
        cpu_v, cpu_sample = cpu_baseline(name, cores=1, envs_per_core=4, steps=16) if world == 1 else (None, None)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u8/f64', 'data': 'synthetic',
            'config': {'workload': f'{workload_label(name)}, {N} envs/GPU, {R}-step random-policy rollouts',
                       'step': f'one bench step = one {R}-step rollout of the {N}-env batch (reset, {R} transitions, '
                               f'return statistics) = {N * R} env-steps per GPU',
                       'envs_per_gpu': N, 'global_envs': N * world, 'grid': n, 'rollout_len': R,
                       'api': 'rddl_rollout (R transitions per call); the per-transition rddl_step loop is step_api_value',
                       'step_api_value': step_api_value,
                       'l2': f'inputs larger than L2: the action sequence every rollout reads is '
                             f'{R * act_bytes / 1e6:.0f} MB (L2 = 126 MB); state tensors {state_bytes / 1e6:.1f} MB',
                       'parallelism': f'env-shard x{world}, return statistics all-gathered (NCCL) once per rollout, '
                                      f'on a side stream overlapping the next rollout',
                       'launches': [l['kernel'] for l in sim.table.launches if l['plan'] == 0],
                       'specialised_kernels': sim.specialised_kernels,
                       'cuda_graph': graph is not None,
                       'return_stats': stats},
            'roofline': roofline,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': float(ms_e2e.item()) / Ke, 'steps': Ke},
            'gpu_launches': launches,
            'clocks': clocks,
        }
        if cpu_v is not None:
            line['cpu_baseline'] = {'value': cpu_v, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'sample': cpu_sample}
        emit(line)
    del graph
    if world > 1:
        # the result is out: never let a stuck NCCL teardown hold the job
        watchdog = threading.Timer(30.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50, help='timed bench steps; one step = one rollout of the batch')
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='wildfire32', choices=sorted(WORKLOADS))
    ap.add_argument('--envs', type=int, default=0)
    ap.add_argument('--sync-collective', action='store_true', help='run the statistics all-gather on the main stream')
    ap.add_argument('--no-graph', action='store_true', help='issue every episode eagerly instead of replaying a CUDA graph')
    ap.add_argument('--rollout', type=int, default=0, help='env-steps per rollout (default: per workload)')
    ap.add_argument('--e2e-steps', type=int, default=10, help='rollouts timed for the end-to-end figure (<= --steps)')
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything a library writes to fd 1 meanwhile (NCCL's
    # version banner, ...) goes to stderr instead
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
