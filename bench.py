"""bench.py -- batched env-steps/sec of the B200 RDDL stepping backend (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload wildfire1024]
    python bench.py --impl reference ...        # the reference's CPU path on the box's host cores

Default workload = the configuration BASELINE.json's target is quoted on: Wildfire 1024x1024, 4096
GLOBAL envs sharded over the N ranks (strong scaling: 4096 / N envs per GPU), random-policy rollouts.
A bench "step" is one rollout of the whole batch (BaseAgent.evaluate's episode for every env: reset,
R transitions with discounted return accumulation, return statistics).

  value     env-steps/s with the random-policy action sequence of the rollout already resident in HBM
            (explicit action planes, one byte per bool cell -- the reference boundary's layout; the
            sequence is larger than L2 and streams from DRAM every rollout)
  e2e       the same rollouts through the host-facing call a user makes (policy.evaluate's body): the
            inputs are a seed and a policy descriptor, the random actions are drawn inside the
            kernels (device-side RandomAgent), the per-env returns and the statistics are read back
            to pinned host memory every rollout
  roofline  the dominant kernel on the algorithmic bytes of SURVEY 8(d), on the bytes of this
            backend's own HBM layout, and the DRAM traffic ncu measured
  config.secondary   BASELINE configs [1] Wildfire 32x32 x 4096 envs, [2] Reservoir R=1000 x 4096 envs,
            [4] pair contraction 512x512 x 1024 envs, each timed the same way in the same run (N = 1)

Multi-GPU (torchrun): contiguous env shards, no data-path collective; one NCCL all-gather of five
numbers per rank per rollout (the return statistics), on a side stream overlapping the next rollout.
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
    # BASELINE.json configs: [3] wildfire1024 (the default: the config the target is stated on),
    # [1] wildfire32, [2] reservoir1000, [4] pair512.  envs = GLOBAL envs (sharded over the ranks);
    # rollout = transitions per bench step (BASELINE config[1]: 100-step rollouts; shorter where the
    # explicit action sequence of `value` would not fit comfortably in HBM: 8.6 GB per step at 1024^2)
    'wildfire1024': dict(kind='wildfire', n=1024, separable=True, envs=4096, rollout=4, p_bool=0.05),
    'wildfire32': dict(kind='wildfire', n=32, separable=False, envs=4096, rollout=100, p_bool=0.05),
    'wildfire32_sep': dict(kind='wildfire', n=32, separable=True, envs=4096, rollout=100, p_bool=0.05),
    'reservoir1000': dict(kind='reservoir', n=1000, envs=4096, rollout=20, bounds={'release': (0.0, 10.0)}),
    'pair512': dict(kind='pair', n=512, envs=1024, rollout=40, bounds={'u': (-1.0, 1.0)}),
}
SECONDARY = ['wildfire32', 'reservoir1000', 'pair512']

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


def make_policy(name):
    from pyrddlgym_b200.policy import RandomPolicy
    w = WORKLOADS[name]
    return RandomPolicy(bounds=w.get('bounds'), p_bool=w.get('p_bool', 0.5))


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
# (dram__bytes_read.sum + dram__bytes_write.sum; profiles/*_ncu_full_*.csv). None = not captured.
MEASURED_TRAFFIC_BYTES_PER_ENV_STEP = {
    # rddl_plane_static, fused 20-step rollout of 4096 envs (r01_ncu_full_plane_static_rollout20_32x32_4096env.csv)
    'wildfire32': (168.882688e6 + 4.095232e6) / (20 * 4096),
    # rddl_plane_static, one step of 4096 envs (r02_ncu_full_plane_static_1024x1024_4096env.csv)
    'wildfire1024': (9.690350e9 + 1.051693e9) / 4096,
    # rddl_level_fused_0, one step of 4096 envs (r02_ncu_full_level_fused_reservoir1000_4096env.csv): more than the
    # algorithmic bytes because the three interm fluents (rain, released, inflow) are written out as well
    'reservoir1000': (68.494080e6 + 81.018624e6) / 4096,
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
    """Random-policy action sequence [k, N, ...] resident on the device (drawn in slices: a 1024x1024
    plane of 4096 envs is 4.3 G cells)."""
    import torch
    w = WORKLOADS[name]
    out = {}
    for an in p.names_of_kind('action'):
        shp = (k, N) + tuple(p.tensor_shape(an))
        if p.tensors[an].dtype == 'b':
            t = torch.empty(shp, dtype=torch.bool, device=dev)
            per = max(1, (1 << 28) // max(1, int(np.prod(shp[2:], dtype=np.int64))))
            for s in range(k):
                for e0 in range(0, N, per):
                    sl = t[s, e0:e0 + per]
                    sl.copy_(torch.rand(sl.shape, device=dev, generator=g) < w.get('p_bool', 0.05))
            out[an] = t
        else:
            lo, hi = w.get('bounds', {}).get(an, (0.0, 1.0))
            out[an] = torch.rand(shp, device=dev, generator=g, dtype=torch.float64) * (hi - lo) + lo
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
# CPU arm: the unmodified reference RDDLSimulator (baseline/_ref, pip-installed copy of
# /root/reference) where it can represent the workload; the NumPy ports of oracle/ elsewhere
# ---------------------------------------------------------------------------------------

def _reference_can_run(name):
    """The reference builds every grounding name and materialises [scope x reduced] temporaries
    (SURVEY fact 8): Wildfire beyond n ~ 48 and the R^2 temporaries of the large synthetic configs are
    out of its reach in a bounded sample."""
    w = WORKLOADS[name]
    return w['kind'] == 'wildfire' and w['n'] <= 48


def _cpu_worker(args):
    """One host core: `steps` transitions of `n_envs` envs of workload `name`. Returns
    (seconds, kind): kind 'reference' = the unmodified RDDLSimulator, one env per simulator object,
    stepped one after the other (the reference has no batch axis); 'port' = oracle/ NumPy code."""
    name, n_envs, steps, seed, want_reference = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    w = WORKLOADS[name]
    rng = np.random.default_rng(seed)
    if want_reference:
        from oracle import ast_builders
        from oracle.ref_import import load_reference
        ref = load_reference()
        spec = make_spec(name)
        sims = [ref.RDDLSimulator(ast_builders.build_model(spec), rng=np.random.default_rng(seed + e), keep_tensors=True)
                for e in range(n_envs)]
        n = w['n']

        def act():
            return {'put-out': rng.random((n, n)) < w['p_bool'], 'cut-out': rng.random((n, n)) < w['p_bool']}
        for s in sims:
            s.reset()
            s.step(act())
        t0 = time.perf_counter()
        for _ in range(steps):
            for s in sims:
                s.step(act())
        return time.perf_counter() - t0, 'reference'
    if w['kind'] == 'wildfire' and w['n'] > 64:
        # grids the reference cannot build: the independent NumPy stencil restatement of the Wildfire CPFs
        from oracle.wildfire_stencil import wildfire_step
        from pyrddlgym_b200 import workloads
        n = w['n']
        target, fire = workloads.wildfire_masks(n)
        b = np.broadcast_to(fire, (n_envs, n, n)).copy()
        f = np.zeros((n_envs, n, n), dtype=bool)
        put, cut = rng.random((n_envs, n, n)) < w['p_bool'], rng.random((n_envs, n, n)) < w['p_bool']
        b, f, _ = wildfire_step(b, f, put, cut, target, rng.random((n_envs, n, n)))
        t0 = time.perf_counter()
        for _ in range(steps):
            b, f, _ = wildfire_step(b, f, put, cut, target, rng.random((n_envs, n, n)))
        return time.perf_counter() - t0, 'port'
    from pyrddlgym_b200 import workloads
    from oracle.oracle_sim import OracleSimulator
    p = workloads.load_program(make_spec(name))
    sim = OracleSimulator(p, n_envs=n_envs, seed=seed)
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
    return time.perf_counter() - t0, 'port'


def cpu_baseline(name, cores=1, envs_per_core=4, steps=8):
    """The CPU path timed on `cores` host cores (one process each, envs_per_core envs, `steps`
    transitions). Returns (env-steps/s aggregated over cores, kind, sample description)."""
    want_ref = False
    if _reference_can_run(name):
        try:
            from oracle.ref_import import reference_available
            want_ref = reference_available()
        except Exception:
            want_ref = False
    w = WORKLOADS[name]
    if w['kind'] == 'wildfire' and w['n'] > 64:
        envs_per_core, steps = 1, max(1, min(steps, 2))
    jobs = [(name, envs_per_core, steps, 100 + i, want_ref) for i in range(cores)]
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context('spawn').Pool(cores) as pool:
            res = pool.map(_cpu_worker, jobs)
    dts, kind = [r[0] for r in res], res[0][1]
    total = cores * envs_per_core * steps
    what = {'reference': 'unmodified pyRDDLGym RDDLSimulator (baseline/_ref), one simulator object per env',
            'port': ('NumPy stencil restatement of the Wildfire CPFs (oracle/wildfire_stencil.py): the reference cannot '
                     'build this grid (2^40 NEIGHBOR groundings / [X,Y,X,Y] temporaries)'
                     if (w['kind'] == 'wildfire' and w['n'] > 64) else
                     'NumPy oracle port (oracle/oracle_sim.py): the reference needs O(R^2)-grounding front-end work '
                     'beyond a bounded sample at this size')}[kind]
    return total / max(dts), kind, f'{cores} process(es) x {envs_per_core} env(s) x {steps} transitions of {workload_label(name)}; {what}'


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
    v, kind, sample = cpu_baseline(name, cores=use, envs_per_core=per_step_envs, steps=max(1, args.steps) * per_step_len)
    wall = time.perf_counter() - t0
    N, R = args.envs or w['envs'], args.rollout or w['rollout']
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * use * per_step_envs * per_step_len / v, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'u8/f64', 'data': 'synthetic',
        'config': {'workload': f'{workload_label(name)}, {N} global envs, {R}-step random-policy rollouts',
                   'step': 'bounded sample of one rollout of the batch on the host cores: ' + sample,
                   'wall_s': round(wall, 2)},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': use, 'kind': kind, 'sample': sample},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    if name == 'wildfire1024' and not args.no_secondary:
        # the largest BASELINE config the unmodified reference can run, timed beside the port
        t1 = time.perf_counter()
        v2, kind2, sample2 = cpu_baseline('wildfire32', cores=use, envs_per_core=2, steps=8)
        line['config']['secondary'] = [{'workload': workload_label('wildfire32'), 'value': v2, 'unit': UNIT,
                                        'cpu_baseline': {'value': v2, 'unit': UNIT, 'cores': use, 'kind': kind2,
                                                         'sample': sample2},
                                        'wall_s': round(time.perf_counter() - t1, 2)}]
    emit(line)


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------

def bench_workload(name, args, world, rank, dev, K, W, Ke, sampler=None, want_step_api=False):
    """Times one workload on this rank's shard; returns the dict of its figures (rank 0 keeps them)."""
    import torch
    import torch.distributed as dist
    from pyrddlgym_b200 import parallel, workloads
    from pyrddlgym_b200.simulator import BatchedSimulator

    w = WORKLOADS[name]
    n = w['n']
    G = args.envs or w['envs']                       # global envs
    N, off = parallel.shard_envs(G, world, rank)     # this rank's contiguous block
    R = args.rollout or w['rollout']                 # transitions per rollout = per bench step
    p = workloads.load_program(make_spec(name))
    # specialize=True: every kernel of the program is specialised on its first launch (during the warm-up)
    sim = BatchedSimulator(p, n_envs=N, device=dev, seed=42, env_offset=off, specialize=True)
    gamma = float(p.discount)
    esz = lambda a: 1 if p.tensors[a].dtype == 'b' else 8
    act_bytes = sum(N * int(np.prod(p.tensor_shape(a), dtype=np.int64)) * esz(a) for a in p.names_of_kind('action'))
    state_bytes = sum(N * int(np.prod(p.tensor_shape(a), dtype=np.int64)) * esz(a) for a in p.next_state)
    desc = make_policy(name).descriptor(sim)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- `value`: explicit random-policy action sequence of one rollout, resident in HBM [R, N, ...]
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    seq = draw_actions(name, p, R, N, dev, g)
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

    def episode(graph=None, local=episode_local):
        if graph is not None:
            graph.replay()
        else:
            local()
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

    def capture(local):
        # The local part of the episode is a fixed launch sequence (reset copies, the rollout kernels,
        # the statistics kernel): captured once into a CUDA graph and replayed per step, so the host's
        # per-launch latency is not part of the step. (The Philox step index of the captured rollout is
        # fixed: every replay repeats the same rollout -- same work, same inputs.)
        if args.no_graph:
            return None
        try:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                local()                      # (the collective stays outside the graph)
            g_ = episode(graph, local)
            join_collectives()
            parallel.finish_return_stats(g_)
            return graph
        except Exception as e:               # capture not possible here: issue the episodes eagerly
            sys.stderr.write('bench: CUDA graph capture failed (%s); running eagerly\n' % e)
            torch.cuda.synchronize()
            g_ = episode(None, local)        # (same number of collectives as the ranks whose capture worked)
            join_collectives()
            parallel.finish_return_stats(g_)
            return None

    graph = capture(episode_local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler is not None:
        sampler.timed = True
    e0.record()
    for _ in range(K):                                             # exactly K timed steps
        gathered = episode(graph)
    if sampler is not None:
        sampler.sample()        # the queue is still draining: at least one sample under load even for tiny K
    join_collectives()
    stats = parallel.finish_return_stats(gathered)
    e1.record()
    barrier()
    if sampler is not None:
        sampler.timed = False
    tms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = G * R * K / (ms * 1e-3)
    launches = K * launches_per_episode + K       # kernels of this library per episode + rddl_return_stats
    del graph

    # ---- the same rollouts with the bool action planes handed over BIT-PACKED (packed_actions=True: int32 words,
    # 32 cells each -- 1/8 of the action bytes of the reference boundary's layout), graph-replayed like `value`
    packed_value = None
    if w['kind'] == 'wildfire' and not args.no_packed:
        psim = BatchedSimulator(p, n_envs=N, device=dev, seed=42, env_offset=off, specialize=True, packed_actions=True)
        pseq = {k: torch.stack([BatchedSimulator.pack_bool(v[t]) for t in range(R)]) for k, v in seq.items()}
        preturns = torch.zeros(N, dtype=torch.float64, device=dev)

        def episode_packed():
            psim.reset()
            preturns.zero_()
            psim.rollout(pseq, R, gamma=gamma, returns=preturns)
            local_stats.copy_(parallel.local_return_stats(preturns))

        for _ in range(3):
            episode_packed()
        torch.cuda.synchronize()
        pgraph = None
        if not args.no_graph:
            try:
                pgraph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(pgraph):
                    episode_packed()
                pgraph.replay()
            except Exception as e:
                sys.stderr.write('bench: CUDA graph capture failed for the packed-action rollout (%s)\n' % e)
                pgraph = None
        Kp = max(1, min(K, 10))
        barrier()
        e0.record()
        for _ in range(Kp):
            if pgraph is not None:
                pgraph.replay()
            else:
                episode_packed()
        e1.record()
        barrier()
        tp = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        packed_value = G * R * Kp / (float(tp.item()) * 1e-3)
        del pgraph, pseq
        psim.close()
        del psim
        torch.cuda.empty_cache()

    # ---- one episode through the per-transition call (rddl_step per env-step), for comparison
    step_api_value = None
    if want_step_api:
        sim.reset()
        per_step = [{k: v[t] for k, v in seq.items()} for t in range(R)]     # (views, made outside the timed loop)
        for t in range(min(3, R)):
            sim.step(per_step[t])
        sim.reset()
        barrier()
        e0.record()
        for t in range(R):
            sim.step(per_step[t])
        e1.record()
        barrier()
        step_api_value = N * R / (e0.elapsed_time(e1) * 1e-3)

    # ---- dominant kernel: per-launch CUDA events on the launch stream (one more episode, same inputs)
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
    peak_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'
    roofline = None
    if top:
        steps_per_launch = R / top['count']
        abytes = algorithmic_bytes_per_env_step(name) * N * steps_per_launch
        lbytes = layout_bytes_per_env_step(name, sim, steps_per_launch) * N * steps_per_launch
        per_launch_ms = top['ms'] / top['count']
        ach = abytes / (per_launch_ms * 1e-3) / 1e9
        traffic = (MEASURED_TRAFFIC_BYTES_PER_ENV_STEP[name] * N * steps_per_launch
                   if name in MEASURED_TRAFFIC_BYTES_PER_ENV_STEP else None)
        roofline = {'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                    'traffic': traffic,
                    'peak_source': peak_src,
                    'kernel': ({'rddl_plane_interp': 'rddl_plane_static (NVRTC-specialised build of rddl_plane_interp)',
                                'rddl_level_interp': (
                                    'rddl_level_fused_%d (NVRTC-specialised: the scalar launches of the step as one env-tile kernel)'
                                    if sim.native.info(9) > 0 else
                                    'rddl_level_static_%d (NVRTC-specialised build of rddl_level_interp)') % top['launch']
                                }.get(top['name'], top['name']) if sim.specialised_kernels else top['name']),
                    'kernel_us': per_launch_ms * 1e3, 'env_steps_per_launch': steps_per_launch,
                    'kernel_share_of_rollout': top['ms'] / max(1e-9, sum(r['ms'] for r in prof)),
                    'algorithmic_bytes_per_launch': abytes,
                    'layout_bytes_per_launch': lbytes,
                    'frac_layout': lbytes / (per_launch_ms * 1e-3) / 1e9 / peak,
                    'frac_traffic': (traffic / (per_launch_ms * 1e-3) / 1e9 / peak) if traffic else None,
                    'note': 'achieved / frac: algorithmic bytes per env-step of SURVEY 8(d) (byte-per-bool state + action '
                            'reads, next-state writes, reward, done). This backend keeps bool state bit-packed in HBM (and on '
                            'chip during a fused multi-step launch), so it moves fewer bytes than that: frac_layout = the bytes '
                            'its own layout needs, frac_traffic = the DRAM bytes ncu measured (profiles/), same kernel time'}
    if w['kind'] == 'pair' and top and roofline:
        # the contraction is the tensor-core kernel of this config: report it against the tensor roofline,
        # for the exact FP64 kernel (what `value` ran) and for the tcgen05 split-bf16 variant
        flops = 2.0 * n * n * N
        tc_sim = BatchedSimulator(p, n_envs=N, device=dev, seed=42, env_offset=off, specialize=True,
                                  contract_precision='bf16x3')
        tc_sim.rollout(seq, R, gamma=gamma)
        tc_sim.reset()
        tc_sim.native.profile(True)
        tc_sim.rollout(seq, R, gamma=gamma)
        torch.cuda.synchronize()
        prof_tc = tc_sim.native.profile_read(tc_sim.table.launches)
        tc_sim.native.profile(False)
        tc_us = [1e3 * r['ms'] / r['count'] for r in prof_tc if r['name'] == 'rddl_contract_f64']
        f64_us = [1e3 * r['ms'] / r['count'] for r in prof if r['name'] == 'rddl_contract_f64']

        def graph_step_us(s):
            # whole transitions (contraction + the scalar launches behind it) replayed from a CUDA graph: the
            # per-launch events above include the host's launch latency between the two kernels of the
            # tcgen05 path, a graph does not
            try:
                s.reset()
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    s.rollout(seq, R, gamma=gamma)
                for _ in range(3):
                    gr.replay()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for _ in range(5):
                    gr.replay()
                g1.record()
                torch.cuda.synchronize()
                return 1e3 * g0.elapsed_time(g1) / (5 * R)
            except Exception:
                torch.cuda.synchronize()
                return None
        step_us = {'f64': graph_step_us(sim), 'bf16x3': graph_step_us(tc_sim)}
        bf16_peak = float(peaks.get('bf16_tflops', 1657.2))
        roofline['tensor'] = {
            'flops_per_launch': flops,
            'fp64_dmma': {'kernel': 'rddl_contract_f64 (mma.sync m8n8k4 f64)', 'kernel_us': f64_us[0] if f64_us else None,
                          'tflops': flops / (f64_us[0] * 1e-6) / 1e12 if f64_us else None},
            'tcgen05_bf16x3': {'kernel': 'rddl_tc_split + rddl_contract_tc (tcgen05.mma kind::f16, 6 bf16 products per '
                                         'f64 product, TMA operand loads, TMEM accumulators)',
                               'kernel_us': tc_us[0] if tc_us else None,
                               'tflops_effective': flops / (tc_us[0] * 1e-6) / 1e12 if tc_us else None,
                               'tflops_issued': 6 * flops / (tc_us[0] * 1e-6) / 1e12 if tc_us else None,
                               'peak_bf16_tflops': bf16_peak,
                               'frac_of_bf16_peak': 6 * flops / (tc_us[0] * 1e-6) / 1e12 / bf16_peak if tc_us else None},
            'transition_us_graph_replay': step_us,
            'note': 'selected by BatchedSimulator(contract_precision=...): f64 (default, exact products) | bf16x3 (1e-5). '
                    'kernel_us: CUDA events around the launch(es) of the contraction, issued eagerly; '
                    'transition_us_graph_replay: the whole transition (contraction + scalar launches) replayed from a CUDA graph'}
        tc_sim.close()
        del tc_sim
    del seq

    # ---- e2e: the host-facing call (policy.evaluate's body) with the random actions drawn in the
    # kernels: inputs = seed + policy descriptor, results = per-env returns + statistics, read back to
    # pinned host memory every rollout
    h_ret = torch.empty(N, dtype=torch.float64).pin_memory()
    h_stats = torch.empty(5 * world, dtype=torch.float64).pin_memory()
    ret_e = torch.zeros(N, dtype=torch.float64, device=dev)

    def e2e_local():
        sim.reset()
        ret_e.zero_()
        sim.rollout_policy(desc, R, gamma=gamma, returns=ret_e)
        local_stats.copy_(parallel.local_return_stats(ret_e))

    def e2e_episode():
        e2e_local()
        gathered = parallel.gather_return_stats(local_stats)
        h_ret.copy_(ret_e, non_blocking=True)
        h_stats.copy_(gathered.reshape(-1), non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller owns the results after the call
        return h_stats

    l0 = sim.launches_issued
    e2e_episode()
    e2e_launches = sim.launches_issued - l0
    e2e_episode()
    barrier()
    e0.record()
    for _ in range(Ke):
        hs = e2e_episode()
    e1.record()
    barrier()
    e2e_stats = parallel.finish_return_stats(hs.view(world, 5))
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_value = G * R * Ke / (float(ms_e2e.item()) * 1e-3)
    import ctypes
    from pyrddlgym_b200 import native
    h2d = ctypes.sizeof(native.Policy) + 8          # policy descriptor + seed (kernel parameters)
    d2h = 8 * N + 40 * world
    # the policy-driven rollout kernel alone (device time of one rollout, CUDA events on the stream)
    sim.reset()
    sim.native.profile(True)
    sim.rollout_policy(desc, R, gamma=gamma)
    torch.cuda.synchronize()
    prof_p = sim.native.profile_read(sim.table.launches)
    sim.native.profile(False)
    policy_us = 1e3 * sum(r['ms'] for r in prof_p)

    out = {
        'workload': f'{workload_label(name)}, {G} global envs ({N} on this GPU), {R}-step random-policy rollouts',
        'value': value, 'unit': UNIT, 'ms_per_step': ms / K, 'steps': K,
        'envs_per_gpu': N, 'global_envs': G, 'grid': n, 'rollout_len': R,
        'step': f'one bench step = one {R}-step rollout of the batch (reset, {R} transitions, return statistics) = '
                f'{G * R} env-steps',
        'l2': f'inputs larger than L2: the explicit action sequence every rollout reads is {R * act_bytes / 1e6:.0f} MB per GPU '
              f'(L2 = 126 MB); state tensors {state_bytes / 1e6:.1f} MB as bytes',
        'launches': [l['kernel'] for l in sim.table.launches if l['plan'] == 0],
        'specialised_kernels': sim.specialised_kernels,
        'cuda_graph': not args.no_graph,
        'return_stats': stats,
        'roofline': roofline,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': float(ms_e2e.item()) / Ke, 'steps': Ke, 'launches_per_step': e2e_launches,
                'device_us_per_rollout': policy_us,
                'api': 'BatchedSimulator.reset + rollout_policy (rddl_rollout_policy: RandomAgent drawn in the kernels) + '
                       'return statistics; per-env returns and statistics copied to pinned host memory, host synchronised, '
                       'every rollout', 'return_stats': e2e_stats},
        'gpu_launches': launches + Ke * (e2e_launches + 1),
    }
    if step_api_value is not None:
        out['step_api_value'] = step_api_value
    if packed_value is not None:
        out['packed_actions_value'] = packed_value
    sim.close()
    del sim
    torch.cuda.empty_cache()
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    name = args.workload
    K, W = args.steps, args.warmup
    Ke = max(1, min(K, args.e2e_steps))
    sampler = ClockSampler(local, period=0.001)
    if rank == 0:
        sampler.start()
    main = bench_workload(name, args, world, rank, dev, K, W, Ke, sampler=sampler if rank == 0 else None,
                          want_step_api=WORKLOADS[name]['n'] <= 64)
    clocks = sampler.stop() if rank == 0 else None
    secondary = []
    if world == 1 and not args.no_secondary and name == 'wildfire1024' and not args.envs:
        for sname in SECONDARY:
            try:
                s = bench_workload(sname, args, world, rank, dev, max(3, min(K, 10)), 3, max(1, min(Ke, 5)),
                                   want_step_api=True)
                secondary.append({k: s[k] for k in ('workload', 'value', 'unit', 'ms_per_step', 'steps', 'roofline', 'e2e',
                                                    'launches', 'step_api_value', 'packed_actions_value') if k in s})
            except Exception as e:      # a secondary entry must never take the headline down
                secondary.append({'workload': workload_label(sname), 'error': repr(e)[:300]})
    if rank == 0:
        cpu = None
        if world == 1:
            try:
                v, kind, sample = cpu_baseline(name, cores=1, envs_per_core=4, steps=16)
                cpu = {'value': v, 'unit': UNIT, 'cores': 1, 'kind': kind, 'sample': sample}
                if secondary and 'error' not in secondary[0]:
                    v2, kind2, sample2 = cpu_baseline('wildfire32', cores=1, envs_per_core=2, steps=8)
                    secondary[0]['cpu_baseline'] = {'value': v2, 'unit': UNIT, 'cores': 1, 'kind': kind2, 'sample': sample2}
            except Exception as e:
                sys.stderr.write('bench: cpu baseline failed: %r\n' % (e,))
        cfg = {k: main[k] for k in ('workload', 'step', 'envs_per_gpu', 'global_envs', 'grid', 'rollout_len', 'l2', 'launches',
                                    'specialised_kernels', 'cuda_graph', 'return_stats')}
        cfg['api'] = 'value: rddl_rollout on explicit action planes resident in HBM; e2e: rddl_rollout_policy (device-side RandomAgent)'
        cfg['parallelism'] = (f'env-shard x{world} (strong scaling: {main["global_envs"]} global envs), return statistics '
                              f'all-gathered (NCCL) once per rollout on a side stream overlapping the next rollout')
        if 'step_api_value' in main:
            cfg['step_api_value'] = main['step_api_value']
        if 'packed_actions_value' in main:
            cfg['packed_actions_value'] = main['packed_actions_value']
            cfg['packed_actions_note'] = ('value with the bool action planes handed over bit-packed (BatchedSimulator(packed_actions=True): '
                                          'int32 words of 32 cells, 1/8 of the action bytes); same action sequence, bit-identical transitions '
                                          '(tests/test_cuda_parity.py::test_bit_packed_action_input_equals_byte_actions)')
        if secondary:
            cfg['secondary'] = secondary
        line = {
            'metric': METRIC, 'value': main['value'], 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': main['ms_per_step'], 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'u8/f64', 'data': 'synthetic', 'config': cfg,
            'roofline': main['roofline'], 'e2e': main['e2e'], 'gpu_launches': main['gpu_launches'], 'clocks': clocks,
        }
        if cpu is not None:
            line['cpu_baseline'] = cpu
        emit(line)
    if world > 1:
        # the result is out: never let a stuck NCCL teardown hold the job
        watchdog = threading.Timer(30.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30, help='timed bench steps; one step = one rollout of the batch')
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='wildfire1024', choices=sorted(WORKLOADS))
    ap.add_argument('--envs', type=int, default=0, help='global envs (default: per workload)')
    ap.add_argument('--sync-collective', action='store_true', help='run the statistics all-gather on the main stream')
    ap.add_argument('--no-graph', action='store_true', help='issue every episode eagerly instead of replaying a CUDA graph')
    ap.add_argument('--no-secondary', action='store_true', help='skip the secondary BASELINE configs')
    ap.add_argument('--no-packed', action='store_true', help='skip the bit-packed-action variant of the Wildfire rollouts')
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
