"""TEST INFRASTRUCTURE ONLY -- a slow, scalar Python emulator of the op-table bytecode.

It exists so that the lowering (pyrddlgym_b200/lowering.py) can be checked against the
oracle on machines without a GPU. It mirrors pyrddlgym_b200/csrc/rddl_vm.cu instruction
by instruction (same decode, same reduction slots, same Philox counters) but is never
used by the product: pyrddlgym_b200 does not import it and has no CPU execution path.
"""
import math

import numpy as np

from oracle import philox_np
from oracle.oracle_sim import lngamma
from pyrddlgym_b200 import optable as ot

_OPNAME = {v: k for k, v in ot.OP.items()}
_NPDT = {0: np.uint8, 1: np.int64, 2: np.float64}


def _sections(blob):
    hdr = np.frombuffer(blob, dtype='<i8', count=16)
    assert hdr[0] == ot.MAGIC and hdr[1] == ot.VERSION
    dts = [ot.TENSOR, ot.LAUNCH, ot.INSN, np.dtype('<u8'), ot.ADDR, ot.CSR, np.dtype('<i4'), ot.REDSLOT,
           ot.VARIATE]
    off = 128
    out = []
    for i, dt in enumerate(dts):
        n = int(hdr[2 + i])
        out.append(np.frombuffer(blob, dtype=dt, count=n, offset=off))
        b = n * dt.itemsize
        off += (b + 7) & ~7
    return out


def _f(bits):
    return np.array([bits], dtype='<u8').view('<f8')[0]


def _bits_f(v):
    return int(np.array([v], dtype='<f8').view('<u8')[0])


def _i(bits):
    return int(np.array([bits], dtype='<u8').view('<i8')[0])


def _bits_i(v):
    return int(np.array([int(v) & 0xFFFFFFFFFFFFFFFF], dtype='<u8')[0])


_RED_ID = {0: 0, 1: 0.0, 2: 1, 3: 1.0, 4: np.iinfo(np.int64).max, 5: math.inf,
           6: np.iinfo(np.int64).min, 7: -math.inf, 8: 1, 9: 0}


def _red_comb(op, a, b):
    if op in (0, 1):
        return a + b
    if op in (2, 3):
        return a * b
    if op == 4:
        return min(a, b)
    if op == 6:
        return max(a, b)
    if op == 5:
        return math.nan if (math.isnan(a) or math.isnan(b)) else min(a, b)
    if op == 7:
        return math.nan if (math.isnan(a) or math.isnan(b)) else max(a, b)
    if op == 8:
        return a & b
    return a | b


class Emulator:

    def __init__(self, optable):
        (self.tensors, self.launches, self.insns, self.consts, self.addrs, self.csrs, self.pool,
         self.redslots, self.variates) = _sections(optable.blob)
        self.meta = optable

    def run(self, plan, bufs, n_envs, status, injected=None, seed=0, step=0, env_offset=0):
        """bufs: tensor id -> flat numpy array (shared: [nelem]; per-env: [n_envs*nelem])."""
        injected = injected or {}
        red = {}
        for L in self.launches:
            if int(L['plan']) != plan:
                continue
            self._launch(L, bufs, n_envs, status, injected, seed, step, env_offset, red)

    def _addr(self, ad, idx, regs, status, env):
        off = int(ad['c0'])
        for i in range(int(ad['naxis'])):
            off += int(ad['coef'][i]) * idx[int(ad['axis'][i])]
        for i in range(int(ad['nreg'])):
            v = _i(regs[int(ad['reg'][i])])
            ext = int(ad['rextent'][i])
            if v < 0 or v >= ext:
                status[env] |= 8
                v = 0 if v < 0 else ext - 1
            off += v * int(ad['rstride'][i])
        return off

    # ---- bit-plane launches (mirrors csrc/rddl_plane.cu at the level of whole bool planes)
    def _fin(self, ci):
        m = int(self.consts[ci]) & 0xFF
        regs = int(self.consts[ci + 1])
        idx = [(regs >> (8 * i)) & 0xFF for i in range(m)]
        truth = int(self.consts[ci + 2])
        bad, always = int(self.consts[ci + 3]), int(self.consts[ci + 4])
        thr = np.array(self.consts[ci + 5:ci + 5 + (1 << m)], dtype=np.uint64)
        vals = np.array(self.consts[ci + 5 + (1 << m):ci + 5 + 2 * (1 << m)], dtype=np.uint64)
        return m, idx, truth, bad, always, thr, vals

    def _plane_launch(self, L, bufs, N, status, injected, seed, step, env_offset, red):
        E = int(L['E'])
        rank = int(L['rank'])
        H, W = (int(L['extent'][0]), int(L['extent'][1])) if rank == 2 else (1, int(L['extent'][0]))
        chunks = int(L['chunks'])
        tile_words = int(L['ipt'])
        rows_per_tile = (tile_words * 32) // W
        P = {}
        shared = {}

        def index_of(idx):
            v = np.zeros((N, E), dtype=np.int64)
            for i, r in enumerate(idx):
                v |= P[r].astype(np.int64) << i
            return v

        for ins in self.insns[int(L['pc_begin']):int(L['pc_end'])]:
            op = _OPNAME[int(ins['op'])]
            dst, a, b, imm = int(ins['dst']), int(ins['a']), int(ins['b']), int(ins['imm'])
            if op == 'PLD':
                P[dst] = bufs[imm].reshape(N, E) != 0
            elif op == 'PLDC':
                words = self.pool[imm:imm + E // 32].view('<u4')
                bits = np.unpackbits(words.view(np.uint8), bitorder='little').astype(bool)
                P[dst] = np.broadcast_to(bits, (N, E))
            elif op == 'PST':
                bufs[imm][:] = P[a].reshape(-1).astype(bufs[imm].dtype)
            elif op == 'PCONST':
                P[dst] = np.full((N, E), bool(imm))
            elif op == 'PLUT':
                m, idx, truth, *_ = self._fin(imm)
                tt = np.array([(truth >> v) & 1 for v in range(1 << m)], dtype=bool)
                P[dst] = tt[index_of(idx)]
            elif op == 'PSHARE':
                shared[b] = P[a]
            elif op == 'PCOUNT':
                src = shared[b].reshape(N, H, W).astype(np.int64)
                pad = np.pad(src, ((0, 0), (1, 1), (1, 1)))
                cnt = np.zeros((N, H, W), dtype=np.int64)
                for dx in (-1, 0, 1):
                    for dy in (-1, 0, 1):
                        if imm >> ((dx + 1) * 3 + (dy + 1)) & 1:
                            cnt += pad[:, 1 + dx:1 + dx + H, 1 + dy:1 + dy + W]
                cnt = cnt.reshape(N, E)
                for i in range(4):
                    P[dst + i] = ((cnt >> i) & 1).astype(bool)
            elif op == 'PBERN':
                m, idx, truth, bad, always, thr, vals = self._fin(imm)
                v = index_of(idx)
                pvals = vals.view(np.float64)[v]
                badv = np.array([(bad >> k) & 1 for k in range(1 << m)], dtype=bool)[v]
                status[badv.any(axis=1)] |= 1
                inj = injected.get(a, None)
                if inj is not None:
                    P[dst] = np.asarray(inj, dtype=np.float64).reshape(N, E) <= pvals
                else:
                    node = int(self.variates[a]['node'])
                    env_ids = env_offset + np.arange(N)
                    P[dst] = philox_np.bernoulli_bitsliced(seed, env_ids, step, node, pvals)
            elif op == 'PRED':
                m, idx, truth, bad, always, thr, vals = self._fin(imm)
                rop = int(L['red_op'][b])
                v = index_of(idx).reshape(N, H, W)
                for env in range(N):
                    parts = []
                    for c in range(chunks):
                        blk = v[env, c * rows_per_tile:(c + 1) * rows_per_tile] if chunks > 1 else v[env]
                        cnt = np.bincount(blk.reshape(-1), minlength=1 << m)
                        if rop == 1:
                            acc = np.float64(0.0)
                            for k in range(1 << m):
                                acc = acc + vals.view(np.float64)[k] * np.float64(cnt[k])
                            parts.append(float(acc))
                        else:
                            parts.append(int((vals.view(np.int64) * cnt).sum()))
                    red[(int(L['red_slot'][b]), env)] = parts
            else:
                raise NotImplementedError(op)

    def _launch(self, L, bufs, N, status, injected, seed, step, env_offset, red):
        if int(L['pad']) == 1:
            return self._plane_launch(L, bufs, N, status, injected, seed, step, env_offset, red)
        if int(L['pad']) == 3:      # dense FP64 contraction out[e, n] = sum_k x[e, k] W(k, n)
            Nn, K = int(L['extent'][0]), int(L['extent'][1])
            w, x, o = [int(v) for v in L['red_slot'][:3]]
            sk, sn = int(L['red_op'][0]), int(L['red_op'][1])
            W = np.lib.stride_tricks.as_strided(bufs[w], shape=(K, Nn), strides=(8 * sk, 8 * sn))
            bufs[o][:] = (bufs[x].reshape(N, K) @ W).reshape(-1)
            return
        E = int(L['E'])
        rank = int(L['rank'])
        ext = [int(e) for e in L['extent'][:rank]]
        prog = self.insns[int(L['pc_begin']):int(L['pc_end'])]
        nred = int(L['nred'])
        chunks = int(L['chunks'])
        span = int(L['G']) * int(L['ipt'])
        for env in range(N):
            partial = [[_RED_ID[int(L['red_op'][j])] for _ in range(chunks)] for j in range(nred)]
            for elem in range(E):
                idx = [0] * ot.MAX_IDX
                rem = elem
                for k in range(rank - 1, -1, -1):
                    idx[k] = rem % ext[k]
                    rem //= ext[k]
                self._element(prog, L, idx, env, elem // span, bufs, status, injected, seed, step,
                              env_offset, red, partial, N)
            for j in range(nred):
                red[(int(L['red_slot'][j]), env)] = partial[j]

    def _element(self, prog, L, idx, env, chunk, bufs, status, injected, seed, step, env_offset, red,
                 partial, N):
        r = [0] * ot.MAX_REGS
        jcur, jend = [0] * 4, [0] * 4
        pc = 0
        npc = len(prog)
        with np.errstate(all='ignore'):
            while pc < npc:
                ins = prog[pc]
                op = _OPNAME[int(ins['op'])]
                dst, a, b, c, imm = int(ins['dst']), int(ins['a']), int(ins['b']), int(ins['c']), int(ins['imm'])
                pc += 1
                if op == 'NOP':
                    pass
                elif op == 'CONST':
                    r[dst] = int(self.consts[imm])
                elif op == 'AXIS':
                    r[dst] = _bits_i(idx[a])
                elif op == 'MOV':
                    r[dst] = r[a]
                elif op == 'I2F':
                    r[dst] = _bits_f(float(_i(r[a])))
                elif op == 'F2I':
                    r[dst] = _bits_i(int(_f(r[a])))
                elif op in ('LOAD', 'STORE'):
                    ad = self.addrs[imm]
                    off = self._addr(ad, idx, r, status, env) + env * int(ad['env_stride'])
                    t = int(ad['tensor'])
                    dt = int(self.tensors[t]['dtype'])
                    buf = bufs[t]
                    if op == 'LOAD':
                        if dt in (0, 3):
                            r[dst] = int(buf[off] != 0)
                        else:
                            r[dst] = int(buf.view('<u8')[off])
                    else:
                        if dt in (0, 3):
                            buf[off] = 1 if r[a] != 0 else 0
                        else:
                            buf.view('<u8')[off] = r[a]
                elif op in ('RANDU', 'RANDN', 'RANDE', 'RANDC'):
                    ad = self.addrs[imm]
                    off = self._addr(ad, idx, r, status, env)
                    inj = injected.get(a, None)
                    if inj is not None:
                        v = float(np.asarray(inj).reshape(-1)[env * int(ad['env_stride']) + off])
                    else:
                        node = int(self.variates[a]['node'])
                        nelem = int(self.variates[a]['nelem'])
                        kind = 'UNEC'[int(self.variates[a]['kind'])]
                        allv = philox_np.base_variates(kind, seed, [env_offset + env], step, node, (nelem,))
                        v = float(allv[0, off])
                    r[dst] = _bits_f(v)
                elif op == 'REDOUT':
                    rop = int(L['red_op'][b])
                    v = _f(r[a]) if rop in (1, 3, 5, 7) else _i(r[a])
                    partial[b][chunk] = _red_comb(rop, partial[b][chunk], v)
                elif op == 'REDIN':
                    rop = int(self.redslots[imm]['op'])
                    vals = red[(imm, env)]
                    acc = vals[0]
                    for v in vals[1:]:
                        acc = _red_comb(rop, acc, v)
                    r[dst] = _bits_f(acc) if rop in (1, 3, 5, 7) else _bits_i(acc)
                elif op == 'SAMPLE':
                    # same distributions as csrc/rddl_samplers.cuh, NumPy's generators (other stream)
                    ad = self.addrs[imm]
                    off = self._addr(ad, idx, r, status, env)
                    node = int(self.variates[c]['node'])
                    rng = np.random.default_rng([seed & 0xFFFFFFFF, env_offset + env, step, node, off])
                    dist = int(ins['flags'])
                    fa, fb = float(_f(r[a])), float(_f(r[b]))
                    if dist == 0:
                        v = _bits_i(rng.poisson(max(fa, 0.0)))
                    elif dist == 1:
                        v = _bits_f(fb * rng.standard_gamma(fa if fa > 0 else 1.0))
                    elif dist == 2:
                        v = _bits_i(rng.binomial(max(_i(r[a]), 0), min(max(fb, 0.0), 1.0)))
                    elif dist == 3:
                        v = _bits_i(rng.negative_binomial(fa if fa > 0 else 1.0, min(max(fb, 1e-12), 1.0)))
                    elif dist == 4:
                        v = _bits_f(rng.beta(fa if fa > 0 else 1.0, fb if fb > 0 else 1.0))
                    elif dist == 5:
                        v = _bits_i(rng.geometric(min(max(fa, 1e-12), 1.0)))
                    elif dist == 6:
                        v = _bits_f(rng.standard_t(fa if fa > 0 else 1.0))
                    else:
                        v = _bits_f(rng.chisquare(fa if fa > 0 else 1.0))
                    r[dst] = v
                elif op == 'SUMLOOP':
                    ins2 = prog[pc]
                    pc += 1
                    arg = int(ins2['dst']) | (int(ins2['a']) << 16)
                    is_csr, two, flt = b & 1, b & 2, b & 4
                    terms = []
                    for ai in ([imm, int(ins2['imm'])] if two else [imm]):
                        ad = self.addrs[ai]
                        off = int(ad['c0']) + env * int(ad['env_stride'])
                        st = 0
                        for i in range(int(ad['naxis'])):
                            if int(ad['axis'][i]) == a:
                                st = int(ad['coef'][i])
                            else:
                                off += int(ad['coef'][i]) * idx[int(ad['axis'][i])]
                        t = int(ad['tensor'])
                        terms.append((bufs[t], int(self.tensors[t]['dtype']), off, st))
                    if is_csr:
                        cs = self.csrs[arg]
                        row = int(cs['row_c0'])
                        for i in range(int(cs['row_naxis'])):
                            row += int(cs['row_coef'][i]) * idx[int(cs['row_axis'][i])]
                        j0 = int(self.pool[int(cs['rowptr_off']) + row])
                        j1 = int(self.pool[int(cs['rowptr_off']) + row + 1])
                        ks = [int(self.pool[int(cs['col_off'][0]) + j]) for j in range(j0, j1)]
                    else:
                        ks = range(arg)
                    facc, iacc = np.float64(0.0), 0
                    for k in ks:
                        vals = []
                        for (buf, dt, off, st) in terms:
                            v = buf[off + st * k]
                            vals.append(np.float64(v) if dt == 2 else int(v != 0) if dt in (0, 3) else int(v))
                        if flt:
                            fv = [np.float64(x) for x in vals]
                            facc = facc + (fv[0] * fv[1] if two else fv[0])
                        else:
                            iacc += vals[0] * vals[1] if two else vals[0]
                    r[dst] = _bits_f(facc) if flt else _bits_i(iacc)
                elif op == 'STATUS':
                    if r[a]:
                        status[env] |= imm
                elif op == 'LOOP':
                    idx[a] = 0
                    if imm == 0:
                        pc = b
                elif op == 'ENDLOOP':
                    idx[a] += 1
                    if idx[a] < imm:
                        pc = b
                elif op in ('CSRLOOP', 'CSREND'):
                    cs = self.csrs[imm]
                    if op == 'CSRLOOP':
                        row = int(cs['row_c0'])
                        for i in range(int(cs['row_naxis'])):
                            row += int(cs['row_coef'][i]) * idx[int(cs['row_axis'][i])]
                        j = int(self.pool[int(cs['rowptr_off']) + row])
                        jend[a] = int(self.pool[int(cs['rowptr_off']) + row + 1])
                        jcur[a] = j
                        if j >= jend[a]:
                            pc = b
                            continue
                    else:
                        jcur[a] += 1
                        j = jcur[a]
                        if j >= jend[a]:
                            continue
                        pc = b
                    for cc in range(int(cs['ncols'])):
                        idx[int(cs['slot'][cc])] = int(self.pool[int(cs['col_off'][cc]) + j])
                elif op == 'SEL':
                    r[dst] = r[b] if r[a] else r[c]
                elif op in ('AND', 'OR', 'XOR'):
                    x, y = r[a], r[b]
                    r[dst] = {'AND': x & y, 'OR': x | y, 'XOR': x ^ y}[op] & 1
                elif op == 'NOT':
                    r[dst] = int(r[a] == 0)
                elif op == 'IMPLY':
                    r[dst] = (int(r[a] == 0) | r[b]) & 1
                elif op[0] == 'I' and op not in ('I2F',):
                    r[dst] = self._int_op(op, _i(r[a]), _i(r[b]))
                else:
                    r[dst] = self._float_op(op, _f(r[a]), _f(r[b]))

    @staticmethod
    def _int_op(op, x, y):
        x, y = np.int64(x), np.int64(y)
        if op == 'IADD':
            v = x + y
        elif op == 'ISUB':
            v = x - y
        elif op == 'IMUL':
            v = x * y
        elif op == 'INEG':
            v = -x
        elif op == 'IABS':
            v = abs(x)
        elif op == 'IMIN':
            v = min(x, y)
        elif op == 'IMAX':
            v = max(x, y)
        elif op == 'IPOW':
            v = np.power(x, y) if y >= 0 else 0
        elif op == 'IFLOORDIV':
            v = 0 if y == 0 else np.floor_divide(x, y)
        elif op == 'IMOD':
            v = 0 if y == 0 else np.mod(x, y)
        elif op == 'ISGN':
            v = np.sign(x)
        elif op in ('ILT', 'ILE', 'IGT', 'IGE', 'IEQ', 'INE'):
            v = {'ILT': x < y, 'ILE': x <= y, 'IGT': x > y, 'IGE': x >= y, 'IEQ': x == y, 'INE': x != y}[op]
        else:
            raise NotImplementedError(op)
        return _bits_i(int(v))

    @staticmethod
    def _float_op(op, x, y):
        x, y = np.float64(x), np.float64(y)
        cmp = {'FLT': lambda: x < y, 'FLE': lambda: x <= y, 'FGT': lambda: x > y, 'FGE': lambda: x >= y,
               'FEQ': lambda: x == y, 'FNE': lambda: x != y}
        if op in cmp:
            return int(bool(cmp[op]()))
        toint = {'FSGN': np.sign, 'FROUND': np.round, 'FFLOOR': np.floor, 'FCEIL': np.ceil}
        if op in toint:
            return _bits_i(int(toint[op](x)))
        two = {'FADD': np.add, 'FSUB': np.subtract, 'FMUL': np.multiply, 'FDIV': np.divide,
               'FMIN': np.minimum, 'FMAX': np.maximum, 'FPOW': np.power, 'FMOD': np.mod,
               'FHYPOT': np.hypot}
        if op in two:
            return _bits_f(two[op](x, y))
        if op == 'FLOGB':
            return _bits_f(np.log(x) / np.log(y))
        one = {'FNEG': np.negative, 'FABS': np.abs, 'COS': np.cos, 'SIN': np.sin, 'TAN': np.tan,
               'ACOS': np.arccos, 'ASIN': np.arcsin, 'ATAN': np.arctan, 'COSH': np.cosh, 'SINH': np.sinh,
               'TANH': np.tanh, 'EXP': np.exp, 'LN': np.log, 'SQRT': np.sqrt, 'EXPM1': np.expm1,
               'LOG1P': np.log1p, 'LNGAMMA': lngamma, 'GAMMA': lambda v: np.exp(lngamma(v))}
        if op in one:
            return _bits_f(float(one[op](x)))
        raise NotImplementedError(op)


class EmulatedSimulator:
    """Drives the emulator with the same buffer protocol as the CUDA binding (state/next
    pointer swap after each step)."""

    def __init__(self, program, optable, n_envs, seed=0, env_offset=0):
        self.p = program
        self.ot = optable
        self.N = n_envs
        self.seed = seed
        self.env_offset = env_offset
        self.em = Emulator(optable)
        self.step_count = 0      # Philox step index: not rewound by reset (BatchedSimulator.reset)
        self.reset()

    def reset(self):
        self.bufs = {}
        for name, tid in self.ot.tensor_id.items():
            dt, kind, nelem = self.ot.tensor_meta[name]
            npdt = {'b': np.uint8, 'i': np.int64, 'f': np.float64}[dt]
            if name in self.p.init:
                v = np.asarray(self.p.init[name]).astype(npdt).reshape(-1)
                self.bufs[tid] = v.copy() if kind == ot.KIND_SHARED else np.tile(v, self.N)
            else:
                self.bufs[tid] = np.zeros(self.N * nelem, dtype=npdt)
        self.status = np.zeros(self.N, dtype=np.uint32)
        self.em.run(ot.PLAN_TERMINAL, self.bufs, self.N, self.status, None, self.seed, self.step_count, self.env_offset)
        return self.fluent('__done').astype(bool).reshape(-1)

    def fluent(self, name):
        tid = self.ot.tensor_id[name]
        dt, kind, nelem = self.ot.tensor_meta[name]
        shape = self.p.tensor_shape(name) if name in self.p.tensors else (nelem,)
        v = self.bufs[tid]
        if kind == ot.KIND_SHARED:
            return v.reshape(shape)
        return v.reshape((self.N,) + tuple(shape))

    def set_actions(self, actions):
        for name, v in actions.items():
            tid = self.ot.tensor_id[name]
            self.bufs[tid][:] = np.asarray(v).astype(self.bufs[tid].dtype).reshape(-1)

    def step(self, actions, variates=None):
        self.set_actions(actions)
        inj = {}
        if variates:
            for slot, (node, _, _) in enumerate(self.ot.variates):
                if node in variates:
                    inj[slot] = np.asarray(variates[node], dtype=np.float64)
        self.em.run(ot.PLAN_STEP, self.bufs, self.N, self.status, inj, self.seed, self.step_count,
                    self.env_offset)
        for s, ns in self.p.next_state.items():
            a, b = self.ot.tensor_id[s], self.ot.tensor_id[ns]
            self.bufs[a], self.bufs[b] = self.bufs[b], self.bufs[a]
        self.step_count += 1
        return self.fluent('__reward').reshape(-1).copy(), self.fluent('__done').reshape(-1).astype(bool)

    def check(self, plan, actions=None):
        if actions:
            self.set_actions(actions)
        self.em.run(plan, self.bufs, self.N, self.status, None, self.seed, self.step_count, self.env_offset)
        key = 'invariant' if plan == ot.PLAN_INVARIANT else 'precondition'
        n = self.ot.n_checks[key]
        return self.fluent('__chk')[:, :n].astype(bool).all(axis=1)
