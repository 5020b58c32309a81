"""Bit-plane lowering: the fast path for launches whose whole program is boolean-grid work.

Many RDDL domains (Wildfire is the canonical one, /root/reference/pyRDDLGym/examples/
run_build.py:7-44) are cellular: every fluent of a level is a bool plane over the same
grid, the only aggregations are neighbour counts through a non-fluent adjacency mask,
and real arithmetic appears only as a function of that small count feeding
``Bernoulli(.)`` or a reward coefficient. For such a launch every value is a function
of a few *bits per cell*, so instead of interpreting ~100 scalar ops per cell the
launch is lowered to ops over packed bit-planes (32 cells per 32-bit word):

* ``Fin`` abstract values: (index planes, table) -- the value at a cell is
  ``table[bits of the index planes at that cell]``. Logical / arithmetic / relational /
  if nodes combine tables on the host with NumPy (the same float64 ufuncs the reference
  applies per element, /root/reference/pyRDDLGym/core/simulator.py:60-127), so e.g.
  ``1 / (1 + exp[4.5 - k])`` becomes a 16-entry table computed exactly as the reference
  would compute it.
* neighbour aggregations whose CSR mask is a shift-invariant 3x3 stencil become PCOUNT
  (bit-sliced carry-save adders over shifted planes staged in shared memory).
* Bernoulli(p) with p = table[idx planes] becomes PBERN: a bit-serial comparison of a
  counter-based uniform against the per-cell threshold, one Philox word per 32 cells and
  level (oracle/philox_np.py:bernoulli_bitsliced restates it); with injected variates it
  is the reference's ``U <= p`` on the same float64 values (simulator.py:1045-1053).
* reward sums of ``coefficient * bool-expression`` become per-env popcounts.

Anything outside this class raises ``NotPlanar`` and the launch falls back to the
general scalar interpreter (same results, slower).
"""
import numpy as np

from pyrddlgym_b200 import optable as ot
from pyrddlgym_b200.optable import OP, RED

MAX_M = 5            # index planes of an intermediate Fin
MAX_M_BOOL = 4       # ... of a boolean combination (a PLUT over m planes costs 2^m - 1 LOP3)
MAX_M_BERN = 4
MAX_M_RED = 3
import os as _os

# cells per CTA tile: 256 threads x WPT words x 32 cells (WPT in {1, 2, 4}; RDDL_PLANE_WPT overrides)
TILE_CELLS = 8192 * int(_os.environ.get('RDDL_PLANE_WPT', '4'))
MAX_PREGS = 32
BERN_LEVELS = 8      # bit-serial levels before the per-cell tail (csrc/rddl_plane.cuh)


class NotPlanar(Exception):
    pass


class Fin:
    """Finite-domain value: value(cell) = table[sum_i bit(idx[i], cell) << i]."""
    __slots__ = ('idx', 'table', 'care')

    def __init__(self, idx, table, care=None):
        self.idx = tuple(idx)
        self.table = np.asarray(table)
        assert self.table.shape == (1 << len(self.idx),)
        # care[v] == False: index value v cannot occur (e.g. a neighbour count above the
        # stencil size); such entries may be filled freely to simplify the device tables
        self.care = np.ones(self.table.shape, dtype=bool) if care is None else np.asarray(care, dtype=bool)

    @property
    def m(self):
        return len(self.idx)

    @property
    def cls(self):
        k = self.table.dtype.kind
        return 'b' if k == 'b' else ('f' if k == 'f' else 'i')


def _one(x):
    return x.astype(np.int64) if x.dtype == np.bool_ else x


def _expand(f, joint):
    """table of f re-indexed over the joint index plane list."""
    pos = [joint.index(r) for r in f.idx]
    v = np.arange(1 << len(joint))
    sub = np.zeros_like(v)
    for i, p in enumerate(pos):
        sub |= ((v >> p) & 1) << i
    return f.table[sub], f.care[sub]


def stencil_offsets(M, H, W):
    """M: bool [H, W, H, W]. Returns the offset set D (subset of the 3x3 window) if
    M[x,y,x2,y2] == ((x2-x, y2-y) in D) for all in-grid pairs, else None."""
    xs, ys, x2s, y2s = np.nonzero(M)
    if len(xs) == 0:
        return frozenset()
    dx, dy = x2s - xs, y2s - ys
    if np.abs(dx).max() > 1 or np.abs(dy).max() > 1:
        return None
    D = frozenset(zip(dx.tolist(), dy.tolist()))
    exp = np.zeros_like(M)
    for (ddx, ddy) in D:
        x = np.arange(max(0, -ddx), min(H, H - ddx))
        y = np.arange(max(0, -ddy), min(W, W - ddy))
        exp[x[:, None], y[None, :], (x + ddx)[:, None], (y + ddy)[None, :]] = True
    return D if np.array_equal(exp, M) else None


def band_offsets(B):
    """B: bool [n, n]; returns offsets d (|d| <= 1) if B[a, a2] == (a2 - a in set)."""
    n = B.shape[0]
    a, a2 = np.nonzero(B)
    if len(a) == 0:
        return frozenset()
    d = a2 - a
    if np.abs(d).max() > 1:
        return None
    D = frozenset(d.tolist())
    exp = np.zeros_like(B)
    for dd in D:
        i = np.arange(max(0, -dd), min(n, n - dd))
        exp[i, i + dd] = True
    return D if np.array_equal(exp, B) else None


class PlaneBuilder:
    """Builds the plane program of one launch group (list of lowering._Item)."""

    kind = 1

    def __init__(self, lw, plan, scope):
        from pyrddlgym_b200 import lowering as _lw
        self._lwmod = _lw
        self.lw = lw
        self.plan = plan
        self.scope = list(scope)
        self.extents = lw.p.shape_of(scope)
        if len(self.extents) == 2:
            self.H, self.W = self.extents
        elif len(self.extents) == 1:
            self.H, self.W = 1, self.extents[0]
        else:
            raise NotPlanar('scope rank')
        if self.W % 32 or self.W > 32768:
            raise NotPlanar('row length must be a multiple of 32')
        self.E = self.H * self.W
        self.insns = []
        self.ops = []
        self.reds = []
        self.stored = {}
        self.written = set()
        self.produced_slots = set()
        self.nregs = 0
        self.free = []
        self.loads = {}
        self.counts = {}
        self.shared_slots = 0
        self.variate_nodes = []
        if self.E <= TILE_CELLS:
            self.envs_per_tile = min(32, TILE_CELLS // self.E)
            self.tile_words = self.envs_per_tile * self.E // 32
            self.chunks = 1
        else:
            rows = max(1, TILE_CELLS // self.W)
            self.envs_per_tile = 1
            self.tile_words = rows * self.W // 32
            self.chunks = -(-self.H // rows)

    # ---- registers / instructions
    def alloc(self, n=1):
        if n == 1 and self.free:
            return self.free.pop()
        r = self.nregs
        self.nregs += n
        if self.nregs > MAX_PREGS:
            raise NotPlanar('too many plane registers')
        return r

    def ins(self, op, dst=0, a=0, b=0, c=0, imm=0, flags=0, fin=None):
        """Records one plane op over *virtual* registers; ``fin``: (Fin, kwargs of fin_const),
        serialised by finalize() once physical registers are known."""
        self.ops.append({'op': op, 'dst': dst, 'a': a, 'b': b, 'imm': imm, 'fin': fin})

    def finalize(self):
        """Liveness-based register allocation (the virtual register file lives in shared
        memory: fewer registers = more resident CTAs), then emission of the instructions."""
        # every plane load first: they have no inputs, and issued back to back their DRAM / L2
        # latencies overlap instead of each stalling the ops behind it
        # (non-fluent planes, then fluent planes: the kernel batches consecutive fluent byte planes)
        state = set(self.lw.p.next_state)
        is_state = lambda o: self.lw.tensor_names_by_id[o['imm']] in state
        ops = ([o for o in self.ops if o['op'] == 'PLD' and is_state(o)] +
               [o for o in self.ops if o['op'] == 'PLD' and not is_state(o)] +
               [o for o in self.ops if o['op'] not in ('PLD', 'PLDC')])
        # pre-packed non-fluent planes (PLDC) are staged in shared memory by TMA at the start of the tile:
        # each is picked up right before its first use, so that the copy overlaps the fluent loads and
        # the work in between (and the plane's register is live for as short as possible)
        for c in reversed([o for o in self.ops if o['op'] == 'PLDC']):
            if _os.environ.get('RDDL_PLDC_EARLY') == '1':       # experiment knob: non-fluent planes first
                ops.insert(0, c)
                continue
            first = len(ops)
            for i, o in enumerate(ops):
                used = ([o['a']] if o['op'] in ('PST', 'PSHARE') else []) + (list(o['fin'][0].idx) if o['fin'] is not None else [])
                if c['dst'] in used:
                    first = i
                    break
            ops.insert(first, c)
        self.ops = ops
        nxt = set(self.lw.p.next_state.values())
        INF = len(ops) + 1
        defs, uses = [], []
        for o in ops:
            op = o['op']
            d = [o['dst'] + k for k in range(4)] if op == 'PCOUNT' else (
                [o['dst']] if op in ('PLD', 'PLDC', 'PCONST', 'PLUT', 'PBERN') else [])
            u = [o['a']] if op in ('PST', 'PSHARE') else []
            if o['fin'] is not None:
                u += list(o['fin'][0].idx)
            defs.append(d)
            uses.append(u)
        last = {}
        for i, (d, u) in enumerate(zip(defs, uses)):
            for v in d + u:
                last[v] = max(last.get(v, -1), i)
        for i, o in enumerate(ops):
            # a state plane stored this step is re-read by next step's PLD in fused rollouts
            if o['op'] == 'PST' and self.lw.tensor_names_by_id[o['imm']] in nxt:
                last[o['a']] = INF
        phys, free, top = {}, set(), 0
        for i, (d, u) in enumerate(zip(defs, uses)):
            for v in set(u):
                if last[v] == i and v in phys:
                    free.add(phys[v])
            if len(d) == 4:
                k = 0
                while not all((k + t) in free or (k + t) >= top for t in range(4)):
                    k += 1
                for t in range(4):
                    free.discard(k + t)
                    phys[d[t]] = k + t
                top = max(top, k + 4)
            elif d:
                if last[d[0]] == INF:
                    # carried across steps: a register of its own, never shared with an earlier
                    # value (that value's next-step write would clobber the carry before its PLD)
                    r = top
                    top += 1
                elif free:
                    r = min(free)
                    free.discard(r)
                else:
                    r = top
                    top += 1
                phys[d[0]] = r
            for v in d:
                if last[v] == i:
                    free.add(phys[v])          # defined but never read
        if top > MAX_PREGS:
            raise NotPlanar('too many plane registers')
        self.nregs = top
        self.insns = []
        for o in ops:
            op = o['op']
            dst = phys.get(o['dst'], 0) if op in ('PLD', 'PLDC', 'PCONST', 'PLUT', 'PBERN', 'PCOUNT') else o['dst']
            a = phys[o['a']] if op in ('PST', 'PSHARE') else o['a']
            imm = o['imm']
            if o['fin'] is not None:
                f, kw = o['fin']
                imm = self.fin_const(Fin([phys[r] for r in f.idx], f.table, f.care), **kw)
            self.insns.append((OP[op], 0, dst, a, o['b'], 0, 0, imm))

    def fin_const(self, f, thresholds=False, values=None):
        """Serialises a Fin descriptor into the const pool; returns its index."""
        lw = self.lw
        m = f.m
        regs = 0
        for i, r in enumerate(f.idx):
            regs |= int(r) << (8 * i)
        truth = 0
        if f.cls == 'b':
            for v in range(1 << m):
                if f.table[v]:
                    truth |= 1 << v
        words = [m, regs, truth, 0, 0]
        if thresholds:
            p = np.asarray(f.table, dtype=np.float64).copy()
            if m >= 2:
                # fill impossible entries so that the upper half of the table is constant when
                # it can be: the device then evaluates an (m-1)-input multiplexer + 1 select
                half = 1 << (m - 1)
                hi_care = f.care[half:]
                if hi_care.any():
                    vals_hi = p[half:][hi_care]
                    if np.all(vals_hi == vals_hi[0]) or len(vals_hi) == 1:
                        p[half:] = np.where(hi_care, p[half:], vals_hi[0])
                if np.all(p[half:] == p[half]):
                    words[0] |= 1 << 8          # hi-const flag
            bad = always = 0
            thr = []
            for v in range(1 << m):
                pv = float(p[v])
                if not (pv >= 0.0 and pv <= 1.0):
                    bad |= 1 << v
                    pv = 0.0 if not (pv > 0.0) else 1.0
                if pv >= 1.0:
                    always |= 1 << v
                    thr.append((1 << 64) - 1)
                else:
                    thr.append(int(pv * float(1 << 64)) if pv > 0 else 0)
            words[3], words[4] = bad, always
            words += thr
            words += [int(x) for x in p.view(np.uint64)]
            # truth table of threshold bit (63 - j) over the index domain, for the first levels
            for j in range(BERN_LEVELS):
                words.append(sum(((t >> (63 - j)) & 1) << v for v, t in enumerate(thr)))
        elif values is not None:
            vals = np.ascontiguousarray(values)
            words += [0] * (1 << m)
            words += [int(x) for x in vals.view(np.uint64)]
        start = len(lw.consts)
        lw.consts.extend(int(w) & 0xFFFFFFFFFFFFFFFF for w in words)   # raw, not deduplicated
        return start

    # ---- Fin algebra
    def combine(self, fins, fn):
        joint = []
        for f in fins:
            for r in f.idx:
                if r not in joint:
                    joint.append(r)
        all_bool = all(f.cls == 'b' for f in fins)
        if len(joint) > (MAX_M_BOOL if all_bool else MAX_M):
            # materialise boolean operands (largest first) until the joint domain fits
            order = sorted(range(len(fins)), key=lambda i: -fins[i].m)
            fins = list(fins)
            for i in order:
                if fins[i].cls == 'b' and fins[i].m > 1:
                    fins[i] = self.as_plane_fin(fins[i])
                    return self.combine(fins, fn)
            raise NotPlanar('value depends on more than %d bits per cell' % MAX_M)
        exp = [_expand(f, joint) for f in fins]
        care = np.ones(1 << len(joint), dtype=bool)
        for (_, c) in exp:
            care &= c
        with np.errstate(all='ignore'):
            out = fn(*[t for (t, _) in exp])
        return Fin(joint, np.asarray(out), care)

    def materialise(self, f):
        """Returns a plane register holding the boolean Fin f."""
        if f.cls != 'b':
            raise NotPlanar('non-boolean value where a plane is needed')
        if f.m == 1 and bool(f.table[1]) and not bool(f.table[0]):
            return f.idx[0]
        d = self.alloc()
        if f.m == 0:
            self.ins('PCONST', dst=d, imm=int(bool(f.table[0])))
        else:
            self.ins('PLUT', dst=d, fin=(f, {}))
        return d

    def as_plane_fin(self, f):
        r = self.materialise(f)
        return Fin((r,), np.array([False, True]))

    # ---- HIR -> Fin
    def identity_index(self, n, base=0):
        rank = len(self.scope)
        return [tuple(x) for x in n.index] == [('a', base + i) for i in range(rank)]

    def load_plane(self, name):
        lw = self.lw
        if name in self.stored:
            return self.stored[name]
        if name in self.loads:
            return self.loads[name]
        decl = lw.p.tensors[name]
        if decl.dtype != 'b' or list(decl.params) != self.scope:
            raise NotPlanar(f'{name} is not a bool plane over the launch scope')
        r = self.alloc()
        if decl.kind == 'nonfluent':
            bits = np.packbits(np.asarray(lw.p.init[name], dtype=bool).reshape(-1), bitorder='little')
            off = lw.pool_add(bits.view('<u4').astype('<u4').view('<i4'), align=4)
            self.ins('PLDC', dst=r, imm=off)
        else:
            self.ins('PLD', dst=r, imm=lw.tensor_id[name])
        f = Fin((r,), np.array([False, True]))
        self.loads[name] = f
        self.load_tensor_of_reg = getattr(self, 'load_tensor_of_reg', {})
        if decl.kind != 'nonfluent':
            self.load_tensor_of_reg[r] = lw.tensor_id[name]
        return f

    def ev(self, n, base=0):
        """base: axis offset (rank) when evaluating an aggregation body at the reduced axes."""
        k = n.kind
        lw = self.lw
        if k == 'const':
            return Fin((), np.array([n.value], dtype={'b': np.bool_, 'i': np.int64, 'f': np.float64}[n.dtype]))
        if k == 'load':
            name = lw.rename.get(n.tensor, n.tensor)
            decl = lw.p.tensors[name]
            if decl.kind == 'nonfluent' and not n.index:
                v = np.asarray(lw.p.init[name]).reshape(())
                return Fin((), np.array([v[()]]))
            if not self.identity_index(n, base):
                raise NotPlanar(f'gather of {name}')
            return self.load_plane(name)
        if k == 'un':
            a = self.ev(n.args[0], base)
            op = n.op
            fns = {'not': np.logical_not, 'neg': lambda x: -1 * x, 'abs': lambda x: np.abs(_one(x)),
                   'sgn': lambda x: np.sign(_one(x)).astype(np.int64),
                   'round': lambda x: np.round(_one(x)).astype(np.int64),
                   'floor': lambda x: np.floor(_one(x)).astype(np.int64),
                   'ceil': lambda x: np.ceil(_one(x)).astype(np.int64),
                   'ln': lambda x: np.log(_one(x)), 'exp': lambda x: np.exp(_one(x)),
                   'sqrt': lambda x: np.sqrt(_one(x)), 'cos': lambda x: np.cos(_one(x)),
                   'sin': lambda x: np.sin(_one(x)), 'tan': lambda x: np.tan(_one(x)),
                   'tanh': lambda x: np.tanh(_one(x)), 'cosh': lambda x: np.cosh(_one(x)),
                   'sinh': lambda x: np.sinh(_one(x)), 'atan': lambda x: np.arctan(_one(x)),
                   'acos': lambda x: np.arccos(_one(x)), 'asin': lambda x: np.arcsin(_one(x))}
            if op not in fns:
                raise NotPlanar(f'unary {op}')
            return self.combine([a], fns[op])
        if k in ('bin', 'nary'):
            args = [self.ev(a, base) for a in n.args]
            op = n.op
            logic = {'and': np.logical_and, 'or': np.logical_or, 'xor': np.logical_xor,
                     'imply': lambda a, b: np.logical_or(np.logical_not(a), b), 'eqv': np.equal}
            arith = {'add': np.add, 'sub': np.subtract, 'mul': np.multiply, 'div': np.divide,
                     'lt': np.less, 'le': np.less_equal, 'gt': np.greater, 'ge': np.greater_equal,
                     'eq': np.equal, 'ne': np.not_equal, 'min': np.minimum, 'max': np.maximum,
                     'pow': np.power, 'fmod': np.mod, 'hypot': np.hypot,
                     'floordiv': lambda a, b: np.floor_divide(a, b).astype(np.int64),
                     'mod': lambda a, b: np.mod(a, b).astype(np.int64),
                     'log': lambda a, b: np.log(a) / np.log(b)}
            if op in logic:
                fn = logic[op]
                wrap = fn
            elif op in arith:
                fn = arith[op]
                wrap = lambda *xs: fn(*[_one(x) for x in xs])
            else:
                raise NotPlanar(f'binary {op}')
            out = args[0]
            for b in args[1:]:
                out = self.combine([out, b], wrap)
            return out
        if k == 'if':
            c, a, b = [self.ev(x, base) for x in n.args]
            return self.combine([c, a, b], lambda cc, aa, bb: np.where(cc, aa, bb))
        if k == 'delta':
            return self.ev(n.args[0], base)
        if k == 'agg':
            return self.ev_agg(n, base)
        if k == 'rand':
            return self.ev_rand(n, base)
        raise NotPlanar(f'node kind {k}')

    def ev_agg(self, n, base):
        if base != 0 or n.op not in ('sum', 'exists') or len(n.scope) != len(self.scope):
            raise NotPlanar('aggregation shape')
        lw = self.lw
        rank = len(self.scope)
        if len(n.red) != rank or list(lw.p.shape_of(n.args[0].scope)[rank:]) != list(self.extents):
            raise NotPlanar('aggregation does not range over the grid')
        loops, body = lw.plan_loops(n)
        D = self.stencil_of(loops, rank)
        src = self.ev(body, base=rank)
        if src.cls != 'b':
            raise NotPlanar('aggregation body is not boolean')
        reg = self.materialise(src)
        key = (reg, D)
        if key not in self.counts:
            mask = 0
            for (dx, dy) in D:
                mask |= 1 << ((dx + 1) * 3 + (dy + 1))
            slot = self.shared_slots
            self.shared_slots += 1
            if slot >= 4:
                raise NotPlanar('more than 4 stencil source planes')
            tensor = getattr(self, 'load_tensor_of_reg', {}).get(reg, None)
            if self.chunks > 1 and tensor is None:
                raise NotPlanar('stencil over a computed plane needs whole envs per tile')
            self.ins('PSHARE', a=reg, b=slot, imm=0xFFFFFFFF if tensor is None else tensor)
            k0 = self.alloc(4)
            self.ins('PCOUNT', dst=k0, b=slot, imm=mask)
            self.counts[key] = (k0, k0 + 1, k0 + 2, k0 + 3)
        planes = self.counts[key]
        if n.op == 'sum':
            return Fin(planes, np.arange(16, dtype=np.int64), np.arange(16) <= len(D))
        return Fin(planes, np.arange(16) > 0, np.arange(16) <= len(D))

    def stencil_of(self, loops, rank):
        lw = self.lw
        H, W = self.H, self.W
        if any(lp[0] != 'csr' for lp in loops):
            raise NotPlanar('dense aggregation')
        csr = [lw.csr_masks[lp[1]] for lp in loops]
        if rank == 2 and len(csr) == 1:
            axes, row_axes, M = csr[0]
            if list(axes) != [2, 3] or list(row_axes) != [0, 1]:
                raise NotPlanar('mask axes')
            D = stencil_offsets(M.reshape(H, W, H, W), H, W)
        elif rank == 2 and len(csr) == 2:
            (a0, r0, M0), (a1, r1, M1) = csr
            if list(a0) != [2] or list(r0) != [0] or list(a1) != [3] or list(r1) != [1]:
                raise NotPlanar('mask axes')
            Dx, Dy = band_offsets(M0.reshape(H, H)), band_offsets(M1.reshape(W, W))
            D = None if (Dx is None or Dy is None) else frozenset((dx, dy) for dx in Dx for dy in Dy)
        elif rank == 1 and len(csr) == 1:
            axes, row_axes, M = csr[0]
            if list(axes) != [1] or list(row_axes) != [0]:
                raise NotPlanar('mask axes')
            Dy = band_offsets(M.reshape(W, W))
            D = None if Dy is None else frozenset((0, dy) for dy in Dy)
        else:
            raise NotPlanar('mask structure')
        if D is None:
            raise NotPlanar('neighbour mask is not a 3x3 shift-invariant stencil')
        return D

    def ev_rand(self, n, base):
        if n.op != 'Bernoulli' or base != 0 or len(n.scope) != len(self.scope):
            raise NotPlanar(f'distribution {n.op}')
        p = self.ev(n.args[0], base)
        if p.cls == 'b':
            p = Fin(p.idx, p.table.astype(np.float64), p.care)
        p = Fin(p.idx, np.asarray(p.table, dtype=np.float64), p.care)
        slot = self.lw.variate_slot(n.id, 'U', self.E)
        self.variate_nodes.append(int(n.id))
        return self.bernoulli(p, slot)

    def bernoulli(self, p, slot):
        """PBERN of a probability table; tables over more than MAX_M_BERN planes are split on
        their last index plane x: Bern(p) = x ? Bern(p | x=1) : Bern(p | x=0). Both halves read the
        same Philox words (same node, same counters), so the split does not change the draw."""
        if p.m > MAX_M_BERN:
            if p.m > MAX_M_BERN + 3:
                raise NotPlanar('Bernoulli parameter depends on too many bits')
            half = 1 << (p.m - 1)
            x = p.idx[-1]
            b0 = self.bernoulli(Fin(p.idx[:-1], p.table[:half], p.care[:half]), slot)
            b1 = self.bernoulli(Fin(p.idx[:-1], p.table[half:], p.care[half:]), slot)
            sel = Fin((x,), np.array([False, True]))
            return self.as_plane_fin(self.combine([sel, b1, b0], lambda c, a, b: np.where(c, a, b)))
        d = self.alloc()
        self.ins('PBERN', dst=d, a=slot, fin=(p, {'thresholds': True}))
        return Fin((d,), np.array([False, True]))

    # ---- sinks
    def add_item(self, item):
        lw = self.lw
        if item.scope != self.scope:
            raise NotPlanar('scope')
        f = self.ev(item.expr)
        if item.store is not None:
            decl = lw.p.tensors[item.store]
            if decl.dtype != 'b' or list(decl.params) != self.scope or item.store_off:
                raise NotPlanar('store target')
            r = self.materialise(f)
            self.ins('PST', a=r, imm=lw.tensor_id[item.store])
            self.stored[item.store] = Fin((r,), np.array([False, True]))
            self.written.add(item.store)
        else:
            slot, rop = item.red
            if rop in (RED['AND'], RED['OR']) and f.cls == 'b':
                raise NotPlanar('forall/exists plane reductions')      # TODO popcount compare
            if rop not in (RED['SUM_I'], RED['SUM_F']):
                raise NotPlanar('reduction op')
            if f.m > 1:
                # value = c * [boolean expression]: reduce the popcount of the materialised plane
                tab = np.asarray(_one(f.table))
                nz = tab[tab != 0]
                if len(nz) and not np.all(nz == nz[0]):
                    raise NotPlanar('reduction value is not coefficient * boolean')
                c = nz[0] if len(nz) else tab.dtype.type(0)
                r = self.materialise(Fin(f.idx, tab != 0))
                f = Fin((r,), np.array([tab.dtype.type(0), c]))
            if rop == RED['SUM_F']:
                vals = np.asarray(_one(f.table), dtype=np.float64)
            else:
                if f.cls == 'f':
                    raise NotPlanar('int reduction of real value')
                vals = np.asarray(_one(f.table), dtype=np.int64)
            if len(self.reds) >= ot.MAX_RED:
                raise NotPlanar('too many reductions')
            self.ins('PRED', b=len(self.reds), fin=(f, {'values': vals}))
            self.reds.append((rop, slot))
            self.produced_slots.add(slot)
