"""HIR -> op table (the bytecode the CUDA level interpreter executes).

Replaces the recursive tree walk of ``RDDLSimulator._sample``
(/root/reference/pyRDDLGym/core/simulator.py:508-537) and the per-CPF loop of ``step``
(:452-456) by a static schedule:

* Roots (CPFs in level order, reward, terminations on the new state) are grouped into
  *launches*. A launch is one kernel over (env x scope element); consecutive roots of the
  same scope fuse into one launch as long as every value they read from a root of the same
  launch is read at the *same* element (then it is forwarded in a register and never
  re-read from HBM). A gather across elements, or a per-env reduction produced in the same
  launch, forces a launch boundary (= grid-wide barrier).
* Scope-free aggregations (reward sums, forall/exists checks) are hoisted into per-env
  block reductions (REDOUT in the producing launch, REDIN in the consumer).
* Aggregations nested in an element scope become per-thread loops. When the body is
  ``MASK ^ fluent`` / ``MASK * fluent`` with MASK built from non-fluents only, the mask is
  evaluated once on the host and turned into CSR neighbour lists (joint over all reduced
  axes, or one list per axis when the mask is separable), so the loop visits only the
  non-zeros -- the reference materialises the dense [scope x reduced] temporary instead
  (simulator.py:713-733,766-779).
* Evaluation is always-evaluate: both branches of if/switch are computed and selected,
  which gives the same values as the reference's data-dependent short-circuit
  (simulator.py:624-637,713-733,903-923) for valid programs.
"""
import hashlib
import math

import numpy as np

from pyrddlgym_b200 import hir, optable as ot
from pyrddlgym_b200.hir import Node
from pyrddlgym_b200.optable import OP, RED

_ORDER = 'bif'
_HOIST_MIN = 16
_CSR_DENSE_LIMIT = 1 << 27


def _bank_schedule(rowptr, cols):
    """Order of the entries inside each row of a padded CSR list (the order in which every float SUMLOOP adds
    them). The fused env-tile kernels gather the terms from a vector staged in shared memory as
    [element][4 envs]: the bank group an entry hits is ``col % 4``, and the four rows a half warp sums in
    lock-step (rows 4 b .. 4 b + 3) conflict at a position whenever two of their entries share a group. The
    four rows of a block are therefore ordered together: at every position the assignment of distinct groups
    to rows that serves the most rows (and, among those, leaves the fullest groups for later) is taken; a
    row that cannot be served takes an entry of its fullest group (ncu, Reservoir R = 1000: 3.45
    shared-memory wavefronts per 8-byte gather in index order, 1.77 without conflicts)."""
    import itertools
    perms = list(itertools.permutations(range(4)))
    out = np.array(cols, dtype='<i4', copy=True)
    nrows = len(rowptr) - 1
    for b0 in range(0, nrows, 4):
        rows = list(range(b0, min(b0 + 4, nrows)))
        groups = []
        for r in rows:
            row = out[int(rowptr[r]):int(rowptr[r + 1])]
            gs = [list(row[row % 4 == g]) for g in range(4)]
            for g in gs:
                g.reverse()                      # pop() takes entries in index order
            groups.append(gs)
        orders = [[] for _ in rows]
        left = [int(rowptr[r + 1] - rowptr[r]) for r in rows]
        while any(left):
            best, best_score = None, None
            for pm in perms:
                served = fill = 0
                for i in range(len(rows)):
                    if left[i] and groups[i][pm[i]]:
                        served += 1
                        fill += len(groups[i][pm[i]])
                score = (served, fill)
                if best_score is None or score > best_score:
                    best, best_score = pm, score
            for i in range(len(rows)):
                if not left[i]:
                    continue
                g = groups[i][best[i]]
                if not g:
                    g = max(groups[i], key=len)
                orders[i].append(g.pop())
                left[i] -= 1
        for i, r in enumerate(rows):
            out[int(rowptr[r]):int(rowptr[r + 1])] = orders[i]
    return out


def _promote(a, b):
    return _ORDER[max(_ORDER.index(a), _ORDER.index(b))]


def _num(t):
    return 'i' if t == 'b' else t


class LoweringError(NotImplementedError):
    pass


# ---------------------------------------------------------------------------
# host-side evaluation of non-fluent-only boolean masks
# ---------------------------------------------------------------------------

def _axes_used(n, acc=None):
    acc = set() if acc is None else acc
    if n.kind == 'axis':
        acc.add(n.value)
    if n.kind == 'load':
        for (k, v) in n.index:
            if k == 'a':
                acc.add(v)
    for c in n.children():
        _axes_used(c, acc)
    return acc


class _MaskEval:
    """NumPy evaluation of an expression over non-fluents, broadcast over a fixed scope
    (size-1 axes where the expression does not depend on the axis)."""

    def __init__(self, program, scope):
        self.p = program
        self.ext = program.shape_of(scope)
        self.rank = len(scope)

    def supported(self, n):
        if n.kind in ('const', 'axis'):
            return True
        if n.kind == 'load':
            return (self.p.tensors[n.tensor].kind == 'nonfluent'
                    and all(k in ('a', 'c') for (k, _) in n.index))
        if n.kind in ('un', 'bin', 'nary'):
            if n.kind == 'un' and n.op not in ('not', 'neg', 'abs'):
                return False
            if n.kind == 'bin' and n.op not in ('and', 'or', 'xor', 'imply', 'eqv', 'lt', 'le', 'gt',
                                                'ge', 'eq', 'ne', 'add', 'sub', 'mul'):
                return False
            return all(self.supported(a) for a in n.args)
        return False

    def _ax(self, k):
        shp = [1] * self.rank
        shp[k] = self.ext[k]
        return np.arange(self.ext[k], dtype=np.int64).reshape(shp)

    def ev(self, n):
        if n.kind == 'const':
            return np.asarray(n.value, dtype=hir.NP_DTYPE[n.dtype]).reshape((1,) * self.rank)
        if n.kind == 'axis':
            return self._ax(n.value)
        if n.kind == 'load':
            V = np.asarray(self.p.init[n.tensor])
            if not n.index:
                return V.reshape((1,) * self.rank)
            idx = [self._ax(v) if k == 'a' else np.asarray(v).reshape((1,) * self.rank)
                   for (k, v) in n.index]
            return V[tuple(idx)]
        args = [self.ev(a) for a in n.args]
        one = lambda x: x.astype(np.int64) if x.dtype == np.bool_ else x
        op = n.op
        if op == 'not':
            return ~args[0]
        if op == 'neg':
            return -one(args[0])
        if op == 'abs':
            return np.abs(one(args[0]))
        if n.kind == 'nary':
            out = args[0]
            for a in args[1:]:
                out = {'and': np.logical_and, 'or': np.logical_or,
                       'add': lambda x, y: one(x) + one(y),
                       'mul': lambda x, y: one(x) * one(y)}[op](out, a)
            return out
        a, b = args
        if op in ('and', 'or', 'xor', 'imply', 'eqv'):
            return {'and': a & b, 'or': a | b, 'xor': a ^ b, 'imply': (~a) | b, 'eqv': a == b}[op]
        a, b = one(a), one(b)
        return {'lt': np.less, 'le': np.less_equal, 'gt': np.greater, 'ge': np.greater_equal,
                'eq': np.equal, 'ne': np.not_equal, 'add': np.add, 'sub': np.subtract,
                'mul': np.multiply}[op](a, b)


def _nonfluent_only(p, n):
    for x in n.walk():
        if x.kind in ('rand', 'agg', 'switch', 'delta', 'redin'):
            return False
        if x.kind == 'load' and p.tensors[x.tensor].kind != 'nonfluent':
            return False
    return True


def _flatten(n, kind_op):
    """Flattens nested binary/nary ``kind_op`` ('and' | 'mul') into a list of operands."""
    if n.kind in ('bin', 'nary') and n.op == kind_op:
        out = []
        for a in n.args:
            out += _flatten(a, kind_op)
        return out
    return [n]


# ---------------------------------------------------------------------------
# work items and launches
# ---------------------------------------------------------------------------

class _Item:
    """One schedulable root: evaluate ``expr`` over ``scope`` and sink the value."""

    def __init__(self, scope, expr, store=None, store_off=0, red=None, rename=None, contract=None):
        self.contract = contract    # (W name, x name, out name, stride_k, stride_n, N, K) or None
        self.scope = list(scope)
        self.expr = expr
        self.store = store          # tensor name or None
        self.store_off = store_off  # constant element offset (check outputs)
        self.red = red              # (slot, RED op) or None
        self.rename = rename or {}  # tensor redirection (terminations read the new state)


class _LaunchBuilder:

    def __init__(self, lw, plan, scope):
        self.lw = lw
        self.plan = plan
        self.scope = list(scope)
        self.extents = lw.p.shape_of(scope)
        self.E = int(np.prod(self.extents, dtype=np.int64)) if self.extents else 1
        self.insns = []
        self.reds = []              # (RED op, slot)
        self.stored = {}            # tensor -> (reg, cls): forwarded values of this launch
        self.memo = {}
        self.free = []
        self.nregs = 0
        self.pinned = {}            # reg -> pin count
        self.guard = None           # register: the enclosing scalar-`if` branches are taken for this env
        self.depth = 0
        self.loop_id = 0
        self.produced_slots = set()

    # ---- registers
    def alloc(self):
        if self.free:
            return self.free.pop()
        r = self.nregs
        self.nregs += 1
        if self.nregs > ot.MAX_REGS:
            raise LoweringError('expression needs more than %d virtual registers' % ot.MAX_REGS)
        return r

    def release(self, r):
        if r is not None and r not in self.pinned:
            self.free.append(r)

    def pin(self, r):
        self.pinned[r] = self.pinned.get(r, 0) + 1

    def unpin(self, r):
        c = self.pinned.get(r, 0) - 1
        if c <= 0:
            self.pinned.pop(r, None)
        else:
            self.pinned[r] = c

    # ---- instruction emission
    def ins(self, op, dst=0, a=0, b=0, c=0, imm=0, flags=0):
        self.insns.append((OP[op], flags, dst, a, b, c, 0, imm))
        return len(self.insns) - 1

    def const(self, value, cls):
        r = self.alloc()
        self.ins('CONST', dst=r, imm=self.lw.const_index(value, cls))
        return r

    def to_f(self, r, cls):
        if cls == 'f':
            return r
        d = self.alloc()
        self.ins('I2F', dst=d, a=r)
        self.release(r)
        return d

    def conv(self, r, cls, want):
        """Converts a value of class cls to class want (b -> i is free; -> f is I2F)."""
        if want == 'f':
            return self.to_f(r, cls)
        return r

    def op2(self, op, a, b):
        d = self.alloc()
        self.ins(op, dst=d, a=a, b=b)
        self.release(a)
        self.release(b)
        return d

    def op1(self, op, a):
        d = self.alloc()
        self.ins(op, dst=d, a=a)
        self.release(a)
        return d

    def sel(self, c, a, b):
        d = self.alloc()
        self.ins('SEL', dst=d, a=c, b=a, c=b)
        for r in (c, a, b):
            self.release(r)
        return d

    def flag_unless(self, ok, bit):
        bad = self.op1('NOT', ok)
        if self.guard is not None:
            # inside a branch of an `if` whose predicate is one value per env: the reference short-circuits
            # such an `if` (simulator.py:912-919), so a parameter error in the branch it does not take is
            # not an error
            g = self.keep(self.guard)
            bad = self.op2('AND', bad, g)
        self.ins('STATUS', a=bad, imm=bit)
        self.release(bad)

    def keep(self, r):
        """Returns a register holding a copy that the caller may consume while r stays live."""
        d = self.alloc()
        self.ins('MOV', dst=d, a=r)
        return d

    # ---- expression emission: returns (reg, cls)
    def emit(self, n):
        k = n.kind
        if k == 'const':
            return self.const(n.value, n.dtype), n.dtype
        if k == 'axis':
            r = self.alloc()
            self.ins('AXIS', dst=r, a=n.value)
            return r, 'i'
        if k == 'load':
            return self.emit_load(n)
        if k == 'redin':
            r = self.alloc()
            self.ins('REDIN', dst=r, imm=n.value)
            return r, n.dtype
        if k == 'un':
            return self.emit_unary(n)
        if k == 'bin':
            return self.emit_binary(n.op, n.args[0], n.args[1], n.dtype)
        if k == 'nary':
            r, cls = self.emit(n.args[0])
            for a in n.args[1:]:
                r2, c2 = self.emit(a)
                r, cls = self.binary_regs(n.op, r, cls, r2, c2)
            if n.op in ('add', 'mul'):
                cls = _num(cls)
            return r, cls
        if k == 'if':
            c, _ = self.emit(n.args[0])
            if len(n.args[0].scope) == 0:
                outer = self.guard
                self.pin(c)                      # live across both branches
                gt = self.keep(c) if outer is None else self.op2('AND', self.keep(c), self.keep(outer))
                self.pin(gt)
                self.guard = gt
                a, ca = self.emit(n.args[1])
                self.unpin(gt)
                self.release(gt)
                nc = self.op1('NOT', self.keep(c))
                ge = nc if outer is None else self.op2('AND', nc, self.keep(outer))
                self.pin(ge)
                self.guard = ge
                b, cb = self.emit(n.args[2])
                self.unpin(ge)
                self.release(ge)
                self.guard = outer
                self.unpin(c)
            else:
                a, ca = self.emit(n.args[1])
                b, cb = self.emit(n.args[2])
            t = _promote(ca, cb)
            a = self.conv(a, ca, t)
            b = self.conv(b, cb, t)
            return self.sel(c, a, b), t
        if k == 'switch':
            return self.emit_switch(n)
        if k == 'delta':
            return self.emit(n.args[0])
        if k == 'agg':
            return self.emit_agg(n)
        if k == 'rand':
            return self.emit_rand(n)
        raise LoweringError(f'HIR node kind {k!r}')

    # ---- loads / stores
    def addr_for(self, tensor_id, shape, index, env_stride, held):
        """Builds an address descriptor; nested index expressions are emitted here and
        their registers appended to ``held`` (released by the caller after the access)."""
        strides = []
        s = 1
        for e in reversed(shape):
            strides.append(s)
            s *= int(e)
        strides.reverse()
        c0 = 0
        coef = {}
        regs = []
        for (j, (kind, v)) in enumerate(index):
            if kind == 'a':
                coef[v] = coef.get(v, 0) + strides[j]
            elif kind == 'c':
                c0 += int(v) * strides[j]
            else:
                r, cls = self.emit(v)
                if cls == 'f':
                    raise LoweringError('object index expression evaluates to real')
                regs.append((r, strides[j], int(shape[j])))
                held.append(r)
        if len(regs) > 4:
            raise LoweringError('more than 4 nested object indices in one pvariable reference')
        return self.lw.addr_index(tensor_id, env_stride, c0, coef, regs)

    def is_identity(self, n):
        decl = self.lw.p.tensors[n.tensor]
        if list(decl.params) != self.scope:
            return False
        return [tuple(x) for x in n.index] == [('a', i) for i in range(len(self.scope))]

    def emit_load(self, n):
        lw = self.lw
        p = lw.p
        name = lw.rename.get(n.tensor, n.tensor)
        decl = p.tensors[name]
        if decl.kind == 'nonfluent' and not n.index:
            v = np.asarray(p.init[name]).reshape(())[()]
            return self.const(v, decl.dtype), decl.dtype
        if name in self.stored:
            probe = Node('load', tensor=name, index=n.index)
            if self.is_identity(probe):
                r, cls = self.stored[name]
                return r, cls
            raise AssertionError(f'gather of {name} inside its producing launch (scheduler bug)')
        simple = all(kd in ('a', 'c') for (kd, _) in n.index)
        key = None
        if simple and self.depth == 0:
            key = (name, tuple((kd, int(v)) for (kd, v) in n.index))
            if key in self.memo:
                return self.memo[key]
        held = []
        shape = p.tensor_shape(name)
        nelem = int(np.prod(shape, dtype=np.int64)) if shape else 1
        env_stride = 0 if decl.kind == 'nonfluent' else nelem
        ai = self.addr_for(lw.tensor_id[name], shape, n.index, env_stride, held)
        r = self.alloc()
        self.ins('LOAD', dst=r, imm=ai)
        for h in held:
            self.release(h)
        if key is not None:
            self.pin(r)
            self.memo[key] = (r, decl.dtype)
        return r, decl.dtype

    def emit_store(self, item, r, cls):
        lw = self.lw
        decl = lw.p.tensors[item.store]
        if _ORDER.index(cls) > _ORDER.index(decl.dtype):
            raise LoweringError(f'{item.store}: value class {cls} does not fit declared {decl.dtype}')
        r = self.conv(r, cls, decl.dtype)
        shape = lw.p.tensor_shape(item.store)
        nelem = lw.out.tensor_meta[item.store][2]
        if item.scope and list(decl.params) == item.scope:
            index = [('a', i) for i in range(len(shape))]
            ai = self.addr_for(lw.tensor_id[item.store], shape, index, nelem, [])
        else:
            ai = lw.addr_index(lw.tensor_id[item.store], nelem, item.store_off, {}, [])
        self.ins('STORE', a=r, imm=ai)
        if item.scope == list(decl.params) and item.store_off == 0:
            self.pin(r)
            self.stored[item.store] = (r, decl.dtype)
        else:
            self.release(r)

    # ---- operators
    def emit_unary(self, n):
        op = n.op
        r, cls = self.emit(n.args[0])
        if op == 'not':
            return self.op1('NOT', r), 'b'
        cls = _num(cls)
        if op == 'neg':
            return self.op1('FNEG' if cls == 'f' else 'INEG', r), cls
        if op == 'abs':
            return self.op1('FABS' if cls == 'f' else 'IABS', r), cls
        if op == 'sgn':
            return self.op1('FSGN' if cls == 'f' else 'ISGN', r), 'i'
        if op in ('round', 'floor', 'ceil'):
            if cls == 'i':
                return r, 'i'
            return self.op1({'round': 'FROUND', 'floor': 'FFLOOR', 'ceil': 'FCEIL'}[op], r), 'i'
        r = self.to_f(r, cls)
        return self.op1({'ln': 'LN'}.get(op, op.upper()), r), 'f'

    def emit_binary(self, op, na, nb, dtype=None):
        a, ca = self.emit(na)
        b, cb = self.emit(nb)
        return self.binary_regs(op, a, ca, b, cb)

    def binary_regs(self, op, a, ca, b, cb):
        if op in ('and', 'or', 'xor', 'imply', 'eqv'):
            return self.op2({'and': 'AND', 'or': 'OR', 'xor': 'XOR', 'imply': 'IMPLY',
                             'eqv': 'IEQ'}[op], a, b), 'b'
        ca, cb = _num(ca), _num(cb)
        t = _promote(ca, cb)
        if op in ('div', 'log', 'hypot'):
            t = 'f'
        if op in ('floordiv', 'mod') and t != 'i':
            raise LoweringError(f'{op} requires int arguments')
        a = self.conv(a, ca, t)
        b = self.conv(b, cb, t)
        f = t == 'f'
        if op in ('lt', 'le', 'gt', 'ge', 'eq', 'ne'):
            return self.op2(('F' if f else 'I') + op.upper(), a, b), 'b'
        table = {'add': ('IADD', 'FADD'), 'sub': ('ISUB', 'FSUB'), 'mul': ('IMUL', 'FMUL'),
                 'div': (None, 'FDIV'), 'floordiv': ('IFLOORDIV', None), 'mod': ('IMOD', None),
                 'fmod': ('IMOD', 'FMOD'), 'min': ('IMIN', 'FMIN'), 'max': ('IMAX', 'FMAX'),
                 'pow': ('IPOW', 'FPOW'), 'log': (None, 'FLOGB'), 'hypot': (None, 'FHYPOT')}
        return self.op2(table[op][1 if f else 0], a, b), t

    def emit_switch(self, n):
        pred, _ = self.emit(n.args[0])
        self.pin(pred)
        t = n.dtype
        present = [(k, c) for (k, c) in enumerate(n.cases) if c is not None]
        # A predicate that is one value per env: the reference short-circuits the switch only when the
        # predicate equals bool(predicate), i.e. for the first two cases (simulator.py:933-943:
        # `first_elem = bool(...)`, `all_equal = np.all(sample_pred == first_elem)`); then it evaluates
        # just that case, otherwise every case and the default. Parameter errors are reported for what
        # the reference evaluates (see flag_unless).
        scalar = len(n.args[0].scope) == 0
        outer = self.guard

        def eq_const(k):
            kk = self.const(k, 'i')
            hit = self.alloc()
            self.ins('IEQ', dst=hit, a=pred, b=kk)
            self.release(kk)
            return hit

        def case_guard(keys):
            """guard register: (pred in keys) OR (pred not in {0, 1}), AND the enclosing guard."""
            g = self.op1('NOT', self.op2('OR', eq_const(0), eq_const(1)))
            for k in keys:
                g = self.op2('OR', g, eq_const(k))
            if outer is not None:
                g = self.op2('AND', g, self.keep(outer))
            self.pin(g)
            return g

        def guarded(keys, negate, node):
            if not scalar:
                return self.emit(node)
            if negate:      # the default: selected by the short-circuit when case 0 / 1 has no expression
                keys = [k for k in (0, 1) if k >= len(n.cases) or n.cases[k] is None]
            g = case_guard(keys)
            self.guard = g
            out = self.emit(node)
            self.guard = outer
            self.unpin(g)
            self.release(g)
            return out

        if n.default is not None:
            acc, ca = guarded([k for (k, _) in present], True, n.default)
            acc = self.conv(acc, ca, t)
            rest = present
        else:
            acc, ca = guarded([present[-1][0]], False, present[-1][1])
            acc = self.conv(acc, ca, t)
            rest = present[:-1]
        for (k, c) in rest:
            v, cv = guarded([k], False, c)
            v = self.conv(v, cv, t)
            kk = self.const(k, 'i')
            hit = self.alloc()
            self.ins('IEQ', dst=hit, a=pred, b=kk)
            self.release(kk)
            acc = self.sel(hit, v, acc)
        self.unpin(pred)
        self.release(pred)
        return acc, t

    # ---- aggregations (simulator.py:766-779) as per-thread loops
    def open_loops(self, loops):
        """loops: list of ('dense', slot, extent) | ('csr', csr_id). Returns patch list."""
        opened = []
        for lp in loops:
            if lp[0] == 'dense':
                _, slot, extent = lp
                pc = self.ins('LOOP', a=slot, imm=int(extent))
                opened.append(('dense', slot, int(extent), pc))
            else:
                lid = self.loop_id
                self.loop_id += 1
                if lid >= 4:
                    raise LoweringError('more than 4 nested sparse loops')
                pc = self.ins('CSRLOOP', a=lid, imm=lp[1])
                opened.append(('csr', lid, lp[1], pc))
            self.depth += 1
        return opened

    def close_loops(self, opened):
        for lp in reversed(opened):
            self.depth -= 1
            if lp[0] == 'dense':
                _, slot, extent, pc = lp
                self.ins('ENDLOOP', a=slot, b=pc + 1, imm=extent)
            else:
                _, lid, cid, pc = lp
                self.ins('CSREND', a=lid, b=pc + 1, imm=cid)
                self.loop_id -= 1
            end = len(self.insns)
            o = list(self.insns[pc])
            o[4] = end                      # field b = pc after the loop
            self.insns[pc] = tuple(o)

    def emit_sumloop(self, n, loops, body):
        """``sum_{v} t1`` / ``sum_{v} t1 * t2`` with t1, t2 plain pvariable reads and a single
        (dense or CSR) loop becomes one SUMLOOP superinstruction: the kernel runs the
        multiply-accumulate loop natively instead of interpreting ~5 ops per iteration
        (the Reservoir inflow sum and the pair sum-product contraction)."""
        lw = self.lw
        p = lw.p
        if len(loops) != 1:
            return None
        terms = [body]
        if body.kind == 'bin' and body.op == 'mul':
            terms = list(body.args)
        if not 1 <= len(terms) <= 2:
            return None
        for t in terms:
            if t.kind != 'load' or any(k not in ('a', 'c') for (k, _) in t.index):
                return None
            name = lw.rename.get(t.tensor, t.tensor)
            if name in self.stored or (p.tensors[name].kind == 'nonfluent' and not t.index):
                return None
        cls = _num(body.dtype)
        if cls not in ('i', 'f'):
            return None
        lp = loops[0]
        if lp[0] == 'csr' and int(lw.csrs[lp[1]]['ncols']) != 1:
            return None
        ais = []
        for t in terms:
            name = lw.rename.get(t.tensor, t.tensor)
            decl = p.tensors[name]
            shape = p.tensor_shape(name)
            nelem = int(np.prod(shape, dtype=np.int64)) if shape else 1
            env_stride = 0 if decl.kind == 'nonfluent' else nelem
            ais.append(self.addr_for(lw.tensor_id[name], shape, t.index, env_stride, []))
        flags = (1 if lp[0] == 'csr' else 0) | (2 if len(terms) == 2 else 0) | (4 if cls == 'f' else 0)
        slot = lp[1] if lp[0] == 'dense' else int(lw.csrs[lp[1]]['slot'][0])
        arg = int(lp[2]) if lp[0] == 'dense' else int(lp[1])        # extent | csr id
        d = self.alloc()
        self.ins('SUMLOOP', dst=d, a=slot, b=flags, imm=ais[0])
        self.ins('NOP', dst=arg & 0xFFFF, a=(arg >> 16) & 0xFFFF, imm=ais[1] if len(ais) == 2 else 0)
        return d, cls

    def emit_agg(self, n):
        lw = self.lw
        op = n.op
        outer = len(n.scope)
        body = n.args[0]
        R = len(n.red)
        if outer + R > ot.MAX_IDX:
            raise LoweringError('scope + aggregation variables exceed %d axes' % ot.MAX_IDX)
        ext = lw.p.shape_of(body.scope)
        count = int(np.prod(ext[outer:], dtype=np.int64))
        loops, body = lw.plan_loops(n)
        bcls = body.dtype
        fused = self.emit_sumloop(n, loops, body) if op == 'sum' else None
        if fused is not None:
            return fused
        if op in ('forall', 'exists'):
            acc = self.const(op == 'forall', 'b')
            cls = 'b'
        elif op == 'sum':
            cls = _num(bcls)
            acc = self.const(0.0 if cls == 'f' else 0, cls)
        elif op == 'avg':
            cls = 'f'
            acc = self.const(0.0, 'f')
        elif op == 'prod':
            cls = _num(bcls)
            acc = self.const(1.0 if cls == 'f' else 1, cls)
        elif op in ('min', 'max', 'argmin', 'argmax'):
            cls = _num(bcls)
            big = (math.inf if cls == 'f' else np.iinfo(np.int64).max)
            lo = op in ('min', 'argmin')
            init = big if lo else (-math.inf if cls == 'f' else np.iinfo(np.int64).min)
            acc = self.const(init, cls)
        else:
            raise LoweringError(op)
        self.pin(acc)
        best_i = None
        if op in ('argmin', 'argmax'):
            best_i = self.const(0, 'i')
            self.pin(best_i)
        opened = self.open_loops(loops)
        v, cv = self.emit(body)
        f = cls == 'f'
        if op == 'forall':
            self.ins('AND', dst=acc, a=acc, b=v)
        elif op == 'exists':
            self.ins('OR', dst=acc, a=acc, b=v)
        elif op in ('sum', 'avg', 'prod'):
            v = self.conv(v, _num(cv), cls)
            self.ins({'sum': 'FADD' if f else 'IADD', 'avg': 'FADD',
                      'prod': 'FMUL' if f else 'IMUL'}[op], dst=acc, a=acc, b=v)
        elif op in ('min', 'max'):
            self.ins({'min': 'FMIN' if f else 'IMIN', 'max': 'FMAX' if f else 'IMAX'}[op],
                     dst=acc, a=acc, b=v)
        else:
            cmp = ('F' if f else 'I') + ('LT' if op == 'argmin' else 'GT')
            hit = self.alloc()
            self.ins(cmp, dst=hit, a=v, b=acc)
            self.ins('SEL', dst=acc, a=hit, b=v, c=acc)
            ax = self.alloc()
            self.ins('AXIS', dst=ax, a=outer)
            self.ins('SEL', dst=best_i, a=hit, b=ax, c=best_i)
            self.release(hit)
            self.release(ax)
        self.release(v)
        self.close_loops(opened)
        self.unpin(acc)
        if op == 'avg':
            cnt = self.const(float(count), 'f')
            acc = self.op2('FDIV', acc, cnt)
        if best_i is not None:
            self.unpin(best_i)
            self.release(acc)
            return best_i, 'i'
        return acc, cls

    # ---- random variables (simulator.py:958-1268) from base variates
    def base(self, n, kind, scope=None):
        lw = self.lw
        scope = n.scope if scope is None else scope
        shape = lw.p.shape_of(scope)
        nelem = int(np.prod(shape, dtype=np.int64)) if shape else 1
        slot = lw.variate_slot(n.id, kind, nelem)
        index = [('a', i) for i in range(len(shape))]
        ai = self.addr_for(-(1 + slot), shape, index, nelem, [])
        r = self.alloc()
        self.ins({'U': 'RANDU', 'N': 'RANDN', 'E': 'RANDE', 'C': 'RANDC'}[kind], dst=r, a=slot, imm=ai)
        return r

    def emit_f(self, n):
        r, c = self.emit(n)
        return self.to_f(r, _num(c))

    def check(self, cmp, a, b_val, bit):
        """flags ``bit`` unless (a cmp b_val); a stays live."""
        b = self.const(float(b_val), 'f')
        ok = self.alloc()
        self.ins(cmp, dst=ok, a=a, b=b)
        self.release(b)
        self.flag_unless(ok, bit)

    def emit_rand(self, n):
        op = n.op
        if op in ('Discrete', 'UnnormDiscrete', 'Discrete(p)', 'UnnormDiscrete(p)'):
            return self.emit_discrete(n)
        RANGE = ot.STATUS_OUT_OF_RANGE
        if op in hir.DIST_NATIVE:
            return self.emit_sample(n)
        if op == 'Uniform':
            lb = self.emit_f(n.args[0])
            ub = self.emit_f(n.args[1])
            ok = self.alloc()
            self.ins('FLE', dst=ok, a=lb, b=ub)
            self.flag_unless(ok, ot.STATUS_BAD_BOUNDS)
            u = self.base(n, 'U')
            w = self.alloc()
            self.ins('FSUB', dst=w, a=ub, b=lb)
            self.release(ub)
            t = self.op2('FMUL', w, u)
            return self.op2('FADD', lb, t), 'f'
        if op == 'Bernoulli':
            p = self.emit_f(n.args[0])
            self.check('FGE', p, 0.0, RANGE)
            self.check('FLE', p, 1.0, RANGE)
            u = self.base(n, 'U')
            return self.op2('FLE', u, p), 'b'
        if op == 'Normal':
            m = self.emit_f(n.args[0])
            var = self.emit_f(n.args[1])
            self.check('FGE', var, 0.0, RANGE)
            s = self.op1('SQRT', var)
            z = self.base(n, 'N')
            return self.op2('FADD', m, self.op2('FMUL', s, z)), 'f'
        if op == 'Exponential':
            s = self.emit_f(n.args[0])
            self.check('FGT', s, 0.0, RANGE)
            e = self.base(n, 'E')
            return self.op2('FMUL', s, e), 'f'
        a = self.emit_f(n.args[0])
        b = self.emit_f(n.args[1])
        if op in ('Weibull', 'Pareto', 'Gompertz', 'Kumaraswamy'):
            self.check('FGT', a, 0.0, RANGE)
        self.check('FGT', b, 0.0, RANGE)
        one = lambda: self.const(1.0, 'f')
        if op == 'Weibull':                      # scale * E ** (1/shape)
            e = self.base(n, 'E')
            inv = self.op2('FDIV', one(), a)
            return self.op2('FMUL', b, self.op2('FPOW', e, inv)), 'f'
        if op == 'Pareto':                       # scale * expm1(E / shape)
            e = self.base(n, 'E')
            return self.op2('FMUL', b, self.op1('EXPM1', self.op2('FDIV', e, a))), 'f'
        if op == 'Gumbel':                       # m - s * ln(-ln(1 - U))
            u = self.base(n, 'U')
            t = self.op1('LN', self.op2('FSUB', one(), u))
            t = self.op1('LN', self.op1('FNEG', t))
            return self.op2('FSUB', a, self.op2('FMUL', b, t)), 'f'
        if op == 'Laplace':
            u = self.base(n, 'U')
            self.pin(u)
            self.pin(a)
            self.pin(b)
            two = self.const(2.0, 'f')
            t = self.op2('FSUB', two, u)
            t2 = self.alloc()
            self.ins('FSUB', dst=t2, a=t, b=u)
            self.release(t)
            l1 = self.op1('LN', t2)
            m1 = self.alloc()
            self.ins('FMUL', dst=m1, a=b, b=l1)
            self.release(l1)
            hi = self.alloc()
            self.ins('FSUB', dst=hi, a=a, b=m1)
            self.release(m1)
            uu = self.alloc()
            self.ins('FADD', dst=uu, a=u, b=u)
            l2 = self.op1('LN', uu)
            m2 = self.alloc()
            self.ins('FMUL', dst=m2, a=b, b=l2)
            self.release(l2)
            lo = self.alloc()
            self.ins('FADD', dst=lo, a=a, b=m2)
            self.release(m2)
            half = self.const(0.5, 'f')
            c = self.alloc()
            self.ins('FGE', dst=c, a=u, b=half)
            self.release(half)
            for r in (u, a, b):
                self.unpin(r)
                self.release(r)
            return self.sel(c, hi, lo), 'f'
        if op == 'Cauchy':
            c = self.base(n, 'C')
            return self.op2('FADD', a, self.op2('FMUL', b, c)), 'f'
        if op == 'Gompertz':                     # ln(1 - log1p(-U)/shape) / scale
            u = self.base(n, 'U')
            t = self.op1('LOG1P', self.op1('FNEG', u))
            t = self.op2('FSUB', one(), self.op2('FDIV', t, a))
            return self.op2('FDIV', self.op1('LN', t), b), 'f'
        if op == 'Kumaraswamy':                  # (1 - U**(1/b)) ** (1/a)
            u = self.base(n, 'U')
            t = self.op2('FPOW', u, self.op2('FDIV', one(), b))
            t = self.op2('FSUB', one(), t)
            return self.op2('FPOW', t, self.op2('FDIV', one(), a)), 'f'
        raise LoweringError(f'distribution {op}')

    def emit_sample(self, n):
        """Poisson / Gamma / Binomial / NegativeBinomial / Beta / Geometric / Student / ChiSquare:
        parameter checks as status bits (simulator.py:1066-1221), then one SAMPLE op."""
        op = n.op
        RANGE = ot.STATUS_OUT_OF_RANGE
        arity, dt = hir.DIST_NATIVE[op]
        if op == 'Binomial':
            a, ca = self.emit(n.args[0])
            af = self.alloc()
            self.ins('I2F', dst=af, a=a)
            self.check('FGE', af, 0.0, RANGE)
            self.release(af)
        else:
            a = self.emit_f(n.args[0])
            self.check('FGE' if op == 'Poisson' else 'FGT', a, 0.0, RANGE) if op != 'Geometric' else None
        b = a
        if arity == 2:
            b = self.emit_f(n.args[1])
        if op in ('Binomial', 'NegativeBinomial'):
            self.check('FGE', b, 0.0, RANGE)
            self.check('FLE', b, 1.0, RANGE)
        elif op == 'Geometric':
            self.check('FGE', a, 0.0, RANGE)
            self.check('FLE', a, 1.0, RANGE)
        elif op in ('Gamma', 'Beta'):
            self.check('FGT', b, 0.0, RANGE)
        lw = self.lw
        shape = lw.p.shape_of(n.scope)
        nelem = int(np.prod(shape, dtype=np.int64)) if shape else 1
        slot = lw.variate_slot(n.id, 'U', nelem)
        ai = self.addr_for(-(1 + slot), shape, [('a', i) for i in range(len(shape))], nelem, [])
        d = self.alloc()
        self.ins('SAMPLE', dst=d, a=a, b=b, c=slot, imm=ai, flags=ot.SAMPLE_DIST[op])
        self.release(a)
        if b != a:
            self.release(b)
        return d, dt

    def emit_discrete(self, n):
        """simulator.py:1240-1256: cdf = cumsum(pdf) [/ cdf[-1]]; argmax(U < cdf)."""
        unnorm = n.op.startswith('Unnorm')
        lw = self.lw
        u = self.base(n, 'U')
        self.pin(u)
        res = self.const(0, 'i')
        found = self.const(False, 'b')
        cdf = self.const(0.0, 'f')
        for r in (res, found, cdf):
            self.pin(r)
        total = None

        def accumulate(pk, k_reg):
            # pk: F register with the k-th probability (consumed)
            ok = self.alloc()
            z = self.const(0.0, 'f')
            self.ins('FGE', dst=ok, a=pk, b=z)
            self.release(z)
            self.flag_unless(ok, ot.STATUS_OUT_OF_RANGE)
            self.ins('FADD', dst=cdf, a=cdf, b=pk)
            self.release(pk)
            if total is not None:
                c = self.alloc()
                self.ins('FDIV', dst=c, a=cdf, b=total)
            else:
                c = cdf
            lt = self.alloc()
            self.ins('FLT', dst=lt, a=u, b=c)
            if c is not cdf:
                self.release(c)
            nf = self.alloc()
            self.ins('NOT', dst=nf, a=found)
            hit = self.op2('AND', lt, nf)
            self.ins('SEL', dst=res, a=hit, b=k_reg, c=res)
            self.ins('OR', dst=found, a=found, b=hit)
            self.release(hit)
            self.release(k_reg)

        if n.op.endswith('(p)'):
            body = n.args[0]
            slot = len(n.scope)
            extent = lw.p.types[n.red[0]]
            if unnorm:
                total = self.const(0.0, 'f')
                self.pin(total)
                opened = self.open_loops([('dense', slot, extent)])
                pk = self.emit_f(body)
                self.ins('FADD', dst=total, a=total, b=pk)
                self.release(pk)
                self.close_loops(opened)
            opened = self.open_loops([('dense', slot, extent)])
            pk = self.emit_f(body)
            kr = self.alloc()
            self.ins('AXIS', dst=kr, a=slot)
            accumulate(pk, kr)
            self.close_loops(opened)
        else:
            ps = []
            if unnorm:
                total = self.const(0.0, 'f')
                self.pin(total)
                for c in n.cases:
                    pk = self.emit_f(c)
                    self.ins('FADD', dst=total, a=total, b=pk)
                    self.release(pk)
            for (k, c) in enumerate(n.cases):
                pk = self.emit_f(c)
                accumulate(pk, self.const(k, 'i'))
        # validity: np.allclose(cdf[-1], 1.0) -> |last - 1| <= 1e-8 + 1e-5
        last = self.alloc()
        if total is not None:
            self.ins('FDIV', dst=last, a=cdf, b=total)
        else:
            self.ins('MOV', dst=last, a=cdf)
        one = self.const(1.0, 'f')
        d = self.op1('FABS', self.op2('FSUB', last, one))
        tol = self.const(1e-8 + 1e-5, 'f')
        ok = self.op2('FLE', d, tol)
        self.flag_unless(ok, ot.STATUS_BAD_PDF)
        for r in (u, found, cdf, res) + ((total,) if total is not None else ()):
            self.unpin(r)
        for r in (u, found, cdf) + ((total,) if total is not None else ()):
            self.release(r)
        return res, 'i'

    # ---- sinks
    def add_item(self, item):
        r, cls = self.emit(item.expr)
        if item.store is not None:
            self.emit_store(item, r, cls)
        else:
            slot, rop = item.red
            want_f = rop in (RED['SUM_F'], RED['PROD_F'], RED['MIN_F'], RED['MAX_F'])
            r = self.conv(r, _num(cls), 'f' if want_f else 'i')
            if len(self.reds) >= ot.MAX_RED:
                raise AssertionError('scheduler exceeded MAX_RED')
            self.ins('REDOUT', a=r, b=len(self.reds), imm=slot)
            self.reds.append((rop, slot))
            self.produced_slots.add(slot)
            self.release(r)


# ---------------------------------------------------------------------------
# the lowering driver
# ---------------------------------------------------------------------------

class OpTable:
    """Result of lowering: the packed blob + the host-side metadata the binding needs."""

    def __init__(self):
        self.blob = b''
        self.tensor_names = []       # id -> name (program tensors then specials)
        self.tensor_id = {}
        self.tensor_meta = {}        # name -> (dtype char, kind, nelem)
        self.variates = []           # slot -> (node id, kind, nelem)
        self.launches = []           # dicts, for logging / tests
        self.n_checks = {}
        self.n_insns = 0
        self.plane_nodes = []        # ids of Bernoulli nodes lowered to the bit-sliced sampler
        self.plane_rejects = []      # (plan, scope, reason) of groups that stayed on the scalar path
        self.packed = []             # bool state fluents stored bit-packed in HBM (dtype code 3)


SPECIALS = ['__reward', '__done', '__chk']


class Lowerer:

    def __init__(self, program, use_planes=True, use_contract=True, use_packed=True, packed_actions=False):
        self.p = program
        self.use_packed = use_packed
        self.packed_actions = packed_actions
        self.use_planes = use_planes
        self.use_contract = use_contract
        self.contract_items = {}
        self.csr_masks = {}
        self._plane_ctx = None
        self.tensor_names_by_id = {}
        self.rename = {}
        self.consts = []
        self.const_map = {}
        self.addrs = []
        self.addr_map = {}
        self.csrs = []
        self.csr_map = {}
        self.pool = []
        self.pool_len = 0
        self.redslots = []
        self.variates = []
        self.variate_map = {}
        self.launch_rows = []
        self.insns = []
        self.tensor_id = {}
        self.tensor_rows = []
        self.out = OpTable()
        self._mask_cache = {}

    # ---- tables
    def const_index(self, value, cls):
        if cls == 'f':
            bits = np.array([value], dtype='<f8').view('<u8')[0]
        else:
            bits = np.array([int(value)], dtype='<i8').view('<u8')[0]
        bits = int(bits)
        if bits not in self.const_map:
            self.const_map[bits] = len(self.consts)
            self.consts.append(bits)
        return self.const_map[bits]

    def addr_index(self, tensor, env_stride, c0, coef, regs):
        axes = sorted(k for k in coef if coef[k] != 0)
        key = (tensor, env_stride, c0, tuple((k, coef[k]) for k in axes), tuple(regs))
        if key in self.addr_map:
            return self.addr_map[key]
        row = np.zeros((), dtype=ot.ADDR)
        row['tensor'] = tensor
        row['env_stride'] = env_stride
        row['c0'] = c0
        row['naxis'] = len(axes)
        for i, k in enumerate(axes):
            if k >= ot.MAX_IDX:
                raise LoweringError('axis index out of range')
            row['axis'][i] = k
            row['coef'][i] = coef[k]
        row['nreg'] = len(regs)
        for i, (r, s, e) in enumerate(regs):
            row['reg'][i] = r
            row['rstride'][i] = s
            row['rextent'][i] = e
        self.addrs.append(row)
        self.addr_map[key] = len(self.addrs) - 1
        return self.addr_map[key]

    def variate_slot(self, node_id, kind, nelem):
        if node_id is None:
            raise LoweringError('random node without id')
        if node_id in self.variate_map:
            return self.variate_map[node_id]
        self.variate_map[node_id] = len(self.variates)
        self.variates.append((int(node_id), 'UNEC'.index(kind), int(nelem)))
        return self.variate_map[node_id]

    def pool_add(self, arr, align=1):
        """Appends an int32 array to the pool; ``align`` (in words) pads the pool first (pre-packed
        non-fluent planes start on 16 bytes: the plane kernels copy them with TMA)."""
        arr = np.ascontiguousarray(arr, dtype='<i4')
        if align > 1 and self.pool_len % align:
            pad = np.zeros(align - self.pool_len % align, dtype='<i4')
            self.pool.append(pad)
            self.pool_len += len(pad)
        off = self.pool_len
        self.pool.append(arr)
        self.pool_len += len(arr)
        return off

    def new_slot(self, rop, chunks):
        self.redslots.append((rop, chunks))
        return len(self.redslots) - 1

    # ---- sparse loop planning
    def plan_loops(self, n):
        """Returns (loops, body') for aggregation node n (see module docstring)."""
        p = self.p
        outer = len(n.scope)
        body = n.args[0]
        R = len(n.red)
        ext = p.shape_of(body.scope)
        red_axes = list(range(outer, outer + R))
        dense = [('dense', a, ext[a]) for a in red_axes]
        op = n.op
        if op == 'exists' or (op == 'sum' and body.dtype == 'b'):
            parts = _flatten(body, 'and')
            mode = 'and'
        elif op == 'sum':
            parts = _flatten(body, 'mul')
            mode = 'mul'
        else:
            return dense, body
        me = _MaskEval(p, body.scope)
        masks, rest = [], []
        for q in parts:
            if (q.dtype == 'b' and _nonfluent_only(p, q) and me.supported(q)
                    and (_axes_used(q) & set(red_axes))):
                masks.append(q)
            else:
                rest.append(q)
        if not masks:
            return dense, body
        used = [(_axes_used(m), m) for m in masks]
        joint = any(len(u & set(red_axes)) > 1 for (u, _) in used)
        loops = []
        if joint:
            groups = [(red_axes, used)]
        else:
            groups = [([a], [(u, m) for (u, m) in used if a in u]) for a in red_axes]
        for (axes, ms) in groups:
            if not ms:
                loops.append(('dense', axes[0], ext[axes[0]]))
                continue
            row_axes = sorted(set().union(*[u for (u, _) in ms]) - set(axes))
            dims = [ext[a] for a in row_axes + axes]
            if int(np.prod(dims, dtype=np.int64)) > _CSR_DENSE_LIMIT:
                return dense, body
            M = None
            for (_, m) in ms:
                v = me.ev(m)
                M = v if M is None else (M & v)
            full = [ext[a] if a in row_axes + axes else 1 for a in range(len(ext))]
            M = np.broadcast_to(M, full).reshape([ext[a] for a in sorted(row_axes + axes)])
            # axes are ascending already (row axes < red axes is not guaranteed): reorder
            order = sorted(row_axes + axes)
            perm = [order.index(a) for a in row_axes + axes]
            M = np.transpose(M, perm)
            nrows = int(np.prod([ext[a] for a in row_axes], dtype=np.int64)) if row_axes else 1
            M2 = M.reshape(nrows, -1)
            rows, flat = np.nonzero(M2)
            rowptr = np.zeros(nrows + 1, dtype=np.int64)
            np.add.at(rowptr, rows + 1, 1)
            rowptr = np.cumsum(rowptr)
            cols = np.unravel_index(flat, [ext[a] for a in axes])
            ci = self.csr_index(rowptr, cols, axes, row_axes, ext)
            self.csr_masks.setdefault(ci, (tuple(axes), tuple(row_axes), M2))
            loops.append(('csr', ci))
        # rebuild the body from the remaining operands
        if mode == 'and':
            if not rest:
                newb = Node('const', value=True, dtype='b', scope=body.scope)
            else:
                newb = rest[0]
                for q in rest[1:]:
                    newb = Node('bin', op='and', args=[newb, q], dtype='b', scope=body.scope)
        else:
            if not rest:
                newb = Node('const', value=1, dtype='i', scope=body.scope)
            else:
                newb = rest[0]
                dt = _num(newb.dtype)
                for q in rest[1:]:
                    dt = _promote(dt, _num(q.dtype))
                    newb = Node('bin', op='mul', args=[newb, q], dtype=dt, scope=body.scope)
                if len(rest) == 1 and newb.dtype == 'b':
                    one = Node('const', value=1, dtype='i', scope=body.scope)
                    newb = Node('bin', op='mul', args=[one, newb], dtype='i', scope=body.scope)
        return loops, newb

    def csr_index(self, rowptr, cols, axes, row_axes, ext):
        h = hashlib.sha1()
        h.update(rowptr.tobytes())
        for c in cols:
            h.update(np.ascontiguousarray(c).tobytes())
        key = (h.hexdigest(), tuple(axes), tuple(row_axes))
        if key in self.csr_map:
            return self.csr_map[key]
        if len(axes) > 4:
            raise LoweringError('sparse aggregation over more than 4 variables')
        if rowptr[-1] >= 2 ** 31:
            raise LoweringError('sparse aggregation mask too large')
        row = np.zeros((), dtype=ot.CSR)
        row['rowptr_off'] = self.pool_add(rowptr)
        row['ncols'] = len(axes)
        for i, (a, c) in enumerate(zip(axes, cols)):
            row['col_off'][i] = self.pool_add(c)
            row['slot'][i] = a
        if len(axes) == 1 and len(rowptr) > 1:
            # padded copy for the sparse sums (rddl_device.cuh, SUMLOOP): rows start on 16-byte boundaries and
            # hold a multiple of four int32 (or eight uint16) entries, the filler being the extent of the
            # reduced axis -- one past the last element, where the staged vector keeps a zero
            rp = np.asarray(rowptr, dtype=np.int64)
            n = np.diff(rp)
            sent = int(ext[axes[0]])
            u16 = sent <= 0xffff                  # eight 16-bit indices per 128-bit load instead of four
            per = 8 if u16 else 4
            n4 = (n + per - 1) // per * per
            rp4 = np.concatenate([[0], np.cumsum(n4)])
            if rp4[-1] < 2 ** 31:
                c4 = np.full(int(rp4[-1]), sent, dtype='<u2' if u16 else '<i4')
                src = np.asarray(cols[0], dtype='<i4')
                pos = np.repeat(rp4[:-1] - rp[:-1], n) + np.arange(len(src))
                c4[pos] = _bank_schedule(rp, src) if len(n) <= 2048 else src      # (the fused env-tile kernels take scopes up to 1408)
                row['rowptr4_off'] = self.pool_add(rp4.astype('<i4'), align=4)
                row['col4_off'] = self.pool_add(c4.view('<i4'), align=4)
                row['sentinel4'] = sent
                row['col4_u16'] = 1 if u16 else 0
                if row['rowptr4_off'] == 0:       # (0 means absent: an empty pool gets a filler first)
                    raise LoweringError('internal: padded list at pool offset 0')
        s = 1
        coefs = []
        for a in reversed(row_axes):
            coefs.append((a, s))
            s *= ext[a]
        coefs.sort()
        row['row_naxis'] = len(coefs)
        for i, (a, c) in enumerate(coefs):
            row['row_axis'][i] = a
            row['row_coef'][i] = c
        self.csrs.append(row)
        self.csr_map[key] = len(self.csrs) - 1
        return self.csr_map[key]

    # ---- hoisting of scope-free aggregations into per-env block reductions
    def geometry(self, E, has_red):
        if E <= 256:
            G = 1
            while G < E:
                G *= 2
            return G, 1, 1
        ipt = 4 if (has_red and E >= 1024) else 1
        return 256, ipt, -(-E // (256 * ipt))

    def hoist(self, expr, items):
        """Replaces scope-free aggregations by REDIN nodes; appends the producing items."""
        p = self.p

        def rec(n):
            if n.kind == 'agg' and not n.scope and n.op in ('sum', 'avg', 'prod', 'min', 'max',
                                                             'forall', 'exists'):
                body = n.args[0]
                ext = p.shape_of(body.scope)
                E = int(np.prod(ext, dtype=np.int64))
                if E >= _HOIST_MIN:
                    body = rec(body)
                    bt = _num(body.dtype)
                    if n.op in ('forall', 'exists'):
                        rop, dt = RED['AND' if n.op == 'forall' else 'OR'], 'b'
                    elif n.op == 'avg':
                        rop, dt = RED['SUM_F'], 'f'
                    else:
                        key = {'sum': 'SUM', 'prod': 'PROD', 'min': 'MIN', 'max': 'MAX'}[n.op]
                        rop, dt = RED[key + ('_F' if bt == 'f' else '_I')], bt
                    _, _, chunks = self.geometry(E, True)
                    slot = self.new_slot(rop, chunks)
                    items.append(_Item(body.scope, body, red=(slot, rop)))
                    rin = Node('redin', value=slot, dtype=dt, scope=[])
                    if n.op == 'avg':
                        cnt = Node('const', value=float(E), dtype='f', scope=[])
                        return Node('bin', op='div', args=[rin, cnt], dtype='f', scope=[])
                    return rin
            if n.kind in ('const', 'axis', 'redin'):
                return n
            m = Node(n.kind, op=n.op, args=[rec(a) for a in n.args], dtype=n.dtype, scope=n.scope,
                     id=n.id, value=n.value, tensor=n.tensor, red=n.red, obj_type=n.obj_type)
            if n.index is not None:
                m.index = [(k, (rec(v) if k == 'e' else v)) for (k, v) in n.index]
            if n.cases is not None:
                m.cases = [None if c is None else rec(c) for c in n.cases]
            if n.default is not None:
                m.default = rec(n.default)
            return m

        return rec(expr)

    # ---- scheduling
    def reads(self, expr):
        """(tensor name, node) for every fluent load, and redin slots."""
        loads, slots = [], []
        for x in expr.walk():
            if x.kind == 'load':
                loads.append((self.rename.get(x.tensor, x.tensor), x))
            elif x.kind == 'redin':
                slots.append(x.value)
        return loads, slots

    def group_items(self, items):
        """Splits the item list into launch groups (pure dependency analysis, no emission)."""
        groups = []
        cur = None
        for it in items:
            if it.contract is not None:
                groups.append({'scope': list(it.scope), 'items': [it], 'written': {it.contract[2]},
                               'stored': set(), 'slots': set(), 'nred': 0, 'contract': True})
                cur = None
                continue
            self.rename = it.rename
            loads, slots = self.reads(it.expr)
            fits = cur is not None and cur['scope'] == it.scope
            if fits:
                for (name, node) in loads:
                    if name in cur['written']:
                        # a value produced in this launch may only be read at the same element
                        decl = self.p.tensors[name]
                        ident = (list(decl.params) == it.scope and
                                 [tuple(x) for x in node.index] == [('a', i) for i in range(len(it.scope))])
                        if name not in cur['stored'] or not ident:
                            fits = False
                            break
                if any(s in cur['slots'] for s in slots):
                    fits = False
                if it.red is not None and cur['nred'] >= ot.MAX_RED:
                    fits = False
            if not fits:
                cur = {'scope': list(it.scope), 'items': [], 'written': set(), 'stored': set(),
                       'slots': set(), 'nred': 0}
                groups.append(cur)
            cur['items'].append(it)
            if it.store is not None:
                cur['written'].add(it.store)
                decl = self.p.tensors[it.store]
                if it.scope == list(decl.params) and it.store_off == 0:
                    cur['stored'].add(it.store)
            if it.red is not None:
                cur['slots'].add(it.red[0])
                cur['nred'] += 1
        self.rename = {}
        return groups

    def schedule(self, plan, items):
        from pyrddlgym_b200 import plane_lowering
        for grp in self.group_items(items):
            if grp.get('contract'):
                self.finish_contract(plan, grp['items'][0])
                continue
            lb = None
            if self.use_planes and len(grp['scope']) in (1, 2):
                mark = (len(self.consts), len(self.pool), self.pool_len)
                try:
                    lb = plane_lowering.PlaneBuilder(self, plan, grp['scope'])
                    for it in grp['items']:
                        self.rename = it.rename
                        lb.add_item(it)
                    lb.finalize()
                except plane_lowering.NotPlanar as e:
                    self.out.plane_rejects.append((plan, list(grp['scope']), str(e)))
                    # table entries (consts, CSR lists) added by the failed attempt stay: they are
                    # shared with / deduplicated against the scalar lowering of the same group
                    lb = None
            if lb is None:
                lb = _LaunchBuilder(self, plan, grp['scope'])
                for it in grp['items']:
                    self.rename = it.rename
                    lb.add_item(it)
            self.finish_launch(lb)
        self.rename = {}

    def finish_contract(self, plan, it):
        (w, x, out, sk, sn, N, K) = it.contract
        row = np.zeros((), dtype=ot.LAUNCH)
        row['plan'] = plan
        row['rank'] = 1
        row['extent'][0], row['extent'][1] = N, K
        row['E'] = N
        row['pc_begin'] = row['pc_end'] = len(self.insns)
        row['G'], row['ipt'], row['chunks'] = 1, 1, 1
        row['red_slot'][0], row['red_slot'][1], row['red_slot'][2] = (self.tensor_id[w], self.tensor_id[x],
                                                                      self.tensor_id[out])
        row['red_op'][0], row['red_op'][1] = sk, sn
        row['nregs'] = 1
        row['pad'] = 3
        self._plane_ctx = None
        self.launch_rows.append(row)
        self.out.launches.append({'plan': plan, 'scope': it.scope, 'E': N, 'kind': 3, 'kernel': 'rddl_contract_f64',
                                  'n_insns': 0, 'nregs': 0, 'stores': [out], 'nred': 0, 'K': K})

    def finish_launch(self, lb):
        row = np.zeros((), dtype=ot.LAUNCH)
        row['plan'] = lb.plan
        row['rank'] = len(lb.scope)
        for i, e in enumerate(lb.extents):
            row['extent'][i] = e
        row['E'] = lb.E
        row['pc_begin'] = len(self.insns)
        self.insns.extend(lb.insns)
        row['pc_end'] = len(self.insns)
        row['nred'] = len(lb.reds)
        kind = getattr(lb, 'kind', 0)
        if kind == 1:
            G, ipt, chunks = lb.envs_per_tile, lb.tile_words, lb.chunks
        else:
            G, ipt, chunks = self.geometry(lb.E, bool(lb.reds))
        row['G'], row['ipt'], row['chunks'] = G, ipt, chunks
        # a scope-free program that only consumes reductions of the preceding single-tile plane
        # launch runs in that kernel's epilogue (kind 2): one launch per step instead of two
        if kind == 1:
            self._plane_ctx = None
            if chunks == 1:
                self._plane_ctx = (lb.plan, set(lb.produced_slots),
                                   {self.tensor_id[n] for n in lb.written})
        elif self._plane_ctx is not None and self._plane_ctx[0] == lb.plan and lb.E == 1 \
                and not lb.reds and lb.nregs <= 40:
            redin = {i[7] for i in lb.insns if i[0] == OP['REDIN']}
            touched = {int(self.addrs[i[7]]['tensor']) for i in lb.insns if i[0] in (OP['LOAD'], OP['STORE'])}
            if redin <= self._plane_ctx[1] and not (touched & self._plane_ctx[2]):
                kind = 2
            else:
                self._plane_ctx = None
        else:
            self._plane_ctx = None
        row['pad'] = kind
        for i, (rop, slot) in enumerate(lb.reds):
            row['red_op'][i] = rop
            row['red_slot'][i] = slot
            self.redslots[slot] = (self.redslots[slot][0], chunks)
        row['nregs'] = max(lb.nregs, 1)
        self.launch_rows.append(row)
        self.out.launches.append({'plan': lb.plan, 'scope': lb.scope, 'E': lb.E, 'kind': kind,
                                  'kernel': {0: 'rddl_level_interp', 1: 'rddl_plane_interp',
                                             2: 'fused into rddl_plane_interp'}[kind],
                                  'n_insns': len(lb.insns), 'nregs': lb.nregs,
                                  'stores': sorted(lb.stored), 'nred': len(lb.reds)})
        if kind == 1:
            self.out.plane_nodes.extend(lb.variate_nodes)

    # ---- top level
    # ---- dense two-operand contractions  out(?b) = sum_{?a} W(?a, ?b) * x(?a)
    def _extract_contractions(self):
        """Rewrites the program (on a copy): every dense real contraction over a shared object
        axis with a non-fluent matrix operand is computed by the FP64 tensor-core kernel
        (csrc/rddl_contract.cuh) into a scratch fluent that the CPF then reads at the same element.
        Small contractions stay on the scalar SUMLOOP path."""
        import copy
        src = self.p
        p = copy.copy(src)
        p.tensors = dict(src.tensors)
        p.init = dict(src.init)
        self.contract_items = {}       # root index -> [_Item]
        counter = [0]

        def match(n):
            if n.kind != 'agg' or n.op != 'sum' or len(n.scope) != 1 or len(n.red) != 1:
                return None
            body = n.args[0]
            if body.kind != 'bin' or body.op != 'mul' or body.dtype != 'f':
                return None
            ops = list(body.args)
            if any(o.kind != 'load' or o.dtype != 'f' for o in ops):
                return None
            for (w, x) in (ops, ops[::-1]):
                dw, dx = p.tensors.get(w.tensor), p.tensors.get(x.tensor)
                if dw is None or dx is None or dw.kind != 'nonfluent' or dx.kind == 'nonfluent':
                    continue
                if [tuple(i) for i in x.index] != [('a', 1)]:
                    continue
                wi = [tuple(i) for i in w.index]
                if sorted(wi) != [('a', 0), ('a', 1)]:
                    continue
                N = p.types[n.scope[0]]
                K = p.types[n.red[0]]
                if N < 32 or K < 32:
                    return None
                if wi == [('a', 1), ('a', 0)]:      # W[a][b]: k-major
                    sk, sn = N, 1
                else:                                # W[b][a]
                    sk, sn = 1, K
                return (w.tensor, x.tensor, sk, sn, N, K)
            return None

        def rec(n, root_index):
            m = match(n)
            if m is not None:
                name = '__ct%d' % counter[0]
                counter[0] += 1
                p.tensors[name] = hir.TensorDecl(name, 'interm', 'f', [n.scope[0]], 'real')
                p.init[name] = np.zeros(m[4], dtype=np.float64)
                self.contract_items.setdefault(root_index, []).append(
                    _Item(n.scope, None, contract=(m[0], m[1], name, m[2], m[3], m[4], m[5])))
                return Node('load', tensor=name, index=[('a', 0)], dtype='f', scope=n.scope, id=n.id)
            if n.kind in ('const', 'axis', 'redin'):
                return n
            q = Node(n.kind, op=n.op, args=[rec(a, root_index) for a in n.args], dtype=n.dtype, scope=n.scope,
                     id=n.id, value=n.value, tensor=n.tensor, red=n.red, obj_type=n.obj_type)
            if n.index is not None:
                q.index = [(k, (rec(v, root_index) if k == 'e' else v)) for (k, v) in n.index]
            if n.cases is not None:
                q.cases = [None if c is None else rec(c, root_index) for c in n.cases]
            if n.default is not None:
                q.default = rec(n.default, root_index)
            return q

        if self.use_contract:
            p.cpfs = [hir.Root(r.name, r.level, r.scope, rec(r.expr, i)) for i, r in enumerate(src.cpfs)]
        self.p = p

    def lower(self):
        self._extract_contractions()
        p = self.p
        names = list(p.tensors)
        for i, name in enumerate(names):
            d = p.tensors[name]
            shape = p.tensor_shape(name)
            nelem = int(np.prod(shape, dtype=np.int64)) if shape else 1
            self.tensor_id[name] = i
            self.tensor_names_by_id[i] = name
            kind = ot.KIND_SHARED if d.kind == 'nonfluent' else ot.KIND_ENV
            pair = names.index(p.next_state[name]) if name in p.next_state else -1
            self.tensor_rows.append((ot.DT[d.dtype], pair, kind, nelem))
            self.out.tensor_meta[name] = (d.dtype, kind, nelem)
        n_chk = max(1, len(p.invariants), len(p.preconditions), len(p.terminations))
        for (name, dt, nelem) in [('__reward', 'f', 1), ('__done', 'b', 1), ('__chk', 'b', n_chk)]:
            self.tensor_id[name] = len(self.tensor_rows)
            self.tensor_rows.append((ot.DT[dt], -1, ot.KIND_ENV, nelem))
            p.tensors[name] = hir.TensorDecl(name, 'output', dt, [], 'special')
            self.out.tensor_meta[name] = (dt, ot.KIND_ENV, nelem)
            names.append(name)
        if len(names) > ot.MAX_TENSORS:
            raise LoweringError('too many pvariables (%d > %d)' % (len(names), ot.MAX_TENSORS))
        try:
            self._lower_plans()
        finally:
            for s in SPECIALS:
                del p.tensors[s]
        self._choose_packed()
        out = self.out
        out.tensor_names = names
        out.tensor_id = dict(self.tensor_id)
        out.variates = list(self.variates)
        out.n_checks = {'invariant': len(p.invariants), 'precondition': len(p.preconditions),
                        'termination': len(p.terminations), 'width': n_chk}
        out.n_insns = len(self.insns)
        pool = np.concatenate(self.pool) if self.pool else np.zeros(0, dtype='<i4')
        out.blob = ot.pack(
            np.array(self.tensor_rows, dtype=ot.TENSOR) if self.tensor_rows else np.zeros(0, ot.TENSOR),
            np.array(self.launch_rows, dtype=ot.LAUNCH) if self.launch_rows else np.zeros(0, ot.LAUNCH),
            np.array(self.insns, dtype=ot.INSN) if self.insns else np.zeros(0, ot.INSN),
            np.array(self.consts, dtype='<u8'),
            np.array(self.addrs, dtype=ot.ADDR) if self.addrs else np.zeros(0, ot.ADDR),
            np.array(self.csrs, dtype=ot.CSR) if self.csrs else np.zeros(0, ot.CSR),
            pool,
            np.array(self.redslots, dtype=ot.REDSLOT) if self.redslots else np.zeros(0, ot.REDSLOT),
            np.array(self.variates, dtype=ot.VARIATE) if self.variates else np.zeros(0, ot.VARIATE))
        return out

    def _choose_packed(self):
        """Bool state fluents that only the plane interpreter writes are kept BIT-PACKED in HBM
        (32 cells per uint32 word; tensor dtype code 3): a step then moves 1/8 of the state bytes
        and skips the byte<->bit conversion. The byte layout of the reference boundary is
        produced on demand by the host mirror (BatchedSimulator.fluent). A state / next-state
        pair is packed together or not at all."""
        self.out.packed = []
        if not self.use_packed:
            return
        p = self.p
        plane_rw, scalar_w, other = set(), set(), set()
        for L in self.launch_rows:
            kind = int(L['pad'])
            ins = self.insns[int(L['pc_begin']):int(L['pc_end'])]
            if kind == 1:
                for i in ins:
                    if i[0] in (OP['PLD'], OP['PST']):
                        plane_rw.add(int(i[7]))
                    elif i[0] == OP['PSHARE'] and int(i[7]) != 0xFFFFFFFF:
                        plane_rw.add(int(i[7]))
            elif kind == 3:
                other.update(int(t) for t in L['red_slot'][:3])
            else:
                k = 0
                while k < len(ins):
                    i = ins[k]
                    if i[0] == OP['STORE']:
                        scalar_w.add(int(self.addrs[i[7]]['tensor']))
                    elif i[0] == OP['SUMLOOP']:
                        other.add(int(self.addrs[i[7]]['tensor']))
                        if i[4] & 2:
                            other.add(int(self.addrs[ins[k + 1][7]]['tensor']))
                        k += 1
                    k += 1
        for s_name, n_name in p.next_state.items():
            ids = (self.tensor_id[s_name], self.tensor_id[n_name])
            d = p.tensors[s_name]
            nelem = self.out.tensor_meta[s_name][2]
            if d.dtype != 'b' or nelem % 32 or any(t in scalar_w or t in other for t in ids):
                continue
            if not all(t in plane_rw for t in ids):
                continue
            for name, t in ((s_name, ids[0]), (n_name, ids[1])):
                row = list(self.tensor_rows[t])
                row[0] = 3
                self.tensor_rows[t] = tuple(row)
                self.out.packed.append(name)
        if self.packed_actions:
            # opt-in: bool action fluents the plane kernels read are taken bit-packed from the caller
            # (uint32 [n_envs, cells / 32]: 1/8 of the bytes of the reference boundary's bool arrays);
            # scalar programs that read them as well extract the bit (OP_LOAD of a packed tensor)
            for name in p.names_of_kind('action'):
                t = self.tensor_id[name]
                nelem = self.out.tensor_meta[name][2]
                if p.tensors[name].dtype != 'b' or nelem % 32 or t in scalar_w or t in other or t not in plane_rw:
                    continue
                row = list(self.tensor_rows[t])
                row[0] = 3
                self.tensor_rows[t] = tuple(row)
                self.out.packed.append(name)

    def _or_all(self, exprs):
        out = exprs[0]
        for e in exprs[1:]:
            out = Node('bin', op='or', args=[out, e], dtype='b', scope=[])
        return out

    def _lower_plans(self):
        p = self.p
        # --- step plan (simulator.py:442-502)
        items = []
        for ri, root in enumerate(p.cpfs):
            items.extend(self.contract_items.get(ri, []))
            e = self.hoist(root.expr, items)
            items.append(_Item(root.scope, e, store=root.name))
        e = self.hoist(p.reward, items)
        items.append(_Item([], e, store='__reward'))
        # terminations are evaluated on the NEW state (simulator.py:463-466,501): state
        # fluents are read from the next-state values written by this step
        ren = dict(p.next_state)
        titems = []
        if p.terminations:
            e = self.hoist(self._or_all(list(p.terminations)), titems)
        else:
            e = Node('const', value=False, dtype='b', scope=[])
        titems.append(_Item([], e, store='__done'))
        for it in titems:
            it.rename = ren
        self.schedule(ot.PLAN_STEP, items + titems)
        # --- terminal plan on the current state (reset, simulator.py:391-399,440)
        items = []
        if p.terminations:
            e = self.hoist(self._or_all(list(p.terminations)), items)
        else:
            e = Node('const', value=False, dtype='b', scope=[])
        items.append(_Item([], e, store='__done'))
        self.schedule(ot.PLAN_TERMINAL, items)
        # --- invariants / preconditions: one bool per expression (simulator.py:362-389)
        for (plan, exprs) in [(ot.PLAN_INVARIANT, p.invariants), (ot.PLAN_PRECONDITION, p.preconditions)]:
            items = []
            for i, ex in enumerate(exprs):
                e = self.hoist(ex, items)
                items.append(_Item([], e, store='__chk', store_off=i))
            if items:
                self.schedule(plan, items)


def lower(program, use_planes=True, use_contract=True, use_packed=True, packed_actions=False):
    """Lowers an HIR ``Program`` to an ``OpTable``. ``use_planes=False`` keeps every launch on
    the general scalar interpreter (used by tests to cross-check the two paths)."""
    return Lowerer(program, use_planes=use_planes, use_contract=use_contract, use_packed=use_packed,
                   packed_actions=packed_actions).lower()
