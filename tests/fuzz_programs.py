"""TEST INFRASTRUCTURE ONLY -- random HIR programs for differential testing of the lowering and
the kernels against the NumPy oracle (property tests over random
expression trees). Programs are built directly as HIR (pyrddlgym_b200/hir.py), the form both
the oracle and the lowering consume; every operator family of
/root/reference/pyRDDLGym/core/simulator.py:584-952 and the base-variate distributions appear."""
import numpy as np

from pyrddlgym_b200 import hir
from pyrddlgym_b200.hir import Node, Program, Root, TensorDecl

TA, TB = 'ta', 'tb'


class Gen:

    def __init__(self, seed, na=5, nb=4):
        self.rng = np.random.default_rng(seed)
        self.na, self.nb = na, nb
        self.next_id = 0
        self.p = Program()
        p = self.p
        p.name = f'fuzz{seed}'
        p.types = {TA: na, TB: nb}
        p.objects = {TA: [f'a{i}' for i in range(na)], TB: [f'b{i}' for i in range(nb)]}
        rng = self.rng
        decl = [
            ('NF', 'nonfluent', 'f', [TA]), ('NM', 'nonfluent', 'f', [TA, TB]), ('NB', 'nonfluent', 'b', [TA, TA]),
            ('NI', 'nonfluent', 'i', [TA]), ('NP', 'nonfluent', 'i', [TA]),
            ('sf', 'state', 'f', [TA]), ('si', 'state', 'i', [TA]), ('sb', 'state', 'b', [TA]),
            ('sm', 'state', 'f', [TA, TB]), ('ss', 'state', 'f', []),
            ("sf'", 'next', 'f', [TA]), ("si'", 'next', 'i', [TA]), ("sb'", 'next', 'b', [TA]),
            ("sm'", 'next', 'f', [TA, TB]), ("ss'", 'next', 'f', []),
            ('af', 'action', 'f', [TA]), ('ab', 'action', 'b', [TA]),
            ('t0', 'interm', 'f', [TA]), ('t1', 'interm', 'f', [TA, TB]), ('t2', 'interm', 'f', []),
        ]
        for (name, kind, dt, params) in decl:
            p.tensors[name] = TensorDecl(name, kind, dt, params, {'f': 'real', 'i': 'int', 'b': 'bool'}[dt])
            shape = p.shape_of(params)
            if dt == 'f':
                v = rng.uniform(-1.5, 1.5, size=shape)
            elif dt == 'i':
                v = rng.integers(0, na, size=shape) if name == 'NP' else rng.integers(-3, 6, size=shape)
            else:
                v = rng.random(shape) < 0.4
            p.init[name] = np.asarray(v, dtype=hir.NP_DTYPE[dt])
        p.next_state = {'sf': "sf'", 'si': "si'", 'sb': "sb'", 'sm': "sm'", 'ss': "ss'"}
        p.horizon, p.discount, p.max_allowed_actions = 10, 0.9, na * 2

    # ---- node helpers
    def nid(self):
        self.next_id += 1
        return self.next_id

    def const(self, dt, scope):
        rng = self.rng
        v = {'f': float(np.round(rng.uniform(-2, 2), 3)), 'i': int(rng.integers(-3, 5)), 'b': bool(rng.random() < 0.5)}[dt]
        return Node('const', value=v, dtype=dt, scope=scope, id=self.nid())

    def load(self, name, scope, index):
        return Node('load', tensor=name, index=index, dtype=self.p.tensors[name].dtype, scope=scope, id=self.nid())

    def leaf(self, dt, scope, avail):
        """scope: list of types; axis i has type scope[i]. avail: readable tensors."""
        rng = self.rng
        cands = []
        for name in avail:
            d = self.p.tensors[name]
            if d.dtype != dt:
                continue
            pos = []
            ok = True
            for t in d.params:
                axes = [i for i, s in enumerate(scope) if s == t]
                if not axes:
                    ok = False
                    break
                pos.append(axes)
            if ok:
                cands.append((name, pos))
        if not cands or rng.random() < 0.25:
            return self.const(dt, scope)
        name, pos = cands[rng.integers(len(cands))]
        index = []
        for axes in pos:
            r = rng.random()
            if r < 0.8:
                index.append(('a', int(axes[rng.integers(len(axes))])))
            elif r < 0.9:
                t = self.p.tensors[name].params[len(index)]
                index.append(('c', int(rng.integers(self.p.types[t]))))
            else:
                # nested object index: NP(?a) is a ta-valued non-fluent
                t = self.p.tensors[name].params[len(index)]
                ta_axes = [i for i, s in enumerate(scope) if s == TA]
                if t == TA and ta_axes:
                    index.append(('e', self.load('NP', scope, [('a', int(ta_axes[0]))])))
                else:
                    index.append(('a', int(axes[0])))
        return self.load(name, scope, index)

    def expr(self, dt, scope, avail, depth):
        rng = self.rng
        if depth <= 0 or rng.random() < 0.15:
            return self.leaf(dt, scope, avail)
        d = depth - 1
        E = lambda t: self.expr(t, scope, avail, d)
        r = rng.random()
        if dt == 'b':
            if r < 0.35:
                op = ['and', 'or', 'xor', 'imply', 'eqv'][rng.integers(5)]
                return Node('bin', op=op, args=[E('b'), E('b')], dtype='b', scope=scope, id=self.nid())
            if r < 0.45:
                return Node('un', op='not', args=[E('b')], dtype='b', scope=scope, id=self.nid())
            if r < 0.8:
                op = ['lt', 'le', 'gt', 'ge', 'eq', 'ne'][rng.integers(6)]
                t = 'if'[rng.integers(2)]
                return Node('bin', op=op, args=[E(t), E('if'[rng.integers(2)])], dtype='b', scope=scope, id=self.nid())
            if r < 0.9 and len(scope) < 3:
                return self.agg(['forall', 'exists'][rng.integers(2)], 'b', scope, avail, d)
            return Node('if', args=[E('b'), E('b'), E('b')], dtype='b', scope=scope, id=self.nid())
        if dt == 'i':
            if r < 0.3:
                op = ['add', 'sub', 'mul', 'min', 'max'][rng.integers(5)]
                a, b = E('ib'[rng.integers(2)]), E('i')
                return Node('bin', op=op, args=[a, b], dtype='i', scope=scope, id=self.nid())
            if r < 0.4:
                op = ['floordiv', 'mod'][rng.integers(2)]
                den = Node('bin', op='add', args=[Node('un', op='abs', args=[E('i')], dtype='i', scope=scope, id=self.nid()),
                                                   Node('const', value=1, dtype='i', scope=scope, id=self.nid())],
                           dtype='i', scope=scope, id=self.nid())
                return Node('bin', op=op, args=[E('i'), den], dtype='i', scope=scope, id=self.nid())
            if r < 0.55:
                op = ['sgn', 'round', 'floor', 'ceil'][rng.integers(4)]
                return Node('un', op=op, args=[E('f')], dtype='i', scope=scope, id=self.nid())
            if r < 0.65:
                return Node('un', op=['neg', 'abs'][rng.integers(2)], args=[E('i')], dtype='i', scope=scope, id=self.nid())
            if r < 0.8 and len(scope) < 3:
                return self.agg(['sum', 'min', 'max', 'prod', 'argmax', 'argmin'][rng.integers(6)], 'i', scope, avail, d)
            return Node('if', args=[E('b'), E('i'), E('i')], dtype='i', scope=scope, id=self.nid())
        # real
        if r < 0.3:
            op = ['add', 'sub', 'mul', 'min', 'max'][rng.integers(5)]
            a, b = E('f'), E('fib'[rng.integers(3)])
            if rng.random() < 0.5:
                a, b = b, a
            return Node('bin', op=op, args=[a, b], dtype='f', scope=scope, id=self.nid())
        if r < 0.38:
            den = Node('bin', op='add', args=[Node('un', op='abs', args=[E('f')], dtype='f', scope=scope, id=self.nid()),
                                               Node('const', value=0.5, dtype='f', scope=scope, id=self.nid())],
                       dtype='f', scope=scope, id=self.nid())
            return Node('bin', op='div', args=[E('fi'[rng.integers(2)]), den], dtype='f', scope=scope, id=self.nid())
        if r < 0.55:
            op = ['cos', 'sin', 'tanh', 'atan', 'exp', 'abs', 'neg'][rng.integers(7)]
            arg = E('f')
            if op == 'exp':
                arg = Node('un', op='tanh', args=[arg], dtype='f', scope=scope, id=self.nid())
            return Node('un', op=op, args=[arg], dtype='f', scope=scope, id=self.nid())
        if r < 0.62:
            pos = Node('bin', op='add', args=[Node('un', op='abs', args=[E('f')], dtype='f', scope=scope, id=self.nid()),
                                               Node('const', value=0.25, dtype='f', scope=scope, id=self.nid())],
                       dtype='f', scope=scope, id=self.nid())
            op = ['ln', 'sqrt', 'lngamma'][rng.integers(3)]
            return Node('un', op=op, args=[pos], dtype='f', scope=scope, id=self.nid())
        if r < 0.75 and len(scope) < 3:
            return self.agg(['sum', 'avg', 'min', 'max', 'prod'][rng.integers(5)], 'f', scope, avail, d)
        if r < 0.87:
            return self.rand(scope, avail, d)
        return Node('if', args=[E('b'), E('f'), E('fi'[rng.integers(2)])], dtype='f', scope=scope, id=self.nid())

    def agg(self, op, dt, scope, avail, depth):
        rng = self.rng
        t = [TA, TB][rng.integers(2)]
        inner = scope + [t]
        if op in ('forall', 'exists'):
            body = self.expr('b', inner, avail, depth)
        elif op in ('argmax', 'argmin'):
            body = self.expr('f', inner, avail, depth)
        elif dt == 'i':
            body = self.expr('ib'[rng.integers(2)] if op == 'sum' else 'i', inner, avail, min(depth, 1))
            if op == 'prod':   # keep integer products small
                body = Node('bin', op='mod', args=[body, Node('const', value=3, dtype='i', scope=inner, id=self.nid())],
                            dtype='i', scope=inner, id=self.nid())
        else:
            body = self.expr('f', inner, avail, depth)
            if op == 'prod':
                body = Node('un', op='tanh', args=[body], dtype='f', scope=inner, id=self.nid())
            if op == 'sum' and rng.random() < 0.4 and t == TA and scope and scope[0] == TA:
                mask = self.load('NB', inner, [('a', 0), ('a', len(scope))])
                body = Node('bin', op='mul', args=[mask, body], dtype='f', scope=inner, id=self.nid())
        out_dt = {'forall': 'b', 'exists': 'b', 'argmax': 'i', 'argmin': 'i', 'avg': 'f'}.get(
            op, 'i' if body.dtype == 'b' else body.dtype)
        return Node('agg', op=op, args=[body], red=[t], dtype=out_dt, scope=scope, id=self.nid())

    def rand(self, scope, avail, depth):
        rng = self.rng
        E = lambda: self.expr('f', scope, avail, min(depth, 1))
        mk = lambda op, a, b: Node('bin', op=op, args=[a, b], dtype='f', scope=scope, id=self.nid())
        absn = lambda x: Node('un', op='abs', args=[x], dtype='f', scope=scope, id=self.nid())
        c = lambda v: Node('const', value=v, dtype='f', scope=scope, id=self.nid())
        pos = lambda: mk('add', absn(E()), c(0.3))
        kind = ['Uniform', 'Normal', 'Exponential', 'Bernoulli', 'Weibull', 'Gumbel', 'Laplace', 'Gompertz',
                'Kumaraswamy', 'Pareto'][rng.integers(10)]
        if kind == 'Uniform':
            lo = E()
            return Node('rand', op=kind, args=[lo, mk('add', lo, pos())], dtype='f', scope=scope, id=self.nid())
        if kind == 'Normal':
            return Node('rand', op=kind, args=[E(), pos()], dtype='f', scope=scope, id=self.nid())
        if kind == 'Exponential':
            return Node('rand', op=kind, args=[pos()], dtype='f', scope=scope, id=self.nid())
        if kind == 'Bernoulli':
            pr = mk('mul', c(0.5), mk('add', Node('un', op='tanh', args=[E()], dtype='f', scope=scope, id=self.nid()), c(1.0)))
            b = Node('rand', op=kind, args=[pr], dtype='b', scope=scope, id=self.nid())
            return Node('if', args=[b, c(1.0), c(-1.0)], dtype='f', scope=scope, id=self.nid())
        if kind in ('Gumbel', 'Laplace'):
            return Node('rand', op=kind, args=[E(), pos()], dtype='f', scope=scope, id=self.nid())
        return Node('rand', op=kind, args=[pos(), pos()], dtype='f', scope=scope, id=self.nid())

    # ---- whole program
    def build(self, depth=3):
        p = self.p
        readable = ['NF', 'NM', 'NB', 'NI', 'sf', 'si', 'sb', 'sm', 'ss', 'af', 'ab']
        roots = [('t0', [TA], 'f'), ('t1', [TA, TB], 'f'), ('t2', [], 'f'), ("sf'", [TA], 'f'), ("si'", [TA], 'i'),
                 ("sb'", [TA], 'b'), ("sm'", [TA, TB], 'f'), ("ss'", [], 'f')]
        avail = list(readable)
        for level, (name, scope, dt) in enumerate(roots):
            e = self.expr(dt, list(scope), avail, depth)
            if dt == 'f':   # keep the state bounded over several steps
                e = Node('un', op='tanh', args=[e], dtype='f', scope=list(scope), id=self.nid())
            if dt == 'i':
                e = Node('bin', op='mod', args=[e, Node('const', value=7, dtype='i', scope=list(scope), id=self.nid())],
                         dtype='i', scope=list(scope), id=self.nid())
            p.cpfs.append(Root(name, level, list(scope), e))
            avail.append(name)
        p.reward = self.expr('f', [], avail, depth)
        p.invariants = [self.expr('b', [], ['sf', 'si', 'sb', 'sm', 'ss', 'NF', 'NI'], 2) for _ in range(2)]
        p.preconditions = [self.expr('b', [], ['sf', 'si', 'sb', 'af', 'ab', 'NF'], 2)]
        p.terminations = [self.expr('b', [], ['sf', 'si', 'sb', 'ss'], 2)]
        return p


def random_program(seed, depth=3):
    return Gen(seed).build(depth)


# ---------------------------------------------------------------------------------------
# random cellular (boolean-grid) programs: the class pyrddlgym_b200/plane_lowering.py targets
# ---------------------------------------------------------------------------------------

class GridGen:
    """Random Wildfire-like programs over an H x W grid: bool planes, 3x3-stencil neighbour
    counts through separable band masks, Bernoulli of a function of count and planes, reward
    sums of coefficient * boolean expression."""

    def __init__(self, seed, H=4, W=32):
        self.rng = rng = np.random.default_rng(seed)
        self.next_id = 0
        self.H, self.W = H, W
        self.p = p = Program()
        p.name = f'gridfuzz{seed}'
        p.types = {'x_pos': H, 'y_pos': W}
        p.objects = {'x_pos': [f'x{i}' for i in range(H)], 'y_pos': [f'y{i}' for i in range(W)]}
        XY = ['x_pos', 'y_pos']

        def band(n, offs):
            B = np.zeros((n, n), dtype=bool)
            for d in offs:
                i = np.arange(max(0, -d), min(n, n - d))
                B[i, i + d] = True
            return B

        offx = [d for d in (-1, 0, 1) if rng.random() < 0.8] or [0]
        offy = [d for d in (-1, 0, 1) if rng.random() < 0.8] or [1]
        decl = [('XN', 'nonfluent', 'b', ['x_pos', 'x_pos'], band(H, offx)),
                ('YN', 'nonfluent', 'b', ['y_pos', 'y_pos'], band(W, offy)),
                ('NA', 'nonfluent', 'b', XY, rng.random((H, W)) < 0.3),
                ('C0', 'nonfluent', 'f', [], np.float64(np.round(rng.uniform(-9, -1), 1))),
                ('C1', 'nonfluent', 'f', [], np.float64(np.round(rng.uniform(0.5, 4), 2)))]
        for k in range(3):
            decl.append((f's{k}', 'state', 'b', XY, rng.random((H, W)) < 0.25))
            decl.append((f"s{k}'", 'next', 'b', XY, np.zeros((H, W), dtype=bool)))
        for k in range(2):
            decl.append((f'a{k}', 'action', 'b', XY, np.zeros((H, W), dtype=bool)))
        for (name, kind, dt, params, init) in decl:
            p.tensors[name] = TensorDecl(name, kind, dt, params, {'f': 'real', 'b': 'bool'}[dt])
            p.init[name] = np.asarray(init, dtype=hir.NP_DTYPE[dt])
        p.next_state = {f's{k}': f"s{k}'" for k in range(3)}
        p.horizon, p.discount, p.max_allowed_actions = 10, 1.0, 2 * H * W
        self.XY = XY

    def nid(self):
        self.next_id += 1
        return self.next_id

    def plane(self, name, scope=None, at=(0, 1)):
        scope = self.XY if scope is None else scope
        return Node('load', tensor=name, index=[('a', at[0]), ('a', at[1])], dtype='b', scope=scope, id=self.nid())

    def count(self, src):
        """sum_{x2,y2} XN(x,x2) ^ YN(y,y2) ^ src(x2,y2) -> int count in 0..9."""
        inner = self.XY + self.XY
        xn = Node('load', tensor='XN', index=[('a', 0), ('a', 2)], dtype='b', scope=inner, id=self.nid())
        yn = Node('load', tensor='YN', index=[('a', 1), ('a', 3)], dtype='b', scope=inner, id=self.nid())
        body = Node('bin', op='and', args=[Node('bin', op='and', args=[xn, yn], dtype='b', scope=inner, id=self.nid()),
                                           self.plane(src, inner, (2, 3))], dtype='b', scope=inner, id=self.nid())
        return body, inner

    def boolean(self, depth, avail):
        rng = self.rng
        r = rng.random()
        if depth <= 0 or r < 0.2:
            return self.plane(avail[rng.integers(len(avail))])
        B = lambda: self.boolean(depth - 1, avail)
        mk = lambda **kw: Node(scope=self.XY, id=self.nid(), **kw)
        if r < 0.5:
            return mk(kind='bin', op=['and', 'or', 'xor', 'imply', 'eqv'][rng.integers(5)], args=[B(), B()], dtype='b')
        if r < 0.6:
            return mk(kind='un', op='not', args=[B()], dtype='b')
        if r < 0.7:
            return mk(kind='if', args=[B(), B(), B()], dtype='b')
        src = [s for s in avail if s.startswith('s') and not s.endswith("'")] or ['s0']
        body, inner = self.count(src[rng.integers(len(src))])
        if r < 0.8:
            return Node('agg', op='exists', args=[body], red=self.XY, dtype='b', scope=self.XY, id=self.nid())
        cnt = Node('agg', op='sum', args=[body], red=self.XY, dtype='i', scope=self.XY, id=self.nid())
        if r < 0.9:
            k = mk(kind='const', value=int(rng.integers(0, 4)), dtype='i')
            return mk(kind='bin', op=['gt', 'ge', 'eq', 'lt'][rng.integers(4)], args=[cnt, k], dtype='b')
        # Bernoulli(1 / (1 + exp(c1 - k * w)))  [optionally modulated by a plane]
        c = lambda v: mk(kind='const', value=float(v), dtype='f')
        w = mk(kind='bin', op='mul', args=[cnt, c(np.round(rng.uniform(0.3, 1.5), 2))], dtype='f')
        e = mk(kind='un', op='exp', args=[mk(kind='bin', op='sub', args=[Node('load', tensor='C1', index=[], dtype='f', scope=self.XY, id=self.nid()), w], dtype='f')], dtype='f')
        pr = mk(kind='bin', op='div', args=[c(1.0), mk(kind='bin', op='add', args=[c(1.0), e], dtype='f')], dtype='f')
        if rng.random() < 0.4:
            pr = mk(kind='if', args=[self.plane(avail[rng.integers(len(avail))]), pr, c(np.round(rng.uniform(0, 1), 2))], dtype='f')
        return mk(kind='rand', op='Bernoulli', args=[pr], dtype='b')

    def build(self, depth=3):
        p = self.p
        avail = ['s0', 's1', 's2', 'a0', 'a1', 'NA']
        for k in range(3):
            p.cpfs.append(Root(f"s{k}'", 0, list(self.XY), self.boolean(depth, avail)))
        terms = []
        for k in range(int(self.rng.integers(1, 4))):
            coef = Node('load', tensor='C0', index=[], dtype='f', scope=self.XY, id=self.nid()) if k == 0 else \
                Node('const', value=float(np.round(self.rng.uniform(-5, 5), 1)), dtype='f', scope=self.XY, id=self.nid())
            body = Node('bin', op='mul', args=[coef, self.boolean(1, avail)], dtype='f', scope=self.XY, id=self.nid())
            terms.append(Node('agg', op='sum', args=[body], red=list(self.XY), dtype='f', scope=[], id=self.nid()))
        r = terms[0]
        for t in terms[1:]:
            r = Node('bin', op='add', args=[r, t], dtype='f', scope=[], id=self.nid())
        p.reward = r
        return p


def random_grid_program(seed, H=4, W=32, depth=3):
    return GridGen(seed, H, W).build(depth)
