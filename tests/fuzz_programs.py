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
