"""TEST INFRASTRUCTURE ONLY -- random RDDL domains in the neutral spec form of
pyrddlgym_b200/workloads.py (tuples shaped like the reference parser's AST), so that the SAME random
program can be given to the unmodified reference RDDLSimulator (oracle/ast_builders.py) and -- through
the front end -- to the oracle, the lowering and the kernels. Expressions are typed (real / int / bool)
and scoped over ?a : ta, ?b : tb; arguments of partial functions are made safe by construction."""
import numpy as np

from pyrddlgym_b200.workloads import (add, agg, discrete, div, eq, equiv, func, ge, gt, implies, ite, land, le, lnot,
                                      lor, lt, mul, ne, neg, num, pv, rand, sub, switch, xor, boolean)

VARS = {'?a': 'ta', '?b': 'tb'}


class SpecGen:
    def __init__(self, seed, n=4, m=3):
        self.rng = np.random.default_rng(seed)
        self.n, self.m = n, m
        self.allow_rand = True
        # name -> (dtype, params)
        self.readable = {
            'K': ('f', ['?a']), 'M': ('f', ['?a', '?b']), 'IK': ('i', ['?a']), 'FLAG': ('b', ['?a']),
            'fa': ('f', ['?a']), 'ia': ('i', ['?a']), 'ba': ('b', ['?a']), 'fab': ('f', ['?a', '?b']), 's0': ('f', []),
            'act': ('f', ['?a']), 'bact': ('b', ['?a']),
        }

    def pick(self, xs):
        return xs[int(self.rng.integers(0, len(xs)))]

    def leaf(self, dt, scope):
        rng = self.rng
        cands = [k for k, (d, ps) in self.readable.items() if d == dt and all(p in scope for p in ps)]
        if cands and rng.random() < 0.75:
            k = self.pick(cands)
            return pv(k, *self.readable[k][1])
        if dt == 'f':
            return num(float(np.round(rng.uniform(-2, 2), 3)))
        if dt == 'i':
            return num(int(rng.integers(-3, 5)))
        return boolean(rng.random() < 0.5)

    def gen(self, dt, scope, depth):
        rng = self.rng
        if depth <= 0 or rng.random() < 0.12:
            return self.leaf(dt, scope)
        g = lambda d, s=scope, k=depth - 1: self.gen(d, s, k)
        free = [v for v in VARS if v not in scope]
        if dt == 'f':
            c = int(rng.integers(0, 12))
            if c == 0:
                return self.pick([add, sub, mul])(g('f'), g(self.pick(['f', 'i'])))
            if c == 1:
                return div(g('f'), add(func('abs', g('f')), num(0.5)))
            if c == 2:
                return func(self.pick(['abs', 'cos', 'sin', 'tanh', 'atan']), g('f'))
            if c == 3:
                return self.pick([lambda x: func('exp', mul(func('tanh', x), num(0.5))),
                                  lambda x: func('sqrt', add(func('abs', x), num(0.1))),
                                  lambda x: func('ln', add(func('abs', x), num(1.0)))])(g('f'))
            if c == 4:
                k = self.pick(['min', 'max', 'hypot', 'fmod', 'log'])
                if k == 'fmod':
                    return func('fmod', g('f'), num(float(self.pick([0.7, -0.7, 1.3]))))
                if k == 'log':
                    return func('log', add(func('abs', g('f')), num(1.1)), num(2.0))
                return func(k, g('f'), g(self.pick(['f', 'i'])))
            if c == 5:
                return func('pow', add(func('abs', g('f')), num(0.1)), num(float(np.round(rng.uniform(0.5, 2.0), 2))))
            if c == 6:
                return ite(g('b'), g('f'), g('f'))
            if c == 7 and free:
                v = self.pick(free)
                op = self.pick(['sum', 'avg', 'max', 'min', 'prod'])
                body = self.gen('f', scope + [v], depth - 1)
                if op == 'prod':
                    body = add(num(1.0), mul(num(0.1), func('tanh', body)))
                return agg(op, [(v, VARS[v])], body)
            if c == 8 and self.allow_rand:
                # (x appears twice below: it must not contain a random node itself -- two occurrences are two draws)
                self.allow_rand = False
                x = func('tanh', g('f'))
                self.allow_rand = not getattr(self, 'deterministic', False)
                pos = add(func('abs', x), num(0.2))
                return self.pick([lambda: rand('Uniform', x, add(x, num(1.0))),
                                  lambda: rand('Normal', x, add(func('abs', x), num(0.1))),
                                  lambda: rand('Exponential', pos),
                                  lambda: rand('Weibull', num(1.5), pos),
                                  lambda: rand('Gumbel', x, num(0.5)),
                                  lambda: rand('Laplace', x, pos),
                                  lambda: rand('Cauchy', x, num(0.01)),
                                  lambda: rand('Gompertz', num(1.2), pos),
                                  lambda: rand('Kumaraswamy', num(2.0), add(pos, num(1.0))),
                                  lambda: rand('Pareto', num(3.0), pos),
                                  lambda: rand('DiracDelta', x)])()
            if c == 9:
                k = int(rng.integers(0, 6))
                if k == 0 and '?a' in scope:        # nested object index (non-fluent and state pointers)
                    return pv('fa', pv(self.pick(['NEXT', 'ptr']), '?a'))
                if k == 1 and '?a' in scope:        # enum-valued fluent: switch with / without default
                    if rng.random() < 0.5:
                        return switch(pv('col', '?a'), [('@red', g('f')), ('@blue', g('f'))], default=pv('CW', '@green'))
                    return switch(pv('col', '?a'), [('@red', g('f')), ('@green', g('f')), ('@blue', g('f'))])
                if k == 2 and '?a' in scope and '?b' not in scope:     # argmax / argmin as an index
                    body = self.gen('f', scope + ['?b'], depth - 1)
                    return pv('fab', '?a', agg(self.pick(['argmax', 'argmin']), [('?b', 'tb')], body))
                if k == 3 and '?a' in scope:
                    return pv('CW', pv('col', '?a'))
                if k == 4 and '?a' in scope:        # object-valued state fluent as an index
                    return pv('fab', '?a', pv('pick', '?a'))
                return mul(num(1.0), g('i'))
            if c == 10:
                return neg(g('f'))
            return mul(num(0.5), add(g('f'), g('f')))
        if dt == 'i':
            c = int(rng.integers(0, 8))
            if c == 0:
                return self.pick([add, sub])(g('i'), g('i'))
            if c == 1:
                return mul(g('i'), num(int(rng.integers(-2, 3))))
            if c == 2:
                return func(self.pick(['abs', 'sgn']), g('i'))
            if c == 3:
                return func(self.pick(['round', 'floor', 'ceil', 'sgn']), mul(func('tanh', g('f')), num(3.0)))
            if c == 4:
                return func(self.pick(['div', 'mod']), g('i'), num(int(self.pick([2, 3, 5]))))
            if c == 5:
                k = self.pick(['min', 'max', 'pow', 'kron'])
                if k == 'pow':
                    return func('pow', func('max', num(-3), func('min', num(3), g('i'))), num(2))
                if k == 'kron' and self.allow_rand:
                    return rand('KronDelta', g('i'))
                return func(self.pick(['min', 'max']), g('i'), g('i'))
            if c == 6:
                return ite(g('b'), g('i'), g('i'))
            if free:
                v = self.pick(free)
                return agg('sum', [(v, VARS[v])], self.gen('b', scope + [v], depth - 1))
            return add(g('i'), g('b'))
        c = int(rng.integers(0, 9))
        if c == 8 and '?a' in scope:
            return self.pick([lambda: eq(pv('col', '?a'), pv(self.pick(['@red', '@green', '@blue']))),
                              lambda: eq(pv('ptr', '?a'), pv('NEXT', '?a')),
                              lambda: ne(pv('ptr', '?a'), pv('?a'))])()
        if c == 0:
            return self.pick([lt, le, gt, ge])(g('f'), g('f'))
        if c == 1:
            return self.pick([eq, ne, lt, ge])(g('i'), g('i'))
        if c == 2:
            return self.pick([land, lor, xor, implies, equiv])(g('b'), g('b'))
        if c == 3:
            return lnot(g('b'))
        if c == 4 and free:
            v = self.pick(free)
            return agg(self.pick(['forall', 'exists']), [(v, VARS[v])], self.gen('b', scope + [v], depth - 1))
        if c == 5 and self.allow_rand:
            return rand('Bernoulli', mul(num(0.5), add(num(1.0), func('tanh', g('f')))))
        if c == 6:
            return ite(g('b'), g('b'), g('b'))
        return gt(g('f'), num(0.0))


def random_spec(seed, depth=3, n=4, m=3, deterministic=False):
    """deterministic=True: no random variables anywhere (the drop-in tests compare native runs)."""
    G = SpecGen(seed, n, m)
    G.deterministic = deterministic
    G.allow_rand = not deterministic
    rng = G.rng
    As = [f'a{i + 1}' for i in range(n)]
    Bs = [f'b{i + 1}' for i in range(m)]
    a, b = '?a', '?b'
    colors = ['@red', '@green', '@blue']
    pvars = [
        ('K', 'non-fluent', 'real', ['ta'], 1.0), ('M', 'non-fluent', 'real', ['ta', 'tb'], 0.0),
        ('IK', 'non-fluent', 'int', ['ta'], 1), ('FLAG', 'non-fluent', 'bool', ['ta'], False),
        ('fa', 'state-fluent', 'real', ['ta'], 0.5), ('ia', 'state-fluent', 'int', ['ta'], 2),
        ('ba', 'state-fluent', 'bool', ['ta'], False), ('fab', 'state-fluent', 'real', ['ta', 'tb'], 0.25),
        ('s0', 'state-fluent', 'real', None, 1.5),
        ('NEXT', 'non-fluent', 'ta', ['ta'], None), ('CW', 'non-fluent', 'real', ['color'], 1.0),
        ('ptr', 'state-fluent', 'ta', ['ta'], None), ('col', 'state-fluent', 'color', ['ta'], '@red'),
        ('act', 'action-fluent', 'real', ['ta'], 0.0), ('bact', 'action-fluent', 'bool', ['ta'], False),
        ('t0', 'interm-fluent', 'real', ['ta'], 0.0), ('t1', 'interm-fluent', 'real', ['ta', 'tb'], 0.0),
        ('t2', 'interm-fluent', 'real', None, 0.0), ('tb0', 'interm-fluent', 'bool', ['ta'], False),
        ('ti0', 'interm-fluent', 'int', ['ta'], 0), ('t3', 'interm-fluent', 'real', ['ta'], 0.0),
        ('iact', 'action-fluent', 'int', None, 0), ('pick', 'state-fluent', 'tb', ['ta'], None),
    ]
    cpfs = [(('t0', [a]), G.gen('f', [a], depth)), (('t1', [a, b]), G.gen('f', [a, b], depth)),
            (('tb0', [a]), G.gen('b', [a], depth)), (('ti0', [a]), G.gen('i', [a], depth))]
    G.readable.update({'t0': ('f', ['?a']), 't1': ('f', ['?a', '?b']), 'tb0': ('b', ['?a']), 'ti0': ('i', ['?a'])})
    G.readable['iact'] = ('i', [])
    cpfs.append((('t3', [a]), G.gen('f', [a], depth)))          # second interm level: may read t0 / t1 / tb0 / ti0
    G.readable['t3'] = ('f', ['?a'])
    cpfs.append((('t2', None), G.gen('f', [], depth)))
    G.readable['t2'] = ('f', [])
    cpfs += [
        (("fa'", [a]), func('tanh', G.gen('f', [a], depth))),
        (("ia'", [a]), func('max', num(-5), func('min', num(9), G.gen('i', [a], depth)))),
        (("ba'", [a]), G.gen('b', [a], depth)),
        (("fab'", [a, b]), func('tanh', G.gen('f', [a, b], depth))),
        (("s0'", None), func('tanh', G.gen('f', [], depth))),
        (("ptr'", [a]), ite(G.gen('b', [a], 2), pv('NEXT', a), pv('ptr', pv('NEXT', a)))),
        # object-valued sample from a normalised / unnormalised probability vector over tb
        (("pick'", [a]), agg('argmin', [(b, 'tb')], G.gen('f', [a, b], 2)) if deterministic else G.pick([
            lambda: ('randomvar', ('UnnormDiscrete(p)', (('typed_var', (b, 'tb')),
                                   (add(num(0.1), func('abs', func('tanh', G.gen('f', [a, b], 2)))),)))),
            lambda: ('randomvar', ('Discrete(p)', (('typed_var', (b, 'tb')), (num(1.0 / m),)))),
            lambda: agg('argmax', [(b, 'tb')], G.gen('f', [a, b], 2))])()),
        (("col'", [a]), ite(G.gen('b', [a], 2), pv('col', a), pv('@green')) if deterministic else G.pick([
            lambda: discrete('color', [('@red', num(0.2)), ('@green', num(0.3)), ('@blue', num(0.5))]),
            lambda: discrete('color', [('@red', mul(num(0.5), add(num(1.0), func('tanh', G.gen('f', [a], 2))))),
                                       ('@green', num(1.0)), ('@blue', num(0.5))], unnorm=True),
            lambda: ite(G.gen('b', [a], 2), pv('col', a), pv('@blue'))])()),
    ]
    G.readable.update({"fa'": ('f', ['?a']), "ba'": ('b', ['?a'])})
    reward = add(G.gen('f', [], depth), agg('sum', [(a, 'ta')], mul(num(-0.1), func('abs', pv('act', a)))))
    # terminations may only read state fluents and non-fluents (levels.py VALID_DEPENDENCIES)
    full = dict(G.readable)
    G.readable = {k: v for k, v in full.items() if k in ('K', 'M', 'IK', 'FLAG', 'fa', 'ia', 'ba', 'fab', 's0')}
    G.allow_rand = False     # (the harness resets the reference before any variates exist)
    terminal = G.gen('b', [], 2)
    G.allow_rand = not deterministic
    G.readable = full
    init_nf, init_state = [], []
    for i in range(n):
        init_nf.append((('K', [As[i]]), float(np.round(rng.random() + 0.5, 3))))
        init_nf.append((('IK', [As[i]]), int(rng.integers(0, 4))))
        init_nf.append((('FLAG', [As[i]]), bool(rng.random() < 0.5)))
        init_state.append((('fa', [As[i]]), float(np.round(rng.standard_normal() * 0.5, 3))))
        init_state.append((('ia', [As[i]]), int(rng.integers(-3, 8))))
        init_state.append((('ba', [As[i]]), bool(rng.random() < 0.5)))
        init_nf.append((('NEXT', [As[i]]), As[(i + 1) % n]))
        init_state.append((('ptr', [As[i]]), As[int(rng.integers(0, n))]))
        init_state.append((('col', [As[i]]), colors[int(rng.integers(0, 3))]))
        init_state.append((('pick', [As[i]]), Bs[int(rng.integers(0, m))]))
        for j in range(m):
            init_nf.append((('M', [As[i], Bs[j]]), float(np.round(rng.standard_normal(), 3))))
            init_state.append((('fab', [As[i], Bs[j]]), float(np.round(rng.random() - 0.5, 3))))
    for (c_, w_) in zip(colors, (0.7, 1.3, 2.1)):
        init_nf.append((('CW', [c_]), w_))
    return dict(
        name=f'fuzzspec{seed}', types=[('ta', 'object'), ('tb', 'object'), ('color', colors)], pvariables=pvars, cpfs=cpfs, reward=reward,
        preconds=[agg('forall', [(a, 'ta')], le(func('abs', pv('act', a)), num(5.0)))],
        invariants=[agg('forall', [(a, 'ta')], le(func('abs', pv('fa', a)), num(1.0)))],
        terminals=[terminal],
        objects=[('ta', As), ('tb', Bs)], init_non_fluent=init_nf, init_state=init_state,
        max_nondef_actions='pos-inf', horizon=10, discount=0.9,
    )


# ---------------------------------------------------------------------------------------------
# Wildfire-like boolean grid domains (the bit-plane path of the lowering) in spec form
# ---------------------------------------------------------------------------------------------

class GridSpecGen:
    """Random domains over an H x W grid: bool planes, 3x3-stencil neighbour counts through separable
    band masks XN(x,x2) ^ YN(y,y2), Bernoulli of a logistic function of the count, reward sums of
    coefficient * boolean expression (the spec-form twin of tests/fuzz_programs.py:GridGen)."""

    def __init__(self, seed, H, W):
        self.rng = np.random.default_rng(seed)
        self.H, self.W = H, W

    def plane(self, name, x='?x', y='?y'):
        return pv(name, x, y)

    def nb(self, src):
        return land(land(pv('XN', '?x', '?x2'), pv('YN', '?y', '?y2')), pv(src, '?x2', '?y2'))

    def boolean(self, depth, avail):
        rng = self.rng
        r = rng.random()
        if depth <= 0 or r < 0.2:
            return self.plane(avail[int(rng.integers(len(avail)))])
        B = lambda: self.boolean(depth - 1, avail)
        if r < 0.5:
            return [land, lor, xor, implies, equiv][int(rng.integers(5))](B(), B())
        if r < 0.6:
            return lnot(B())
        if r < 0.7:
            return ite(B(), B(), B())
        tv = [('?x2', 'x_pos'), ('?y2', 'y_pos')]
        src = [s for s in avail if s.startswith('s')] or ['s0']
        body = self.nb(src[int(rng.integers(len(src)))])
        if r < 0.8:
            return agg('exists', tv, body)
        cnt = agg('sum', tv, body)
        if r < 0.9:
            return [gt, ge, eq, lt][int(rng.integers(4))](cnt, num(int(rng.integers(0, 4))))
        w = float(np.round(rng.uniform(0.3, 1.5), 2))
        pr = div(num(1.0), add(num(1.0), func('exp', sub(pv('C1'), mul(cnt, num(w))))))
        if rng.random() < 0.4:
            pr = ite(self.plane(avail[int(rng.integers(len(avail)))]), pr, num(float(np.round(rng.uniform(0, 1), 2))))
        return rand('Bernoulli', pr)


def random_grid_spec(seed, H=3, W=32, depth=3):
    G = GridSpecGen(seed, H, W)
    rng = G.rng
    xs = [f'x{i + 1}' for i in range(H)]
    ys = [f'y{i + 1}' for i in range(W)]
    offx = [d for d in (-1, 0, 1) if rng.random() < 0.8] or [0]
    offy = [d for d in (-1, 0, 1) if rng.random() < 0.8] or [1]
    pvars = [
        ('XN', 'non-fluent', 'bool', ['x_pos', 'x_pos'], False), ('YN', 'non-fluent', 'bool', ['y_pos', 'y_pos'], False),
        ('NA', 'non-fluent', 'bool', ['x_pos', 'y_pos'], False),
        ('C0', 'non-fluent', 'real', None, float(np.round(rng.uniform(-9, -1), 1))),
        ('C1', 'non-fluent', 'real', None, float(np.round(rng.uniform(0.5, 4), 2))),
    ]
    for k in range(3):
        pvars.append((f's{k}', 'state-fluent', 'bool', ['x_pos', 'y_pos'], False))
    for k in range(2):
        pvars.append((f'a{k}', 'action-fluent', 'bool', ['x_pos', 'y_pos'], False))
    init_nf, init_state = [], []
    for i in range(H):
        for d in offx:
            if 0 <= i + d < H:
                init_nf.append((('XN', [xs[i], xs[i + d]]), True))
    for j in range(W):
        for d in offy:
            if 0 <= j + d < W:
                init_nf.append((('YN', [ys[j], ys[j + d]]), True))
    for i in range(H):
        for j in range(W):
            if rng.random() < 0.3:
                init_nf.append((('NA', [xs[i], ys[j]]), True))
            for k in range(3):
                if rng.random() < 0.25:
                    init_state.append(((f's{k}', [xs[i], ys[j]]), True))
    avail = ['s0', 's1', 's2', 'a0', 'a1', 'NA']
    xy = ['?x', '?y']
    cpfs = [((f"s{k}'", xy), G.boolean(depth, avail)) for k in range(3)]
    axy = [('?x', 'x_pos'), ('?y', 'y_pos')]
    terms = []
    for k in range(int(rng.integers(1, 4))):
        coef = pv('C0') if k == 0 else num(float(np.round(rng.uniform(-5, 5), 1)))
        terms.append(agg('sum', axy, mul(coef, G.boolean(1, avail))))
    reward = terms[0]
    for t_ in terms[1:]:
        reward = add(reward, t_)
    return dict(
        name=f'gridspec{seed}', types=[('x_pos', 'object'), ('y_pos', 'object')], pvariables=pvars, cpfs=cpfs,
        reward=reward, preconds=[], invariants=[], terminals=[],
        objects=[('x_pos', xs), ('y_pos', ys)], init_non_fluent=init_nf, init_state=init_state,
        max_nondef_actions='pos-inf', horizon=10, discount=1.0,
    )
