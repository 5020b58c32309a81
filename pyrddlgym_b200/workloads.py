"""Synthetic RDDL workloads for the five BASELINE.json configs (+ an operator zoo).

Each factory returns a *neutral domain spec*: plain Python tuples in exactly the
shapes the reference's PLY grammar produces (SURVEY.md appendix C;
/root/reference/pyRDDLGym/core/parser/parser.py:705-980), but without using any
reference class, so the spec can be built anywhere. ``oracle/ast_builders.py``
wraps a spec into the reference's own ``Expression/PVariable/Domain/...``
classes when the reference is importable (build container); the compiled
programs for the GPU box are serialised from there (``pyrddlgym_b200/programs``).

No RDDL text exists in the reference tree except Wildfire
(/root/reference/pyRDDLGym/examples/run_build.py:7-79); every other domain here
is a synthetic restatement authored for this repo and labelled as such.
"""
import math

import numpy as np

# ---------------------------------------------------------------------------
# tiny expression DSL producing the parser's tuple shapes
# ---------------------------------------------------------------------------


def num(v):
    return ('number', v)


def boolean(v):
    return ('boolean', bool(v))


def pv(name, *args):
    """pvariable reference f(?x, @lit, <nested expr>) or scalar f / free ?x."""
    return ('pvar_expr', (name, list(args) if args else None))


def _bin(op):
    return lambda a, b: (op, (a, b))


add, sub, mul, div = _bin('+'), _bin('-'), _bin('*'), _bin('/')
lt, le, gt, ge, eq, ne = _bin('<'), _bin('<='), _bin('>'), _bin('>='), _bin('=='), _bin('~=')
land, lor, xor, implies, equiv = _bin('^'), _bin('|'), _bin('~'), _bin('=>'), _bin('<=>')


def neg(a):
    return ('-', (a,))


def lnot(a):
    return ('~', (a,))


def func(name, *args):
    return ('func', (name, list(args)))


def agg(op, typed_vars, body):
    """op in sum prod avg max min forall exists argmax argmin;
    typed_vars = [('?x','type'), ...]."""
    return (op, tuple(('typed_var', (v, t)) for (v, t) in typed_vars) + (body,))


def ite(c, a, b):
    return ('if', (c, a, b))


def switch(pred, cases, default=None):
    items = [('case', (lit, e)) for (lit, e) in cases]
    if default is not None:
        items.append(('default', default))
    return ('switch', (pred,) + tuple(items))


def rand(name, *args):
    return ('randomvar', (name, tuple(args)))


def discrete(enum_type, cases, unnorm=False):
    name = 'UnnormDiscrete' if unnorm else 'Discrete'
    return ('randomvar', (name, (('enum_type', enum_type),) +
                          tuple(('lconst', (lit, e)) for (lit, e) in cases)))


def discrete_pvar(var, typ, body, unnorm=False):
    name = 'UnnormDiscrete(p)' if unnorm else 'Discrete(p)'
    return ('randomvar', (name, (('typed_var', (var, typ)), (body,))))


def sum_all(terms):
    out = terms[0]
    for t in terms[1:]:
        out = add(out, t)
    return out


# ---------------------------------------------------------------------------
# C1: CartPole (continuous force), gym physics -- synthetic restatement
# ---------------------------------------------------------------------------

def cartpole():
    G, MC, MP, L, FMAX, DT, XLIM, ALIM = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02, 2.4, 0.21
    pvars = [
        ('GRAVITY', 'non-fluent', 'real', None, G),
        ('CART-MASS', 'non-fluent', 'real', None, MC),
        ('POLE-MASS', 'non-fluent', 'real', None, MP),
        ('POLE-LEN', 'non-fluent', 'real', None, L),
        ('FORCE-MAX', 'non-fluent', 'real', None, FMAX),
        ('TIME-STEP', 'non-fluent', 'real', None, DT),
        ('POS-LIMIT', 'non-fluent', 'real', None, XLIM),
        ('ANG-LIMIT', 'non-fluent', 'real', None, ALIM),
        ('temp', 'interm-fluent', 'real', None, 0.0),
        ('acc', 'interm-fluent', 'real', None, 0.0),
        ('ang-acc', 'interm-fluent', 'real', None, 0.0),
        ('pos', 'state-fluent', 'real', None, 0.0),
        ('ang-pos', 'state-fluent', 'real', None, 0.0),
        ('vel', 'state-fluent', 'real', None, 0.0),
        ('ang-vel', 'state-fluent', 'real', None, 0.0),
        ('force', 'action-fluent', 'real', None, 0.0),
    ]
    total = add(pv('CART-MASS'), pv('POLE-MASS'))
    sin_a, cos_a = func('sin', pv('ang-pos')), func('cos', pv('ang-pos'))
    temp = div(add(pv('force'),
                   mul(mul(mul(pv('POLE-MASS'), pv('POLE-LEN')),
                           mul(pv('ang-vel'), pv('ang-vel'))), sin_a)), total)
    ang_acc = div(sub(mul(pv('GRAVITY'), sin_a), mul(cos_a, pv('temp'))),
                  mul(pv('POLE-LEN'),
                      sub(div(num(4.0), num(3.0)),
                          div(mul(pv('POLE-MASS'), mul(cos_a, cos_a)), total))))
    acc = sub(pv('temp'),
              div(mul(mul(mul(pv('POLE-MASS'), pv('POLE-LEN')), pv('ang-acc')), cos_a), total))
    cpfs = [
        (('temp', None), temp),
        (('ang-acc', None), ang_acc),
        (('acc', None), acc),
        (("pos'", None), add(pv('pos'), mul(pv('TIME-STEP'), pv('vel')))),
        (("ang-pos'", None), add(pv('ang-pos'), mul(pv('TIME-STEP'), pv('ang-vel')))),
        (("vel'", None), add(pv('vel'), mul(pv('TIME-STEP'), pv('acc')))),
        (("ang-vel'", None), add(pv('ang-vel'), mul(pv('TIME-STEP'), pv('ang-acc')))),
    ]
    in_bounds = land(land(le(pv("pos'"), pv('POS-LIMIT')), ge(pv("pos'"), neg(pv('POS-LIMIT')))),
                     land(le(pv("ang-pos'"), pv('ANG-LIMIT')), ge(pv("ang-pos'"), neg(pv('ANG-LIMIT')))))
    reward = ite(in_bounds, num(1.0), num(0.0))
    return dict(
        name='cart_pole_continuous',
        types=[],
        pvariables=pvars,
        cpfs=cpfs,
        reward=reward,
        preconds=[ge(pv('force'), neg(pv('FORCE-MAX'))), le(pv('force'), pv('FORCE-MAX'))],
        invariants=[],
        terminals=[gt(pv('pos'), pv('POS-LIMIT')), lt(pv('pos'), neg(pv('POS-LIMIT'))),
                   gt(pv('ang-pos'), pv('ANG-LIMIT')), lt(pv('ang-pos'), neg(pv('ANG-LIMIT')))],
        objects=[],
        init_non_fluent=[],
        init_state=[],
        max_nondef_actions='pos-inf', horizon=200, discount=1.0,
    )


# ---------------------------------------------------------------------------
# C2 / C4: Wildfire -- domain text of /root/reference/pyRDDLGym/examples/run_build.py:7-44,
# instance recipe of run_build.py:46-79 scaled to n x n
# ---------------------------------------------------------------------------

def wildfire_masks(n):
    """TARGET and initial-fire boolean grids per run_build.py:61-74, rescaled from the
    30x30 recipe to n x n (disc centres scale with n/30; radii are kept >= 1)."""
    s = n / 30.0
    target = np.zeros((n, n), dtype=bool)
    for x0, y0, k in [(25, 25, 3), (27, 19, 2), (5, 23, 2)]:
        cx, cy = min(n - 1, int(round(x0 * s))), min(n - 1, int(round(y0 * s)))
        k = max(1, int(math.ceil(k * max(1.0, s))))
        for x in range(max(0, cx - k), min(n, cx + k + 1)):
            for y in range(max(0, cy - k), min(n, cy + k + 1)):
                if (x - cx) ** 2 + (y - cy) ** 2 < k ** 2:
                    target[x, y] = True
    burning = np.zeros((n, n), dtype=bool)
    cx, cy = min(n - 1, int(round(6 * s))), min(n - 1, int(round(6 * s)))
    k = max(1, int(math.ceil(2 * max(1.0, s))))
    for x in range(max(0, cx - k), min(n, cx + k + 1)):
        for y in range(max(0, cy - k), min(n, cy + k + 1)):
            if (x - cx) ** 2 + (y - cy) ** 2 < k ** 2:
                burning[x, y] = True
    return target, burning


def wildfire(n=32, separable=False, max_nondef='pos-inf', horizon=100):
    """Wildfire n x n. ``separable=False`` is the reference's 4-D NEIGHBOR(x,y,x2,y2)
    formulation (8-neighbourhood, run_build.py:51-59). ``separable=True`` replaces it by
    XN(x,x2) ^ YN(y,y2) (3-wide bands, *including* the cell itself); the two agree on the
    only branch that reads the neighbour aggregate (~burning(?x,?y)), see SURVEY.md section 7."""
    if separable:
        def nb(x2, y2):
            return land(land(pv('XN', '?x', x2), pv('YN', '?y', y2)), pv('burning', x2, y2))
    else:
        def nb(x2, y2):
            return land(pv('NEIGHBOR', '?x', '?y', x2, y2), pv('burning', x2, y2))
    tv = [('?x2', 'x_pos'), ('?y2', 'y_pos')]
    xy = ('?x', '?y')
    burning = ite(
        pv('put-out', *xy), boolean(False),
        ite(land(lnot(pv('out-of-fuel', *xy)), lnot(pv('burning', *xy))),
            ite(land(pv('TARGET', *xy), lnot(agg('exists', tv, nb('?x2', '?y2')))),
                boolean(False),
                rand('Bernoulli',
                     div(num(1.0), add(num(1.0), func('exp', sub(num(4.5), agg('sum', tv, nb('?x2', '?y2')))))))),
            pv('burning', *xy)))
    oof = lor(lor(pv('out-of-fuel', *xy), pv('burning', *xy)),
              land(lnot(pv('TARGET', *xy)), pv('cut-out', *xy)))
    axy = [('?x', 'x_pos'), ('?y', 'y_pos')]
    reward = sum_all([
        agg('sum', axy, mul(pv('COST_CUTOUT'), pv('cut-out', *xy))),
        agg('sum', axy, mul(pv('COST_PUTOUT'), pv('put-out', *xy))),
        agg('sum', axy, mul(pv('PENALTY_TARGET_BURN'),
                            land(lor(pv('burning', *xy), pv('out-of-fuel', *xy)), pv('TARGET', *xy)))),
        agg('sum', axy, mul(pv('PENALTY_NONTARGET_BURN'),
                            land(pv('burning', *xy), lnot(pv('TARGET', *xy))))),
    ])
    pvars = [
        ('COST_CUTOUT', 'non-fluent', 'real', None, -5.0),
        ('COST_PUTOUT', 'non-fluent', 'real', None, -10.0),
        ('PENALTY_TARGET_BURN', 'non-fluent', 'real', None, -100.0),
        ('PENALTY_NONTARGET_BURN', 'non-fluent', 'real', None, -5.0),
        ('TARGET', 'non-fluent', 'bool', ['x_pos', 'y_pos'], False),
        ('burning', 'state-fluent', 'bool', ['x_pos', 'y_pos'], False),
        ('out-of-fuel', 'state-fluent', 'bool', ['x_pos', 'y_pos'], False),
        ('put-out', 'action-fluent', 'bool', ['x_pos', 'y_pos'], False),
        ('cut-out', 'action-fluent', 'bool', ['x_pos', 'y_pos'], False),
    ]
    xs = [f'x{i + 1}' for i in range(n)]
    ys = [f'y{i + 1}' for i in range(n)]
    init_nf = []
    if separable:
        pvars.insert(4, ('XN', 'non-fluent', 'bool', ['x_pos', 'x_pos'], False))
        pvars.insert(5, ('YN', 'non-fluent', 'bool', ['y_pos', 'y_pos'], False))
        for a in range(n):
            for d in (-1, 0, 1):
                if 0 <= a + d < n:
                    init_nf.append((('XN', [xs[a], xs[a + d]]), True))
                    init_nf.append((('YN', [ys[a], ys[a + d]]), True))
    else:
        pvars.insert(4, ('NEIGHBOR', 'non-fluent', 'bool', ['x_pos', 'y_pos', 'x_pos', 'y_pos'], False))
        for x1 in range(n):
            for y1 in range(n):
                for dx in (-1, 0, 1):
                    for dy in (-1, 0, 1):
                        x2, y2 = x1 + dx, y1 + dy
                        if 0 <= x2 < n and 0 <= y2 < n and not (dx == 0 and dy == 0):
                            init_nf.append((('NEIGHBOR', [xs[x1], ys[y1], xs[x2], ys[y2]]), True))
    target, fire = wildfire_masks(n)
    for (x, y) in zip(*np.nonzero(target)):
        init_nf.append((('TARGET', [xs[x], ys[y]]), True))
    init_state = [(('burning', [xs[x], ys[y]]), True) for (x, y) in zip(*np.nonzero(fire))]
    return dict(
        name='wildfire_sep' if separable else 'wildfire',
        types=[('x_pos', 'object'), ('y_pos', 'object')],
        pvariables=pvars,
        cpfs=[(("burning'", ['?x', '?y']), burning), (("out-of-fuel'", ['?x', '?y']), oof)],
        reward=reward, preconds=[], invariants=[], terminals=[],
        objects=[('x_pos', xs), ('y_pos', ys)],
        init_non_fluent=init_nf, init_state=init_state,
        max_nondef_actions=max_nondef, horizon=horizon, discount=1.0,
    )


# ---------------------------------------------------------------------------
# C3: Reservoir control -- synthetic restatement
# ---------------------------------------------------------------------------

def reservoir_tensors(R, seed=0):
    rng = np.random.default_rng(seed)
    connect = np.zeros((R, R), dtype=bool)
    for u in range(R - 1):
        connect[u, u + 1] = True                      # chain
    extra = rng.random((R, R)) < 0.02                 # 2% random extra edges
    np.fill_diagonal(extra, False)
    connect |= extra
    top = 80.0 + 40.0 * rng.random(R)
    mean = 2.0 + 6.0 * rng.random(R)
    var = 0.5 + 2.0 * rng.random(R)
    level = 30.0 + 40.0 * rng.random(R)
    return connect, top, mean, var, level


def reservoir(R=1000, rain='normal', seed=0):
    rs = [f'r{i + 1}' for i in range(R)]
    connect, top, mean, var, level = reservoir_tensors(R, seed)
    pvars = [
        ('CONNECT', 'non-fluent', 'bool', ['res', 'res'], False),
        ('TOP', 'non-fluent', 'real', ['res'], 100.0),
        ('RAIN_MEAN', 'non-fluent', 'real', ['res'], 5.0),
        ('RAIN_VAR', 'non-fluent', 'real', ['res'], 1.0),
        ('TARGET_LEVEL', 'non-fluent', 'real', None, 50.0),
        ('rain', 'interm-fluent', 'real', ['res'], 0.0),
        ('released', 'interm-fluent', 'real', ['res'], 0.0),
        ('inflow', 'interm-fluent', 'real', ['res'], 0.0),
        ('rlevel', 'state-fluent', 'real', ['res'], 50.0),
        ('release', 'action-fluent', 'real', ['res'], 0.0),
    ]
    if rain == 'normal':
        rain_e = func('abs', rand('Normal', pv('RAIN_MEAN', '?r'), pv('RAIN_VAR', '?r')))
    else:
        rain_e = rand('Exponential', pv('RAIN_MEAN', '?r'))
    cpfs = [
        (('rain', ['?r']), rain_e),
        (('released', ['?r']), func('max', num(0.0), func('min', pv('rlevel', '?r'), pv('release', '?r')))),
        (('inflow', ['?r']), agg('sum', [('?u', 'res')], mul(pv('CONNECT', '?u', '?r'), pv('released', '?u')))),
        (("rlevel'", ['?r']), func('min', pv('TOP', '?r'),
                                   func('max', num(0.0),
                                        add(sub(add(pv('rlevel', '?r'), pv('rain', '?r')), pv('released', '?r')),
                                            pv('inflow', '?r'))))),
    ]
    reward = neg(agg('sum', [('?r', 'res')], func('abs', sub(pv("rlevel'", '?r'), pv('TARGET_LEVEL')))))
    init_nf = []
    for (u, v) in zip(*np.nonzero(connect)):
        init_nf.append((('CONNECT', [rs[u], rs[v]]), True))
    for i in range(R):
        init_nf.append((('TOP', [rs[i]]), float(top[i])))
        init_nf.append((('RAIN_MEAN', [rs[i]]), float(mean[i])))
        init_nf.append((('RAIN_VAR', [rs[i]]), float(var[i])))
    init_state = [(('rlevel', [rs[i]]), float(level[i])) for i in range(R)]
    return dict(
        name='reservoir_' + rain,
        types=[('res', 'object')],
        pvariables=pvars, cpfs=cpfs, reward=reward,
        preconds=[agg('forall', [('?r', 'res')], ge(pv('release', '?r'), num(0.0)))],
        invariants=[agg('forall', [('?r', 'res')], ge(pv('rlevel', '?r'), num(0.0)))],
        terminals=[],
        objects=[('res', rs)],
        init_non_fluent=init_nf, init_state=init_state,
        max_nondef_actions='pos-inf', horizon=100, discount=1.0,
    )


# ---------------------------------------------------------------------------
# C5: object-pair sum-product contraction + state invariants -- synthetic
# ---------------------------------------------------------------------------

def pair_contraction(X=512, seed=0):
    rng = np.random.default_rng(seed)
    W = rng.standard_normal((X, X)) / math.sqrt(X)
    x0 = rng.standard_normal(X)
    ns = [f'n{i + 1}' for i in range(X)]
    pvars = [
        ('W', 'non-fluent', 'real', ['node', 'node'], 0.0),
        ('BOUND', 'non-fluent', 'real', None, 10.0),
        ('GAIN', 'non-fluent', 'real', None, 0.1),
        ('x', 'state-fluent', 'real', ['node'], 0.0),
        ('u', 'action-fluent', 'real', ['node'], 0.0),
    ]
    cpfs = [
        (("x'", ['?b']), add(func('tanh', agg('sum', [('?a', 'node')], mul(pv('W', '?a', '?b'), pv('x', '?a')))),
                             mul(pv('GAIN'), pv('u', '?b')))),
    ]
    reward = neg(agg('sum', [('?b', 'node')], mul(pv("x'", '?b'), pv("x'", '?b'))))
    init_nf = [(('W', [ns[a], ns[b]]), float(W[a, b])) for a in range(X) for b in range(X)]
    init_state = [(('x', [ns[a]]), float(x0[a])) for a in range(X)]
    return dict(
        name='pair_contraction',
        types=[('node', 'object')],
        pvariables=pvars, cpfs=cpfs, reward=reward,
        preconds=[agg('forall', [('?b', 'node')], land(ge(pv('u', '?b'), neg(pv('BOUND'))),
                                                        le(pv('u', '?b'), pv('BOUND'))))],
        invariants=[agg('forall', [('?b', 'node')], le(pv('x', '?b'), pv('BOUND'))),
                    agg('forall', [('?b', 'node')], ge(pv('x', '?b'), neg(pv('BOUND'))))],
        terminals=[],
        objects=[('node', ns)],
        init_non_fluent=init_nf, init_state=init_state,
        max_nondef_actions='pos-inf', horizon=100, discount=1.0,
    )


# ---------------------------------------------------------------------------
# operator zoo: one small lifted domain touching every evaluator family of
# /root/reference/pyRDDLGym/core/simulator.py:584-1268 that this backend lowers
# ---------------------------------------------------------------------------

def opzoo(n=5, m=4, seed=3):
    rng = np.random.default_rng(seed)
    As = [f'a{i + 1}' for i in range(n)]
    Bs = [f'b{i + 1}' for i in range(m)]
    colors = ['@red', '@green', '@blue']
    P = lambda name, *a: pv(name, *a)
    pvars = [
        ('K', 'non-fluent', 'real', ['ta'], 1.0),
        ('M', 'non-fluent', 'real', ['ta', 'tb'], 0.0),
        ('ADJ', 'non-fluent', 'bool', ['ta', 'ta'], False),
        ('NEXT', 'non-fluent', 'ta', ['ta'], None),
        ('IK', 'non-fluent', 'int', ['ta'], 1),
        ('CW', 'non-fluent', 'real', ['color'], 1.0),
        ('fa', 'state-fluent', 'real', ['ta'], 0.5),
        ('ia', 'state-fluent', 'int', ['ta'], 2),
        ('ba', 'state-fluent', 'bool', ['ta'], False),
        ('fab', 'state-fluent', 'real', ['ta', 'tb'], 0.25),
        ('sq', 'state-fluent', 'real', ['ta', 'ta'], 0.0),
        ('col', 'state-fluent', 'color', ['ta'], '@red'),
        ('ptr', 'state-fluent', 'ta', ['ta'], None),
        ('s0', 'state-fluent', 'real', None, 1.5),
        ('act', 'action-fluent', 'real', ['ta'], 0.0),
        ('bact', 'action-fluent', 'bool', ['ta'], False),
        ('iact', 'action-fluent', 'int', None, 0),
        ('t-arith', 'interm-fluent', 'real', ['ta'], 0.0),
        ('t-int', 'interm-fluent', 'int', ['ta'], 0),
        ('t-rel', 'interm-fluent', 'bool', ['ta'], False),
        ('t-logic', 'interm-fluent', 'bool', ['ta'], False),
        ('t-func', 'interm-fluent', 'real', ['ta'], 0.0),
        ('t-ifunc', 'interm-fluent', 'int', ['ta'], 0),
        ('t-agg', 'interm-fluent', 'real', ['ta'], 0.0),
        ('t-aggb', 'interm-fluent', 'bool', ['ta'], False),
        ('t-arg', 'interm-fluent', 'tb', ['ta'], None),
        ('t-nest', 'interm-fluent', 'real', ['ta'], 0.0),
        ('t-diag', 'interm-fluent', 'real', ['ta'], 0.0),
        ('t-trans', 'interm-fluent', 'real', ['tb', 'ta'], 0.0),
        ('t-switch', 'interm-fluent', 'real', ['ta'], 0.0),
        ('t-lit', 'interm-fluent', 'real', None, 0.0),
        ('t-scal', 'interm-fluent', 'real', None, 0.0),
        ('t-free', 'interm-fluent', 'bool', ['ta', 'ta'], False),
        ('t-rand', 'interm-fluent', 'real', ['ta'], 0.0),
        ('t-tot', 'interm-fluent', 'real', None, 0.0),
    ]
    a = '?a'
    fa, ia, ba = P('fa', a), P('ia', a), P('ba', a)
    t_arith = add(sub(mul(fa, P('K', a)), div(ia, num(3))), mul(neg(fa), ba))
    t_int = sub(add(mul(ia, num(3)), ba), P('IK', a))
    t_rel = lor(land(lt(fa, num(0.5)), ge(ia, num(2))), lor(eq(ia, P('IK', a)), ne(fa, P('act', a))))
    t_logic = xor(implies(ba, P('bact', a)), equiv(lnot(ba), land(P('t-rel', a), lor(ba, P('bact', a)))))
    t_func = sum_all([
        func('abs', fa), func('cos', fa), func('sin', fa), func('tan', mul(fa, num(0.3))),
        func('acos', mul(func('tanh', fa), num(0.9))), func('asin', mul(func('tanh', fa), num(0.9))),
        func('atan', fa), func('cosh', mul(fa, num(0.2))), func('sinh', mul(fa, num(0.2))),
        func('exp', mul(fa, num(0.1))), func('ln', add(func('abs', fa), num(1.0))),
        func('sqrt', add(func('abs', fa), num(0.25))),
        func('lngamma', add(func('abs', fa), num(0.5))), func('gamma', add(func('abs', fa), num(1.5))),
        func('min', fa, P('act', a)), func('max', fa, ia), func('pow', add(func('abs', fa), num(0.1)), num(1.7)),
        func('log', add(func('abs', fa), num(1.1)), num(2.0)), func('hypot', fa, P('act', a)),
        func('fmod', fa, num(0.7)), func('fmod', neg(fa), num(0.7)),
    ])
    t_ifunc = sum_all([
        func('sgn', fa), func('round', mul(fa, num(2.5))), func('floor', mul(fa, num(3.0))),
        func('ceil', mul(fa, num(3.0))), func('div', sub(ia, num(7)), num(2)), func('mod', sub(ia, num(7)), num(3)),
        func('abs', sub(ia, num(9))), func('sgn', sub(ia, num(2))), func('min', ia, num(3)), func('max', ia, P('IK', a)),
        func('pow', ia, num(2)), func('fmod', sub(ia, num(7)), num(3)),
    ])
    b = '?b'
    t_agg = sum_all([
        agg('sum', [(b, 'tb')], mul(P('M', a, b), P('fab', a, b))),
        agg('avg', [(b, 'tb')], P('fab', a, b)),
        agg('prod', [(b, 'tb')], add(num(1.0), mul(num(0.1), P('fab', a, b)))),
        agg('min', [(b, 'tb')], P('fab', a, b)),
        agg('max', [(b, 'tb')], P('M', a, b)),
        agg('sum', [('?a2', 'ta')], land(P('ADJ', a, '?a2'), P('ba', '?a2'))),
        agg('sum', [('?a2', 'ta'), (b, 'tb')], mul(P('ADJ', a, '?a2'), P('fab', '?a2', b))),
        agg('avg', [('?a2', 'ta')], P('ba', '?a2')),
    ])
    t_aggb = lor(agg('exists', [('?a2', 'ta')], land(P('ADJ', a, '?a2'), P('ba', '?a2'))),
                 agg('forall', [(b, 'tb')], gt(P('fab', a, b), num(0.0))))
    t_arg = agg('argmax', [(b, 'tb')], P('fab', a, b))
    t_nest = add(add(P('fa', P('NEXT', a)), P('fa', P('ptr', a))),
                 add(P('fab', a, P('t-arg', a)), P('fab', P('NEXT', a), agg('argmin', [(b, 'tb')], P('fab', a, b)))))
    t_diag = add(P('sq', a, a), agg('sum', [('?a2', 'ta')], mul(P('sq', '?a2', a), P('sq', a, '?a2'))))
    t_trans = add(P('fab', a, b), P('M', a, b))
    t_switch = switch(P('col', a), [('@red', fa), ('@blue', mul(fa, num(2.0)))], default=P('CW', '@green'))
    t_lit = add(P('CW', '@blue'), mul(P('s0'), num(2)))
    t_scal = ite(gt(P('s0'), num(1.0)), add(P('s0'), P('iact')), sub(P('s0'), P('t-lit')))
    t_free = lor(eq(pv('?a'), pv('?c')), land(eq(P('col', '?a'), pv('@green')), eq(P('ptr', '?a'), pv('?c'))))
    t_rand = sum_all([
        rand('Uniform', fa, add(fa, num(1.0))),
        rand('Normal', fa, add(func('abs', fa), num(0.1))),
        rand('Exponential', add(func('abs', fa), num(0.2))),
        ite(rand('Bernoulli', func('min', num(1.0), func('abs', mul(fa, num(0.5))))), num(1.0), num(0.0)),
        rand('Weibull', num(1.5), add(func('abs', fa), num(0.2))),
        rand('Gumbel', fa, num(0.5)),
        rand('Laplace', fa, num(0.5)),
        rand('Cauchy', fa, num(0.01)),
        rand('Gompertz', num(1.2), num(0.8)),
        rand('Kumaraswamy', num(2.0), num(3.0)),
        rand('Pareto', num(3.0), num(1.0)),
        rand('DiracDelta', fa),
        mul(num(1.0), rand('KronDelta', ia)),
    ])
    t_tot = sum_all([agg('sum', [(a, 'ta')], P('t-arith', a)), agg('max', [(a, 'ta')], P('t-func', a)),
                     agg('sum', [(a, 'ta'), (b, 'tb')], P('t-trans', b, a)), P('t-scal'),
                     agg('avg', [(a, 'ta')], P('t-ifunc', a))])
    cpfs = [
        (('t-arith', [a]), t_arith), (('t-int', [a]), t_int), (('t-rel', [a]), t_rel),
        (('t-logic', [a]), t_logic), (('t-func', [a]), t_func), (('t-ifunc', [a]), t_ifunc),
        (('t-agg', [a]), t_agg), (('t-aggb', [a]), t_aggb), (('t-arg', [a]), t_arg),
        (('t-nest', [a]), t_nest), (('t-diag', [a]), t_diag), (('t-trans', [b, a]), t_trans),
        (('t-switch', [a]), t_switch), (('t-lit', None), t_lit), (('t-scal', None), t_scal),
        (('t-free', ['?a', '?c']), t_free), (('t-rand', [a]), t_rand), (('t-tot', None), t_tot),
        (("fa'", [a]), func('tanh', add(mul(fa, num(0.9)), mul(num(0.05), add(P('t-arith', a), P('act', a)))))),
        (("ia'", [a]), func('max', num(-5), func('min', num(9), add(ia, sub(P('t-int', a), mul(num(3), ia)))))),
        (("ba'", [a]), ite(P('bact', a), lnot(ba), xor(P('t-logic', a), P('t-aggb', a)))),
        (("fab'", [a, b]), add(mul(P('fab', a, b), num(0.5)), mul(num(0.01), add(P('t-trans', b, a), P('t-nest', a))))),
        (("sq'", [a, '?c']), ite(P('t-free', a, '?c'), add(P('sq', '?c', a), num(0.125)), mul(P('sq', a, '?c'), num(0.5)))),
        (("col'", [a]), discrete('color', [('@red', num(0.2)), ('@green', add(num(0.3), mul(num(0.0), fa))),
                                            ('@blue', num(0.5))])),
        (("ptr'", [a]), ite(ba, P('NEXT', a), P('ptr', P('NEXT', a)))),
        (("s0'", None), add(mul(P('s0'), num(0.99)), mul(num(0.001), P('t-tot')))),
    ]
    reward = add(add(P('t-tot'), agg('sum', [(a, 'ta')], mul(num(-0.1), func('abs', P('act', a))))),
                 agg('sum', [(a, 'ta')], mul(P("fa'", a), P('ba', a))))
    init_nf, init_state = [], []
    Kv = rng.random(n) + 0.5
    Mv = rng.standard_normal((n, m))
    adj = rng.random((n, n)) < 0.4
    for i in range(n):
        init_nf.append((('K', [As[i]]), float(Kv[i])))
        init_nf.append((('NEXT', [As[i]]), As[(i + 1) % n]))
        init_nf.append((('IK', [As[i]]), int(rng.integers(0, 4))))
        init_state.append((('fa', [As[i]]), float(rng.standard_normal())))
        init_state.append((('ia', [As[i]]), int(rng.integers(-3, 8))))
        init_state.append((('ba', [As[i]]), bool(rng.random() < 0.5)))
        init_state.append((('ptr', [As[i]]), As[int(rng.integers(0, n))]))
        init_state.append((('col', [As[i]]), colors[int(rng.integers(0, 3))]))
        for j in range(m):
            init_nf.append((('M', [As[i], Bs[j]]), float(Mv[i, j])))
            init_state.append((('fab', [As[i], Bs[j]]), float(rng.random() + 0.05)))
        for j in range(n):
            if adj[i, j]:
                init_nf.append((('ADJ', [As[i], As[j]]), True))
            init_state.append((('sq', [As[i], As[j]]), float(rng.standard_normal() * 0.3)))
    for (c, w) in zip(colors, (0.7, 1.3, 2.1)):
        init_nf.append((('CW', [c]), w))
    return dict(
        name='opzoo',
        types=[('ta', 'object'), ('tb', 'object'), ('color', colors)],
        pvariables=pvars, cpfs=cpfs, reward=reward,
        preconds=[agg('forall', [(a, 'ta')], land(ge(P('act', a), num(-2.0)), le(P('act', a), num(2.0)))),
                  ge(P('iact'), num(0))],
        invariants=[agg('forall', [(a, 'ta')], land(ge(P('fa', a), num(-1.0)), le(P('fa', a), num(1.0)))),
                    le(P('s0'), num(100.0))],
        terminals=[gt(agg('sum', [(a, 'ta')], P('ba', a)), num(n - 1)),
                   agg('exists', [(a, 'ta')], gt(P('ia', a), num(8)))],
        objects=[('ta', As), ('tb', Bs)],
        init_non_fluent=init_nf, init_state=init_state,
        max_nondef_actions='pos-inf', horizon=20, discount=0.9,
    )


# ---------------------------------------------------------------------------
# distribution zoo: one state fluent per natively sampled distribution
# (/root/reference/pyRDDLGym/core/simulator.py:1066-1221), parameters from non-fluents
# ---------------------------------------------------------------------------

DISTZOO_PARAMS = {
    'poisson-small': ('Poisson', (3.5,)), 'poisson-large': ('Poisson', (47.0,)),
    'gamma-small': ('Gamma', (0.6, 2.0)), 'gamma-large': ('Gamma', (4.5, 0.5)),
    'binomial-small': ('Binomial', (12, 0.3)), 'binomial-large': ('Binomial', (400, 0.62)),
    'negbinomial': ('NegativeBinomial', (5.0, 0.4)), 'beta': ('Beta', (2.0, 5.0)),
    'geometric': ('Geometric', (0.2,)), 'student': ('Student', (7.0,)), 'chisquare': ('ChiSquare', (3.0,)),
}


def distzoo(n=4):
    objs = [f'o{i + 1}' for i in range(n)]
    pvars, cpfs = [], []
    for key, (dist, params) in DISTZOO_PARAMS.items():
        is_int = dist in ('Poisson', 'Binomial', 'NegativeBinomial', 'Geometric')
        pvars.append((key, 'state-fluent', 'int' if is_int else 'real', ['obj'], 0 if is_int else 0.0))
        args = []
        for j, v in enumerate(params):
            pname = f'P-{key}-{j}'
            pvars.append((pname, 'non-fluent', 'int' if isinstance(v, int) else 'real', None, v))
            args.append(pv(pname))
        cpfs.append(((key + "'", ['?o']), rand(dist, *args)))
    pvars.append(('nudge', 'action-fluent', 'real', None, 0.0))
    reward = add(pv('nudge'), agg('sum', [('?o', 'obj')], pv("student'", '?o')))
    return dict(
        name='distzoo', types=[('obj', 'object')], pvariables=pvars, cpfs=cpfs, reward=reward,
        preconds=[], invariants=[], terminals=[], objects=[('obj', objs)], init_non_fluent=[], init_state=[],
        max_nondef_actions='pos-inf', horizon=10, discount=1.0,
    )


WORKLOADS = {
    'cartpole': cartpole,
    'wildfire': wildfire,
    'reservoir': reservoir,
    'pair_contraction': pair_contraction,
    'opzoo': opzoo,
    'distzoo': distzoo,
}


# ---------------------------------------------------------------------------
# instance tensors + compiled (lifted) programs
# ---------------------------------------------------------------------------

_DEFAULTS = {'bool': False, 'int': 0, 'real': 0.0}
_NP_RANGE = {'bool': np.bool_, 'int': np.int64, 'real': np.float64}


def instance_tensors(spec):
    """Initial tensors of every pvariable of ``spec`` -- the same values
    RDDLValueInitializer.initialize() produces
    (/root/reference/pyRDDLGym/core/compiler/initializer.py:45-96,160-214): declared
    default, overridden by the instance's init-non-fluent / init-state entries; object
    valued fluents become canonical object indices (None -> 0); C order over parameters.
    Primed copies of state fluents start from the default of their range."""
    objects = {t: list(objs) for (t, objs) in spec['objects']}
    enum_types = []
    for (t, kind) in spec['types']:
        if kind != 'object':
            objects[t] = [o.lstrip('@') for o in kind]
            enum_types.append(t)
    index = {}
    for t, objs in objects.items():
        for i, o in enumerate(objs):
            index[o.lstrip('@')] = i
    sizes = {t: len(o) for t, o in objects.items()}
    out = {}
    meta = {}
    for (name, ftype, rng, params, default) in (pv[:5] for pv in spec['pvariables']):
        shape = tuple(sizes[t] for t in (params or []))
        if rng in _NP_RANGE:
            dt = _NP_RANGE[rng]
            dv = _DEFAULTS[rng] if default is None else default
        else:
            dt = np.int64
            dv = 0 if default is None else index[str(default).lstrip('@')]
        out[name] = np.full(shape, dv, dtype=dt)
        meta[name] = (ftype, rng, params or [])
        if ftype == 'state-fluent':
            # (the primed copy is not in the instance: range default, initializer.py:160-196)
            out[name + "'"] = np.full(shape, _DEFAULTS.get(rng, 0), dtype=dt)
    for ((name, objs), value) in list(spec['init_non_fluent']) + list(spec['init_state']):
        rng = meta[name][1]
        if rng not in _NP_RANGE:
            value = index[str(value).lstrip('@')]
        pos = tuple(index[o.lstrip('@')] for o in (objs or []))
        out[name][pos] = value
    return out, objects, enum_types


_PROGRAM_DIR = __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)),
                                          'programs')


def program_path(spec_name):
    import os
    return os.path.join(_PROGRAM_DIR, spec_name + '.json')


def load_program(spec):
    """Compiled program for ``spec`` without the reference front end: the lifted HIR
    (size independent; produced once from the reference's traced model by
    oracle/gen_fixtures.py and stored in pyrddlgym_b200/programs/) re-instantiated with
    this instance's object counts and initial tensors."""
    import json
    from pyrddlgym_b200 import hir
    with open(program_path(spec['name'])) as f:
        text = f.read()
    init, objects, enum_types = instance_tensors(spec)
    p = hir.Program.from_json(text, {})
    p.types = {t: len(o) for t, o in objects.items()}
    p.objects = objects
    p.enum_types = sorted(enum_types)
    p.init = {name: init[name] for name in p.tensors}
    p.horizon = int(spec['horizon'])
    p.discount = float(spec['discount'])
    n_act = sum(int(np.prod(p.tensor_shape(n), dtype=np.int64)) for n in p.names_of_kind('action'))
    p.max_allowed_actions = n_act if spec['max_nondef_actions'] == 'pos-inf' else int(spec['max_nondef_actions'])
    return p
