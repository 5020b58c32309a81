"""Reference front end -> HIR.

Input is exactly what ``RDDLSimulator._compile`` produces
(/root/reference/pyRDDLGym/core/simulator.py:136-174): the planning model ``rddl``, the
level-ordered CPF list ``cpfs``, ``levels``, the tracer cache ``traced`` and
``init_values``. Parser, level analysis, tracer and initializer are the reference's own,
reused unchanged; this module only *reads* their output:

* per-node scope            : ``traced.cached_objects_in_scope`` (tracer.py:42-44)
* free objects / enum consts: ``cached_sim_info`` ``(True, value)`` (tracer.py:357-405)
* aggregation axes          : new iteration variables are appended as trailing axes
                              (tracer.py:786-790)
* switch / Discrete cases   : canonical object order (tracer.py:914-952)

The pvariable gather is emitted in index form straight from the argument
list -- scope axis / literal / nested expression per parameter -- which covers the
reference's NOOP, TRANSPOSE, EINSUM and NESTED_SLICE cases (tracer.py:440-623) uniformly.

dtype inference restates the NumPy promotion the reference evaluator produces
(``1 * x`` promotions simulator.py:603-604,673-674,777,794,807-808; int64 casts :96-99,
:116-117; true division :64).
"""
import numpy as np

from pyrddlgym_b200 import hir
from pyrddlgym_b200.hir import Node

_RANGE_DTYPE = {'bool': 'b', 'int': 'i', 'real': 'f'}
_KIND = {'non-fluent': 'nonfluent', 'state-fluent': 'state', 'next-state-fluent': 'next',
         'action-fluent': 'action', 'interm-fluent': 'interm', 'derived-fluent': 'derived',
         'observ-fluent': 'observ'}
_ARITH = {'+': 'add', '-': 'sub', '*': 'mul', '/': 'div'}
_REL = {'<': 'lt', '<=': 'le', '>': 'gt', '>=': 'ge', '==': 'eq', '~=': 'ne'}
_LOGIC = {'^': 'and', '&': 'and', '|': 'or', '~': 'xor', '=>': 'imply', '<=>': 'eqv'}
_AGG = {'sum': 'sum', 'avg': 'avg', 'prod': 'prod', 'minimum': 'min', 'maximum': 'max',
        'forall': 'forall', 'exists': 'exists', 'argmin': 'argmin', 'argmax': 'argmax'}
_UNARY = {'abs', 'sgn', 'round', 'floor', 'ceil', 'cos', 'sin', 'tan', 'acos', 'asin', 'atan',
          'cosh', 'sinh', 'tanh', 'exp', 'ln', 'sqrt', 'lngamma', 'gamma'}
_BINARY = {'div': 'floordiv', 'mod': 'mod', 'fmod': 'fmod', 'min': 'min', 'max': 'max',
           'pow': 'pow', 'log': 'log', 'hypot': 'hypot'}


def _exc():
    from pyRDDLGym.core.debug import exception as exc
    return exc


def promote(a, b):
    order = 'bif'
    return order[max(order.index(a), order.index(b))]


def num(t):
    """dtype after the reference's ``1 * x`` promotion."""
    return 'i' if t == 'b' else t


class _Converter:

    def __init__(self, rddl, traced):
        self.rddl = rddl
        self.traced = traced

    def scope_of(self, expr):
        objs = self.traced.cached_objects_in_scope(expr)
        return [t for (_, t) in objs]

    def fail_type(self, msg):
        raise _exc().RDDLTypeError(msg)

    def require_bool(self, node, what):
        if node.dtype != 'b':
            self.fail_type(f'{what} must evaluate to bool, got dtype {node.dtype!r}.')

    def conv(self, expr):
        etype, op = expr.etype
        scope = self.scope_of(expr)
        nid = expr.id
        rddl = self.rddl
        exc = _exc()
        if etype == 'constant':
            v = expr.args
            if isinstance(v, (bool, np.bool_)):
                return Node('const', value=bool(v), dtype='b', scope=scope, id=nid)
            if isinstance(v, (int, np.integer)):
                return Node('const', value=int(v), dtype='i', scope=scope, id=nid)
            return Node('const', value=float(v), dtype='f', scope=scope, id=nid)

        if etype == 'pvar':
            var, args = expr.args
            is_value, info = self.traced.cached_sim_info(expr)
            obj_type = self.traced.cached_object_type(expr)
            if is_value:
                objs = self.traced.cached_objects_in_scope(expr)
                if rddl.is_free_object(var):
                    k = [v for (v, _) in objs].index(var)
                    return Node('axis', value=k, dtype='i', scope=scope, id=nid, obj_type=obj_type)
                lit = rddl.strip_literal(var)
                return Node('const', value=int(rddl.object_to_index[lit]), dtype='i',
                            scope=scope, id=nid, obj_type=obj_type)
            objs = self.traced.cached_objects_in_scope(expr)
            names = [v for (v, _) in objs]
            index = []
            for arg in (args or []):
                if hasattr(arg, 'etype'):
                    index.append(('e', self.conv(arg)))
                elif arg == '_':
                    raise exc.RDDLNotImplementedError(
                        f'Random-vector placeholder _ in <{var}> is not supported by the B200 backend.')
                elif arg in names and rddl.is_free_object(arg):
                    index.append(('a', names.index(arg)))
                else:
                    index.append(('c', int(rddl.object_to_index[rddl.strip_literal(arg)])))
            prange = rddl.variable_ranges[var]
            return Node('load', tensor=var, index=index, dtype=_RANGE_DTYPE.get(prange, 'i'),
                        scope=scope, id=nid, obj_type=obj_type)

        if etype == 'arithmetic':
            args = [self.conv(a) for a in expr.args]
            n = len(args)
            if n == 1 and op == '-':
                return Node('un', op='neg', args=args, dtype=num(args[0].dtype), scope=scope, id=nid)
            if n == 2:
                a, b = args
                dt = 'f' if op == '/' else promote(num(a.dtype), num(b.dtype))
                return Node('bin', op=_ARITH[op], args=args, dtype=dt, scope=scope, id=nid)
            if n > 0 and not scope and op in ('*', '+'):
                dt = 'i'
                for a in args:
                    dt = promote(dt, num(a.dtype))
                return Node('nary', op=_ARITH[op], args=args, dtype=dt, scope=scope, id=nid)
            raise exc.RDDLInvalidNumberOfArgumentsError(
                f'Arithmetic operator {op} does not have the required number of arguments.')

        if etype == 'relational':
            args = [self.conv(a) for a in expr.args]
            if len(args) != 2:
                raise exc.RDDLInvalidNumberOfArgumentsError(f'{op} requires 2 argument(s).')
            return Node('bin', op=_REL[op], args=args, dtype='b', scope=scope, id=nid)

        if etype == 'boolean':
            args = [self.conv(a) for a in expr.args]
            n = len(args)
            for a in args:
                self.require_bool(a, f'Argument of {op}')
            if n == 1 and op == '~':
                return Node('un', op='not', args=args, dtype='b', scope=scope, id=nid)
            if n == 2:
                return Node('bin', op=_LOGIC[op], args=args, dtype='b', scope=scope, id=nid)
            if n > 0 and op in ('^', '&', '|') and not scope:
                return Node('nary', op=_LOGIC[op], args=args, dtype='b', scope=scope, id=nid)
            raise exc.RDDLInvalidNumberOfArgumentsError(
                f'Logical operator {op} does not have the required number of arguments.')

        if etype == 'aggregation':
            *pvars, arg = expr.args
            red = [t for (_, (_, t)) in pvars]
            body = self.conv(arg)
            aop = _AGG[op]
            if aop in ('forall', 'exists'):
                self.require_bool(body, f'Argument of {op}')
                dt = 'b'
            elif aop in ('argmin', 'argmax'):
                dt = 'i'
            elif aop == 'avg':
                dt = 'f'
            else:
                dt = num(body.dtype)
            return Node('agg', op=aop, args=[body], red=red, dtype=dt, scope=scope, id=nid,
                        obj_type=self.traced.cached_object_type(expr))

        if etype == 'func':
            args = [self.conv(a) for a in expr.args]
            if op in _UNARY:
                if len(args) != 1:
                    raise exc.RDDLInvalidNumberOfArgumentsError(f'{op} requires 1 argument(s).')
                t = num(args[0].dtype)
                if op == 'abs':
                    dt = t
                elif op in hir.UNARY_TO_INT:
                    dt = 'i'
                else:
                    dt = 'f'
                return Node('un', op=op, args=args, dtype=dt, scope=scope, id=nid)
            if op in _BINARY:
                if len(args) != 2:
                    raise exc.RDDLInvalidNumberOfArgumentsError(f'{op} requires 2 argument(s).')
                ta, tb = num(args[0].dtype), num(args[1].dtype)
                bop = _BINARY[op]
                if bop in ('floordiv', 'mod'):
                    if ta != 'i' or tb != 'i':
                        self.fail_type(f'Arguments of {op} must evaluate to int.')
                    dt = 'i'
                elif bop in ('fmod', 'min', 'max', 'pow'):
                    dt = promote(ta, tb)
                else:
                    dt = 'f'
                return Node('bin', op=bop, args=args, dtype=dt, scope=scope, id=nid)
            raise exc.RDDLNotImplementedError(f'Function {op} is not supported.')

        if etype == 'control':
            if op == 'if':
                c, a, b = [self.conv(x) for x in expr.args]
                self.require_bool(c, 'If predicate')
                return Node('if', args=[c, a, b], dtype=promote(a.dtype, b.dtype), scope=scope,
                            id=nid, obj_type=self.traced.cached_object_type(expr))
            pred, *_ = expr.args
            cases, default = self.traced.cached_sim_info(expr)
            p = self.conv(pred)
            cs = [None if c is None else self.conv(c) for c in cases]
            df = None if default is None else self.conv(default)
            dt = 'b'
            for c in cs + [df]:
                if c is not None:
                    dt = promote(dt, c.dtype)
            return Node('switch', args=[p], cases=cs, default=df, dtype=dt, scope=scope, id=nid,
                        obj_type=self.traced.cached_object_type(expr))

        if etype == 'randomvar':
            return self.conv_random(expr, op, scope, nid)

        raise exc.RDDLNotImplementedError(
            f'Expression type {etype} ({op}) is not supported by the B200 backend '
            f'(no CPU fallback).')

    def conv_random(self, expr, name, scope, nid):
        exc = _exc()
        if name in ('KronDelta', 'DiracDelta'):
            arg, = expr.args
            a = self.conv(arg)
            if name == 'KronDelta' and a.dtype == 'f':
                self.fail_type('Argument of KronDelta must evaluate to bool or int.')
            return Node('delta', op=name, args=[a], dtype=a.dtype, scope=scope, id=nid,
                        obj_type=a.obj_type)
        if name in ('Discrete', 'UnnormDiscrete'):
            sorted_args = self.traced.cached_sim_info(expr)
            cs = [self.conv(a) for a in sorted_args]
            return Node('rand', op=name, cases=cs, dtype='i', scope=scope, id=nid,
                        obj_type=self.traced.cached_object_type(expr))
        if name in ('Discrete(p)', 'UnnormDiscrete(p)'):
            *pvars, args = expr.args
            (_, (_, typ)), = pvars
            body, = args
            return Node('rand', op=name, args=[self.conv(body)], red=[typ], dtype='i', scope=scope,
                        id=nid, obj_type=self.traced.cached_object_type(expr))
        if name in hir.DIST_NATIVE:
            arity, dt = hir.DIST_NATIVE[name]
            args = [self.conv(a) for a in expr.args]
            if len(args) != arity:
                raise exc.RDDLInvalidNumberOfArgumentsError(f'{name} requires {arity} argument(s).')
            if name == 'Binomial' and args[0].dtype == 'f':
                self.fail_type('Binomial count must evaluate to int.')
            return Node('rand', op=name, args=args, dtype=dt, scope=scope, id=nid)
        if name not in hir.DIST_BASE:
            raise exc.RDDLNotImplementedError(
                f'Distribution {name} is not supported by the B200 backend (no CPU fallback).')
        args = [self.conv(a) for a in expr.args]
        arity = {'Bernoulli': 1, 'Exponential': 1}.get(name, 2)
        if len(args) != arity:
            raise exc.RDDLInvalidNumberOfArgumentsError(f'{name} requires {arity} argument(s).')
        return Node('rand', op=name, args=args, dtype=hir.DIST_DTYPE.get(name, 'f'),
                    scope=scope, id=nid)


def program_from_reference(rddl, cpfs, levels, traced, init_values, name=None):
    """Builds the HIR ``Program`` from the reference's compiled artefacts."""
    p = hir.Program()
    p.name = name or getattr(rddl, 'domain_name', None) or 'rddl'
    p.types = {t: len(objs) for (t, objs) in rddl.type_to_objects.items()}
    p.objects = {t: list(objs) for (t, objs) in rddl.type_to_objects.items()}
    p.enum_types = sorted(rddl.enum_types)
    for (var, vtype) in rddl.variable_types.items():
        prange = rddl.variable_ranges[var]
        p.tensors[var] = hir.TensorDecl(var, _KIND[vtype], _RANGE_DTYPE.get(prange, 'i'),
                                        rddl.variable_params[var], prange)
    for (var, value) in init_values.items():
        t = p.tensors[var]
        p.init[var] = np.ascontiguousarray(np.asarray(value, dtype=hir.NP_DTYPE[t.dtype]))
    # next-state buffers start from the declared default (they are written before being read)
    cv = _Converter(rddl, traced)
    cpf_level = {c: lvl for (lvl, cs) in levels.items() for c in cs}
    for (cpf, expr, _) in cpfs:
        objects, _ = rddl.cpfs[cpf]
        scope = [t for (_, t) in objects]
        node = cv.conv(expr)
        decl = p.tensors[cpf]
        order = 'bif'
        if order.index(node.dtype) > order.index(decl.dtype):
            # simulator.py:190-201,455 -- np.can_cast(sample, declared dtype) must hold
            raise _exc().RDDLTypeError(
                f'{cpf} must evaluate to {hir.NP_DTYPE[decl.dtype]}, got dtype {node.dtype!r}.')
        p.cpfs.append(hir.Root(cpf, cpf_level[cpf], scope, node))
    p.reward = cv.conv(rddl.reward)
    for (lst, exprs, what) in [(p.invariants, rddl.invariants, 'Invariant'),
                               (p.preconditions, rddl.preconditions, 'Precondition'),
                               (p.terminations, rddl.terminations, 'Termination')]:
        for i, e in enumerate(exprs):
            n = cv.conv(e)
            if n.dtype != 'b' or n.scope:
                raise _exc().RDDLTypeError(f'{what} {i} must evaluate to a scalar bool.')
            lst.append(n)
    p.next_state = dict(rddl.next_state)
    p.horizon = int(rddl.horizon)
    p.discount = float(rddl.discount)
    p.max_allowed_actions = int(rddl.max_allowed_actions)
    return p


def program_from_simulator(sim, name=None):
    """``sim`` is any reference ``RDDLSimulator`` instance (its ``_compile`` has run)."""
    return program_from_reference(sim.rddl, sim.cpfs, sim.levels, sim.traced, sim.init_values,
                                  name=name)


# ---------------------------------------------------------------------------------------
# front end for huge instances
# ---------------------------------------------------------------------------------------

def _literal_objects(node, out):
    """Object / enum literals (@name) written in an expression tree of a neutral spec."""
    if isinstance(node, str):
        if node.startswith('@'):
            out.add(node[1:])
    elif isinstance(node, (tuple, list)):
        for x in node:
            _literal_objects(x, out)


def program_from_spec(spec, surrogate_objects=3, name=None):
    """HIR ``Program`` of a neutral domain spec (``workloads`` / ``rddl_reader.parse_rddl`` format) with
    front-end cost O(#instance entries) instead of O(#groundings).

    The reference front end materialises every grounding *name* of every pvariable
    (RDDLLiftedModel, /root/reference/pyRDDLGym/core/compiler/model.py:915-923) and fills the initial
    tensors through per-grounding Python lists (initializer.py:160-196): minutes and gigabytes at a
    million cells, and 2^40 strings for a 4-parameter mask over a 1024 x 1024 grid. But everything the
    lowering needs from it -- CPF levels, traced scopes, types, switch / Discrete case orders -- is a
    property of the *lifted* domain, not of how many objects the instance has. So the unmodified
    reference front end (parse tree -> RDDLLiftedModel -> RDDLLevelAnalysis -> RDDLObjectsTracer) runs
    on a surrogate instance that keeps, per object type, the first ``surrogate_objects`` objects (and
    every object an expression names literally), and the resulting lifted program is re-instantiated
    with the real object lists and with initial tensors built directly from the instance's entries
    (``workloads.instance_tensors``: the values RDDLValueInitializer would produce). Equal, op table
    byte for byte, to the program of the full front end wherever that can run
    (tests/test_frontend_scaling.py)."""
    from pyRDDLGym.core.compiler.model import RDDLLiftedModel
    from pyRDDLGym.core.simulator import RDDLSimulator
    from pyrddlgym_b200 import rddl_reader, workloads
    lits = set()
    for (_, e) in spec['cpfs']:
        _literal_objects(e, lits)
    for e in [spec['reward']] + list(spec.get('preconds', [])) + list(spec.get('invariants', [])) + \
            list(spec.get('terminals', [])) + list(spec.get('constraints', [])):
        _literal_objects(e, lits)
    keep = {}
    for (t, objs) in spec['objects']:
        k = min(len(objs), max(1, int(surrogate_objects)))
        for i, o in enumerate(objs):
            if o in lits:
                k = max(k, i + 1)
        keep[t] = list(objs[:k])
    kept = set(o for objs in keep.values() for o in objs)
    type_of = {o: t for (t, objs) in spec['objects'] for o in objs}

    def shrink(entries):
        # entries whose arguments lie in the surrogate; an object *value* outside it is replaced by the
        # type's first object (object-valued fluents have no default, model.py:975-1040: every surrogate
        # grounding still needs a value, and which one does not matter for the lifted program)
        out = []
        for ((var, objs), value) in entries:
            if not all(o.startswith('@') or o in kept for o in (objs or [])):
                continue
            if isinstance(value, str) and not value.startswith('@') and value not in kept and value in type_of:
                value = keep[type_of[value]][0]
            out.append(((var, objs), value))
        return out

    small = dict(spec)
    small['objects'] = [(t, keep[t]) for (t, _) in spec['objects']]
    small['init_non_fluent'] = shrink(spec['init_non_fluent'])
    small['init_state'] = shrink(spec['init_state'])
    sim = RDDLSimulator(RDDLLiftedModel(rddl_reader.reference_ast(small)), keep_tensors=True)
    p = program_from_simulator(sim, name=name or spec['name'])
    init, objects, enum_types = workloads.instance_tensors(spec)
    p.types = {t: len(o) for t, o in objects.items()}
    p.objects = objects
    p.enum_types = sorted(enum_types)
    p.init = {n: init[n] for n in p.tensors}
    p.horizon = int(spec['horizon'])
    p.discount = float(spec['discount'])
    n_act = sum(int(np.prod(p.tensor_shape(n), dtype=np.int64)) for n in p.names_of_kind('action'))
    p.max_allowed_actions = n_act if spec['max_nondef_actions'] == 'pos-inf' else int(spec['max_nondef_actions'])
    return p
