"""HIR: the typed, scope-annotated expression DAG that the reference's traced model is
converted into (``frontend.py``) and that ``lowering.py`` turns into the op table the
CUDA level interpreter executes.

A ``Program`` is self-contained (tensor declarations, initial values, expression roots
in the reference's evaluation order) and serialisable, which is how compiled problems
travel to machines where pyRDDLGym is not installed -- the analogue of the reference's
``RDDLSimulatorPrecompiled`` (/root/reference/pyRDDLGym/core/simulator.py:1444-1496).

dtypes follow /root/reference/pyRDDLGym/core/compiler/initializer.py:16-23:
'b' = bool (uint8 in HBM), 'i' = int64, 'f' = float64. Object/enum valued fluents are
'i' (canonical object indices, initializer.py:68-79).
"""
import io
import json
from typing import Any, Dict, List, Optional

import numpy as np

NP_DTYPE = {'b': np.bool_, 'i': np.int64, 'f': np.float64}

# ops by family --------------------------------------------------------------
UNARY_F2F = ['cos', 'sin', 'tan', 'acos', 'asin', 'atan', 'cosh', 'sinh', 'tanh', 'exp',
             'ln', 'sqrt', 'lngamma', 'gamma']
UNARY_TO_INT = ['sgn', 'round', 'floor', 'ceil']
ARITH = ['add', 'sub', 'mul', 'div']
RELATIONAL = ['lt', 'le', 'gt', 'ge', 'eq', 'ne']
LOGICAL = ['and', 'or', 'xor', 'imply', 'eqv']
BINARY_FUNC = ['floordiv', 'mod', 'fmod', 'min', 'max', 'pow', 'log', 'hypot']
AGGREGATIONS = ['sum', 'avg', 'prod', 'min', 'max', 'forall', 'exists', 'argmin', 'argmax']

# distributions and the base variate each one consumes per element:
#   'U' uniform [0,1) ; 'N' standard normal ; 'E' standard exponential ; 'C' standard cauchy
DIST_BASE = {
    'Uniform': 'U', 'Bernoulli': 'U', 'Normal': 'N', 'Exponential': 'E', 'Weibull': 'E',
    'Pareto': 'E', 'Gumbel': 'U', 'Laplace': 'U', 'Cauchy': 'C', 'Gompertz': 'U',
    'Kumaraswamy': 'U', 'Discrete': 'U', 'UnnormDiscrete': 'U',
    'Discrete(p)': 'U', 'UnnormDiscrete(p)': 'U',
}
# distributions NumPy draws by rejection / search: sampled natively (no shared base variate);
#   name -> (arity, result dtype)
DIST_NATIVE = {'Poisson': (1, 'i'), 'Gamma': (2, 'f'), 'Binomial': (2, 'i'), 'NegativeBinomial': (2, 'i'),
               'Beta': (2, 'f'), 'Geometric': (1, 'i'), 'Student': (1, 'f'), 'ChiSquare': (1, 'f')}
DIST_DTYPE = {'Bernoulli': 'b', 'Discrete': 'i', 'UnnormDiscrete': 'i',
              'Discrete(p)': 'i', 'UnnormDiscrete(p)': 'i'}


class Node:
    """One HIR expression node.

    kind: const | axis | load | un | bin | nary | agg | if | switch | rand | delta
    scope: list of object-type names; the node's value has one axis per entry
           (extent = number of objects of that type), plus the implicit env axis.
    """
    __slots__ = ('kind', 'op', 'args', 'dtype', 'scope', 'id', 'value', 'tensor',
                 'index', 'red', 'cases', 'default', 'obj_type')

    def __init__(self, kind, op=None, args=None, dtype=None, scope=None, id=None,
                 value=None, tensor=None, index=None, red=None, cases=None,
                 default=None, obj_type=None):
        self.kind = kind
        self.op = op
        self.args = args or []
        self.dtype = dtype
        self.scope = list(scope or [])
        self.id = id
        self.value = value
        self.tensor = tensor
        self.index = index          # load: list of ('a', k) | ('c', idx) | ('e', Node)
        self.red = red              # agg / discrete(p): list of reduced type names
        self.cases = cases          # switch / discrete: list of Node|None in canonical order
        self.default = default
        self.obj_type = obj_type

    # ---- traversal helpers
    def children(self):
        out = list(self.args)
        if self.index:
            out += [v for (k, v) in self.index if k == 'e']
        if self.cases:
            out += [c for c in self.cases if c is not None]
        if self.default is not None:
            out.append(self.default)
        return out

    def walk(self):
        yield self
        for c in self.children():
            yield from c.walk()

    # ---- (de)serialisation
    def to_dict(self):
        d = {'k': self.kind, 't': self.dtype, 's': self.scope}
        if self.op is not None:
            d['op'] = self.op
        if self.args:
            d['a'] = [a.to_dict() for a in self.args]
        if self.id is not None:
            d['id'] = int(self.id)
        if self.value is not None:
            v = self.value
            d['v'] = bool(v) if isinstance(v, (bool, np.bool_)) else (
                int(v) if isinstance(v, (int, np.integer)) else float(v))
        if self.tensor is not None:
            d['tn'] = self.tensor
        if self.index is not None:
            d['ix'] = [[k, (v.to_dict() if k == 'e' else int(v))] for (k, v) in self.index]
        if self.red is not None:
            d['red'] = self.red
        if self.cases is not None:
            d['cs'] = [None if c is None else c.to_dict() for c in self.cases]
        if self.default is not None:
            d['df'] = self.default.to_dict()
        if self.obj_type is not None:
            d['ot'] = self.obj_type
        return d

    @staticmethod
    def from_dict(d):
        n = Node(d['k'], op=d.get('op'), dtype=d['t'], scope=d['s'], id=d.get('id'),
                 value=d.get('v'), tensor=d.get('tn'), red=d.get('red'), obj_type=d.get('ot'))
        n.args = [Node.from_dict(a) for a in d.get('a', [])]
        if 'ix' in d:
            n.index = [(k, (Node.from_dict(v) if k == 'e' else int(v))) for (k, v) in d['ix']]
        if 'cs' in d:
            n.cases = [None if c is None else Node.from_dict(c) for c in d['cs']]
        if 'df' in d:
            n.default = Node.from_dict(d['df'])
        return n


class TensorDecl:
    """A pvariable's storage. kind: nonfluent | state | next | action | interm | derived | observ.
    Non-fluents are shared by all envs; every other kind has a leading env axis."""
    __slots__ = ('name', 'kind', 'dtype', 'params', 'range')

    def __init__(self, name, kind, dtype, params, range_):
        self.name = name
        self.kind = kind
        self.dtype = dtype
        self.params = list(params)
        self.range = range_       # 'bool' | 'int' | 'real' | object/enum type name

    def to_dict(self):
        return {'name': self.name, 'kind': self.kind, 'dtype': self.dtype,
                'params': self.params, 'range': self.range}

    @staticmethod
    def from_dict(d):
        return TensorDecl(d['name'], d['kind'], d['dtype'], d['params'], d['range'])


class Root:
    """A CPF: ``tensor[scope] = expr`` evaluated at ``level`` (levels.py:427-439 order)."""
    __slots__ = ('name', 'level', 'scope', 'expr')

    def __init__(self, name, level, scope, expr):
        self.name = name
        self.level = level
        self.scope = list(scope)
        self.expr = expr


class Program:

    def __init__(self):
        self.name = ''
        self.types: Dict[str, int] = {}            # object type -> #objects
        self.objects: Dict[str, List[str]] = {}    # object type -> object names (canonical order)
        self.enum_types: List[str] = []
        self.tensors: Dict[str, TensorDecl] = {}
        self.init: Dict[str, np.ndarray] = {}      # non-fluents, state, action (noop), interm defaults
        self.cpfs: List[Root] = []                 # evaluation order
        self.reward: Optional[Node] = None
        self.invariants: List[Node] = []
        self.preconditions: List[Node] = []
        self.terminations: List[Node] = []
        self.next_state: Dict[str, str] = {}       # state -> state'
        self.horizon = 0
        self.discount = 1.0
        self.max_allowed_actions = 0

    # ---- shape helpers
    def shape_of(self, scope):
        return tuple(self.types[t] for t in scope)

    def tensor_shape(self, name):
        return self.shape_of(self.tensors[name].params)

    def names_of_kind(self, *kinds):
        return [n for (n, t) in self.tensors.items() if t.kind in kinds]

    def all_roots(self):
        """(label, node) for every expression root in reference trace order
        (tracer.py:139-144)."""
        for r in self.cpfs:
            yield (f'cpf:{r.name}', r.expr)
        yield ('reward', self.reward)
        for i, e in enumerate(self.invariants):
            yield (f'invariant:{i}', e)
        for i, e in enumerate(self.preconditions):
            yield (f'precondition:{i}', e)
        for i, e in enumerate(self.terminations):
            yield (f'termination:{i}', e)

    def random_nodes(self):
        out = []
        for _, root in self.all_roots():
            for n in root.walk():
                if n.kind == 'rand':
                    out.append(n)
        return out

    # ---- (de)serialisation
    def to_json(self):
        d = {
            'name': self.name, 'types': self.types, 'objects': self.objects,
            'enum_types': self.enum_types,
            'tensors': [t.to_dict() for t in self.tensors.values()],
            'cpfs': [{'name': r.name, 'level': r.level, 'scope': r.scope,
                      'expr': r.expr.to_dict()} for r in self.cpfs],
            'reward': self.reward.to_dict(),
            'invariants': [e.to_dict() for e in self.invariants],
            'preconditions': [e.to_dict() for e in self.preconditions],
            'terminations': [e.to_dict() for e in self.terminations],
            'next_state': self.next_state, 'horizon': self.horizon,
            'discount': self.discount, 'max_allowed_actions': self.max_allowed_actions,
        }
        return json.dumps(d)

    @staticmethod
    def from_json(s, init):
        d = json.loads(s)
        p = Program()
        p.name = d['name']
        p.types = {k: int(v) for k, v in d['types'].items()}
        p.objects = d.get('objects', {})
        p.enum_types = d.get('enum_types', [])
        for t in d['tensors']:
            p.tensors[t['name']] = TensorDecl.from_dict(t)
        p.cpfs = [Root(c['name'], c['level'], c['scope'], Node.from_dict(c['expr']))
                  for c in d['cpfs']]
        p.reward = Node.from_dict(d['reward'])
        p.invariants = [Node.from_dict(e) for e in d['invariants']]
        p.preconditions = [Node.from_dict(e) for e in d['preconditions']]
        p.terminations = [Node.from_dict(e) for e in d['terminations']]
        p.next_state = d['next_state']
        p.horizon = d['horizon']
        p.discount = d['discount']
        p.max_allowed_actions = d['max_allowed_actions']
        p.init = dict(init)
        return p

    def save(self, path):
        arrays = {f'init__{k}': np.asarray(v) for k, v in self.init.items()}
        np.savez_compressed(path, hir=np.frombuffer(self.to_json().encode(), dtype=np.uint8),
                            **arrays)

    @staticmethod
    def load(path):
        with np.load(path, allow_pickle=False) as z:
            s = bytes(z['hir']).decode()
            init = {k[len('init__'):]: z[k] for k in z.files if k.startswith('init__')}
        return Program.from_json(s, init)
