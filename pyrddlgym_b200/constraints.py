"""Box bounds of action / state fluents from the action preconditions and state invariants -- the
counterpart of /root/reference/pyRDDLGym/core/constraints.py (RDDLConstraints :107-229, which RDDLEnv
turns into its gym spaces, env.py:152-193) on the HIR, with the bounds kept as device tensors for the
batched environment.

Like the reference, only constraints of the shape  [forall_{...}] [^] fluent(args) <op> rhs  are box
constraints: <op> in <=, <, >=, > (strict ones move the limit by ``inequality_tol``), rhs a
deterministic function of non-fluents only -- evaluated once on the host with the NumPy ufuncs the
reference applies (its ``sim._sample(const_expr, sim.subs)``, constraints.py:181-191); anything else is
reported as "not a box constraint" and ignored.
"""
import numpy as np

from pyrddlgym_b200 import hir

_REL = {'le': ('<=', 0.0), 'lt': ('<', 1.0), 'ge': ('>=', 0.0), 'gt': ('>', 1.0)}


class _NonFluentEval:
    """NumPy value of a non-fluent-only expression over a scope (one axis per scope entry)."""

    def __init__(self, program, scope):
        self.p = program
        self.ext = program.shape_of(scope)
        self.rank = len(scope)

    def ok(self, n):
        if n.kind in ('const', 'axis'):
            return True
        if n.kind == 'load':
            return self.p.tensors[n.tensor].kind == 'nonfluent' and all(k in ('a', 'c') for (k, _) in n.index)
        if n.kind in ('un', 'bin', 'nary', 'if'):
            return all(self.ok(a) for a in n.args)
        return False

    def _ax(self, k):
        shp = [1] * self.rank
        shp[k] = self.ext[k]
        return np.arange(self.ext[k], dtype=np.int64).reshape(shp)

    def ev(self, n):
        one = lambda x: x.astype(np.int64) if x.dtype == np.bool_ else x
        if n.kind == 'const':
            return np.asarray(n.value, dtype=hir.NP_DTYPE[n.dtype]).reshape((1,) * self.rank)
        if n.kind == 'axis':
            return self._ax(n.value)
        if n.kind == 'load':
            V = np.asarray(self.p.init[n.tensor])
            if not n.index:
                return V.reshape((1,) * self.rank)
            idx = [self._ax(v) if k == 'a' else np.asarray(v).reshape((1,) * self.rank) for (k, v) in n.index]
            return V.reshape(self.p.tensor_shape(n.tensor))[tuple(idx)]
        args = [self.ev(a) for a in n.args]
        op = n.op
        if n.kind == 'if':
            return np.where(args[0], args[1], args[2])
        if n.kind == 'un':
            a = args[0]
            table = {'not': lambda x: ~x, 'neg': lambda x: -1 * one(x), 'abs': lambda x: np.abs(one(x)),
                     'sgn': lambda x: np.sign(one(x)).astype(np.int64), 'round': lambda x: np.round(one(x)).astype(np.int64),
                     'floor': lambda x: np.floor(one(x)).astype(np.int64), 'ceil': lambda x: np.ceil(one(x)).astype(np.int64),
                     'cos': np.cos, 'sin': np.sin, 'tan': np.tan, 'acos': np.arccos, 'asin': np.arcsin, 'atan': np.arctan,
                     'cosh': np.cosh, 'sinh': np.sinh, 'tanh': np.tanh, 'exp': np.exp, 'ln': np.log, 'sqrt': np.sqrt}
            if op not in table:
                raise NotImplementedError(op)
            return table[op](a if op == 'not' else one(a))
        table = {'add': np.add, 'sub': np.subtract, 'mul': np.multiply, 'div': np.divide, 'min': np.minimum,
                 'max': np.maximum, 'pow': np.power, 'floordiv': np.floor_divide, 'mod': np.mod, 'fmod': np.mod,
                 'hypot': np.hypot, 'log': lambda a, b: np.log(a) / np.log(b),
                 'lt': np.less, 'le': np.less_equal, 'gt': np.greater, 'ge': np.greater_equal, 'eq': np.equal, 'ne': np.not_equal,
                 'and': np.logical_and, 'or': np.logical_or, 'xor': np.logical_xor,
                 'imply': lambda a, b: np.logical_or(np.logical_not(a), b), 'eqv': np.equal}
        if op not in table:
            raise NotImplementedError(op)
        logic = op in ('and', 'or', 'xor', 'imply', 'eqv')
        out = args[0] if logic else one(args[0])
        for b in args[1:]:
            out = table[op](out, b if logic else one(b))
        return out


class BoxBounds:
    """``bounds[name] = (lower, upper)``: float64 arrays of the fluent's shape for every state, observ and
    action fluent (-max_bound / +max_bound where nothing constrains it), ``is_box_preconditions`` /
    ``is_box_invariants`` as in the reference."""

    def __init__(self, program, max_bound=np.inf, inequality_tol=1e-5):
        self.p = program
        self.BigM, self.epsilon = float(max_bound), float(inequality_tol)
        self.bounds = {}
        for name, d in program.tensors.items():
            if d.kind in ('state', 'observ', 'action'):
                shape = program.tensor_shape(name)
                self.bounds[name] = [np.full(shape, -self.BigM), np.full(shape, +self.BigM)]
        actions = set(program.names_of_kind('action'))
        states = set(program.names_of_kind('state'))
        self.is_box_preconditions = [self._parse(e, [], actions) for e in program.preconditions]
        self.is_box_invariants = [self._parse(e, [], states) for e in program.invariants]
        for name, (lo, hi) in self.bounds.items():
            if np.any(lo > hi):                     # RDDLSimulator._check_bounds (constraints.py:60)
                raise ValueError(f'Variable <{name}>: bounds lower > upper.')
        self.bounds = {k: (v[0], v[1]) for k, v in self.bounds.items()}

    def _parse(self, n, scope, search):
        """constraints.py:107-139. ``scope``: the object types of the enclosing forall variables."""
        if n.kind == 'agg' and n.op == 'forall':
            return self._parse(n.args[0], list(n.args[0].scope), search)
        if n.kind in ('bin', 'nary') and n.op == 'and':
            ok = True
            for a in n.args:
                ok = self._parse(a, scope, search) and ok
            return ok
        if n.kind == 'bin' and n.op in _REL:
            return self._relational(n, scope, search)
        return False

    def _relational(self, n, scope, search):
        """constraints.py:141-229."""
        left, right = n.args
        is_var = lambda x: x.kind == 'load' and x.tensor in search
        lv, rv = is_var(left), is_var(right)
        if lv == rv:                      # fluents on both sides, or on neither: not a bound
            return False
        var, const = (left, right) if lv else (right, left)
        ev = _NonFluentEval(self.p, scope)
        if not ev.ok(const) or any(k == 'e' for (k, _) in var.index):
            return False
        try:
            lim = np.asarray(ev.ev(const), dtype=np.float64)
        except NotImplementedError:
            return False
        strict = n.op in ('lt', 'gt')
        upper = (n.op in ('le', 'lt')) == lv            # fluent <= rhs, or rhs >= fluent
        if strict:
            lim = lim - self.epsilon if upper else lim + self.epsilon
        # place the limit (one axis per scope variable) into the fluent's own axes
        free = [v for (k, v) in var.index if k == 'a']
        if len(set(free)) != len(free):
            return False
        lim = np.broadcast_to(lim, self.p.shape_of(scope)) if scope else lim.reshape(())
        other = [a for a in range(len(scope)) if a not in free]
        if other:                          # the bound must hold for every value of the unused variables
            lim = lim.max(axis=tuple(other)) if not upper else lim.min(axis=tuple(other))
            remaining = [a for a in range(len(scope)) if a in free]
        else:
            remaining = list(range(len(scope)))
        lim = np.transpose(lim, [remaining.index(a) for a in free]) if free else lim
        sl = tuple(slice(None) if k == 'a' else int(v) for (k, v) in var.index)
        cur = self.bounds[var.tensor][1 if upper else 0]
        if sl:
            cur[sl] = (np.minimum if upper else np.maximum)(cur[sl], lim)
        else:
            self.bounds[var.tensor][1 if upper else 0] = (np.minimum if upper else np.maximum)(cur, lim)
        return True

    def to_device(self, device):
        """name -> (lower, upper) torch tensors on ``device``."""
        import torch
        return {k: (torch.from_numpy(np.ascontiguousarray(lo)).to(device), torch.from_numpy(np.ascontiguousarray(hi)).to(device))
                for k, (lo, hi) in self.bounds.items()}
