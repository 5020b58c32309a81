"""TEST INFRASTRUCTURE ONLY -- CPU (NumPy) restatement of the reference hot path.

This is the oracle the CUDA level interpreter is checked against. It re-states the
algorithm of ``RDDLSimulator.step`` and the ``_sample_*`` evaluators of
/root/reference/pyRDDLGym/core/simulator.py over the HIR (pyrddlgym_b200/hir.py), with
one extra leading axis: N independent environments (the reference has no batch axis;
"batched" == N independent reference simulators fed the same inputs).

Parity status: the reference ships no tests or golden vectors for this path (SURVEY.md
section 4), so the oracle is pinned against *outputs of the reference itself* generated in
the build container by ``oracle/gen_fixtures.py`` / ``oracle/gen_fuzz_fixtures.py`` and committed
under ``tests/golden/`` (tests/test_oracle_golden.py replays them; tests/test_fuzz_reference.py
compares live against the imported reference when it is present).

Only tests/, __graft_entry__.smoke() and bench.py's cpu-baseline leg may import this
module; the product package never does.

Evaluation is always-evaluate (no data-dependent short-circuit): values are identical to
the reference for valid programs; parameter-range errors are reported per env as status
bits instead of exceptions.
"""
import math

import numpy as np

from oracle import philox_np

NP = {'b': np.bool_, 'i': np.int64, 'f': np.float64}

STATUS_OUT_OF_RANGE = 1      # RDDLValueOutOfRangeError: _check_positive/_check_range (simulator.py:229-253)
STATUS_BAD_BOUNDS = 2        # _check_bounds (simulator.py:240-246)
STATUS_BAD_PDF = 4           # Discrete pdf negative / not summing to 1 (simulator.py:1240-1252)


def lngamma(x):
    """simulator.py:1425-1441, element-wise: the reference recurses on the *array minimum*
    (shift by 2 until min >= 7); lngamma(x+2) - ln(x) - ln(x+1) is an identity, so applying
    the shift per element until that element is >= 7 gives the same function value up to
    rounding. Requires x > 0."""
    x = np.asarray(x, dtype=np.float64)
    corr = np.zeros_like(x)
    y = x.copy()
    for _ in range(4):
        small = y < 7
        corr = corr + np.where(small, np.log(np.where(small, y, 1.0)) + np.log(np.where(small, y + 1, 1.0)), 0.0)
        y = np.where(small, y + 2, y)
    y2 = y * y
    res = (y - 0.5) * np.log(y) - y + 0.5 * np.log(2 * np.pi) + \
        1 / (12 * y) * (
            1 + 1 / (30 * y2) * (
                -1 + 1 / (7 * y2 / 2) * (
                    1 + 1 / (4 * y2 / 3) * (
                        -1 + 1 / (99 * y2 / 140) * (
                            1 + 1 / (910 * y2 / 3))))))
    return res - corr


def _one(x):
    """``1 * x`` promotion of the reference (bool -> int64), simulator.py:603-604."""
    x = np.asarray(x)
    return x.astype(np.int64) if x.dtype == np.bool_ else x


class OracleSimulator:
    """Batched CPU oracle. State lives in ``self.subs[name]``: arrays [N, *shape] for
    fluents, [*shape] for non-fluents."""

    def __init__(self, program, n_envs=1, seed=0, env_offset=0, bitsliced_nodes=()):
        self.p = program
        # Bernoulli nodes the lowering mapped to the bit-sliced sampler (OpTable.plane_nodes):
        # their *native* draw is defined by philox_np.bernoulli_bitsliced
        self.bitsliced_nodes = set(int(x) for x in bitsliced_nodes)
        self.N = int(n_envs)
        self.seed = int(seed)
        self.env_offset = int(env_offset)
        self.step_count = 0
        self._nf_cache = {}
        self._guard = None        # per-env mask of the scalar if-branches being evaluated (status reporting)
        self.status = np.zeros(self.N, dtype=np.uint32)
        self.reset()

    # ------------------------------------------------------------------ state
    def reset(self):
        """simulator.py:405-440: subs = init_values.copy(); done = check_terminal_states()."""
        p = self.p
        self.subs = {}
        for name, decl in p.tensors.items():
            v = np.asarray(p.init[name], dtype=NP[decl.dtype])
            if decl.kind == 'nonfluent':
                self.subs[name] = v
            else:
                self.subs[name] = np.broadcast_to(v, (self.N,) + v.shape).copy()
        # (the Philox step index is not rewound: episodes after a reset draw fresh noise, like the
        # reference's generator stream, simulator.py:129-134; set_seed restarts it)
        self.status[:] = 0
        return self.states(), self.check_terminal_states()

    def set_seed(self, seed):
        self.seed = int(seed)
        self.step_count = 0

    def states(self):
        return {s: self.subs[s] for s in self.p.next_state}

    # ------------------------------------------------------------------ step
    def step(self, actions, variates=None):
        """simulator.py:442-502. ``actions``: name -> [N, *shape]. ``variates``: node id ->
        base variates [N, *scope]; missing ones come from Philox."""
        p = self.p
        self._variates = variates or {}
        for name, v in actions.items():
            decl = p.tensors[name]
            self.subs[name] = np.broadcast_to(np.asarray(v, dtype=NP[decl.dtype]),
                                              (self.N,) + p.tensor_shape(name)).copy()
        for root in p.cpfs:
            val = self.ev(root.expr)
            decl = p.tensors[root.name]
            full = (self.N,) + p.shape_of(root.scope)
            self.subs[root.name] = np.ascontiguousarray(
                np.broadcast_to(val, full)).astype(NP[decl.dtype])
        reward = np.broadcast_to(np.asarray(self.ev(p.reward), dtype=np.float64).reshape(-1), (self.N,)).copy()
        for s, ns in p.next_state.items():
            self.subs[s] = self.subs[ns]
        # terminations on the new state (simulator.py:501); a random variable inside a termination
        # draws like any other node of this step (injected variate of its node id, else Philox)
        done = self.check_terminal_states(keep_variates=True)
        self.step_count += 1
        return self.states(), reward, done

    def _all_of(self, exprs):
        ok = np.ones(self.N, dtype=bool)
        each = []
        for e in exprs:
            v = np.broadcast_to(np.asarray(self.ev(e)).reshape(-1), (self.N,))
            each.append(v.copy())
            ok &= v
        return ok, each

    def check_state_invariants(self):
        """simulator.py:362-373 -> bool[N] (all invariants hold)."""
        self._variates = {}
        return self._all_of(self.p.invariants)[0]

    def check_action_preconditions(self, actions):
        """simulator.py:375-389: subs.update(actions) then all preconditions."""
        for name, v in actions.items():
            decl = self.p.tensors[name]
            self.subs[name] = np.broadcast_to(np.asarray(v, dtype=NP[decl.dtype]),
                                              (self.N,) + self.p.tensor_shape(name)).copy()
        self._variates = {}
        return self._all_of(self.p.preconditions)[0]

    def check_terminal_states(self, keep_variates=False):
        """simulator.py:391-399 -> bool[N] (any termination true)."""
        if not keep_variates:
            self._variates = {}
        done = np.zeros(self.N, dtype=bool)
        for e in self.p.terminations:
            done |= np.broadcast_to(np.asarray(self.ev(e)).reshape(-1), (self.N,))
        return done

    # ------------------------------------------------------------------ evaluator
    def _shape(self, scope):
        return self.p.shape_of(scope)

    def _flag(self, bad, scope_rank, bit):
        """bad: bool array broadcastable to [N, *scope(+extra)] -> per-env status bit."""
        bad = np.asarray(bad)
        if bad.ndim == 0:
            bad = bad.reshape(1)
        if bad.shape[0] == 1:
            bad = np.broadcast_to(bad, (self.N,) + bad.shape[1:])
        per_env = bad.reshape(self.N, -1).any(axis=1)
        if self._guard is not None:
            per_env = per_env & self._guard      # inside a branch the reference does not evaluate for this env
        self.status[per_env] |= np.uint32(bit)

    def ev(self, n):
        k = n.kind
        rank = len(n.scope)
        if k == 'const':                                  # simulator.py:543-544
            return np.asarray(n.value, dtype=NP[n.dtype]).reshape((1,) * (rank + 1))
        if k == 'axis':                                   # tracer.py:357-380
            ext = self._shape(n.scope)[n.value]
            shp = [1] * (rank + 1)
            shp[1 + n.value] = ext
            return np.arange(ext, dtype=np.int64).reshape(shp)
        if k == 'load':
            return self._load(n)
        if k == 'un':
            return self._unary(n)
        if k == 'bin':
            return self._binary(n)
        if k == 'nary':                                   # simulator.py:613-618,735-760
            vals = [self.ev(a) for a in n.args]
            if n.op in ('add', 'mul'):
                out = _one(vals[0])
                for v in vals[1:]:
                    out = (out + _one(v)) if n.op == 'add' else (out * _one(v))
                return out
            out = vals[0]
            for v in vals[1:]:
                out = np.logical_and(out, v) if n.op == 'and' else np.logical_or(out, v)
            return out
        if k == 'agg':
            return self._agg(n)
        if k == 'if':                                     # simulator.py:903-923
            if len(n.args[0].scope) == 0:
                # a predicate that is one value per env: the reference short-circuits (all elements of the
                # predicate tensor are equal, :912-919) and never evaluates the other branch, so parameter
                # errors there are not errors. (Tensor predicates: both branches are evaluated whenever
                # the predicate is mixed; their errors are reported unconditionally here.)
                c = self.ev(n.args[0])
                cg = np.broadcast_to(np.asarray(c, dtype=bool).reshape(-1), (self.N,))
                g = self._guard
                self._guard = cg if g is None else (g & cg)
                a = self.ev(n.args[1])
                self._guard = ~cg if g is None else (g & ~cg)
                b = self.ev(n.args[2])
                self._guard = g
                return np.where(c, a, b)
            c, a, b = [self.ev(x) for x in n.args]
            return np.where(c, a, b)
        if k == 'switch':                                 # simulator.py:925-952
            pred = self.ev(n.args[0])
            if len(n.args[0].scope) == 0:
                # one predicate value per env: the reference short-circuits only when the predicate equals
                # bool(predicate) (:933-943), i.e. for the first two cases, and then evaluates just that case;
                # otherwise every case and the default
                pg = np.broadcast_to(np.asarray(pred).reshape(-1), (self.N,))
                g = self._guard
                general = (pg != 0) & (pg != 1)

                def guarded(mask, node):
                    mask = mask | general
                    self._guard = mask if g is None else (g & mask)
                    out = self.ev(node)
                    self._guard = g
                    return out

                dmask = np.zeros(self.N, dtype=bool)
                for i in (0, 1):
                    if i >= len(n.cases) or n.cases[i] is None:
                        dmask = dmask | (pg == i)
                df = None if n.default is None else guarded(dmask, n.default)
                cases = [(df if c is None else guarded(pg == i, c)) for i, c in enumerate(n.cases)]
            else:
                df = None if n.default is None else self.ev(n.default)
                cases = [(df if c is None else self.ev(c)) for c in n.cases]
            full = (self.N,) + self._shape(n.scope)
            stack = np.stack([np.broadcast_to(c, full) for c in cases], axis=0)
            idx = np.broadcast_to(pred, full)[np.newaxis, ...]
            return np.take_along_axis(stack, idx, axis=0)[0]
        if k == 'delta':                                  # simulator.py:1015-1033
            return self.ev(n.args[0])
        if k == 'rand':
            return self._random(n)
        raise NotImplementedError(k)

    def _load(self, n):
        """simulator.py:546-578 in index form: out[i...] = val[idx_0(i), ..., idx_{a-1}(i)]."""
        rank = len(n.scope)
        decl = self.p.tensors[n.tensor]
        V = self.subs[n.tensor]
        shared = decl.kind == 'nonfluent'
        const_idx = all(kind != 'e' for (kind, _) in n.index)
        if shared and const_idx and n.id in self._nf_cache:
            return self._nf_cache[n.id]
        ext = self._shape(n.scope)
        idx = []
        for (kind, v) in n.index:
            if kind == 'a':
                shp = [1] * (rank + 1)
                shp[1 + v] = ext[v]
                idx.append(np.arange(ext[v], dtype=np.int64).reshape(shp))
            elif kind == 'c':
                idx.append(np.asarray(v, dtype=np.int64).reshape((1,) * (rank + 1)))
            else:
                idx.append(np.asarray(self.ev(v), dtype=np.int64))
        if shared:
            if not idx:
                out = np.asarray(V).reshape((1,) * (rank + 1))
            else:
                out = V[tuple(idx)]
            if const_idx:
                self._nf_cache[n.id] = out
            return out
        env = np.arange(self.N, dtype=np.int64).reshape((self.N,) + (1,) * rank)
        if not idx:
            return V.reshape((self.N,) + (1,) * rank)
        return V[(env,) + tuple(idx)]

    def _unary(self, n):
        x = self.ev(n.args[0])
        op = n.op
        if op == 'not':                                   # simulator.py:686-690
            return np.logical_not(x)
        if op == 'neg':                                   # simulator.py:593-595
            return -1 * x
        x = _one(x)                                       # simulator.py:794
        with np.errstate(all='ignore'):
            if op == 'abs':
                return np.abs(x)
            if op == 'sgn':
                return np.sign(x).astype(np.int64)
            if op == 'round':
                return np.round(x).astype(np.int64)
            if op == 'floor':
                return np.floor(x).astype(np.int64)
            if op == 'ceil':
                return np.ceil(x).astype(np.int64)
            if op == 'ln':
                return np.log(x)
            if op == 'lngamma':
                return lngamma(x)
            if op == 'gamma':
                return np.exp(lngamma(x))
            fn = {'cos': np.cos, 'sin': np.sin, 'tan': np.tan, 'acos': np.arccos, 'asin': np.arcsin,
                  'atan': np.arctan, 'cosh': np.cosh, 'sinh': np.sinh, 'tanh': np.tanh,
                  'exp': np.exp, 'sqrt': np.sqrt}[op]
            return fn(x)

    def _binary(self, n):
        a = self.ev(n.args[0])
        b = self.ev(n.args[1])
        op = n.op
        if op in ('and', 'or', 'xor', 'imply', 'eqv'):    # simulator.py:74-81
            if op == 'and':
                return np.logical_and(a, b)
            if op == 'or':
                return np.logical_or(a, b)
            if op == 'xor':
                return np.logical_xor(a, b)
            if op == 'imply':
                return np.logical_or(np.logical_not(a), b)
            return np.equal(a, b)
        a, b = _one(a), _one(b)
        with np.errstate(all='ignore'):
            if op == 'add':                               # simulator.py:60-65
                return np.add(a, b)
            if op == 'sub':
                return np.subtract(a, b)
            if op == 'mul':
                return np.multiply(a, b)
            if op == 'div':
                return np.divide(a, b)
            if op in ('lt', 'le', 'gt', 'ge', 'eq', 'ne'):  # simulator.py:66-73
                return {'lt': np.less, 'le': np.less_equal, 'gt': np.greater,
                        'ge': np.greater_equal, 'eq': np.equal, 'ne': np.not_equal}[op](a, b)
            if op == 'floordiv':                          # simulator.py:116
                return np.floor_divide(a, b).astype(np.int64)
            if op == 'mod':                               # simulator.py:117
                return np.mod(a, b).astype(np.int64)
            if op == 'fmod':                              # simulator.py:118 (np.mod, not C fmod)
                return np.mod(a, b)
            if op == 'min':
                return np.minimum(a, b)
            if op == 'max':
                return np.maximum(a, b)
            if op == 'pow':
                return np.power(a, b)
            if op == 'log':                               # simulator.py:122
                return np.log(a) / np.log(b)
            if op == 'hypot':
                return np.hypot(a, b)
        raise NotImplementedError(op)

    def _agg(self, n):
        """simulator.py:766-779: reduce over the trailing axes appended for the iteration
        variables (tracer.py:786-790)."""
        body = n.args[0]
        v = self.ev(body)
        full = (self.N,) + self._shape(body.scope)
        v = np.broadcast_to(v, full)
        nred = len(n.red)
        axes = tuple(range(len(full) - nred, len(full)))
        op = n.op
        if op in ('forall', 'exists'):
            out = np.all(v, axis=axes) if op == 'forall' else np.any(v, axis=axes)
        else:
            v = _one(v)
            if op == 'sum':
                out = np.sum(v, axis=axes)
            elif op == 'avg':
                out = np.mean(v, axis=axes)
            elif op == 'prod':
                out = np.prod(v, axis=axes)
            elif op == 'min':
                out = np.min(v, axis=axes)
            elif op == 'max':
                out = np.max(v, axis=axes)
            elif op == 'argmin':
                out = np.argmin(v, axis=axes[0])
            elif op == 'argmax':
                out = np.argmax(v, axis=axes[0])
            else:
                raise NotImplementedError(op)
        return out

    # ------------------------------------------------------------------ random
    def _base(self, n, kind, scope=None):
        scope = n.scope if scope is None else scope
        shp = self._shape(scope)
        v = self._variates.get(n.id, None)
        if v is not None:
            return np.asarray(v, dtype=np.float64).reshape((self.N,) + shp)
        env_ids = self.env_offset + np.arange(self.N)
        return philox_np.base_variates(kind, self.seed, env_ids, self.step_count, n.id, shp)

    def _native(self, n, op, args, rank):
        """simulator.py:1066-1221: the distributions NumPy draws by rejection / search. No base
        variate can be shared with the device samplers (csrc/rddl_samplers.cuh), so this draws
        from NumPy's own generators on a stream keyed by (seed, step, node): same distribution,
        different numbers -- compared through moments / KS tests only."""
        full = (self.N,) + self._shape(n.scope)
        rng = np.random.default_rng([self.seed & 0xFFFFFFFF, self.step_count, int(n.id), self.env_offset])
        a = np.broadcast_to(args[0], full)
        b = np.broadcast_to(args[1], full) if len(args) > 1 else None
        pos = lambda x, strict: ~((x > 0) if strict else (x >= 0))
        if op == 'Poisson':
            self._flag(pos(a, False), rank, STATUS_OUT_OF_RANGE)
            return rng.poisson(np.maximum(a, 0.0))
        if op == 'Gamma':
            self._flag(pos(a, True) | pos(b, True), rank, STATUS_OUT_OF_RANGE)
            return rng.gamma(np.where(a > 0, a, 1.0), np.where(b > 0, b, 1.0))
        if op == 'Binomial':
            self._flag(pos(a, False) | ~((b >= 0) & (b <= 1)), rank, STATUS_OUT_OF_RANGE)
            return rng.binomial(np.maximum(a, 0), np.clip(b, 0.0, 1.0))
        if op == 'NegativeBinomial':
            self._flag(pos(a, True) | ~((b >= 0) & (b <= 1)), rank, STATUS_OUT_OF_RANGE)
            return rng.negative_binomial(np.where(a > 0, a, 1.0), np.clip(b, 1e-12, 1.0))
        if op == 'Beta':
            self._flag(pos(a, True) | pos(b, True), rank, STATUS_OUT_OF_RANGE)
            return rng.beta(np.where(a > 0, a, 1.0), np.where(b > 0, b, 1.0))
        if op == 'Geometric':
            self._flag(~((a >= 0) & (a <= 1)), rank, STATUS_OUT_OF_RANGE)
            return rng.geometric(np.clip(a, 1e-12, 1.0))
        if op == 'Student':
            self._flag(pos(a, True), rank, STATUS_OUT_OF_RANGE)
            return rng.standard_t(np.where(a > 0, a, 1.0))
        self._flag(pos(a, True), rank, STATUS_OUT_OF_RANGE)
        return rng.chisquare(np.where(a > 0, a, 1.0))

    def _random(self, n):
        op = n.op
        rank = len(n.scope)
        if op in ('Discrete', 'UnnormDiscrete', 'Discrete(p)', 'UnnormDiscrete(p)'):
            if op.endswith('(p)'):
                body = n.args[0]
                pdf = np.broadcast_to(self.ev(body), (self.N,) + self._shape(body.scope))
            else:
                full = (self.N,) + self._shape(n.scope)
                pdf = np.stack([np.broadcast_to(np.asarray(self.ev(c), dtype=np.float64), full)
                                for c in n.cases], axis=-1)
            pdf = np.asarray(pdf, dtype=np.float64)
            self._flag(~(pdf >= 0), rank + 1, STATUS_OUT_OF_RANGE)
            cdf = np.cumsum(pdf, axis=-1)                 # simulator.py:1244
            if op.startswith('Unnorm'):
                with np.errstate(all='ignore'):
                    cdf = cdf / cdf[..., -1:]
            last = cdf[..., -1]
            self._flag(~(np.abs(last - 1.0) <= 1e-8 + 1e-5), rank, STATUS_BAD_PDF)   # np.allclose
            U = self._base(n, 'U')[..., np.newaxis]
            return np.argmax(U < cdf, axis=-1)            # simulator.py:1255-1256
        args = [np.asarray(self.ev(a)) for a in n.args]
        if op in ('Poisson', 'Gamma', 'Binomial', 'NegativeBinomial', 'Beta', 'Geometric', 'Student', 'ChiSquare'):
            return self._native(n, op, args, rank)
        with np.errstate(all='ignore'):
            if op == 'Uniform':                           # simulator.py:1035-1043
                lb, ub = args
                self._flag(~(lb <= ub), rank, STATUS_BAD_BOUNDS)
                return lb + (ub - lb) * self._base(n, 'U')
            if op == 'Bernoulli':                         # simulator.py:1045-1053 (note <=)
                p, = args
                self._flag(~((p >= 0) & (p <= 1)), rank, STATUS_OUT_OF_RANGE)
                if n.id in self.bitsliced_nodes and self._variates.get(n.id, None) is None:
                    shp = self._shape(n.scope)
                    pf = np.broadcast_to(np.asarray(p, dtype=np.float64), (self.N,) + shp).reshape(self.N, -1)
                    env_ids = self.env_offset + np.arange(self.N)
                    draw = philox_np.bernoulli_bitsliced(self.seed, env_ids, self.step_count, n.id, pf)
                    return draw.reshape((self.N,) + shp)
                return self._base(n, 'U') <= p
            if op == 'Normal':                            # simulator.py:1055-1064 (variance)
                m, var = args
                self._flag(~(var >= 0), rank, STATUS_OUT_OF_RANGE)
                return m + np.sqrt(var) * self._base(n, 'N')
            if op == 'Exponential':                       # simulator.py:1075-1082 (scale)
                s, = args
                self._flag(~(s > 0), rank, STATUS_OUT_OF_RANGE)
                return s * self._base(n, 'E')
            if op == 'Weibull':                           # simulator.py:1084-1093
                shape, scale = args
                self._flag(~(shape > 0) | ~(scale > 0), rank, STATUS_OUT_OF_RANGE)
                return scale * np.power(self._base(n, 'E'), 1.0 / shape)
            if op == 'Pareto':                            # simulator.py:1149-1158
                shape, scale = args
                self._flag(~(shape > 0) | ~(scale > 0), rank, STATUS_OUT_OF_RANGE)
                return scale * np.expm1(self._base(n, 'E') / shape)
            if op == 'Gumbel':                            # simulator.py:1169-1177
                m, s = args
                self._flag(~(s > 0), rank, STATUS_OUT_OF_RANGE)
                return m - s * np.log(-np.log(1.0 - self._base(n, 'U')))
            if op == 'Laplace':                           # simulator.py:1179-1187
                m, s = args
                self._flag(~(s > 0), rank, STATUS_OUT_OF_RANGE)
                U = self._base(n, 'U')
                return np.where(U >= 0.5, m - s * np.log(2.0 - U - U), m + s * np.log(U + U))
            if op == 'Cauchy':                            # simulator.py:1189-1199
                m, s = args
                self._flag(~(s > 0), rank, STATUS_OUT_OF_RANGE)
                return m + s * self._base(n, 'C')
            if op == 'Gompertz':                          # simulator.py:1201-1212
                shape, scale = args
                self._flag(~(shape > 0) | ~(scale > 0), rank, STATUS_OUT_OF_RANGE)
                U = self._base(n, 'U')
                return np.log(1.0 - np.log1p(-U) / shape) / scale
            if op == 'Kumaraswamy':                       # simulator.py:1223-1234
                a, b = args
                self._flag(~(a > 0) | ~(b > 0), rank, STATUS_OUT_OF_RANGE)
                U = self._base(n, 'U')
                return (1.0 - U ** (1.0 / b)) ** (1.0 / a)
        raise NotImplementedError(op)
