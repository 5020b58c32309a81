"""TEST INFRASTRUCTURE ONLY -- wraps a neutral domain spec (pyrddlgym_b200.workloads)
into the reference's own AST classes and builds the reference model/simulator.

The reference ships no .rddl files except Wildfire text inside examples/run_build.py
and its text parser needs ``ply`` (absent), so models are built by constructing the
AST directly -- the tuple shapes follow the grammar actions in
/root/reference/pyRDDLGym/core/parser/parser.py:550-556,705-713,745-747,778-927 and are
consumed by /root/reference/pyRDDLGym/core/compiler/model.py:820-875,877-929,993-1240.
"""
import numpy as np

from oracle.ref_import import load_reference

_EXPR_HEADS = {
    'number', 'boolean', 'pvar_expr', 'randomvar', 'randomvector', 'func', 'pyfunc',
    '+', '-', '*', '/', '^', '&', '|', '~', '=>', '<=>', '>=', '<=', '<', '>', '==', '~=',
    'sum', 'prod', 'avg', 'max', 'min', 'forall', 'exists', 'argmin', 'argmax',
    'if', 'switch', 'det', 'inverse', 'pinverse', 'cholesky',
}


def to_expression(node, ref=None):
    """Recursively wraps tuple nodes whose head is an expression head into reference
    ``Expression`` objects; structural tuples (typed_var, case, default, lconst,
    enum_type) stay plain tuples. Always creates fresh objects (tracer.py:124-137)."""
    ref = ref or load_reference()
    E = ref.Expression
    if isinstance(node, tuple):
        if len(node) == 2 and isinstance(node[0], str) and node[0] in _EXPR_HEADS:
            head, payload = node
            if head in ('number', 'boolean'):
                return E((head, payload))
            if head in ('func', 'randomvar', 'pvar_expr', 'pyfunc', 'randomvector'):
                # payload = (name, args): the name is never an expression head
                name, args = payload
                return E((head, (name, to_expression(args, ref))))
            return E((head, to_expression(payload, ref)))
        return tuple(to_expression(x, ref) for x in node)
    if isinstance(node, list):
        return [to_expression(x, ref) for x in node]
    return node


def build_ast(spec, ref=None):
    ref = ref or load_reference()
    pvars = [ref.PVariable(name, ftype, rng, params, default)
             for (name, ftype, rng, params, default) in spec['pvariables']]
    cpfs = [ref.CPF(('pvar_expr', (name, params)), to_expression(expr, ref))
            for ((name, params), expr) in spec['cpfs']]
    sections = {
        'types': list(spec['types']),
        'pvariables': pvars,
        'cpfs': ('cpfs', cpfs),
        'reward': to_expression(spec['reward'], ref),
        'preconds': [to_expression(e, ref) for e in spec.get('preconds', [])],
        'invariants': [to_expression(e, ref) for e in spec.get('invariants', [])],
        'terminals': [to_expression(e, ref) for e in spec.get('terminals', [])],
    }
    domain = ref.Domain(spec['name'], [], sections)
    nf = ref.NonFluents('nf_' + spec['name'], {
        'domain': spec['name'],
        'objects': list(spec['objects']),
        'init_non_fluent': list(spec['init_non_fluent']),
    })
    inst = ref.Instance('inst_' + spec['name'], {
        'domain': spec['name'],
        'non_fluents': 'nf_' + spec['name'],
        'init_state': list(spec['init_state']),
        'max_nondef_actions': spec['max_nondef_actions'],
        'horizon': spec['horizon'],
        'discount': spec['discount'],
    })
    return ref.RDDL({'domain': domain, 'non_fluents': nf, 'instance': inst})


def build_model(spec, ref=None):
    ref = ref or load_reference()
    return ref.RDDLLiftedModel(build_ast(spec, ref))


_REF_SIM_CLS = None


def ref_simulator_class():
    """Reference RDDLSimulator whose *base variates* can be injected per AST node.

    The reference consumes NumPy's PCG64 stream in a state-dependent order
    (short-circuits, simulator.py:624-637,713-733,903-923), so stream-level parity with a
    counter-based GPU generator is impossible by construction. Instead, each sampler is
    re-expressed through its base variate using identities that hold bit-exactly in NumPy
    (uniform(lo,hi) = lo + (hi-lo)*random(), normal(m,s) = m + s*standard_normal(), ...;
    SURVEY.md 8c) and the base variate comes from ``self.variates[expr.id]`` -- an array of
    the node's full scope shape. Everything else is the untouched reference code.
    """
    global _REF_SIM_CLS
    if _REF_SIM_CLS is not None:
        return _REF_SIM_CLS
    ref = load_reference()
    Base = ref.RDDLSimulator

    class InjectedVariateSimulator(Base):

        def __init__(self, *args, **kwargs):
            self.variates = {}
            super().__init__(*args, **kwargs)

        def _variate(self, expr):
            v = self.variates.get(expr.id, None)
            if v is None:
                raise KeyError(f'no injected variate for node {expr.id}')
            return v

        # simulator.py:1035-1043
        def _sample_uniform(self, expr, subs):
            lb, ub = expr.args
            sample_lb = self._sample(lb, subs)
            sample_ub = self._sample(ub, subs)
            Base._check_bounds(sample_lb, sample_ub, 'Uniform', expr)
            return sample_lb + (sample_ub - sample_lb) * self._variate(expr)

        # simulator.py:1045-1053
        def _sample_bernoulli(self, expr, subs):
            pr, = expr.args
            sample_pr = self._sample(pr, subs)
            Base._check_range(sample_pr, 0, 1, 'Bernoulli p', expr)
            return self._variate(expr) <= sample_pr

        # simulator.py:1055-1064
        def _sample_normal(self, expr, subs):
            mean, var = expr.args
            sample_mean = self._sample(mean, subs)
            sample_var = self._sample(var, subs)
            Base._check_positive(sample_var, False, 'Normal variance', expr)
            return sample_mean + np.sqrt(sample_var) * self._variate(expr)

        # simulator.py:1075-1082
        def _sample_exponential(self, expr, subs):
            scale, = expr.args
            sample_scale = self._sample(scale, subs)
            Base._check_positive(sample_scale, True, 'Exponential rate', expr)
            return sample_scale * self._variate(expr)

        # simulator.py:1084-1093 ; rng.weibull(a) == standard_exponential() ** (1/a)
        def _sample_weibull(self, expr, subs):
            shape, scale = expr.args
            sample_shape = self._sample(shape, subs)
            sample_scale = self._sample(scale, subs)
            Base._check_positive(sample_shape, True, 'Weibull shape', expr)
            Base._check_positive(sample_scale, True, 'Weibull scale', expr)
            return sample_scale * np.power(self._variate(expr), 1.0 / sample_shape)

        # simulator.py:1149-1158 ; rng.pareto(a) == expm1(standard_exponential() / a)
        def _sample_pareto(self, expr, subs):
            shape, scale = expr.args
            sample_shape = self._sample(shape, subs)
            sample_scale = self._sample(scale, subs)
            Base._check_positive(sample_shape, True, 'Pareto shape', expr)
            Base._check_positive(sample_scale, True, 'Pareto scale', expr)
            return sample_scale * np.expm1(self._variate(expr) / sample_shape)

        # simulator.py:1169-1177 ; rng.gumbel(loc, scale) == loc - scale*log(-log(1-U))
        def _sample_gumbel(self, expr, subs):
            mean, scale = expr.args
            sample_mean = self._sample(mean, subs)
            sample_scale = self._sample(scale, subs)
            Base._check_positive(sample_scale, True, 'Gumbel scale', expr)
            U = self._variate(expr)
            return sample_mean - sample_scale * np.log(-np.log(1.0 - U))

        # simulator.py:1179-1187 ; numpy legacy-free laplace: U>=0.5 ? loc - scale*log(2-2U)
        #                                                           : loc + scale*log(2U)
        def _sample_laplace(self, expr, subs):
            mean, scale = expr.args
            sample_mean = self._sample(mean, subs)
            sample_scale = self._sample(scale, subs)
            Base._check_positive(sample_scale, True, 'Laplace scale', expr)
            U = self._variate(expr)
            with np.errstate(divide='ignore', invalid='ignore'):
                hi = sample_mean - sample_scale * np.log(2.0 - U - U)
                lo = sample_mean + sample_scale * np.log(U + U)
            return np.where(U >= 0.5, hi, lo)

        # simulator.py:1189-1199 ; base variate is the standard Cauchy draw itself
        def _sample_cauchy(self, expr, subs):
            mean, scale = expr.args
            sample_mean = self._sample(mean, subs)
            sample_scale = self._sample(scale, subs)
            Base._check_positive(sample_scale, True, 'Cauchy scale', expr)
            return sample_mean + sample_scale * self._variate(expr)

        # simulator.py:1201-1212
        def _sample_gompertz(self, expr, subs):
            shape, scale = expr.args
            sample_shape = self._sample(shape, subs)
            sample_scale = self._sample(scale, subs)
            Base._check_positive(sample_shape, True, 'Gompertz shape', expr)
            Base._check_positive(sample_scale, True, 'Gompertz scale', expr)
            U = self._variate(expr)
            return np.log(1.0 - np.log1p(-U) / sample_shape) / sample_scale

        # simulator.py:1223-1234
        def _sample_kumaraswamy(self, expr, subs):
            a, b = expr.args
            sample_a = self._sample(a, subs)
            sample_b = self._sample(b, subs)
            Base._check_positive(sample_a, True, 'Kumaraswamy a', expr)
            Base._check_positive(sample_b, True, 'Kumaraswamy b', expr)
            U = self._variate(expr)
            return (1.0 - U ** (1.0 / sample_b)) ** (1.0 / sample_a)

        # simulator.py:1240-1256
        def _sample_discrete_helper(self, pdf, unnorm, expr):
            Base._check_positive(pdf, False, 'Discrete probabilities', expr)
            cdf = np.cumsum(pdf, axis=-1)
            if unnorm:
                cdf = cdf / cdf[..., -1:]
            if not np.allclose(cdf[..., -1], 1.0):
                raise ref.exc.RDDLValueOutOfRangeError('Discrete probabilities must sum to 1')
            U = np.asarray(self._variate(expr))[..., np.newaxis]
            return np.argmax(U < cdf, axis=-1)

    _REF_SIM_CLS = InjectedVariateSimulator
    return _REF_SIM_CLS
