"""The no-PLY RDDL reader (pyrddlgym_b200/rddl_reader.py) against the reference's grammar
(/root/reference/pyRDDLGym/core/parser/parser.py:283-1233): token rules, operator precedence, every
expression form, and -- where the reference is importable -- the in-tree Wildfire text
(examples/run_build.py:7-79, produced by the reference's own RDDLBuilder) parsed, traced by the
reference front end and lowered to the same op table as ``workloads.wildfire``."""
import numpy as np
import pytest

from pyrddlgym_b200 import rddl_reader, workloads
from pyrddlgym_b200.workloads import add, agg, boolean, div, func, ite, land, lnot, lor, mul, neg, num, pv, rand, sub

DOMAIN = '''
// a comment
domain zoo {
    requirements = { reward-deterministic, concurrent };
    types { res : object; mode : {@low, @mid, @high}; };
    pvariables {
        CAP(res)        : { non-fluent, real, default = 10.5 };
        LINK(res, res)  : { non-fluent, bool, default = false };
        K               : { non-fluent, int, default = -3 };
        level(res)      : { state-fluent, real, default = 1.0 };
        m               : { state-fluent, mode, default = @low };
        flow(res)       : { interm-fluent, real, level = 2 };
        seen            : { observ-fluent, bool };
        open-valve(res) : { action-fluent, bool, default = false };
        push            : { action-fluent, real, default = 0.0 };
    };
    cpfs {
        flow(?r) = sum_{?u : res} [LINK(?u, ?r) * level(?u)] / 2 - -K * 3 + min[level(?r), CAP(?r)];
        level'(?r) = if (open-valve(?r) ^ ~(level(?r) >= CAP(?r)) | exists_{?u : res} LINK(?r, ?u) ^ level(?u) > 1)
                        then Normal(flow(?r), 0.5) else level(?r) + push;
        m' = switch (m) { case @low : Discrete(mode, @low : 0.5, @mid : 0.25, @high : 0.25), case @mid : @high, default : m };
        seen = Bernoulli(0.3) <=> (forall_{?r : res} level'(?r) < 3.0 => level'(?r) ~= 0);
    };
    reward = -sum_{?r : res} abs[level'(?r) - CAP(?r)] + (m == @high) * 2;
    action-preconditions { push >= -1.0; push <= 1.0; };
    state-invariants { forall_{?r : res} [level(?r) >= 0]; };
    termination { (sum_{?r : res} level(?r)) > 100; };
}
non-fluents nf_zoo { domain = zoo; objects { res : {r1, r2, r3}; };
    non-fluents { CAP(r1) = 4.0; LINK(r1, r2); ~LINK(r2, r3); K = 7; }; }
instance zoo_inst { domain = zoo; non-fluents = nf_zoo;
    init-state { level(r2) = 2.5; m = @mid; };
    max-nondef-actions = pos-inf; horizon = 20; discount = 0.9; }
'''


def test_tokens_follow_the_reference_lexer():
    toks = [(k, v) for (k, v, _) in rddl_reader.tokenize("out-of-fuel'(?x2, @a-b) sum_{ 1.5 .25 12 a-b - c <=> => ~= // x\n y")]
    assert toks == [('IDENT', "out-of-fuel'"), ('OP', '('), ('VAR', '?x2'), ('OP', ','), ('ENUM_VAL', '@a-b'), ('OP', ')'),
                    ('IDENT', 'sum'), ('OP', '_'), ('OP', '{'), ('DOUBLE', 1.5), ('DOUBLE', 0.25), ('INTEGER', 12),
                    ('IDENT', 'a-b'), ('OP', '-'), ('IDENT', 'c'), ('OP', '<=>'), ('OP', '=>'), ('OP', '~='), ('IDENT', 'y')]
    with pytest.raises(rddl_reader.RDDLSyntaxError):
        rddl_reader.tokenize('a # b')


def test_precedence_and_expression_shapes():
    s = rddl_reader.parse_rddl(DOMAIN)
    cpf = {name: (params, e) for ((name, params), e) in s['cpfs']}
    # * and / bind tighter than + -, left associative; unary minus tighter than *; an aggregation's body
    # runs to the end of the expression (its rule precedence is below every operator: yacc shifts), the
    # brackets only group the first factor
    tv = ('typed_var', ('?u', 'res'))
    body = add(sub(div(mul(pv('LINK', '?u', '?r'), pv('level', '?u')), num(2)), mul(neg(pv('K')), num(3))),
               func('min', pv('level', '?r'), pv('CAP', '?r')))
    assert cpf['flow'] == (['?r'], ('sum', (tv, body)))
    # ~ binds like unary minus; ^ tighter than |; relational tighter than ^; quantifier bodies extend to the right
    cond = ('|', (land(pv('open-valve', '?r'), lnot(('>=', (pv('level', '?r'), pv('CAP', '?r'))))),
                  ('exists', (tv, land(pv('LINK', '?r', '?u'), ('>', (pv('level', '?u'), num(1))))))))
    assert cpf["level'"][1] == ite(cond, rand('Normal', pv('flow', '?r'), num(0.5)), add(pv('level', '?r'), pv('push')))
    sw = cpf["m'"][1]
    assert sw[0] == 'switch' and sw[1][0] == pv('m')
    assert sw[1][1] == ('case', ('@low', ('randomvar', ('Discrete', (('enum_type', 'mode'), ('lconst', ('@low', num(0.5))),
                                                                      ('lconst', ('@mid', num(0.25))), ('lconst', ('@high', num(0.25))))))))
    assert sw[1][2] == ('case', ('@mid', pv('@high'))) and sw[1][3] == ('default', pv('m'))
    # <=> lowest, then =>; forall absorbs the implication
    tr = ('typed_var', ('?r', 'res'))
    assert cpf['seen'][1] == ('<=>', (rand('Bernoulli', num(0.3)),
                                      ('forall', (tr, ('=>', (('<', (pv("level'", '?r'), num(3.0))),
                                                              ('~=', (pv("level'", '?r'), num(0)))))))))
    # unary minus applies to the whole aggregation, whose body runs to the end of the expression
    assert s['reward'] == neg(('sum', (tr, add(func('abs', sub(pv("level'", '?r'), pv('CAP', '?r'))),
                                               mul(('==', (pv('m'), pv('@high'))), num(2))))))
    assert s['pvariables'][0] == ('CAP', 'non-fluent', 'real', ['res'], 10.5, None)
    assert s['pvariables'][2] == ('K', 'non-fluent', 'int', None, -3, None)
    assert s['pvariables'][5] == ('flow', 'interm-fluent', 'real', ['res'], None, 2)
    assert s['types'] == [('res', 'object'), ('mode', ['@low', '@mid', '@high'])]
    assert s['init_non_fluent'] == [(('CAP', ['r1']), 4.0), (('LINK', ['r1', 'r2']), True), (('LINK', ['r2', 'r3']), False),
                                    (('K', None), 7)]
    assert s['init_state'] == [(('level', ['r2']), 2.5), (('m', None), '@mid')]
    assert (s['max_nondef_actions'], s['horizon'], s['discount']) == ('pos-inf', 20, 0.9)
    assert len(s['preconds']) == 2 and len(s['invariants']) == 1 and len(s['terminals']) == 1
    assert s['requirements'] == ['reward-deterministic', 'concurrent']


def test_other_expression_forms_and_errors():
    def e(txt):
        return rddl_reader._Parser(txt).expr()
    assert e('Discrete_{?x : t} p(?x)') == ('randomvar', ('Discrete(p)', (('typed_var', ('?x', 't')), (pv('p', '?x'),))))
    assert e('f(g(?x), argmax_{?y : t} q(?y), @c)') == pv('f', pv('g', '?x'), ('argmax', (('typed_var', ('?y', 't')), pv('q', '?y'))), '@c')
    assert e('$ext[?x, _](a(?x), b)') == ('pyfunc', ('ext', (['?x', '_'], [pv('a', '?x'), pv('b')])))
    assert e('MultivariateNormal[?i](mu(?i), cov(?i, _))')[0] == 'randomvector'
    assert e('inverse[row = ?a, col = ?b][M(?a, ?b)]') == ('inverse', (['?a', '?b'], pv('M', '?a', '?b')))
    assert e('det_{?a : t, ?b : t} M(?a, ?b)')[0] == 'det'
    assert e('a & b ^ c') == land(('&', (pv('a'), pv('b'))), pv('c'))
    assert e('+x - (y)') == sub(('+', (pv('x'),)), pv('y'))
    assert e('true | false') == lor(boolean(True), boolean(False))
    for bad in ('a +', 'if (a) then b', 'Normal(1)', 'sum_{?x} a', 'f(', 'a b'):
        with pytest.raises(rddl_reader.RDDLSyntaxError):
            p = rddl_reader._Parser(bad)
            p.expr()
            if p.peek()[0] != 'EOF':
                p.fail('end of expression')
    with pytest.raises(rddl_reader.RDDLSyntaxError):
        rddl_reader.parse_rddl('domain d { pvariables { }; cpfs { }; }')          # no reward, no instance


# ------------------------------------------------------------------------------ with the reference

def _wildfire_text(n, max_nondef=3, horizon=60):
    """examples/run_build.py:5-95 with numx = numy = n, through the reference's own RDDLBuilder."""
    from oracle.ref_import import load_reference
    load_reference()
    from pyRDDLGym.core.builder import RDDLBuilder
    b = RDDLBuilder()
    b.add_object_type('x_pos')
    b.add_object_type('y_pos')
    for name, v in (('COST_CUTOUT', -5.0), ('COST_PUTOUT', -10.0), ('PENALTY_TARGET_BURN', -100.0), ('PENALTY_NONTARGET_BURN', -5.0)):
        b.add_pvariable(name, [], 'non-fluent', 'real', v)
    b.add_pvariable('NEIGHBOR', ['x_pos', 'y_pos', 'x_pos', 'y_pos'], 'non-fluent', 'bool', 'false')
    b.add_pvariable('TARGET', ['x_pos', 'y_pos'], 'non-fluent', 'bool', 'false')
    b.add_pvariable('burning', ['x_pos', 'y_pos'], 'state-fluent', 'bool', 'false')
    b.add_pvariable('out-of-fuel', ['x_pos', 'y_pos'], 'state-fluent', 'bool', 'false')
    b.add_pvariable('put-out', ['x_pos', 'y_pos'], 'action-fluent', 'bool', 'false')
    b.add_pvariable('cut-out', ['x_pos', 'y_pos'], 'action-fluent', 'bool', 'false')
    b.add_cpf('burning', ['?x', '?y'],
              '''if ( put-out(?x, ?y) )
				then false
            else if (~out-of-fuel(?x, ?y) ^ ~burning(?x, ?y))
              then [if (TARGET(?x, ?y) ^ ~(exists_{?x2: x_pos, ?y2: y_pos} (NEIGHBOR(?x, ?y, ?x2, ?y2) ^ burning(?x2, ?y2))))
                    then false
                    else Bernoulli( 1.0 / (1.0 + exp[4.5 - (sum_{?x2: x_pos, ?y2: y_pos} (NEIGHBOR(?x, ?y, ?x2, ?y2) ^ burning(?x2, ?y2)))]) ) ]
			else
				burning(?x, ?y)''')
    b.add_cpf('out-of-fuel', ['?x', '?y'], 'out-of-fuel(?x, ?y) | burning(?x,?y) | (~TARGET(?x, ?y) ^ cut-out(?x, ?y))')
    b.add_reward('''[sum_{?x: x_pos, ?y: y_pos} [ COST_CUTOUT*cut-out(?x, ?y) ]]
        + [sum_{?x: x_pos, ?y: y_pos} [ COST_PUTOUT*put-out(?x, ?y) ]]
        + [sum_{?x: x_pos, ?y: y_pos} [ PENALTY_TARGET_BURN*[ (burning(?x, ?y) | out-of-fuel(?x, ?y)) ^ TARGET(?x, ?y) ]]]
        + [sum_{?x: x_pos, ?y: y_pos} [ PENALTY_NONTARGET_BURN*[ burning(?x, ?y) ^ ~TARGET(?x, ?y) ]]]''')
    b.add_object_values('x_pos', [f'x{i + 1}' for i in range(n)])
    b.add_object_values('y_pos', [f'y{i + 1}' for i in range(n)])
    spec = workloads.wildfire(n)
    for ((name, objs), v) in spec['init_non_fluent']:
        b.add_nonfluent_init(name, objs, 'true')
    for ((name, objs), v) in spec['init_state']:
        b.add_init_state(name, objs, 'true')
    b.add_max_nondef_actions(max_nondef)
    b.add_horizon(horizon)
    b.add_discount(1.0)
    return b.build_domain('wildfire') + '\n\n' + b.build_instance('wildfire', 'wildfire_large', 'nf_wildfire_large')


def _reference_present():
    from oracle.ref_import import reference_available
    return reference_available()


@pytest.mark.skipif(not _reference_present(), reason='reference pyRDDLGym not present')
def test_wildfire_text_parses_to_the_program_of_the_workload_spec():
    """The Wildfire domain + instance text of the reference tree -> this reader -> the reference's AST
    classes -> RDDLLiftedModel -> tracer -> HIR -> op table == the op table of workloads.wildfire(n)
    (the neutral spec the bench and parity workloads are built from), byte for byte; and the two
    reference simulators built from the parsed text and from the spec step identically."""
    from oracle import ast_builders
    from oracle.ref_import import load_reference
    from pyrddlgym_b200 import frontend, lowering
    ref = load_reference()
    n = 30                                             # the size of examples/run_build.py
    spec = rddl_reader.parse_rddl(_wildfire_text(n))
    assert spec['name'] == 'wildfire' and spec['max_nondef_actions'] == 3 and spec['horizon'] == 60
    sims = []
    for model in (ref.RDDLLiftedModel(rddl_reader.reference_ast(spec)),
                  ast_builders.build_model(workloads.wildfire(n, max_nondef=3, horizon=60))):
        sims.append(ref.RDDLSimulator(model, rng=np.random.default_rng(7), keep_tensors=True))
    pa, pb = (frontend.program_from_simulator(s, name='wildfire') for s in sims)
    import json
    assert json.loads(pa.to_json()) == json.loads(pb.to_json())    # (RDDLBuilder emits its object types in set order)
    for k in pa.init:
        assert np.array_equal(pa.init[k], pb.init[k]), k
    assert lowering.lower(pa).blob == lowering.lower(pb).blob
    rng = np.random.default_rng(0)
    for s in sims:
        s.reset()
    for t in range(5):
        act = {'put-out': rng.random((n, n)) < 0.05, 'cut-out': rng.random((n, n)) < 0.05}
        (oa, ra, da), (ob, rb, db) = sims[0].step(act), sims[1].step(act)
        assert ra == rb and da == db and all(np.array_equal(oa[k], ob[k]) for k in oa)


@pytest.mark.skipif(not _reference_present(), reason='reference pyRDDLGym not present')
def test_parsed_zoo_domain_builds_and_steps_in_the_reference():
    """Enum types, switch / Discrete, interm levels, observ-fluents, preconditions, invariants and
    terminations survive the trip through the reference front end."""
    from oracle.ref_import import load_reference
    ref = load_reference()
    spec = rddl_reader.parse_rddl(DOMAIN)
    model = ref.RDDLLiftedModel(rddl_reader.reference_ast(spec))
    sim = ref.RDDLSimulator(model, rng=np.random.default_rng(1), keep_tensors=True)
    obs, done = sim.reset()
    assert not done and sim.is_pomdp
    for t in range(4):
        act = sim.prepare_actions_for_sim({'push': 0.5, 'open-valve___r1': True})
        assert sim.check_action_preconditions(act, silent=True)
        obs, r, done = sim.step(act)
        assert np.isfinite(r)
    assert sim.check_state_invariants(silent=True) in (True, False)
