"""RDDL text reader without PLY.

The reference parses RDDL with a PLY LALR grammar (/root/reference/pyRDDLGym/core/parser/parser.py:
lexer :28-243, precedence :260-275, grammar :283-1233) and needs the third-party ``ply`` package at
import time. This module is a hand-written tokenizer + recursive-descent / precedence-climbing parser
for the same language that produces

* ``parse_rddl(text) -> spec``: a neutral domain spec (plain tuples / lists / dicts, exactly the
  shapes the grammar actions build: ``('pvar_expr', (name, args))``, ``('+', (a, b))``,
  ``('sum', (('typed_var', (v, t)), ..., body))``, ``('if', (c, a, b))``, ``('func', (name, [args]))``,
  ``('randomvar', (name, (args...)))`` ... -- the format of ``pyrddlgym_b200.workloads``), and
* ``reference_ast(spec)``: the *reference's own* AST classes (RDDL / Domain / Instance / NonFluents /
  PVariable / CPF / Expression) built from such a spec when pyRDDLGym is importable, ready for
  ``RDDLLiftedModel`` -- so a ``.rddl`` file can be loaded where ``ply`` is absent without touching
  the reused front end.

Operator precedence follows parser.py:260-275 as yacc applies it: ``<=>`` < ``=>`` < ``|`` < ``^ &`` <
relational < ``+ -`` < ``* /`` < unary ``- + ~`` (``NOT expr %prec UMINUS``); the bodies of
``if``-``else``, quantifiers, aggregations and ``Discrete_{...}`` extend as far right as possible
(their rule precedence is below every binary operator, so yacc always shifts).
pyfunc (``$f[...]``), random vectors and matrix operators are parsed as well; the simulator backends
decide what they can run.
"""
import re

_ALPHA = r'[A-Za-z]'
_DIGIT = r'[0-9]'
# parser.py:20-26
_IDENT = r'(' + _ALPHA + r')((' + _ALPHA + r'|' + _DIGIT + r'|\-|\_)*(' + _ALPHA + r'|' + _DIGIT + r'))?(\')?'
_VAR = r'\?(' + _ALPHA + r'|' + _DIGIT + r'|\-|\_)*(' + _ALPHA + r'|' + _DIGIT + r')'
_ENUM = r'\@(' + _ALPHA + r'|' + _DIGIT + r'|\-|\_)*(' + _ALPHA + r'|' + _DIGIT + r')'
_DOUBLE = _DIGIT + r'*\.' + _DIGIT + r'+'
_INTEGER = _DIGIT + r'+'

# PLY tries function tokens in definition order, then string tokens by decreasing pattern length
_TOKEN_RE = re.compile('|'.join([
    r'(?P<WS>[ \t\r\n]+)', r'(?P<COMMENT>//[^\r\n]*)',
    r'(?P<IDENT>' + _IDENT + r')', r'(?P<VAR>' + _VAR + r')', r'(?P<ENUM_VAL>' + _ENUM + r')',
    r'(?P<DOUBLE>' + _DOUBLE + r')', r'(?P<INTEGER>' + _INTEGER + r')',
    r'(?P<OP><=>|=>|~=|<=|>=|==|[\^|~+*(){}.,_\[\]<>=/\-:;$?&])',
]))

DISTRIBUTIONS = {
    'KronDelta': 1, 'DiracDelta': 1, 'Uniform': 2, 'Bernoulli': 1, 'Normal': 2, 'Poisson': 1, 'Exponential': 1,
    'Weibull': 2, 'Gamma': 2, 'Binomial': 2, 'NegativeBinomial': 2, 'Beta': 2, 'Geometric': 1, 'Pareto': 2,
    'Student': 1, 'Gumbel': 2, 'Laplace': 2, 'Cauchy': 2, 'Gompertz': 2, 'ChiSquare': 1, 'Kumaraswamy': 2,
}
RANDOM_VECTORS = {'MultivariateNormal': 2, 'MultivariateStudent': 3, 'Dirichlet': 1, 'Multinomial': 2}
MATRIX_OPS = ('inverse', 'pinverse', 'cholesky')
FLUENT_TYPES = ('non-fluent', 'state-fluent', 'action-fluent', 'interm-fluent', 'derived-fluent', 'observ-fluent',
                'param-fluent')
# identifiers the reference lexer reserves (parser.py:33-116): never pvariable / function names
_KEYWORDS = {'if', 'then', 'else', 'switch', 'case', 'default', 'otherwise', 'true', 'false', 'forall', 'exists',
             'argmax', 'argmin', 'Discrete', 'UnnormDiscrete', 'det', 'row', 'col'} | set(DISTRIBUTIONS) | \
    set(RANDOM_VECTORS) | set(MATRIX_OPS)


class RDDLSyntaxError(ValueError):
    """Malformed RDDL text (the reference raises RDDLParseError / a yacc error, parser.py:1235-1297)."""


def tokenize(text):
    """[(kind, value, line)] with kind in IDENT VAR ENUM_VAL DOUBLE INTEGER OP; comments and blanks dropped."""
    out, pos, line = [], 0, 1
    while pos < len(text):
        m = _TOKEN_RE.match(text, pos)
        if m is None:
            raise RDDLSyntaxError(f'line {line}: illegal character {text[pos]!r}')
        kind, val = m.lastgroup, m.group(m.lastgroup)
        if kind == 'DOUBLE':
            out.append((kind, float(val), line))
        elif kind == 'INTEGER':
            out.append((kind, int(val), line))
        elif kind not in ('WS', 'COMMENT'):
            out.append((kind, val, line))
        line += val.count('\n') if kind in ('WS', 'COMMENT') else 0
        pos = m.end()
    return out


class _Parser:

    def __init__(self, text):
        self.toks = tokenize(text)
        self.i = 0

    # ---- token helpers
    def peek(self, k=0):
        j = self.i + k
        return self.toks[j] if j < len(self.toks) else ('EOF', None, self.toks[-1][2] if self.toks else 0)

    def fail(self, what):
        kind, val, line = self.peek()
        raise RDDLSyntaxError(f'line {line}: expected {what}, found {val!r}')

    def at(self, val, k=0):
        t = self.peek(k)
        return t[0] in ('OP', 'IDENT') and t[1] == val

    def accept(self, val):
        if self.at(val):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            self.fail(repr(val))

    def take(self, kind):
        t = self.peek()
        if t[0] != kind:
            self.fail(kind)
        self.i += 1
        return t[1]

    def ident(self):
        return self.take('IDENT')

    # ---- top level (parser.py:283-312)
    def parse(self):
        blocks = {}
        while self.peek()[0] != 'EOF':
            if self.at('domain'):
                blocks['domain'] = self.domain_block()
            elif self.at('non-fluents'):
                blocks['non_fluents'] = self.nonfluents_block()
            elif self.at('instance'):
                blocks['instance'] = self.instance_block()
            elif self.at('policy'):
                raise RDDLSyntaxError('policy {...} blocks are not supported by this reader')
            else:
                self.fail("'domain', 'non-fluents' or 'instance'")
        return blocks

    # ---- domain (parser.py:318-687)
    def domain_block(self):
        self.expect('domain')
        d = {'name': self.ident(), 'requirements': [], 'types': [], 'pvariables': [], 'cpfs': [], 'reward': None,
             'preconds': [], 'invariants': [], 'terminals': [], 'constraints': []}
        self.expect('{')
        while not self.at('}'):
            if self.accept('requirements'):
                self.accept('=')
                self.expect('{')
                while not self.at('}'):
                    d['requirements'].append(self.ident())
                    self.accept(',')
                self.expect('}')
                self.expect(';')
            elif self.accept('types'):
                self.expect('{')
                while not self.at('}'):
                    name = self.ident()
                    self.expect(':')
                    if self.accept('object'):
                        d['types'].append((name, 'object'))
                    else:
                        self.expect('{')
                        vals = []
                        while not self.at('}'):
                            vals.append(self.take('ENUM_VAL'))
                            self.accept(',')
                        self.expect('}')
                        d['types'].append((name, vals))
                    self.expect(';')
                self.expect('}')
                self.expect(';')
            elif self.accept('pvariables'):
                self.expect('{')
                while not self.at('}'):
                    d['pvariables'].append(self.pvar_def())
                self.expect('}')
                self.expect(';')
            elif self.at('cpfs') or self.at('cdfs'):
                self.i += 1
                self.expect('{')
                while not self.at('}'):
                    name = self.ident()
                    params = None
                    if self.accept('('):
                        params = []
                        while not self.at(')'):
                            params.append(self.take('VAR'))
                            self.accept(',')
                        self.expect(')')
                    self.expect('=')
                    d['cpfs'].append(((name, params), self.expr()))
                    self.expect(';')
                self.expect('}')
                self.expect(';')
            elif self.accept('reward'):
                self.expect('=')
                d['reward'] = self.expr()
                self.expect(';')
            elif self.at('termination') or self.at('action-preconditions') or self.at('state-invariants') \
                    or self.at('state-action-constraints'):
                key = {'termination': 'terminals', 'action-preconditions': 'preconds',
                       'state-invariants': 'invariants', 'state-action-constraints': 'constraints'}[self.peek()[1]]
                self.i += 1
                self.expect('{')
                while not self.at('}'):
                    d[key].append(self.expr())
                    self.expect(';')
                self.expect('}')
                self.expect(';')
            else:
                self.fail('a domain section')
        self.expect('}')
        return d

    def range_const(self):
        """parser.py:512-540: true / false, [-] number, pos-inf / neg-inf, @enum, object name."""
        if self.accept('true'):
            return True
        if self.accept('false'):
            return False
        if self.accept('pos-inf'):
            return 'pos-inf'
        if self.accept('neg-inf'):
            return 'neg-inf'
        neg = self.accept('-')
        t = self.peek()
        if t[0] in ('DOUBLE', 'INTEGER'):
            self.i += 1
            return -t[1] if neg else t[1]
        if neg:
            self.fail('a number')
        if t[0] in ('ENUM_VAL', 'IDENT'):
            self.i += 1
            return t[1]
        self.fail('a constant')

    def pvar_def(self):
        """parser.py:420-510 -> (name, fluent type, range, params | None, default, level)."""
        name = self.ident()
        params = None
        if self.accept('('):
            params = []
            while not self.at(')'):
                params.append(self.ident())
                self.accept(',')
            self.expect(')')
        self.expect(':')
        self.expect('{')
        ftype = self.ident()
        if ftype not in FLUENT_TYPES:
            self.fail('a fluent type')
        self.expect(',')
        rng = self.ident()
        default, level = None, None
        if self.accept(','):
            if self.accept('default'):
                self.expect('=')
                default = self.range_const()
            elif self.accept('level'):
                self.expect('=')
                level = self.range_const()
            else:
                self.fail("'default' or 'level'")
        elif ftype in ('interm-fluent', 'derived-fluent'):
            level = 1                                   # parser.py:479-489
        self.expect('}')
        self.expect(';')
        return (name, ftype, rng, params, default, level)

    # ---- instance / non-fluents (parser.py:985-1159)
    def lconst_list(self):
        out = []
        while not self.at(')'):
            t = self.peek()
            if t[0] not in ('IDENT', 'ENUM_VAL'):
                self.fail('an object')
            self.i += 1
            out.append(t[1])
            self.accept(',')
        return out

    def pvar_inst_list(self):
        out = []
        while not self.at('}'):
            neg = self.accept('~')
            name = self.ident()
            objs = None
            if self.accept('('):
                objs = self.lconst_list()
                self.expect(')')
            value = not neg
            if not neg and self.accept('='):
                value = self.range_const()
            self.expect(';')
            out.append(((name, objs), value))
        return out

    def objects_section(self):
        out = []
        self.expect('{')
        while not self.at('}'):
            t = self.ident()
            self.expect(':')
            self.expect('{')
            objs = []
            while not self.at('}'):
                objs.append(self.ident())
                self.accept(',')
            self.expect('}')
            self.expect(';')
            out.append((t, objs))
        self.expect('}')
        self.expect(';')
        return out

    def nonfluents_block(self):
        self.expect('non-fluents')
        nf = {'name': self.ident(), 'domain': None, 'objects': [], 'init_non_fluent': []}
        self.expect('{')
        while not self.at('}'):
            if self.accept('domain'):
                self.expect('=')
                nf['domain'] = self.ident()
                self.expect(';')
            elif self.accept('objects'):
                nf['objects'] = self.objects_section()
            elif self.accept('non-fluents'):
                self.expect('{')
                nf['init_non_fluent'] = self.pvar_inst_list()
                self.expect('}')
                self.expect(';')
            else:
                self.fail('a non-fluents section')
        self.expect('}')
        return nf

    def instance_block(self):
        self.expect('instance')
        inst = {'name': self.ident(), 'domain': None, 'non_fluents': None, 'init_state': [],
                'max_nondef_actions': 'pos-inf', 'horizon': None, 'discount': 1.0}
        self.expect('{')
        while not self.at('}'):
            if self.accept('domain'):
                self.expect('=')
                inst['domain'] = self.ident()
                self.expect(';')
            elif self.at('non-fluents') and self.at('=', 1):
                self.i += 2
                inst['non_fluents'] = self.ident()
                self.expect(';')
            elif self.accept('non-fluents'):         # 2018-style instance: non-fluents { ... }; inline (parser.py:989-1003)
                self.expect('{')
                inst['init_non_fluent'] = self.pvar_inst_list()
                self.expect('}')
                self.expect(';')
            elif self.accept('objects'):
                inst['objects'] = self.objects_section()
            elif self.accept('init-state'):
                self.expect('{')
                inst['init_state'] = self.pvar_inst_list()
                self.expect('}')
                self.expect(';')
            elif self.accept('max-nondef-actions'):
                self.expect('=')
                inst['max_nondef_actions'] = 'pos-inf' if self.accept('pos-inf') else self.take('INTEGER')
                self.expect(';')
            elif self.accept('horizon'):
                self.expect('=')
                if self.accept('terminate-when'):
                    self.expect('(')
                    inst['horizon'] = self.expr()
                    self.expect(')')
                    self.accept(';')
                else:
                    inst['horizon'] = 'pos-inf' if self.accept('pos-inf') else self.take('INTEGER')
                    self.expect(';')
            elif self.accept('discount'):
                self.expect('=')
                inst['discount'] = self.take('DOUBLE')
                self.expect(';')
            else:
                self.fail('an instance section')
        self.expect('}')
        return inst

    # ---- expressions (parser.py:688-980), precedence climbing
    _BINARY = [('<=>',), ('=>',), ('|',), ('^', '&'), ('==', '~=', '<', '<=', '>', '>='), ('+', '-'), ('*', '/')]

    def expr(self, level=0):
        if level == len(self._BINARY):
            return self.unary()
        left = self.expr(level + 1)
        while self.peek()[0] == 'OP' and self.peek()[1] in self._BINARY[level]:
            op = self.take('OP')
            right = self.expr(level + 1)
            left = (op, (left, right))
        return left

    def unary(self):
        if self.peek()[0] == 'OP' and self.peek()[1] in ('-', '+', '~'):
            op = self.take('OP')
            return (op, (self.unary(),))
        return self.primary()

    def typed_vars(self):
        self.expect('_')
        self.expect('{')
        out = []
        while True:
            v = self.take('VAR')
            self.expect(':')
            out.append(('typed_var', (v, self.ident())))
            if not self.accept(','):
                break
        self.expect('}')
        return out

    def term_list(self, close):
        """Arguments of a pvariable reference (parser.py:716-738): ?x and @o stay strings, nested
        pvariable / argmax / Discrete_{...} references are expressions."""
        out = []
        while not self.at(close):
            t = self.peek()
            if t[0] in ('VAR', 'ENUM_VAL'):
                self.i += 1
                out.append(t[1])
            elif self.at('argmax') or self.at('argmin') or ((self.at('Discrete') or self.at('UnnormDiscrete')) and self.at('_', 1)):
                out.append(self.primary())
            elif t[0] == 'IDENT':
                out.append(self.pvar_ref())
            else:
                self.fail('an argument')
            if not self.accept(','):
                break
        return out

    def pvar_ref(self):
        name = self.ident()
        if self.accept('('):
            args = self.term_list(')')
            self.expect(')')
            return ('pvar_expr', (name, args))
        return ('pvar_expr', (name, None))

    def vector_terms(self, close):
        out = []
        while not self.at(close):
            t = self.peek()
            if t[0] in ('VAR', 'ENUM_VAL') or (t[0] == 'OP' and t[1] == '_'):
                self.i += 1
                out.append(t[1])
            else:
                self.fail('?var, @object or _')
            if not self.accept(','):
                break
        return out

    def vector_pvar(self):
        name = self.ident()
        if self.accept('('):
            args = self.vector_terms(')')
            self.expect(')')
            return ('pvar_expr', (name, args))
        return ('pvar_expr', (name, None))

    def primary(self):
        kind, val, _ = self.peek()
        if kind in ('DOUBLE', 'INTEGER'):
            self.i += 1
            return ('number', val)
        if kind == 'VAR' or kind == 'ENUM_VAL':
            self.i += 1
            return ('pvar_expr', (val, None))
        if kind == 'OP':
            if val in ('(', '['):
                self.i += 1
                e = self.expr()
                self.expect(')' if val == '(' else ']')
                return e
            if val == '$':                              # external Python function (parser.py:768-775)
                self.i += 1
                name = self.ident()
                self.expect('[')
                captured = self.vector_terms(']')
                self.expect(']')
                args = []
                if self.accept('('):
                    while not self.at(')'):
                        args.append(self.vector_pvar())
                        self.accept(',')
                    self.expect(')')
                return ('pyfunc', (name, (captured, args)))
            self.fail('an expression')
        if kind != 'IDENT':
            self.fail('an expression')
        if val in ('true', 'false'):
            self.i += 1
            return ('boolean', val == 'true')
        if val == 'if':
            self.i += 1
            self.expect('(')
            c = self.expr()
            self.expect(')')
            self.expect('then')
            a = self.expr()
            self.expect('else')
            return ('if', (c, a, self.expr()))
        if val == 'switch':
            self.i += 1
            self.expect('(')
            pred = self.expr()
            self.expect(')')
            self.expect('{')
            cases = []
            while True:
                if self.accept('case'):
                    t = self.peek()
                    if t[0] in ('VAR', 'ENUM_VAL'):
                        self.i += 1
                        lit = t[1]
                    else:
                        lit = self.pvar_ref()
                    self.expect(':')
                    cases.append(('case', (lit, self.expr())))
                elif self.accept('default'):
                    self.expect(':')
                    cases.append(('default', self.expr()))
                else:
                    self.fail("'case' or 'default'")
                if not self.accept(','):
                    break
            self.expect('}')
            return ('switch', (pred,) + tuple(cases))
        if val in ('forall', 'exists', 'argmax', 'argmin'):
            self.i += 1
            tv = self.typed_vars()
            return (val, tuple(tv) + (self.expr(),))
        if val in ('Discrete', 'UnnormDiscrete'):
            self.i += 1
            if self.at('_'):                            # Discrete_{?x : t} expr  (parser.py:905-908)
                tv = self.typed_vars()
                return ('randomvar', (val + '(p)', tuple(tv) + ((self.expr(),),)))
            self.expect('(')
            enum = self.ident()
            cases = []
            while self.accept(','):
                t = self.peek()
                if t[0] not in ('IDENT', 'ENUM_VAL'):
                    self.fail('an enum value')
                self.i += 1
                self.expect(':')
                cases.append(('lconst', (t[1], 'otherwise' if self.accept('otherwise') else self.expr())))
            self.expect(')')
            return ('randomvar', (val, (('enum_type', enum),) + tuple(cases)))
        if val in RANDOM_VECTORS and self.at('[', 1):
            self.i += 1
            self.expect('[')
            dims = self.vector_terms(']')
            self.expect(']')
            self.expect('(')
            args = []
            while not self.at(')'):
                args.append(self.vector_pvar())
                self.accept(',')
            self.expect(')')
            return ('randomvector', (val, (dims, tuple(args))))
        if val == 'Dirichlet' and self.at('(', 1):       # scalar form Dirichlet(type, expr) (parser.py:869)
            self.i += 2
            enum = self.ident()
            self.expect(',')
            e = self.expr()
            self.expect(')')
            return ('randomvar', (val, (enum, e)))
        if val in DISTRIBUTIONS:
            self.i += 1
            self.expect('(')
            args = [self.expr()]
            while self.accept(','):
                args.append(self.expr())
            self.expect(')')
            if len(args) != DISTRIBUTIONS[val]:
                self.fail(f'{DISTRIBUTIONS[val]} argument(s) of {val}')
            return ('randomvar', (val, tuple(args)))
        if val == 'det':
            self.i += 1
            self.expect('_')
            self.expect('{')
            tv = []
            for k in range(2):
                v = self.take('VAR')
                self.expect(':')
                tv.append(('typed_var', (v, self.ident())))
                if k == 0:
                    self.expect(',')
            self.expect('}')
            return ('det', (tv[0], tv[1], self.expr()))
        if val in MATRIX_OPS:
            self.i += 1
            self.expect('[')
            self.expect('row')
            self.expect('=')
            r = self.take('VAR')
            self.expect(',')
            self.expect('col')
            self.expect('=')
            c = self.take('VAR')
            self.expect(']')
            self.expect('[')
            e = self.expr()
            self.expect(']')
            return (val, ([r, c], e))
        if val in _KEYWORDS:
            self.fail('an expression')
        # IDENT: function f[...], aggregation f_{...} body, or pvariable reference
        if self.at('[', 1):
            self.i += 2
            args = [self.expr()]
            while self.accept(','):
                args.append(self.expr())
            self.expect(']')
            return ('func', (val, args))
        if self.at('_', 1) and self.at('{', 2):
            self.i += 1
            tv = self.typed_vars()
            return (val, tuple(tv) + (self.expr(),))
        return self.pvar_ref()


def parse_blocks(text):
    """{'domain': ..., 'non_fluents': ..., 'instance': ...} as plain dicts."""
    return _Parser(text).parse()


def parse_rddl(text):
    """RDDL text (domain + non-fluents + instance blocks, in one string) -> neutral domain spec, the
    format of ``pyrddlgym_b200.workloads`` (see module docstring)."""
    b = parse_blocks(text)
    if 'domain' not in b:
        raise RDDLSyntaxError('domain {...} block is missing')
    if 'instance' not in b:
        raise RDDLSyntaxError('instance {...} block is missing')
    d, inst = b['domain'], b['instance']
    nf = b.get('non_fluents') or {'name': 'nf__' + d['name'], 'objects': [], 'init_non_fluent': []}
    if 'init_non_fluent' in inst:                    # 2018-style instance carrying its own non-fluents (parser.py:989-1003)
        nf = {'name': 'nf__' + (inst['domain'] or d['name']), 'objects': inst.get('objects', []),
              'init_non_fluent': inst['init_non_fluent']}
    elif 'objects' in inst and not nf['objects']:
        nf = dict(nf, objects=inst['objects'])
    if d['reward'] is None:
        raise RDDLSyntaxError('reward expression is missing')
    return dict(
        name=d['name'], instance_name=inst['name'], nonfluents_name=nf['name'], requirements=d['requirements'],
        types=d['types'], pvariables=d['pvariables'], cpfs=d['cpfs'], reward=d['reward'],
        preconds=d['preconds'], invariants=d['invariants'], terminals=d['terminals'], constraints=d['constraints'],
        objects=nf['objects'], init_non_fluent=nf['init_non_fluent'], init_state=inst['init_state'],
        max_nondef_actions=inst['max_nondef_actions'], horizon=inst['horizon'], discount=inst['discount'],
    )


def read_rddl(domain_path, instance_path=None):
    """RDDLReader + RDDLParser.parse of the reference (reader.py:8-76, parser.py:1287-1297) for files."""
    with open(domain_path, encoding='utf-8', errors='replace') as f:
        text = f.read()
    if instance_path is not None:
        with open(instance_path, encoding='utf-8', errors='replace') as f:
            text += '\n\n' + f.read()
    return parse_rddl(text)


# ---------------------------------------------------------------------------------------
# neutral spec -> the reference's own AST classes
# ---------------------------------------------------------------------------------------

_EXPR_HEADS = {
    'number', 'boolean', 'pvar_expr', 'randomvar', 'randomvector', 'func', 'pyfunc',
    '+', '-', '*', '/', '^', '&', '|', '~', '=>', '<=>', '>=', '<=', '<', '>', '==', '~=',
    'sum', 'prod', 'avg', 'max', 'min', 'forall', 'exists', 'argmin', 'argmax',
    'if', 'switch', 'det', 'inverse', 'pinverse', 'cholesky',
}


def _is_expr(node):
    return (isinstance(node, tuple) and len(node) == 2 and isinstance(node[0], str)
            and (node[0] in _EXPR_HEADS or (isinstance(node[1], tuple) and node[1] and isinstance(node[1][0], tuple)
                                            and node[1][0][:1] == ('typed_var',))))


def _wrap(node, E):
    """Every expression tuple becomes a fresh reference ``Expression`` (parser.py:688-703 wraps each
    ``expr`` reduction); structural tuples (typed_var, case, default, lconst, enum_type) stay tuples."""
    if _is_expr(node):
        head, payload = node
        if head in ('number', 'boolean'):
            return E((head, payload))
        if head in ('func', 'randomvar', 'pvar_expr', 'pyfunc', 'randomvector'):
            name, args = payload
            return E((head, (name, _wrap(args, E))))
        return E((head, _wrap(payload, E)))
    if isinstance(node, tuple):
        return tuple(_wrap(x, E) for x in node)
    if isinstance(node, list):
        return [_wrap(x, E) for x in node]
    return node


def reference_ast(spec):
    """The reference's ``RDDL`` object (parser/rddl.py:32-36) for a neutral spec: what
    ``RDDLParser().parse(text)`` returns, accepted by ``RDDLLiftedModel``. Needs pyRDDLGym importable
    (its parser *module* is not imported, so ``ply`` is not needed)."""
    from pyRDDLGym.core.parser.cpf import CPF
    from pyRDDLGym.core.parser.domain import Domain
    from pyRDDLGym.core.parser.expr import Expression
    from pyRDDLGym.core.parser.instance import Instance
    from pyRDDLGym.core.parser.nonfluents import NonFluents
    from pyRDDLGym.core.parser.pvariable import PVariable
    from pyRDDLGym.core.parser.rddl import RDDL
    pvars = []
    for pv in spec['pvariables']:
        name, ftype, rng, params, default = pv[:5]
        level = pv[5] if len(pv) > 5 else None
        pvars.append(PVariable(name=name, fluent_type=ftype, range_type=rng, param_types=params, default=default,
                               level=level))
    cpfs = [CPF(('pvar_expr', (name, params)), _wrap(expr, Expression)) for ((name, params), expr) in spec['cpfs']]
    sections = {
        'types': list(spec['types']),
        'pvariables': pvars,
        'cpfs': ('cpfs', cpfs),
        'reward': _wrap(spec['reward'], Expression),
        'preconds': [_wrap(e, Expression) for e in spec.get('preconds', [])],
        'invariants': [_wrap(e, Expression) for e in spec.get('invariants', [])],
        'terminals': [_wrap(e, Expression) for e in spec.get('terminals', [])],
    }
    if spec.get('constraints'):
        sections['constraints'] = [_wrap(e, Expression) for e in spec['constraints']]
    name = spec['name']
    nf_name = spec.get('nonfluents_name') or 'nf_' + name
    domain = Domain(name, list(spec.get('requirements', [])), sections)
    nf = NonFluents(nf_name, {'domain': name, 'objects': list(spec['objects']),
                              'init_non_fluent': list(spec['init_non_fluent'])})
    inst = Instance(spec.get('instance_name') or 'inst_' + name, {
        'domain': name, 'non_fluents': nf_name, 'init_state': list(spec['init_state']),
        'max_nondef_actions': spec['max_nondef_actions'], 'horizon': spec['horizon'], 'discount': spec['discount']})
    return RDDL({'domain': domain, 'non_fluents': nf, 'instance': inst})
