"""TEST INFRASTRUCTURE ONLY -- imports the unmodified reference (pyRDDLGym) read-only.

Only ``tests/``, ``oracle/gen_fixtures.py`` / ``gen_fuzz_fixtures.py`` (fixture generation, run in
the build container) and ``bench.py --impl reference`` may import this module. The product
package ``pyrddlgym_b200`` never does (its drop-in backend imports ``pyRDDLGym``
directly and only when the user already has it installed).

The reference is pure Python + NumPy, but ``pyRDDLGym/__init__.py:1-2`` eagerly
imports ``core/env.py`` which needs gymnasium / pygame / matplotlib / ply, none
of which exist in this image. The numerical path (``core/simulator.py`` and
``core/compiler/*``) needs only NumPy, so the absent modules are stubbed in
``sys.modules``. The stubs are installed only for modules
that are genuinely missing.
"""
import importlib
import importlib.util
import os
import sys
import types

# Where the unmodified reference lives: $RDDL_REFERENCE_ROOT, else the read-only checkout of the build
# container, else the copy `pip install --target baseline/_ref` made from it (oracle/install_reference.py;
# git-ignored, but it travels to the GPU box with the tree -- there it is the only one present).
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    for cand in (os.environ.get('RDDL_REFERENCE_ROOT'), '/root/reference', os.path.join(_REPO, 'baseline', '_ref')):
        if cand and os.path.isdir(os.path.join(cand, 'pyRDDLGym')):
            return cand
    return os.environ.get('RDDL_REFERENCE_ROOT', '/root/reference')


REFERENCE_ROOT = _find_root()

_STUBS = ['gymnasium', 'gymnasium.spaces', 'pygame', 'matplotlib',
          'matplotlib.pyplot', 'ply', 'ply.lex', 'ply.yacc']


def reference_available() -> bool:
    if importlib.util.find_spec('pyRDDLGym') is not None:
        return True
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'pyRDDLGym'))


def _install_stubs():
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            if importlib.util.find_spec(name) is not None:
                continue
        except (ImportError, ValueError, ModuleNotFoundError):
            pass
        sys.modules[name] = types.ModuleType(name)
    g = sys.modules['gymnasium']
    if not hasattr(g, 'Env'):
        g.Env = object
    if not hasattr(g, 'spaces'):
        g.spaces = sys.modules['gymnasium.spaces']
    sp = sys.modules['gymnasium.spaces']
    for n in ['Box', 'Dict', 'Discrete']:
        if not hasattr(sp, n):
            setattr(sp, n, type(n, (), {}))
    m = sys.modules['matplotlib']
    if not hasattr(m, 'use'):
        m.use = lambda *a, **k: None
    if not hasattr(m, 'pyplot'):
        m.pyplot = sys.modules['matplotlib.pyplot']
    p = sys.modules['ply']
    if not hasattr(p, 'lex'):
        p.lex = sys.modules['ply.lex']
    if not hasattr(p, 'yacc'):
        p.yacc = sys.modules['ply.yacc']
    if not hasattr(sys.modules['ply.lex'], 'TOKEN'):
        # parser.py uses @lex.TOKEN(...) at import time
        sys.modules['ply.lex'].TOKEN = lambda r: (lambda f: f)


_loaded = None


def load_reference():
    """Returns a namespace with the reference classes used by the test tooling."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise ImportError(
            f'reference pyRDDLGym not found (looked in {REFERENCE_ROOT}); '
            'it exists only in the build container')
    _install_stubs()
    if importlib.util.find_spec('pyRDDLGym') is None:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    from pyRDDLGym.core.simulator import RDDLSimulator
    from pyRDDLGym.core.compiler.model import RDDLLiftedModel, RDDLPlanningModel
    from pyRDDLGym.core.compiler.tracer import RDDLObjectsTracer
    from pyRDDLGym.core.compiler.initializer import RDDLValueInitializer
    from pyRDDLGym.core.compiler.levels import RDDLLevelAnalysis
    from pyRDDLGym.core.parser.expr import Expression
    from pyRDDLGym.core.parser.pvariable import PVariable
    from pyRDDLGym.core.parser.cpf import CPF
    from pyRDDLGym.core.parser.domain import Domain
    from pyRDDLGym.core.parser.instance import Instance
    from pyRDDLGym.core.parser.nonfluents import NonFluents
    from pyRDDLGym.core.parser.rddl import RDDL
    from pyRDDLGym.core.debug import exception as exc
    ns.RDDLSimulator = RDDLSimulator
    ns.RDDLLiftedModel = RDDLLiftedModel
    ns.RDDLPlanningModel = RDDLPlanningModel
    ns.RDDLObjectsTracer = RDDLObjectsTracer
    ns.RDDLValueInitializer = RDDLValueInitializer
    ns.RDDLLevelAnalysis = RDDLLevelAnalysis
    ns.Expression = Expression
    ns.PVariable = PVariable
    ns.CPF = CPF
    ns.Domain = Domain
    ns.Instance = Instance
    ns.NonFluents = NonFluents
    ns.RDDL = RDDL
    ns.exc = exc
    _loaded = ns
    return ns
