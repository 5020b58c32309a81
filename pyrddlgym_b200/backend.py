"""Drop-in backend for pyRDDLGym: ``RDDLEnv(domain, instance, backend=B200RDDLSimulator,
backend_kwargs={'device': 'cuda:0'})`` (``RDDLEnv`` drives one environment, n_envs == 1; batches
of environments go through ``B200RDDLSimulator(model, n_envs=N)`` directly or through
``pyrddlgym_b200.vector_env.BatchedRDDLEnv``).

The plug-in point is the reference's own (/root/reference/pyRDDLGym/core/env.py:47-48,
107-110: ``backend(self.model, logger=..., keep_tensors=..., **backend_kwargs)``; "a backend is
a subclass of RDDLSimulator"). This class reuses the parent's front-end calls in ``_compile``
(initializer, level analysis, tracer -- simulator.py:136-174), converts their product to the
HIR, lowers it to the op table and hands stepping to the CUDA kernels through the C ABI.

* ``n_envs == 1`` (what ``RDDLEnv`` drives): same argument and return types as the reference --
  dict of NumPy arrays / grounded scalars, ``float`` reward, ``bool`` done, the reference's
  exception types for violated constraints and invalid distribution parameters.
* ``n_envs > 1``: tensors carry a leading env axis ``[n_envs, ...]``; ``step`` returns
  (dict of torch tensors, reward tensor, done tensor) and the ``check_*`` methods return bool[N].

pyRDDLGym is imported lazily: the rest of this package (HIR, lowering, kernels) does not need it.
"""
import numpy as np

from pyrddlgym_b200 import frontend

_CLS = None


def _default_engine(program, n_envs, device, seed):
    from pyrddlgym_b200.simulator import BatchedSimulator
    return BatchedSimulator(program, n_envs=n_envs, device=device, seed=seed)


def backend_class():
    """Builds (once) and returns the ``RDDLSimulator`` subclass."""
    global _CLS
    if _CLS is not None:
        return _CLS
    from pyRDDLGym.core.simulator import RDDLSimulator
    from pyRDDLGym.core.compiler.initializer import RDDLValueInitializer
    from pyRDDLGym.core.debug.exception import (
        RDDLActionPreconditionNotSatisfiedError, RDDLInvalidActionError, RDDLNotImplementedError,
        RDDLStateInvariantNotSatisfiedError, RDDLValueOutOfRangeError)

    class B200RDDLSimulator(RDDLSimulator):

        def __init__(self, rddl, allow_synchronous_state=True, rng=None, logger=None, keep_tensors=False,
                     objects_as_strings=True, python_functions=None, n_envs=1, device='cuda:0', seed=0,
                     engine_factory=None):
            self.n_envs = int(n_envs)
            self.device = device
            self._seed0 = int(seed)
            self._engine_factory = engine_factory or _default_engine
            super().__init__(rddl, allow_synchronous_state=allow_synchronous_state,
                             rng=np.random.default_rng(seed) if rng is None else rng, logger=logger,
                             keep_tensors=keep_tensors, objects_as_strings=objects_as_strings,
                             python_functions=python_functions)

        # ---- compile: reference front end, then HIR -> op table -> device
        def _compile(self):
            super()._compile()
            if self.python_functions:
                raise RDDLNotImplementedError(
                    'External Python functions cannot run on the device (no CPU fallback).')
            self.program = frontend.program_from_simulator(self)
            self.engine = self._engine_factory(self.program, self.n_envs, self.device, self._seed0)
            if self.logger is not None:
                rows = '\n\t'.join(f"plan {l['plan']}: {l['kernel']} scope={l['scope']} E={l['E']} "
                                   f"insns={l['n_insns']}" for l in self.engine.table.launches)
                self.logger.log(f'[info] B200 op table: {self.engine.table.n_insns} instructions, '
                                f'launches:\n\t{rows}\n')

        def seed(self, seed):
            super().seed(seed)
            self.engine.seed(seed)

        # ---- helpers
        @staticmethod
        def _host(x):
            return x.detach().cpu().numpy() if hasattr(x, 'detach') else np.asarray(x)

        def _raise_status(self):
            try:
                self.engine.raise_for_status()
            except Exception as e:  # StatusError -> the reference's exception type (simulator.py:229-253)
                if type(e).__name__ == 'StatusError':
                    raise RDDLValueOutOfRangeError(str(e))
                raise

        def _publish(self, names):
            """Single-env mode: mirrors device fluents into ``self.subs`` with the reference's
            host types (array per parameterised fluent, NumPy scalar otherwise)."""
            rddl = self.rddl
            for name in names:
                v = self._host(self.engine.fluent(name))[0]
                self.subs[name] = v if rddl.variable_params[name] else v[()]

        def _state_dict(self, fluents):
            """The obs / state dict the reference builds in reset and step (simulator.py:412-426,
            463-479): optional object strings, optional grounding."""
            rddl = self.rddl
            out = {}
            for var in fluents:
                values = self.subs[var]
                if self.objects_as_strings:
                    ptype = rddl.variable_ranges[var]
                    if ptype not in RDDLValueInitializer.NUMPY_TYPES:
                        values = rddl.index_to_object_string_array(ptype, values)
                if self.keep_tensors:
                    out[var] = values
                else:
                    out.update(rddl.ground_var_with_values(var, values))
            return out

        def _batch(self, actions):
            if self.n_envs > 1:
                return actions
            return {k: np.asarray(v)[np.newaxis, ...] for (k, v) in actions.items()}

        # ---- RDDLSimulator interface
        def reset(self):
            rddl = self.rddl
            self.subs = self.init_values.copy()
            _, done = self.engine.reset()
            if self.n_envs > 1:
                self.state = self.engine.states
                return self.state, done
            self.state = self._state_dict(rddl.state_fluents)
            if self._pomdp:
                if self.keep_tensors:
                    obs = {var: None for var in rddl.observ_fluents}
                else:
                    obs = {}
                    for var in rddl.observ_fluents:
                        obs.update(rddl.ground_var_with_value(var, None))
            else:
                obs = self.state
            return obs, bool(self._host(done)[0])

        def step(self, actions):
            rddl = self.rddl
            states, reward, done = self.engine.step(self._batch(actions))
            if self.n_envs > 1:
                self.state = states
                return states, reward, done
            self._raise_status()
            self.subs.update(actions)
            self._publish([n for n, d in self.program.tensors.items() if d.kind not in ('nonfluent', 'action', 'next')])
            for (state, next_state) in rddl.next_state.items():
                self.subs[next_state] = self.subs[state]
            self.state = self._state_dict(rddl.next_state.keys())
            obs = self._state_dict(rddl.observ_fluents) if self._pomdp else self.state
            return obs, float(self._host(reward)[0]), bool(self._host(done)[0])

        def check_default_action_count(self, actions, enforce_for_non_bool=True):
            """simulator.py:329-360. Batched mode: one device reduction over the action tensors
            ([n_envs, ...]) instead of np.count_nonzero per fluent; any env above max-nondef-actions
            raises the reference's RDDLInvalidActionError."""
            if self.n_envs == 1:
                return super().check_default_action_count(actions, enforce_for_non_bool)
            try:
                self.engine.check_default_action_count(actions, enforce_for_non_bool)
            except ValueError as e:
                if type(e).__name__ == 'InvalidActionCount':
                    raise RDDLInvalidActionError(str(e))
                raise

        def check_state_invariants(self, silent=False):
            ok = self.engine.check_state_invariants()
            if self.n_envs > 1:
                return ok
            if bool(self._host(ok)[0]):
                return True
            if not silent:
                each = self._host(self.engine.check_each(2))[0]
                bad = int(np.argmin(each))
                raise RDDLStateInvariantNotSatisfiedError(f'{self.invariant_names[bad]} is not satisfied.')
            return False

        def check_action_preconditions(self, actions, silent=False):
            ok = self.engine.check_action_preconditions(self._batch(actions))
            if self.n_envs > 1:
                return ok
            self.subs.update(actions)
            if bool(self._host(ok)[0]):
                return True
            if not silent:
                each = self._host(self.engine.check_each(3))[0]
                bad = int(np.argmin(each))
                raise RDDLActionPreconditionNotSatisfiedError(
                    f'{self.precond_names[bad]} is not satisfied for actions {actions}.')
            return False

        def check_terminal_states(self):
            done = self.engine.check_terminal_states()
            return done if self.n_envs > 1 else bool(self._host(done)[0])

        def sample_reward(self):
            r = self.engine.fluent('__reward')
            return r if self.n_envs > 1 else float(self._host(r)[0])

        def rollout(self, actions, n_steps, gamma=None):
            """Batched counterpart of the BaseAgent.evaluate loop (policy.py:85-117)."""
            return self.engine.rollout(actions, n_steps, gamma=gamma)

    _CLS = B200RDDLSimulator
    return _CLS


def __getattr__(name):
    if name == 'B200RDDLSimulator':
        return backend_class()
    raise AttributeError(name)
