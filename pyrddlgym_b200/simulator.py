"""Batched RDDL simulator on B200: the host-side mirror of ``RDDLSimulator``'s
reset / step / check_* interface (/root/reference/pyRDDLGym/core/simulator.py:362-502)
with a leading env axis on every fluent.

State lives in torch CUDA tensors (bool / int64 / float64, C order -- the dtypes of
compiler/initializer.py:16-23); each ``step`` issues the launches of the lowered op
table through the C ABI (include/rddl_b200.h) on torch's current stream and then swaps
the state / next-state buffers (the reference aliases them, simulator.py:463-466).
There is no CPU execution path.
"""
import collections.abc
import ctypes

import numpy as np
import torch

from pyrddlgym_b200 import lowering, native, optable as ot

_TORCH_DT = {'b': torch.bool, 'i': torch.int64, 'f': torch.float64}


class StatusError(RuntimeError):
    """Raised by ``raise_for_status`` when a kernel flagged an invalid distribution parameter
    (the reference raises RDDLValueOutOfRangeError from inside _sample_*, simulator.py:229-253)."""

    def __init__(self, msg, status):
        super().__init__(msg)
        self.status = status


class InvalidActionCount(ValueError):
    """More non-default actions than max-nondef-actions allows (the reference raises
    RDDLInvalidActionError, simulator.py:356-360; the drop-in subclass converts)."""

    def __init__(self, msg, counts):
        super().__init__(msg)
        self.counts = counts


class _States(collections.abc.Mapping):
    """Lazy name -> tensor view of the current state (bit-packed fluents unpack on access)."""

    def __init__(self, sim):
        self._sim = sim

    def __getitem__(self, name):
        if name not in self._sim.state_names:
            raise KeyError(name)
        return self._sim.fluent(name)

    def __iter__(self):
        return iter(self._sim.state_names)

    def __len__(self):
        return len(self._sim.state_names)

    def copy(self):
        """Shallow dict copy, like ``RDDLSimulator.states`` (simulator.py:176-179): name -> tensor."""
        return {name: self[name] for name in self}


class _TensorTable(dict):
    """name -> tensor; remembers whether the C pointer array has to be rebuilt."""

    def __init__(self):
        super().__init__()
        self.dirty = True

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        self.dirty = True


class BatchedSimulator:
    """N independent environments of one compiled RDDL problem, stepped in lock-step."""

    def __init__(self, program, n_envs=1, device='cuda:0', seed=0, env_offset=0, optable=None,
                 use_planes=True, use_contract=True, use_packed=True, specialize=None, contract_precision='f64',
                 packed_actions=False):
        """``specialize``: run-time specialised kernels (NVRTC, csrc/rddl_jit.h) for this program.
        None = the library default (on unless RDDL_B200_JIT=0: plane kernels on first use, scalar
        kernels once the program is hot); True = everything on first use; False = interpreter kernels.
        ``contract_precision``: 'f64' (default) = sum-product contractions on the exact FP64 tensor-core
        kernel; 'bf16x3' = the tcgen05 kernel on three-term bfloat16 splits with float32 accumulation in
        tensor memory (csrc/rddl_contract_tc.cuh; within the 1e-5 relative tolerance for reals).
        ``packed_actions``: bool action fluents the bit-plane kernels read are taken BIT-PACKED (int32
        [n_envs, cells / 32], cell i -> bit i % 32 of word i / 32; ``pack_bool`` converts): 1/8 of the bytes of
        the reference boundary's bool arrays. ``set_actions`` / ``step`` / ``rollout`` then accept either the packed
        words (used in place) or bool arrays (packed on the way in); ``fluent`` still returns bools."""
        self.program = program
        self.n_envs = int(n_envs)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise native.NativeError('the B200 backend runs on CUDA devices only (no CPU fallback)')
        self.seed_value = int(seed)
        self.env_offset = int(env_offset)
        self.table = optable if optable is not None else lowering.lower(
            program, use_planes=use_planes, use_contract=use_contract, use_packed=use_packed,
            packed_actions=packed_actions)
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.native = native.NativeProgram(self.table.blob, index)
        if specialize is not None:
            self.native.option(native.OPTION_SPECIALIZE, 1 if specialize else 0)
            if specialize:
                self.native.option(native.OPTION_SCALAR_HOT, 0)
        if contract_precision not in ('f64', 'bf16x3'):
            raise ValueError("contract_precision must be 'f64' or 'bf16x3'")
        if contract_precision == 'bf16x3':
            self.native.option(native.OPTION_CONTRACT_TC, 1)
        self._jit_warned = False
        self._jit_checks_left = 100
        self.names = list(self.table.tensor_names)
        self.packed = set(self.table.packed)
        self.tid = self.table.tensor_id
        self.state_names = list(program.next_state)
        self.action_names = program.names_of_kind('action')
        self.t = _TensorTable()
        with torch.cuda.device(self.device):
            for name in self.names:
                dt, kind, nelem = self.table.tensor_meta[name]
                if name in program.tensors:
                    shape = tuple(program.tensor_shape(name))
                else:
                    shape = (nelem,) if nelem > 1 else ()
                if kind == ot.KIND_SHARED:
                    v = torch.from_numpy(np.ascontiguousarray(program.init[name]))
                    self.t[name] = v.to(self.device, dtype=_TORCH_DT[dt]).contiguous()
                elif name in self.packed:
                    # bool state kept bit-packed in HBM (32 cells per int32 word); see fluent()
                    self.t[name] = torch.zeros((self.n_envs, nelem // 32), dtype=torch.int32, device=self.device)
                else:
                    self.t[name] = torch.zeros((self.n_envs,) + shape, dtype=_TORCH_DT[dt], device=self.device)
            self.status = torch.zeros(self.n_envs, dtype=torch.int32, device=self.device)
        self._init = {name: torch.from_numpy(np.ascontiguousarray(program.init[name])).to(
            self.device, dtype=_TORCH_DT[program.tensors[name].dtype])
            for name in program.tensors if program.tensors[name].kind != 'nonfluent'}
        self._owned_actions = {name: self.t[name] for name in self.action_names}
        self._owned_is_default = set()        # action fluents whose owned storage holds the defaults
        self._init_packed = {}
        self._ptrs = (ctypes.c_void_p * len(self.names))()
        self._inj = (ctypes.c_void_p * max(1, len(self.table.variates)))()
        self._external = {}
        self._state_pairs = list(program.next_state.items())
        # (state, next-state, their op-table ids): what a transition swaps, resolved once
        self._swap_plan = [(s_, ns_, self.tid[s_], self.tid[ns_]) for s_, ns_ in self._state_pairs]
        self._inj_used = False
        # per-transition call: everything that does not change between steps is resolved here (the C entry
        # point, the status pointer, the raw-stream getter), so that step() is a handful of attribute reads
        self._step_fn = self.native.lib.rddl_step
        self._handle = self.native.handle
        self._status_ptr = ctypes.c_void_p(self.status.data_ptr())
        self._dev_index = index
        self._raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)
        self._states_view = _States(self)
        self.step_count = 0          # Philox step index: advances with every transition, across resets
        self.episode_step = 0        # transitions since the last reset
        self.reset()

    # ------------------------------------------------------------------ plumbing
    def _refresh_ptrs(self):
        """The device-pointer array of the C ABI (one entry per op-table tensor id); rebuilt only when a
        tensor of the table was replaced since the last call."""
        if not self.t.dirty:
            return
        for i, name in enumerate(self.names):
            self._ptrs[i] = self.t[name].data_ptr()
        self.t.dirty = False

    def _swap_state(self):
        """state <-> next-state after a transition (the reference aliases them, simulator.py:463-466):
        the table entries and -- when it is current -- the pointer array entries, without a rebuild."""
        t, ptrs, put = self.t, self._ptrs, dict.__setitem__
        clean = not t.dirty
        for s, ns, i, j in self._swap_plan:
            a, b = t[s], t[ns]
            put(t, s, b)
            put(t, ns, a)
            if clean:
                ptrs[i], ptrs[j] = ptrs[j], ptrs[i]

    def _stream(self):
        if self._raw_stream is not None:
            return self._raw_stream(self._dev_index)
        return torch.cuda.current_stream(self.device).cuda_stream

    def seed(self, seed):
        """simulator.py:129-134. Also restarts the Philox step index (see ``reset``)."""
        self.seed_value = int(seed)
        self.step_count = 0

    @property
    def launches_issued(self):
        return self.native.info(4)

    @property
    def specialised_kernels(self):
        """Number of run-time specialised plane kernels this program runs (0 = interpreter kernels only)."""
        return self.native.info(native.INFO_JIT_BUILT)

    def _jit_check(self):
        # a failed specialisation is not an error (the interpreter kernel ran, same results) but it is
        # several times slower: say so once (kernels are specialised within the first few dozen launches
        # of a program: later calls need not look any more)
        self._jit_checks_left -= 1
        if self._jit_checks_left <= 0:
            self._jit_warned = True
            return
        if not self._jit_warned and self.native.info(native.INFO_JIT_FAILED) > 0:
            self._jit_warned = True
            import warnings
            warnings.warn('pyrddlgym_b200: plane kernel specialisation failed, running the interpreter kernel: '
                          + self.native.jit_error()[:500], RuntimeWarning)

    # ------------------------------------------------------------------ reference API
    def reset(self):
        """simulator.py:405-440 -> (states, done[N])."""
        for name, own in self._owned_actions.items():
            self.t[name] = own          # never write into caller-owned action tensors
        for name, v in self._init.items():
            if name in self._owned_is_default:
                continue                # untouched since the last reset (zero-copy actions): nothing to rewrite
            if name in self.packed:
                if name not in self._init_packed:
                    self._init_packed[name] = self._pack(v.reshape(1, -1))
                self.t[name].copy_(self._init_packed[name].expand_as(self.t[name]))
            else:
                self.t[name].copy_(v.expand_as(self.t[name]))
        self._owned_is_default = set(self.action_names)
        self.status.zero_()
        # step_count -- the Philox step index -- is NOT rewound: like the reference's generator, whose
        # stream keeps advancing across episodes (simulator.py:129-134 seeds it once), every episode
        # after a reset draws fresh noise. Only seed() restarts it.
        self.episode_step = 0
        done = self.check_terminal_states()
        return self.states, done

    def reset_envs(self, mask):
        """Re-initialises the environments selected by ``mask`` (bool[N]) in place -- the rows of every
        fluent tensor go back to their initial values, status words are cleared -- while the others
        keep running (vector-env auto-reset). The Philox step index stays global: draws remain
        unique per (env, step)."""
        mask = mask.to(self.device, dtype=torch.bool)
        for name, own in self._owned_actions.items():
            if self.t[name] is not own:
                own.copy_(self.t[name])       # keep the live values of the other envs, in owned storage
                self.t[name] = own
                self._owned_is_default.discard(name)
        for name, v in self._init.items():
            t = self.t[name]
            if name in self.packed:
                if name not in self._init_packed:
                    self._init_packed[name] = self._pack(v.reshape(1, -1))
                init = self._init_packed[name].expand_as(t)
            else:
                init = v.expand_as(t)
            m = mask.reshape((self.n_envs,) + (1,) * (t.dim() - 1))
            t.copy_(torch.where(m, init, t))
        self.status.masked_fill_(mask, 0)

    @staticmethod
    def _pack(x):
        """bool [n, cells] -> int32 [n, cells / 32], cell i -> bit (i % 32) of word i / 32."""
        n = x.shape[0]
        x = x.reshape(n, -1, 32)
        # (row blocks: the int64 intermediate is 8 bytes per cell)
        rows = max(1, (1 << 27) // max(1, x.shape[1] * 32))
        sh = torch.arange(32, device=x.device, dtype=torch.int64)
        out = torch.empty((n, x.shape[1]), dtype=torch.int32, device=x.device)
        for r0 in range(0, n, rows):
            out[r0:r0 + rows] = (x[r0:r0 + rows].to(torch.int64) << sh).sum(-1).to(torch.int32)
        return out

    @staticmethod
    def pack_bool(x):
        """bool [..., *shape] with prod(shape) % 32 == 0 over the LAST axes flattened per leading row ->
        int32 [leading, cells / 32]: the layout of bit-packed fluents (leading = x.shape[0])."""
        return BatchedSimulator._pack(x.reshape(x.shape[0], -1))

    @staticmethod
    def _unpack(w, shape):
        bits = (w.unsqueeze(-1) >> torch.arange(32, device=w.device, dtype=torch.int32)) & 1
        return bits.to(torch.bool).reshape((w.shape[0],) + tuple(shape))

    @property
    def states(self):
        """name -> tensor [N, *shape] of the current state. Byte / int64 / float64 fluents are
        views of the live buffers (valid until the next step, like the reference's aliasing dict,
        simulator.py:176-179); bit-packed bool fluents are unpacked into a fresh tensor."""
        return self._states_view

    def fluent(self, name):
        if name in self.packed:
            return self._unpack(self.t[name], self.program.tensor_shape(name))
        return self.t[name]

    def set_actions(self, actions):
        """simulator.py:259-327 for tensor inputs: each entry must be [N, *shape] (or
        broadcastable); host arrays are copied to the device (pinned staging is the caller's
        business), device tensors of the right dtype/layout are used in place."""
        for name, v in actions.items():
            if name not in self.tid or name not in self.action_names:
                raise KeyError(f'<{name}> is not an action-fluent')
            dst = self._owned_actions[name]
            if name in self.packed and not (isinstance(v, torch.Tensor) and v.dtype is torch.int32):
                # bool values for a bit-packed action fluent: packed on the way in
                b = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
                shape = (self.n_envs,) + tuple(self.program.tensor_shape(name))
                v = self.pack_bool(b.to(self.device).to(torch.bool).expand(shape))
            if isinstance(v, torch.Tensor):
                if (v.dtype is dst.dtype and v.shape == dst.shape and v.is_contiguous()
                        and v.device == dst.device):
                    # zero-copy: read the caller's buffer directly (table entry + its C pointer, in place)
                    dict.__setitem__(self.t, name, v)
                    if not self.t.dirty:
                        self._ptrs[self.tid[name]] = v.data_ptr()
                    continue
                dst.copy_(v.to(dst.dtype).expand_as(dst) if v.shape != dst.shape else v, non_blocking=True)
            else:
                a = torch.from_numpy(np.array(np.broadcast_to(np.asarray(v), tuple(dst.shape)), order='C'))
                dst.copy_(a, non_blocking=True)
            self._owned_is_default.discard(name)
            self.t[name] = dst

    def step(self, actions=None, variates=None):
        """simulator.py:442-502 for all envs -> (states, reward f64[N], done bool[N]).
        ``variates``: AST node id -> pre-drawn base variates [N, *scope] (parity testing);
        nodes not listed draw Philox."""
        if actions:
            self.set_actions(actions)
        keep = []
        for slot, (node, _, nelem) in (enumerate(self.table.variates) if (variates or self._inj_used) else ()):
            v = None if variates is None else variates.get(node, None)
            if v is None:
                self._inj[slot] = None
            else:
                tv = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
                tv = tv.to(self.device, dtype=torch.float64).contiguous()
                if tv.numel() != self.n_envs * nelem:
                    raise ValueError(f'variates for node {node}: expected {self.n_envs * nelem} values')
                keep.append(tv)
                self._inj[slot] = tv.data_ptr()
        self._inj_used = bool(variates)
        if self.t.dirty:
            self._refresh_ptrs()
        rc = self._step_fn(self._handle, self._ptrs, self._inj if variates else None, self._status_ptr,
                           self.n_envs, self.env_offset, self.seed_value, self.step_count, self._stream())
        if rc != 0:
            self.native._check(rc, 'rddl_step')
        if not self._jit_warned:
            self._jit_check()
        self._swap_state()
        self.step_count += 1
        self.episode_step += 1
        self._keep = keep
        t = self.t
        return self._states_view, t['__reward'], t['__done']

    def rollout(self, actions, n_steps, gamma=None, returns=None, keep_rewards=False):
        """``n_steps`` transitions with pre-drawn action sequences -- the loop of
        BaseAgent.evaluate (/root/reference/pyRDDLGym/core/policy.py:85-117) for all envs.

        actions: name -> device tensor [n_steps, N, *shape] (one slice per step) or [N, *shape]
        (same every step). Returns (returns f64[N], rewards f64[n_steps, N] or None); the
        discounted return is accumulated while an env is not done. Identical, step for step, to
        ``n_steps`` calls of ``step`` (same Philox counters)."""
        gamma = float(self.program.discount) if gamma is None else float(gamma)
        strides = (ctypes.c_int64 * len(self.names))()
        keep = []
        for name, v in actions.items():
            if name not in self.action_names:
                raise KeyError(f'<{name}> is not an action-fluent')
            dst = self._owned_actions[name]
            v = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
            if name in self.packed and v.dtype is not torch.int32:
                # bool sequence [n_steps, N, *shape] (or [N, *shape]) for a bit-packed fluent
                b = v.to(self.device).to(torch.bool)
                lead = b.shape[:b.dim() - len(self.program.tensor_shape(name))]
                v = self.pack_bool(b.reshape((-1,) + tuple(self.program.tensor_shape(name)))).reshape(tuple(lead) + (-1,))
            v = v.to(self.device, dtype=dst.dtype).contiguous()
            if tuple(v.shape) == (n_steps,) + tuple(dst.shape):
                strides[self.tid[name]] = dst.numel() * dst.element_size()
            elif tuple(v.shape) != tuple(dst.shape):
                raise ValueError(f'<{name}>: expected shape {(n_steps,) + tuple(dst.shape)} or {tuple(dst.shape)}')
            keep.append(v)
            self.t[name] = v
        if returns is None:
            returns = torch.zeros(self.n_envs, dtype=torch.float64, device=self.device)
        rewards = torch.empty((n_steps, self.n_envs), dtype=torch.float64, device=self.device) if keep_rewards else None
        self._refresh_ptrs()
        swapped = self.native.rollout(
            self._ptrs, ctypes.c_void_p(self.status.data_ptr()), self.n_envs, self.env_offset, self.seed_value,
            self.step_count, int(n_steps), strides, gamma, ctypes.c_void_p(returns.data_ptr()),
            ctypes.c_void_p(rewards.data_ptr()) if rewards is not None else None, self._stream())
        if not self._jit_warned:
            self._jit_check()
        if swapped:
            self._swap_state()
        self.step_count += int(n_steps)
        self.episode_step += int(n_steps)
        # the action fluents now hold the last step's actions: a read-only view of the caller's
        # sequence (zero-copy, like set_actions; reset / set_actions go back to owned storage)
        for name in actions:
            last = self.t[name]
            self.t[name] = last[-1] if last.dim() == self._owned_actions[name].dim() + 1 else last
        self._keep = keep
        return returns, rewards

    # ------------------------------------------------------------------ device-side policies
    def policy_descriptor(self, spec, max_nondef=None):
        """Builds the C descriptor (``native.Policy``, rddl_policy_t) of a device-side random policy.

        spec: action fluent name -> ('bernoulli', p) | ('uniform', lo, hi) | ('default',); fluents not
        listed are read from their buffers like explicit actions. ``max_nondef``: None = the instance's
        max-nondef-actions (``program.max_allowed_actions``); the kernels keep that many randomly chosen
        groundings per (env, step) when it is smaller than the number of groundings
        (RandomAgent, /root/reference/pyRDDLGym/core/policy.py:166-178)."""
        pol = native.Policy()
        if len(spec) > native.MAX_POLICY_ENTRIES:
            raise ValueError('a device-side policy drives at most %d action fluents' % native.MAX_POLICY_ENTRIES)
        for i, (name, how) in enumerate(spec.items()):
            if name not in self.action_names:
                raise KeyError(f'<{name}> is not an action-fluent')
            if name in self.packed:
                raise ValueError(f'<{name}> is taken bit-packed (packed_actions=True): device-side policies draw into '
                                 'the byte layout; build the simulator without packed_actions for them')
            dt = self.program.tensors[name].dtype
            e = pol.entries[i]
            e.tensor = self.tid[name]
            e.dflt = float(np.asarray(self.program.init[name]).reshape(-1)[0])
            if np.any(np.asarray(self.program.init[name]) != np.asarray(self.program.init[name]).reshape(-1)[0]):
                raise ValueError(f'<{name}>: per-grounding default values are not supported by device-side policies')
            kind = how[0]
            if kind == 'default':
                e.kind = native.POLICY_DEFAULT
            elif kind == 'bernoulli':
                if dt != 'b':
                    raise ValueError(f'<{name}> is not a bool fluent')
                e.kind, e.lo = native.POLICY_BERNOULLI, float(how[1])
            elif kind == 'uniform':
                if dt == 'b':
                    raise ValueError(f'<{name}> is a bool fluent: use ("bernoulli", p)')
                e.kind = native.POLICY_UNIFORM_INT if dt == 'i' else native.POLICY_UNIFORM_REAL
                e.lo, e.hi = float(how[1]), float(how[2])
            else:
                raise ValueError(f'unknown policy kind {kind!r}')
        pol.n_entries = len(spec)
        k = self.program.max_allowed_actions if max_nondef is None else max_nondef
        total = sum(int(self._owned_actions[n][0].numel()) for n in spec)
        pol.max_nondef = -1 if (k is None or k >= total) else int(k)
        return pol

    def rollout_policy(self, policy, n_steps, gamma=None, returns=None, keep_rewards=False):
        """``n_steps`` transitions under a device-side policy (``policy_descriptor``): the loop of
        BaseAgent.evaluate with a RandomAgent (/root/reference/pyRDDLGym/core/policy.py:85-117) for all
        envs, with the actions drawn inside the kernels -- no action tensor is produced, copied or read.
        Returns (returns f64[N], rewards f64[n_steps, N] or None). Identical to materialising each
        step's actions with ``sample_policy_actions`` and stepping with them."""
        gamma = float(self.program.discount) if gamma is None else float(gamma)
        for name, own in self._owned_actions.items():
            self.t[name] = own
            # kernels that read an action through a raw pointer get it materialised in the owned buffer;
            # otherwise the policy's draws live in registers only and the buffer keeps its defaults
            if self.native.info(1000 + self.tid[name]) == 1:
                self._owned_is_default.discard(name)
        if returns is None:
            returns = torch.zeros(self.n_envs, dtype=torch.float64, device=self.device)
        rewards = torch.empty((n_steps, self.n_envs), dtype=torch.float64, device=self.device) if keep_rewards else None
        self._refresh_ptrs()
        swapped = self.native.rollout_policy(
            policy, self._ptrs, ctypes.c_void_p(self.status.data_ptr()), self.n_envs, self.env_offset, self.seed_value,
            self.step_count, int(n_steps), gamma, ctypes.c_void_p(returns.data_ptr()),
            ctypes.c_void_p(rewards.data_ptr()) if rewards is not None else None, self._stream())
        if not self._jit_warned:
            self._jit_check()
        if swapped:
            self._swap_state()
        self.step_count += int(n_steps)
        self.episode_step += int(n_steps)
        return returns, rewards

    def sample_policy_actions(self, policy, step=None):
        """Materialises the actions the device-side policy takes at Philox step ``step`` (default: the
        next transition) into the simulator's action buffers -> {name: tensor [N, *shape]}
        (RandomAgent.sample_action for every env, policy.py:166-178)."""
        for name, own in self._owned_actions.items():
            self.t[name] = own
        self._owned_is_default.clear()
        self._refresh_ptrs()
        self.native.policy_sample(policy, self._ptrs, self.n_envs, self.env_offset, self.seed_value,
                                  self.step_count if step is None else int(step), self._stream())
        ids = {policy.entries[i].tensor for i in range(policy.n_entries)}
        return {name: self.t[name] for name in self.action_names if self.tid[name] in ids}

    def count_nondefault_actions(self, actions=None, enforce_for_non_bool=True):
        """int32[N]: groundings whose value differs from the fluent's default, per env -- the counting
        part of check_default_action_count (simulator.py:329-360) as one device reduction."""
        if actions:
            self.set_actions(actions)
        names = [n for n in self.action_names if enforce_for_non_bool or self.program.tensors[n].dtype == 'b']
        counts = torch.zeros(self.n_envs, dtype=torch.int32, device=self.device)
        for k in range(0, len(names), 16):
            part = names[k:k + 16]
            c = counts if k == 0 else torch.zeros_like(counts)
            self._refresh_ptrs()
            self.native.action_count(self._ptrs, [self.tid[n] for n in part],
                                     [float(np.asarray(self.program.init[n]).reshape(-1)[0]) for n in part],
                                     ctypes.c_void_p(c.data_ptr()), self.n_envs, self._stream())
            if k:
                counts += c
        return counts

    def check_default_action_count(self, actions=None, enforce_for_non_bool=True):
        """simulator.py:329-360 for all envs: raises ``InvalidActionCount`` if any env has more than
        ``max_allowed_actions`` non-default action groundings (one device reduction + one scalar read)."""
        limit = self.program.max_allowed_actions
        counts = self.count_nondefault_actions(actions, enforce_for_non_bool)
        worst = int(counts.max().item())
        if worst > limit:
            env = int(counts.argmax().item())
            raise InvalidActionCount(f'Expected at most {limit} non-default actions, got {worst} (env {env}).', counts)

    def _run_check(self, plan):
        self._refresh_ptrs()
        self.native.check(plan, self._ptrs, ctypes.c_void_p(self.status.data_ptr()), self.n_envs,
                          self.env_offset, self.seed_value, self.step_count, self._stream())

    def check_state_invariants(self):
        """simulator.py:362-373 -> bool[N]."""
        n = self.table.n_checks['invariant']
        if n == 0:
            return torch.ones(self.n_envs, dtype=torch.bool, device=self.device)
        self._run_check(ot.PLAN_INVARIANT)
        return self.t['__chk'].reshape(self.n_envs, -1)[:, :n].all(dim=1)

    def check_action_preconditions(self, actions=None):
        """simulator.py:375-389 (subs.update(actions) first) -> bool[N]."""
        if actions:
            self.set_actions(actions)
        n = self.table.n_checks['precondition']
        if n == 0:
            return torch.ones(self.n_envs, dtype=torch.bool, device=self.device)
        self._run_check(ot.PLAN_PRECONDITION)
        return self.t['__chk'].reshape(self.n_envs, -1)[:, :n].all(dim=1)

    def check_each(self, plan):
        """bool[N, #expressions] of the last invariant / precondition evaluation."""
        key = 'invariant' if plan == ot.PLAN_INVARIANT else 'precondition'
        return self.t['__chk'].reshape(self.n_envs, -1)[:, :self.table.n_checks[key]]

    def check_terminal_states(self):
        """simulator.py:391-399 -> bool[N]."""
        self._run_check(ot.PLAN_TERMINAL)
        return self.t['__done']

    def raise_for_status(self):
        """Synchronises and raises if any env flagged an invalid distribution parameter."""
        st = self.status.cpu().numpy().astype(np.uint32)
        bad = np.nonzero(st)[0]
        if len(bad):
            bits = int(np.bitwise_or.reduce(st))
            what = [n for (b, n) in [(ot.STATUS_OUT_OF_RANGE, 'distribution parameter out of range'),
                                     (ot.STATUS_BAD_BOUNDS, 'Uniform bounds lb > ub'),
                                     (ot.STATUS_BAD_PDF, 'Discrete probabilities do not sum to 1'),
                                     (ot.STATUS_BAD_INDEX, 'object index out of range')] if bits & b]
            raise StatusError('%s in %d env(s), first env %d' % ('; '.join(what), len(bad), int(bad[0])), st)

    def close(self):
        self.native.close()
