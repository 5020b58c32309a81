"""ctypes binding of the C ABI declared in include/rddl_b200.h (librddl_b200.so).

The library is built in-tree by ``__graft_entry__.build()`` / ``pyrddlgym_b200.build``.
There is no CPU execution path: if the library is missing or the device is not CUDA the
binding raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('RDDL_B200_LIB') or os.path.join(_HERE, 'csrc', 'librddl_b200.so')

_lib = None

POLICY_BERNOULLI, POLICY_UNIFORM_REAL, POLICY_UNIFORM_INT, POLICY_DEFAULT = 1, 2, 3, 4
MAX_POLICY_ENTRIES = 8
MAX_POLICY_SELECT = 32


class PolicyEntry(ctypes.Structure):
    """rddl_policy_entry_t (include/rddl_b200.h)."""
    _fields_ = [('tensor', ctypes.c_int32), ('kind', ctypes.c_int32), ('lo', ctypes.c_double),
                ('hi', ctypes.c_double), ('dflt', ctypes.c_double)]


class Policy(ctypes.Structure):
    """rddl_policy_t (include/rddl_b200.h)."""
    _fields_ = [('n_entries', ctypes.c_int32), ('max_nondef', ctypes.c_int32),
                ('entries', PolicyEntry * MAX_POLICY_ENTRIES)]


class NativeError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            '(the B200 backend has no CPU fallback)')
    lib = ctypes.CDLL(LIB_PATH)
    vp, i64, u64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64
    lib.rddl_program_create.argtypes = [vp, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(vp)]
    lib.rddl_program_create.restype = ctypes.c_int
    lib.rddl_program_destroy.argtypes = [vp]
    lib.rddl_program_destroy.restype = None
    lib.rddl_last_error.argtypes = [vp]
    lib.rddl_last_error.restype = ctypes.c_char_p
    lib.rddl_program_info.argtypes = [vp, ctypes.c_int]
    lib.rddl_program_info.restype = i64
    lib.rddl_step.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), vp, i64, i64, u64, u64, vp]
    lib.rddl_step.restype = ctypes.c_int
    lib.rddl_check.argtypes = [vp, ctypes.c_int, ctypes.POINTER(vp), vp, i64, i64, u64, u64, vp]
    lib.rddl_check.restype = ctypes.c_int
    lib.rddl_rollout.argtypes = [vp, ctypes.POINTER(vp), vp, i64, i64, u64, u64, ctypes.c_int32,
                                 ctypes.POINTER(i64), ctypes.c_double, vp, vp, ctypes.POINTER(ctypes.c_int32), vp]
    lib.rddl_rollout.restype = ctypes.c_int
    lib.rddl_rollout_policy.argtypes = [vp, ctypes.POINTER(Policy), ctypes.POINTER(vp), vp, i64, i64, u64, u64,
                                        ctypes.c_int32, ctypes.c_double, vp, vp, ctypes.POINTER(ctypes.c_int32), vp]
    lib.rddl_rollout_policy.restype = ctypes.c_int
    lib.rddl_policy_sample.argtypes = [vp, ctypes.POINTER(Policy), ctypes.POINTER(vp), i64, i64, u64, u64, vp]
    lib.rddl_policy_sample.restype = ctypes.c_int
    lib.rddl_action_count.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_int32),
                                      ctypes.POINTER(ctypes.c_double), ctypes.c_int32, vp, i64, vp]
    lib.rddl_action_count.restype = ctypes.c_int
    lib.rddl_policy_precompile.argtypes = [ctypes.c_char_p, ctypes.c_size_t, i64, ctypes.POINTER(Policy)]
    lib.rddl_policy_precompile.restype = ctypes.c_int
    lib.rddl_profile.argtypes = [vp, ctypes.c_int]
    lib.rddl_profile.restype = ctypes.c_int
    lib.rddl_profile_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64), ctypes.c_int]
    lib.rddl_profile_read.restype = ctypes.c_int
    lib.rddl_return_stats.argtypes = [vp, i64, vp, vp]
    lib.rddl_return_stats.restype = ctypes.c_int
    lib.rddl_program_option.argtypes = [vp, ctypes.c_int, i64]
    lib.rddl_program_option.restype = ctypes.c_int
    lib.rddl_jit_error.argtypes = [vp]
    lib.rddl_jit_error.restype = ctypes.c_char_p
    lib.rddl_plane_precompile.argtypes = [ctypes.c_char_p, ctypes.c_size_t, i64]
    lib.rddl_plane_precompile.restype = ctypes.c_int
    lib.rddl_plane_source.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, i64, ctypes.c_char_p,
                                      ctypes.c_size_t]
    lib.rddl_plane_source.restype = i64
    _lib = lib
    return lib


EXPORTS = ['rddl_program_create', 'rddl_program_destroy', 'rddl_last_error', 'rddl_program_info',
           'rddl_step', 'rddl_check', 'rddl_rollout', 'rddl_profile', 'rddl_profile_read',
           'rddl_return_stats', 'rddl_program_option', 'rddl_jit_error', 'rddl_plane_precompile', 'rddl_plane_source',
           'rddl_rollout_policy', 'rddl_policy_sample', 'rddl_action_count', 'rddl_policy_precompile']

OPTION_SPECIALIZE = 0
OPTION_SCALAR_HOT = 1       # scalar launches a program issues before its scalar kernels are specialised
OPTION_CONTRACT_TC = 2      # 1: sum-product contractions on the tcgen05 split-bf16 kernel (rddl_contract_tc.cuh)
INFO_JIT_BUILT, INFO_JIT_FAILED, INFO_JIT_CACHE_HITS = 5, 6, 7


def return_stats(returns, out5, stream):
    """out5 (f64[5], device) <- {sum, sum of squares, count, min, max} of returns (f64[n], device)."""
    rc = load().rddl_return_stats(ctypes.c_void_p(returns.data_ptr()), int(returns.numel()),
                                  ctypes.c_void_p(out5.data_ptr()), ctypes.c_void_p(stream))
    if rc != 0:
        raise NativeError('rddl_return_stats failed (%d)' % rc)


def precompile(blob, n_envs):
    """Compiles the specialised plane kernels of an op table for batches of ``n_envs`` into the on-disk
    cubin cache (host only: NVRTC needs no device). Returns the number of kernels."""
    lib = load()
    blob = bytes(blob)
    rc = lib.rddl_plane_precompile(blob, len(blob), int(n_envs))
    if rc < 0:
        raise NativeError('rddl_plane_precompile failed (%d): %s' % (rc, lib.rddl_last_error(None).decode()))
    return rc


def precompile_policy(blob, n_envs, policy):
    """Like ``precompile`` for the plane kernels a device-side policy (``Policy``) specialises."""
    lib = load()
    blob = bytes(blob)
    rc = lib.rddl_policy_precompile(blob, len(blob), int(n_envs), ctypes.byref(policy))
    if rc < 0:
        raise NativeError('rddl_policy_precompile failed (%d): %s' % (rc, lib.rddl_last_error(None).decode()))
    return rc


def plane_source(blob, launch, n_envs):
    """The generated constants ("plane_static_data.inc") of one plane launch, for inspection."""
    lib = load()
    blob = bytes(blob)
    buf = ctypes.create_string_buffer(1 << 20)
    rc = lib.rddl_plane_source(blob, len(blob), int(launch), int(n_envs), buf, len(buf))
    if rc < 0:
        raise NativeError('rddl_plane_source failed (%d): %s' % (rc, lib.rddl_last_error(None).decode()))
    return buf.value.decode()


class NativeProgram:
    """Owns one ``rddl_program_t``."""

    def __init__(self, blob, device_index):
        self.lib = load()
        self.handle = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(blob, len(blob))
        rc = self.lib.rddl_program_create(ctypes.cast(buf, ctypes.c_void_p), len(blob), int(device_index),
                                          ctypes.byref(self.handle))
        if rc != 0:
            raise NativeError('rddl_program_create failed (%d): %s'
                              % (rc, self.lib.rddl_last_error(None).decode()))

    def _check(self, rc, what):
        if rc != 0:
            raise NativeError('%s failed (%d): %s' % (what, rc, self.lib.rddl_last_error(self.handle).decode()))

    def info(self, what):
        return int(self.lib.rddl_program_info(self.handle, what))

    def option(self, what, value):
        self._check(self.lib.rddl_program_option(self.handle, int(what), int(value)), 'rddl_program_option')

    def jit_error(self):
        return self.lib.rddl_jit_error(self.handle).decode()

    def step(self, ptrs, injected, status_ptr, n_envs, env_offset, seed, step, stream):
        self._check(self.lib.rddl_step(self.handle, ptrs, injected, status_ptr, n_envs, env_offset,
                                       seed, step, stream), 'rddl_step')

    def check(self, which, ptrs, status_ptr, n_envs, env_offset, seed, step, stream):
        self._check(self.lib.rddl_check(self.handle, which, ptrs, status_ptr, n_envs, env_offset, seed, step,
                                        stream), 'rddl_check')

    def rollout(self, ptrs, status_ptr, n_envs, env_offset, seed, step0, n_steps, strides, gamma,
                returns_ptr, rewards_ptr, stream):
        swapped = ctypes.c_int32(0)
        self._check(self.lib.rddl_rollout(self.handle, ptrs, status_ptr, n_envs, env_offset, seed, step0,
                                          n_steps, strides, gamma, returns_ptr, rewards_ptr,
                                          ctypes.byref(swapped), stream), 'rddl_rollout')
        return bool(swapped.value)

    def rollout_policy(self, policy, ptrs, status_ptr, n_envs, env_offset, seed, step0, n_steps, gamma,
                       returns_ptr, rewards_ptr, stream):
        swapped = ctypes.c_int32(0)
        self._check(self.lib.rddl_rollout_policy(self.handle, ctypes.byref(policy), ptrs, status_ptr, n_envs,
                                                 env_offset, seed, step0, n_steps, gamma, returns_ptr, rewards_ptr,
                                                 ctypes.byref(swapped), stream), 'rddl_rollout_policy')
        return bool(swapped.value)

    def policy_sample(self, policy, ptrs, n_envs, env_offset, seed, step, stream):
        self._check(self.lib.rddl_policy_sample(self.handle, ctypes.byref(policy), ptrs, n_envs, env_offset, seed,
                                                step, stream), 'rddl_policy_sample')

    def action_count(self, ptrs, tensor_ids, defaults, counts_ptr, n_envs, stream):
        n = len(tensor_ids)
        ids = (ctypes.c_int32 * max(1, n))(*tensor_ids)
        dfl = (ctypes.c_double * max(1, n))(*defaults)
        self._check(self.lib.rddl_action_count(self.handle, ptrs, ids, dfl, n, counts_ptr, n_envs, stream),
                    'rddl_action_count')

    def profile(self, enable):
        self._check(self.lib.rddl_profile(self.handle, 1 if enable else 0), 'rddl_profile')

    def profile_read(self, launches=None):
        """[{launch, ms, count, name}] for every launch index that ran since the last read."""
        n = 256
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_int64 * n)()
        self._check(self.lib.rddl_profile_read(self.handle, ms, cnt, n), 'rddl_profile_read')
        out = []
        for i in range(n):
            if cnt[i]:
                name = 'rddl_level_interp[launch %d]' % i
                if launches is not None and i < len(launches):
                    name = launches[i].get('kernel', name)
                out.append({'launch': i, 'ms': float(ms[i]), 'count': int(cnt[i]), 'name': name})
        return out

    def close(self):
        if self.handle:
            self.lib.rddl_program_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
