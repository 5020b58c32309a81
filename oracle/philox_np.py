"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the counter-based generator used by
the CUDA samplers (pyrddlgym_b200/csrc/philox.cuh), so that the native (non-injected)
random path can be checked element-for-element.

Philox4x32-10 (Salmon et al., SC'11; the published algorithm, constants below).
The reference itself draws from NumPy's PCG64 stream in a state-dependent order
(/root/reference/pyRDDLGym/core/simulator.py:1043-1256), which no counter-based
generator can reproduce; stream parity is therefore neither possible nor required
(SURVEY.md 8c). What *is* pinned is the mapping counter -> base variate:

  key     = (seed & 0xffffffff, seed >> 32)
  counter = (elem & 0xffffffff, global_env_id, step, node_id | ((elem >> 32) << 16))
  U = ((r0 >> 5) * 2^26 + (r1 >> 6)) * 2^-53                 in [0, 1)
  N = sqrt(-2 ln(1 - U(r0,r1))) * cos(2 pi U(r2,r3))         (Box-Muller)
  E = -log1p(-U)          C = tan(pi (U - 1/2))
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments uint32 arrays (broadcastable); returns four uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over='ignore'):
        for r in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _u53(a, b):
    return ((a >> np.uint32(5)).astype(np.float64) * 67108864.0 +
            (b >> np.uint32(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)


def base_variates(kind, seed, env_ids, step, node_id, scope_shape):
    """Base variates of ``kind`` ('U','N','E','C') for every (env, scope element).
    Returns float64 array of shape [len(env_ids), *scope_shape]."""
    n_elem = int(np.prod(scope_shape, dtype=np.int64)) if len(scope_shape) else 1
    elem = np.arange(n_elem, dtype=np.uint64)
    env = np.asarray(env_ids, dtype=np.uint64)[:, None]
    c0 = (elem & MASK).astype(np.uint32)[None, :]
    c1 = (env & MASK).astype(np.uint32)
    c2 = np.uint32(step & 0xFFFFFFFF)
    c3 = (np.uint64(node_id & 0xFFFF) | ((elem >> np.uint64(32)) << np.uint64(16))).astype(np.uint32)[None, :]
    r0, r1, r2, r3 = philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = _u53(r0, r1)
    if kind == 'U':
        out = u
    elif kind == 'N':
        u2 = _u53(r2, r3)
        out = np.sqrt(-2.0 * np.log(1.0 - u)) * np.cos(2.0 * np.pi * u2)
    elif kind == 'E':
        out = -np.log1p(-u)
    elif kind == 'C':
        out = np.tan(np.pi * (u - 0.5))
    else:
        raise ValueError(kind)
    return out.reshape((len(env_ids),) + tuple(scope_shape))
