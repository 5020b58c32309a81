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
    Returns float64 array of shape [len(env_ids), *scope_shape].

    Normals come in Box-Muller pairs: elements e and e + 8 (bit 3 of e clear) share the Philox block of counter
    q = e with bit 3 removed and take the cosine and the sine branch (csrc/rddl_device.cuh: rddl_randn_pair,
    rddl_randn_counter) -- the fused scalar kernels compute both from one logarithm and one square root."""
    n_elem = int(np.prod(scope_shape, dtype=np.int64)) if len(scope_shape) else 1
    elem = np.arange(n_elem, dtype=np.uint64)
    ctr = (((elem >> np.uint64(4)) << np.uint64(3)) | (elem & np.uint64(7))) if kind == 'N' else elem
    env = np.asarray(env_ids, dtype=np.uint64)[:, None]
    c0 = (ctr & MASK).astype(np.uint32)[None, :]
    c1 = (env & MASK).astype(np.uint32)
    c2 = np.uint32(step & 0xFFFFFFFF)
    c3 = (np.uint64(node_id & 0xFFFF) | ((ctr >> np.uint64(32)) << np.uint64(16))).astype(np.uint32)[None, :]
    r0, r1, r2, r3 = philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = _u53(r0, r1)
    if kind == 'U':
        out = u
    elif kind == 'N':
        u2 = _u53(r2, r3)
        rad = np.sqrt(-2.0 * np.log(1.0 - u))
        odd = ((elem >> np.uint64(3)) & np.uint64(1)).astype(bool)[None, :]
        out = rad * np.where(odd, np.sin(2.0 * np.pi * u2), np.cos(2.0 * np.pi * u2))
    elif kind == 'E':
        out = -np.log1p(-u)
    elif kind == 'C':
        out = np.tan(np.pi * (u - 0.5))
    else:
        raise ValueError(kind)
    return out.reshape((len(env_ids),) + tuple(scope_shape))


def bernoulli_thresholds(p):
    """float64 probabilities -> (T uint64 = floor(p * 2^64) clipped, always = p >= 1).
    Mirrors pyrddlgym_b200/plane_lowering.py:PlaneBuilder.fin_const."""
    p = np.asarray(p, dtype=np.float64)
    flat = p.reshape(-1)
    T = np.zeros(flat.shape, dtype=np.uint64)
    always = np.zeros(flat.shape, dtype=bool)
    for i, pv in enumerate(flat.tolist()):
        if not (pv >= 0.0 and pv <= 1.0):
            pv = 0.0 if not (pv > 0.0) else 1.0
        if pv >= 1.0:
            always[i] = True
            T[i] = np.uint64((1 << 64) - 1)
        elif pv > 0.0:
            T[i] = np.uint64(int(pv * float(1 << 64)))
    return T.reshape(p.shape), always.reshape(p.shape)


BERN_LEVELS = 8


def bitsliced_less_than(seed, env_ids, step, node_id, T):
    """[U < T] per cell for the bit-sliced uniform U of the plane kernels (csrc/rddl_plane.cuh).

    T: uint64 [len(env_ids), n_cells] with n_cells % 32 == 0 (C-order cells of the scope). Cell c of
    env e first compares T's top BERN_LEVELS bits, MSB first, against random bits: bit j of the cell
    is bit (c % 32) of word (j % 4) of
    Philox4x32-10(key = seed, counter = (c // 32, e, step, (node & 0xffff) | ((0x8000 + j // 4) << 16))).
    The first level where the random bit differs decides (random bit 0 -> True). A cell still
    tied after BERN_LEVELS levels draws u64 = (x << 32) | y from
    Philox4x32-10(counter = (c, e, step, (node & 0xffff) | (0xC000 << 16))) and is True iff
    u64 < (T << BERN_LEVELS) mod 2^64. Comparisons of one cell against several thresholds under the
    same node share these bits (one uniform per cell)."""
    T = np.asarray(T, dtype=np.uint64)
    N, n_cells = T.shape
    assert n_cells % 32 == 0
    G = n_cells // 32
    env = (np.asarray(env_ids, dtype=np.uint64)[:, None] & MASK).astype(np.uint32)
    g = np.arange(G, dtype=np.uint32)[None, :]
    lane = np.arange(32, dtype=np.uint32)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    st = np.uint32(step & 0xFFFFFFFF)
    T3 = T.reshape(N, G, 32)
    res = np.zeros((N, G, 32), dtype=bool)
    und = np.ones((N, G, 32), dtype=bool)
    for b in range(BERN_LEVELS // 4):
        c3 = np.uint32((node_id & 0xFFFF) | ((0x8000 + b) << 16))
        words = philox4x32_10(g, env, st, c3, k0, k1)
        for q in range(4):
            j = 4 * b + q
            r = ((words[q][:, :, None] >> lane[None, None, :]) & np.uint32(1)).astype(bool)
            t = ((T3 >> np.uint64(63 - j)) & np.uint64(1)).astype(bool)
            res |= und & ~r & t
            und &= ~(r ^ t)
    cell = np.arange(n_cells, dtype=np.uint32)[None, :]
    c3 = np.uint32((node_id & 0xFFFF) | (0xC000 << 16))
    x, y, _, _ = philox4x32_10(cell, env, st, c3, k0, k1)
    u64 = (x.astype(np.uint64) << np.uint64(32)) | y.astype(np.uint64)
    tail = u64 < (T << np.uint64(BERN_LEVELS))
    return np.where(und.reshape(N, n_cells), tail, res.reshape(N, n_cells))


def bernoulli_bitsliced(seed, env_ids, step, node_id, p):
    """Bit-sliced Bernoulli sampler of the plane interpreter (csrc/rddl_plane.cuh: PBERN):
    [U < floor(p * 2^64)] with the bit-sliced uniform of ``bitsliced_less_than``; cells with p >= 1
    are always True. p: float64 [len(env_ids), n_cells], n_cells % 32 == 0."""
    T, always = bernoulli_thresholds(np.asarray(p, dtype=np.float64))
    return bitsliced_less_than(seed, env_ids, step, node_id, T) | always


# ---------------------------------------------------------------------------------------
# Device-side random policy (pyrddlgym_b200/csrc/rddl_device.cuh: rddl_policy_value,
# rddl_policy_pick, rddl_policy_floyd; include/rddl_b200.h: rddl_policy_t) -- the batched
# counterpart of RandomAgent.sample_action (/root/reference/pyRDDLGym/core/policy.py:166-178):
# every action grounding is drawn from its range, then -- under a max-nondef rule -- only k
# groundings chosen uniformly without replacement keep their draw.
# ---------------------------------------------------------------------------------------

POLICY_NODE = 0xF000
POLICY_SELECT_NODE = 0xFF00
POL_BERNOULLI, POL_UNIFORM_REAL, POL_UNIFORM_INT, POL_DEFAULT = 1, 2, 3, 4


def policy_select(seed, env_ids, step, k, total):
    """int64 [len(env_ids), k]: the groundings of [0, total) that keep their draw. Candidate i is
    umulhi64((x << 32) | y, total - k + i + 1) of Philox(counter = (i, env, step, POLICY_SELECT_NODE));
    Floyd's rule: a candidate already chosen is replaced by total - k + i."""
    env = (np.asarray(env_ids, dtype=np.uint64) & MASK).astype(np.uint32)[:, None]
    i = np.arange(k, dtype=np.uint32)[None, :]
    x, y, _, _ = philox4x32_10(i, env, np.uint32(step & 0xFFFFFFFF), np.uint32(POLICY_SELECT_NODE),
                               seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    sel = np.zeros((len(env_ids), k), dtype=np.int64)
    for e in range(len(env_ids)):
        chosen = []
        for q in range(k):
            j = total - k + q
            u64 = (int(x[e, q]) << 32) | int(y[e, q])
            t = (u64 * (j + 1)) >> 64
            if t in chosen:
                t = j
            chosen.append(t)
        sel[e] = chosen
    return sel


def policy_actions(entries, max_nondef, seed, env_ids, step):
    """entries: list of dicts {tensor (id), kind, lo, hi, dflt, shape, dtype ('b' | 'i' | 'f')} in
    policy order. Returns {tensor id: array [len(env_ids), *shape]} -- what rddl_policy_sample writes
    and what the kernels draw in registers under rddl_rollout_policy."""
    N = len(env_ids)
    sizes = [int(np.prod(e['shape'], dtype=np.int64)) if len(e['shape']) else 1 for e in entries]
    total = sum(sizes)
    k = 0
    kinds = [e['kind'] for e in entries]
    if max_nondef is not None and 0 <= max_nondef < total:
        if max_nondef == 0:
            kinds = [POL_DEFAULT] * len(entries)
        else:
            k = int(max_nondef)
    sel = policy_select(seed, env_ids, step, k, total) if k else None
    # Two consecutive Bernoulli entries of the same size are drawn from ONE uniform per cell (half the
    # random bits): a = [U < T(p)], b = [T(p) - T(pq) <= U < T(p) - T(pq) + T(q)], T(x) = floor(x 2^64) --
    # P(a) = p, P(b) = q, P(a and b) = pq: independent, like the agent's per-key draws. U lives under
    # the first entry's node. (Not when the two intervals would come within 2^-32 of covering [0, 1).)
    pair = {}
    i = 0
    while i + 1 < len(entries):
        a, b = entries[i], entries[i + 1]
        if (kinds[i] == POL_BERNOULLI and kinds[i + 1] == POL_BERNOULLI and sizes[i] == sizes[i + 1]
                and a['lo'] < 1.0 and b['lo'] < 1.0 and (1.0 - a['lo']) * (1.0 - b['lo']) >= 2.0 ** -32):
            pair[i + 1] = i
            i += 2
        else:
            i += 1
    out = {}
    base = 0
    for idx, (e, kind, n) in enumerate(zip(entries, kinds, sizes)):
        npdt = {'b': np.bool_, 'i': np.int64, 'f': np.float64}[e['dtype']]
        dflt = np.asarray(e['dflt']).astype(npdt)
        node = POLICY_NODE + int(e['tensor'])
        if kind == POL_DEFAULT:
            v = np.full((N, n), dflt, dtype=npdt)
        elif kind == POL_BERNOULLI:
            pad = (-n) % 32      # the bit-sliced mapping is defined per 32-cell word; a ragged last word is cut
            thr = lambda x: int(bernoulli_thresholds(np.array([x]))[0][0])
            if idx in pair:
                first = entries[pair[idx]]
                node = POLICY_NODE + int(first['tensor'])
                lo = thr(float(first['lo'])) - thr(float(first['lo']) * float(e['lo']))
                hi = lo + thr(float(e['lo']))
                below = bitsliced_less_than(seed, env_ids, step, node, np.full((N, n + pad), lo, dtype=np.uint64))
                v = (bitsliced_less_than(seed, env_ids, step, node, np.full((N, n + pad), hi, dtype=np.uint64)) & ~below)[:, :n]
            else:
                p = np.full((N, n + pad), float(e['lo']), dtype=np.float64)
                v = bernoulli_bitsliced(seed, env_ids, step, node, p)[:, :n]
        else:
            u = base_variates('U', seed, env_ids, step, node, (n,))
            if kind == POL_UNIFORM_REAL:
                v = float(e['lo']) + (float(e['hi']) - float(e['lo'])) * u
            else:
                lo, hi = int(e['lo']), int(e['hi'])
                v = np.minimum(lo + np.floor(u * float(hi - lo + 1)).astype(np.int64), hi)
        if k and kind != POL_DEFAULT:
            g = base + np.arange(n, dtype=np.int64)[None, :]
            keep = (sel[:, :, None] == g[:, None, :]).any(axis=1)
            v = np.where(keep, v, dflt)
        out[e['tensor']] = np.asarray(v, dtype=npdt).reshape((N,) + tuple(e['shape']))
        base += n
    return out
