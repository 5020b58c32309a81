"""CPU tests of the padded, bank-scheduled copies of single-column CSR lists (lowering.csr_index,
lowering._bank_schedule; csrc/rddl_device.cuh SUMLOOP reads them) and of the Box-Muller pairing of
normal variates (oracle/philox_np.base_variates restates csrc/rddl_device.cuh: rddl_randn_pair)."""
import numpy as np

from oracle import philox_np
from pyrddlgym_b200 import lowering, workloads
import optable_emulator


def _sections(blob):
    """(csrs, pool) of a packed op table."""
    secs = optable_emulator._sections(blob)
    return secs[5], secs[6]


def _rows(pool, rowptr_off, col_off, nrows, u16=False):
    rp = pool[rowptr_off:rowptr_off + nrows + 1]
    cols = pool[col_off:].view('<u2') if u16 else pool[col_off:]
    return [np.asarray(cols[rp[r]:rp[r + 1]], dtype=np.int64) for r in range(nrows)], rp


def test_bank_schedule_permutes_rows_and_separates_bank_groups():
    rng = np.random.default_rng(3)
    n = rng.integers(0, 40, size=203)
    rp = np.concatenate([[0], np.cumsum(n)])
    cols = np.concatenate([np.sort(rng.choice(1000, size=k, replace=False)) for k in n]).astype('<i4')
    out = lowering._bank_schedule(rp, cols)
    conflicts = lambda c: sum(
        len(g) - len(set(g)) for b in range(0, 200, 4) for k in range(40)
        for g in [[int(c[rp[r] + k]) % 4 for r in range(b, b + 4) if k < n[r]]])
    for r in range(203):
        assert sorted(out[rp[r]:rp[r + 1]]) == sorted(cols[rp[r]:rp[r + 1]])
    assert conflicts(out) < 0.45 * conflicts(cols)


def test_reservoir_padded_list_is_a_padded_permutation_of_the_list():
    R = 1000
    tab = lowering.lower(workloads.load_program(workloads.reservoir(R)))
    csrs, pool = _sections(tab.blob)
    padded = [c for c in csrs if c['rowptr4_off'] != 0]
    assert padded, 'the inflow sum of Reservoir is a single-column CSR list'
    for c in padded:
        assert c['ncols'] == 1 and c['sentinel4'] == R and c['col4_u16'] == 1      # 1000 fits 16 bits
        assert c['col4_off'] % 4 == 0 and c['rowptr4_off'] % 4 == 0                # 16-byte aligned in the pool
        rows, _ = _rows(pool, int(c['rowptr_off']), int(c['col_off'][0]), R)
        rows4, rp4 = _rows(pool, int(c['rowptr4_off']), int(c['col4_off']), R, u16=True)
        assert np.all(rp4 % 8 == 0)                                                # whole 128-bit groups
        for a, b in zip(rows, rows4):
            real = b[b != R]
            assert sorted(real) == sorted(a) and len(b) - len(real) < 8
            assert np.all(b[len(real):] == R)                                      # filler only at the end


def test_bank_schedule_keeps_wide_indices():
    rp = np.array([0, 2, 5], dtype=np.int64)
    cols = np.array([70000, 3, 1, 2, 69999], dtype='<i4')
    out = lowering._bank_schedule(rp, cols)
    assert sorted(out[:2]) == [3, 70000] and sorted(out[2:]) == [1, 2, 69999]


def test_normal_variates_pair_elements_eight_apart():
    env = np.arange(3)
    z = philox_np.base_variates('N', 11, env, 5, 77, (40,))
    # partners e, e + 8 (bit 3 of e clear) are the cosine and sine branch of one Box-Muller transform:
    # z[e]^2 + z[e + 8]^2 = -2 ln(1 - u1) of the shared Philox block, and elements of different blocks differ
    r2 = z[:, :8] ** 2 + z[:, 8:16] ** 2
    c0 = np.arange(8, dtype=np.uint32)[None, :]
    c1 = env.astype(np.uint32)[:, None]
    r0, r1, _, _ = philox_np.philox4x32_10(c0, c1, np.uint32(5), np.uint32(77), 11, 0)
    np.testing.assert_allclose(r2, -2.0 * np.log(1.0 - philox_np._u53(r0, r1)), rtol=1e-12)
    assert len(np.unique(np.round(z, 12))) == z.size
    big = philox_np.base_variates('N', 1, np.arange(64), 0, 9, (4096,))
    assert abs(big.mean()) < 0.01 and abs(big.std() - 1.0) < 0.01
    assert abs(np.corrcoef(big[:, :8].ravel(), big[:, 8:16].ravel())[0, 1]) < 0.1


def test_packed_actions_is_opt_in_and_covers_the_plane_read_action_fluents():
    """lowering.lower(packed_actions=True): bool action fluents the plane launch reads get the packed dtype code."""
    p = workloads.load_program(workloads.wildfire(32))
    plain = lowering.lower(p)
    packed = lowering.lower(p, packed_actions=True)
    assert not ({'put-out', 'cut-out'} & set(plain.packed))
    assert {'put-out', 'cut-out'} <= set(packed.packed)
    tens = optable_emulator._sections(packed.blob)[0]
    for name in ('put-out', 'cut-out'):
        assert int(tens[packed.tensor_id[name]]['dtype']) == 3
    # a real-valued action fluent is never packed
    r = lowering.lower(workloads.load_program(workloads.reservoir(8)), packed_actions=True)
    assert 'release' not in set(r.packed)
