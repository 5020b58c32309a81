"""TEST INFRASTRUCTURE ONLY -- independent NumPy restatement of the Wildfire CPFs
(/root/reference/pyRDDLGym/examples/run_build.py:20-43) as a 3x3 stencil, for grids the
reference cannot build (hard part 1). Cross-checked against the
reference-generated goldens at n = 8 and 32 in tests/test_oracle_golden.py."""
import numpy as np


def wildfire_step(burning, oof, put, cut, target, u, costs=(-5.0, -10.0, -100.0, -5.0)):
    """All planes [N, n, n] bool (target [n, n]); u: injected uniforms of the Bernoulli node.
    Returns (burning', out-of-fuel', reward)."""
    N, H, W = burning.shape
    pad = np.pad(burning.astype(np.int64), ((0, 0), (1, 1), (1, 1)))
    k = np.zeros((N, H, W), dtype=np.int64)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            if (dx, dy) != (0, 0):
                k += pad[:, 1 + dx:1 + dx + H, 1 + dy:1 + dy + W]
    p = 1.0 / (1.0 + np.exp(4.5 - k))
    spark = u <= p
    spread = np.where(target & ~(k > 0), False, spark)
    b1 = np.where(put, False, np.where(~oof & ~burning, spread, burning))
    f1 = oof | burning | (~target & cut)
    c_cut, c_put, c_tb, c_nb = costs
    reward = (c_cut * cut.sum((1, 2)) + c_put * put.sum((1, 2))
              + c_tb * ((burning | oof) & target).sum((1, 2)) + c_nb * (burning & ~target).sum((1, 2)))
    return b1, f1, reward.astype(np.float64)
