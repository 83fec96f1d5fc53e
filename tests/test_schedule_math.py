"""The arithmetic behind the two-key band pruning of k_schedule (csrc/kernels.cuh, K2 / K2b), checked on the CPU:
  (1) |A xor B| >= |s_A - s_B| + |(c_A - s_A) - (c_B - s_B)| for any column subset H (c = |A|, s = |A n H|),
  (2) for D = c_B - c_A, the s_B allowed by (1) at distance <= d are exactly s_A - (d - D)//2 .. s_A + (D + d)//2,
the integer divisions the kernel uses."""
import numpy as np
import pytest


def test_symmetric_difference_dominates_the_two_half_cardinality_gaps():
    rng = np.random.default_rng(3)
    for _ in range(2000):
        n_cols = int(rng.integers(4, 60))
        in_h = rng.random(n_cols) < rng.random()
        a = rng.random(n_cols) < rng.random()
        b = a.copy()
        flip = rng.choice(n_cols, size=int(rng.integers(0, min(6, n_cols))), replace=False)
        b[flip] = ~b[flip]
        dist = int((a ^ b).sum())
        ca, cb, sa, sb = int(a.sum()), int(b.sum()), int((a & in_h).sum()), int((b & in_h).sum())
        assert dist >= abs(sa - sb) + abs((ca - sa) - (cb - sb))


@pytest.mark.parametrize("d", [0, 1, 2, 3, 7])
def test_partner_interval_of_the_second_key(d):
    for D in range(-d, d + 1):
        feasible = [ds for ds in range(-3 * d - 2, 3 * d + 3) if abs(ds) + abs(D - ds) <= d]
        lo, hi = -((d - D) // 2), (D + d) // 2
        assert feasible == list(range(lo, hi + 1)), (d, D, feasible, lo, hi)
    # outside the cardinality band nothing is feasible
    for D in (-d - 1, d + 1):
        assert not [ds for ds in range(-3 * d - 2, 3 * d + 3) if abs(ds) + abs(D - ds) <= d]
