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


@pytest.mark.parametrize("max_dist", [1, 2, 3])
@pytest.mark.parametrize("n_keys", [1, 2, 3])
def test_schedule_mirror_loses_no_edge(max_dist, n_keys):
    """tools/schedule_sim.py restates k_schedule on the CPU (the GPU suite pins the kernel's tile-pair counter to it):
    every edge the oracle finds must lie in a tile pair the schedule lists, and more keys never list more pairs"""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import oracle
    import schedule_sim
    from breakfast_b200 import synth
    indptr, indices, _ = synth.generate(6000, seed=11).csr()
    order, runs = schedule_sim.schedule(indptr, indices, max_dist, n_keys)
    pos = np.empty(len(order), dtype=np.int64)
    pos[order] = np.arange(len(order))
    covered = set()
    for tile, first, end in runs:
        covered.update((tile, j) for j in range(first, end))
    src, dst = oracle.edges(indptr, indices, max_dist)
    assert src.size > 100
    a, b = np.minimum(pos[src], pos[dst]) // 128, np.maximum(pos[src], pos[dst]) // 128
    assert all((int(i), int(j)) in covered for i, j in zip(a, b))
    if n_keys > 1:
        assert len(covered) <= schedule_sim.tile_pairs(indptr, indices, max_dist, n_keys - 1)


def test_packed_pair_threshold_arithmetic_of_the_tensor_core_level1():
    """CPU restatement of k_pairs_l1_imma2's survivor test (csrc/kernels.cuh): one accumulator holds
    acc = d1 + 64 d2 (d = 32 - 2 popc of the two pairs); w = (acc + 63) * 66560 puts a monotone image of acc in the high
    16 bits and ((d1 + 63) mod 64) << 10 in the low 16 bits; both halves are compared as signed 16-bit numbers.
    The test must pass exactly the accumulators with d1 >= thr or d2 >= thr, plus the alias d1 = -32."""
    pop = np.arange(0, 33)
    d = 32 - 2 * pop
    d1, d2 = np.meshgrid(d, d, indexing="ij")
    acc = (d1 + 64 * d2).astype(np.int64)
    w = ((acc + 63) * 66560).astype(np.int64)
    assert np.abs(w).max() < 2 ** 31
    hi = w >> 16                                    # arithmetic shift: floor
    lo = ((w & 0xFFFF) ^ 0x8000) - 0x8000           # low half as a signed 16-bit number
    assert hi.max() < 2 ** 15 and hi.min() >= -2 ** 15
    for max_dist in range(0, 32):
        thr = 32 - 2 * max_dist
        hi_thr, lo_thr = 65 * thr + 31, 1024 * (thr - 1)
        assert -2 ** 15 <= lo_thr < 2 ** 15
        got = (hi >= hi_thr) | (lo >= lo_thr)
        want = (d1 >= thr) | (d2 >= thr)
        assert not (want & ~got).any(), f"max_dist {max_dist}: a survivor is lost"
        extra = got & ~want
        assert (d1[extra] == -32).all(), f"max_dist {max_dist}: false positives beyond the complement alias"
    # the packed int8 operand: e1 + 64 e2 with e = +1 / -1 fits an int8 and the byte trick of expand_pm1_pair gives it
    for a in (0, 1):
        for b in (0, 1):
            byte = 0x41 ^ (0x7E if a else 0) ^ (0x80 if b else 0)
            val = byte - 256 if byte >= 128 else byte
            assert val == (-1 if a else 1) + 64 * (-1 if b else 1)


def test_packed_level1_end_to_end_on_random_folds():
    """The whole level-1 chain of k_pairs_l1_imma2 on the CPU: 32-bit folds -> +-1 int8 row operand, packed int8 column
    operand (expand_pm1_pair: two rows per int8 row), int8 dot products as the MMA computes them (s32 accumulation),
    one multiply + the two signed 16-bit halves, thresholds.  Survivor <=> popc(fa ^ fb) <= max_dist for either of the
    two packed column rows (plus the complement alias, which only adds survivors)."""
    rng = np.random.default_rng(11)
    n_rows, n_cols = 96, 64
    fa = rng.integers(0, 2 ** 32, size=n_rows, dtype=np.uint64).astype(np.uint32)
    fb = rng.integers(0, 2 ** 32, size=n_cols, dtype=np.uint64).astype(np.uint32)
    # plant close pairs, identical folds and a complement
    fb[0] = fa[0]
    fb[1] = fa[1] ^ np.uint32(1 << 7)
    fb[2] = fa[2] ^ np.uint32((1 << 3) | (1 << 30))
    fb[3] = ~fa[3]
    fb[5] = fa[4] ^ np.uint32(1 << 31)

    def pm1(f):   # bit k set -> -1, clear -> +1
        bits = (f[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1
        return (1 - 2 * bits.astype(np.int32)).astype(np.int8)

    ea, eb = pm1(fa), pm1(fb)
    packed = (eb[0::2].astype(np.int32) + 64 * eb[1::2].astype(np.int32))
    assert packed.min() >= -128 and packed.max() <= 127
    packed = packed.astype(np.int8)
    acc = ea.astype(np.int32) @ packed.astype(np.int32).T            # [rows, packed columns], as the s32 accumulators
    popc = np.array([[bin(int(a) ^ int(b)).count("1") for b in fb] for a in fa])
    d = 32 - 2 * popc
    assert np.array_equal(acc, d[:, 0::2] + 64 * d[:, 1::2])
    w = (acc.astype(np.int64) * 66560 + 63 * 66560)
    hi = w >> 16
    lo = ((w & 0xFFFF) ^ 0x8000) - 0x8000
    for max_dist in (0, 1, 2, 3, 5, 8, 15, 16, 31):
        thr = 32 - 2 * max_dist
        got = (hi >= 65 * thr + 31) | (lo >= 1024 * (thr - 1))
        want = (popc[:, 0::2] <= max_dist) | (popc[:, 1::2] <= max_dist)
        assert not (want & ~got).any()
        extra = got & ~want
        assert (popc[:, 0::2][extra] == 32).all()


def test_compact_form_hash_identities_of_the_sketch_pass():
    """k_pack_sketch_rows16 never rebuilds a 17-bit column: it hashes the low 16 bits and adds C << 16 for the columns
    from 65536 on (the multiplicative hash distributes), takes the sketch bit from the top log2(m) bits and the two
    sort-key halves from hash bits 31 and 30.  CPU restatement against the plain-column formulas of k_pack_sketch_rows
    (fold_hash) and tools/schedule_sim.row_keys."""
    C = 2654435761
    cols = np.concatenate([np.arange(0, 131072, 7), [0, 1, 65535, 65536, 65537, 131071]]).astype(np.uint64)
    h_plain = (cols * C) & 0xFFFFFFFF
    lo, hi = cols & 0xFFFF, cols >> 16
    h_compact = (lo * C + hi * ((C << 16) & 0xFFFFFFFF)) & 0xFFFFFFFF
    assert np.array_equal(h_plain, h_compact)
    for log2m in (7, 8):
        bit_plain = h_plain >> (32 - log2m)                       # fold_hash
        word, bit = h_compact >> (37 - log2m), (h_compact >> (32 - log2m)) & 31
        assert np.array_equal(bit_plain, word * 32 + bit)
    # quadrant counters of a row: quad += h >> 30 and in_h1 += h >> 31  ->  s = |row n H1|, t = |row n H2|
    quad, in_h1 = int((h_plain >> 30).sum()), int((h_plain >> 31).sum())
    s, t = int(((h_plain >> 31) & 1).sum()), int(((h_plain >> 30) & 1).sum())
    assert (in_h1, quad - 2 * in_h1) == (s, t)
