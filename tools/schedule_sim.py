"""CPU mirror of the band-pruned tile schedule (csrc/kernels.cuh, k_schedule) — counts the tile pairs a pass evaluates.

  python tools/schedule_sim.py [n_profiles] [max_dist]

prints the tile pairs of the plain cardinality band (1 key), of the shipped schedule (2 keys: cardinality and the
cardinality on the columns whose hash has its top bit set) and of a 3-key variant (next hash bit as third key) that
is not built.  tests/test_gpu_parity.py compares the 2-key count with the counter of the real kernel.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
TILE = 128
KEY_CLAMP = 65535


def row_keys(indptr, indices):
    """(c, s, t) per row: cardinality and the cardinalities on the columns whose multiplicative hash has bit 31 / 30 set"""
    h = (indices.astype(np.uint64) * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
    c = np.diff(indptr).astype(np.int64)
    out = [c]
    for bit in (31, 30):
        b = ((h >> np.uint64(bit)) & np.uint64(1)).astype(np.int64)
        cs = np.concatenate([[0], np.cumsum(b)])
        out.append(cs[indptr[1:]] - cs[indptr[:-1]])
    return out


def _feasible_dt(d, D, Ds):
    return [Dt for Dt in range(-d, d + 1)
            if any(abs(u) + abs(Ds - u) + abs(Dt - u) + abs(D - Ds - Dt + u) <= d for u in range(-d, d + 1))]


def tile_pairs(indptr, indices, max_dist: int, n_keys: int = 2) -> int:
    """tile pairs (I, J >= I) listed by the schedule with n_keys sort keys, triangular case, clamping ignored
    (valid while every cardinality is below 65535)"""
    return sum(end - first for _, first, end in schedule(indptr, indices, max_dist, n_keys)[1])


def schedule(indptr, indices, max_dist: int, n_keys: int = 2):
    """(order, runs): order = row indices in sort-key order, runs = [(I, first, end)] column-tile runs per row tile"""
    d = max_dist
    keys = row_keys(indptr, indices)[:n_keys]
    n = len(keys[0])
    assert n == 0 or keys[0].max() < KEY_CLAMP
    order = np.lexsort(keys[::-1])
    K = [k[order] for k in keys]
    comp = K[0].copy()
    for k in K[1:]:
        comp = comp * 65536 + k
    T = (n + TILE - 1) // TILE
    b_min = comp[::TILE]
    b_max = comp[np.minimum(np.arange(1, T + 1) * TILE, n) - 1]
    width = 65536 ** (n_keys - 1)
    runs = []
    for I in range(T):
        lo_i = [int(k[I * TILE]) for k in K]
        hi_i = [int(k[min((I + 1) * TILE, n) - 1]) for k in K]
        ranges = []
        if n_keys >= 3 and lo_i[0] == hi_i[0] and lo_i[1] == hi_i[1]:
            for D in range(-d, d + 1):
                for Ds in range(-((d - D) // 2), (D + d) // 2 + 1):
                    dt = _feasible_dt(d, D, Ds)
                    if dt and lo_i[0] + D >= 0 and lo_i[1] + Ds >= 0:
                        head = ((lo_i[0] + D) * 65536 + lo_i[1] + Ds) * 65536
                        ranges.append((head + max(lo_i[2] + min(dt), 0), head + hi_i[2] + max(dt)))
        elif n_keys >= 2 and lo_i[0] == hi_i[0]:
            for D in range(-d, d + 1):
                if lo_i[0] + D < 0:
                    continue
                head = (lo_i[0] + D) * 65536
                s_lo, s_hi = max(lo_i[1] - (d - D) // 2, 0), hi_i[1] + (D + d) // 2
                sub = width // 65536
                ranges.append(((head + s_lo) * sub, (head + s_hi) * sub + sub - 1))
        else:
            ranges.append((max(lo_i[0] - d, 0) * width, (hi_i[0] + d) * width + width - 1))
        prev = I
        for lo, hi in sorted(ranges):
            first = max(int(np.searchsorted(b_max, lo, "left")), prev)
            end = max(int(np.searchsorted(b_min, hi, "right")), first)
            if end > first:
                runs.append((I, first, end))
            prev = end
    return order, runs


if __name__ == "__main__":
    from breakfast_b200 import synth
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    indptr, indices, _ = synth.generate(n, seed=1).csr()
    T = (n + TILE - 1) // TILE
    print(f"{n} profiles, max_dist {d}: {T * (T + 1) // 2} tile pairs in the triangle")
    for k in (1, 2, 3):
        print(f"  {k} key(s): {tile_pairs(indptr, indices, d, k)} tile pairs")
