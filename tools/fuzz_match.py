#!/usr/bin/env python3
"""Differential fuzz of the brute-force top-2 search (viorb_hamming_top2): random query / map sizes around the kernel's
tile, slice and small-query boundaries, heavy ties, index bases; every record against the CPU oracle."""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import oracle_py as O  # noqa: E402
from viorb_b200 import api  # noqa: E402


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    m = api.ORBmatcher(ctx=api.Context(0))
    bad = 0
    for c in range(cases):
        Q = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 127, 128, 129, 255, 256, 257, 1000, int(rng.integers(1, 1500))]))
        M = int(rng.choice([1, 2, 255, 256, 257, 1023, 4097, int(rng.integers(1, 300000))]))
        dmap = rng.integers(0, 256, (M, 32)).astype(np.uint8)
        if rng.random() < 0.5:                      # few distinct descriptors: ties everywhere
            dmap = dmap[rng.integers(0, max(1, M // 50 + 1), M)]
        q = rng.integers(0, 256, (Q, 32)).astype(np.uint8)
        k = Q // 2
        if k:
            q[:k] = dmap[rng.integers(0, M, k)]
        base = int(rng.choice([0, 0, 12345, 2 ** 30]))
        got = m.hamming_top2(q, dmap, base)
        want = O.hamming_top2(q, dmap, base, nthreads=8)
        ok = all((got[f] == want[f]).all() for f in ("d1", "i1", "d2", "i2"))
        if not ok:
            bad += 1
            print("MISMATCH case %d: Q=%d M=%d base=%d" % (c, Q, M, base))
    print("fuzz_match: %d cases, %d mismatches" % (cases, bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
