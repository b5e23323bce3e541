#!/usr/bin/env python3
"""Golden digests of BASELINE.json configs[2] -- all 4096 synthetic 752x480 frames (seed = frame index), EuRoC parameters --
written by the REFERENCE ITSELF (oracle/_ref: src/ORBextractor.cc compiled unmodified, ascending-address allocator
convention, OpenCV >= 3.4 Gaussian taps), run in the build container.

  ref_extract_batch4096.json   {"frames": 4096, "n": [keypoints per frame], "digest16": [first 16 hex digits of
                               sha256(keypoints || descriptors) per frame], "total": sha256 over the concatenated full
                               per-frame digests}

bench.py recomputes the same digests from what the GPU path returned on every rank and asserts equality, so the timed
workload is also a parity workload at full size and at every N.   Run:  python tests/golden/make_ref_batch_golden.py
"""
import hashlib
import json
import os
import sys
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle_py as O, ref_py  # noqa: E402
from util import extraction_digest  # noqa: E402
from viorb_b200 import synth  # noqa: E402

FRAMES = 4096


def main():
    ref_py.set_allocator(1)
    nthreads = os.cpu_count() or 1
    digests, counts = [None] * FRAMES, [0] * FRAMES

    def work(t):
        ex = O.Extractor(1000, 1.2, 8, 20, 7, _lib=ref_py.lib())
        for i in range(t, FRAMES, nthreads):
            k, d = ex(synth.frame(480, 752, i))
            digests[i], counts[i] = extraction_digest(k, d), len(k)

    th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert ref_py.arena_overflows() == 0
    total = hashlib.sha256("".join(digests).encode()).hexdigest()
    with open(os.path.join(HERE, "ref_extract_batch4096.json"), "w") as f:
        json.dump({"frames": FRAMES, "n": counts, "digest16": [d[:16] for d in digests], "total": total}, f)
    print("frames", FRAMES, "keypoints", sum(counts), "total", total)


if __name__ == "__main__":
    main()
