#!/usr/bin/env python3
"""Differential fuzz of the whole ORBextractor::operator(): random shapes, parameters and image statistics, GPU (C ABI)
against the CPU oracle, every output byte.  usage: tools/fuzz_extract.py [cases] [seed]"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import oracle_py as O  # noqa: E402
from viorb_b200 import api  # noqa: E402
from util import fuzz_extract_cases  # noqa: E402


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    verbose = len(sys.argv) > 3
    ctx = api.Context(0)
    bad = 0
    skipped = 0
    total_kp = 0
    for c, img, (nf, sf, nl, it, mt) in fuzz_extract_cases(cases, seed):
        h, w = img.shape
        if verbose:
            print("case %d: %dx%d nf=%d sf=%.3f nl=%d th=%d/%d mode=%d" % (c, w, h, nf, sf, nl, it, mt, 1 + c % 2), flush=True)
        try:
            ex = api.ORBextractor(nf, sf, nl, it, mt, ctx=ctx)
            ex.set_describe_mode(1 + c % 2)      # odd cases blur whole levels, even cases blur per keypoint
            kg, dg = ex(img)
        except api.ViorbError as e:
            if e.code == -4:         # outside the documented envelope (tiny top level, portrait aspect): refused, not guessed
                skipped += 1
                continue
            raise
        ref = O.Extractor(nf, sf, nl, it, mt)
        kr, dr = ref(img)
        total_kp += len(kr)
        ok = len(kg) == len(kr) and kg.tobytes() == kr.tobytes() and (dg == dr).all()
        if ok and c % 4 == 0 and h * w <= 600 * 800:
            # the batch kernels (128 x 64 pyramid tiles, several keypoints per warp, whole-level blur by default) on the same
            # geometry: nine frames in passes of nine, the case's image and its mirror alternating
            ex.set_describe_mode(c % 3)
            imgs = np.stack([img if i % 2 == 0 else np.ascontiguousarray(img[:, ::-1]) for i in range(9)])
            bk, bd, bn = ex.extract_batch(imgs)
            km, dm = ref(imgs[1])
            for i in range(9):
                kx, dx = (kr, dr) if i % 2 == 0 else (km, dm)
                ok = ok and bn[i] == len(kx) and bk[i, :bn[i]].tobytes() == kx.tobytes() and (bd[i, :bn[i]] == dx).all()
            total_kp += 4 * len(kr) + 4 * len(km)
        if not ok:
            bad += 1
            print("MISMATCH case %d: %dx%d nf=%d sf=%.2f nl=%d th=%d/%d n=%d/%d" % (c, w, h, nf, sf, nl, it, mt, len(kg), len(kr)))
        ex.close()
    print("fuzz: %d cases, %d outside the envelope, %d keypoints compared, %d mismatches" % (cases, skipped, total_kp, bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
