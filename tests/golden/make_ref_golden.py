#!/usr/bin/env python3
"""Generates golden outputs from the REFERENCE ITSELF, run in the build container.

oracle/_ref/libviorb_ref.so is /root/reference/src/ORBextractor.cc (and the matcher sources) compiled unmodified by
oracle/refbuild/Makefile against stand-in OpenCV headers; this script runs it on the BASELINE.json configurations and on
the differential-fuzz corpus and commits what it returns, so that the GPU box (where /root/reference does not exist)
can compare against reference outputs even without the prebuilt library:

  ref_extract_<cfg>_seed<k>.npz    keypoints + descriptors + per-level pyramid hashes of ORBextractor::operator()
  ref_extract_hashes.json          sha256(keypoints || descriptors) per case: the large configs and the fuzz corpus

Allocator convention: ascending heap addresses (oracle/refbuild/ref_extractor_capi.cpp), Gaussian taps of OpenCV >= 3.4.
Run:  python tests/golden/make_ref_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle_py as O, ref_py  # noqa: E402
from util import CONFIGS, extraction_digest, fuzz_extract_cases, reference_defined  # noqa: E402
from viorb_b200 import synth  # noqa: E402

FUZZ_CASES, FUZZ_SEED = 200, 0


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run(cfg, img, variant=0):
    h, w, nf, sf, nl, it, mt = CONFIGS[cfg]
    ref_py.set_gaussian_variant(variant)
    ex = O.Extractor(nf, sf, nl, it, mt, _lib=ref_py.lib())
    k, d = ex(img)
    pyr = [sha(ex.pyramid(l)) for l in range(nl)]
    ref_py.set_gaussian_variant(0)
    return k, d, pyr


def main():
    ref_py.set_allocator(1)
    hashes = {"configs": {}, "fuzz": {"cases": FUZZ_CASES, "seed": FUZZ_SEED, "digest": {}}, "cv24": {}}
    for cfg, seed in (("euroc", 0), ("odd", 5), ("kitti12", 7)):
        h, w = CONFIGS[cfg][:2]
        img = synth.frame(h, w, seed)
        k, d, pyr = run(cfg, img)
        np.savez_compressed(os.path.join(HERE, "ref_extract_%s_seed%d.npz" % (cfg, seed)), image_sha=sha(img), keypoints=k,
                            descriptors=d, pyramid_sha=np.array(pyr))
        print(cfg, seed, len(k))
    # KITTI stereo pair of BASELINE configs[1]
    left, right, _ = synth.stereo_pair(376, 1241, 7)
    for name, img in (("kitti_left", left), ("kitti_right", right)):
        k, d, pyr = run("kitti", img)
        hashes["configs"][name] = {"n": len(k), "digest": extraction_digest(k, d), "pyramid": sha("".join(pyr).encode())}
    for cfg, seed in (("euroc", 1), ("euroc", 4095), ("hd", 0), ("uhd", 0)):
        h, w = CONFIGS[cfg][:2]
        k, d, pyr = run(cfg, synth.frame(h, w, seed))
        hashes["configs"]["%s_seed%d" % (cfg, seed)] = {"n": len(k), "digest": extraction_digest(k, d), "pyramid": sha("".join(pyr).encode())}
        print(cfg, seed, len(k))
    # the OpenCV-2.4 Gaussian taps (the OpenCV the reference pins): descriptors change, keypoints do not
    k, d, _ = run("euroc", synth.frame(480, 752, 0), variant=1)
    hashes["cv24"]["euroc_seed0"] = {"n": len(k), "digest": extraction_digest(k, d)}
    for c, img, (nf, sf, nl, it, mt) in fuzz_extract_cases(FUZZ_CASES, FUZZ_SEED):
        if not reference_defined(img.shape[0], img.shape[1], sf, nl):
            continue
        k, d = O.Extractor(nf, sf, nl, it, mt, _lib=ref_py.lib())(img)
        hashes["fuzz"]["digest"][str(c)] = extraction_digest(k, d)
    assert ref_py.arena_overflows() == 0
    with open(os.path.join(HERE, "ref_extract_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1, sort_keys=True)
    print("fuzz cases hashed:", len(hashes["fuzz"]["digest"]))


if __name__ == "__main__":
    main()
