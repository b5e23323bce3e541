#!/usr/bin/env python3
"""Generates the committed golden fixtures from REAL OpenCV (python cv2 4.13) in the build container.

  primitives.npz            -- cv2.resize / GaussianBlur / copyMakeBorder / FAST / fastAtan2 on a small image
  extract_<cfg>_seed<k>.npz -- full ORBextractor::operator() output of tests/cv2_restatement.py (cv2 primitives
                               glued as in src/ORBextractor.cc) for synthetic frame (h, w, seed)

Run:  python tests/golden/make_golden.py      (needs cv2; the fixtures travel to the GPU box, cv2 need not)
"""
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cv2_restatement as R  # noqa: E402
from util import CONFIGS, load_pattern  # noqa: E402
from viorb_b200 import synth  # noqa: E402

KEYPOINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def primitives():
    cv2.setNumThreads(1)
    img = synth.frame(96, 128, 42)
    out = dict(img=img)
    out["resize_107x80"] = cv2.resize(img, (107, 80), interpolation=cv2.INTER_LINEAR)
    out["resize_53x41"] = cv2.resize(img, (53, 41), interpolation=cv2.INTER_LINEAR)
    out["blur7"] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    out["border19"] = cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    for t in (20, 7):
        k = cv2.FastFeatureDetector_create(t, True).detect(img)
        out["fast%d" % t] = np.array([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in k], np.int32).reshape(-1, 3)
    rng = np.random.default_rng(5)
    yx = rng.integers(-1248480, 1248481, size=(4000, 2)).astype(np.float32)
    yx[:8] = [[0, 0], [0, 1], [1, 0], [0, -1], [-1, 0], [1, 1], [-1, -1], [5, -5]]
    out["atan2_yx"] = yx
    out["atan2_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), **out)


def extraction(cfg, seed):
    h, w, nf, sf, nl, it, mt = CONFIGS[cfg]
    pat, _ = load_pattern()
    img = synth.frame(h, w, seed)
    p = R.Params(nf, sf, nl, it, mt)
    levels, cand, kps, desc = R.extract(p, img, pat)
    k = np.zeros(len(kps), KEYPOINT)
    for i, d in enumerate(kps):
        k[i] = (d["x"], d["y"], d["size"], d["angle"], d["response"], d["octave"], -1)
    np.savez_compressed(
        os.path.join(HERE, "extract_%s_seed%d.npz" % (cfg, seed)),
        image_sha=sha(img), keypoints=k, descriptors=desc,
        pyramid_sha=np.array([sha(l) for l in levels]),
        cand_count=np.array([len(c) for c in cand], np.int32),
        cand_sha=np.array([sha(np.array([(c[0], c[1], c[2]) for c in cl], np.int32).reshape(-1, 3)) for cl in cand]),
        quota=np.array(p.quota, np.int32), umax=np.array(p.umax, np.int32))
    print(cfg, seed, len(k))


if __name__ == "__main__":
    primitives()
    extraction("euroc", 0)
    extraction("odd", 5)
    extraction("kitti12", 7)
