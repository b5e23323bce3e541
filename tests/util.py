"""Shared helpers for the test-suite."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_pattern():
    """512 sampling points (x, y) as int array [512, 2] parsed from include/viorb_orb_pattern.h"""
    txt = open(os.path.join(ROOT, "include", "viorb_orb_pattern.h")).read()
    body = txt.split("VIORB_ORB_PATTERN_INIT {", 1)[1].split("}", 1)[0]
    nums = np.array([int(v) for v in re.findall(r"-?\d+", body)], np.int8)
    assert nums.size == 1024
    return nums.reshape(512, 2).astype(np.int32), nums


CONFIGS = {
    # name: (h, w, nfeatures, scale, levels, iniTh, minTh)   -- BASELINE.json configs / reference YAMLs
    "euroc": (480, 752, 1000, 1.2, 8, 20, 7),     # Examples/Monocular/EuRoC.yaml:29-42
    "kitti": (376, 1241, 2000, 1.2, 8, 20, 7),    # Examples/Stereo/KITTI00-02.yaml:38-51
    "hd": (1080, 1920, 5000, 1.2, 8, 20, 7),
    "uhd": (2160, 3840, 5000, 1.2, 8, 20, 7),
    "odd": (360, 640, 777, 1.2, 8, 20, 7),
    "kitti12": (376, 1241, 2000, 1.2, 8, 12, 7),  # Examples/Stereo/KITTI04-12.yaml:50
}


# ---------------------------------------------------------------- differential-fuzz corpus of the extractor
def fuzz_image(rng, h, w):
    """five image statistics: synthetic scene, pure noise, low contrast (forces the 20 -> 7 retry), checkerboard, half flat"""
    from viorb_b200 import synth
    kind = rng.integers(0, 5)
    if kind == 0:
        return synth.frame(h, w, int(rng.integers(0, 1 << 30)))
    if kind == 1:
        return rng.integers(0, 256, (h, w)).astype(np.uint8)
    if kind == 2:
        img = synth.frame(h, w, int(rng.integers(0, 1 << 30))).astype(np.int32)
        return np.clip((img - 128) * 0.15 + 128, 0, 255).astype(np.uint8)
    if kind == 3:
        yy, xx = np.mgrid[0:h, 0:w]
        s = int(rng.integers(3, 17))
        return (((yy // s + xx // s) & 1) * int(rng.integers(30, 255))).astype(np.uint8)
    img = synth.frame(h, w, int(rng.integers(0, 1 << 30)))
    img[:, : w // 2] = int(rng.integers(0, 256))
    return img


def fuzz_extract_cases(cases, seed, max_h=900, max_w=1400):
    """yields (index, image, (nfeatures, scale, levels, iniTh, minTh)) -- random shapes, parameters and image statistics;
    the same stream as tools/fuzz_extract.py so the corpus hashed in tests/golden/ref_extract_hashes.json is reproducible"""
    rng = np.random.default_rng(seed)
    for c in range(cases):
        h, w = int(rng.integers(96, max_h)), int(rng.integers(128, max_w))
        nl = int(rng.integers(2, 9))
        sf = float(np.float32(rng.choice([1.1, 1.2, 1.2, 1.25, 1.33, 1.5])))
        nf = int(rng.integers(100, 4000))
        it = int(rng.integers(8, 128)) if rng.random() < 0.3 else int(rng.integers(8, 40))
        mt = int(rng.integers(2, it + 1))
        img = fuzz_image(rng, h, w)
        yield c, img, (nf, sf, nl, it, mt)


def extraction_digest(kps, desc):
    """sha256 over the keypoint records (28 B each) followed by the descriptor rows: one value per frame result"""
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(kps).tobytes())
    h.update(np.ascontiguousarray(desc, np.uint8).tobytes())
    return h.hexdigest()


def reference_defined(h, w, sf, nl):
    """False where the reference itself is undefined: a level whose detection window is smaller than one 30-px cell
    (nCols or nRows = 0, src/ORBextractor.cc:784-787) or taller than wide enough that round(W/H) = 0 quadtree roots
    (:543, vpIniNodes[...] then indexes an empty vector).  The product refuses exactly these with VIORB_ERR_UNSUPPORTED."""
    s = np.float32(1.0)
    for l in range(nl):
        if l:
            s = np.float32(s * np.float32(sf))
        inv = np.float32(1.0) / s
        wl, hl = int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))
        W, H = wl - 32, hl - 32
        if W < 30 or H < 30:
            return False
        if int(np.floor(np.float32(W) / np.float32(H) + np.float32(0.5))) < 1:
            return False
    return True
