"""CPU tests that PIN the oracle (no GPU needed).

 1. every OpenCV primitive restated in oracle/orb_oracle.cpp == the committed cv2-4.13 golden vectors
    (tests/golden/primitives.npz) and, when cv2 is importable, == cv2 live on fresh inputs;
 2. the whole ORBextractor::operator() restatement == the golden outputs produced by
    tests/cv2_restatement.py (real cv2 primitives glued as in src/ORBextractor.cc);
 3. orc_sincosf == this image's glibc sincosf on a dense sample of [0, 2*pi];
 4. constructor tables (quotas, umax) == SURVEY.md Appendix B / A1.
"""
import ctypes
import hashlib
import os

import numpy as np
import pytest

from util import CONFIGS, ROOT, load_pattern
from viorb_b200 import synth

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def corners_to_array(c):
    return np.stack([c["x"], c["y"], c["score"]], 1).astype(np.int32).reshape(-1, 3)


def test_pattern_hash():
    _, raw = load_pattern()
    assert hashlib.sha256(raw.tobytes()).hexdigest() == \
        "2164181aea6ff9ac426ca512d5130d15e1f6e3cd47b1cbdd568bbe1e55d49023"


def test_primitives_vs_golden(oracle):
    g = np.load(os.path.join(GOLD, "primitives.npz"))
    img = g["img"]
    assert (synth.frame(96, 128, 42) == img).all(), "synthetic generator changed: regenerate goldens"
    assert (oracle.resize_linear(img, 107, 80) == g["resize_107x80"]).all()
    assert (oracle.resize_linear(img, 53, 41) == g["resize_53x41"]).all()
    assert (oracle.gaussian7(img) == g["blur7"]).all()
    assert (oracle.border_reflect101(img) == g["border19"]).all()
    for t in (20, 7):
        mine = corners_to_array(oracle.fast9(img, t))
        assert mine.shape == g["fast%d" % t].shape and (mine == g["fast%d" % t]).all()
    yx = g["atan2_yx"]
    mine = np.array([oracle.fast_atan2(y, x) for y, x in yx], np.float32)
    assert (mine.view(np.uint32) == g["atan2_deg"].view(np.uint32)).all()


@pytest.mark.parametrize("shape,seed", [((480, 752), 11), ((313, 1034), 12), ((97, 131), 13)])
def test_primitives_vs_cv2_live(oracle, shape, seed):
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    h, w = shape
    img = synth.frame(h, w, seed)
    dw, dh = int(round(w / 1.2)), int(round(h / 1.2))
    assert (oracle.resize_linear(img, dw, dh) == cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)).all()
    assert (oracle.gaussian7(img) == cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)).all()
    assert (oracle.border_reflect101(img) == cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)).all()
    for t in (20, 12, 7):
        k = cv2.FastFeatureDetector_create(t, True).detect(img)
        ref = np.array([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in k], np.int32).reshape(-1, 3)
        mine = corners_to_array(oracle.fast9(img, t))
        assert mine.shape == ref.shape and (mine == ref).all()


def test_fast_small_cells_vs_cv2(oracle):
    """cell-sized sub-images with a row stride (how ComputeKeyPointsOctTree calls FAST, :809)"""
    cv2 = pytest.importorskip("cv2")
    img = synth.frame(200, 300, 21)
    rng = np.random.default_rng(0)
    for _ in range(60):
        cw, ch = int(rng.integers(7, 45)), int(rng.integers(7, 45))
        x0, y0 = int(rng.integers(0, 300 - cw)), int(rng.integers(0, 200 - ch))
        t = int(rng.choice([20, 7]))
        cell = img[y0:y0 + ch, x0:x0 + cw]
        k = cv2.FastFeatureDetector_create(t, True).detect(np.ascontiguousarray(cell))
        ref = np.array([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in k], np.int32).reshape(-1, 3)
        out = np.zeros(4096, oracle.CORNER)
        n = oracle.lib().orc_fast9_16(cell.ctypes.data, cw, ch, img.strides[0], t, 1, out.ctypes.data, 4096)
        mine = corners_to_array(out[:n])
        assert mine.shape == ref.shape and (mine == ref).all()


def test_sincosf_vs_glibc(oracle):
    libm = ctypes.CDLL("libm.so.6")
    libm.sincosf.argtypes = [ctypes.c_float, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(0, 6.2832, 20000), np.arange(0, 360, 0.25) * np.pi / 180,
                         [0.0, 1e-5, 2.0 ** -12, 0.78539816, 0.7853982, 6.2831855]]).astype(np.float32)
    s, c = ctypes.c_float(), ctypes.c_float()
    for x in xs:
        libm.sincosf(float(x), ctypes.byref(s), ctypes.byref(c))
        ms, mc = oracle.sincosf(x)
        assert (ms, mc) == (s.value, c.value), x


def test_fast_atan2_vs_cv2_dense(oracle):
    """orc_fast_atan2 against the scalar cv2.fastAtan2 (the call of ORBextractor.cc:103) on integer moments."""
    cv2 = pytest.importorskip("cv2")
    lim = 255 * 4896
    rng = np.random.default_rng(3)
    small = np.arange(-30, 31, dtype=np.int32)
    m01 = np.concatenate([np.repeat(small, small.size), rng.integers(-lim, lim + 1, 200000, dtype=np.int32), [lim, -lim, lim, 0]])
    m10 = np.concatenate([np.tile(small, small.size), rng.integers(-lim, lim + 1, 200000, dtype=np.int32), [lim, lim, -lim, -lim]])
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(m01, m10)], np.float32)
    mine = oracle.orientation_sweep(m01, m10)
    assert (mine.view(np.uint32) == ref.view(np.uint32)).all()


def test_sincosf_restatement_exhaustive(oracle):
    """every angle the extractor can produce is a float in [0, 360] degrees; the restatement must equal
    this image's libm on all 1.1e9 of them (bit patterns 0 .. bits(360.0f))."""
    last = int(np.array([360.0], np.float32).view(np.uint32)[0])
    bad, chunk = 0, 1 << 27
    for first in range(0, last + 1, chunk):
        bad += oracle.steering_sweep(first, min(chunk, last + 1 - first), outputs=False)[2]
    assert bad == 0


def test_tables(oracle):
    e = oracle.Extractor(1000, 1.2, 8, 20, 7)
    assert e.quotas() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert e.umax() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert oracle.Extractor(2000, 1.2, 8, 20, 7).quotas() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert oracle.Extractor(5000, 1.2, 8, 20, 7).quotas() == [1086, 905, 754, 628, 524, 436, 364, 303]
    sf = e.scale_factors()
    assert abs(sf[7] - 3.5831816196) < 1e-6


@pytest.mark.parametrize("cfg,seed", [("euroc", 0), ("odd", 5), ("kitti12", 7)])
def test_extract_vs_golden(oracle, cfg, seed):
    h, w, nf, sf, nl, it, mt = CONFIGS[cfg]
    g = np.load(os.path.join(GOLD, "extract_%s_seed%d.npz" % (cfg, seed)))
    img = synth.frame(h, w, seed)
    assert sha(img) == str(g["image_sha"])
    e = oracle.Extractor(nf, sf, nl, it, mt)
    kps, desc = e(img)
    assert e.quotas() == list(g["quota"]) and e.umax() == list(g["umax"])
    for l in range(nl):
        assert sha(e.pyramid(l)) == str(g["pyramid_sha"][l]), "pyramid level %d" % l
        c = e.candidates(l)
        assert len(c) == g["cand_count"][l]
        assert sha(corners_to_array(c)) == str(g["cand_sha"][l]), "candidates level %d" % l
    gk = g["keypoints"]
    assert len(kps) == len(gk)
    for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
        assert (kps[f] == gk[f]).all(), f
    assert (desc == g["descriptors"]).all()


def test_extract_vs_cv2_restatement_live(oracle):
    pytest.importorskip("cv2")
    import cv2_restatement as R
    pat, _ = load_pattern()
    h, w, nf, sf, nl, it, mt = CONFIGS["odd"]
    img = synth.frame(h, w, 101)
    p = R.Params(nf, sf, nl, it, mt)
    levels, cand, kps, desc = R.extract(p, img, pat)
    e = oracle.Extractor(nf, sf, nl, it, mt)
    k2, d2 = e(img)
    assert len(kps) == len(k2)
    ka = np.array([(k["x"], k["y"], k["angle"], k["response"], k["octave"]) for k in kps], np.float32)
    kb = np.stack([k2["x"], k2["y"], k2["angle"], k2["response"], k2["octave"].astype(np.float32)], 1)
    assert (ka == kb).all() and (desc == d2).all()


def test_edge_cases(oracle):
    e = oracle.Extractor(1000, 1.2, 8, 20, 7)
    flat = np.full((480, 752), 77, np.uint8)
    k, d = e(flat)
    assert len(k) == 0 and d.shape == (0, 32)          # nkeypoints == 0 path, ORBextractor.cc:1064-1065
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, (240, 320)).astype(np.uint8)
    k, d = e(noise)
    assert len(k) > 0 and len(k) == len(d)
    # determinism
    k2, d2 = e(noise)
    assert (k == k2).all() and (d == d2).all()


def test_octree_list_order_model(oracle):
    """stand-alone octree: selected count never below min(N, #occupied cells) and indices are unique"""
    rng = np.random.default_rng(9)
    for n, N in ((50, 100), (500, 100), (3000, 217), (1, 10), (2, 1)):
        c = np.zeros(n, oracle.CORNER)
        c["x"] = rng.integers(0, 720, n)
        c["y"] = rng.integers(0, 448, n)
        c["score"] = rng.integers(7, 120, n)
        sel = oracle.distribute_octree(c, 16, 736, 16, 464, N)
        assert len(set(sel.tolist())) == len(sel)
        assert len(sel) <= N + 3 or n <= N


EUROC_K = (458.654, 457.296, 367.215, 248.375)                 # Examples/Monocular/EuRoC.yaml
EUROC_DIST = [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]
TUM1_K = (517.306408, 516.469215, 318.643040, 255.313989)     # Examples/Monocular/TUM1.yaml (with k3)
TUM1_DIST = [0.262383, -0.953104, -0.005358, 0.002628, 1.163314]


@pytest.mark.parametrize("K,dist,size", [(EUROC_K, EUROC_DIST, (752, 480)), (TUM1_K, TUM1_DIST, (640, 480)),
                                         (EUROC_K, [0.0, 0.1, 0.0, 0.0], (752, 480))])
def test_undistort_vs_cv2(oracle, K, dist, size):
    """Frame::UndistortKeyPoints / ComputeImageBounds restatement is bit-equal to cv2.undistortPoints (the call of
    Frame.cc:602 with R = I, P = K), incl. the no-distortion shortcut of :586"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    n = 4000
    kps = np.zeros(n, oracle.KEYPOINT)
    kps["x"] = rng.uniform(0, size[0], n).astype(np.float32)
    kps["y"] = rng.uniform(0, size[1], n).astype(np.float32)
    kps["octave"] = rng.integers(0, 8, n)
    got = oracle.undistort_keypoints(kps, *K, dist)
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32)
    D = np.array(dist, np.float32).reshape(-1, 1)
    if D[0, 0] != 0.0:
        pts = np.stack([kps["x"], kps["y"]], 1).reshape(-1, 1, 2)
        ref = cv2.undistortPoints(pts, Km, D, None, Km).reshape(-1, 2)
        corners = np.array([[0, 0], [size[0], 0], [0, size[1]], [size[0], size[1]]], np.float32).reshape(-1, 1, 2)
        c = cv2.undistortPoints(corners, Km, D, None, Km).reshape(-1, 2)
        bref = np.array([min(c[0, 0], c[2, 0]), max(c[1, 0], c[3, 0]), min(c[0, 1], c[1, 1]), max(c[2, 1], c[3, 1])], np.float32)
    else:
        ref = np.stack([kps["x"], kps["y"]], 1)
        bref = np.array([0, size[0], 0, size[1]], np.float32)
    assert (got["x"].view(np.uint32) == ref[:, 0].view(np.uint32)).all()
    assert (got["y"].view(np.uint32) == ref[:, 1].view(np.uint32)).all()
    assert (got["octave"] == kps["octave"]).all()
    b = oracle.compute_image_bounds(size[0], size[1], *K, dist)
    assert (b.view(np.uint32) == bref.view(np.uint32)).all()
