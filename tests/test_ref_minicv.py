"""Pins the stand-in OpenCV (oracle/refbuild/minicv) that the reference's sources are compiled against on the REAL OpenCV
that runs in this image (python cv2 4.13): every cv:: function the ORB path calls, bit for bit, on fresh random inputs.

These are the conventions the reference inherits from its un-vendored third-party dependency; the
expression forms of src/ORBmatcher.cc (`Rcw*x3Dw+tcw`, `-Rcw.t()*tcw`, `sRcw/scw` ...) are checked against cv2.gemm with the
flags and scale factors OpenCV's MatExpr produces for them (core/matop.cpp).
"""
import ctypes as C

import numpy as np
import pytest

from viorb_b200 import synth

cv2 = pytest.importorskip("cv2")
f32 = np.float32


@pytest.fixture(scope="module")
def R():
    from oracle import ref_py
    if not ref_py.available():
        pytest.skip("no reference library")
    L = ref_py.lib()._l
    vp, i32, sz, dbl = C.c_void_p, C.c_int, C.c_size_t, C.c_double
    L.ref_cv_resize.argtypes = [vp, i32, i32, sz, vp, i32, i32, sz]
    L.ref_cv_copy_make_border.argtypes = [vp, i32, i32, sz, i32, i32, i32, i32, i32, i32, vp, sz]
    L.ref_cv_copy_make_border_inplace.argtypes = [vp, i32, i32, sz, i32]
    L.ref_cv_gaussian7.argtypes = [vp, i32, i32, sz, vp, sz]
    L.ref_cv_fast.argtypes = [vp, i32, i32, sz, i32, i32, vp, i32]
    L.ref_cv_fast_atan2.argtypes = [C.c_float, C.c_float]
    L.ref_cv_fast_atan2.restype = C.c_float
    L.ref_cv_round.argtypes = [dbl]
    L.ref_cv_gemm.argtypes = [vp, i32, i32, vp, i32, i32, dbl, vp, i32, i32, dbl, i32, vp]
    L.ref_cv_norm.argtypes = [vp, i32, i32]
    L.ref_cv_norm.restype = dbl
    L.ref_cv_norm_diff.argtypes = [vp, vp, i32, i32, i32]
    L.ref_cv_norm_diff.restype = dbl
    L.ref_cv_dot.argtypes = [vp, vp, i32]
    L.ref_cv_dot.restype = dbl
    L.ref_cv_undistort.argtypes = [vp, i32, vp, vp, i32, vp]
    L.ref_cv_expr.argtypes = [i32, vp, vp, vp, C.c_float, vp]
    cv2.setNumThreads(1)
    return L


def bits(a):
    return np.ascontiguousarray(a, f32).view(np.uint32)


def gemm(R, A, B, alpha=1.0, Cm=None, beta=0.0, flags=0):
    A, B = np.ascontiguousarray(A, f32), np.ascontiguousarray(B, f32)
    m = A.shape[1] if flags & 1 else A.shape[0]
    n = B.shape[0] if flags & 2 else B.shape[1]
    D = np.zeros((m, n), f32)
    cc = np.ascontiguousarray(Cm, f32) if Cm is not None else None
    R.ref_cv_gemm(A.ctypes.data, A.shape[0], A.shape[1], B.ctypes.data, B.shape[0], B.shape[1], alpha,
                  cc.ctypes.data if cc is not None else None, cc.shape[0] if cc is not None else 0,
                  cc.shape[1] if cc is not None else 0, beta, flags, D.ctypes.data)
    return D


def test_resize_border_blur(R):
    for h, w, seed in ((480, 752, 1), (97, 131, 2), (376, 1241, 3)):
        img = synth.frame(h, w, seed)
        dh, dw = int(round(h / 1.2)), int(round(w / 1.2))
        d = np.zeros((dh, dw), np.uint8)
        R.ref_cv_resize(img.ctypes.data, h, w, img.strides[0], d.ctypes.data, dh, dw, d.strides[0])
        assert (d == cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)).all()
        b = np.zeros((h, w), np.uint8)
        R.ref_cv_gaussian7(img.ctypes.data, h, w, img.strides[0], b.ctypes.data, b.strides[0])
        assert (b == cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)).all()
        # a view inside a parent.  BORDER_ISOLATED: the view is the whole image (what cv2 sees for any numpy view).
        x0, y0, vw, vh = 9, 5, w - 40, h - 30
        out = np.zeros((vh + 38, vw + 38), np.uint8)
        R.ref_cv_copy_make_border(img.ctypes.data, h, w, img.strides[0], x0, y0, vw, vh, 19, 1, out.ctypes.data, out.strides[0])
        assert (out == cv2.copyMakeBorder(img[y0:y0 + vh, x0:x0 + vw], 19, 19, 19, 19, cv2.BORDER_REFLECT_101)).all()
        # without it OpenCV borrows the parent's pixels as far as they exist (copy.cpp: locateROI) and reflects the rest
        # about the grown source; python cannot hand cv2 a sub-matrix, so this is checked against that definition
        R.ref_cv_copy_make_border(img.ctypes.data, h, w, img.strides[0], x0, y0, vw, vh, 19, 0, out.ctypes.data, out.strides[0])
        grown = img[0:min(h, y0 + vh + 19), 0:min(w, x0 + vw + 19)]           # 5 rows / 9 columns exist above / left
        want = cv2.copyMakeBorder(grown, 19 - y0, 19 - (grown.shape[0] - y0 - vh), 19 - x0, 19 - (grown.shape[1] - x0 - vw),
                                  cv2.BORDER_REFLECT_101)
        assert (out == want).all()
        pad = np.zeros((h + 38, w + 38), np.uint8)
        pad[19:19 + h, 19:19 + w] = img
        R.ref_cv_copy_make_border_inplace(pad.ctypes.data, h, w, pad.strides[0], 19)
        assert (pad == cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)).all()


def test_fast_cells(R):
    img = synth.frame(240, 320, 4)
    rng = np.random.default_rng(0)
    for _ in range(80):
        cw, ch = int(rng.integers(7, 46)), int(rng.integers(7, 46))
        x0, y0 = int(rng.integers(0, 320 - cw)), int(rng.integers(0, 240 - ch))
        t = int(rng.choice([20, 12, 7]))
        cell = img[y0:y0 + ch, x0:x0 + cw]
        out = np.zeros((4096, 3), f32)
        n = R.ref_cv_fast(cell.ctypes.data, ch, cw, cell.strides[0], t, 1, out.ctypes.data, 4096)
        k = cv2.FastFeatureDetector_create(t, True).detect(np.ascontiguousarray(cell))
        want = np.array([(p.pt[0], p.pt[1], p.response) for p in k], f32).reshape(-1, 3)
        assert n == len(want) and (out[:n] == want).all()


def test_atan2_and_round(R):
    rng = np.random.default_rng(5)
    yx = rng.integers(-1248480, 1248481, size=(20000, 2)).astype(f32)
    got = np.array([R.ref_cv_fast_atan2(float(y), float(x)) for y, x in yx], f32)
    want = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], f32)
    assert (bits(got) == bits(want)).all()
    for v in (0.5, 1.5, 2.5, -0.5, -1.5, 3.4999, 1e6 + 0.5, -2.5):
        assert R.ref_cv_round(v) == int(np.rint(v))                  # cvRound = round half to even


def test_small_gemm_forms(R):
    """cv::gemm on the 3x3 / 3x1 CV_32F shapes of src/ORBmatcher.cc, all three code paths of core/matmul.cpp"""
    rng = np.random.default_rng(1)
    for _ in range(3000):
        A = rng.standard_normal((3, 3)).astype(f32)
        B3 = rng.standard_normal((3, 3)).astype(f32)
        x = (rng.standard_normal((3, 1)) * 10).astype(f32)
        t = rng.standard_normal((3, 1)).astype(f32)
        s = float(f32(rng.uniform(0.5, 2.0)))
        # R*x + t: ONE gemm (MatOp_GEMM::add), the hand-unrolled float block
        assert (bits(gemm(R, A, x, 1.0, t, 1.0)) == bits(cv2.gemm(A, x, 1.0, t, 1.0))).all()
        assert (bits(gemm(R, A, x)) == bits(cv2.gemm(A, x, 1.0, None, 0.0))).all()
        assert (bits(gemm(R, A, B3)) == bits(cv2.gemm(A, B3, 1.0, None, 0.0))).all()
        assert (bits(gemm(R, A, x, -1.0)) == bits(cv2.gemm(A, x, -1.0, None, 0.0))).all()
        # -R.t()*t: GEMM_1_T, alpha = -1 -> the generic path accumulating in double
        assert (bits(gemm(R, A, t, -1.0, None, 0.0, 1)) == bits(cv2.gemm(A, t, -1.0, None, 0.0, flags=cv2.GEMM_1_T))).all()
        assert (bits(gemm(R, A, t, s, x, 1.0, 1)) == bits(cv2.gemm(A, t, s, x, 1.0, flags=cv2.GEMM_1_T))).all()
        # the same forms through the stand-in MatExpr
        out = np.zeros((3, 1), f32)
        R.ref_cv_expr(0, A.ctypes.data, x.ctypes.data, t.ctypes.data, s, out.ctypes.data)
        assert (bits(out) == bits(cv2.gemm(A, x, 1.0, t, 1.0))).all()
        R.ref_cv_expr(1, A.ctypes.data, x.ctypes.data, t.ctypes.data, s, out.ctypes.data)
        assert (bits(out) == bits(cv2.gemm(A, t, -1.0, None, 0.0, flags=cv2.GEMM_1_T))).all()
        R.ref_cv_expr(5, A.ctypes.data, x.ctypes.data, t.ctypes.data, s, out.ctypes.data)
        assert (bits(out) == bits(cv2.gemm(A, t, -1.0, None, 0.0))).all()
        R.ref_cv_expr(6, A.ctypes.data, x.ctypes.data, t.ctypes.data, s, out.ctypes.data)
        assert (bits(out) == bits(cv2.subtract(x, t))).all()
        o9 = np.zeros((3, 3), f32)
        # M/s = convertTo(alpha = 1./s): float multiply by (float)(1./s)  (matop.cpp operator/ ; convert_scale)
        R.ref_cv_expr(2, A.ctypes.data, x.ctypes.data, t.ctypes.data, s, o9.ctypes.data)
        assert (bits(o9) == bits(cv2.multiply(A, np.full((3, 3), f32(1.0 / s), f32)))).all()
        R.ref_cv_expr(3, A.ctypes.data, x.ctypes.data, t.ctypes.data, s, o9.ctypes.data)
        assert (bits(o9) == bits(cv2.multiply(A, np.full((3, 3), f32(s), f32)))).all()
        R.ref_cv_expr(4, A.ctypes.data, x.ctypes.data, t.ctypes.data, s, o9.ctypes.data)
        assert (bits(o9) == bits(cv2.multiply(np.ascontiguousarray(A.T), np.full((3, 3), f32(1.0 / s), f32)))).all()


def test_norm_dot(R):
    rng = np.random.default_rng(2)
    for _ in range(5000):
        v = (rng.standard_normal(3) * 7).astype(f32)
        w = (rng.standard_normal(3) * 7).astype(f32)
        assert R.ref_cv_norm(v.ctypes.data, 3, 4) == cv2.norm(v.reshape(3, 1))
        got = R.ref_cv_dot(v.ctypes.data, w.ctypes.data, 3)
        want = 0.0
        for k in range(3):
            want += float(v[k]) * float(w[k])                      # dotProd_32f scalar tail: double products, in order
        assert got == want
        assert f32(got) == cv2.gemm(v.reshape(1, 3), w.reshape(1, 3), 1.0, None, 0.0, flags=cv2.GEMM_2_T)[0, 0]
    a = rng.integers(-255, 256, (11, 11)).astype(f32)
    b = rng.integers(-255, 256, (11, 11)).astype(f32)
    assert R.ref_cv_norm_diff(a.ctypes.data, b.ctypes.data, 11, 11, 2) == cv2.norm(a, b, cv2.NORM_L1)


def test_undistort_points(R):
    rng = np.random.default_rng(3)
    K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]], f32)        # Examples/Monocular/EuRoC.yaml
    for dist in ([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], [-0.2, 0.05, 0.001, -0.0005, 0.01]):
        d = np.array(dist, f32)
        pts = np.stack([rng.uniform(0, 752, 4000), rng.uniform(0, 480, 4000)], 1).astype(f32)
        out = np.zeros_like(pts)
        R.ref_cv_undistort(pts.ctypes.data, len(pts), K.ctypes.data, d.ctypes.data, len(d), out.ctypes.data)
        want = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, d, None, None, K).reshape(-1, 2)
        assert (bits(out) == bits(want)).all()
