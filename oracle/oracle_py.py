"""ctypes binding of the CPU oracle (oracle/orb_oracle.h).  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (viorb_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

KEYPOINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
CORNER = np.dtype([("x", "<i4"), ("y", "<i4"), ("score", "<i4")])
TOP2 = np.dtype([("d1", "<i4"), ("i1", "<i4"), ("d2", "<i4"), ("i2", "<i4")])
assert KEYPOINT.itemsize == 28


class _Image(C.Structure):
    _fields_ = [("data", C.c_void_p), ("w", C.c_int32), ("h", C.c_int32), ("step", C.c_size_t)]


_lib = None
_override = None          # set by using(): routes every wrapper of this module to another library (oracle/ref_py.py)


class using:
    """with using(other_lib): ...  -- the wrappers below call `other_lib` (same orc_* names and signatures)."""

    def __init__(self, other):
        self.other = other

    def __enter__(self):
        global _override
        self.prev, _override = _override, self.other
        return self.other

    def __exit__(self, *a):
        global _override
        _override = self.prev


def lib(path=None):
    global _lib
    if _override is not None and path is None:
        return _override
    if _lib is not None and path is None:
        return _lib
    if path is None:
        path = os.path.join(_HERE, "_build", "liborb_oracle.so")
        srcs = [os.path.join(_HERE, f) for f in ("orb_oracle.cpp", "match_oracle.cpp", "bow_oracle.cpp", "orb_oracle.h")]
        if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    L = C.CDLL(path)
    vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    L.orc_resize_linear_u8.argtypes = [vp, i32, i32, sz, vp, i32, i32, sz]
    L.orc_border_reflect101_u8.argtypes = [vp, i32, i32, sz, vp, sz, i32]
    L.orc_fast9_16.argtypes = [vp, i32, i32, sz, i32, i32, vp, i32]
    L.orc_gaussian7_u8.argtypes = [vp, i32, i32, sz, vp, sz]
    L.orc_gaussian7_u8_variant.argtypes = [vp, i32, i32, sz, vp, sz, i32]
    L.orc_fast_atan2.argtypes = [f32, f32]
    L.orc_fast_atan2.restype = f32
    L.orc_sincosf.argtypes = [f32, C.POINTER(f32), C.POINTER(f32)]
    L.orc_orientation_sweep.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
    L.orc_orientation_sweep.restype = None
    L.orc_steering_sweep.argtypes = [C.c_uint32, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.orc_steering_sweep.restype = C.c_int64
    L.orc_cv_round_f.argtypes = [f32]
    L.orc_extractor_create.argtypes = [i32, f32, i32, i32, i32]
    L.orc_extractor_create.restype = vp
    L.orc_extractor_destroy.argtypes = [vp]
    L.orc_extractor_set_gaussian.argtypes = [vp, i32]
    L.orc_extractor_set_gaussian.restype = None
    L.orc_extract.argtypes = [vp, vp, i32, i32, sz, vp, vp, i32]
    L.orc_extractor_levels.argtypes = [vp]
    L.orc_extractor_quota.argtypes = [vp, i32]
    L.orc_extractor_scale.argtypes = [vp, i32]
    L.orc_extractor_scale.restype = f32
    L.orc_extractor_umax.argtypes = [vp, i32]
    for name in ("orc_extractor_pyramid", "orc_extractor_blurred"):
        fn = getattr(L, name)
        fn.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(sz)]
        fn.restype = vp
    L.orc_extractor_candidates.argtypes = [vp, i32, vp, i32]
    L.orc_extractor_level_keypoints.argtypes = [vp, i32, vp, i32]
    L.orc_extractor_stage_seconds.argtypes = [vp, C.POINTER(C.c_double)]
    L.orc_extractor_retried_cells.argtypes = [vp]
    L.orc_distribute_octree.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, i32]
    L.orc_descriptor_distance.argtypes = [vp, vp]
    L.orc_descriptor_distance_popcnt.argtypes = [vp, vp]
    L.orc_hamming_top2.argtypes = [vp, i32, vp, C.c_int64, C.c_int64, vp, i32, i32]
    L.orc_hamming_top2.restype = None
    L.orc_top2_merge.argtypes = [vp, i32, i32, vp]
    L.orc_top2_merge.restype = None
    L.orc_stereo_match.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp, i32, vp, vp, f32, f32, vp, vp, vp, vp]
    L.orc_grid_create.argtypes = [vp, i32, f32, f32, f32, f32]
    L.orc_grid_create.restype = vp
    L.orc_grid_destroy.argtypes = [vp]
    L.orc_grid_features_in_area.argtypes = [vp, f32, f32, f32, i32, i32, vp, i32]
    L.orc_search_by_projection_local.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32,
                                                 f32, f32, vp]
    L.orc_search_by_projection_frame.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32,
                                                 f32, f32, i32, i32, i32, vp]
    L.orc_search_for_triangulation.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp,
                                               vp, i32, vp, f32, f32, vp, vp, i32, i32, vp]
    L.orc_distinctive_descriptor.argtypes = [vp, i32, C.POINTER(i32)]
    L.orc_search_by_bow.argtypes = [i32, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, f32, i32, vp]
    L.orc_search_for_initialization.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, i32, f32, i32, vp]
    L.orc_undistort_keypoints.argtypes = [vp, i32, f32, f32, f32, f32, vp, i32, vp]
    L.orc_undistort_keypoints.restype = None
    L.orc_compute_image_bounds.argtypes = [i32, i32, f32, f32, f32, f32, vp, i32, vp]
    L.orc_compute_image_bounds.restype = None
    L.orc_search_window_top1.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, f32, i32, vp, vp, vp]
    L.orc_search_window_top1.restype = None
    L.orc_search_by_sim3.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, vp]
    L.orc_vocabulary_create.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp]
    L.orc_vocabulary_create.restype = vp
    L.orc_vocabulary_destroy.argtypes = [vp]
    L.orc_bow_transform.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp, C.POINTER(i32), vp, vp]
    L.orc_num_threads.argtypes = []
    if path.endswith(os.path.join("_build", "liborb_oracle.so")):
        _lib = L
    return L


def _p(a):
    return a.ctypes.data if a is not None else None


def _supported(n, what):
    """adapters of oracle/refbuild return -2 / -3 for argument combinations that exist in no reference function
    (e.g. the frame-to-frame overload with a threshold other than TH_HIGH): nothing to compare against"""
    if n < 0 and _override is not None:
        import pytest
        pytest.skip("%s: no reference function takes this combination (adapter code %d)" % (what, n))
    return n


def _c(a, dtype=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


# ---------------------------------------------------------------- primitives
def resize_linear(src, dw, dh):
    src = _c(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def border_reflect101(src, border=19):
    src = _c(src, np.uint8)
    h, w = src.shape
    dst = np.empty((h + 2 * border, w + 2 * border), np.uint8)
    lib().orc_border_reflect101_u8(_p(src), w, h, src.strides[0], _p(dst), dst.strides[0], border)
    return dst


def fast9(img, threshold, nms=True, cap=1 << 16):
    img = _c(img, np.uint8)
    out = np.zeros(cap, CORNER)
    n = lib().orc_fast9_16(_p(img), img.shape[1], img.shape[0], img.strides[0], threshold, int(nms), _p(out), cap)
    assert n <= cap
    return out[:n]


def gaussian7(src, variant=0):
    """variant 0 = OpenCV >= 3.4 taps (cv2-pinned), 1 = OpenCV 2.4 taps"""
    src = _c(src, np.uint8)
    dst = np.empty_like(src)
    lib().orc_gaussian7_u8_variant(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0], int(variant))
    return dst


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


def sincosf(x):
    s, c = C.c_float(), C.c_float()
    lib().orc_sincosf(float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def orientation_sweep(m01, m10, threads=None):
    import os
    m01 = np.ascontiguousarray(m01, np.int32)
    m10 = np.ascontiguousarray(m10, np.int32)
    out = np.empty(len(m01), np.float32)
    lib().orc_orientation_sweep(m01.ctypes.data, m10.ctypes.data, len(m01), out.ctypes.data, threads or max(1, os.cpu_count() or 1))
    return out


def steering_sweep(first_bits, n, which=0, outputs=True, threads=None):
    """(sin, cos, mismatches) of the steering pair for n consecutive float bit patterns as angles in degrees;
    which=0 -> this image's libm, which=1 -> the restatement; mismatches counts restatement != libm."""
    import os
    threads = threads or max(1, os.cpu_count() or 1)
    s = np.empty(n, np.float32) if outputs else None
    c = np.empty(n, np.float32) if outputs else None
    bad = lib().orc_steering_sweep(int(first_bits), int(n), int(which), s.ctypes.data if outputs else None,
                                   c.ctypes.data if outputs else None, threads)
    return s, c, int(bad)


# ---------------------------------------------------------------- extractor
class Extractor:
    """Mirror of ORB_SLAM2::ORBextractor (ORBextractor.h:45-111) on the oracle."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, _lib=None):
        self.L = _lib or lib()
        self.h = self.L.orc_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nfeatures, self.nlevels = nfeatures, nlevels

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_extractor_destroy(self.h)
            self.h = None

    def __call__(self, img):
        img = _c(img, np.uint8)
        cap = self.nfeatures * 2 + 64
        kps = np.zeros(cap, KEYPOINT)
        desc = np.zeros((cap, 32), np.uint8)
        n = self.L.orc_extract(self.h, _p(img), img.shape[0], img.shape[1], img.strides[0], _p(kps), _p(desc), cap)
        assert 0 <= n <= cap, n
        return kps[:n].copy(), desc[:n].copy()

    def set_gaussian_variant(self, variant):
        """0 = OpenCV >= 3.4 taps (default), 1 = OpenCV 2.4 taps"""
        self.L.orc_extractor_set_gaussian(self.h, int(variant))

    def quotas(self):
        return [self.L.orc_extractor_quota(self.h, l) for l in range(self.nlevels)]

    def scale_factors(self):
        return np.array([self.L.orc_extractor_scale(self.h, l) for l in range(self.nlevels)], np.float32)

    def umax(self):
        return [self.L.orc_extractor_umax(self.h, v) for v in range(16)]

    def _image(self, fn, level):
        w, h, step = C.c_int(), C.c_int(), C.c_size_t()
        ptr = fn(self.h, level, C.byref(w), C.byref(h), C.byref(step))
        if not ptr:
            return None
        ph, pw = h.value + 38, w.value + 38
        buf = (C.c_uint8 * (ph * step.value)).from_address(ptr)
        return np.frombuffer(buf, np.uint8).reshape(ph, step.value)[:, :pw].copy()

    def pyramid(self, level):
        """padded level image (border 19) of the last call"""
        return self._image(self.L.orc_extractor_pyramid, level)

    def blurred(self, level):
        return self._image(self.L.orc_extractor_blurred, level)

    def candidates(self, level):
        n = self.L.orc_extractor_candidates(self.h, level, None, 0)
        out = np.zeros(max(n, 1), CORNER)
        self.L.orc_extractor_candidates(self.h, level, _p(out), n)
        return out[:n]

    def level_keypoints(self, level):
        n = self.L.orc_extractor_level_keypoints(self.h, level, None, 0)
        out = np.zeros(max(n, 1), KEYPOINT)
        self.L.orc_extractor_level_keypoints(self.h, level, _p(out), n)
        return out[:n]

    def stage_seconds(self):
        a = (C.c_double * 6)()
        self.L.orc_extractor_stage_seconds(self.h, a)
        return dict(zip(("pyramid", "fast", "octree", "orient", "blur", "desc"), list(a)))

    def retried_cells(self):
        return self.L.orc_extractor_retried_cells(self.h)


def distribute_octree(cand, min_x, max_x, min_y, max_y, n_target):
    cand = _c(cand, CORNER)
    out = np.zeros(n_target + 64, np.int32)
    n = lib().orc_distribute_octree(_p(cand), len(cand), min_x, max_x, min_y, max_y, n_target, _p(out), len(out))
    return out[:n]


# ---------------------------------------------------------------- matcher
def descriptor_distance(a, b, popcnt=False):
    a, b = _c(a, np.uint8), _c(b, np.uint8)
    f = lib().orc_descriptor_distance_popcnt if popcnt else lib().orc_descriptor_distance
    return f(_p(a), _p(b))


def hamming_top2(q, dmap, index_base=0, popcnt=True, nthreads=1, _lib=None):
    q, dmap = _c(q, np.uint8), _c(dmap, np.uint8)
    out = np.zeros(len(q), TOP2)
    (_lib or lib()).orc_hamming_top2(_p(q), len(q), _p(dmap), len(dmap), index_base, _p(out), int(popcnt), nthreads)
    return out


def top2_merge(parts):
    parts = _c(parts, TOP2)
    nparts, q = parts.shape
    out = np.zeros(q, TOP2)
    lib().orc_top2_merge(_p(parts), nparts, q, _p(out))
    return out


def _images(levels):
    """levels: list of padded numpy images (border 19) -> ctypes array of ROI views"""
    arr = (_Image * len(levels))()
    for i, im in enumerate(levels):
        arr[i].data = im.ctypes.data + 19 * im.strides[0] + 19
        arr[i].w = im.shape[1] - 38
        arr[i].h = im.shape[0] - 38
        arr[i].step = im.strides[0]
    return arr


def stereo_match(kl, dl, kr, dr, pyr_l, pyr_r, scale_factors, mbf, mb):
    kl, kr = _c(kl, KEYPOINT), _c(kr, KEYPOINT)
    dl, dr = _c(dl, np.uint8), _c(dr, np.uint8)
    sf = _c(scale_factors, np.float32)
    isf = (np.float32(1.0) / sf).astype(np.float32)
    pl = [np.ascontiguousarray(p) for p in pyr_l]
    pr = [np.ascontiguousarray(p) for p in pyr_r]
    il, ir = _images(pl), _images(pr)
    ur = np.zeros(len(kl), np.float32)
    depth = np.zeros(len(kl), np.float32)
    bd = np.zeros(len(kl), np.int32)
    bi = np.zeros(len(kl), np.int32)
    n = lib().orc_stereo_match(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), C.addressof(il), C.addressof(ir),
                               len(pl), _p(sf), _p(isf), mbf, mb, _p(ur), _p(depth), _p(bd), _p(bi))
    return ur, depth, bd, bi, n


class Grid:
    def __init__(self, kps_un, minx, maxx, miny, maxy):
        self.kps = _c(kps_un, KEYPOINT)
        self.L = lib()                      # the handle belongs to the library that made it (see using())
        self.g = self.L.orc_grid_create(_p(self.kps), len(self.kps), minx, maxx, miny, maxy)

    def __del__(self):
        if getattr(self, "g", None):
            self.L.orc_grid_destroy(self.g)
            self.g = None

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        out = np.zeros(len(self.kps) + 1, np.int32)
        n = self.L.orc_grid_features_in_area(self.g, x, y, r, min_level, max_level, _p(out), len(out))
        return out[:n]


def search_by_projection_local(grid, fdesc, fu_right, frame_mp_obs, scale_factors, proj_x, proj_y, proj_xr,
                               pred_level, view_cos, valid, mp_nobs, mpdesc, th, nnratio):
    f32, i32, u8 = np.float32, np.int32, np.uint8
    fdesc, mpdesc = _c(fdesc, u8), _c(mpdesc, u8)
    obs = _c(frame_mp_obs, i32).copy()
    match = np.full(len(grid.kps), -1, i32)
    args = [_c(a, t) for a, t in ((fu_right, f32), (scale_factors, f32), (proj_x, f32), (proj_y, f32), (proj_xr, f32),
                                  (pred_level, i32), (view_cos, f32), (valid, u8), (mp_nobs, i32))]
    n = grid.L.orc_search_by_projection_local(grid.g, _p(grid.kps), _p(fdesc), _p(args[0]), _p(obs), len(grid.kps),
                                             _p(args[1]), _p(args[2]), _p(args[3]), _p(args[4]), _p(args[5]),
                                             _p(args[6]), _p(args[7]), _p(args[8]), _p(mpdesc), len(mpdesc),
                                             th, nnratio, _p(match))
    return n, match, obs


def search_by_projection_frame(grid, fdesc, fu_right, frame_mp_obs, scale_factors, u, v, invz, last_octave,
                               last_angle, valid, mp_nobs, mpdesc, th, mbf, mode, check_ori=True, th_high=100):
    f32, i32, u8 = np.float32, np.int32, np.uint8
    fdesc, mpdesc = _c(fdesc, u8), _c(mpdesc, u8)
    obs = _c(frame_mp_obs, i32).copy()
    match = np.full(len(grid.kps), -1, i32)
    a = [_c(x, t) for x, t in ((fu_right, f32), (scale_factors, f32), (u, f32), (v, f32), (invz, f32),
                               (last_octave, i32), (last_angle, f32), (valid, u8), (mp_nobs, i32))]
    n = grid.L.orc_search_by_projection_frame(grid.g, _p(grid.kps), _p(fdesc), _p(a[0]), _p(obs), len(grid.kps),
                                             _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]), _p(a[5]), _p(a[6]), _p(a[7]),
                                             _p(a[8]), _p(mpdesc), len(mpdesc), th, mbf, mode, int(check_ori),
                                             th_high, _p(match))
    return _supported(n, "search_by_projection_frame"), match, obs


def search_for_triangulation(k1, d1, ur1, has_mp1, k2, d2, ur2, has_mp2, fv1, fv2, F12, ex, ey, scale2, sigma2_2,
                             only_stereo=False, check_ori=False):
    """fv1/fv2 = (node_ids, node_ptr, idx) CSR feature vectors"""
    f32, i32, u8 = np.float32, np.int32, np.uint8
    k1, k2 = _c(k1, KEYPOINT), _c(k2, KEYPOINT)
    d1, d2 = _c(d1, u8), _c(d2, u8)
    a = [_c(x, t) for x, t in ((ur1, f32), (has_mp1, u8), (ur2, f32), (has_mp2, u8), (fv1[0], i32), (fv1[1], i32),
                               (fv1[2], i32), (fv2[0], i32), (fv2[1], i32), (fv2[2], i32), (F12, f32), (scale2, f32),
                               (sigma2_2, f32))]
    m12 = np.full(len(k1), -1, i32)
    n = lib().orc_search_for_triangulation(_p(k1), _p(d1), _p(a[0]), _p(a[1]), len(k1), _p(k2), _p(d2), _p(a[2]),
                                           _p(a[3]), len(k2), _p(a[4]), _p(a[5]), _p(a[6]), len(a[4]), _p(a[7]),
                                           _p(a[8]), _p(a[9]), len(a[7]), _p(a[10]), ex, ey, _p(a[11]), _p(a[12]),
                                           int(only_stereo), int(check_ori), _p(m12))
    return n, m12


def distinctive_descriptors(obs_desc, obs_ptr):
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:249-314) per point of a CSR batch -> (best, median)."""
    obs_desc = _c(obs_desc, np.uint8).reshape(-1, 32)
    obs_ptr = np.asarray(obs_ptr, np.int64)
    best = np.zeros(len(obs_ptr) - 1, np.int32)
    med = np.zeros(len(obs_ptr) - 1, np.int32)
    for p in range(len(best)):
        rows = _c(obs_desc[obs_ptr[p]:obs_ptr[p + 1]])
        m = C.c_int()
        best[p] = lib().orc_distinctive_descriptor(_p(rows) if len(rows) else None, len(rows), C.byref(m))
        med[p] = m.value
    return best, med


class Vocabulary:
    """DBoW2 TemplatedVocabulary<FORB> restatement (bow_oracle.cpp): transform() only."""

    def __init__(self, k, L, parent, node_desc, node_weight, weighting=0, scoring=0):
        par, d, w = _c(parent, np.int32), _c(node_desc, np.uint8), _c(node_weight, np.float64)
        self.L = lib()
        self.h = self.L.orc_vocabulary_create(k, L, weighting, scoring, len(par), _p(par), _p(d), _p(w))

    def __del__(self):
        try:
            self.L.orc_vocabulary_destroy(self.h)
        except Exception:
            pass

    def transform(self, desc, levelsup=4):
        d = _c(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        ids, vals = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64)
        fvn, fvp, fvi = np.zeros(max(n, 1), np.int32), np.zeros(n + 1, np.int32), np.zeros(max(n, 1), np.int32)
        wo, no = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
        nf = C.c_int()
        nb = self.L.orc_bow_transform(self.h, _p(d), n, levelsup, _p(ids), _p(vals), _p(fvn), _p(fvp), _p(fvi), C.byref(nf),
                                     _p(wo), _p(no))
        nfv = nf.value
        return ((ids[:nb], vals[:nb]), (fvn[:nfv], fvp[:nfv + 1], fvi[:fvp[nfv] if nfv else 0]), wo[:n], no[:n])


def search_by_bow(mode, k1, d1, valid1, k2, d2, valid2, fv1, fv2, nnratio, check_ori=True):
    """ORBmatcher::SearchByBoW: mode 0 = (KeyFrame, Frame) -> match per frame keypoint; 1 = (KF1, KF2) -> per KF1 keypoint"""
    i32, u8 = np.int32, np.uint8
    k1, k2 = _c(k1, KEYPOINT), _c(k2, KEYPOINT)
    d1, d2, v1 = _c(d1, u8), _c(d2, u8), _c(valid1, u8)
    v2 = _c(valid2, u8) if valid2 is not None else None
    a = [_c(x, i32) for x in (fv1[0], fv1[1], fv1[2], fv2[0], fv2[1], fv2[2])]
    match = np.full(len(k2) if mode == 0 else len(k1), -1, i32)
    n = lib().orc_search_by_bow(mode, _p(k1), _p(d1), _p(v1), len(k1), _p(k2), _p(d2), _p(v2), len(k2), _p(a[0]), _p(a[1]),
                                _p(a[2]), len(a[0]), _p(a[3]), _p(a[4]), _p(a[5]), len(a[3]), nnratio, int(check_ori),
                                _p(match))
    return n, match


def search_for_initialization(grid2, d2, k1, d1, prev, window, nnratio, check_ori=True):
    i32, u8 = np.int32, np.uint8
    k1, d1, d2 = _c(k1, KEYPOINT), _c(d1, u8), _c(d2, u8)
    prev = np.ascontiguousarray(prev, np.float32).copy()
    m12 = np.full(len(k1), -1, i32)
    n = grid2.L.orc_search_for_initialization(grid2.g, _p(grid2.kps), _p(d2), len(grid2.kps), _p(k1), _p(d1), len(k1), _p(prev),
                                            int(window), nnratio, int(check_ori), _p(m12))
    return n, m12, prev


def undistort_keypoints(kps, fx, fy, cx, cy, dist):
    """Frame::UndistortKeyPoints (Frame.cc:584-614)"""
    kps = _c(kps, KEYPOINT)
    dist = _c(dist, np.float32).ravel()
    out = np.zeros(len(kps), KEYPOINT)
    lib().orc_undistort_keypoints(_p(kps), len(kps), fx, fy, cx, cy, _p(dist) if len(dist) else None, len(dist), _p(out))
    return out


def compute_image_bounds(cols, rows, fx, fy, cx, cy, dist):
    """Frame::ComputeImageBounds (Frame.cc:616-645) -> (mnMinX, mnMaxX, mnMinY, mnMaxY)"""
    dist = _c(dist, np.float32).ravel()
    b = np.zeros(4, np.float32)
    lib().orc_compute_image_bounds(cols, rows, fx, fy, cx, cy, _p(dist) if len(dist) else None, len(dist), _p(b))
    return b


def search_window_top1(grid, kdesc, u_right, scale_factors, u, v, ur, pred_level, valid, mp_desc, th, th_dist,
                       inv_level_sigma2=None):
    """search loop of Fuse x2 / SearchBySim3 (ORBmatcher.cc:883-943, :1043-1073, :1193-1226) -> (best_idx, best_dist)"""
    f32, i32, u8 = np.float32, np.int32, np.uint8
    n = len(u)
    a = [_c(x, t) for x, t in ((kdesc, u8), (u_right, f32), (scale_factors, f32), (u, f32), (v, f32), (pred_level, i32),
                               (valid, u8), (mp_desc, u8))]
    urr = _c(ur, f32) if ur is not None else None
    inv = _c(inv_level_sigma2, f32) if inv_level_sigma2 is not None else None
    bi, bd = np.full(n, -1, i32), np.zeros(n, i32)
    grid.L.orc_search_window_top1(grid.g, _p(grid.kps), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]), _p(urr), _p(a[5]),
                                 _p(a[6]), _p(a[7]), n, th, th_dist, _p(inv), _p(bi), _p(bd))
    return bi, bd


def search_by_sim3(g1, d1, sf1, g2, d2, sf2, q12, q21, th):
    """q12 = (u, v, level, valid, mp_desc) of KF1's map points projected into KF2; q21 the reverse"""
    f32, i32, u8 = np.float32, np.int32, np.uint8
    def prep(q):
        return [_c(q[0], f32), _c(q[1], f32), _c(q[2], i32), _c(q[3], u8), _c(q[4], u8)]
    a, b = prep(q12), prep(q21)
    d1, d2, sf1, sf2 = _c(d1, u8), _c(d2, u8), _c(sf1, f32), _c(sf2, f32)
    m = np.full(len(g1.kps), -1, i32)
    n = g1.L.orc_search_by_sim3(g1.g, _p(g1.kps), _p(d1), _p(sf1), len(g1.kps), g2.g, _p(g2.kps), _p(d2), _p(sf2), len(g2.kps),
                                 _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]), _p(b[0]), _p(b[1]), _p(b[2]), _p(b[3]),
                                 _p(b[4]), th, _p(m))
    return n, m
