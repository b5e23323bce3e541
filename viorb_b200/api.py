"""ctypes binding of libviorb_b200.so (include/viorb_gpu.h) for the tests and bench.py.

This is harness code, not the product: the product is the C-ABI library and the C++ shims in
viorb_b200/host/.  The classes below mirror the reference's operator surface (ORBextractor.h:45-111,
ORBmatcher.h:37-102) so the parity tests read like calls into ORB_SLAM2.  There is no fallback: if the
CUDA library is missing or no B200 is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

KEYPOINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
TOP2 = np.dtype([("d1", "<i4"), ("i1", "<i4"), ("d2", "<i4"), ("i2", "<i4")])

LIB_PATH = os.path.join(_build.LIBDIR, "libviorb_b200.so")
_lib = None
EXPORTED = ["viorb_last_error", "viorb_ctx_launch_count"]


class ViorbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("viorb error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Loads the in-tree CUDA library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ViorbError(-2, "%s not built: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
    pp = C.POINTER(C.c_void_p)
    pi = C.POINTER(C.c_int)
    sig = {
        "viorb_ctx_create": [i32, vp, pp],
        "viorb_ctx_destroy": [vp],
        "viorb_ctx_synchronize": [vp],
        "viorb_device_count": [],
        "viorb_host_alloc": [sz, pp],
        "viorb_host_free": [vp],
        "viorb_extractor_create": [vp, i32, f32, i32, i32, i32, pp],
        "viorb_extractor_destroy": [vp],
        "viorb_extractor_configure": [vp, i32, i32],
        "viorb_extractor_set_gaussian": [vp, i32],
        "viorb_extractor_set_describe_mode": [vp, i32],
        "viorb_extractor_pyramid_download_all": [vp, i32, vp, vp, i32],
        "viorb_extractor_set_copy_mode": [vp, i32],
        "viorb_extractor_tables": [vp, pi, vp, vp, vp, vp, vp],
        "viorb_extractor_profile": [vp, i32],
        "viorb_extractor_stage_ms": [vp, vp, pi],
        "viorb_extract": [vp, vp, i32, i32, sz, vp, vp, i32, pi],
        "viorb_extract_batch": [vp, vp, i32, i32, i32, sz, sz, vp, vp, i32, vp],
        "viorb_extract_batch_device": [vp, vp, i32, i32, i32, sz, sz, vp, vp, i32, vp],
        "viorb_extractor_check": [vp],
        "viorb_extractor_pyramid_info": [vp, i32, pi, pi],
        "viorb_extractor_pyramid_download": [vp, i32, i32, vp, sz],
        "viorb_extractor_pyramid_device": [vp, i32, i32, pp, C.POINTER(sz)],
        "viorb_extractor_resident": [vp, pi, pi],
        "viorb_extractor_debug_candidates": [vp, i32, i32, vp, i32, pi],
        "viorb_extractor_debug_selected": [vp, i32, i32, vp, i32, pi],
        "viorb_debug_steering": [vp, C.c_uint32, i64, vp, vp],
        "viorb_debug_orientation": [vp, vp, vp, i64, vp],
        "viorb_descriptor_distance": [vp, vp, vp, i32, vp],
        "viorb_hamming_top2": [vp, vp, i32, vp, i64, i64, vp],
        "viorb_hamming_top2_device": [vp, vp, i32, vp, i64, i64, vp],
        "viorb_top2_merge_device": [vp, vp, i32, i32, vp],
        "viorb_stereo_match": [vp, i32, vp, i32, vp, vp, i32, vp, vp, i32, f32, f32, vp, vp],
        "viorb_frame_index_create": [vp, vp, vp, vp, i32, f32, f32, f32, f32, vp, i32, pp],
        "viorb_frame_index_destroy": [vp],
        "viorb_frame_index_create_distorted": [vp, vp, vp, vp, i32, f32, f32, f32, f32, vp, i32, i32, i32, vp, i32, pp],
        "viorb_frame_index_create_device": [vp, vp, vp, vp, i32, f32, f32, f32, f32, vp, i32, i32, i32, vp, i32, pp],
        "viorb_frame_index_keys": [vp, vp, vp],
        "viorb_frame_index_grid": [vp, vp, vp],
        "viorb_undistort_keypoints": [vp, vp, i32, f32, f32, f32, f32, vp, i32, vp],
        "viorb_compute_image_bounds": [vp, i32, i32, f32, f32, f32, f32, vp, i32, vp],
        "viorb_frame_features_in_area": [vp, f32, f32, f32, i32, i32, vp, i32, pi],
        "viorb_search_by_projection_local": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, f32, f32, vp, pi],
        "viorb_search_by_projection_frame": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, f32, f32, i32, i32, i32, vp, pi],
        "viorb_distinctive_descriptors": [vp, vp, vp, i32, vp, vp],
        "viorb_search_window_top1": [vp, vp, vp, vp, vp, vp, vp, i32, f32, i32, vp, vp, vp],
        "viorb_search_by_sim3": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, vp, pi],
        "viorb_search_by_bow": [vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, f32, i32, vp, pi],
        "viorb_search_for_initialization": [vp, vp, vp, i32, vp, i32, f32, i32, vp, pi],
        "viorb_vocabulary_create": [vp, i32, i32, i32, i32, i32, vp, vp, vp, pp],
        "viorb_vocabulary_destroy": [vp],
        "viorb_vocabulary_info": [vp, pi, pi],
        "viorb_bow_transform": [vp, vp, i32, i32, vp, vp, pi, vp, vp, vp, pi, vp, vp],
        "viorb_search_for_triangulation": [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp,
                                           i32, vp, f32, f32, vp, vp, i32, i32, i32, vp, pi],
    }
    for name, args in sig.items():
        fn = getattr(L, name)          # AttributeError here == a symbol of include/viorb_gpu.h is missing
        fn.argtypes = args
        fn.restype = i32
        EXPORTED.append(name)
    L.viorb_last_error.restype = C.c_char_p
    L.viorb_last_error.argtypes = []
    L.viorb_ctx_launch_count.restype = i64
    L.viorb_ctx_launch_count.argtypes = [vp]
    _lib = L
    return L


def _ck(rc):
    if rc != 0:
        raise ViorbError(rc, lib().viorb_last_error().decode())


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):          # torch tensor (device or pinned host memory)
        return a.data_ptr()
    raise TypeError(type(a))


class Context:
    """One CUDA device + stream.  stream: raw cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""

    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        _ck(lib().viorb_ctx_create(device, stream, C.byref(h)))
        self.h = h

    def synchronize(self):
        _ck(lib().viorb_ctx_synchronize(self.h))

    def launch_count(self):
        return int(lib().viorb_ctx_launch_count(self.h))

    def debug_orientation(self, m01, m10):
        """IC_Angle's fastAtan2 (degrees) for arrays of integer patch moments."""
        m01, m10 = np.ascontiguousarray(m01, np.int32), np.ascontiguousarray(m10, np.int32)
        out = np.empty(len(m01), np.float32)
        _ck(lib().viorb_debug_orientation(self.h, _ptr(m01), _ptr(m10), len(m01), _ptr(out)))
        return out

    def debug_steering(self, first_bits, n):
        """(sin, cos) the descriptor kernel uses for n consecutive float bit patterns taken as angles in degrees."""
        s, c = np.empty(n, np.float32), np.empty(n, np.float32)
        _ck(lib().viorb_debug_steering(self.h, int(first_bits), int(n), _ptr(s), _ptr(c)))
        return s, c

    def close(self):
        if getattr(self, "h", None):
            lib().viorb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """numpy array backed by pinned host memory (cudaHostAlloc)"""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    _ck(lib().viorb_host_alloc(max(n, 1), C.byref(p)))
    buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.ctypes.data] = p.value
    return arr


_PINNED = {}


def pinned_free(arr):
    p = _PINNED.pop(arr.ctypes.data, None)
    if p:
        lib().viorb_host_free(p)


class ORBextractor:
    """Mirror of ORB_SLAM2::ORBextractor (include/ORBextractor.h:45-111)."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, ctx=None):
        self.ctx = ctx or Context()
        h = C.c_void_p()
        _ck(lib().viorb_extractor_create(self.ctx.h, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, C.byref(h)))
        self.h = h
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.cap = nfeatures + 8 * nlevels + 64          # octree overshoot
        n = C.c_int()
        self._scale = np.zeros(nlevels, np.float32)
        self._inv = np.zeros(nlevels, np.float32)
        self._s2 = np.zeros(nlevels, np.float32)
        self._is2 = np.zeros(nlevels, np.float32)
        self._quota = np.zeros(nlevels, np.int32)
        _ck(lib().viorb_extractor_tables(self.h, C.byref(n), _ptr(self._scale), _ptr(self._inv), _ptr(self._s2),
                                         _ptr(self._is2), _ptr(self._quota)))

    def close(self):
        if getattr(self, "h", None):
            lib().viorb_extractor_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # getters, ORBextractor.h:63-83
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactors(self):
        return self._scale.copy()

    def GetInverseScaleFactors(self):
        return self._inv.copy()

    def GetScaleSigmaSquares(self):
        return self._s2.copy()

    def GetInverseScaleSigmaSquares(self):
        return self._is2.copy()

    def features_per_level(self):
        return self._quota.copy()

    def configure(self, chunk_frames=0, cand_div=0):
        _ck(lib().viorb_extractor_configure(self.h, chunk_frames, cand_div))

    def set_copy_mode(self, mode):
        """0 = input and output copies of extract_batch on two streams (default), 1 = on one (viorb_extractor_set_copy_mode)"""
        _ck(lib().viorb_extractor_set_copy_mode(self.h, int(mode)))

    def set_gaussian(self, opencv_variant):
        """0 = OpenCV >= 3.4 taps (default), 1 = OpenCV 2.4 taps (viorb_extractor_set_gaussian)"""
        _ck(lib().viorb_extractor_set_gaussian(self.h, int(opencv_variant)))

    def set_describe_mode(self, mode):
        """0 = automatic, 1 = blur per keypoint (fused kernel), 2 = blur whole levels (viorb_extractor_set_describe_mode)"""
        _ck(lib().viorb_extractor_set_describe_mode(self.h, int(mode)))

    def __call__(self, image, mask=None):
        """operator()(image, mask, keypoints, descriptors): returns (keypoints[KEYPOINT], descriptors[N,32])."""
        if image is None or image.size == 0:
            return np.zeros(0, KEYPOINT), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2, "image.type() == CV_8UC1"
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        kps = np.zeros(self.cap, KEYPOINT)
        desc = np.zeros((self.cap, 32), np.uint8)
        n = C.c_int()
        _ck(lib().viorb_extract(self.h, _ptr(image), image.shape[0], image.shape[1], image.strides[0], _ptr(kps),
                                _ptr(desc), self.cap, C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images, kps=None, desc=None, counts=None):
        """images: uint8 [B, rows, cols] host array (pinned for speed).  Returns (kps[B,cap], desc[B,cap,32], counts[B])."""
        B, rows, cols = images.shape
        assert images.dtype == np.uint8 and images.strides[2] == 1
        if kps is None:
            kps = np.zeros((B, self.cap), KEYPOINT)
        if desc is None:
            desc = np.zeros((B, self.cap, 32), np.uint8)
        if counts is None:
            counts = np.zeros(B, np.int32)
        _ck(lib().viorb_extract_batch(self.h, _ptr(images), B, rows, cols, images.strides[1], images.strides[0],
                                      _ptr(kps), _ptr(desc), self.cap, _ptr(counts)))
        return kps, desc, counts

    def extract_batch_device(self, d_images, B, rows, cols, d_kps, d_desc, d_counts, step=None, frame_stride=None):
        """all pointers are device addresses (ints or torch CUDA tensors); asynchronous"""
        step = step or cols
        frame_stride = frame_stride or step * rows
        _ck(lib().viorb_extract_batch_device(self.h, _ptr(d_images), B, rows, cols, step, frame_stride, _ptr(d_kps),
                                             _ptr(d_desc), self.cap, _ptr(d_counts)))

    def check(self):
        _ck(lib().viorb_extractor_check(self.h))

    def profile(self, enable=True):
        _ck(lib().viorb_extractor_profile(self.h, int(enable)))

    def stage_ms(self):
        """{stage: ms summed over the passes since the last query}, passes"""
        ms = np.zeros(4, np.float32)
        n = C.c_int()
        _ck(lib().viorb_extractor_stage_ms(self.h, _ptr(ms), C.byref(n)))
        return dict(zip(("pyramid", "fast", "octree", "describe"), [float(v) for v in ms])), n.value

    def pyramid_all(self, frame=0):
        """all padded levels of one frame with a single device-to-host copy (viorb_extractor_pyramid_download_all)"""
        outs = []
        for l in range(self.nlevels):
            w, h = C.c_int(), C.c_int()
            _ck(lib().viorb_extractor_pyramid_info(self.h, l, C.byref(w), C.byref(h)))
            outs.append(np.zeros((h.value + 38, w.value + 38), np.uint8))
        ptrs = (C.c_void_p * self.nlevels)(*[o.ctypes.data for o in outs])
        steps = (C.c_size_t * self.nlevels)(*[o.strides[0] for o in outs])
        _ck(lib().viorb_extractor_pyramid_download_all(self.h, frame, ptrs, steps, self.nlevels))
        return outs

    # mvImagePyramid, ORBextractor.h:85
    def pyramid(self, level, frame=0):
        w, h = C.c_int(), C.c_int()
        _ck(lib().viorb_extractor_pyramid_info(self.h, level, C.byref(w), C.byref(h)))
        out = np.zeros((h.value + 38, w.value + 38), np.uint8)
        _ck(lib().viorb_extractor_pyramid_download(self.h, frame, level, _ptr(out), out.strides[0]))
        return out

    def debug_candidates(self, level, frame=0, cap=1 << 17):
        out = np.zeros((cap, 3), np.int32)
        n = C.c_int()
        _ck(lib().viorb_extractor_debug_candidates(self.h, frame, level, _ptr(out), cap, C.byref(n)))
        return out[:min(n.value, cap)]

    def debug_selected(self, level, frame=0, cap=1 << 16):
        out = np.zeros((cap, 3), np.int32)
        n = C.c_int()
        _ck(lib().viorb_extractor_debug_selected(self.h, frame, level, _ptr(out), cap, C.byref(n)))
        return out[:min(n.value, cap)]


class FrameIndex:
    """Device copy of the matching state of a Frame: mvKeysUn, mDescriptors, mvuRight, mGrid (Frame.h)."""

    def __init__(self, ctx, kps_un, desc, u_right, bounds, scale_factors):
        self.ctx = ctx
        self.kps = np.ascontiguousarray(kps_un, KEYPOINT)
        self.desc = np.ascontiguousarray(desc, np.uint8)
        self.u_right = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        sf = np.ascontiguousarray(scale_factors, np.float32)
        h = C.c_void_p()
        _ck(lib().viorb_frame_index_create(ctx.h, _ptr(self.kps), _ptr(self.desc), _ptr(self.u_right), len(self.kps),
                                           bounds[0], bounds[1], bounds[2], bounds[3], _ptr(sf), len(sf), C.byref(h)))
        self.h = h
        self.n = len(self.kps)

    @classmethod
    def from_distorted(cls, ctx, kps, desc, u_right, K, dist_coef, image_size, scale_factors):
        """Frame::UndistortKeyPoints + ComputeImageBounds + AssignFeaturesToGrid on the device (Frame.cc:584-645,410-425).
        K = (fx, fy, cx, cy); image_size = (cols, rows)."""
        self = cls.__new__(cls)
        self.ctx = ctx
        self.kps = np.ascontiguousarray(kps, KEYPOINT)
        self.desc = np.ascontiguousarray(desc, np.uint8)
        self.u_right = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        sf = np.ascontiguousarray(scale_factors, np.float32)
        dc = np.ascontiguousarray(dist_coef, np.float32).ravel()
        h = C.c_void_p()
        _ck(lib().viorb_frame_index_create_distorted(ctx.h, _ptr(self.kps), _ptr(self.desc), _ptr(self.u_right), len(self.kps),
                                                     K[0], K[1], K[2], K[3], _ptr(dc) if len(dc) else None, len(dc),
                                                     image_size[0], image_size[1], _ptr(sf), len(sf), C.byref(h)))
        self.h = h
        self.n = len(self.kps)
        return self

    @classmethod
    def from_device(cls, ctx, d_kps, d_desc, n, K, dist_coef, image_size, scale_factors, d_u_right=None):
        """the same from device-resident raw keypoints / descriptors (a frame's slice of extract_batch_device outputs)"""
        self = cls.__new__(cls)
        self.ctx = ctx
        sf = np.ascontiguousarray(scale_factors, np.float32)
        dc = np.ascontiguousarray(dist_coef, np.float32).ravel()
        h = C.c_void_p()
        _ck(lib().viorb_frame_index_create_device(ctx.h, _ptr(d_kps), _ptr(d_desc), _ptr(d_u_right), n, K[0], K[1], K[2], K[3],
                                                  _ptr(dc) if len(dc) else None, len(dc), image_size[0], image_size[1],
                                                  _ptr(sf), len(sf), C.byref(h)))
        self.h = h
        self.n = n
        return self

    def keys(self):
        """-> (mvKeysUn, (mnMinX, mnMaxX, mnMinY, mnMaxY))"""
        k = np.zeros(self.n, KEYPOINT)
        b = np.zeros(4, np.float32)
        _ck(lib().viorb_frame_index_keys(self.h, _ptr(k) if self.n else None, _ptr(b)))
        return k, b

    def grid(self):
        """Frame::mGrid as CSR: (cell_start[64*48+1], cell_items), cell id = ix*48 + iy"""
        cs = np.zeros(64 * 48 + 1, np.int32)
        ci = np.zeros(max(self.n, 1), np.int32)
        _ck(lib().viorb_frame_index_grid(self.h, _ptr(cs), _ptr(ci)))
        return cs, ci[:cs[-1]]

    def close(self):
        if getattr(self, "h", None):
            lib().viorb_frame_index_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def GetFeaturesInArea(self, x, y, r, minLevel=-1, maxLevel=-1):
        out = np.zeros(self.n + 1, np.int32)
        n = C.c_int()
        _ck(lib().viorb_frame_features_in_area(self.h, x, y, r, minLevel, maxLevel, _ptr(out), len(out), C.byref(n)))
        return out[:n.value]


class ORBmatcher:
    """Mirror of ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:37-102) on flattened inputs."""
    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30

    def __init__(self, nnratio=0.6, checkOri=True, ctx=None):
        self.mfNNratio, self.mbCheckOrientation = nnratio, checkOri
        self.ctx = ctx or Context()

    def DescriptorDistance(self, a, b):
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        out = np.zeros(len(a), np.int32)
        _ck(lib().viorb_descriptor_distance(self.ctx.h, _ptr(a), _ptr(b), len(a), _ptr(out)))
        return int(out[0]) if len(out) == 1 else out

    def hamming_top2(self, queries, dmap, index_base=0):
        q = np.ascontiguousarray(queries, np.uint8)
        m = np.ascontiguousarray(dmap, np.uint8)
        out = np.zeros(len(q), TOP2)
        _ck(lib().viorb_hamming_top2(self.ctx.h, _ptr(q), len(q), _ptr(m), len(m), index_base, _ptr(out)))
        return out

    def hamming_top2_device(self, d_q, Q, d_map, M, index_base, d_out):
        _ck(lib().viorb_hamming_top2_device(self.ctx.h, _ptr(d_q), Q, _ptr(d_map), M, index_base, _ptr(d_out)))

    def top2_merge_device(self, d_parts, nparts, Q, d_out):
        _ck(lib().viorb_top2_merge_device(self.ctx.h, _ptr(d_parts), nparts, Q, _ptr(d_out)))

    def SearchByProjectionLocal(self, fi, frame_mp_obs, proj_x, proj_y, proj_xr, pred_level, view_cos, valid, nobs,
                                mp_desc, th):
        f32, i32, u8 = np.float32, np.int32, np.uint8
        obs = np.ascontiguousarray(frame_mp_obs, i32).copy()
        a = [np.ascontiguousarray(x, t) for x, t in ((proj_x, f32), (proj_y, f32), (proj_xr, f32), (pred_level, i32),
                                                     (view_cos, f32), (valid, u8), (nobs, i32), (mp_desc, u8))]
        match = np.full(fi.n, -1, i32)
        n = C.c_int()
        _ck(lib().viorb_search_by_projection_local(fi.h, _ptr(obs), *[_ptr(x) for x in a], len(a[0]), th,
                                                   self.mfNNratio, _ptr(match), C.byref(n)))
        return n.value, match, obs

    def SearchByProjectionFrame(self, fi, frame_mp_obs, u, v, invz, last_octave, last_angle, valid, nobs, mp_desc, th,
                                mbf, mode, th_high=100):
        f32, i32, u8 = np.float32, np.int32, np.uint8
        obs = np.ascontiguousarray(frame_mp_obs, i32).copy()
        a = [np.ascontiguousarray(x, t) for x, t in ((u, f32), (v, f32), (invz, f32), (last_octave, i32),
                                                     (last_angle, f32), (valid, u8), (nobs, i32), (mp_desc, u8))]
        match = np.full(fi.n, -1, i32)
        n = C.c_int()
        _ck(lib().viorb_search_by_projection_frame(fi.h, _ptr(obs), *[_ptr(x) for x in a], len(a[0]), th, mbf, mode,
                                                   int(self.mbCheckOrientation), th_high, _ptr(match), C.byref(n)))
        match[match == -2] = -1      # -2 = "assigned, then removed by the rotation check": an empty slot for the flat comparison
        return n.value, match, obs

    def SearchForTriangulation(self, k1, d1, ur1, has_mp1, k2, d2, ur2, has_mp2, fv1, fv2, F12, ex, ey, scale2,
                               sigma2_2, bOnlyStereo=False):
        f32, i32, u8 = np.float32, np.int32, np.uint8
        k1, k2 = np.ascontiguousarray(k1, KEYPOINT), np.ascontiguousarray(k2, KEYPOINT)
        a = [np.ascontiguousarray(x, t) for x, t in ((d1, u8), (ur1, f32), (has_mp1, u8), (d2, u8), (ur2, f32),
                                                     (has_mp2, u8), (fv1[0], i32), (fv1[1], i32), (fv1[2], i32),
                                                     (fv2[0], i32), (fv2[1], i32), (fv2[2], i32), (F12, f32),
                                                     (scale2, f32), (sigma2_2, f32))]
        m12 = np.full(len(k1), -1, i32)
        n = C.c_int()
        _ck(lib().viorb_search_for_triangulation(
            self.ctx.h, _ptr(k1), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), len(k1), _ptr(k2), _ptr(a[3]), _ptr(a[4]),
            _ptr(a[5]), len(k2), _ptr(a[6]), _ptr(a[7]), _ptr(a[8]), len(a[6]), _ptr(a[9]), _ptr(a[10]), _ptr(a[11]),
            len(a[9]), _ptr(a[12]), ex, ey, _ptr(a[13]), _ptr(a[14]), len(a[13]), int(bOnlyStereo),
            int(self.mbCheckOrientation), _ptr(m12), C.byref(n)))
        return n.value, m12

    def SearchByBoW(self, mode, k1, d1, valid1, k2, d2, valid2, fv1, fv2):
        """mode 0: SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) -> (nmatches, index in pKF per frame keypoint);
        mode 1: SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) -> (nmatches, keypoint of pKF2 per keypoint of pKF1)"""
        i32, u8 = np.int32, np.uint8
        k1, k2 = np.ascontiguousarray(k1, KEYPOINT), np.ascontiguousarray(k2, KEYPOINT)
        d1, d2, v1 = np.ascontiguousarray(d1, u8), np.ascontiguousarray(d2, u8), np.ascontiguousarray(valid1, u8)
        v2 = np.ascontiguousarray(valid2, u8) if valid2 is not None else None
        a = [np.ascontiguousarray(x, i32) for x in (fv1[0], fv1[1], fv1[2], fv2[0], fv2[1], fv2[2])]
        match = np.full(len(k2) if mode == 0 else len(k1), -1, i32)
        n = C.c_int()
        _ck(lib().viorb_search_by_bow(self.ctx.h, mode, _ptr(k1), _ptr(d1), _ptr(v1), len(k1), _ptr(k2), _ptr(d2), _ptr(v2),
                                      len(k2), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), len(a[0]), _ptr(a[3]), _ptr(a[4]),
                                      _ptr(a[5]), len(a[3]), self.mfNNratio, int(self.mbCheckOrientation), _ptr(match),
                                      C.byref(n)))
        return n.value, match

    def SearchForInitialization(self, fi2, k1_un, d1, prev_matched, windowSize=10):
        """SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize); fi2 = FrameIndex of F2.
        -> (nmatches, vnMatches12, updated vbPrevMatched)"""
        k1 = np.ascontiguousarray(k1_un, KEYPOINT)
        d1 = np.ascontiguousarray(d1, np.uint8)
        prev = np.ascontiguousarray(prev_matched, np.float32).copy()
        m12 = np.full(len(k1), -1, np.int32)
        n = C.c_int()
        _ck(lib().viorb_search_for_initialization(fi2.h, _ptr(k1), _ptr(d1), len(k1), _ptr(prev), int(windowSize),
                                                  self.mfNNratio, int(self.mbCheckOrientation), _ptr(m12), C.byref(n)))
        return n.value, m12, prev

    def SearchWindowTop1(self, kf, u, v, ur, pred_level, valid, mp_desc, th, th_dist, inv_level_sigma2=None):
        """search loop of Fuse(KeyFrame*, vpMapPoints, th) (ur given) / Fuse(KeyFrame*, Scw, ...) (ur None):
        -> (bestIdx per map point or -1, bestDist)"""
        f32, i32, u8 = np.float32, np.int32, np.uint8
        a = [np.ascontiguousarray(x, t) for x, t in ((u, f32), (v, f32), (pred_level, i32), (valid, u8), (mp_desc, u8))]
        urr = np.ascontiguousarray(ur, f32) if ur is not None else None
        inv = np.ascontiguousarray(inv_level_sigma2, f32) if inv_level_sigma2 is not None else None
        n = len(a[0])
        bi, bd = np.full(n, -1, i32), np.zeros(n, i32)
        _ck(lib().viorb_search_window_top1(kf.h, _ptr(a[0]), _ptr(a[1]), _ptr(urr), _ptr(a[2]), _ptr(a[3]), _ptr(a[4]), n, th,
                                           th_dist, _ptr(inv), _ptr(bi), _ptr(bd)))
        return bi, bd

    def SearchBySim3(self, kf1, kf2, q12, q21, th):
        """q12 = (u, v, level, valid, mp_desc) of KF1's map points projected into KF2, q21 the reverse
        -> (nFound, idx2 per keypoint of KF1 or -1)"""
        f32, i32, u8 = np.float32, np.int32, np.uint8
        def prep(q):
            return [np.ascontiguousarray(q[0], f32), np.ascontiguousarray(q[1], f32), np.ascontiguousarray(q[2], i32),
                    np.ascontiguousarray(q[3], u8), np.ascontiguousarray(q[4], u8)]
        a, b = prep(q12), prep(q21)
        m = np.full(kf1.n, -1, i32)
        n = C.c_int()
        _ck(lib().viorb_search_by_sim3(kf1.h, kf2.h, *[_ptr(x) for x in a], *[_ptr(x) for x in b], th, _ptr(m), C.byref(n)))
        return n.value, m

    def ComputeDistinctiveDescriptors(self, obs_desc, obs_ptr):
        """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:249-314) over a CSR batch of map points:
        returns (BestIdx per point, BestMedian per point)."""
        d = np.ascontiguousarray(obs_desc, np.uint8).reshape(-1, 32)
        p = np.ascontiguousarray(obs_ptr, np.int32)
        best = np.zeros(len(p) - 1, np.int32)
        med = np.zeros(len(p) - 1, np.int32)
        _ck(lib().viorb_distinctive_descriptors(self.ctx.h, _ptr(d) if len(d) else None, _ptr(p), len(best),
                                                _ptr(best), _ptr(med)))
        return best, med


def UndistortKeyPoints(ctx, kps, K, dist_coef):
    """Frame::UndistortKeyPoints (Frame.cc:584-614); K = (fx, fy, cx, cy)"""
    k = np.ascontiguousarray(kps, KEYPOINT)
    dc = np.ascontiguousarray(dist_coef, np.float32).ravel()
    out = np.zeros(len(k), KEYPOINT)
    _ck(lib().viorb_undistort_keypoints(ctx.h, _ptr(k), len(k), K[0], K[1], K[2], K[3], _ptr(dc) if len(dc) else None, len(dc),
                                        _ptr(out)))
    return out


def ComputeImageBounds(ctx, cols, rows, K, dist_coef):
    """Frame::ComputeImageBounds (Frame.cc:616-645) -> (mnMinX, mnMaxX, mnMinY, mnMaxY)"""
    dc = np.ascontiguousarray(dist_coef, np.float32).ravel()
    b = np.zeros(4, np.float32)
    _ck(lib().viorb_compute_image_bounds(ctx.h, cols, rows, K[0], K[1], K[2], K[3], _ptr(dc) if len(dc) else None, len(dc),
                                         _ptr(b)))
    return b


class ORBVocabulary:
    """Mirror of ORB_SLAM2::ORBVocabulary (DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>) for transform():
    the tree as loadFromTextFile reads it -- parent[i], descriptor[i], weight[i] per node, node 0 = root."""
    TF_IDF, TF, IDF, BINARY = 0, 1, 2, 3
    L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT = range(6)

    def __init__(self, k, L, parent, node_desc, node_weight, weighting=0, scoring=0, ctx=None):
        self.ctx = ctx or Context()
        par = np.ascontiguousarray(parent, np.int32)
        d = np.ascontiguousarray(node_desc, np.uint8).reshape(-1, 32)
        w = np.ascontiguousarray(node_weight, np.float64)
        assert len(par) == len(d) == len(w)
        h = C.c_void_p()
        _ck(lib().viorb_vocabulary_create(self.ctx.h, k, L, weighting, scoring, len(par), _ptr(par), _ptr(d), _ptr(w),
                                          C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().viorb_vocabulary_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        a, b = C.c_int(), C.c_int()
        _ck(lib().viorb_vocabulary_info(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def transform(self, desc, levelsup=4):
        """-> (BowVector (ids, values), FeatureVector (node ids, ptr, feature idx), word per feature, node per feature)"""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        ids, vals = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64)
        fvn, fvp, fvi = np.zeros(max(n, 1), np.int32), np.zeros(n + 1, np.int32), np.zeros(max(n, 1), np.int32)
        wo, no = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
        nb, nf = C.c_int(), C.c_int()
        _ck(lib().viorb_bow_transform(self.h, _ptr(d) if n else None, n, levelsup, _ptr(ids), _ptr(vals), C.byref(nb),
                                      _ptr(fvn), _ptr(fvp), _ptr(fvi), C.byref(nf), _ptr(wo), _ptr(no)))
        nfv = nf.value
        return ((ids[:nb.value], vals[:nb.value]), (fvn[:nfv], fvp[:nfv + 1], fvi[:fvp[nfv] if nfv else 0]), wo[:n], no[:n])


def ComputeStereoMatches(ex_left, ex_right, kps_l, desc_l, kps_r, desc_r, mbf, mb, frame_l=0, frame_r=0):
    """Frame::ComputeStereoMatches (Frame.cc:646-820): returns (mvuRight, mvDepth)."""
    kl, kr = np.ascontiguousarray(kps_l, KEYPOINT), np.ascontiguousarray(kps_r, KEYPOINT)
    dl, dr = np.ascontiguousarray(desc_l, np.uint8), np.ascontiguousarray(desc_r, np.uint8)
    ur = np.zeros(len(kl), np.float32)
    depth = np.zeros(len(kl), np.float32)
    _ck(lib().viorb_stereo_match(ex_left.h, frame_l, ex_right.h, frame_r, _ptr(kl), _ptr(dl), len(kl), _ptr(kr),
                                 _ptr(dr), len(kr), mbf, mb, _ptr(ur), _ptr(depth)))
    return ur, depth
