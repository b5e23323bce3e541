"""ctypes binding of oracle/_ref/libviorb_ref.so: the REFERENCE's own source files (src/ORBextractor.cc, src/ORBmatcher.cc,
the ORB members of src/Frame.cc / KeyFrame.cc / MapPoint.cc, Thirdparty/DBoW2) compiled unmodified by oracle/refbuild/Makefile.
TEST INFRASTRUCTURE ONLY -- same import rules as oracle_py.

The library exports ref_X for the orc_X entry points of oracle/orb_oracle.h with identical signatures, so

    with O.using(ref_py.lib()):
        kps, desc = O.Extractor(...)(img)          # runs the reference

drives every wrapper of oracle_py through the reference instead of the restatement.

/root/reference exists only in the build container.  There the library is (re)built on demand; on the GPU box the prebuilt
file that travelled with the snapshot is loaded as is; with neither, available() is False and callers skip.
"""
import ctypes as C
import os
import subprocess

from . import oracle_py as O

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("VIORB_REFERENCE", "/root/reference")
PATH = os.path.join(_HERE, "_ref", "libviorb_ref.so")


_STATELESS = {"orc_hamming_top2", "orc_top2_merge", "orc_orientation_sweep", "orc_steering_sweep", "orc_sincosf", "orc_cv_round_f",
              "orc_num_threads", "orc_resize_linear_u8", "orc_border_reflect101_u8", "orc_fast9_16", "orc_gaussian7_u8",
              "orc_gaussian7_u8_variant", "orc_fast_atan2"}


class RefLib:
    """Attribute orc_X resolves to the library's ref_X, typed like the oracle's orc_X."""

    def __init__(self, path, fallback=False):
        self._l = C.CDLL(path) if isinstance(path, str) else path
        self._o = O.lib(os.path.join(_HERE, "_build", "liborb_oracle.so")) if O._lib is None else O._lib
        self._fallback = fallback

    def __getattr__(self, name):
        if name.startswith("orc_"):
            o = getattr(self._o, name)
            try:
                f = getattr(self._l, "ref_" + name[4:])
            except AttributeError:
                if not self._fallback:
                    raise
                if name in _STATELESS:
                    # no counterpart in the reference (the brute-force top-2 of BASELINE configs[4] is not a reference
                    # function; sweeps of single primitives): the restatement answers
                    f = o
                else:
                    # handle-based entry points must never mix libraries: the calling test is skipped for this checker
                    def f(*a, _n=name, **k):
                        import pytest
                        pytest.skip("the reference has no counterpart of %s" % _n)
            else:
                f.argtypes, f.restype = o.argtypes, o.restype
            setattr(self, name, f)
            return f
        return getattr(self._l, name)


_lib = None
_lib_fallback = None


def build():
    """make -C oracle/refbuild; only possible where the reference sources are."""
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "refbuild"), "REF=" + REF_ROOT], stdout=subprocess.DEVNULL)
    return PATH


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "src")) or os.path.exists(PATH)


def lib(fallback=False):
    """fallback=True: orc_X names the reference does not implement resolve to the restatement instead of raising"""
    global _lib, _lib_fallback
    if fallback:
        if _lib_fallback is None:
            _lib_fallback = RefLib(lib()._l, fallback=True)
        return _lib_fallback
    if _lib is None:
        O.lib()
        if os.path.isdir(os.path.join(REF_ROOT, "src")):
            build()
        if not os.path.exists(PATH):
            raise RuntimeError("oracle/_ref/libviorb_ref.so is missing and %s is not here to build it from" % REF_ROOT)
        _lib = RefLib(PATH)
        _lib._l.ref_arena_overflows.restype = C.c_long
    return _lib


def set_allocator(mode):
    """0 = process allocator, 1 = ascending addresses (the oracle's convention), 2 = descending addresses"""
    lib()._l.ref_set_allocator(int(mode))


def set_gaussian_variant(variant):
    """0 = OpenCV >= 3.4 taps, 1 = OpenCV 2.4 taps (the version the reference pins)"""
    lib()._l.ref_set_gaussian_variant(int(variant))


def arena_overflows():
    return int(lib()._l.ref_arena_overflows())


class PlainExtractor:
    """The reference's ORBextractor::operator() (src/ORBextractor.cc:1043-1105) and nothing else, on the process allocator,
    into preallocated buffers: what bench.py's CPU arm times.  One instance per thread."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        l = lib()._l
        l.ref_extractor_create.restype = C.c_void_p
        l.ref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        l.ref_extractor_destroy.argtypes = [C.c_void_p]
        l.ref_extract_plain.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
        l.ref_extract_plain.restype = C.c_int
        self._l = l
        self.h = l.ref_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.cap = nfeatures * 2 + 64
        import numpy as np
        self.kps = np.zeros(self.cap, O.KEYPOINT)
        self.desc = np.zeros((self.cap, 32), np.uint8)

    def __call__(self, img):
        assert img.dtype.name == "uint8" and img.strides[1] == 1
        n = self._l.ref_extract_plain(self.h, img.ctypes.data, img.shape[0], img.shape[1], img.strides[0],
                                      self.kps.ctypes.data, self.desc.ctypes.data, self.cap)
        return self.kps[:n], self.desc[:n]

    def close(self):
        if self.h:
            self._l.ref_extractor_destroy(self.h)
            self.h = None
