"""Synthetic inputs (frames, stereo pairs, descriptor maps) -- see csrc/synth.cpp."""
import ctypes
import os

import numpy as np

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        path = _build.build_synth()          # rebuilt only when synth.cpp is newer than the .so
        _lib = ctypes.CDLL(path)
        _lib.viorb_synth_frame.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p]
        _lib.viorb_synth_stereo.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.viorb_synth_frames.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p,
                                            ctypes.c_int]
        _lib.viorb_synth_bytes.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t]
    return _lib


def frame(h, w, seed, out=None):
    if out is None:
        out = np.empty((h, w), np.uint8)
    assert out.flags["C_CONTIGUOUS"] and out.shape == (h, w)
    _load().viorb_synth_frame(h, w, int(seed), out.ctypes.data)
    return out


def frames(n, h, w, seed0=0, out=None, nthreads=None):
    if out is None:
        out = np.empty((n, h, w), np.uint8)
    assert out.flags["C_CONTIGUOUS"] and out.shape == (n, h, w)
    _load().viorb_synth_frames(n, h, w, int(seed0), out.ctypes.data, nthreads or os.cpu_count() or 1)
    return out


def stereo_pair(h, w, seed, nbands=8, dmin=4, dmax=64):
    left = np.empty((h, w), np.uint8)
    right = np.empty((h, w), np.uint8)
    disp = np.zeros(nbands, np.int32)
    _load().viorb_synth_stereo(h, w, int(seed), nbands, dmin, dmax, left.ctypes.data, right.ctypes.data,
                               disp.ctypes.data)
    return left, right, disp


def random_bytes(n, seed):
    out = np.empty(n, np.uint8)
    _load().viorb_synth_bytes(int(seed), out.ctypes.data, n)
    return out


def descriptor_map(m, seed=1234):
    return random_bytes(m * 32, seed).reshape(m, 32)


def queries_from_map(dmap, q, seed=99, max_flips=40):
    """Half the queries are map rows with k in [0, max_flips] random bit flips, half are fresh random."""
    rng = np.random.default_rng(seed)
    out = random_bytes(q * 32, seed + 7).reshape(q, 32).copy()
    rows = rng.integers(0, dmap.shape[0], size=q // 2)
    for i, r in enumerate(rows):
        d = dmap[r].copy()
        k = int(rng.integers(0, max_flips + 1))
        bits = rng.integers(0, 256, size=k)
        for b in bits:
            d[b >> 3] ^= np.uint8(1 << (b & 7))
        out[i] = d
    return out
