"""Pins the oracle on the REFERENCE'S OWN CODE (CPU tier, no GPU).

oracle/_ref/libviorb_ref.so is /root/reference/src/ORBextractor.cc compiled UNMODIFIED (oracle/refbuild/Makefile) against
stand-in OpenCV headers whose primitives are the cv2-pinned ones (tests/test_ref_minicv.py).  Here:

  1. oracle == reference, every output byte (keypoint records, descriptors, padded pyramid, per-level keypoint lists,
     constructor tables) on the BASELINE.json configurations and on the differential-fuzz corpus;
  2. oracle == the committed outputs of the reference (tests/golden/ref_extract_*.npz, ref_extract_hashes.json, written by
     tests/golden/make_ref_golden.py) -- this part also runs where the reference library is absent;
  3. DistributeOctTree in isolation, including heavy size ties;
  4. the pointer tie-break of src/ORBextractor.cc:684 measured separately: under the process allocator or a descending
     allocator the reference's OWN output moves by 1-3 % of the keypoints, which is why parity is stated against the
     ascending-address convention.
"""
import json
import os

import numpy as np
import pytest

from util import CONFIGS, ROOT, extraction_digest, fuzz_extract_cases, reference_defined
from viorb_b200 import synth

GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_py
    if not ref_py.available():
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref/libviorb_ref.so")
    ref_py.lib()
    ref_py.set_allocator(1)
    ref_py.set_gaussian_variant(0)
    return ref_py


def both(oracle, ref, params, img):
    eo = oracle.Extractor(*params)
    er = oracle.Extractor(*params, _lib=ref.lib())
    return eo, eo(img), er, er(img)


@pytest.mark.parametrize("cfg,seed", [("euroc", 0), ("euroc", 3), ("kitti", 7), ("kitti12", 7), ("odd", 5), ("hd", 0)])
def test_oracle_equals_reference_on_configs(oracle, ref, cfg, seed):
    h, w, nf, sf, nl, it, mt = CONFIGS[cfg]
    img = synth.frame(h, w, seed)
    eo, (ko, do), er, (kr, dr) = both(oracle, ref, (nf, sf, nl, it, mt), img)
    assert len(ko) == len(kr) > 0.9 * nf
    assert ko.tobytes() == kr.tobytes(), "keypoint records differ from the reference"
    assert (do == dr).all(), "descriptors differ from the reference"
    assert eo.quotas() == er.quotas() and eo.umax() == er.umax()
    assert (eo.scale_factors().view(np.uint32) == er.scale_factors().view(np.uint32)).all()
    for l in range(nl):
        assert (eo.pyramid(l) == er.pyramid(l)).all(), "pyramid level %d" % l
        a, b = eo.level_keypoints(l), er.level_keypoints(l)
        assert a.tobytes() == b.tobytes(), "level %d keypoint list (order included)" % l
        a, b = eo.candidates(l), er.candidates(l)          # every cv::FAST call the reference made, cell by cell
        assert a.tobytes() == b.tobytes(), "level %d FAST candidates (vToDistributeKeys)" % l
    assert eo.retried_cells() == er.retried_cells() > 0


def test_oracle_equals_reference_stereo_pair(oracle, ref):
    """BASELINE configs[1]: both images of the synthetic KITTI-shape pair"""
    left, right, _ = synth.stereo_pair(376, 1241, 7)
    for img in (left, right):
        _, (ko, do), _, (kr, dr) = both(oracle, ref, CONFIGS["kitti"][2:], img)
        assert ko.tobytes() == kr.tobytes() and (do == dr).all()


@pytest.mark.parametrize("kind", ["flat", "noise", "checker", "strided"])
def test_oracle_equals_reference_adversarial(oracle, ref, kind):
    rng = np.random.default_rng(3)
    if kind == "flat":
        img = np.full((480, 752), 77, np.uint8)
    elif kind == "noise":
        img = rng.integers(0, 256, (240, 376)).astype(np.uint8)
    elif kind == "checker":
        yy, xx = np.mgrid[0:300, 0:400]
        img = (((yy // 8 + xx // 8) & 1) * 200).astype(np.uint8)
    else:
        img = synth.frame(500, 800, 9)[7:487, 13:765]          # a view with a row stride
    eo = oracle.Extractor(1000, 1.2, 8, 20, 7)
    er = oracle.Extractor(1000, 1.2, 8, 20, 7, _lib=ref.lib())
    if kind == "strided":
        # without BORDER_ISOLATED the reference reads the level-0 border from the parent image of a view
        # (src/ORBextractor.cc:1127); the oracle and the C ABI see (pointer, step) only, i.e. an isolated image
        ko, do = eo(np.ascontiguousarray(img))
        kr, dr = er(np.ascontiguousarray(img))
    else:
        ko, do = eo(img)
        kr, dr = er(img)
    if kind == "flat":
        assert len(ko) == len(kr) == 0
    assert ko.tobytes() == kr.tobytes() and (do == dr).all()


def test_oracle_equals_reference_fuzz(oracle, ref):
    """random shapes, parameters and image statistics (the stream of tools/fuzz_extract.py)"""
    compared = kp = 0
    for c, img, (nf, sf, nl, it, mt) in fuzz_extract_cases(60, 1):
        if not reference_defined(img.shape[0], img.shape[1], sf, nl):
            continue
        _, (ko, do), _, (kr, dr) = both(oracle, ref, (nf, sf, nl, it, mt), img)
        assert ko.tobytes() == kr.tobytes() and (do == dr).all(), "fuzz case %d (%s)" % (c, (img.shape, nf, sf, nl, it, mt))
        compared += 1
        kp += len(ko)
    assert compared >= 35 and kp > 30000
    assert ref.arena_overflows() == 0


def test_oracle_equals_committed_reference_outputs(oracle):
    """runs without the reference library: fixtures written by tests/golden/make_ref_golden.py from the reference itself"""
    for cfg, seed in (("euroc", 0), ("odd", 5), ("kitti12", 7)):
        g = np.load(os.path.join(GOLD, "ref_extract_%s_seed%d.npz" % (cfg, seed)))
        h, w, nf, sf, nl, it, mt = CONFIGS[cfg]
        k, d = oracle.Extractor(nf, sf, nl, it, mt)(synth.frame(h, w, seed))
        assert k.tobytes() == g["keypoints"].tobytes() and (d == g["descriptors"]).all(), cfg
    hashes = json.load(open(os.path.join(GOLD, "ref_extract_hashes.json")))
    left, right, _ = synth.stereo_pair(376, 1241, 7)
    for name, img in (("kitti_left", left), ("kitti_right", right)):
        k, d = oracle.Extractor(*CONFIGS["kitti"][2:])(img)
        assert extraction_digest(k, d) == hashes["configs"][name]["digest"], name
    k, d = oracle.Extractor(*CONFIGS["euroc"][2:])(synth.frame(480, 752, 4095))
    assert extraction_digest(k, d) == hashes["configs"]["euroc_seed4095"]["digest"]
    n = 0
    for c, img, params in fuzz_extract_cases(40, hashes["fuzz"]["seed"]):
        if str(c) not in hashes["fuzz"]["digest"]:
            continue
        k, d = oracle.Extractor(*params)(img)
        assert extraction_digest(k, d) == hashes["fuzz"]["digest"][str(c)], "fuzz case %d" % c
        n += 1
    assert n >= 25


def test_octree_equals_reference(oracle, ref):
    """DistributeOctTree alone (src/ORBextractor.cc:539-763): random candidate clouds, few distinct responses => ties
    in both the node-size sort and the per-node maximum"""
    rng = np.random.default_rng(11)
    from oracle import oracle_py as O
    for case in range(60):
        W, H = int(rng.integers(60, 1300)), int(rng.integers(40, 420))
        if round(np.float32(W) / np.float32(H)) < 1:
            continue
        n = int(rng.integers(1, 6000))
        N = int(rng.integers(1, 1200))
        cand = np.zeros(n, O.CORNER)
        if case % 3 == 0:      # clustered
            cx, cy = rng.integers(0, W, 8), rng.integers(0, H, 8)
            pick = rng.integers(0, 8, n)
            cand["x"] = np.clip(cx[pick] + rng.normal(0, 12, n), 0, W - 1).astype(np.int32)
            cand["y"] = np.clip(cy[pick] + rng.normal(0, 12, n), 0, H - 1).astype(np.int32)
        else:
            cand["x"], cand["y"] = rng.integers(0, W, n), rng.integers(0, H, n)
        cand["score"] = rng.integers(7, 7 + int(rng.choice([2, 5, 60])), n)
        # cv::FAST lists corners row-major inside a cell and cells row-major: any order is legal input for the quadtree
        a = oracle.distribute_octree(cand, 16, 16 + W, 16, 16 + H, N)
        with O.using(ref.lib()):
            b = O.distribute_octree(cand, 16, 16 + W, 16, 16 + H, N)
        assert len(a) == len(b) and (a == b).all(), "case %d: n=%d N=%d %dx%d" % (case, n, N, W, H)


def test_pointer_tie_break_is_allocator_dependent(oracle, ref, capsys):
    """Not a parity check: documents how far the reference's own output moves when only the heap layout changes."""
    img = synth.frame(480, 752, 0)
    ko, _ = oracle.Extractor(1000, 1.2, 8, 20, 7)(img)
    base = set(zip(ko["x"].tolist(), ko["y"].tolist(), ko["octave"].tolist()))
    share = {}
    try:
        for mode, name in ((1, "ascending"), (2, "descending"), (0, "glibc malloc")):
            ref.set_allocator(mode)
            kr, _ = oracle.Extractor(1000, 1.2, 8, 20, 7, _lib=ref.lib())(img)
            got = set(zip(kr["x"].tolist(), kr["y"].tolist(), kr["octave"].tolist()))
            share[name] = len(base & got) / len(base)
    finally:
        ref.set_allocator(1)
    with capsys.disabled():
        print("\n[pointer tie-break, src/ORBextractor.cc:684] share of the oracle's keypoints the reference reproduces: %s" % share)
    assert share["ascending"] == 1.0
    assert 0.9 < share["descending"] < 1.0, "descending addresses must flip some ties (SURVEY.md Appendix C.1)"
    assert share["glibc malloc"] > 0.9


def test_opencv24_gaussian_variant(oracle, ref):
    """the 2.4 taps [18,34,49,55,49,34,18]: keypoints unchanged, descriptors differ; oracle == reference in both variants"""
    img = synth.frame(480, 752, 0)
    from oracle import oracle_py as O
    a = O.gaussian7(img, 0).astype(np.int32)
    b = O.gaussian7(img, 1).astype(np.int32)
    assert np.abs(a - b).max() <= 2 and (a != b).mean() > 0.5
    try:
        ref.set_gaussian_variant(1)
        kr, dr = oracle.Extractor(1000, 1.2, 8, 20, 7, _lib=ref.lib())(img)
    finally:
        ref.set_gaussian_variant(0)
    eo = oracle.Extractor(1000, 1.2, 8, 20, 7)
    eo.set_gaussian_variant(1)
    ko, do = eo(img)
    assert ko.tobytes() == kr.tobytes() and (do == dr).all()
    k0, d0 = oracle.Extractor(1000, 1.2, 8, 20, 7)(img)
    assert k0.tobytes() == ko.tobytes() and (d0 != do).any()
    hashes = json.load(open(os.path.join(GOLD, "ref_extract_hashes.json")))
    assert extraction_digest(ko, do) == hashes["cv24"]["euroc_seed0"]["digest"]
