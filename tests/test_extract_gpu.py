"""GPU parity tests of the extraction path: libviorb_b200.so (through the C ABI) vs the CPU oracle.

Bar (BASELINE.json north_star): pyramid, FAST scores/candidates and descriptors bit-exact; selected
keypoint sets >= 99.9 % (we require 100 % incl. order); IC_Angle within 1e-3 rad (we require bit-equal).
"""
import os

import numpy as np
import pytest

from util import CONFIGS, ROOT
from viorb_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def api():
    from viorb_b200 import api
    api.lib()
    return api


@pytest.fixture(scope="module")
def ctx(api):
    c = api.Context(0)
    yield c
    c.close()


def sort_rows(a):
    a = np.asarray(a).reshape(-1, 3)
    return a[np.lexsort((a[:, 2], a[:, 0], a[:, 1]))]


def assert_same_output(k_gpu, d_gpu, k_ref, d_ref):
    assert len(k_gpu) == len(k_ref), (len(k_gpu), len(k_ref))
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert (k_gpu[f] == k_ref[f]).all(), f
    # angle: bar is 1e-3 rad; the implementation is bit-exact
    assert np.max(np.abs(k_gpu["angle"] - k_ref["angle"]), initial=0) * np.pi / 180 <= 1e-3
    assert (k_gpu["angle"].view(np.uint32) == k_ref["angle"].view(np.uint32)).all()
    assert (d_gpu == d_ref).all()


@pytest.mark.parametrize("cfg,seed", [("euroc", 0), ("odd", 5), ("kitti12", 7), ("kitti", 3)])
def test_stages_vs_oracle(api, ctx, oracle, cfg, seed):
    h, w, nf, sf, nl, it, mt = CONFIGS[cfg]
    img = synth.frame(h, w, seed)
    ref = oracle.Extractor(nf, sf, nl, it, mt)
    k_ref, d_ref = ref(img)
    ex = api.ORBextractor(nf, sf, nl, it, mt, ctx=ctx)
    k_gpu, d_gpu = ex(img)
    assert (ex.features_per_level() == ref.quotas()).all()
    assert (ex.GetScaleFactors() == ref.scale_factors()).all()
    levels = ex.pyramid_all()                 # one copy for all levels (viorb_extractor_pyramid_download_all)
    for l in range(nl):
        assert (ex.pyramid(l) == ref.pyramid(l)).all(), "pyramid level %d" % l
        assert (levels[l] == ref.pyramid(l)).all(), "pyramid level %d, single-copy download" % l
        c_ref = ref.candidates(l)
        c_ref = np.stack([c_ref["x"] + 16, c_ref["y"] + 16, c_ref["score"]], 1)
        c_gpu = ex.debug_candidates(l)
        assert len(c_gpu) == len(c_ref), "candidate count level %d" % l
        assert (sort_rows(c_gpu) == sort_rows(c_ref)).all(), "candidates level %d" % l
        s_ref = ref.level_keypoints(l)
        s_ref = np.stack([s_ref["x"], s_ref["y"], s_ref["response"]], 1).astype(np.int32)
        s_gpu = ex.debug_selected(l)
        assert len(s_gpu) == len(s_ref), "selected count level %d" % l
        assert (s_gpu == s_ref).all(), "selected keypoints (list order) level %d" % l
    assert_same_output(k_gpu, d_gpu, k_ref, d_ref)
    ex.close()


@pytest.mark.parametrize("cfg,seed", [("euroc", 0), ("odd", 5), ("kitti12", 7)])
def test_vs_golden_opencv(api, ctx, cfg, seed):
    """against the committed outputs of real OpenCV primitives (tests/golden/make_golden.py)"""
    h, w, nf, sf, nl, it, mt = CONFIGS[cfg]
    g = np.load(os.path.join(GOLD, "extract_%s_seed%d.npz" % (cfg, seed)))
    ex = api.ORBextractor(nf, sf, nl, it, mt, ctx=ctx)
    k, d = ex(synth.frame(h, w, seed))
    assert_same_output(k, d, g["keypoints"], g["descriptors"])
    ex.close()


def test_steering_exhaustive_vs_libm(ctx, oracle):
    """the kernel's sincosf (glibc flt-32 algorithm, no FMA) against this image's libm for every float
    angle in [0, 360] degrees -- all 1 135 869 953 bit patterns, outputs compared bit for bit."""
    last = int(np.array([360.0], np.float32).view(np.uint32)[0])
    chunk = 1 << 26
    for first in range(0, last + 1, chunk):
        n = min(chunk, last + 1 - first)
        s, c = ctx.debug_steering(first, n)
        rs, rc, bad = oracle.steering_sweep(first, n, which=0)
        assert bad == 0
        assert (s.view(np.uint32) == rs.view(np.uint32)).all() and (c.view(np.uint32) == rc.view(np.uint32)).all(), hex(first)


def test_orientation_dense_vs_oracle(ctx, oracle):
    """fastAtan2 on integer moments (|m| <= 255*4896): 2^27 random pairs, every pair of a small-magnitude
    square (ties, axes, diagonals, zero) and the extreme corners, bit for bit."""
    lim = 255 * 4896
    rng = np.random.default_rng(11)
    small = np.arange(-300, 301, dtype=np.int32)
    sets = [(np.repeat(small, small.size), np.tile(small, small.size)),
            (np.array([0, 0, lim, -lim, lim, -lim, lim, -lim], np.int32), np.array([0, lim, 0, 0, lim, lim, -lim, -lim], np.int32))]
    for _ in range(4):
        sets.append((rng.integers(-lim, lim + 1, 1 << 25, dtype=np.int32), rng.integers(-lim, lim + 1, 1 << 25, dtype=np.int32)))
    for m01, m10 in sets:
        g = ctx.debug_orientation(m01, m10)
        r = oracle.orientation_sweep(m01, m10)
        assert (g.view(np.uint32) == r.view(np.uint32)).all()


def test_batch_matches_single(api, ctx, oracle):
    h, w, nf, sf, nl, it, mt = CONFIGS["euroc"]
    B = 9
    imgs = synth.frames(B, h, w, seed0=100)
    ex = api.ORBextractor(nf, sf, nl, it, mt, ctx=ctx)
    ex.configure(chunk_frames=4)          # 3 passes, last one ragged
    kps, desc, counts = ex.extract_batch(imgs)
    ref = oracle.Extractor(nf, sf, nl, it, mt)
    for b in range(B):
        k_ref, d_ref = ref(imgs[b])
        n = counts[b]
        assert_same_output(kps[b, :n], desc[b, :n], k_ref, d_ref)
    # the pyramids of the last pass stay resident (mvImagePyramid)
    assert (ex.pyramid(3, frame=0) == ref.pyramid(3)).all()
    ex.close()


def test_host_batch_default_pass_size(api, ctx, oracle):
    """Host-buffer batch call without configure(): the pass size follows the batch (32..128 frames), here 75 small
    frames -> passes of 32, 32 and 11 through three staging slots."""
    B, h, w = 75, 120, 160
    imgs = synth.frames(B, h, w, seed0=900)
    ex = api.ORBextractor(300, 1.2, 3, 20, 7, ctx=ctx)
    kps, desc, counts = ex.extract_batch(imgs)
    ref = oracle.Extractor(300, 1.2, 3, 20, 7)
    for b in range(B):
        k_ref, d_ref = ref(imgs[b])
        assert_same_output(kps[b, :counts[b]], desc[b, :counts[b]], k_ref, d_ref)
    ex.close()


@pytest.mark.parametrize("cfg", ["hd", "uhd"])
def test_large_frames(api, ctx, oracle, cfg):
    """BASELINE config 4: 1920x1080 and 3840x2160, 5000 features"""
    h, w, nf, sf, nl, it, mt = CONFIGS[cfg]
    img = synth.frame(h, w, 2)
    ref = oracle.Extractor(nf, sf, nl, it, mt)
    k_ref, d_ref = ref(img)
    ex = api.ORBextractor(nf, sf, nl, it, mt, ctx=ctx)
    k, d = ex(img)
    assert_same_output(k, d, k_ref, d_ref)
    ex.close()


def test_edge_cases(api, ctx, oracle):
    ex = api.ORBextractor(1000, 1.2, 8, 20, 7, ctx=ctx)
    ref = oracle.Extractor(1000, 1.2, 8, 20, 7)
    # empty image: silent return (ORBextractor.cc:1046-1047)
    k, d = ex(np.zeros((0, 0), np.uint8))
    assert len(k) == 0
    # flat image: no keypoints (:1064-1065)
    k, d = ex(np.full((480, 752), 77, np.uint8))
    assert len(k) == 0 and d.shape == (0, 32)
    # saturated checkerboard and pure noise (retry path everywhere / candidate-heavy)
    yy, xx = np.mgrid[0:480, 0:752]
    checker = (((yy // 8 + xx // 8) & 1) * 255).astype(np.uint8)
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, (480, 752)).astype(np.uint8)
    for img in (checker, noise):
        k_ref, d_ref = ref(img)
        k, d = ex(img)
        assert_same_output(k, d, k_ref, d_ref)
    # a sub-matrix view with a row stride
    big = synth.frame(500, 800, 8)
    view = big[10:490, 20:772]
    k_ref, d_ref = ref(np.ascontiguousarray(view))
    k, d = ex(view)
    assert_same_output(k, d, k_ref, d_ref)
    ex.close()


def test_bad_arguments(api, ctx):
    with pytest.raises(api.ViorbError):
        api.ORBextractor(0, 1.2, 8, 20, 7, ctx=ctx)
    with pytest.raises(api.ViorbError):
        api.ORBextractor(1000, 1.2, 64, 20, 7, ctx=ctx)
    ex = api.ORBextractor(1000, 1.2, 8, 20, 7, ctx=ctx)
    with pytest.raises(api.ViorbError):
        ex(np.zeros((40, 40), np.uint8))        # smaller than one FAST cell at the top level
    ex.close()


# (h, w, nfeatures, scale, levels, iniTh, minTh): odd sizes (unaligned rows, ragged tiles, single-cell levels), the
# scale-factor envelope (1.1 .. 1.5), few/many levels, thresholds far apart and equal
SWEEP = [(97, 131, 200, 1.2, 3, 20, 7), (333, 250, 500, 1.25, 5, 15, 5), (480, 640, 1500, 1.5, 4, 20, 7),
         (241, 1023, 800, 1.1, 8, 25, 10), (720, 1280, 3000, 1.3, 6, 20, 20), (128, 128, 300, 1.2, 4, 30, 3),
         (479, 751, 1000, 1.2, 8, 20, 7), (600, 799, 1200, 1.41, 5, 12, 7),
         # level 3 is 181x91: one row of 38x59 cells whose last group starts at another byte shift than the first -- two
         # instantiations of the FAST kernel, each with its own slice of the group table (found by the fuzz, seed 31)
         (158, 312, 236, 1.2, 6, 12, 7)]


@pytest.mark.parametrize("cfg", SWEEP, ids=lambda c: "%dx%d_s%.2f_l%d" % (c[1], c[0], c[3], c[4]))
def test_parameter_sweep(api, ctx, oracle, cfg):
    """whole operator() on shapes and parameters away from the benchmark configs (strided view of a wider buffer too)"""
    h, w, nf, sf, nl, it, mt = cfg
    img = synth.frame(h, w, h * 7 + w)
    ref = oracle.Extractor(nf, sf, nl, it, mt)
    k_ref, d_ref = ref(img)
    ex = api.ORBextractor(nf, sf, nl, it, mt, ctx=ctx)
    k_gpu, d_gpu = ex(img)
    assert len(k_ref) > 20
    assert_same_output(k_gpu, d_gpu, k_ref, d_ref)
    for l in range(nl):
        assert (ex.pyramid(l) == ref.pyramid(l)).all(), "pyramid level %d" % l
    wide = np.zeros((h, w + 13), np.uint8)
    wide[:, 5:5 + w] = img
    k2, d2 = ex(wide[:, 5:5 + w])          # unaligned base pointer and row stride
    assert_same_output(k2, d2, k_ref, d_ref)
    ex.close()


@pytest.mark.parametrize("mode", [1, 2], ids=["per_keypoint", "whole_levels"])
def test_describe_modes(api, ctx, oracle, mode):
    """viorb_extractor_set_describe_mode: the GaussianBlur evaluated per keypoint (fused kernel) or once per level
    (blur_levels_kernel + describe_blurred_kernel) -- forced either way on single frames, odd shapes, unaligned strided
    views, saturated / noisy / flat images, both Gaussian variants and a batch with a ragged last pass; every byte against
    the oracle.  (Without the call the library picks per geometry and pass size.)"""
    for cfg in SWEEP + [CONFIGS["kitti"]]:
        h, w, nf, sf, nl, it, mt = cfg
        img = synth.frame(h, w, h * 7 + w)
        ref = oracle.Extractor(nf, sf, nl, it, mt)
        k_ref, d_ref = ref(img)
        ex = api.ORBextractor(nf, sf, nl, it, mt, ctx=ctx)
        ex.set_describe_mode(mode)
        k, d = ex(img)
        assert_same_output(k, d, k_ref, d_ref)
        wide = np.zeros((h, w + 13), np.uint8)
        wide[:, 5:5 + w] = img
        k2, d2 = ex(wide[:, 5:5 + w])
        assert_same_output(k2, d2, k_ref, d_ref)
        ex.close()
    ex = api.ORBextractor(1000, 1.2, 8, 20, 7, ctx=ctx)
    ex.set_describe_mode(mode)
    ref = oracle.Extractor(1000, 1.2, 8, 20, 7)
    yy, xx = np.mgrid[0:480, 0:752]
    checker = (((yy // 8 + xx // 8) & 1) * 255).astype(np.uint8)
    noise = np.random.default_rng(3).integers(0, 256, (480, 752)).astype(np.uint8)
    sat = synth.frame(480, 752, 0).copy()
    sat[sat > 140] = 255
    for img in (checker, noise, sat):
        k_ref, d_ref = ref(img)
        k, d = ex(img)
        assert_same_output(k, d, k_ref, d_ref)
    k, d = ex(np.full((480, 752), 77, np.uint8))
    assert len(k) == 0
    # a batch: 13 frames in passes of 5, 5 and 3 (and 13 > 8: the automatic mode would blur whole levels here)
    imgs = synth.frames(13, 480, 752, seed0=300)
    ex.configure(chunk_frames=5)
    kps, desc, counts = ex.extract_batch(imgs)
    for b in range(13):
        k_ref, d_ref = ref(imgs[b])
        assert_same_output(kps[b, :counts[b]], desc[b, :counts[b]], k_ref, d_ref)
    ex.close()
    # OpenCV 2.4 taps: digest written by the reference itself
    import json
    from util import extraction_digest
    hashes = json.load(open(os.path.join(GOLD, "ref_extract_hashes.json")))
    ex = api.ORBextractor(*CONFIGS["euroc"][2:], ctx=ctx)
    ex.set_describe_mode(mode)
    ex.set_gaussian(1)
    for img in (synth.frame(480, 752, 0),):
        k, d = ex(img)
        assert extraction_digest(k, d) == hashes["cv24"]["euroc_seed0"]["digest"]
    imgs = synth.frames(9, 480, 752, seed0=0)
    kps, desc, counts = ex.extract_batch(imgs)
    assert extraction_digest(kps[0, :counts[0]], desc[0, :counts[0]]) == hashes["cv24"]["euroc_seed0"]["digest"]
    ex.close()


def test_device_batch_path_and_device_frame_index(api, ctx, oracle):
    """viorb_extract_batch_device (device frames in, device outputs, ragged last pass) returns the bytes of the host
    path, and a frame index built from its device outputs (undistortion + grid on the device, no host round trip)
    equals the oracle's undistorted keypoints and grid queries"""
    torch = pytest.importorskip("torch")
    h, w, nf, sf, nl, it, mt = CONFIGS["euroc"]
    B = 9
    imgs = synth.frames(B, h, w, seed0=40)
    ex = api.ORBextractor(nf, sf, nl, it, mt, ctx=ctx)
    ex.configure(chunk_frames=4)
    kps_h, desc_h, cnt_h = ex.extract_batch(imgs)
    cap = ex.cap
    d_imgs = torch.from_numpy(imgs).cuda()
    d_kps = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros((B,), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ex.extract_batch_device(d_imgs, B, h, w, d_kps, d_desc, d_cnt)
    ex.check()
    cnt_d = d_cnt.cpu().numpy()
    assert (cnt_d == cnt_h).all()
    kd = d_kps.cpu().numpy().view(np.uint8).reshape(B, cap, 28)
    dd = d_desc.cpu().numpy()
    kh = np.ascontiguousarray(kps_h).view(np.uint8).reshape(B, cap, 28)
    for b in range(B):
        n = int(cnt_h[b])
        assert (kd[b, :n] == kh[b, :n]).all() and (dd[b, :n] == np.asarray(desc_h)[b, :n]).all()
    # frame 3 straight from the device outputs into a matcher index
    K, dist = (458.654, 457.296, 367.215, 248.375), [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]
    b, n = 3, int(cnt_h[3])
    sfs = ex.GetScaleFactors()
    fi = api.FrameIndex.from_device(ctx, d_kps[b], d_desc[b], n, K, dist, (w, h), sfs)
    k_un, bounds = fi.keys()
    raw = np.ascontiguousarray(kps_h[b][:n])
    ref = oracle.undistort_keypoints(raw, *K, dist)
    bref = oracle.compute_image_bounds(w, h, *K, dist)
    assert k_un.tobytes() == ref.tobytes() and (bounds.view(np.uint32) == bref.view(np.uint32)).all()
    g = oracle.Grid(ref, *[float(v) for v in bref])
    rng = np.random.default_rng(0)
    for _ in range(30):
        x, y, r = float(rng.uniform(0, w)), float(rng.uniform(0, h)), float(rng.choice([8.0, 30.0]))
        a, c = fi.GetFeaturesInArea(x, y, r), g.features_in_area(x, y, r)
        assert len(a) == len(c) and (a == c).all()
    fi.close()
    ex.close()


def test_two_extractors_with_different_quotas_interleaved(api, ctx, oracle):
    """the quadtree kernel's opt-in shared-memory limit is per kernel, not per extractor: a small-quota extractor
    created after a large-quota one must not break the large one"""
    big_cfg, small_cfg = CONFIGS["hd"], (240, 320, 150, 1.2, 4, 20, 7)
    img_b, img_s = synth.frame(big_cfg[0], big_cfg[1], 2), synth.frame(small_cfg[0], small_cfg[1], 3)
    ex_b = api.ORBextractor(*big_cfg[2:], ctx=ctx)
    kb, db = ex_b(img_b)
    ex_s = api.ORBextractor(*small_cfg[2:], ctx=ctx)
    ks, ds = ex_s(img_s)
    kb2, db2 = ex_b(img_b)                       # the large extractor again, after the small one was set up
    assert kb2.tobytes() == kb.tobytes() and (db2 == db).all()
    rs = oracle.Extractor(*small_cfg[2:])
    k_ref, d_ref = rs(img_s)
    assert_same_output(ks, ds, k_ref, d_ref)
    ex_b.close()
    ex_s.close()


def test_two_extractors_in_two_threads(api, oracle):
    """the stereo Frame constructor runs the left and right extractor in two threads (src/Frame.cc:258-261): distinct
    instances (each with its own context / stream / graph) must be safe to enter concurrently"""
    import threading
    h, w, nf, sf, nl, it, mt = CONFIGS["kitti"]
    left, right, _ = synth.stereo_pair(h, w, 11)
    ref = oracle.Extractor(nf, sf, nl, it, mt)
    want = [ref(left), ref(right)]
    got, errs = [None, None], []

    def work(i, img):
        try:
            c = api.Context(0)
            ex = api.ORBextractor(nf, sf, nl, it, mt, ctx=c)
            for _ in range(25):
                got[i] = ex(img)
            ex.close()
            c.close()
        except Exception as e:      # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(0, left)), threading.Thread(target=work, args=(1, right))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for i in range(2):
        assert_same_output(got[i][0], got[i][1], want[i][0], want[i][1])


def test_differential_fuzz():
    """tools/fuzz_extract.py: random shapes / parameters / image statistics (noise, low contrast, checkerboards, half-flat
    frames), every output byte against the oracle"""
    import subprocess
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_extract.py"), "80", "3"], capture_output=True,
                       text=True, cwd=ROOT, timeout=600)
    assert p.returncode == 0 and " 0 mismatches" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.parametrize("kind", ["scene", "saturated"])
def test_opencv24_gaussian_variant(api, ctx, oracle, kind):
    """viorb_extractor_set_gaussian(VIORB_GAUSSIAN_OPENCV24): the 8-bit kernel of the OpenCV the reference pins
    ([18,34,49,55,49,34,18], sum 257, saturating) -- descriptors bit-equal to the oracle / the reference built with
    those taps, and back to the default taps afterwards"""
    img = synth.frame(480, 752, 0)
    if kind == "saturated":                                   # 255-valued regions drive the 257/256 gain into saturation
        img = img.copy()
        img[img > 140] = 255
    ex = api.ORBextractor(1000, 1.2, 8, 20, 7, ctx=ctx)
    k0, d0 = ex(img)
    ex.set_gaussian(1)
    k1, d1 = ex(img)
    ref = oracle.Extractor(1000, 1.2, 8, 20, 7)
    from oracle import oracle_py, ref_py
    under_reference = oracle_py._override is not None
    if under_reference:
        ref_py.set_gaussian_variant(1)
    else:
        ref.set_gaussian_variant(1)
    try:
        kr, dr = ref(img)
    finally:
        if under_reference:
            ref_py.set_gaussian_variant(0)
    assert len(k1) == len(kr) > 500 and k1.tobytes() == kr.tobytes() and (d1 == dr).all()
    assert k0.tobytes() == k1.tobytes() and (d0 != d1).any()
    ex.set_gaussian(0)
    k2, d2 = ex(img)
    assert (d2 == d0).all()
    ex.close()


def test_extract_equals_committed_reference_outputs(api, ctx):
    """GPU output against files written by the REFERENCE ITSELF (src/ORBextractor.cc compiled unmodified and run in the
    build container, tests/golden/make_ref_golden.py): full records for three configurations, digests for the KITTI
    stereo pair, 1080p, 4K and the differential-fuzz corpus.  Needs neither the oracle nor the reference library."""
    import json
    from util import extraction_digest, fuzz_extract_cases
    gold = os.path.join(ROOT, "tests", "golden")
    for cfg, seed in (("euroc", 0), ("odd", 5), ("kitti12", 7)):
        g = np.load(os.path.join(gold, "ref_extract_%s_seed%d.npz" % (cfg, seed)))
        h, w, nf, sf, nl, it, mt = CONFIGS[cfg]
        ex = api.ORBextractor(nf, sf, nl, it, mt, ctx=ctx)
        k, d = ex(synth.frame(h, w, seed))
        assert k.tobytes() == g["keypoints"].tobytes() and (d == g["descriptors"]).all(), cfg
        ex.close()
    hashes = json.load(open(os.path.join(gold, "ref_extract_hashes.json")))
    left, right, _ = synth.stereo_pair(376, 1241, 7)
    ex = api.ORBextractor(*CONFIGS["kitti"][2:], ctx=ctx)
    for name, img in (("kitti_left", left), ("kitti_right", right)):
        k, d = ex(img)
        assert extraction_digest(k, d) == hashes["configs"][name]["digest"], name
    ex.close()
    for cfg, seed in (("euroc", 1), ("euroc", 4095), ("hd", 0), ("uhd", 0)):
        h, w = CONFIGS[cfg][:2]
        ex = api.ORBextractor(*CONFIGS[cfg][2:], ctx=ctx)
        k, d = ex(synth.frame(h, w, seed))
        assert extraction_digest(k, d) == hashes["configs"]["%s_seed%d" % (cfg, seed)]["digest"], (cfg, seed)
        ex.close()
    ex = api.ORBextractor(*CONFIGS["euroc"][2:], ctx=ctx)
    ex.set_gaussian(1)
    k, d = ex(synth.frame(480, 752, 0))
    assert extraction_digest(k, d) == hashes["cv24"]["euroc_seed0"]["digest"]
    ex.close()
    compared = 0
    for c, img, params in fuzz_extract_cases(hashes["fuzz"]["cases"], hashes["fuzz"]["seed"]):
        if str(c) not in hashes["fuzz"]["digest"]:
            continue
        try:
            ex = api.ORBextractor(*params, ctx=ctx)
            k, d = ex(img)
        except api.ViorbError as e:
            assert e.code == -4, e          # outside the documented envelope: refused, never guessed
            continue
        assert extraction_digest(k, d) == hashes["fuzz"]["digest"][str(c)], "fuzz case %d %s" % (c, (img.shape, params))
        ex.close()
        compared += 1
    assert compared >= 120


def test_serial_copy_mode_equals_duplex(api, ctx):
    """viorb_extractor_set_copy_mode(VIORB_COPY_SERIAL): the output copies of viorb_extract_batch queue behind the input copies
    on one stream (two passes late); results must not depend on the order -- ragged last pass, more passes than slots"""
    imgs = synth.frames(150, 240, 320, seed0=40)
    ex = api.ORBextractor(300, 1.2, 6, 20, 7, ctx=ctx)
    ex.configure(chunk_frames=16)
    k0, d0, c0 = ex.extract_batch(imgs)
    ex.set_copy_mode(1)
    k1, d1, c1 = ex.extract_batch(imgs)
    ex.set_copy_mode(0)
    assert (c0 == c1).all() and c0.min() > 100
    for b in range(len(imgs)):
        n = c0[b]
        assert k0[b, :n].tobytes() == k1[b, :n].tobytes() and (d0[b, :n] == d1[b, :n]).all(), b
    ex.close()
