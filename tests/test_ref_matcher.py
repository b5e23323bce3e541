"""Pins the matcher oracle on the REFERENCE'S OWN CODE (CPU tier, no GPU).

oracle/_ref/libviorb_ref.so holds src/ORBmatcher.cc compiled whole and unmodified, the ORB members of src/Frame.cc,
src/KeyFrame.cc and src/MapPoint.cc compiled from verbatim line ranges, and Thirdparty/DBoW2 (oracle/refbuild/Makefile).
Each test runs one scenario through the restatement (oracle/match_oracle.cpp, bow_oracle.cpp) and through the reference
and compares every output: match tables, counts, the updated "already matched" state, float results as raw bits.
"""
import ctypes as C

import numpy as np
import pytest

import scenarios as S
from util import CONFIGS
from viorb_b200 import synth


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_py
    if not ref_py.available():
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref/libviorb_ref.so")
    ref_py.lib()
    return ref_py


@pytest.fixture(scope="module")
def pair(oracle):
    """KITTI-shape synthetic stereo pair (BASELINE configs[1]) extracted by the oracle"""
    h, w, nf, sf, nl, it, mt = CONFIGS["kitti"]
    left, right, _ = synth.stereo_pair(h, w, 7)
    e1, e2 = oracle.Extractor(nf, sf, nl, it, mt), oracle.Extractor(nf, sf, nl, it, mt)
    k1, d1 = e1(left)
    k2, d2 = e2(right)
    return dict(k1=k1, d1=d1, k2=k2, d2=d2, e1=e1, e2=e2, sf=e1.scale_factors(), shape=(h, w), nl=nl,
                bounds=(0.0, float(w), 0.0, float(h)))


def snap(ref, sc):
    """(u, v, 1/z) the adapters' trivial camera reproduces exactly (oracle/refbuild/ref_matcher_capi.cpp header)"""
    u, v, iz = (np.ascontiguousarray(sc[k], np.float32).copy() for k in ("proj_x", "proj_y", "invz"))
    ref.lib()._l.ref_snap_projection(C.c_void_p(u.ctypes.data), C.c_void_p(v.ctypes.data), C.c_void_p(iz.ctypes.data), len(u))
    return u, v, iz


def both(oracle, ref, fn, *args, **kw):
    a = fn(*args, **kw)
    with oracle.using(ref.lib()):
        b = fn(*args, **kw)
    return a, b


def test_descriptor_distance(oracle, ref):
    rng = np.random.default_rng(0)
    d = rng.integers(0, 256, (400, 32)).astype(np.uint8)
    d[1] = d[0]
    d[3] = ~d[2]
    for i in range(0, 400, 2):
        a, b = both(oracle, ref, oracle.descriptor_distance, d[i], d[i + 1])
        assert a == b == int(np.unpackbits(d[i] ^ d[i + 1]).sum())


@pytest.mark.parametrize("cfg,seed", [("kitti", 7), ("kitti12", 3), ("euroc", 2)])
def test_compute_stereo_matches(oracle, ref, cfg, seed):
    """Frame::ComputeStereoMatches, src/Frame.cc:646-820: mvuRight and mvDepth bit for bit"""
    h, w, nf, sf, nl, it, mt = CONFIGS[cfg]
    left, right, _ = synth.stereo_pair(h, w, seed)
    ol, orr = oracle.Extractor(nf, sf, nl, it, mt), oracle.Extractor(nf, sf, nl, it, mt)
    kl, dl = ol(left)
    kr, dr = orr(right)
    fx = S.KITTI_FX if cfg != "euroc" else 435.2
    mbf, mb = S.KITTI_BF, S.KITTI_BF / fx
    pl, pr = [ol.pyramid(l) for l in range(nl)], [orr.pyramid(l) for l in range(nl)]
    (ur, depth, _, _, n), (ur2, depth2, _, _, n2) = both(oracle, ref, oracle.stereo_match, kl, dl, kr, dr, pl, pr,
                                                         ol.scale_factors(), mbf, mb)
    assert n > 100 and n == n2
    assert (ur.view(np.uint32) == ur2.view(np.uint32)).all()
    assert (depth.view(np.uint32) == depth2.view(np.uint32)).all()
    # no right keypoints at all: every left keypoint stays unmatched
    (ur, depth, _, _, n), (ur2, depth2, _, _, n2) = both(oracle, ref, oracle.stereo_match, kl, dl, kr[:0], dr[:0], pl, pr,
                                                         ol.scale_factors(), mbf, mb)
    assert n == n2 == 0 and (ur == ur2).all() and (ur == -1).all()


def test_frame_grid(oracle, ref, pair):
    """AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (src/Frame.cc:410-425, 507-572): same indices, same order"""
    f = pair
    h, w = f["shape"]
    g = oracle.Grid(f["k1"], *f["bounds"])
    with oracle.using(ref.lib()):
        g2 = oracle.Grid(f["k1"], *f["bounds"])
        rng = np.random.default_rng(5)
        for _ in range(300):
            x, y = float(rng.uniform(-30, w + 30)), float(rng.uniform(-30, h + 30))
            r = float(rng.choice([0.5, 3.0, 10.0, 25.0, 60.0, 400.0]))
            lo, hi = [(-1, -1), (0, 3), (2, -1), (1, 2), (4, 4), (0, 0), (7, -1)][int(rng.integers(0, 7))]
            b = g2.features_in_area(x, y, r, lo, hi)
            with oracle.using(None):
                a = g.features_in_area(x, y, r, lo, hi)
            assert a.tolist() == b.tolist()
    # keypoints outside of the image bounds (undistortion can push them out) are not in the grid
    k = f["k1"].copy()
    k["x"][:50] -= 40.0
    k["y"][50:90] += 400.0
    a = oracle.Grid(k, *f["bounds"]).features_in_area(20.0, 100.0, 80.0)
    with oracle.using(ref.lib()):
        b = oracle.Grid(k, *f["bounds"]).features_in_area(20.0, 100.0, 80.0)
    assert a.tolist() == b.tolist()


@pytest.mark.parametrize("th,nnratio,seed", [(1.0, 0.8, 11), (3.0, 0.8, 12), (5.0, 0.6, 13), (3.0, 0.9, 14)])
def test_search_by_projection_local(oracle, ref, pair, th, nnratio, seed):
    """ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th), src/ORBmatcher.cc:45-129"""
    f = pair
    sc = S.projection_scenario(f["k1"], f["d1"], f["sf"], seed=seed)

    def run():
        return oracle.search_by_projection_local(oracle.Grid(f["k1"], *f["bounds"]), f["d1"], sc["u_right"], sc["obs0"], f["sf"],
                                                 sc["proj_x"], sc["proj_y"], sc["proj_xr"], sc["pred_level"], sc["view_cos"],
                                                 sc["valid"], sc["nobs"], sc["mp_desc"], th, nnratio)
    (n, m, obs), (n2, m2, obs2) = both(oracle, ref, run)
    assert n > 50 and n == n2 and (m == m2).all() and (obs == obs2).all()


def test_search_by_projection_local_many_points(oracle, ref, pair):
    """a local map of 5000 points (Tracking::SearchLocalPoints sizes), most of them competing for the same keypoints"""
    f = pair
    sc = S.projection_scenario(f["k1"], f["d1"], f["sf"], seed=21, n_mp=5000, conflicts=1500)

    def run():
        return oracle.search_by_projection_local(oracle.Grid(f["k1"], *f["bounds"]), f["d1"], sc["u_right"], sc["obs0"], f["sf"],
                                                 sc["proj_x"], sc["proj_y"], sc["proj_xr"], sc["pred_level"], sc["view_cos"],
                                                 sc["valid"], sc["nobs"], sc["mp_desc"], 3.0, 0.8)
    (n, m, obs), (n2, m2, obs2) = both(oracle, ref, run)
    assert n > 300 and n == n2 and (m == m2).all() and (obs == obs2).all()


@pytest.mark.parametrize("mode,th,check_ori", [(0, 15.0, True), (1, 7.0, True), (2, 7.0, False), (0, 7.0, False), (1, 15.0, False)])
def test_search_by_projection_last_frame(oracle, ref, pair, mode, th, check_ori):
    """ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono), src/ORBmatcher.cc:1328-1471, with the forward /
    backward level windows and the rotation histogram (ComputeThreeMaxima :1602-1643)"""
    f = pair
    sc = S.projection_scenario(f["k1"], f["d1"], f["sf"], seed=11)
    u, v, iz = snap(ref, sc)

    def run():
        return oracle.search_by_projection_frame(oracle.Grid(f["k1"], *f["bounds"]), f["d1"], sc["u_right"], sc["obs0"], f["sf"], u, v,
                                                 iz, sc["last_octave"], sc["last_angle"], sc["valid"], sc["nobs"], sc["mp_desc"],
                                                 th, S.KITTI_BF, mode, check_ori, 100)
    (n, m, obs), (n2, m2, obs2) = both(oracle, ref, run)
    assert n2 >= 0, "the adapter refused the scenario (%d)" % n2
    assert n > 30 and n == n2 and (m == m2).all() and (obs == obs2).all()


@pytest.mark.parametrize("th,orb_dist,check_ori", [(10.0, 100, True), (3.0, 64, True), (10.0, 100, False)])
def test_search_by_projection_relocalisation(oracle, ref, pair, th, orb_dist, check_ori):
    """ORBmatcher::SearchByProjection(Frame&, KeyFrame*, set&, th, ORBdist), src/ORBmatcher.cc:1473-1600 (incl. the level
    predicted by MapPoint::PredictScale, src/MapPoint.cc:409-424)"""
    f = pair
    sc = S.projection_scenario(f["k1"], f["d1"], f["sf"], seed=11)
    u, v, iz = snap(ref, sc)
    obs0 = (sc["obs0"] > 0).astype(np.int32)
    nobs = np.ones(len(sc["valid"]), np.int32)

    def run():
        return oracle.search_by_projection_frame(oracle.Grid(f["k1"], *f["bounds"]), f["d1"], sc["u_right"], obs0, f["sf"], u, v, iz,
                                                 sc["pred_level"], sc["last_angle"], sc["valid"], nobs, sc["mp_desc"], th,
                                                 S.KITTI_BF, 0 | 8, check_ori, orb_dist)
    (n, m, obs), (n2, m2, obs2) = both(oracle, ref, run)
    assert n2 >= 0, "the adapter refused the scenario (%d)" % n2
    assert n > 30 and n == n2 and (m == m2).all() and (obs == obs2).all()


def test_search_by_projection_sim3(oracle, ref, pair):
    """ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th), src/ORBmatcher.cc:290-403, through
    KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:906-945)"""
    f = pair
    sc = S.projection_scenario(f["k1"], f["d1"], f["sf"], seed=11)
    obs0 = (sc["obs0"] > 0).astype(np.int32)
    nobs = np.ones(len(sc["valid"]), np.int32)
    inside = (sc["proj_x"] < f["bounds"][1]) & (sc["proj_y"] < f["bounds"][3])        # KeyFrame::IsInImage is half open
    valid = (sc["valid"] & inside).astype(np.uint8)

    def run():
        return oracle.search_by_projection_frame(oracle.Grid(f["k1"], *f["bounds"]), f["d1"], sc["u_right"], obs0, f["sf"], sc["proj_x"],
                                                 sc["proj_y"], sc["invz"], sc["pred_level"], sc["last_angle"], valid, nobs,
                                                 sc["mp_desc"], 10.0, S.KITTI_BF, 3 | 8, False, 50)
    (n, m, obs), (n2, m2, obs2) = both(oracle, ref, run)
    assert n > 30 and n == n2 and (m == m2).all() and (obs == obs2).all()


@pytest.mark.parametrize("only_stereo,check_ori", [(False, False), (False, True), (True, False)])
def test_search_for_triangulation(oracle, ref, pair, only_stereo, check_ori):
    """ORBmatcher::SearchForTriangulation + CheckDistEpipolarLine, src/ORBmatcher.cc:657-823, 140-157"""
    f = pair
    k1, d1, k2, d2 = f["k1"], f["d1"], f["k2"], f["d2"]
    rng = np.random.default_rng(4)
    ur1 = np.where(rng.random(len(k1)) < 0.4, k1["x"] - 10, -1).astype(np.float32)
    ur2 = np.where(rng.random(len(k2)) < 0.4, k2["x"] - 10, -1).astype(np.float32)
    mp1 = (rng.random(len(k1)) < 0.2).astype(np.uint8)
    mp2 = (rng.random(len(k2)) < 0.2).astype(np.uint8)
    fv1 = S.feature_vector(k1, S.row_band_nodes())
    fv2 = S.feature_vector(k2, S.row_band_nodes(drop_every=5))
    sigma2 = (f["sf"] * f["sf"]).astype(np.float32)
    args = (k1, d1, ur1, mp1, k2, d2, ur2, mp2, fv1, fv2, S.RECTIFIED_F12, 600.0, 180.0, f["sf"], sigma2, only_stereo, check_ori)
    (n, m), (n2, m2) = both(oracle, ref, oracle.search_for_triangulation, *args)
    assert n > (5 if only_stereo else 40) and n == n2 and (m == m2).all()


@pytest.mark.parametrize("window,check_ori,nnratio", [(100, True, 0.9), (30, False, 0.9), (60, True, 0.7)])
def test_search_for_initialization(oracle, ref, pair, window, check_ori, nnratio):
    """ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:405-520, incl. matches stolen by later keypoints"""
    f = pair
    h, w = f["shape"]
    prev = np.stack([f["k1"]["x"], f["k1"]["y"]], 1).astype(np.float32)

    def run():
        return oracle.search_for_initialization(oracle.Grid(f["k2"], 0.0, float(w), 0.0, float(h)), f["d2"], f["k1"], f["d1"], prev,
                                                window, nnratio, check_ori)
    (n, m, p), (n2, m2, p2) = both(oracle, ref, run)
    assert n > 20 and n == n2 and (m == m2).all() and (p.view(np.uint32) == p2.view(np.uint32)).all()


@pytest.mark.parametrize("mode,check_ori,nnratio", [(0, True, 0.7), (0, False, 0.9), (1, True, 0.75), (1, False, 0.6)])
def test_search_by_bow(oracle, ref, pair, mode, check_ori, nnratio):
    """both ORBmatcher::SearchByBoW overloads, src/ORBmatcher.cc:159-288 and :522-655"""
    f = pair
    k1, d1, v1, k2, d2, v2, fv1, fv2 = S.bow_inputs(f["k1"], f["d1"], f["k2"], f["d2"], 5 + mode)
    (n, m), (n2, m2) = both(oracle, ref, oracle.search_by_bow, mode, k1, d1, v1, k2, d2, v2 if mode else None, fv1, fv2, nnratio, check_ori)
    assert n > 50 and n == n2 and (m == m2).all()


def test_distinctive_descriptors(oracle, ref):
    """MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:249-314: least median distance, first row wins ties"""
    desc, ptr = S.distinctive_batch(3, nmp=80)
    (best, med), (best2, med2) = both(oracle, ref, oracle.distinctive_descriptors, desc, ptr)
    counts = np.diff(ptr)
    keep = counts > 0                      # the reference returns early and leaves mDescriptor untouched for 0 observations
    assert keep.sum() >= 79 and (best[keep] == best2[keep]).all() and (med[keep] == med2[keep]).all()


def test_undistort_and_image_bounds(oracle, ref, pair):
    """Frame::UndistortKeyPoints / ComputeImageBounds, src/Frame.cc:584-644"""
    f = pair
    for dist in ([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], [0.0, 0.0, 0.0, 0.0], [-0.2, 0.05, 0.001, -0.0005, 0.01]):
        fx, fy, cx, cy = 458.654, 457.296, 367.215, 248.375
        a, b = both(oracle, ref, oracle.undistort_keypoints, f["k1"], fx, fy, cx, cy, dist)
        assert a.tobytes() == b.tobytes()
        a, b = both(oracle, ref, oracle.compute_image_bounds, 752, 480, fx, fy, cx, cy, dist)
        assert (a.view(np.uint32) == b.view(np.uint32)).all()


@pytest.mark.parametrize("weighting,scoring,levelsup", [(0, 0, 4), (1, 1, 2), (2, 5, 4), (3, 0, 1), (0, 5, 3)])
def test_bow_transform(oracle, ref, weighting, scoring, levelsup):
    """DBoW2 TemplatedVocabulary<FORB>::transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1138-1272), loaded through the
    reference's own loadFromTextFile: BowVector ids and double values bit for bit, FeatureVector, word per feature"""
    parent, desc, weight = S.vocabulary(11, k=5, L=4, unbalanced=False)
    feats = S.vocabulary_features(12, (parent, desc, weight), n=400)

    def run():
        voc = oracle.Vocabulary(5, 4, parent, desc, weight, weighting, scoring)
        return voc.transform(feats, levelsup)
    ((ids, vals), (fvn, fvp, fvi), wo, no), ((ids2, vals2), (fvn2, fvp2, fvi2), wo2, no2) = both(oracle, ref, run)
    assert len(ids) > 20 and (ids == ids2).all() and (vals.view(np.uint64) == vals2.view(np.uint64)).all()
    assert (fvn == fvn2).all() and (fvp == fvp2).all() and (fvi == fvi2).all()
    assert (wo == wo2).all() and (no == no2).all()


def test_bow_transform_unbalanced_tree(oracle, ref):
    """leaves above level L - levelsup: the reference leaves *nid unset there (its FeatureVector node is then undefined,
    DESIGN.md C.8), everything else must agree"""
    parent, desc, weight = S.vocabulary(21, k=6, L=4, unbalanced=True)
    feats = S.vocabulary_features(22, (parent, desc, weight), n=500)

    def run():
        return oracle.Vocabulary(6, 4, parent, desc, weight, 0, 0).transform(feats, 2)
    ((ids, vals), _, wo, _), ((ids2, vals2), _, wo2, _) = both(oracle, ref, run)
    assert (ids == ids2).all() and (vals.view(np.uint64) == vals2.view(np.uint64)).all() and (wo == wo2).all()
