"""GPU parity tests: stereo matcher, frame grid and the windowed Hamming searches vs the CPU oracle.
Everything is compared bit-for-bit (float outputs as raw bits, indices and counts exactly)."""
import numpy as np
import pytest

import scenarios as S
from util import CONFIGS
from viorb_b200 import synth

pytestmark = pytest.mark.gpu


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


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def stereo(api, ctx, oracle):
    """KITTI-shape synthetic pair (BASELINE config 2) extracted by both implementations"""
    h, w, nf, sf, nl, it, mt = CONFIGS["kitti"]
    left, right, disp = synth.stereo_pair(h, w, 7)
    exl, exr = api.ORBextractor(nf, sf, nl, it, mt, ctx=ctx), api.ORBextractor(nf, sf, nl, it, mt, ctx=ctx)
    kl, dl = exl(left)
    kr, dr = exr(right)
    ol, orr = oracle.Extractor(nf, sf, nl, it, mt), oracle.Extractor(nf, sf, nl, it, mt)
    kl2, dl2 = ol(left)
    kr2, dr2 = orr(right)
    assert (kl == kl2).all() and (dl == dl2).all() and (kr == kr2).all() and (dr == dr2).all()
    yield dict(exl=exl, exr=exr, kl=kl, dl=dl, kr=kr, dr=dr, ol=ol, orr=orr, disp=disp, shape=(h, w), nl=nl)
    exl.close()
    exr.close()


def test_compute_stereo_matches(api, oracle, stereo):
    s = stereo
    mbf, mb = S.KITTI_BF, S.KITTI_BF / S.KITTI_FX
    ur, depth = api.ComputeStereoMatches(s["exl"], s["exr"], s["kl"], s["dl"], s["kr"], s["dr"], mbf, mb)
    pl = [s["ol"].pyramid(l) for l in range(s["nl"])]
    pr = [s["orr"].pyramid(l) for l in range(s["nl"])]
    ur_ref, depth_ref, bd, bi, n = oracle.stereo_match(s["kl"], s["dl"], s["kr"], s["dr"], pl, pr,
                                                       s["ol"].scale_factors(), mbf, mb)
    assert n > 300                                        # the synthetic pair really has stereo matches
    assert (bits(ur) == bits(ur_ref)).all()
    assert (bits(depth) == bits(depth_ref)).all()
    # sanity: recovered disparities are the synthetic band disparities
    ok = ur >= 0
    band = (s["kl"]["y"][ok] * len(s["disp"]) / s["shape"][0]).astype(int)
    err = np.abs((s["kl"]["x"][ok] - ur[ok]) - s["disp"][band])
    assert np.median(err) < 1.0


def test_stereo_edge_cases(api, stereo):
    s = stereo
    mbf, mb = S.KITTI_BF, S.KITTI_BF / S.KITTI_FX
    ur, depth = api.ComputeStereoMatches(s["exl"], s["exr"], s["kl"][:0], s["dl"][:0], s["kr"], s["dr"], mbf, mb)
    assert len(ur) == 0
    ur, depth = api.ComputeStereoMatches(s["exl"], s["exr"], s["kl"], s["dl"], s["kr"][:0], s["dr"][:0], mbf, mb)
    assert (ur == -1).all() and (depth == -1).all()


@pytest.fixture(scope="module")
def frame(api, ctx, stereo):
    s = stereo
    h, w = s["shape"]
    sf = s["ol"].scale_factors()
    sc = S.projection_scenario(s["kl"], s["dl"], sf, seed=11)
    fi = api.FrameIndex(ctx, s["kl"], s["dl"], sc["u_right"], (0.0, float(w), 0.0, float(h)), sf)
    yield dict(fi=fi, sc=sc, sf=sf, bounds=(0.0, float(w), 0.0, float(h)))
    fi.close()


def test_get_features_in_area(oracle, stereo, frame):
    g = oracle.Grid(stereo["kl"], *frame["bounds"])
    rng = np.random.default_rng(1)
    h, w = stereo["shape"]
    for _ in range(60):
        x, y = float(rng.uniform(-20, w + 20)), float(rng.uniform(-20, h + 20))
        r = float(rng.choice([3.0, 10.0, 25.0, 60.0]))
        lo, hi = [(-1, -1), (0, 3), (2, -1), (1, 2), (4, 4)][int(rng.integers(0, 5))]
        got = frame["fi"].GetFeaturesInArea(x, y, r, lo, hi)
        ref = g.features_in_area(x, y, r, lo, hi)
        assert len(got) == len(ref) and (got == ref).all()       # same indices in the reference's order


@pytest.mark.parametrize("th,nnratio", [(1.0, 0.8), (3.0, 0.8), (5.0, 0.6)])
def test_search_by_projection_local(api, ctx, oracle, stereo, frame, th, nnratio):
    s, sc = stereo, frame["sc"]
    g = oracle.Grid(s["kl"], *frame["bounds"])
    n_ref, m_ref, obs_ref = oracle.search_by_projection_local(
        g, s["dl"], sc["u_right"], sc["obs0"], frame["sf"], sc["proj_x"], sc["proj_y"], sc["proj_xr"], sc["pred_level"],
        sc["view_cos"], sc["valid"], sc["nobs"], sc["mp_desc"], th, nnratio)
    m = api.ORBmatcher(nnratio, True, ctx=ctx)
    n, match, obs = m.SearchByProjectionLocal(frame["fi"], sc["obs0"], sc["proj_x"], sc["proj_y"], sc["proj_xr"],
                                              sc["pred_level"], sc["view_cos"], sc["valid"], sc["nobs"], sc["mp_desc"], th)
    assert n_ref > 50
    assert n == n_ref and (match == m_ref).all() and (obs == obs_ref).all()


@pytest.mark.parametrize("n_mp", [2500, 6000])
def test_search_by_projection_large_local_map(api, ctx, oracle, stereo, n_mp):
    """Tracking::SearchLocalPoints passes local maps of several thousand points (the other scenarios stay below 1000): the
    scratch arena must hold seven per-point arrays and the descriptors, and the ordered resolve must still equal the
    reference's sequential loop when three map points compete for every keypoint"""
    s = stereo
    h, w = s["shape"]
    sf = s["ol"].scale_factors()
    sc = S.projection_scenario(s["kl"], s["dl"], sf, seed=23, n_mp=n_mp, conflicts=n_mp // 10)
    bounds = (0.0, float(w), 0.0, float(h))
    fi = api.FrameIndex(ctx, s["kl"], s["dl"], sc["u_right"], bounds, sf)
    g = oracle.Grid(s["kl"], *bounds)
    n_ref, m_ref, obs_ref = oracle.search_by_projection_local(
        g, s["dl"], sc["u_right"], sc["obs0"], sf, sc["proj_x"], sc["proj_y"], sc["proj_xr"], sc["pred_level"],
        sc["view_cos"], sc["valid"], sc["nobs"], sc["mp_desc"], 3.0, 0.8)
    m = api.ORBmatcher(0.8, True, ctx=ctx)
    n, match, obs = m.SearchByProjectionLocal(fi, sc["obs0"], sc["proj_x"], sc["proj_y"], sc["proj_xr"], sc["pred_level"],
                                              sc["view_cos"], sc["valid"], sc["nobs"], sc["mp_desc"], 3.0)
    assert n_ref > 500
    assert n == n_ref and (match == m_ref).all() and (obs == obs_ref).all()
    n2, match2, _ = m.SearchByProjectionFrame(fi, sc["obs0"], sc["proj_x"], sc["proj_y"], sc["invz"], sc["last_octave"],
                                              sc["last_angle"], sc["valid"], sc["nobs"], sc["mp_desc"], 7.0, S.KITTI_BF, 0, 100)[:3]
    n2_ref, m2_ref = oracle.search_by_projection_frame(g, s["dl"], sc["u_right"], sc["obs0"], sf, sc["proj_x"], sc["proj_y"], sc["invz"],
                                                       sc["last_octave"], sc["last_angle"], sc["valid"], sc["nobs"], sc["mp_desc"], 7.0,
                                                       S.KITTI_BF, 0, True, 100)[:2]
    assert n2 == n2_ref and (match2 == m2_ref).all()
    fi.close()


@pytest.mark.parametrize("mode,th,check_ori,th_high", [(0, 15.0, True, 100), (1, 7.0, True, 100), (2, 7.0, False, 100),
                                                       (0, 10.0, True, 64), (0 | 8, 10.0, True, 64), (0 | 8, 3.0, True, 100),
                                                       (3 | 8, 10.0, False, 50)])
def test_search_by_projection_frame(api, ctx, oracle, stereo, frame, mode, th, check_ori, th_high):
    s, sc = stereo, frame["sc"]
    g = oracle.Grid(s["kl"], *frame["bounds"])
    n_ref, m_ref, obs_ref = oracle.search_by_projection_frame(
        g, s["dl"], sc["u_right"], sc["obs0"], frame["sf"], sc["proj_x"], sc["proj_y"], sc["invz"], sc["last_octave"],
        sc["last_angle"], sc["valid"], sc["nobs"], sc["mp_desc"], th, S.KITTI_BF, mode, check_ori, th_high)
    m = api.ORBmatcher(0.9, check_ori, ctx=ctx)
    n, match, obs = m.SearchByProjectionFrame(frame["fi"], sc["obs0"], sc["proj_x"], sc["proj_y"], sc["invz"],
                                              sc["last_octave"], sc["last_angle"], sc["valid"], sc["nobs"], sc["mp_desc"],
                                              th, S.KITTI_BF, mode, th_high)
    assert n_ref > 30
    assert n == n_ref and (match == m_ref).all() and (obs == obs_ref).all()


@pytest.mark.parametrize("mode,th,check_ori,th_high", [(0 | 8, 10.0, True, 100), (3 | 8, 10.0, False, 50)])
def test_search_by_projection_keyframe_overloads(api, ctx, oracle, stereo, frame, mode, th, check_ori, th_high):
    """relocalisation (ORBmatcher.cc:1473-1600) and Sim3 (:290-403) overloads: any assigned keypoint blocks,
    every match blocks later map points (nobs = 1), no stereo gate"""
    s, sc = stereo, frame["sc"]
    g = oracle.Grid(s["kl"], *frame["bounds"])
    obs0 = (sc["obs0"] > 0).astype(np.int32)
    nobs = np.ones(len(sc["valid"]), np.int32)
    n_ref, m_ref, obs_ref = oracle.search_by_projection_frame(
        g, s["dl"], sc["u_right"], obs0, frame["sf"], sc["proj_x"], sc["proj_y"], sc["invz"], sc["pred_level"],
        sc["last_angle"], sc["valid"], nobs, sc["mp_desc"], th, S.KITTI_BF, mode, check_ori, th_high)
    m = api.ORBmatcher(0.75, check_ori, ctx=ctx)
    n, match, obs = m.SearchByProjectionFrame(frame["fi"], obs0, sc["proj_x"], sc["proj_y"], sc["invz"], sc["pred_level"],
                                              sc["last_angle"], sc["valid"], nobs, sc["mp_desc"], th, S.KITTI_BF, mode, th_high)
    assert n_ref > 30
    assert n == n_ref and (match == m_ref).all() and (obs == obs_ref).all()
    # every keypoint is claimed at most once and previously assigned keypoints are never taken
    assert (match[obs0 > 0] == -1).all()


def test_search_sequential_dependence(api, ctx, oracle, stereo, frame):
    """a chain of map points that all want the same keypoint: the fixed-point iteration must reproduce
    the sequential 'already taken' skips (ORBmatcher.cc:87-89,123) exactly"""
    s = stereo
    k, d = s["kl"], s["dl"]
    g = oracle.Grid(k, *frame["bounds"])
    n = 64
    target = int(np.argmax(k["octave"] == 0))
    sc = dict(proj_x=np.full(n, k["x"][target], np.float32), proj_y=np.full(n, k["y"][target], np.float32),
              proj_xr=np.zeros(n, np.float32), pred_level=np.zeros(n, np.int32), view_cos=np.ones(n, np.float32),
              valid=np.ones(n, np.uint8), nobs=np.arange(n, dtype=np.int32) % 3, mp_desc=np.repeat(d[target:target + 1], n, 0))
    ur = np.full(len(k), -1, np.float32)
    obs0 = np.zeros(len(k), np.int32)
    fi = api.FrameIndex(ctx, k, d, ur, frame["bounds"], frame["sf"])
    n_ref, m_ref, obs_ref = oracle.search_by_projection_local(g, d, ur, obs0, frame["sf"], sc["proj_x"], sc["proj_y"],
                                                              sc["proj_xr"], sc["pred_level"], sc["view_cos"], sc["valid"],
                                                              sc["nobs"], sc["mp_desc"], 5.0, 0.8)
    m = api.ORBmatcher(0.8, True, ctx=ctx)
    n_gpu, match, obs = m.SearchByProjectionLocal(fi, obs0, sc["proj_x"], sc["proj_y"], sc["proj_xr"], sc["pred_level"],
                                                  sc["view_cos"], sc["valid"], sc["nobs"], sc["mp_desc"], 5.0)
    assert n_gpu == n_ref and (match == m_ref).all() and (obs == obs_ref).all()
    fi.close()


@pytest.mark.parametrize("only_stereo,check_ori", [(False, False), (False, True), (True, False)])
def test_search_for_triangulation(api, ctx, oracle, stereo, only_stereo, check_ori):
    s = stereo
    rng = np.random.default_rng(4)
    k1, d1, k2, d2 = s["kl"], s["dl"], s["kr"], s["dr"]
    ur1 = np.where(rng.random(len(k1)) < 0.4, k1["x"] - 10, -1).astype(np.float32)
    ur2 = np.where(rng.random(len(k2)) < 0.4, k2["x"] - 10, -1).astype(np.float32)
    mp1 = (rng.random(len(k1)) < 0.2).astype(np.uint8)
    mp2 = (rng.random(len(k2)) < 0.2).astype(np.uint8)
    fv1 = S.feature_vector(k1, S.row_band_nodes())
    fv2 = S.feature_vector(k2, S.row_band_nodes(drop_every=5))
    sf = s["ol"].scale_factors()
    sigma2 = (sf * sf).astype(np.float32)
    ex, ey = 600.0, 180.0            # an epipole inside the image so the exclusion gate is exercised
    n_ref, m_ref = oracle.search_for_triangulation(k1, d1, ur1, mp1, k2, d2, ur2, mp2, fv1, fv2, S.RECTIFIED_F12, ex, ey,
                                                   sf, sigma2, only_stereo, check_ori)
    m = api.ORBmatcher(0.6, check_ori, ctx=ctx)
    n, m12 = m.SearchForTriangulation(k1, d1, ur1, mp1, k2, d2, ur2, mp2, fv1, fv2, S.RECTIFIED_F12, ex, ey, sf, sigma2,
                                      only_stereo)
    assert n_ref > (5 if only_stereo else 40)
    assert n == n_ref and (m12 == m_ref).all()


def test_distinctive_descriptors(api, ctx, oracle):
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:249-314) batched: BestIdx and BestMedian identical"""
    m = api.ORBmatcher(ctx=ctx)
    for seed in (0, 1):
        desc, ptr = S.distinctive_batch(seed)
        best, med = m.ComputeDistinctiveDescriptors(desc, ptr)
        rb, rm = oracle.distinctive_descriptors(desc, ptr)
        assert (best == rb).all() and (med == rm).all()
    best, med = m.ComputeDistinctiveDescriptors(np.zeros((0, 32), np.uint8), np.zeros(4, np.int32))
    assert (best == -1).all()


@pytest.mark.parametrize("weighting,scoring,k,L,levelsup", [(0, 0, 10, 4, 2), (1, 1, 6, 5, 4), (2, 5, 6, 4, 1), (3, 0, 6, 4, 2),
                                                             (0, 5, 6, 4, 6)])
def test_bow_transform(api, ctx, oracle, weighting, scoring, k, L, levelsup):
    """ORBVocabulary::transform (DBoW2 TemplatedVocabulary.h:1138-1272): BowVector ids and values (doubles, bit for
    bit), FeatureVector node ids and index lists, per-feature word and node identical to the oracle"""
    voc = S.vocabulary(21 + k, k=k, L=L)
    V = api.ORBVocabulary(k, L, *voc, weighting=weighting, scoring=scoring, ctx=ctx)
    R = oracle.Vocabulary(k, L, *voc, weighting=weighting, scoring=scoring)
    for n in (1, 37, 1500):
        feats = S.vocabulary_features(n, voc, n=n)
        (ids, vals), (fvn, fvp, fvi), words, nodes = V.transform(feats, levelsup)
        (rids, rvals), (rfvn, rfvp, rfvi), rwords, rnodes = R.transform(feats, levelsup)
        assert np.array_equal(words, rwords) and np.array_equal(nodes, rnodes)
        assert np.array_equal(ids, rids) and np.array_equal(vals.view(np.uint64), rvals.view(np.uint64))
        assert np.array_equal(fvn, rfvn) and np.array_equal(fvp, rfvp) and np.array_equal(fvi, rfvi)
    (ids, vals), (fvn, fvp, fvi), _, _ = V.transform(np.zeros((0, 32), np.uint8), levelsup)
    assert len(ids) == 0 and len(fvn) == 0
    V.close()


def _bow_inputs(s, seed):
    return S.bow_inputs(s["kl"], s["dl"], s["kr"], s["dr"], seed)


@pytest.mark.parametrize("mode,check_ori,nnratio", [(0, True, 0.7), (0, False, 0.9), (1, True, 0.75), (1, False, 0.6)])
def test_search_by_bow(api, ctx, oracle, stereo, mode, check_ori, nnratio):
    """ORBmatcher::SearchByBoW, both overloads (ORBmatcher.cc:159-288, :522-655)"""
    k1, d1, v1, k2, d2, v2, fv1, fv2 = _bow_inputs(stereo, 5 + mode)
    n_ref, m_ref = oracle.search_by_bow(mode, k1, d1, v1, k2, d2, v2 if mode else None, fv1, fv2, nnratio, check_ori)
    m = api.ORBmatcher(nnratio, check_ori, ctx=ctx)
    n, match = m.SearchByBoW(mode, k1, d1, v1, k2, d2, v2 if mode else None, fv1, fv2)
    assert n_ref > 50
    assert n == n_ref and (match == m_ref).all()


@pytest.mark.parametrize("window,check_ori,nnratio", [(100, True, 0.9), (30, False, 0.9), (100, True, 0.6)])
def test_search_for_initialization(api, ctx, oracle, stereo, window, check_ori, nnratio):
    """ORBmatcher::SearchForInitialization (ORBmatcher.cc:405-520) incl. the stolen-match dependence"""
    s = stereo
    h, w = s["shape"]
    sf = s["ol"].scale_factors()
    k1, d1, k2, d2 = s["kl"], s["dl"], s["kr"], s["dr"]
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    g2 = oracle.Grid(k2, 0.0, float(w), 0.0, float(h))
    n_ref, m_ref, p_ref = oracle.search_for_initialization(g2, d2, k1, d1, prev, window, nnratio, check_ori)
    fi2 = api.FrameIndex(ctx, k2, d2, None, (0.0, float(w), 0.0, float(h)), sf)
    m = api.ORBmatcher(nnratio, check_ori, ctx=ctx)
    n, m12, p = m.SearchForInitialization(fi2, k1, d1, prev, window)
    fi2.close()
    assert n_ref > 20
    assert n == n_ref and (m12 == m_ref).all() and (p.view(np.uint32) == p_ref.view(np.uint32)).all()


@pytest.mark.parametrize("K,dist,size", [((458.654, 457.296, 367.215, 248.375), [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], (752, 480)),
                                         ((517.306408, 516.469215, 318.643040, 255.313989), [0.262383, -0.953104, -0.005358, 0.002628, 1.163314], (640, 480)),
                                         ((718.856, 718.856, 607.1928, 185.2157), [0.0, 0.0, 0.0, 0.0], (1241, 376))])
def test_undistort_and_grid_on_device(api, ctx, oracle, stereo, K, dist, size):
    """Frame::UndistortKeyPoints, ComputeImageBounds and AssignFeaturesToGrid on the device (Frame.cc:584-645, 410-425):
    mvKeysUn and the bounds bit-equal to the oracle (itself pinned on cv2.undistortPoints), and the grid built from
    the raw keypoints answers GetFeaturesInArea like the oracle grid of the undistorted ones"""
    s = stereo
    rng = np.random.default_rng(9)
    kps = s["kl"].copy()
    kps["x"] = (kps["x"] * (size[0] / 1241.0)).astype(np.float32)
    kps["y"] = (kps["y"] * (size[1] / 376.0)).astype(np.float32)
    ref = oracle.undistort_keypoints(kps, *K, dist)
    got = api.UndistortKeyPoints(ctx, kps, K, dist)
    assert got.tobytes() == ref.tobytes()
    bref = oracle.compute_image_bounds(size[0], size[1], *K, dist)
    b = api.ComputeImageBounds(ctx, size[0], size[1], K, dist)
    assert (b.view(np.uint32) == bref.view(np.uint32)).all()
    sf = s["ol"].scale_factors()
    fi = api.FrameIndex.from_distorted(ctx, kps, s["dl"], None, K, dist, size, sf)
    k_un, bounds = fi.keys()
    assert k_un.tobytes() == ref.tobytes() and (bounds.view(np.uint32) == bref.view(np.uint32)).all()
    g = oracle.Grid(ref, *[float(v) for v in bref])
    for _ in range(40):
        x, y = float(rng.uniform(-20, size[0] + 20)), float(rng.uniform(-20, size[1] + 20))
        r = float(rng.choice([5.0, 20.0, 60.0]))
        a, c = fi.GetFeaturesInArea(x, y, r), g.features_in_area(x, y, r)
        assert len(a) == len(c) and (a == c).all()
    fi.close()


def _window_queries(k_src, d_src, k_dst, seed, disp):
    return S.window_queries(k_src, d_src, k_dst, seed, disp)


@pytest.mark.parametrize("gates,th", [(True, 3.0), (False, 4.0), (False, 10.0)])
def test_search_window_top1(api, ctx, oracle, stereo, gates, th):
    """search loop of both ORBmatcher::Fuse overloads (ORBmatcher.cc:883-943 with the chi-square gates, :1043-1073)"""
    s = stereo
    h, w = s["shape"]
    sf = s["ol"].scale_factors()
    inv_sigma2 = (1.0 / (sf * sf)).astype(np.float32)
    rng = np.random.default_rng(2)
    kr, dr, kl, dl = s["kr"], s["dr"], s["kl"], s["dl"]
    u_right = np.where(rng.random(len(kr)) < 0.5, kr["x"] - rng.uniform(2, 40, len(kr)), -1).astype(np.float32)
    u, v, level, valid, mpd = _window_queries(kl, dl, kr, 3, 20.0)
    ur = (u - rng.uniform(2, 40, len(u))).astype(np.float32) if gates else None
    g = oracle.Grid(kr, 0.0, float(w), 0.0, float(h))
    rb, rd = oracle.search_window_top1(g, dr, u_right, sf, u, v, ur, level, valid, mpd, th, 50, inv_sigma2 if gates else None)
    fi = api.FrameIndex(ctx, kr, dr, u_right, (0.0, float(w), 0.0, float(h)), sf)
    m = api.ORBmatcher(ctx=ctx)
    bi, bd = m.SearchWindowTop1(fi, u, v, ur, level, valid, mpd, th, 50, inv_sigma2 if gates else None)
    fi.close()
    assert (rb >= 0).sum() > (5 if gates else 30)
    assert (bi == rb).all() and (bd[rb >= 0] == rd[rb >= 0]).all()


def test_search_by_sim3(api, ctx, oracle, stereo):
    """ORBmatcher::SearchBySim3 (ORBmatcher.cc:1102-1326): both directional searches and the agreement check"""
    s = stereo
    h, w = s["shape"]
    sf = s["ol"].scale_factors()
    kl, dl, kr, dr = s["kl"], s["dl"], s["kr"], s["dr"]
    q12 = _window_queries(kl, dl, kr, 5, 20.0)
    q21 = _window_queries(kr, dr, kl, 6, -20.0)
    g1, g2 = oracle.Grid(kl, 0.0, float(w), 0.0, float(h)), oracle.Grid(kr, 0.0, float(w), 0.0, float(h))
    n_ref, m_ref = oracle.search_by_sim3(g1, dl, sf, g2, dr, sf, q12, q21, 25.0)
    f1 = api.FrameIndex(ctx, kl, dl, None, (0.0, float(w), 0.0, float(h)), sf)
    f2 = api.FrameIndex(ctx, kr, dr, None, (0.0, float(w), 0.0, float(h)), sf)
    n, m12 = api.ORBmatcher(ctx=ctx).SearchBySim3(f1, f2, q12, q21, 25.0)
    f1.close(); f2.close()
    assert n_ref > 20
    assert n == n_ref and (m12 == m_ref).all()


def test_matchers_differential_fuzz():
    """tools/fuzz_search.py: fresh stereo pairs and random scenarios per seed through every matcher entry point"""
    import os
    import subprocess
    import sys
    from util import ROOT
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_search.py"), "8", "500"], capture_output=True, text=True,
                       cwd=ROOT, timeout=900)
    assert p.returncode == 0 and " 0 mismatches" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
