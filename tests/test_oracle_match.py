"""CPU tests of the matcher oracle (pins it against plain numpy) and of the multi-rank host logic."""
import os
import subprocess
import sys

import numpy as np
import pytest

import scenarios as S
from util import CONFIGS, ROOT
from viorb_b200 import sharding, synth


def np_dist(q, m):
    return np.unpackbits(q[:, None, :] ^ m[None, :, :], axis=2).sum(2).astype(np.int32)


def test_descriptor_distance(oracle):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (200, 32)).astype(np.uint8)
    b = rng.integers(0, 256, (200, 32)).astype(np.uint8)
    d = np_dist(a, b)
    for i in range(200):
        assert oracle.descriptor_distance(a[i], b[i]) == d[i, i] == oracle.descriptor_distance(a[i], b[i], popcnt=True)
    assert oracle.descriptor_distance(a[0], a[0]) == 0 and oracle.descriptor_distance(a[0], ~a[0]) == 256


def test_top2_vs_numpy(oracle):
    dmap = synth.descriptor_map(3000, seed=1)
    dmap[100] = dmap[7]            # exact duplicates -> ties
    q = synth.queries_from_map(dmap, 50, seed=2)
    q[0] = dmap[7]
    got = oracle.hamming_top2(q, dmap, nthreads=2)
    d = np_dist(q, dmap)
    order = np.argsort(d, axis=1, kind="stable")
    assert (got["i1"] == order[:, 0]).all() and (got["i2"] == order[:, 1]).all()
    assert (got["d1"] == d[np.arange(50), order[:, 0]]).all() and (got["d2"] == d[np.arange(50), order[:, 1]]).all()
    assert got["i1"][0] == 7 and got["i2"][0] == 100 and got["d2"][0] == 0
    assert (got == oracle.hamming_top2(q, dmap, popcnt=False)).all()
    empty = oracle.hamming_top2(q, dmap[:0])
    assert (empty["i1"] == -1).all() and (empty["d1"] == 256).all()


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_sharded_merge_equals_full_scan(oracle, world):
    dmap = synth.descriptor_map(5000, seed=3)
    dmap[::7] = dmap[0]
    q = synth.queries_from_map(dmap, 40, seed=4)
    full = oracle.hamming_top2(q, dmap)
    parts = np.stack([oracle.hamming_top2(q, dmap[b:e], index_base=b) for b, e in sharding.all_shards(len(dmap), world)])
    assert (oracle.top2_merge(parts) == full).all()


def test_shard_ranges():
    for n in (0, 1, 7, 4096, 10_000_000):
        for w in (1, 2, 4, 8):
            r = sharding.all_shards(n, w)
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))


def test_grid_vs_naive(oracle):
    h, w = 376, 1241
    e = oracle.Extractor(500, 1.2, 8, 20, 7)
    k, d = e(synth.frame(h, w, 3))
    g = oracle.Grid(k, 0.0, float(w), 0.0, float(h))
    rng = np.random.default_rng(5)
    for _ in range(40):
        x, y, r = float(rng.uniform(0, w)), float(rng.uniform(0, h)), float(rng.choice([5.0, 20.0, 50.0]))
        lo, hi = [(-1, -1), (0, 2), (3, -1)][int(rng.integers(0, 3))]
        got = set(g.features_in_area(x, y, r, lo, hi).tolist())
        ok = (np.abs(k["x"] - np.float32(x)) < r) & (np.abs(k["y"] - np.float32(y)) < r)
        if lo > 0 or hi >= 0:
            ok &= k["octave"] >= lo
            if hi >= 0:
                ok &= k["octave"] <= hi
        # the grid only visits cells overlapping the window, which always covers the naive set
        assert got == set(np.nonzero(ok)[0].tolist())


def test_stereo_recovers_synthetic_disparity(oracle):
    h, w, nf, sf, nl, it, mt = CONFIGS["kitti"]
    left, right, disp = synth.stereo_pair(h, w, 7)
    ol, orr = oracle.Extractor(nf, sf, nl, it, mt), oracle.Extractor(nf, sf, nl, it, mt)
    kl, dl = ol(left)
    kr, dr = orr(right)
    mbf, mb = S.KITTI_BF, S.KITTI_BF / S.KITTI_FX
    ur, depth, bd, bi, n = oracle.stereo_match(kl, dl, kr, dr, [ol.pyramid(l) for l in range(nl)],
                                               [orr.pyramid(l) for l in range(nl)], ol.scale_factors(), mbf, mb)
    ok = ur >= 0
    assert n == ok.sum() and n > 300
    band = (kl["y"][ok] * len(disp) / h).astype(int)
    err = np.abs((kl["x"][ok] - ur[ok]) - disp[band])
    assert np.median(err) < 1.0
    assert np.allclose(depth[ok], mbf / (kl["x"][ok] - ur[ok]), rtol=1e-5)


@pytest.mark.parametrize("cfg,seed", [("kitti", 7), ("kitti12", 3)])
def test_stereo_vs_second_restatement(oracle, cfg, seed):
    """orc_stereo_match against tests/stereo_restatement.py, an independent numpy reading of Frame.cc:646-820."""
    import stereo_restatement
    h, w, nf, sf, nl, it, mt = CONFIGS[cfg]
    left, right, _ = synth.stereo_pair(h, w, seed)
    ol, orr = oracle.Extractor(nf, sf, nl, it, mt), oracle.Extractor(nf, sf, nl, it, mt)
    kl, dl = ol(left)
    kr, dr = orr(right)
    mbf, mb = S.KITTI_BF, S.KITTI_BF / S.KITTI_FX
    pl, pr = [ol.pyramid(l) for l in range(nl)], [orr.pyramid(l) for l in range(nl)]
    ur, depth, _, _, n = oracle.stereo_match(kl, dl, kr, dr, pl, pr, ol.scale_factors(), mbf, mb)
    roi = lambda levels: [p[19:-19, 19:-19] for p in levels]        # mvImagePyramid[l] is the ROI of the padded level
    ur2, depth2 = stereo_restatement.compute_stereo_matches(kl, dl, kr, dr, roi(pl), roi(pr), ol.scale_factors(), mbf, mb)
    assert n > 100
    assert (ur.view(np.uint32) == ur2.view(np.uint32)).all()
    assert (depth.view(np.uint32) == depth2.view(np.uint32)).all()


@pytest.fixture(scope="module")
def kitti_frame(oracle):
    h, w, nf, sf, nl, it, mt = CONFIGS["kitti"]
    left, _, _ = synth.stereo_pair(h, w, 7)
    ex = oracle.Extractor(nf, sf, nl, it, mt)
    k, d = ex(left)
    return dict(k=k, d=d, sf=ex.scale_factors(), bounds=(0.0, float(w), 0.0, float(h)), shape=(h, w))


def test_grid_vs_second_restatement(oracle, kitti_frame):
    import search_restatement as R
    f = kitti_frame
    g, g2 = oracle.Grid(f["k"], *f["bounds"]), R.Grid(f["k"], *f["bounds"])
    rng = np.random.default_rng(5)
    h, w = f["shape"]
    for _ in range(80):
        x, y = float(rng.uniform(-20, w + 20)), float(rng.uniform(-20, h + 20))
        r = float(rng.choice([3.0, 10.0, 25.0, 60.0]))
        lo, hi = [(-1, -1), (0, 3), (2, -1), (1, 2), (4, 4)][int(rng.integers(0, 5))]
        assert g.features_in_area(x, y, r, lo, hi).tolist() == g2.features_in_area(x, y, r, lo, hi)


@pytest.mark.parametrize("th,nnratio", [(1.0, 0.8), (3.0, 0.8), (5.0, 0.6)])
def test_search_by_projection_local_vs_second_restatement(oracle, kitti_frame, th, nnratio):
    """orc_search_by_projection_local against tests/search_restatement.py (ORBmatcher.cc:45-137 read a second time)."""
    import search_restatement as R
    f = kitti_frame
    sc = S.projection_scenario(f["k"], f["d"], f["sf"], seed=11)
    args = (f["d"], sc["u_right"], sc["obs0"], f["sf"], sc["proj_x"], sc["proj_y"], sc["proj_xr"], sc["pred_level"],
            sc["view_cos"], sc["valid"], sc["nobs"], sc["mp_desc"], th, nnratio)
    n, m, obs = oracle.search_by_projection_local(oracle.Grid(f["k"], *f["bounds"]), *args)
    n2, m2, obs2 = R.search_by_projection_local(R.Grid(f["k"], *f["bounds"]), *args)
    assert n > 50 and n == n2 and (m == m2).all() and (obs == obs2).all()


@pytest.mark.parametrize("mode,th,check_ori", [(0, 15.0, True), (1, 7.0, True), (2, 7.0, False), (0, 7.0, False)])
def test_search_by_projection_frame_vs_second_restatement(oracle, kitti_frame, mode, th, check_ori):
    """orc_search_by_projection_frame (modes 0/1/2 = neither / forward / backward) against
    tests/search_restatement.py (ORBmatcher.cc:1378-1468 and ComputeThreeMaxima read a second time)."""
    import search_restatement as R
    f = kitti_frame
    sc = S.projection_scenario(f["k"], f["d"], f["sf"], seed=11)
    n, m, obs = oracle.search_by_projection_frame(
        oracle.Grid(f["k"], *f["bounds"]), f["d"], sc["u_right"], sc["obs0"], f["sf"], sc["proj_x"], sc["proj_y"],
        sc["invz"], sc["last_octave"], sc["last_angle"], sc["valid"], sc["nobs"], sc["mp_desc"], th, S.KITTI_BF, mode,
        check_ori, 100)
    n2, m2, obs2 = R.search_by_projection_frame(
        R.Grid(f["k"], *f["bounds"]), f["d"], sc["u_right"], sc["obs0"], f["sf"], sc["proj_x"], sc["proj_y"], sc["invz"],
        sc["last_octave"], sc["last_angle"], sc["valid"], sc["nobs"], sc["mp_desc"], th, S.KITTI_BF, mode == 1, mode == 2,
        check_ori)
    assert n > 30 and n == n2 and (m == m2).all() and (obs == obs2).all()


@pytest.mark.parametrize("only_stereo,check_ori", [(False, False), (False, True), (True, False)])
def test_search_for_triangulation_vs_second_restatement(oracle, only_stereo, check_ori):
    """orc_search_for_triangulation against tests/search_restatement.py (ORBmatcher.cc:140-157, 657-823 read again)."""
    import search_restatement as R
    h, w, nf, sf_, nl, it, mt = CONFIGS["kitti"]
    left, right, _ = synth.stereo_pair(h, w, 7)
    e1, e2 = oracle.Extractor(nf, sf_, nl, it, mt), oracle.Extractor(nf, sf_, nl, it, mt)
    k1, d1 = e1(left)
    k2, d2 = e2(right)
    rng = np.random.default_rng(4)
    ur1 = np.where(rng.random(len(k1)) < 0.4, k1["x"] - 10, -1).astype(np.float32)
    ur2 = np.where(rng.random(len(k2)) < 0.4, k2["x"] - 10, -1).astype(np.float32)
    mp1 = (rng.random(len(k1)) < 0.2).astype(np.uint8)
    mp2 = (rng.random(len(k2)) < 0.2).astype(np.uint8)
    fv1 = S.feature_vector(k1, S.row_band_nodes())
    fv2 = S.feature_vector(k2, S.row_band_nodes(drop_every=5))
    sf = e1.scale_factors()
    sigma2 = (sf * sf).astype(np.float32)
    args = (k1, d1, ur1, mp1, k2, d2, ur2, mp2, fv1, fv2, S.RECTIFIED_F12, 600.0, 180.0, sf, sigma2, only_stereo, check_ori)
    n, m = oracle.search_for_triangulation(*args)
    n2, m2 = R.search_for_triangulation(*args)
    assert n > (5 if only_stereo else 40) and n == n2 and (m == m2).all()


@pytest.mark.parametrize("window,check_ori,nnratio", [(100, True, 0.9), (30, False, 0.9), (60, True, 0.7)])
def test_search_for_initialization_vs_second_restatement(oracle, window, check_ori, nnratio):
    """orc_search_for_initialization against tests/search_restatement.py (ORBmatcher.cc:405-520 read again), including
    the stolen-match dependence between consecutive keypoints."""
    import search_restatement as R
    h, w, nf, sf_, nl, it, mt = CONFIGS["kitti"]
    left, right, _ = synth.stereo_pair(h, w, 7)
    e1, e2 = oracle.Extractor(nf, sf_, nl, it, mt), oracle.Extractor(nf, sf_, nl, it, mt)
    k1, d1 = e1(left)
    k2, d2 = e2(right)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    n, m, p = oracle.search_for_initialization(oracle.Grid(k2, 0.0, float(w), 0.0, float(h)), d2, k1, d1, prev, window,
                                               nnratio, check_ori)
    n2, m2, p2 = R.search_for_initialization(R.Grid(k2, 0.0, float(w), 0.0, float(h)), d2, k1, d1, prev, window, nnratio,
                                             check_ori)
    assert n > 20 and n == n2 and (m == m2).all() and (p.view(np.uint32) == p2.view(np.uint32)).all()


@pytest.mark.parametrize("mode,check_ori,nnratio", [(0, True, 0.7), (0, False, 0.9), (1, True, 0.75), (1, False, 0.6)])
def test_search_by_bow_vs_second_restatement(oracle, mode, check_ori, nnratio):
    """orc_search_by_bow against tests/search_restatement.py (ORBmatcher.cc:159-288 and :522-655 read again)."""
    import search_restatement as R
    h, w, nf, sf_, nl, it, mt = CONFIGS["kitti"]
    left, right, _ = synth.stereo_pair(h, w, 7)
    e1, e2 = oracle.Extractor(nf, sf_, nl, it, mt), oracle.Extractor(nf, sf_, nl, it, mt)
    k1, d1 = e1(left)
    k2, d2 = e2(right)
    k1, d1, v1, k2, d2, v2, fv1, fv2 = S.bow_inputs(k1, d1, k2, d2, 5 + mode)
    n, m = oracle.search_by_bow(mode, k1, d1, v1, k2, d2, v2 if mode else None, fv1, fv2, nnratio, check_ori)
    n2, m2 = R.search_by_bow(mode, k1, d1, v1, k2, d2, v2 if mode else None, fv1, fv2, nnratio, check_ori)
    assert n > 50 and n == n2 and (m == m2).all()


def test_keyframe_projection_overloads_vs_second_restatement(oracle, kitti_frame):
    """relocalisation (ORBmatcher.cc:1473-1600) and Sim3 (:290-403) overloads: the oracle's flag combinations
    (mode 0|8 and 3|8 of orc_search_by_projection_frame) against the overloads read a second time."""
    import search_restatement as R
    f = kitti_frame
    sc = S.projection_scenario(f["k"], f["d"], f["sf"], seed=11)
    obs0 = (sc["obs0"] > 0).astype(np.int32)
    nobs = np.ones(len(sc["valid"]), np.int32)
    g, g2 = oracle.Grid(f["k"], *f["bounds"]), R.Grid(f["k"], *f["bounds"])
    for th, orb_dist, check_ori in ((10.0, 100, True), (3.0, 64, True), (10.0, 100, False)):
        n, m, obs = oracle.search_by_projection_frame(g, f["d"], sc["u_right"], obs0, f["sf"], sc["proj_x"], sc["proj_y"],
                                                      sc["invz"], sc["pred_level"], sc["last_angle"], sc["valid"], nobs,
                                                      sc["mp_desc"], th, S.KITTI_BF, 0 | 8, check_ori, orb_dist)
        n2, m2, a2 = R.search_by_projection_reloc(g2, f["d"], obs0, f["sf"], sc["proj_x"], sc["proj_y"], sc["pred_level"],
                                                  sc["last_angle"], sc["valid"], sc["mp_desc"], th, orb_dist, check_ori)
        assert n > 30 and n == n2 and (m == m2).all() and (obs == a2).all()
    n, m, obs = oracle.search_by_projection_frame(g, f["d"], sc["u_right"], obs0, f["sf"], sc["proj_x"], sc["proj_y"],
                                                  sc["invz"], sc["pred_level"], sc["last_angle"], sc["valid"], nobs,
                                                  sc["mp_desc"], 10.0, S.KITTI_BF, 3 | 8, False, 50)
    n2, m2, a2 = R.search_by_projection_sim3(g2, f["d"], obs0, f["sf"], sc["proj_x"], sc["proj_y"], sc["pred_level"],
                                             sc["valid"], sc["mp_desc"], 10)
    assert n > 30 and n == n2 and (m == m2).all() and (obs == a2).all()


def test_fuse_loop_and_sim3_vs_second_restatement(oracle):
    """orc_search_window_top1 (the Fuse candidate loops, with and without the chi-square gates) and orc_search_by_sim3
    against tests/search_restatement.py (ORBmatcher.cc:883-943, :1043-1073, :1193-1320 read again)."""
    import search_restatement as R
    h, w, nf, sf_, nl, it, mt = CONFIGS["kitti"]
    left, right, _ = synth.stereo_pair(h, w, 7)
    e1, e2 = oracle.Extractor(nf, sf_, nl, it, mt), oracle.Extractor(nf, sf_, nl, it, mt)
    kl, dl = e1(left)
    kr, dr = e2(right)
    sf = e1.scale_factors()
    inv_sigma2 = (1.0 / (sf * sf)).astype(np.float32)
    b = (0.0, float(w), 0.0, float(h))
    for gates, th in ((True, 3.0), (False, 4.0), (False, 10.0)):
        rng = np.random.default_rng(2)
        u_right = np.where(rng.random(len(kr)) < 0.5, kr["x"] - rng.uniform(2, 40, len(kr)), -1).astype(np.float32)
        u, v, level, valid, mpd = S.window_queries(kl, dl, kr, 3, 20.0)
        ur = (u - rng.uniform(2, 40, len(u))).astype(np.float32) if gates else None
        args = (dr, u_right, sf, u, v, ur, level, valid, mpd, th, 50, inv_sigma2 if gates else None)
        bi, bd = oracle.search_window_top1(oracle.Grid(kr, *b), *args)
        bi2, bd2 = R.search_window_top1(R.Grid(kr, *b), *args)
        assert (bi >= 0).sum() > (5 if gates else 30)
        assert (bi == bi2).all() and (bd[bi >= 0] == bd2[bi >= 0]).all()
    q12 = S.window_queries(kl, dl, kr, 5, 20.0)
    q21 = S.window_queries(kr, dr, kl, 6, -20.0)
    n, m = oracle.search_by_sim3(oracle.Grid(kl, *b), dl, sf, oracle.Grid(kr, *b), dr, sf, q12, q21, 25.0)
    n2, m2 = R.search_by_sim3(R.Grid(kl, *b), dl, sf, R.Grid(kr, *b), dr, sf, q12, q21, 25.0)
    assert n > 20 and n == n2 and (m == m2).all()


def test_projection_scenario_is_meaningful(oracle):
    h, w, nf, sf, nl, it, mt = CONFIGS["kitti"]
    e = oracle.Extractor(nf, sf, nl, it, mt)
    k, d = e(synth.frame(h, w, 9))
    sc = S.projection_scenario(k, d, e.scale_factors(), seed=11)
    g = oracle.Grid(k, 0.0, float(w), 0.0, float(h))
    n, match, obs = oracle.search_by_projection_local(g, d, sc["u_right"], sc["obs0"], e.scale_factors(), sc["proj_x"],
                                                      sc["proj_y"], sc["proj_xr"], sc["pred_level"], sc["view_cos"],
                                                      sc["valid"], sc["nobs"], sc["mp_desc"], 3.0, 0.8)
    assert n > 50 and (match >= 0).sum() <= n
    assert ((obs > 0) >= (sc["obs0"] > 0)).all()


def test_two_rank_gloo_top2_merge():
    """world_size-2 gloo run of the sharded matcher host logic (CPU stand-in shard scan = the oracle)"""
    script = os.path.join(ROOT, "tests", "gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", script],
                       env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "RANK0 OK" in p.stdout and "RANK1 OK" in p.stdout


def test_distinctive_descriptor_vs_numpy(oracle):
    """MapPoint.cc:249-314 restatement against an independent numpy statement (sort each row, take [int(0.5*(N-1))],
    first least median)."""
    desc, ptr = S.distinctive_batch(3, nmp=60)
    best, med = oracle.distinctive_descriptors(desc, ptr)
    for p in range(len(ptr) - 1):
        rows = desc[ptr[p]:ptr[p + 1]]
        if len(rows) == 0:
            assert best[p] == -1
            continue
        D = np.sort(np_dist(rows, rows), axis=1)
        m = D[:, int(0.5 * (len(rows) - 1))]
        assert best[p] == int(np.argmin(m)) and med[p] == m.min()


def _np_transform(voc, feats, L, levelsup, weighting, scoring):
    """independent statement of TemplatedVocabulary::transform with python dicts and IEEE doubles"""
    parent, desc, weight = voc
    n = len(parent)
    kids = [[] for _ in range(n)]
    for i in range(1, n):
        kids[parent[i]].append(i)
    wid, nw = {}, 0
    for i in range(1, n):
        if not kids[i]:
            wid[i] = nw
            nw += 1
    bow, fv, words, nodes = {}, {}, [], []
    for i, f in enumerate(feats):
        node, level, nid = 0, 0, 0
        while kids[node]:
            level += 1
            d = np.unpackbits(desc[kids[node]] ^ f[None], axis=1).sum(1)
            node = kids[node][int(np.argmin(d))]
            if level == L - levelsup:
                nid = node
        w = float(weight[node])
        words.append(wid[node]); nodes.append(nid)
        if w > 0:
            if weighting in (0, 1):
                bow[wid[node]] = bow[wid[node]] + w if wid[node] in bow else w
            else:
                bow.setdefault(wid[node], w)
            fv.setdefault(nid, []).append(i)
    ids = sorted(bow)
    vals = [bow[i] for i in ids]
    must = scoring != 5
    if weighting in (0, 1) and vals and not must:
        vals = [v / float(len(vals)) for v in vals]
    if must:
        norm = 0.0
        for v in vals:
            norm = norm + (abs(v) if scoring != 1 else v * v)
        if scoring == 1:
            norm = float(np.sqrt(np.float64(norm)))
        if norm > 0:
            vals = [v / norm for v in vals]
    return ids, vals, fv, words, nodes


@pytest.mark.parametrize("weighting,scoring", [(0, 0), (1, 1), (2, 5), (3, 0), (0, 5)])
def test_bow_transform_vs_numpy(oracle, weighting, scoring):
    """bow_oracle.cpp (DBoW2 TemplatedVocabulary.h:1138-1272 restated) against an independent python statement"""
    voc = S.vocabulary(11, k=5, L=4)
    feats = S.vocabulary_features(12, voc, n=300)
    V = oracle.Vocabulary(5, 4, *voc, weighting=weighting, scoring=scoring)
    (ids, vals), (fvn, fvp, fvi), words, nodes = V.transform(feats, levelsup=2)
    rids, rvals, rfv, rwords, rnodes = _np_transform(voc, feats, 4, 2, weighting, scoring)
    assert list(ids) == rids and list(words) == rwords and list(nodes) == rnodes
    assert np.array_equal(np.asarray(rvals, np.float64).view(np.uint64), vals.view(np.uint64))
    assert list(fvn) == sorted(rfv)
    for j, nid in enumerate(fvn):
        assert list(fvi[fvp[j]:fvp[j + 1]]) == rfv[nid]
