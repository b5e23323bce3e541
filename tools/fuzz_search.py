#!/usr/bin/env python3
"""Differential fuzz of the windowed / per-frame matchers: a fresh synthetic stereo pair and fresh random scenarios per
seed; stereo matcher, the projection searches (all modes), triangulation, SearchByBoW x2, SearchForInitialization,
SearchBySim3, the Fuse search loop, distinctive descriptors and the BoW transform against the CPU oracle.
usage: tools/fuzz_search.py [seeds] [first seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenarios as S  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from viorb_b200 import api, synth  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def main():
    nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    ctx = api.Context(0)
    bad = []

    def check(name, seed, ok):
        if not ok:
            bad.append((name, seed))
            print("MISMATCH", name, "seed", seed)

    for seed in range(first, first + nseeds):
        rng = np.random.default_rng(seed)
        h, w = int(rng.integers(300, 500)), int(rng.integers(900, 1300))
        nf = int(rng.integers(800, 2500))
        left, right, disp = synth.stereo_pair(h, w, seed)
        exl, exr = api.ORBextractor(nf, 1.2, 8, 20, 7, ctx=ctx), api.ORBextractor(nf, 1.2, 8, 20, 7, ctx=ctx)
        kl, dl = exl(left)
        kr, dr = exr(right)
        ol, orr = O.Extractor(nf, 1.2, 8, 20, 7), O.Extractor(nf, 1.2, 8, 20, 7)
        kl2, dl2 = ol(left)
        kr2, dr2 = orr(right)
        check("extract", seed, kl.tobytes() == kl2.tobytes() and (dl == dl2).all() and kr.tobytes() == kr2.tobytes())
        sf = ol.scale_factors()
        bounds = (0.0, float(w), 0.0, float(h))
        mbf, mb = S.KITTI_BF, S.KITTI_BF / S.KITTI_FX
        # stereo
        ur, depth = api.ComputeStereoMatches(exl, exr, kl, dl, kr, dr, mbf, mb)
        pl, pr = [ol.pyramid(l) for l in range(8)], [orr.pyramid(l) for l in range(8)]
        ur_ref, depth_ref, _, _, _ = O.stereo_match(kl, dl, kr, dr, pl, pr, sf, mbf, mb)
        check("stereo", seed, (bits(ur) == bits(ur_ref)).all() and (bits(depth) == bits(depth_ref)).all())
        # projection searches
        sc = S.projection_scenario(kl, dl, sf, seed=seed, n_mp=int(rng.integers(100, 900)), conflicts=int(rng.integers(0, 50)))
        fi = api.FrameIndex(ctx, kl, dl, sc["u_right"], bounds, sf)
        g = O.Grid(kl, *bounds)
        for th, nn in ((1.0, 0.8), (float(rng.uniform(2, 8)), float(rng.uniform(0.5, 0.95)))):
            n_ref, m_ref, obs_ref = O.search_by_projection_local(g, dl, sc["u_right"], sc["obs0"], sf, sc["proj_x"], sc["proj_y"],
                                                                 sc["proj_xr"], sc["pred_level"], sc["view_cos"], sc["valid"],
                                                                 sc["nobs"], sc["mp_desc"], th, nn)
            n, match, obs = api.ORBmatcher(nn, True, ctx=ctx).SearchByProjectionLocal(
                fi, sc["obs0"], sc["proj_x"], sc["proj_y"], sc["proj_xr"], sc["pred_level"], sc["view_cos"], sc["valid"], sc["nobs"],
                sc["mp_desc"], th)
            check("proj_local", seed, n == n_ref and (match == m_ref).all() and (obs == obs_ref).all())
        for mode in (0, 1, 2, 0 | 8, 3 | 8):
            th, co, thh = float(rng.uniform(3, 15)), bool(rng.integers(0, 2)), int(rng.choice([50, 64, 100]))
            n_ref, m_ref, obs_ref = O.search_by_projection_frame(g, dl, sc["u_right"], sc["obs0"], sf, sc["proj_x"], sc["proj_y"],
                                                                 sc["invz"], sc["last_octave"], sc["last_angle"], sc["valid"],
                                                                 sc["nobs"], sc["mp_desc"], th, mbf, mode, co, thh)
            n, match, obs = api.ORBmatcher(0.9, co, ctx=ctx).SearchByProjectionFrame(
                fi, sc["obs0"], sc["proj_x"], sc["proj_y"], sc["invz"], sc["last_octave"], sc["last_angle"], sc["valid"], sc["nobs"],
                sc["mp_desc"], th, mbf, mode, thh)
            check("proj_frame_%d" % mode, seed, n == n_ref and (match == m_ref).all() and (obs == obs_ref).all())
        # window top-1 (Fuse) with and without gates, Sim3
        n1 = len(kl)
        u = (kl["x"] + rng.normal(0, 1.5, n1)).astype(np.float32)
        v = (kl["y"] + rng.normal(0, 1.0, n1)).astype(np.float32)
        lvl = np.clip(kl["octave"] + rng.integers(0, 2, n1), 0, 7).astype(np.int32)
        valid = (rng.random(n1) < 0.85).astype(np.uint8)
        mpd = S.flip_bits(dl, rng, 20)
        urq = (u - rng.uniform(2, 40, n1)).astype(np.float32)
        inv = (1.0 / (sf * sf)).astype(np.float32)
        for gates in (True, False):
            rb, rd = O.search_window_top1(g, dl, sc["u_right"], sf, u, v, urq if gates else None, lvl, valid, mpd, 4.0, 50,
                                          inv if gates else None)
            bi, bd = api.ORBmatcher(ctx=ctx).SearchWindowTop1(fi, u, v, urq if gates else None, lvl, valid, mpd, 4.0, 50,
                                                             inv if gates else None)
            check("window_%d" % gates, seed, (bi == rb).all() and (bd[rb >= 0] == rd[rb >= 0]).all())
        fi2 = api.FrameIndex(ctx, kr, dr, None, bounds, sf)
        g2 = O.Grid(kr, *bounds)
        dmean = float(np.mean(disp))
        q12 = (np.asarray(kl["x"] - dmean, np.float32), kl["y"].copy(), lvl, valid, mpd)
        n2 = len(kr)
        q21 = (np.asarray(kr["x"] + dmean, np.float32), kr["y"].copy(), np.clip(kr["octave"], 0, 7).astype(np.int32),
               (rng.random(n2) < 0.9).astype(np.uint8), S.flip_bits(dr, rng, 20))
        n_ref, m_ref = O.search_by_sim3(g, dl, sf, g2, dr, sf, q12, q21, 30.0)
        n, m12 = api.ORBmatcher(ctx=ctx).SearchBySim3(fi, fi2, q12, q21, 30.0)
        check("sim3", seed, n == n_ref and (m12 == m_ref).all())
        # feature-vector searches
        band = int(rng.integers(12, 40))
        fv1 = S.feature_vector(kl, S.row_band_nodes(band=band))
        fv2 = S.feature_vector(kr, S.row_band_nodes(band=band, drop_every=5))
        v1 = (rng.random(n1) < 0.8).astype(np.uint8)
        v2 = (rng.random(n2) < 0.8).astype(np.uint8)
        for mode in (0, 1):
            nn, co = float(rng.uniform(0.6, 0.95)), bool(rng.integers(0, 2))
            n_ref, m_ref = O.search_by_bow(mode, kl, dl, v1, kr, dr, v2 if mode else None, fv1, fv2, nn, co)
            n, mt = api.ORBmatcher(nn, co, ctx=ctx).SearchByBoW(mode, kl, dl, v1, kr, dr, v2 if mode else None, fv1, fv2)
            check("bow_%d" % mode, seed, n == n_ref and (mt == m_ref).all())
        ur1 = np.where(rng.random(n1) < 0.4, kl["x"] - 10, -1).astype(np.float32)
        ur2 = np.where(rng.random(n2) < 0.4, kr["x"] - 10, -1).astype(np.float32)
        mp1, mp2 = (rng.random(n1) < 0.2).astype(np.uint8), (rng.random(n2) < 0.2).astype(np.uint8)
        sig2 = (sf * sf).astype(np.float32)
        for only, co in ((False, True), (True, False)):
            n_ref, m_ref = O.search_for_triangulation(kl, dl, ur1, mp1, kr, dr, ur2, mp2, fv1, fv2, S.RECTIFIED_F12, 600.0, 180.0, sf,
                                                      sig2, only, co)
            n, m12 = api.ORBmatcher(0.6, co, ctx=ctx).SearchForTriangulation(kl, dl, ur1, mp1, kr, dr, ur2, mp2, fv1, fv2,
                                                                             S.RECTIFIED_F12, 600.0, 180.0, sf, sig2, only)
            check("triangulation", seed, n == n_ref and (m12 == m_ref).all())
        prev = np.stack([kl["x"], kl["y"]], 1).astype(np.float32)
        win = int(rng.choice([20, 50, 100]))
        n_ref, m_ref, p_ref = O.search_for_initialization(g2, dr, kl, dl, prev, win, 0.9, True)
        n, m12, pp = api.ORBmatcher(0.9, True, ctx=ctx).SearchForInitialization(fi2, kl, dl, prev, win)
        check("init", seed, n == n_ref and (m12 == m_ref).all() and (bits(pp) == bits(p_ref)).all())
        # distinctive descriptors, BoW transform
        dd, ptr = S.distinctive_batch(seed, nmp=120)
        b1, m1 = api.ORBmatcher(ctx=ctx).ComputeDistinctiveDescriptors(dd, ptr)
        rb1, rm1 = O.distinctive_descriptors(dd, ptr)
        check("distinctive", seed, (b1 == rb1).all() and (m1 == rm1).all())
        k, Lv = int(rng.integers(3, 11)), int(rng.integers(2, 6))
        voc = S.vocabulary(seed, k=k, L=Lv)
        wgt, scr, lup = int(rng.integers(0, 4)), int(rng.integers(0, 6)), int(rng.integers(0, Lv + 1))
        V = api.ORBVocabulary(k, Lv, *voc, weighting=wgt, scoring=scr, ctx=ctx)
        R = O.Vocabulary(k, Lv, *voc, weighting=wgt, scoring=scr)
        (ids, vals), (fvn, fvp, fvi), wd, nd = V.transform(dl, lup)
        (rids, rvals), (rfvn, rfvp, rfvi), rwd, rnd = R.transform(dl, lup)
        check("bow_transform", seed, np.array_equal(ids, rids) and np.array_equal(vals.view(np.uint64), rvals.view(np.uint64)) and
              np.array_equal(fvn, rfvn) and np.array_equal(fvp, rfvp) and np.array_equal(fvi, rfvi) and np.array_equal(wd, rwd) and
              np.array_equal(nd, rnd))
        V.close(); fi.close(); fi2.close(); exl.close(); exr.close()
    print("fuzz_search: %d seeds, %d mismatches" % (nseeds, len(bad)))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
