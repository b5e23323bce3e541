"""Second, independent numpy/float32 restatement of the windowed searches of the hot path, statement by statement
from the reference (line numbers inline).  Test infrastructure: it pins oracle/match_oracle.cpp the way
tests/cv2_restatement.py pins the extractor oracle.

  Grid                       Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea   src/Frame.cc:410-425,507-572
  search_by_projection_local ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) src/ORBmatcher.cc:45-137
  search_by_projection_frame ORBmatcher::SearchByProjection(Cur, Last, th, bMono)          src/ORBmatcher.cc:1378-1468
  compute_three_maxima       ORBmatcher::ComputeThreeMaxima                                 src/ORBmatcher.cc:1602-1643
  search_for_triangulation   ORBmatcher::SearchForTriangulation + CheckDistEpipolarLine     src/ORBmatcher.cc:140-157,657-823
  search_for_initialization  ORBmatcher::SearchForInitialization                            src/ORBmatcher.cc:405-520
  search_by_bow              ORBmatcher::SearchByBoW, both overloads                        src/ORBmatcher.cc:159-288,522-655
  search_by_projection_reloc / _sim3   the two KeyFrame overloads of SearchByProjection     src/ORBmatcher.cc:1473-1600, 290-403
  search_window_top1         candidate loop of Fuse x2 and SearchBySim3                     src/ORBmatcher.cc:883-943,1043-1073,1193-1226
  search_by_sim3             ORBmatcher::SearchBySim3 (searches + agreement)                src/ORBmatcher.cc:1193-1320

The map-point graph is flattened the way the oracle's interface does it: obs[i] > 0 stands for
"F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0", nobs[k] is Observations() of map point k.
"""
import math

import numpy as np

F = np.float32
GRID_COLS, GRID_ROWS = 64, 48                 # include/Frame.h:41-42
TH_HIGH, HISTO_LENGTH = 100, 30               # src/ORBmatcher.cc:37-39


def _roundf(x):
    x = float(x)
    return int(math.floor(x + 0.5) if x >= 0 else math.ceil(x - 0.5))


def _hamming(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


class Grid:
    def __init__(self, kps_un, min_x, max_x, min_y, max_y):
        self.kps = kps_un
        self.min_x, self.max_x, self.min_y, self.max_y = F(min_x), F(max_x), F(min_y), F(max_y)
        self.inv_w = F(F(GRID_COLS) / F(self.max_x - self.min_x))        # Frame.cc (ctor): mfGridElementWidthInv
        self.inv_h = F(F(GRID_ROWS) / F(self.max_y - self.min_y))
        self.cells = [[[] for _ in range(GRID_ROWS)] for _ in range(GRID_COLS)]
        for i in range(len(kps_un)):                                     # :417-424
            px = _roundf(F(F(kps_un["x"][i] - self.min_x) * self.inv_w))  # :564-565
            py = _roundf(F(F(kps_un["y"][i] - self.min_y) * self.inv_h))
            if px < 0 or px >= GRID_COLS or py < 0 or py >= GRID_ROWS:   # :568
                continue
            self.cells[px][py].append(i)

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        x, y, r = F(x), F(y), F(r)
        out = []
        c0 = max(0, int(math.floor(F(F(F(x - self.min_x) - r) * self.inv_w))))       # :512
        if c0 >= GRID_COLS:
            return out
        c1 = min(GRID_COLS - 1, int(math.ceil(F(F(F(x - self.min_x) + r) * self.inv_w))))
        if c1 < 0:
            return out
        r0 = max(0, int(math.floor(F(F(F(y - self.min_y) - r) * self.inv_h))))
        if r0 >= GRID_ROWS:
            return out
        r1 = min(GRID_ROWS - 1, int(math.ceil(F(F(F(y - self.min_y) + r) * self.inv_h))))
        if r1 < 0:
            return out
        check = (min_level > 0) or (max_level >= 0)                                   # :528
        for ix in range(c0, c1 + 1):
            for iy in range(r0, r1 + 1):
                for j in self.cells[ix][iy]:
                    o = int(self.kps["octave"][j])
                    if check:
                        if o < min_level:
                            continue
                        if max_level >= 0 and o > max_level:
                            continue
                    dx = F(self.kps["x"][j] - x)
                    dy = F(self.kps["y"][j] - y)
                    if abs(dx) < r and abs(dy) < r:                                   # :552
                        out.append(j)
        return out


def search_by_projection_local(grid, fdesc, u_right, obs0, scale_factors, proj_x, proj_y, proj_xr, pred_level,
                               view_cos, valid, nobs, mp_desc, th, nnratio):
    kps = grid.kps
    sf = np.asarray(scale_factors, np.float32)
    obs = np.array(obs0, np.int32)
    match = np.full(len(kps), -1, np.int32)
    th, nnratio = F(th), F(nnratio)
    factor = float(th) != 1.0                                            # :49
    n = 0
    for k in range(len(mp_desc)):
        if not valid[k]:                                                 # :54-58
            continue
        level = int(pred_level[k])
        r = F(2.5) if float(view_cos[k]) > 0.998 else F(4.0)             # :131-137 (double literal vs float)
        if factor:
            r = F(r * th)
        rs = F(r * sf[level])
        idx = grid.features_in_area(proj_x[k], proj_y[k], rs, level - 1, level)       # :68-69
        if not idx:
            continue
        best, best_lvl, best2, best_lvl2, best_i = 256, -1, 256, -1, -1
        for i in idx:
            if obs[i] > 0:                                               # :87-89
                continue
            if u_right[i] > 0:                                           # :91-96
                er = abs(F(F(proj_xr[k]) - F(u_right[i])))
                if er > rs:
                    continue
            d = _hamming(mp_desc[k], fdesc[i])
            if d < best:                                                 # :102-114
                best2, best = best, d
                best_lvl2, best_lvl = best_lvl, int(kps["octave"][i])
                best_i = i
            elif d < best2:
                best_lvl2 = int(kps["octave"][i])
                best2 = d
        if best <= TH_HIGH:                                              # :118-124
            if best_lvl == best_lvl2 and F(best) > F(nnratio * F(best2)):
                continue
            match[best_i] = k
            obs[best_i] = nobs[k]
            n += 1
    return n, match, obs


def compute_three_maxima(histo):
    max1 = max2 = max3 = 0
    ind1 = ind2 = ind3 = -1
    for i, h in enumerate(histo):
        s = len(h)
        if s > max1:
            max3, max2, max1 = max2, max1, s
            ind3, ind2, ind1 = ind2, ind1, i
        elif s > max2:
            max3, max2 = max2, s
            ind3, ind2 = ind2, i
        elif s > max3:
            max3, ind3 = s, i
    if F(max2) < F(F(0.1) * F(max1)):                                    # :1632-1641
        ind2 = ind3 = -1
    elif F(max3) < F(F(0.1) * F(max1)):
        ind3 = -1
    return ind1, ind2, ind3


def search_by_projection_frame(grid, fdesc, u_right, obs0, scale_factors, u, v, invz, last_octave, last_angle, valid,
                               nobs, mp_desc, th, mbf, forward, backward, check_ori):
    kps = grid.kps
    sf = np.asarray(scale_factors, np.float32)
    obs = np.array(obs0, np.int32)
    match = np.full(len(kps), -1, np.int32)
    th, mbf = F(th), F(mbf)
    hist = [[] for _ in range(HISTO_LENGTH)]
    factor = F(F(1.0) / F(HISTO_LENGTH))                                 # :1336
    n = 0
    for i in range(len(mp_desc)):
        if not valid[i]:                                                 # pMP && !mvbOutlier && invzc >= 0 (:1356-1369)
            continue
        ui, vi, iz = F(u[i]), F(v[i]), F(invz[i])
        if ui < grid.min_x or ui > grid.max_x:                           # :1374-1377
            continue
        if vi < grid.min_y or vi > grid.max_y:
            continue
        lo = int(last_octave[i])                                         # :1379
        radius = F(th * sf[lo])                                          # :1382
        if forward:                                                      # :1386-1391
            idx = grid.features_in_area(ui, vi, radius, lo)
        elif backward:
            idx = grid.features_in_area(ui, vi, radius, 0, lo)
        else:
            idx = grid.features_in_area(ui, vi, radius, lo - 1, lo + 1)
        if not idx:
            continue
        best, best_i = 256, -1
        for i2 in idx:                                                   # :1402-1426
            if obs[i2] > 0:
                continue
            if u_right[i2] > 0:
                ur = F(ui - F(mbf * iz))
                er = abs(F(ur - F(u_right[i2])))
                if er > radius:
                    continue
            d = _hamming(mp_desc[i], fdesc[i2])
            if d < best:
                best, best_i = d, i2
        if best <= TH_HIGH:                                              # :1428-1445
            match[best_i] = i
            obs[best_i] = nobs[i]
            n += 1
            if check_ori:
                rot = F(F(last_angle[i]) - F(kps["angle"][best_i]))
                if rot < 0.0:
                    rot = F(rot + F(360.0))
                b = _roundf(F(rot * factor))
                if b == HISTO_LENGTH:
                    b = 0
                hist[b].append(best_i)
    if check_ori:                                                        # :1449-1467
        keep = compute_three_maxima(hist)
        for b in range(HISTO_LENGTH):
            if b not in keep:
                for j in hist[b]:
                    match[j] = -1
                    obs[j] = 0
                    n -= 1
    return n, match, obs


def check_dist_epipolar_line(kp1, kp2, F12, sigma2_2):
    """ORBmatcher::CheckDistEpipolarLine, src/ORBmatcher.cc:140-157; F12[r][c] = F12.at<float>(r, c)."""
    x1, y1, x2, y2 = F(kp1["x"]), F(kp1["y"]), F(kp2["x"]), F(kp2["y"])
    a = F(F(F(x1 * F12[0][0]) + F(y1 * F12[1][0])) + F12[2][0])
    b = F(F(F(x1 * F12[0][1]) + F(y1 * F12[1][1])) + F12[2][1])
    c = F(F(F(x1 * F12[0][2]) + F(y1 * F12[1][2])) + F12[2][2])
    num = F(F(F(a * x2) + F(b * y2)) + c)
    den = F(F(a * a) + F(b * b))
    if den == 0:
        return False
    dsqr = F(F(num * num) / den)
    return float(dsqr) < 3.84 * float(sigma2_2[int(kp2["octave"])])      # double comparison


def search_for_triangulation(k1, d1, ur1, has_mp1, k2, d2, ur2, has_mp2, fv1, fv2, F12, ex, ey, scale2, sigma2_2,
                             only_stereo, check_ori):
    """ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:657-823 (epipole :667-670 computed by the caller).
    fv = (node_ids ascending, node_ptr, idx): the std::map<node, vector<idx>> of DBoW2::FeatureVector."""
    TH_LOW = 50
    F12 = np.asarray(F12, np.float32).reshape(3, 3)
    ex, ey = F(ex), F(ey)
    m12 = np.full(len(k1), -1, np.int32)
    matched2 = np.zeros(len(k2), bool)                                   # never set by the reference (:681)
    hist = [[] for _ in range(HISTO_LENGTH)]
    factor = F(F(1.0) / F(HISTO_LENGTH))
    ids1, ptr1, idx1v = fv1
    ids2, ptr2, idx2v = fv2
    n, i1, i2 = 0, 0, 0
    while i1 < len(ids1) and i2 < len(ids2):                             # :697
        if ids1[i1] == ids2[i2]:
            for a in idx1v[ptr1[i1]:ptr1[i1 + 1]]:                       # :701
                if has_mp1[a]:                                           # :707-709
                    continue
                stereo1 = ur1[a] >= 0
                if only_stereo and not stereo1:
                    continue
                best, best2 = TH_LOW, -1
                for b in idx2v[ptr2[i2]:ptr2[i2 + 1]]:                   # :724
                    if matched2[b] or has_mp2[b]:
                        continue
                    stereo2 = ur2[b] >= 0
                    if only_stereo and not stereo2:
                        continue
                    dist = _hamming(d1[a], d2[b])
                    if dist > TH_LOW or dist > best:                     # :744 (an equal distance replaces the best)
                        continue
                    if not stereo1 and not stereo2:                      # :749-755
                        dx = F(ex - F(k2["x"][b]))
                        dy = F(ey - F(k2["y"][b]))
                        if F(F(dx * dx) + F(dy * dy)) < F(F(100) * F(scale2[int(k2["octave"][b])])):
                            continue
                    if check_dist_epipolar_line(k1[a], k2[b], F12, sigma2_2):
                        best2, best = int(b), dist
                if best2 >= 0:                                           # :764-781
                    m12[a] = best2
                    n += 1
                    if check_ori:
                        rot = F(F(k1["angle"][a]) - F(k2["angle"][best2]))
                        if rot < 0.0:
                            rot = F(rot + F(360.0))
                        bn = _roundf(F(rot * factor))
                        if bn == HISTO_LENGTH:
                            bn = 0
                        hist[bn].append(int(a))
            i1 += 1
            i2 += 1
        elif ids1[i1] < ids2[i2]:                                        # lower_bound (:787-793)
            while i1 < len(ids1) and ids1[i1] < ids2[i2]:
                i1 += 1
        else:
            while i2 < len(ids2) and ids2[i2] < ids1[i1]:
                i2 += 1
    if check_ori:                                                        # :796-814
        keep = compute_three_maxima(hist)
        for bn in range(HISTO_LENGTH):
            if bn in keep:
                continue
            for a in hist[bn]:
                m12[a] = -1
                n -= 1
    return n, m12


def search_for_initialization(grid2, d2, k1, d1, prev_matched, window, nnratio, check_ori):
    """ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:405-520.  grid2 = frame 2 (keypoints + grid),
    prev_matched = vbPrevMatched as an (N1, 2) float array; returns (nmatches, vnMatches12, updated vbPrevMatched)."""
    TH_LOW, INT_MAX = 50, 2147483647
    k2 = grid2.kps
    nnratio = F(nnratio)
    prev = np.array(prev_matched, np.float32)
    m12 = np.full(len(k1), -1, np.int32)
    hist = [[] for _ in range(HISTO_LENGTH)]
    factor = F(F(1.0) / F(HISTO_LENGTH))
    matched_dist = [INT_MAX] * len(k2)                                   # :416-417
    m21 = [-1] * len(k2)
    n = 0
    for i1 in range(len(k1)):
        level1 = int(k1["octave"][i1])
        if level1 > 0:                                                   # :423-424
            continue
        idx = grid2.features_in_area(prev[i1][0], prev[i1][1], F(window), level1, level1)   # :426
        if not idx:
            continue
        best, best2, best_i = INT_MAX, INT_MAX, -1
        for i2 in idx:                                                   # :437-458
            dist = _hamming(d1[i1], d2[i2])
            if matched_dist[i2] <= dist:
                continue
            if dist < best:
                best2, best, best_i = best, dist, i2
            elif dist < best2:
                best2 = dist
        if best <= TH_LOW:                                               # :460-486
            if F(best) < F(F(best2) * nnratio):
                if m21[best_i] >= 0:
                    m12[m21[best_i]] = -1
                    n -= 1
                m12[i1] = best_i
                m21[best_i] = i1
                matched_dist[best_i] = best
                n += 1
                if check_ori:
                    rot = F(F(k1["angle"][i1]) - F(k2["angle"][best_i]))
                    if rot < 0.0:
                        rot = F(rot + F(360.0))
                    b = _roundf(F(rot * factor))
                    if b == HISTO_LENGTH:
                        b = 0
                    hist[b].append(i1)
    if check_ori:                                                        # :490-512
        keep = compute_three_maxima(hist)
        for b in range(HISTO_LENGTH):
            if b in keep:
                continue
            for i1 in hist[b]:
                if m12[i1] >= 0:
                    m12[i1] = -1
                    n -= 1
    for i1 in range(len(k1)):                                            # :515-517
        if m12[i1] >= 0:
            prev[i1][0] = k2["x"][m12[i1]]
            prev[i1][1] = k2["y"][m12[i1]]
    return n, m12, prev


def search_by_bow(mode, k1, d1, valid1, k2, d2, valid2, fv1, fv2, nnratio, check_ori):
    """ORBmatcher::SearchByBoW.  mode 0 = (KeyFrame* pKF, Frame& F), src/ORBmatcher.cc:159-288: set 1 is the key
    frame (valid1[i] = map point present and not bad), set 2 the frame; returns per frame keypoint the key-frame
    keypoint whose map point it received.  mode 1 = (KeyFrame*, KeyFrame*), :522-655: valid2 likewise; returns per
    keypoint of key frame 1 the matched keypoint of key frame 2."""
    TH_LOW = 50
    nnratio = F(nnratio)
    match = np.full(len(k2) if mode == 0 else len(k1), -1, np.int32)
    matched2 = np.zeros(len(k2), bool)
    hist = [[] for _ in range(HISTO_LENGTH)]
    factor = F(F(1.0) / F(HISTO_LENGTH))
    ids1, ptr1, idx1v = fv1
    ids2, ptr2, idx2v = fv2
    n, i1, i2 = 0, 0, 0
    while i1 < len(ids1) and i2 < len(ids2):
        if ids1[i1] == ids2[i2]:
            for a in idx1v[ptr1[i1]:ptr1[i1 + 1]]:
                if not valid1[a]:                                        # :194-200 / :566-573
                    continue
                best1, best2, best_i = 256, 256, -1
                for b in idx2v[ptr2[i2]:ptr2[i2 + 1]]:
                    if mode == 0:
                        if match[b] >= 0:                                # :212-213 vpMapPointMatches[realIdxF]
                            continue
                    else:
                        if matched2[b] or not valid2[b]:                 # :587-591
                            continue
                    dist = _hamming(d1[a], d2[b])
                    if dist < best1:
                        best2, best1, best_i = best1, dist, int(b)
                    elif dist < best2:
                        best2 = dist
                ok = best1 <= TH_LOW if mode == 0 else best1 < TH_LOW     # :232 vs :610
                if ok and F(best1) < F(nnratio * F(best2)):
                    rot = F(F(k1["angle"][a]) - F(k2["angle"][best_i]))
                    if mode == 0:
                        match[best_i] = a
                    else:
                        match[a] = best_i
                        matched2[best_i] = True
                    if check_ori:
                        if rot < 0.0:
                            rot = F(rot + F(360.0))
                        bn = _roundf(F(rot * factor))
                        if bn == HISTO_LENGTH:
                            bn = 0
                        hist[bn].append(best_i if mode == 0 else int(a))
                    n += 1
            i1 += 1
            i2 += 1
        elif ids1[i1] < ids2[i2]:
            while i1 < len(ids1) and ids1[i1] < ids2[i2]:
                i1 += 1
        else:
            while i2 < len(ids2) and ids2[i2] < ids1[i1]:
                i2 += 1
    if check_ori:
        keep = compute_three_maxima(hist)
        for bn in range(HISTO_LENGTH):
            if bn in keep:
                continue
            for j in hist[bn]:
                match[j] = -1
                n -= 1
    return n, match


def search_by_projection_reloc(grid, fdesc, assigned0, scale_factors, u, v, pred_level, kf_angle, valid, mp_desc, th,
                               orb_dist, check_ori):
    """ORBmatcher::SearchByProjection(Frame&, KeyFrame*, set<MapPoint*>&, th, ORBdist), src/ORBmatcher.cc:1473-1600,
    from the image-bounds test (:1505) on; the projection, the distance gate and PredictScale are the caller's
    (valid / pred_level).  assigned0[i] != 0 <=> CurrentFrame.mvpMapPoints[i] != NULL."""
    kps = grid.kps
    sf = np.asarray(scale_factors, np.float32)
    assigned = np.array(assigned0, np.int32)
    match = np.full(len(kps), -1, np.int32)
    th = F(th)
    hist = [[] for _ in range(HISTO_LENGTH)]
    factor = F(F(1.0) / F(HISTO_LENGTH))
    n = 0
    for i in range(len(mp_desc)):
        if not valid[i]:
            continue
        ui, vi = F(u[i]), F(v[i])
        if ui < grid.min_x or ui > grid.max_x or vi < grid.min_y or vi > grid.max_y:       # :1505-1508
            continue
        level = int(pred_level[i])
        radius = F(th * sf[level])                                                          # :1525
        idx = grid.features_in_area(ui, vi, radius, level - 1, level + 1)                   # :1527
        best, best_i = 256, -1
        for i2 in idx:
            if assigned[i2]:                                                                # :1540-1541
                continue
            d = _hamming(mp_desc[i], fdesc[i2])
            if d < best:
                best, best_i = d, i2
        if best <= orb_dist:                                                                # :1554
            match[best_i] = i
            assigned[best_i] = 1
            n += 1
            if check_ori:
                rot = F(F(kf_angle[i]) - F(kps["angle"][best_i]))
                if rot < 0.0:
                    rot = F(rot + F(360.0))
                b = _roundf(F(rot * factor))
                if b == HISTO_LENGTH:
                    b = 0
                hist[b].append(best_i)
    if check_ori:
        keep = compute_three_maxima(hist)
        for b in range(HISTO_LENGTH):
            if b not in keep:
                for j in hist[b]:
                    match[j] = -1
                    assigned[j] = 0
                    n -= 1
    return n, match, assigned


def search_by_projection_sim3(grid, fdesc, assigned0, scale_factors, u, v, pred_level, valid, mp_desc, th):
    """ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th), src/ORBmatcher.cc:290-403, from the
    search radius (:365) on: KeyFrame::GetFeaturesInArea has no level arguments (src/KeyFrame.cc:906-945), the level
    window [l-1, l] is applied inside the candidate loop, any keypoint with a match is skipped, TH_LOW."""
    TH_LOW = 50
    kps = grid.kps
    sf = np.asarray(scale_factors, np.float32)
    assigned = np.array(assigned0, np.int32)
    match = np.full(len(kps), -1, np.int32)
    n = 0
    for i in range(len(mp_desc)):
        if not valid[i]:
            continue
        level = int(pred_level[i])
        radius = F(F(int(th)) * sf[level])                                                  # :365 (int th)
        idx = grid.features_in_area(u[i], v[i], radius)                                     # :367
        best, best_i = 256, -1
        for j in idx:
            if assigned[j]:                                                                 # :378-379
                continue
            lvl = int(kps["octave"][j])
            if lvl < level - 1 or lvl > level:                                              # :383-384
                continue
            d = _hamming(mp_desc[i], fdesc[j])
            if d < best:
                best, best_i = d, j
        if best <= TH_LOW:                                                                  # :396
            match[best_i] = i
            assigned[best_i] = 1
            n += 1
    return n, match, assigned


def search_window_top1(grid, kdesc, u_right, scale_factors, u, v, ur, pred_level, valid, mp_desc, th, th_dist,
                       inv_level_sigma2=None):
    """The candidate loop shared by ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th) (src/ORBmatcher.cc:883-943, with the
    chi-square gates: ur given), Fuse(KeyFrame*, Scw, ...) (:1043-1073) and SearchBySim3 (:1193-1226, :1272-1303):
    KeyFrame::GetFeaturesInArea, level window [l-1, l], best distance <= th_dist.  -> (best index or -1, distance)."""
    INT_MAX = 2147483647
    kps = grid.kps
    sf = np.asarray(scale_factors, np.float32)
    th = F(th)
    n = len(u)
    best_idx = np.full(n, -1, np.int32)
    best_dist = np.full(n, INT_MAX, np.int32)
    for i in range(n):
        if not valid[i]:
            continue
        level = int(pred_level[i])
        radius = F(th * sf[level])
        ui, vi = F(u[i]), F(v[i])
        best, bi = (256 if ur is not None else INT_MAX), -1                               # :885 vs :1045, :1198
        for j in grid.features_in_area(ui, vi, radius):
            lvl = int(kps["octave"][j])
            if lvl < level - 1 or lvl > level:
                continue
            if ur is not None:                                                            # :901-927
                ex = F(ui - F(kps["x"][j]))
                ey = F(vi - F(kps["y"][j]))
                if u_right[j] >= 0:
                    er = F(F(ur[i]) - F(u_right[j]))
                    e2 = F(F(F(ex * ex) + F(ey * ey)) + F(er * er))
                    if float(F(e2 * F(inv_level_sigma2[lvl]))) > 7.8:
                        continue
                else:
                    e2 = F(F(ex * ex) + F(ey * ey))
                    if float(F(e2 * F(inv_level_sigma2[lvl]))) > 5.99:
                        continue
            d = _hamming(mp_desc[i], kdesc[j])
            if d < best:
                best, bi = d, j
        if best <= th_dist:
            best_idx[i], best_dist[i] = bi, best
    return best_idx, best_dist


def search_by_sim3(g1, d1, sf1, g2, d2, sf2, q12, q21, th):
    """ORBmatcher::SearchBySim3, src/ORBmatcher.cc:1102-1326 from the projections on: q12 = (u, v, level, valid,
    descriptor) of key frame 1's map points in key frame 2, q21 the reverse; TH_HIGH; agreement check :1305-1320."""
    m1, _ = search_window_top1(g2, d2, None, sf2, q12[0], q12[1], None, q12[2], q12[3], q12[4], th, TH_HIGH)
    m2, _ = search_window_top1(g1, d1, None, sf1, q21[0], q21[1], None, q21[2], q21[3], q21[4], th, TH_HIGH)
    m12 = np.full(len(g1.kps), -1, np.int32)
    n = 0
    for i1 in range(len(g1.kps)):
        i2 = m1[i1]
        if i2 >= 0 and m2[i2] == i1:
            m12[i1] = i2
            n += 1
    return n, m12
