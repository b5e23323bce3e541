"""Second, independent restatement of Frame::ComputeStereoMatches (/root/reference/src/Frame.cc:646-820) in numpy
float32 arithmetic, statement by statement (line numbers inline).  Test infrastructure: it pins oracle/match_oracle.cpp
(orc_stereo_match) the same way tests/cv2_restatement.py pins the extractor oracle -- two readings of the reference
that must agree bit for bit.  Conventions where the reference would throw or read out of range:
a window that leaves the level image counts as "no match"; an empty match list skips the median cut.
"""
import math

import numpy as np

F = np.float32
TH_HIGH, TH_LOW = 100, 50                      # src/ORBmatcher.cc:37-38


def _roundf(x):
    """C round() on a float: half away from zero."""
    x = float(x)
    return F(math.floor(x + 0.5) if x >= 0 else math.ceil(x - 0.5))


def _hamming(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def compute_stereo_matches(kl, dl, kr, dr, pyr_l, pyr_r, scale_factors, mbf, mb):
    n = len(kl)
    sf = np.asarray(scale_factors, np.float32)
    inv_sf = (F(1.0) / sf).astype(np.float32)                       # mvInvScaleFactors (ORBextractor.cc:425-431)
    mbf, mb = F(mbf), F(mb)
    u_right = np.full(n, -1.0, np.float32)                          # :648-649
    depth = np.full(n, -1.0, np.float32)
    th_orb = (TH_HIGH + TH_LOW) // 2                                # :651
    n_rows = pyr_l[0].shape[0]                                      # :653
    rows = [[] for _ in range(n_rows)]                              # :656
    for ir in range(len(kr)):                                       # :663-673
        y = F(kr["y"][ir])
        r = F(2.0) * sf[kr["octave"][ir]]
        maxr = int(math.ceil(F(y + r)))
        minr = int(math.floor(F(y - r)))
        for yi in range(minr, maxr + 1):
            if 0 <= yi < n_rows:                                    # C.6: cannot leave the table for real keypoints
                rows[yi].append(ir)
    min_d = F(0)                                                    # :676-678
    max_d = F(mbf / mb)
    dist_idx = []
    for il in range(n):                                             # :684
        level = int(kl["octave"][il])
        v_l, u_l = F(kl["y"][il]), F(kl["x"][il])
        cand = rows[int(v_l)]                                       # :691 (truncation)
        if not cand:
            continue
        min_u = F(u_l - max_d)
        max_u = F(u_l - min_d)
        if max_u < 0:
            continue
        best, best_r = TH_HIGH, 0                                   # :702-703
        for ir in cand:                                             # :708-728
            o = int(kr["octave"][ir])
            if o < level - 1 or o > level + 1:
                continue
            u_r = F(kr["x"][ir])
            if min_u <= u_r <= max_u:
                d = _hamming(dl[il], dr[ir])
                if d < best:
                    best, best_r = d, ir
        if best >= th_orb:                                          # :732
            continue
        u_r0 = F(kr["x"][best_r])                                   # :735-739
        s = inv_sf[level]
        su_l = _roundf(F(u_l * s))
        sv_l = _roundf(F(v_l * s))
        su_r0 = _roundf(F(u_r0 * s))
        w, L = 5, 5
        img_l, img_r = pyr_l[level], pyr_r[level]

        def window(img, cu, cv):
            x0, y0 = int(cu) - w, int(cv) - w
            if x0 < 0 or y0 < 0 or x0 + 2 * w + 1 > img.shape[1] or y0 + 2 * w + 1 > img.shape[0]:
                return None
            p = img[y0:y0 + 2 * w + 1, x0:x0 + 2 * w + 1].astype(np.float32)
            return p - p[w, w]                                      # :744-745, :761-762

        il_patch = window(img_l, su_l, sv_l)                        # :743
        if il_patch is None:
            continue
        iniu = F(su_r0 + F(L) - F(w))                               # :753-756
        endu = F(su_r0 + F(L) + F(w) + F(1))
        if iniu < 0 or endu >= img_r.shape[1]:
            continue
        best_sad, best_inc = 2147483647, 0                          # :747-748
        dists = [F(0)] * (2 * L + 1)
        thrown = False
        for inc in range(-L, L + 1):                                # :758-772
            ir_patch = window(img_r, su_r0 + F(inc), sv_l)
            if ir_patch is None:
                thrown = True
                break
            d = F(np.abs(il_patch.astype(np.float64) - ir_patch.astype(np.float64)).sum())   # cv::norm L1 -> float
            if d < F(best_sad):                                     # float vs int compare
                best_sad, best_inc = int(d), inc
            dists[L + inc] = d
        if thrown or best_inc == -L or best_inc == L:               # :774
            continue
        d1, d2, d3 = dists[L + best_inc - 1], dists[L + best_inc], dists[L + best_inc + 1]   # :778-782
        with np.errstate(divide="ignore", invalid="ignore"):
            delta = F(F(d1 - d3) / F(F(2.0) * F(F(d1 + d3) - F(F(2.0) * d2))))
        if delta < -1 or delta > 1:                                 # :784 (a NaN passes here and fails :792, as in C)
            continue
        best_ur = F(sf[level] * F(F(su_r0 + F(best_inc)) + delta))  # :788
        disparity = F(u_l - best_ur)                                # :790
        if disparity >= min_d and disparity < max_d:                # :792 (false for NaN)
            if disparity <= 0:                                      # :794-798 (double constants)
                disparity = F(0.01)
                best_ur = F(float(u_l) - 0.01)
            depth[il] = F(mbf / disparity)
            u_right[il] = best_ur
            dist_idx.append((best_sad, il))
    if dist_idx:                                                    # :806-819
        dist_idx.sort()
        median = F(dist_idx[len(dist_idx) // 2][0])
        th = F(F(F(1.5) * F(1.4)) * median)
        for d, i in reversed(dist_idx):
            if F(d) < th:
                break
            u_right[i] = -1
            depth[i] = -1
    return u_right, depth
