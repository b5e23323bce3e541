/*
 * match_oracle.cpp -- CPU oracle for the Hamming matchers.  TEST INFRASTRUCTURE ONLY (see orb_oracle.h).
 *
 * Restates /root/reference/src/ORBmatcher.cc (DescriptorDistance :1648-1664, SearchByProjection
 * :45-129 and :1328-1471, SearchForTriangulation :657-823, ComputeThreeMaxima :1602-1643),
 * /root/reference/src/Frame.cc (ComputeStereoMatches :646-820, grid :410-425,507-572) and
 * /root/reference/src/MapPoint.cc (ComputeDistinctiveDescriptors :249-314) on plain arrays.  Build with -ffp-contract=off.
 */
#include "orb_oracle.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <thread>
#include <utility>
#include <vector>

namespace {
const int TH_HIGH = 100;      /* ORBmatcher.cc:37 */
const int TH_LOW = 50;        /* :38 */
const int HISTO_LENGTH = 30;  /* :39 */
const int GRID_COLS = 64;     /* Frame.h:41 */
const int GRID_ROWS = 48;     /* Frame.h:42 */
}  // namespace

/* ORBmatcher::DescriptorDistance, ORBmatcher.cc:1648-1664 (Stanford bit hack) */
extern "C" int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int32_t pa[8], pb[8];
    memcpy(pa, a, 32);
    memcpy(pb, b, 32);
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        unsigned int v = pa[i] ^ pb[i];
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

extern "C" int orc_descriptor_distance_popcnt(const uint8_t* a, const uint8_t* b) {
    uint64_t pa[4], pb[4];
    memcpy(pa, a, 32);
    memcpy(pb, b, 32);
    return __builtin_popcountll(pa[0] ^ pb[0]) + __builtin_popcountll(pa[1] ^ pb[1]) +
           __builtin_popcountll(pa[2] ^ pb[2]) + __builtin_popcountll(pa[3] ^ pb[3]);
}

/* Sequential top-2 scan (strict <, first index wins): ORBmatcher.cc:201-226 / :76-115 */
static void top2_range(const uint8_t* q, int q0, int q1, const uint8_t* map, int64_t M, int64_t base,
                       orc_top2* out, int use_popcnt) {
    for (int iq = q0; iq < q1; iq++) {
        const uint8_t* dq = q + (size_t)iq * 32;
        int best1 = 256, best2 = 256, i1 = -1, i2 = -1;
        for (int64_t m = 0; m < M; m++) {
            const uint8_t* dm = map + (size_t)m * 32;
            int dist = use_popcnt ? orc_descriptor_distance_popcnt(dq, dm) : orc_descriptor_distance(dq, dm);
            if (dist < best1) {
                best2 = best1; i2 = i1;
                best1 = dist; i1 = (int)(base + m);
            } else if (dist < best2) {
                best2 = dist; i2 = (int)(base + m);
            }
        }
        out[iq].d1 = best1; out[iq].i1 = i1; out[iq].d2 = best2; out[iq].i2 = i2;
    }
}

extern "C" void orc_hamming_top2(const uint8_t* q, int Q, const uint8_t* map, int64_t M, int64_t base,
                                 orc_top2* out, int use_popcnt, int nthreads) {
    if (nthreads <= 1) { top2_range(q, 0, Q, map, M, base, out, use_popcnt); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
        int q0 = (int)((int64_t)Q * t / nthreads), q1 = (int)((int64_t)Q * (t + 1) / nthreads);
        th.emplace_back(top2_range, q, q0, q1, map, M, base, out, use_popcnt);
    }
    for (auto& t : th) t.join();
}

/* Merge of per-shard top-2 records: the two smallest (distance, global index) pairs under the total
 * order "smaller distance, then lower index" == what one sequential scan over the whole map yields. */
extern "C" void orc_top2_merge(const orc_top2* parts, int nparts, int Q, orc_top2* out) {
    for (int iq = 0; iq < Q; iq++) {
        int d1 = 256, i1 = INT_MAX, d2 = 256, i2 = INT_MAX;   /* INT_MAX index == "none" */
        auto offer = [&](int d, int i) {
            if (i < 0) return;
            if (d < d1 || (d == d1 && i < i1)) {
                d2 = d1; i2 = i1; d1 = d; i1 = i;
            } else if (d < d2 || (d == d2 && i < i2)) {
                d2 = d; i2 = i;
            }
        };
        for (int p = 0; p < nparts; p++) {
            const orc_top2& r = parts[(size_t)p * Q + iq];
            offer(r.d1, r.i1);
            offer(r.d2, r.i2);
        }
        out[iq].d1 = d1; out[iq].i1 = i1 == INT_MAX ? -1 : i1;
        out[iq].d2 = d2; out[iq].i2 = i2 == INT_MAX ? -1 : i2;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Frame::ComputeStereoMatches, Frame.cc:646-820
 * ---------------------------------------------------------------------------------------------- */
extern "C" int orc_stereo_match(const orc_keypoint* kl, const uint8_t* dl, int N,
                                const orc_keypoint* kr, const uint8_t* dr, int Nr,
                                const orc_image* pyr_l, const orc_image* pyr_r, int nlevels,
                                const float* mvScaleFactors, const float* mvInvScaleFactors,
                                float mbf, float mb, float* mvuRight, float* mvDepth,
                                int32_t* best_dist_out, int32_t* best_idx_out) {
    (void)nlevels;
    for (int i = 0; i < N; i++) { mvuRight[i] = -1.0f; mvDepth[i] = -1.0f; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const int nRows = pyr_l[0].h;
    std::vector<std::vector<size_t> > vRowIndices(nRows);
    for (int iR = 0; iR < Nr; iR++) {
        const float kpY = kr[iR].y;
        const float r = 2.0f * mvScaleFactors[kr[iR].octave];
        const int maxr = (int)ceilf(kpY + r);
        const int minr = (int)floorf(kpY - r);
        for (int yi = minr; yi <= maxr; yi++)
            if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);   /* C.6: defensive clamp */
    }
    const float minZ = mb;
    const float minD = 0;
    const float maxD = mbf / minZ;
    std::vector<std::pair<int, int> > vDistIdx;
    int nmatched = 0;
    for (int iL = 0; iL < N; iL++) {
        if (best_dist_out) best_dist_out[iL] = -1;
        if (best_idx_out) best_idx_out[iL] = -1;
        const orc_keypoint& kpL = kl[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y;
        const float uL = kpL.x;
        const std::vector<size_t>& vCandidates = vRowIndices[(size_t)vL];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD;
        const float maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        size_t bestIdxR = 0;
        const uint8_t* dL = dl + (size_t)iL * 32;
        for (size_t iC = 0; iC < vCandidates.size(); iC++) {
            const size_t iR = vCandidates[iC];
            const orc_keypoint& kpR = kr[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = orc_descriptor_distance(dL, dr + iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (best_dist_out) best_dist_out[iL] = bestDist;
        if (best_idx_out) best_idx_out[iL] = bestDist < TH_HIGH ? (int)bestIdxR : -1;
        if (bestDist < thOrbDist) {
            const float uR0 = kr[bestIdxR].x;
            const float scaleFactor = mvInvScaleFactors[kpL.octave];
            const float scaleduL = roundf(kpL.x * scaleFactor);
            const float scaledvL = roundf(kpL.y * scaleFactor);
            const float scaleduR0 = roundf(uR0 * scaleFactor);
            const int w = 5;
            const orc_image& imL = pyr_l[kpL.octave];
            const orc_image& imR = pyr_r[kpL.octave];
            const int cvL = (int)scaledvL, cuL = (int)scaleduL, cuR = (int)scaleduR0;
            /* cv::Mat::rowRange/colRange would throw outside [0,rows]x[0,cols]: treat as no match (C.5) */
            if (cvL - w < 0 || cvL + w + 1 > imL.h || cuL - w < 0 || cuL + w + 1 > imL.w) continue;
            float IL[11][11];
            {
                const float c = (float)imL.data[(size_t)cvL * imL.step + cuL];
                for (int y = 0; y < 11; y++)
                    for (int x = 0; x < 11; x++)
                        IL[y][x] = (float)imL.data[(size_t)(cvL - w + y) * imL.step + (cuL - w + x)] - c;
            }
            int bestDistS = INT_MAX;
            int bestincR = 0;
            const int L = 5;
            float vDists[2 * 5 + 1];
            const float iniu = scaleduR0 + L - w;
            const float endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= imR.w) continue;
            bool thrown = false;
            for (int incR = -L; incR <= +L; incR++) {
                const int c0 = cuR + incR - w, c1 = cuR + incR + w + 1;
                if (c0 < 0 || c1 > imR.w || cvL + w + 1 > imR.h) { thrown = true; break; }
                const float c = (float)imR.data[(size_t)cvL * imR.step + (cuR + incR)];
                double acc = 0;
                for (int y = 0; y < 11; y++)
                    for (int x = 0; x < 11; x++) {
                        float ir = (float)imR.data[(size_t)(cvL - w + y) * imR.step + (c0 + x)] - c;
                        acc += std::fabs(IL[y][x] - ir);
                    }
                float dist = (float)acc;
                if (dist < bestDistS) { bestDistS = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (thrown) continue;
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1];
            const float dist2 = vDists[L + bestincR];
            const float dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = mvScaleFactors[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) {
                    disparity = 0.01;
                    bestuR = uL - 0.01;
                }
                mvDepth[iL] = mbf / disparity;
                mvuRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(bestDistS, iL));
            }
        }
    }
    if (vDistIdx.empty()) return 0;   /* the reference reads vDistIdx[0] of an empty vector here (UB) */
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = (float)vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    nmatched = (int)vDistIdx.size();
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if (vDistIdx[i].first < thDist) break;
        mvuRight[vDistIdx[i].second] = -1;
        mvDepth[vDistIdx[i].second] = -1;
        nmatched--;
    }
    return nmatched;
}

/* ------------------------------------------------------------------------------------------------
 * Frame grid: AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea, Frame.cc:410-425,507-572
 * ---------------------------------------------------------------------------------------------- */
struct orc_grid {
    std::vector<size_t> cell[GRID_COLS][GRID_ROWS];
    const orc_keypoint* kps;
    int n;
    float mnMinX, mnMaxX, mnMinY, mnMaxY, invW, invH;
};

extern "C" orc_grid* orc_grid_create(const orc_keypoint* kps, int n, float minx, float maxx, float miny,
                                     float maxy) {
    orc_grid* g = new orc_grid();
    g->kps = kps; g->n = n;
    g->mnMinX = minx; g->mnMaxX = maxx; g->mnMinY = miny; g->mnMaxY = maxy;
    /* Frame.cc:181-182 */
    g->invW = static_cast<float>(GRID_COLS) / (maxx - minx);
    g->invH = static_cast<float>(GRID_ROWS) / (maxy - miny);
    for (int i = 0; i < n; i++) {
        int posX = (int)roundf((kps[i].x - minx) * g->invW);
        int posY = (int)roundf((kps[i].y - miny) * g->invH);
        if (posX < 0 || posX >= GRID_COLS || posY < 0 || posY >= GRID_ROWS) continue;
        g->cell[posX][posY].push_back(i);
    }
    return g;
}
extern "C" void orc_grid_destroy(orc_grid* g) { delete g; }

static std::vector<size_t> features_in_area(const orc_grid* g, float x, float y, float r, int minLevel,
                                            int maxLevel) {
    std::vector<size_t> vIndices;
    const int nMinCellX = std::max(0, (int)floorf((x - g->mnMinX - r) * g->invW));
    if (nMinCellX >= GRID_COLS) return vIndices;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)ceilf((x - g->mnMinX + r) * g->invW));
    if (nMaxCellX < 0) return vIndices;
    const int nMinCellY = std::max(0, (int)floorf((y - g->mnMinY - r) * g->invH));
    if (nMinCellY >= GRID_ROWS) return vIndices;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)ceilf((y - g->mnMinY + r) * g->invH));
    if (nMaxCellY < 0) return vIndices;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++) {
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const std::vector<size_t>& vCell = g->cell[ix][iy];
            for (size_t j = 0; j < vCell.size(); j++) {
                const orc_keypoint& kpUn = g->kps[vCell[j]];
                if (bCheckLevels) {
                    if (kpUn.octave < minLevel) continue;
                    if (maxLevel >= 0)
                        if (kpUn.octave > maxLevel) continue;
                }
                const float distx = kpUn.x - x;
                const float disty = kpUn.y - y;
                if (std::fabs(distx) < r && std::fabs(disty) < r) vIndices.push_back(vCell[j]);
            }
        }
    }
    return vIndices;
}

extern "C" int orc_grid_features_in_area(const orc_grid* g, float x, float y, float r, int minLevel,
                                         int maxLevel, int32_t* out, int cap) {
    std::vector<size_t> v = features_in_area(g, x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int32_t)v[i];
    return (int)v.size();
}

/* ORBmatcher::ComputeThreeMaxima, ORBmatcher.cc:1602-1643 */
static void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) {
            max3 = max2; max2 = max1; max1 = s;
            ind3 = ind2; ind2 = ind1; ind1 = i;
        } else if (s > max2) {
            max3 = max2; max2 = s;
            ind3 = ind2; ind2 = i;
        } else if (s > max3) {
            max3 = s; ind3 = i;
        }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), ORBmatcher.cc:45-129 */
extern "C" int orc_search_by_projection_local(const orc_grid* grid, const orc_keypoint* kps_un,
                                              const uint8_t* fdesc, const float* mvuRight,
                                              int32_t* frame_mp_obs, int nf, const float* mvScaleFactors,
                                              const float* proj_x, const float* proj_y, const float* proj_xr,
                                              const int32_t* pred_level, const float* view_cos,
                                              const uint8_t* valid, const int32_t* mp_nobs,
                                              const uint8_t* mpdesc, int nmp, float th, float mfNNratio,
                                              int32_t* match_out) {
    (void)nf;
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    for (int iMP = 0; iMP < nmp; iMP++) {
        if (!valid[iMP]) continue;   /* !mbTrackInView || isBad() */
        const int nPredictedLevel = pred_level[iMP];
        float r = view_cos[iMP] > 0.998 ? 2.5f : 4.0f;   /* RadiusByViewingCos :131-137 */
        if (bFactor) r *= th;
        const std::vector<size_t> vIndices = features_in_area(
            grid, proj_x[iMP], proj_y[iMP], r * mvScaleFactors[nPredictedLevel], nPredictedLevel - 1,
            nPredictedLevel);
        if (vIndices.empty()) continue;
        const uint8_t* MPdescriptor = mpdesc + (size_t)iMP * 32;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (size_t k = 0; k < vIndices.size(); k++) {
            const size_t idx = vIndices[k];
            if (frame_mp_obs[idx] > 0) continue;
            if (mvuRight[idx] > 0) {
                const float er = std::fabs(proj_xr[iMP] - mvuRight[idx]);
                if (er > r * mvScaleFactors[nPredictedLevel]) continue;
            }
            const int dist = orc_descriptor_distance(MPdescriptor, fdesc + idx * 32);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist;
                bestLevel2 = bestLevel; bestLevel = kps_un[idx].octave;
                bestIdx = (int)idx;
            } else if (dist < bestDist2) {
                bestLevel2 = kps_un[idx].octave;
                bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > mfNNratio * bestDist2) continue;
            match_out[bestIdx] = iMP;
            frame_mp_obs[bestIdx] = mp_nobs[iMP];
            nmatches++;
        }
    }
    return nmatches;
}

/* ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono), ORBmatcher.cc:1328-1471
 * (projection :1359-1376 is done by the caller; this is the search from :1378 on) */
extern "C" int orc_search_by_projection_frame(const orc_grid* grid, const orc_keypoint* kps_un,
                                              const uint8_t* fdesc, const float* mvuRight,
                                              int32_t* frame_mp_obs, int nf, const float* mvScaleFactors,
                                              const float* pu, const float* pv, const float* pinvz,
                                              const int32_t* last_octave, const float* last_angle,
                                              const uint8_t* valid, const int32_t* mp_nobs,
                                              const uint8_t* mpdesc, int nlast, float th, float mbf, int mode,
                                              int check_ori, int th_high, int32_t* match_out) {
    (void)nf;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    for (int i = 0; i < nlast; i++) {
        if (!valid[i]) continue;
        const float u = pu[i], v = pv[i], invzc = pinvz[i];
        if (u < grid->mnMinX || u > grid->mnMaxX) continue;
        if (v < grid->mnMinY || v > grid->mnMaxY) continue;
        int nLastOctave = last_octave[i];
        float radius = th * mvScaleFactors[nLastOctave];
        std::vector<size_t> vIndices2;
        /* mode & 7: level window of the overload (:1385-1390; 3 = the Sim3 overload's [l-1, l], :375-379, applied
         * there as a filter after KeyFrame::GetFeaturesInArea, which enumerates in the same order);
         * mode & 8: no stereo gate (the KeyFrame overloads :290-403 and :1473-1600 have none) */
        const int lm = mode & 7;
        if (lm == 1) vIndices2 = features_in_area(grid, u, v, radius, nLastOctave, -1);
        else if (lm == 2) vIndices2 = features_in_area(grid, u, v, radius, 0, nLastOctave);
        else if (lm == 3) vIndices2 = features_in_area(grid, u, v, radius, nLastOctave - 1, nLastOctave);
        else vIndices2 = features_in_area(grid, u, v, radius, nLastOctave - 1, nLastOctave + 1);
        if (vIndices2.empty()) continue;
        const uint8_t* dMP = mpdesc + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (size_t k = 0; k < vIndices2.size(); k++) {
            const size_t i2 = vIndices2[k];
            if (frame_mp_obs[i2] > 0) continue;
            if (!(mode & 8) && mvuRight[i2] > 0) {
                const float ur = u - mbf * invzc;
                const float er = std::fabs(ur - mvuRight[i2]);
                if (er > radius) continue;
            }
            const int dist = orc_descriptor_distance(dMP, fdesc + i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = (int)i2; }
        }
        if (bestDist <= th_high) {
            match_out[bestIdx2] = i;
            frame_mp_obs[bestIdx2] = mp_nobs[i];
            nmatches++;
            if (check_ori) {
                float rot = last_angle[i] - kps_un[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)roundf(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i != ind1 && i != ind2 && i != ind3) {
                for (size_t j = 0; j < rotHist[i].size(); j++) {
                    match_out[rotHist[i][j]] = -1;
                    frame_mp_obs[rotHist[i][j]] = 0;
                    nmatches--;
                }
            }
        }
    }
    return nmatches;
}

/* ORBmatcher::CheckDistEpipolarLine, ORBmatcher.cc:140-157 */
static bool check_dist_epipolar(const orc_keypoint& kp1, const orc_keypoint& kp2, const float* F12,
                                const float* sigma2_2) {
    const float a = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
    const float b = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
    const float c = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
    const float num = a * kp2.x + b * kp2.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * sigma2_2[kp2.octave];
}

/* ORBmatcher::SearchForTriangulation, ORBmatcher.cc:657-823 */
extern "C" int orc_search_for_triangulation(const orc_keypoint* k1, const uint8_t* d1, const float* ur1,
                                            const uint8_t* has_mp1, int n1, const orc_keypoint* k2,
                                            const uint8_t* d2, const float* ur2, const uint8_t* has_mp2,
                                            int n2, const int32_t* node_id1, const int32_t* node_ptr1,
                                            const int32_t* idx1v, int nn1, const int32_t* node_id2,
                                            const int32_t* node_ptr2, const int32_t* idx2v, int nn2,
                                            const float* F12, float ex, float ey, const float* scale2,
                                            const float* sigma2_2, int bOnlyStereo, int check_ori,
                                            int32_t* vMatches12) {
    (void)n2;
    int nmatches = 0;
    for (int i = 0; i < n1; i++) vMatches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int f1 = 0, f2 = 0;
    while (f1 < nn1 && f2 < nn2) {
        if (node_id1[f1] == node_id2[f2]) {
            for (int i1 = node_ptr1[f1]; i1 < node_ptr1[f1 + 1]; i1++) {
                const int idx1 = idx1v[i1];
                if (has_mp1[idx1]) continue;
                const bool bStereo1 = ur1[idx1] >= 0;
                if (bOnlyStereo && !bStereo1) continue;
                const orc_keypoint& kp1 = k1[idx1];
                int bestDist = TH_LOW;
                int bestIdx2 = -1;
                for (int i2 = node_ptr2[f2]; i2 < node_ptr2[f2 + 1]; i2++) {
                    const int idx2 = idx2v[i2];
                    if (has_mp2[idx2]) continue;   /* vbMatched2 is never set in the reference */
                    const bool bStereo2 = ur2[idx2] >= 0;
                    if (bOnlyStereo && !bStereo2) continue;
                    const int dist = orc_descriptor_distance(d1 + (size_t)idx1 * 32, d2 + (size_t)idx2 * 32);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const orc_keypoint& kp2 = k2[idx2];
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ex - kp2.x;
                        const float distey = ey - kp2.y;
                        if (distex * distex + distey * distey < 100 * scale2[kp2.octave]) continue;
                    }
                    if (check_dist_epipolar(kp1, kp2, F12, sigma2_2)) {
                        bestIdx2 = idx2;
                        bestDist = dist;
                    }
                }
                if (bestIdx2 >= 0) {
                    const orc_keypoint& kp2 = k2[bestIdx2];
                    vMatches12[idx1] = bestIdx2;
                    nmatches++;
                    if (check_ori) {
                        float rot = kp1.angle - kp2.angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)roundf(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(idx1);
                    }
                }
            }
            f1++; f2++;
        } else if (node_id1[f1] < node_id2[f2]) {
            f1 = (int)(std::lower_bound(node_id1, node_id1 + nn1, node_id2[f2]) - node_id1);
        } else {
            f2 = (int)(std::lower_bound(node_id2, node_id2 + nn2, node_id1[f1]) - node_id2);
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) {
                vMatches12[rotHist[i][j]] = -1;
                nmatches--;
            }
        }
    }
    return nmatches;
}

/* MapPoint::ComputeDistinctiveDescriptors, MapPoint.cc:249-314, for one point: desc = the N descriptors of its
 * non-bad observations in map iteration order (:269-275).  Returns BestIdx (-1 when N == 0, the early return of
 * :277-278) and BestMedian through *median.  The float Distances[N][N] of :284 holds integers <= 256 exactly. */
extern "C" int orc_distinctive_descriptor(const uint8_t* desc, int N, int* median_out) {
    if (median_out) *median_out = INT_MAX;
    if (N <= 0) return -1;
    std::vector<float> D((size_t)N * N);
    for (int i = 0; i < N; i++) {
        D[(size_t)i * N + i] = 0;
        for (int j = i + 1; j < N; j++) {
            const int distij = orc_descriptor_distance(desc + (size_t)i * 32, desc + (size_t)j * 32);
            D[(size_t)i * N + j] = distij;
            D[(size_t)j * N + i] = distij;
        }
    }
    int BestMedian = INT_MAX, BestIdx = 0;
    for (int i = 0; i < N; i++) {
        std::vector<int> vDists(D.begin() + (size_t)i * N, D.begin() + (size_t)(i + 1) * N);
        std::sort(vDists.begin(), vDists.end());
        const int median = vDists[(size_t)(0.5 * (N - 1))];
        if (median < BestMedian) {
            BestMedian = median;
            BestIdx = i;
        }
    }
    if (median_out) *median_out = BestMedian;
    return BestIdx;
}

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), ORBmatcher.cc:159-288 (mode 0) and
 * ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&), ORBmatcher.cc:522-655 (mode 1), on the flattened
 * inputs of the C ABI.  mode 0: match has n2 entries (vpMapPointMatches, as the index of the map point in pKF);
 * mode 1: n1 entries (vpMatches12, as the keypoint index in pKF2). */
extern "C" int orc_search_by_bow(int mode, const orc_keypoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                                 const orc_keypoint* k2, const uint8_t* d2, const uint8_t* valid2, int n2,
                                 const int32_t* node_id1, const int32_t* node_ptr1, const int32_t* idx1v, int nn1,
                                 const int32_t* node_id2, const int32_t* node_ptr2, const int32_t* idx2v, int nn2,
                                 float mfNNratio, int check_ori, int32_t* match) {
    const int nOut = mode == 0 ? n2 : n1;
    for (int i = 0; i < nOut; i++) match[i] = -1;
    std::vector<bool> vbMatched2(n2, false);
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int f1 = 0, f2 = 0;
    while (f1 < nn1 && f2 < nn2) {
        if (node_id1[f1] == node_id2[f2]) {
            for (int e1 = node_ptr1[f1]; e1 < node_ptr1[f1 + 1]; e1++) {
                const int idx1 = idx1v[e1];
                if (!valid1[idx1]) continue;                              /* !pMP || pMP->isBad() */
                const uint8_t* dA = d1 + (size_t)idx1 * 32;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int e2 = node_ptr2[f2]; e2 < node_ptr2[f2 + 1]; e2++) {
                    const int idx2 = idx2v[e2];
                    if (mode == 0) {
                        if (match[idx2] >= 0) continue;                   /* vpMapPointMatches[realIdxF] (:196) */
                    } else {
                        if (vbMatched2[idx2] || !valid2[idx2]) continue;  /* :571-575 */
                    }
                    const int dist = orc_descriptor_distance(dA, d2 + (size_t)idx2 * 32);
                    if (dist < bestDist1) {
                        bestDist2 = bestDist1;
                        bestDist1 = dist;
                        bestIdx2 = idx2;
                    } else if (dist < bestDist2) {
                        bestDist2 = dist;
                    }
                }
                const bool pass = mode == 0 ? bestDist1 <= TH_LOW : bestDist1 < TH_LOW;     /* :216 vs :597 */
                if (pass) {
                    if (static_cast<float>(bestDist1) < mfNNratio * static_cast<float>(bestDist2)) {
                        if (mode == 0) match[bestIdx2] = idx1;
                        else { match[idx1] = bestIdx2; vbMatched2[bestIdx2] = true; }
                        if (check_ori) {
                            float rot = k1[idx1].angle - k2[bestIdx2].angle;
                            if (rot < 0.0) rot += 360.0f;
                            int bin = (int)roundf(rot * factor);
                            if (bin == HISTO_LENGTH) bin = 0;
                            rotHist[bin].push_back(mode == 0 ? bestIdx2 : idx1);
                        }
                        nmatches++;
                    }
                }
            }
            f1++; f2++;
        } else if (node_id1[f1] < node_id2[f2]) {
            f1 = (int)(std::lower_bound(node_id1, node_id1 + nn1, node_id2[f2]) - node_id1);
        } else {
            f2 = (int)(std::lower_bound(node_id2, node_id2 + nn2, node_id1[f1]) - node_id2);
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) {
                match[rotHist[i][j]] = -1;
                nmatches--;
            }
        }
    }
    return nmatches;
}

/* ORBmatcher::SearchForInitialization, ORBmatcher.cc:405-520.  grid2 / k2 / d2 = F2 (mvKeysUn); prev = vbPrevMatched
 * (in/out, n1 x 2). */
extern "C" int orc_search_for_initialization(const orc_grid* grid2, const orc_keypoint* k2, const uint8_t* d2, int n2,
                                             const orc_keypoint* k1, const uint8_t* d1, int n1, float* prev,
                                             int windowSize, float mfNNratio, int check_ori, int32_t* vnMatches12) {
    int nmatches = 0;
    for (int i = 0; i < n1; i++) vnMatches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> vMatchedDistance(n2, INT_MAX);
    std::vector<int> vnMatches21(n2, -1);
    for (int i1 = 0; i1 < n1; i1++) {
        const orc_keypoint kp1 = k1[i1];
        const int level1 = kp1.octave;
        if (level1 > 0) continue;
        std::vector<size_t> vIndices2 = features_in_area(grid2, prev[2 * i1], prev[2 * i1 + 1], (float)windowSize, level1, level1);
        if (vIndices2.empty()) continue;
        const uint8_t* dA = d1 + (size_t)i1 * 32;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (size_t k = 0; k < vIndices2.size(); k++) {
            const size_t i2 = vIndices2[k];
            const int dist = orc_descriptor_distance(dA, d2 + i2 * 32);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) {
                bestDist2 = bestDist;
                bestDist = dist;
                bestIdx2 = (int)i2;
            } else if (dist < bestDist2) {
                bestDist2 = dist;
            }
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * mfNNratio) {
                if (vnMatches21[bestIdx2] >= 0) {
                    vnMatches12[vnMatches21[bestIdx2]] = -1;
                    nmatches--;
                }
                vnMatches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (check_ori) {
                    float rot = k1[i1].angle - k2[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)roundf(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rotHist[bin].push_back(i1);
                }
            }
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) {
                const int idx1 = rotHist[i][j];
                if (vnMatches12[idx1] >= 0) {
                    vnMatches12[idx1] = -1;
                    nmatches--;
                }
            }
        }
    }
    for (int i1 = 0; i1 < n1; i1++)
        if (vnMatches12[i1] >= 0) {
            prev[2 * i1] = k2[vnMatches12[i1]].x;
            prev[2 * i1 + 1] = k2[vnMatches12[i1]].y;
        }
    return nmatches;
}

/* Frame::UndistortKeyPoints (Frame.cc:584-614): cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK).  OpenCV is a
 * third-party dependency (pinned 2.4.x by the reference; 4.13 in this image): cvUndistortPoints normalises with the
 * inverse intrinsics, runs five iterations of the distortion compensation in double precision and re-projects with
 * P*R = K.  Pinned bit for bit against cv2.undistortPoints (tests/test_oracle.py::test_undistort_vs_cv2).
 * dist = k1 k2 p1 p2 [k3 [k4 k5 k6 [s1 s2 s3 s4]]]. */
static void undistort_point(double fx, double fy, double cx, double cy, const double* k, float uf, float vf, float* xo, float* yo) {
    const double ifx = 1. / fx, ify = 1. / fy;
    const double u = uf, v = vf;
    double x = (u - cx) * ifx, y = (v - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0) {       /* OpenCV >= 3.4 only; 2.4 has no such exit (never taken for physical lenses) */
            x = (u - cx) * ifx;
            y = (v - cy) * ify;
            break;
        }
        const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
        const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    const double xx = fx * x + 0. * y + cx;
    const double yy = 0. * x + fy * y + cy;
    const double ww = 1. / (0. * x + 0. * y + 1.);
    *xo = (float)(xx * ww);
    *yo = (float)(yy * ww);
}

extern "C" void orc_undistort_keypoints(const orc_keypoint* kps, int n, float fx, float fy, float cx, float cy,
                                        const float* dist, int ndist, orc_keypoint* out) {
    double k[12] = {0};
    for (int i = 0; i < ndist && i < 12; i++) k[i] = dist[i];
    for (int i = 0; i < n; i++) {
        out[i] = kps[i];
        if (ndist > 0 && dist[0] != 0.0)          /* :586-590 */
            undistort_point(fx, fy, cx, cy, k, kps[i].x, kps[i].y, &out[i].x, &out[i].y);
    }
}

/* Frame::ComputeImageBounds, Frame.cc:616-645 -> {mnMinX, mnMaxX, mnMinY, mnMaxY} */
extern "C" void orc_compute_image_bounds(int cols, int rows, float fx, float fy, float cx, float cy, const float* dist,
                                         int ndist, float* bounds) {
    if (ndist > 0 && dist[0] != 0.0) {
        double k[12] = {0};
        for (int i = 0; i < ndist && i < 12; i++) k[i] = dist[i];
        float m[4][2];
        const float in[4][2] = {{0.0f, 0.0f}, {(float)cols, 0.0f}, {0.0f, (float)rows}, {(float)cols, (float)rows}};
        for (int i = 0; i < 4; i++) undistort_point(fx, fy, cx, cy, k, in[i][0], in[i][1], &m[i][0], &m[i][1]);
        bounds[0] = std::min(m[0][0], m[2][0]);
        bounds[1] = std::max(m[1][0], m[3][0]);
        bounds[2] = std::min(m[0][1], m[1][1]);
        bounds[3] = std::max(m[2][1], m[3][1]);
    } else {
        bounds[0] = 0.0f; bounds[1] = (float)cols; bounds[2] = 0.0f; bounds[3] = (float)rows;
    }
}

/* The per-map-point search loop of ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th) (ORBmatcher.cc:883-943, with ur != NULL),
 * Fuse(KeyFrame*, Scw, ...) (:1043-1073) and SearchBySim3 (:1193-1226, :1272-1303) on a KeyFrame grid
 * (KeyFrame::GetFeaturesInArea, KeyFrame.cc:906-945).  valid[i] = the caller's projection gates passed. */
extern "C" void orc_search_window_top1(const orc_grid* grid, const orc_keypoint* kps_un, const uint8_t* kdesc,
                                       const float* mvuRight, const float* scale_factors, const float* u, const float* v,
                                       const float* ur, const int32_t* pred_level, const uint8_t* valid, const uint8_t* mp_desc,
                                       int n, float th, int th_dist, const float* inv_level_sigma2, int32_t* best_idx,
                                       int32_t* best_dist) {
    for (int i = 0; i < n; i++) {
        best_idx[i] = -1;
        if (best_dist) best_dist[i] = INT_MAX;
        if (!valid[i]) continue;
        const int nPredictedLevel = pred_level[i];
        const float radius = th * scale_factors[nPredictedLevel];
        const std::vector<size_t> vIndices = features_in_area(grid, u[i], v[i], radius, -1, -1);
        if (vIndices.empty()) continue;
        const uint8_t* dMP = mp_desc + (size_t)i * 32;
        int bestDist = ur ? 256 : INT_MAX;       /* :885 vs :1045,:1198 */
        int bestIdx = -1;
        for (size_t k = 0; k < vIndices.size(); k++) {
            const size_t idx = vIndices[k];
            const orc_keypoint& kp = kps_un[idx];
            const int kpLevel = kp.octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            if (ur) {
                if (mvuRight[idx] >= 0) {
                    const float ex = u[i] - kp.x;
                    const float ey = v[i] - kp.y;
                    const float er = ur[i] - mvuRight[idx];
                    const float e2 = ex * ex + ey * ey + er * er;
                    if (e2 * inv_level_sigma2[kpLevel] > 7.8) continue;
                } else {
                    const float ex = u[i] - kp.x;
                    const float ey = v[i] - kp.y;
                    const float e2 = ex * ex + ey * ey;
                    if (e2 * inv_level_sigma2[kpLevel] > 5.99) continue;
                }
            }
            const int dist = orc_descriptor_distance(dMP, kdesc + idx * 32);
            if (dist < bestDist) {
                bestDist = dist;
                bestIdx = (int)idx;
            }
        }
        if (bestDist <= th_dist) {
            best_idx[i] = bestIdx;
            if (best_dist) best_dist[i] = bestDist;
        }
    }
}

/* ORBmatcher::SearchBySim3, ORBmatcher.cc:1102-1326, from the projections on (the matrix algebra of :1105-1191 is the
 * caller's): both searches with TH_HIGH, then the agreement check (:1305-1320). */
extern "C" int orc_search_by_sim3(const orc_grid* g1, const orc_keypoint* k1, const uint8_t* d1, const float* sf1, int n1,
                                  const orc_grid* g2, const orc_keypoint* k2, const uint8_t* d2, const float* sf2, int n2,
                                  const float* u12, const float* v12, const int32_t* level12, const uint8_t* valid12,
                                  const uint8_t* mp_desc1, const float* u21, const float* v21, const int32_t* level21,
                                  const uint8_t* valid21, const uint8_t* mp_desc2, float th, int32_t* match12) {
    std::vector<int32_t> vnMatch1(std::max(n1, 1), -1), vnMatch2(std::max(n2, 1), -1);
    orc_search_window_top1(g2, k2, d2, NULL, sf2, u12, v12, NULL, level12, valid12, mp_desc1, n1, th, TH_HIGH, NULL, vnMatch1.data(), NULL);
    orc_search_window_top1(g1, k1, d1, NULL, sf1, u21, v21, NULL, level21, valid21, mp_desc2, n2, th, TH_HIGH, NULL, vnMatch2.data(), NULL);
    int nFound = 0;
    for (int i1 = 0; i1 < n1; i1++) {
        match12[i1] = -1;
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0) {
            const int idx1 = vnMatch2[idx2];
            if (idx1 == i1) {
                match12[i1] = idx2;
                nFound++;
            }
        }
    }
    return nFound;
}
