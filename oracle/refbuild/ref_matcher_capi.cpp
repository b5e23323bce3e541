/*
 * ref_matcher_capi.cpp -- flat C entry points around the REFERENCE's own matcher code: src/ORBmatcher.cc compiled whole and
 * unmodified, the ORB members of src/Frame.cc / src/KeyFrame.cc / src/MapPoint.cc taken as verbatim line ranges
 * (ranges.sh), DBoW2 from Thirdparty/DBoW2.  TEST INFRASTRUCTURE ONLY.
 *
 * Signatures mirror the orc_* functions of oracle/orb_oracle.h: each adapter builds the Frame / KeyFrame / MapPoint
 * objects the reference function reads from the flat arrays, calls the reference, and flattens what it wrote.
 *
 * The oracle's API is cut AFTER the projection of map points (it takes u, v, 1/z); the reference projects inside the
 * search.  The adapters therefore give the reference a trivial camera -- identity pose, fx = fy = 1, cx = cy = 0 -- and
 * world points (u/invz, v/invz, 1/invz) chosen (by a search over neighbouring floats) so that the reference's own
 * float arithmetic lands exactly on the requested (u, v, 1/z); inputs for which no such point exists are refused with
 * -2 (tests snap their scenario with ref_snap_projection first).  The projection arithmetic itself (general pose) is
 * pinned separately: the cv::gemm forms in tests/test_ref_minicv.py, the shim-level comparison in tests/cpp.
 */
#include <opencv2/core/core.hpp>

#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "ORBextractor.h"
#include "ORBmatcher.h"
#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"
#include "../orb_oracle.h"

using namespace ORB_SLAM2;

namespace {

cv::Mat desc_row(const uint8_t* d) {
    cv::Mat m(1, 32, CV_8UC1);
    memcpy(m.data, d, 32);
    return m;
}

cv::Mat desc_matrix(const uint8_t* d, int n) {
    cv::Mat m(std::max(n, 1), 32, CV_8UC1);
    if (n > 0) memcpy(m.data, d, (size_t)n * 32);
    return m;
}

std::vector<cv::KeyPoint> keys(const orc_keypoint* k, int n) {
    std::vector<cv::KeyPoint> v(n);
    if (n > 0) memcpy((void*)v.data(), k, (size_t)n * sizeof(orc_keypoint));
    return v;
}

cv::Mat vec3(float x, float y, float z) {
    cv::Mat m(3, 1, CV_32F);
    m.at<float>(0) = x; m.at<float>(1) = y; m.at<float>(2) = z;
    return m;
}

struct RefGrid {
    Frame F;                 /* keypoints, bounds and the 64x48 grid filled by the reference's AssignFeaturesToGrid */
};

/* a frame as the Frame constructors leave it (src/Frame.cc:143-207): bounds, grid element sizes, grid */
void init_frame(Frame& F, const orc_keypoint* kps_un, int n, float minx, float maxx, float miny, float maxy) {
    F.N = n;
    F.mvKeysUn = keys(kps_un, n);
    F.mvKeys = F.mvKeysUn;
    F.mnMinX = minx; F.mnMaxX = maxx; F.mnMinY = miny; F.mnMaxY = maxy;
    F.mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(F.mnMaxX - F.mnMinX);      /* :181-182 */
    F.mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(F.mnMaxY - F.mnMinY);
    F.mvpMapPoints.assign(n, static_cast<MapPoint*>(NULL));
    F.mvbOutlier.assign(n, false);
    F.mvuRight.assign(n, -1.0f);
    F.AssignFeaturesToGrid();
}

void fill_frame(Frame& F, const uint8_t* fdesc, const float* fu_right, const float* scale_factors, int nlevels = 8) {
    F.mDescriptors = desc_matrix(fdesc, F.N);
    if (fu_right) F.mvuRight.assign(fu_right, fu_right + F.N);
    F.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
    F.mvInvScaleFactors.resize(nlevels);
    for (int l = 0; l < nlevels; l++) F.mvInvScaleFactors[l] = 1.0f / F.mvScaleFactors[l];
    F.mnScaleLevels = nlevels;
    F.mfLogScaleFactor = log(nlevels > 1 ? scale_factors[1] : 1.2f);
    F.fx = F.fy = 1.f; F.cx = F.cy = 0.f;
    F.mTcw = cv::Mat::eye(4, 4, CV_32F);
}

/* key frame made of a frame: src/KeyFrame.cc:270-306 copies keypoints, bounds and the grid */
void keyframe_from_frame(KeyFrame& K, const Frame& F) {
    K.N = F.N;
    K.mvKeysUn = F.mvKeysUn;
    K.mvuRight = F.mvuRight;
    K.mDescriptors = F.mDescriptors;
    K.mnMinX = F.mnMinX; K.mnMaxX = F.mnMaxX; K.mnMinY = F.mnMinY; K.mnMaxY = F.mnMaxY;
    K.mfGridElementWidthInv = F.mfGridElementWidthInv; K.mfGridElementHeightInv = F.mfGridElementHeightInv;
    K.mnGridCols = FRAME_GRID_COLS; K.mnGridRows = FRAME_GRID_ROWS;
    K.mGrid.resize(K.mnGridCols);
    for (int i = 0; i < K.mnGridCols; i++) {
        K.mGrid[i].resize(K.mnGridRows);
        for (int j = 0; j < K.mnGridRows; j++) K.mGrid[i][j] = F.mGrid[i][j];
    }
    K.mvScaleFactors = F.mvScaleFactors;
    K.mnScaleLevels = F.mnScaleLevels;
    K.mfLogScaleFactor = F.mfLogScaleFactor;
    K.fx = K.fy = 1.f; K.cx = K.cy = 0.f;
    K.mapPoints.assign(K.N, static_cast<MapPoint*>(NULL));
}

/* world point that the reference's projection (identity pose, unit intrinsics) maps exactly onto (u, v, invz):
 *   invzc = 1.0/Z (double division, rounded to float), u = fx*X*invzc + cx = X*invzc, v likewise */
bool unproject_exact(float u, float v, float invz, float out[3]) {
    float Z = (float)(1.0 / (double)invz);
    bool okz = false;
    for (int s = 0; s < 9 && !okz; s++) {
        float z = Z;
        for (int k = 0; k < (s + 1) / 2; k++) z = nextafterf(z, (s & 1) ? INFINITY : -INFINITY);
        const float back = 1.0 / z;
        if (back == invz) { Z = z; okz = true; }
    }
    if (!okz) return false;
    float c[2] = {u, v};
    for (int a = 0; a < 2; a++) {
        const float X0 = c[a] / invz;
        bool ok = false;
        for (int s = 0; s < 17 && !ok; s++) {
            float x = X0;
            for (int k = 0; k < (s + 1) / 2; k++) x = nextafterf(x, (s & 1) ? INFINITY : -INFINITY);
            const float p = 1.f * x * invz + 0.f;
            if (p == c[a]) { out[a] = x; ok = true; }
        }
        if (!ok) return false;
    }
    out[2] = Z;
    return true;
}

struct MapPointPool {
    std::vector<MapPoint*> all;
    ~MapPointPool() { for (size_t i = 0; i < all.size(); i++) delete all[i]; }
    MapPoint* make() { all.push_back(new MapPoint()); return all.back(); }
};

/* frame_mp_obs[i] > 0: keypoint i already holds a map point with that many observations */
void seed_frame_mappoints(std::vector<MapPoint*>& slots, const int32_t* obs, int n, MapPointPool& pool) {
    slots.assign(n, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < n; i++)
        if (obs[i] > 0) { MapPoint* p = pool.make(); p->nObs = obs[i]; slots[i] = p; }
}

void build_feature_vector(DBoW2::FeatureVector& fv, const int32_t* node_id, const int32_t* node_ptr, const int32_t* idx, int nn) {
    for (int f = 0; f < nn; f++)
        for (int e = node_ptr[f]; e < node_ptr[f + 1]; e++) fv.addFeature((DBoW2::NodeId)node_id[f], (unsigned)idx[e]);
}

}  // namespace

extern "C" {

/* ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:1648-1664 */
int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) { return ORBmatcher::DescriptorDistance(desc_row(a), desc_row(b)); }
int ref_descriptor_distance_popcnt(const uint8_t* a, const uint8_t* b) { return ref_descriptor_distance(a, b); }

/* snaps (u, v, invz) onto values the trivial camera of these adapters reproduces exactly (in place); see the header */
void ref_snap_projection(float* u, float* v, float* invz, int n) {
    for (int i = 0; i < n; i++) {
        const float Z = (float)(1.0 / (double)invz[i]);
        const float iz = 1.0 / Z;
        const float X = u[i] / iz, Y = v[i] / iz;
        invz[i] = iz;
        u[i] = 1.f * X * iz + 0.f;
        v[i] = 1.f * Y * iz + 0.f;
    }
}

/* Frame::ComputeStereoMatches, src/Frame.cc:646-820 */
int ref_stereo_match(const orc_keypoint* kl, const uint8_t* dl, int nl, const orc_keypoint* kr, const uint8_t* dr, int nr,
                     const orc_image* pyr_l, const orc_image* pyr_r, int nlevels, const float* scale_factors,
                     const float* inv_scale_factors, float mbf, float mb, float* u_right, float* depth, int32_t* best_dist_out,
                     int32_t* best_idx_out) {
    (void)best_dist_out; (void)best_idx_out;                    /* internals of the restatement; the reference keeps none */
    ORBextractor exL(1000, 1.2f, nlevels, 20, 7), exR(1000, 1.2f, nlevels, 20, 7);
    for (int l = 0; l < nlevels; l++) {
        exL.mvImagePyramid[l] = cv::Mat(pyr_l[l].h, pyr_l[l].w, CV_8UC1, (void*)pyr_l[l].data, pyr_l[l].step);
        exR.mvImagePyramid[l] = cv::Mat(pyr_r[l].h, pyr_r[l].w, CV_8UC1, (void*)pyr_r[l].data, pyr_r[l].step);
    }
    Frame F;
    F.N = nl;
    F.mvKeys = keys(kl, nl);
    F.mvKeysUn = F.mvKeys;
    F.mvKeysRight = keys(kr, nr);
    F.mDescriptors = desc_matrix(dl, nl);
    F.mDescriptorsRight = desc_matrix(dr, nr);
    F.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
    F.mvInvScaleFactors.assign(inv_scale_factors, inv_scale_factors + nlevels);
    F.mbf = mbf; F.mb = mb;
    F.mpORBextractorLeft = &exL; F.mpORBextractorRight = &exR;
    F.ComputeStereoMatches();
    int n = 0;
    for (int i = 0; i < nl; i++) {
        u_right[i] = F.mvuRight[i];
        depth[i] = F.mvDepth[i];
        n += F.mvuRight[i] >= 0;
    }
    return n;
}

/* Frame grid: AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea, src/Frame.cc:410-425,507-572 */
RefGrid* ref_grid_create(const orc_keypoint* kps_un, int n, float minx, float maxx, float miny, float maxy) {
    RefGrid* g = new RefGrid();
    init_frame(g->F, kps_un, n, minx, maxx, miny, maxy);
    return g;
}
void ref_grid_destroy(RefGrid* g) { delete g; }
int ref_grid_features_in_area(const RefGrid* g, float x, float y, float r, int min_level, int max_level, int32_t* out, int cap) {
    const std::vector<size_t> v = g->F.GetFeaturesInArea(x, y, r, min_level, max_level);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int32_t)v[i];
    return (int)v.size();
}

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), src/ORBmatcher.cc:45-129 */
int ref_search_by_projection_local(const RefGrid* grid, const orc_keypoint* kps_un, const uint8_t* fdesc, const float* fu_right,
                                   int32_t* frame_mp_obs, int nf, const float* scale_factors, const float* proj_x,
                                   const float* proj_y, const float* proj_xr, const int32_t* pred_level, const float* view_cos,
                                   const uint8_t* valid, const int32_t* mp_nobs, const uint8_t* mpdesc, int nmp, float th,
                                   float nnratio, int32_t* match_out) {
    (void)kps_un;
    Frame F = grid->F;
    fill_frame(F, fdesc, fu_right, scale_factors);
    MapPointPool pool;
    seed_frame_mappoints(F.mvpMapPoints, frame_mp_obs, nf, pool);
    std::vector<MapPoint*> vp(nmp);
    for (int i = 0; i < nmp; i++) {
        MapPoint* p = pool.make();
        p->mbTrackInView = valid[i] != 0;
        p->mTrackProjX = proj_x[i]; p->mTrackProjY = proj_y[i]; p->mTrackProjXR = proj_xr[i];
        p->mnTrackScaleLevel = pred_level[i]; p->mTrackViewCos = view_cos[i];
        p->nObs = mp_nobs[i];
        p->descriptor = desc_row(mpdesc + (size_t)i * 32);
        vp[i] = p;
    }
    ORBmatcher matcher(nnratio, true);
    const int n = matcher.SearchByProjection(F, vp, th);
    for (int k = 0; k < nf; k++) {
        MapPoint* p = F.mvpMapPoints[k];
        frame_mp_obs[k] = p ? p->Observations() : 0;
        for (int i = 0; i < nmp && p; i++)
            if (vp[i] == p) { match_out[k] = i; break; }
    }
    return n;
}

/* the three overloads that project map points themselves:
 *   mode & 7 = 0/1/2, no bit 8:  SearchByProjection(Frame& Cur, const Frame& Last, th, bMono)      src/ORBmatcher.cc:1328-1471
 *   mode = 0|8:                   SearchByProjection(Frame& Cur, KeyFrame*, sAlreadyFound, th, dist) :1473-1600
 *   mode = 3|8:                   SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)        :290-403 */
int ref_search_by_projection_frame(const RefGrid* grid, const orc_keypoint* kps_un, const uint8_t* fdesc, const float* fu_right,
                                   int32_t* frame_mp_obs, int nf, const float* scale_factors, const float* u, const float* v,
                                   const float* invz, const int32_t* last_octave, const float* last_angle, const uint8_t* valid,
                                   const int32_t* mp_nobs, const uint8_t* mpdesc, int nlast, float th, float mbf, int mode,
                                   int check_ori, int th_high, int32_t* match_out) {
    (void)kps_un;
    Frame Cur = grid->F;
    fill_frame(Cur, fdesc, fu_right, scale_factors);
    Cur.mbf = mbf; Cur.mb = 1.f;
    MapPointPool pool;
    seed_frame_mappoints(Cur.mvpMapPoints, frame_mp_obs, nf, pool);
    std::vector<MapPoint*> pts(nlast, static_cast<MapPoint*>(NULL));
    const int lm = mode & 7;
    const int nlevels = Cur.mnScaleLevels;
    if (mode & 8)            /* the KeyFrame overloads skip ANY occupied keypoint (:1541-1542, :372-373): every match blocks later points */
        for (int i = 0; i < nlast; i++)
            if (valid[i] && mp_nobs[i] != 1) return -3;
    for (int i = 0; i < nlast; i++) {
        if (!valid[i]) continue;
        float w[3];
        if (lm == 3) {
            /* the Sim3 overload rejects z < 0 and never uses 1/z: any depth reproduces (u, v) */
            w[2] = 1.f; w[0] = u[i]; w[1] = v[i];
        } else if (!unproject_exact(u[i], v[i], invz[i], w)) {
            return -2;
        }
        MapPoint* p = pool.make();
        p->worldPos = vec3(w[0], w[1], w[2]);
        p->nObs = mp_nobs[i];
        p->descriptor = desc_row(mpdesc + (size_t)i * 32);
        if (mode & 8) {
            /* these overloads predict the level from the distance (MapPoint::PredictScale): put the point's maximum
             * distance a quarter level below the boundary so that ceil(log(max/dist)/log(s)) == last_octave[i] */
            const float dist = (float)std::sqrt((double)w[0] * w[0] + (double)w[1] * w[1] + (double)w[2] * w[2]);
            const int L = last_octave[i];
            p->mfMaxDistance = dist * std::pow((double)Cur.mvScaleFactors[1], L - 0.25);
            if (L >= nlevels - 1) p->mfMaxDistance = dist * std::pow((double)Cur.mvScaleFactors[1], nlevels + 1.5);
            p->mfMinDistance = 0.f;
            p->normal = vec3(w[0] / dist, w[1] / dist, w[2] / dist);          /* seen head-on: passes the 60 degree gate */
        }
        pts[i] = p;
    }
    int n = 0;
    if (!(mode & 8)) {
        Frame Last;
        Last.N = nlast;
        Last.mvpMapPoints = pts;
        Last.mvbOutlier.assign(nlast, false);
        Last.mvKeys.resize(nlast); Last.mvKeysUn.resize(nlast);
        for (int i = 0; i < nlast; i++) {
            Last.mvKeys[i].octave = last_octave[i]; Last.mvKeysUn[i].octave = last_octave[i];
            Last.mvKeys[i].angle = last_angle[i]; Last.mvKeysUn[i].angle = last_angle[i];
        }
        /* tlc = Rlw*twc + tlw = tlw for an identity current pose: its z against mb (= 1) selects forward / backward (:1346-1349) */
        Last.mTcw = cv::Mat::eye(4, 4, CV_32F);
        Last.mTcw.at<float>(2, 3) = lm == 1 ? 2.f : (lm == 2 ? -2.f : 0.f);
        if (th_high != ORBmatcher::TH_HIGH) return -3;                          /* this overload hard-codes TH_HIGH (:1421) */
        ORBmatcher matcher(0.9f, check_ori != 0);
        n = matcher.SearchByProjection(Cur, Last, th, false);
    } else if (lm == 0) {
        KeyFrame KF;
        KF.N = nlast;
        KF.mapPoints = pts;
        KF.mvKeysUn.resize(nlast);
        for (int i = 0; i < nlast; i++) KF.mvKeysUn[i].angle = last_angle[i];
        ORBmatcher matcher(0.9f, check_ori != 0);
        std::set<MapPoint*> found;
        n = matcher.SearchByProjection(Cur, &KF, found, th, th_high);
    } else if (lm == 3) {
        if (th_high != ORBmatcher::TH_LOW || check_ori) return -3;              /* :389 uses TH_LOW, no rotation check */
        KeyFrame KF;
        keyframe_from_frame(KF, Cur);
        std::vector<MapPoint*> vpPoints, vpMatched(nf, static_cast<MapPoint*>(NULL));
        std::vector<int> src;
        for (int i = 0; i < nlast; i++)
            if (pts[i]) { vpPoints.push_back(pts[i]); src.push_back(i); }
        for (int k = 0; k < nf; k++) vpMatched[k] = Cur.mvpMapPoints[k];        /* :372-373: occupied slots are skipped */
        ORBmatcher matcher(0.9f, false);
        n = matcher.SearchByProjection(&KF, cv::Mat::eye(4, 4, CV_32F), vpPoints, vpMatched, (int)th);
        Cur.mvpMapPoints = vpMatched;
    } else {
        return -3;
    }
    for (int k = 0; k < nf; k++) {
        MapPoint* p = Cur.mvpMapPoints[k];
        frame_mp_obs[k] = p ? p->Observations() : 0;
        for (int i = 0; i < nlast && p; i++)
            if (pts[i] == p) { match_out[k] = i; break; }
    }
    return n;
}

/* ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:657-823 (+ CheckDistEpipolarLine :140-157) */
int ref_search_for_triangulation(const orc_keypoint* k1, const uint8_t* d1, const float* ur1, const uint8_t* has_mp1, int n1,
                                 const orc_keypoint* k2, const uint8_t* d2, const float* ur2, const uint8_t* has_mp2, int n2,
                                 const int32_t* node_id1, const int32_t* node_ptr1, const int32_t* idx1, int nn1,
                                 const int32_t* node_id2, const int32_t* node_ptr2, const int32_t* idx2, int nn2, const float* F12,
                                 float ex, float ey, const float* scale_factors2, const float* level_sigma2_2, int only_stereo,
                                 int check_ori, int32_t* matches12) {
    MapPointPool pool;
    KeyFrame K1, K2;
    K1.N = n1; K1.mvKeysUn = keys(k1, n1); K1.mDescriptors = desc_matrix(d1, n1); K1.mvuRight.assign(ur1, ur1 + n1);
    K2.N = n2; K2.mvKeysUn = keys(k2, n2); K2.mDescriptors = desc_matrix(d2, n2); K2.mvuRight.assign(ur2, ur2 + n2);
    K1.mapPoints.assign(n1, static_cast<MapPoint*>(NULL));
    K2.mapPoints.assign(n2, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < n1; i++) if (has_mp1[i]) K1.mapPoints[i] = pool.make();
    for (int i = 0; i < n2; i++) if (has_mp2[i]) K2.mapPoints[i] = pool.make();
    build_feature_vector(K1.mFeatVec, node_id1, node_ptr1, idx1, nn1);
    build_feature_vector(K2.mFeatVec, node_id2, node_ptr2, idx2, nn2);
    int nlevels = 0;
    for (int i = 0; i < n2; i++) nlevels = std::max(nlevels, K2.mvKeysUn[i].octave + 1);
    nlevels = std::max(nlevels, 8);
    K2.mvScaleFactors.assign(scale_factors2, scale_factors2 + nlevels);
    K2.mvLevelSigma2.assign(level_sigma2_2, level_sigma2_2 + nlevels);
    /* epipole of camera 1 in image 2 (:663-670): C2 = R2w*Cw + t2w with R2w = I, t2w = 0, Cw = (ex, ey, 1), unit intrinsics */
    K1.Ow = vec3(ex, ey, 1.f);
    K2.Rcw = cv::Mat::eye(3, 3, CV_32F);
    K2.tcw = vec3(0.f, 0.f, 0.f);
    K2.fx = K2.fy = 1.f; K2.cx = K2.cy = 0.f;
    cv::Mat F(3, 3, CV_32F);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) F.at<float>(r, c) = F12[3 * r + c];
    std::vector<std::pair<size_t, size_t> > pairs;
    ORBmatcher matcher(0.6f, check_ori != 0);
    const int n = matcher.SearchForTriangulation(&K1, &K2, F, pairs, only_stereo != 0);
    for (size_t i = 0; i < pairs.size(); i++) matches12[pairs[i].first] = (int32_t)pairs[i].second;
    return n;
}

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) :159-288 (mode 0) and SearchByBoW(KeyFrame*, KeyFrame*, ...) :522-655 (mode 1) */
int ref_search_by_bow(int mode, const orc_keypoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1, const orc_keypoint* k2,
                      const uint8_t* d2, const uint8_t* valid2, int n2, const int32_t* node_id1, const int32_t* node_ptr1,
                      const int32_t* idx1v, int nn1, const int32_t* node_id2, const int32_t* node_ptr2, const int32_t* idx2v, int nn2,
                      float mfNNratio, int check_ori, int32_t* match) {
    MapPointPool pool;
    KeyFrame K1;
    K1.N = n1; K1.mvKeysUn = keys(k1, n1); K1.mDescriptors = desc_matrix(d1, n1);
    K1.mapPoints.assign(n1, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < n1; i++) if (valid1[i]) K1.mapPoints[i] = pool.make();
    build_feature_vector(K1.mFeatVec, node_id1, node_ptr1, idx1v, nn1);
    ORBmatcher matcher(mfNNratio, check_ori != 0);
    const int nOut = mode == 0 ? n2 : n1;
    for (int i = 0; i < nOut; i++) match[i] = -1;
    if (mode == 0) {
        Frame F;
        F.N = n2; F.mvKeys = keys(k2, n2); F.mvKeysUn = F.mvKeys; F.mDescriptors = desc_matrix(d2, n2);
        build_feature_vector(F.mFeatVec, node_id2, node_ptr2, idx2v, nn2);
        std::vector<MapPoint*> out;
        const int n = matcher.SearchByBoW(&K1, F, out);
        for (int k = 0; k < n2; k++)
            for (int i = 0; i < n1 && out[k]; i++)
                if (K1.mapPoints[i] == out[k]) { match[k] = i; break; }
        return n;
    }
    KeyFrame K2;
    K2.N = n2; K2.mvKeysUn = keys(k2, n2); K2.mDescriptors = desc_matrix(d2, n2);
    K2.mapPoints.assign(n2, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < n2; i++) if (valid2[i]) K2.mapPoints[i] = pool.make();
    build_feature_vector(K2.mFeatVec, node_id2, node_ptr2, idx2v, nn2);
    std::vector<MapPoint*> out;
    const int n = matcher.SearchByBoW(&K1, &K2, out);
    for (int i = 0; i < n1; i++)
        for (int k = 0; k < n2 && out[i]; k++)
            if (K2.mapPoints[k] == out[i]) { match[i] = k; break; }
    return n;
}

/* ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:405-520 */
int ref_search_for_initialization(const RefGrid* grid2, const orc_keypoint* k2, const uint8_t* d2, int n2, const orc_keypoint* k1,
                                  const uint8_t* d1, int n1, float* prev, int windowSize, float mfNNratio, int check_ori,
                                  int32_t* vnMatches12) {
    (void)k2;
    Frame F2 = grid2->F;
    F2.mDescriptors = desc_matrix(d2, n2);
    Frame F1;
    F1.N = n1; F1.mvKeysUn = keys(k1, n1); F1.mvKeys = F1.mvKeysUn; F1.mDescriptors = desc_matrix(d1, n1);
    std::vector<cv::Point2f> vbPrevMatched(n1);
    for (int i = 0; i < n1; i++) vbPrevMatched[i] = cv::Point2f(prev[2 * i], prev[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher matcher(mfNNratio, check_ori != 0);
    const int n = matcher.SearchForInitialization(F1, F2, vbPrevMatched, m12, windowSize);
    for (int i = 0; i < n1; i++) {
        vnMatches12[i] = m12[i];
        prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y;
    }
    return n;
}

/* MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:249-314, for one point observed in N key frames.  The reference
 * iterates a std::map keyed by KeyFrame*: the key frames sit in one array here, so the iteration follows the row order. */
int ref_distinctive_descriptor(const uint8_t* desc, int N, int* median_out) {
    if (median_out) *median_out = INT_MAX;
    if (N <= 0) return -1;                   /* the reference returns early and leaves mDescriptor untouched (:263-264) */
    std::vector<KeyFrame> kfs(N);
    MapPoint mp;
    for (int i = 0; i < N; i++) {
        kfs[i].N = 1;
        kfs[i].mDescriptors = desc_matrix(desc + (size_t)i * 32, 1);
        mp.mObservations[&kfs[i]] = 0;
    }
    mp.ComputeDistinctiveDescriptors();
    int best = -1;
    for (int i = 0; i < N && best < 0; i++)
        if (memcmp(desc + (size_t)i * 32, mp.mDescriptor.data, 32) == 0) best = i;      /* identical rows have identical medians */
    if (best >= 0 && median_out) {
        std::vector<int> d(N);
        for (int j = 0; j < N; j++) d[j] = ORBmatcher::DescriptorDistance(desc_row(desc + (size_t)best * 32), desc_row(desc + (size_t)j * 32));
        std::sort(d.begin(), d.end());
        *median_out = d[(size_t)(0.5 * (N - 1))];
    }
    return best;
}

/* Frame::UndistortKeyPoints / ComputeImageBounds, src/Frame.cc:584-644 */
void ref_undistort_keypoints(const orc_keypoint* kps, int n, float fx, float fy, float cx, float cy, const float* dist, int ndist,
                             orc_keypoint* out) {
    Frame F;
    F.N = n;
    F.mvKeys = keys(kps, n);
    F.mK = cv::Mat::eye(3, 3, CV_32F);
    F.mK.at<float>(0, 0) = fx; F.mK.at<float>(1, 1) = fy; F.mK.at<float>(0, 2) = cx; F.mK.at<float>(1, 2) = cy;
    F.mDistCoef = cv::Mat(std::max(ndist, 4), 1, CV_32F, cv::Scalar(0));
    for (int i = 0; i < ndist; i++) F.mDistCoef.at<float>(i) = dist[i];
    F.UndistortKeyPoints();
    if (n > 0) memcpy(out, F.mvKeysUn.data(), (size_t)n * sizeof(orc_keypoint));
}
void ref_compute_image_bounds(int cols, int rows, float fx, float fy, float cx, float cy, const float* dist, int ndist, float* bounds) {
    Frame F;
    F.mK = cv::Mat::eye(3, 3, CV_32F);
    F.mK.at<float>(0, 0) = fx; F.mK.at<float>(1, 1) = fy; F.mK.at<float>(0, 2) = cx; F.mK.at<float>(1, 2) = cy;
    F.mDistCoef = cv::Mat(std::max(ndist, 4), 1, CV_32F, cv::Scalar(0));
    for (int i = 0; i < ndist; i++) F.mDistCoef.at<float>(i) = dist[i];
    cv::Mat im(rows, cols, CV_8UC1, (void*)&F, (size_t)cols);            /* only rows / cols are read (:622-625) */
    F.ComputeImageBounds(im);
    bounds[0] = F.mnMinX; bounds[1] = F.mnMaxX; bounds[2] = F.mnMinY; bounds[3] = F.mnMaxY;
}

/* ---- DBoW2: TemplatedVocabulary<FORB::TDescriptor, FORB>::transform, Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1138-1272 */
typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> RefVocabularyBase;
class RefVocabulary : public RefVocabularyBase {
public:
    void one(const cv::Mat& f, DBoW2::WordId& id, DBoW2::WordValue& w, DBoW2::NodeId* nid, int levelsup) const { transform(f, id, w, nid, levelsup); }
};

RefVocabulary* ref_vocabulary_create(int k, int L, int weighting, int scoring, int nnodes, const int32_t* parent, const uint8_t* desc,
                                     const double* weight) {
    /* through the reference's own loader (loadFromTextFile, :1351-1437): "k L scoring weighting" then one line per node
     * "parent isLeaf d0 .. d31 weight"; no trailing newline (the loader's eof loop would add a node for an empty line) */
    std::vector<int> children(nnodes, 0);
    for (int i = 1; i < nnodes; i++) children[parent[i]]++;
    char name[] = "/tmp/viorb_ref_voc_XXXXXX";
    const int fd = mkstemp(name);
    if (fd < 0) return NULL;
    FILE* f = fdopen(fd, "w");
    fprintf(f, "%d %d %d %d", k, L, scoring, weighting);
    for (int i = 1; i < nnodes; i++) {
        fprintf(f, "\n%d %d", parent[i], children[i] == 0 ? 1 : 0);
        for (int b = 0; b < 32; b++) fprintf(f, " %d", desc[(size_t)i * 32 + b]);
        fprintf(f, " %.17g", weight[i]);
    }
    fclose(f);
    RefVocabulary* v = new RefVocabulary();
    const bool ok = v->loadFromTextFile(name);
    remove(name);
    if (!ok) { delete v; return NULL; }
    return v;
}
void ref_vocabulary_destroy(RefVocabulary* v) { delete v; }

int ref_bow_transform(const RefVocabulary* voc, const uint8_t* desc, int n, int levelsup, int32_t* bow_ids, double* bow_values,
                      int32_t* fv_node, int32_t* fv_ptr, int32_t* fv_idx, int* nfv, int32_t* word_of, int32_t* node_of) {
    std::vector<cv::Mat> features(n);
    for (int i = 0; i < n; i++) features[i] = desc_row(desc + (size_t)i * 32);
    DBoW2::BowVector v;
    DBoW2::FeatureVector fv;
    voc->RefVocabularyBase::transform(features, v, fv, levelsup);
    bool undefinedNode = false;
    DBoW2::FeatureVector byFeature;
    for (int i = 0; i < n; i++) {
        DBoW2::WordId id = 0;
        DBoW2::WordValue w = 0;
        DBoW2::NodeId nid = 0;                   /* stays 0 where the reference leaves *nid unset (a leaf above the level) */
        if (n > 0 && !voc->empty()) voc->one(features[i], id, w, &nid, levelsup);
        if (word_of) word_of[i] = (int32_t)id;
        if (node_of) node_of[i] = (int32_t)nid;
        if (w > 0) {
            if (nid == 0 && levelsup > 0) undefinedNode = true;
            byFeature.addFeature(nid, (unsigned)i);
        }
    }
    /* transform(features, v, fv, levelsup) passes an UNINITIALISED NodeId to the per-feature transform, which only writes it
     * when the descent passes level L - levelsup (TemplatedVocabulary.h:1170-1185, :1230-1272): for a word that is a leaf
     * above that level the node in the reference's FeatureVector is whatever the stack held.  Where that happens the
     * FeatureVector is rebuilt from the per-feature calls with node 0 for those features (DESIGN.md convention C.8). */
    if (undefinedNode) fv = byFeature;
    int k = 0;
    for (DBoW2::BowVector::const_iterator it = v.begin(); it != v.end(); ++it, ++k) { bow_ids[k] = (int32_t)it->first; bow_values[k] = it->second; }
    int f = 0, pos = 0;
    fv_ptr[0] = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++f) {
        fv_node[f] = (int32_t)it->first;
        fv_ptr[f] = pos;
        for (size_t j = 0; j < it->second.size(); j++) fv_idx[pos++] = (int32_t)it->second[j];
    }
    fv_ptr[f] = pos;
    *nfv = f;
    return k;
}

}  // extern "C"
