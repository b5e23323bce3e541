/*
 * pose_scenarios.h -- deterministic scenario builders shared by tests/cpp/pose_driver.cc (results) and
 * tests/cpp/matcher_bench.cc (per-call timings).  Written against the members common to the stand-in classes of
 * viorb_b200/host/orbslam_compat.h, so the same code drives the CUDA drop-in and the reference's own ORBmatcher.cc.
 */
#ifndef VIORB_POSE_SCENARIOS_H
#define VIORB_POSE_SCENARIOS_H
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <vector>

#include "ORBmatcher.h"

using namespace ORB_SLAM2;

static uint32_t rng_state = 2463534242u;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 17; rng_state ^= rng_state << 5; return rng_state; }
static float rndf() { return (rnd() >> 8) * (1.0f / 16777216.0f); }

static const float FX = 718.856f, FY = 718.856f, CX = 607.1928f, CY = 185.2157f, BF = 386.1448f;      /* Examples/Stereo/KITTI00-02.yaml */
static const int W = 1241, H = 376, NLEVELS = 8;

struct Pose {
    double R[9], t[3];
};

static Pose random_pose(double maxAngle, double maxT) {
    double ax[3] = {rndf() - 0.5, rndf() - 0.5, rndf() - 0.5};
    const double n = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]) + 1e-9;
    for (int i = 0; i < 3; i++) ax[i] /= n;
    const double a = maxAngle * rndf(), c = std::cos(a), s = std::sin(a), C = 1 - c;
    Pose p;
    const double x = ax[0], y = ax[1], z = ax[2];
    const double R[9] = {c + x * x * C, x * y * C - z * s, x * z * C + y * s, y * x * C + z * s, c + y * y * C, y * z * C - x * s,
                         z * x * C - y * s, z * y * C + x * s, c + z * z * C};
    memcpy(p.R, R, sizeof(R));
    for (int i = 0; i < 3; i++) p.t[i] = maxT * (2 * rndf() - 1);
    return p;
}

static cv::Mat mat33(const double* R) {
    cv::Mat m(3, 3, CV_32F);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) m.at<float>(r, c) = (float)R[3 * r + c];
    return m;
}
static cv::Mat mat31(const double* t) {
    cv::Mat m(3, 1, CV_32F);
    for (int r = 0; r < 3; r++) m.at<float>(r, 0) = (float)t[r];
    return m;
}
static cv::Mat mat44(const Pose& p, double s = 1.0) {
    cv::Mat m(4, 4, CV_32F);
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) m.at<float>(r, c) = r == c ? 1.f : 0.f;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) m.at<float>(r, c) = (float)(s * p.R[3 * r + c]);
        m.at<float>(r, 3) = (float)(s * p.t[r]);
    }
    return m;
}
/* camera centre -R^T t, evaluated in double and rounded once: an INPUT of the searches (KeyFrame::GetCameraCenter) */
static cv::Mat centre(const Pose& p) {
    double o[3];
    for (int r = 0; r < 3; r++) o[r] = -(p.R[r] * p.t[0] + p.R[3 + r] * p.t[1] + p.R[6 + r] * p.t[2]);
    return mat31(o);
}

struct Keys {
    std::vector<cv::KeyPoint> k;
    cv::Mat d;
};

static bool read_keys(FILE* f, Keys& out) {
    int32_t n = 0;
    if (fread(&n, 4, 1, f) != 1) return false;
    out.k.resize(n);
    out.d = cv::Mat(n, 32, CV_8U);
    if (fread((void*)out.k.data(), 28, n, f) != (size_t)n) return false;
    if (fread(out.d.data, 32, n, f) != (size_t)n) return false;
    return true;
}

static std::vector<float> scale_factors() {
    std::vector<float> s(NLEVELS);
    s[0] = 1.f;
    for (int i = 1; i < NLEVELS; i++) s[i] = (float)(s[i - 1] * (double)1.2f);       /* src/ORBextractor.cc:419-423 */
    return s;
}

static void setup_frame(Frame& F, const Keys& K, const Pose& pose) {
    F.N = (int)K.k.size();
    F.mvKeys = K.k; F.mvKeysUn = K.k; F.mDescriptors = K.d;
    F.mvuRight.assign(F.N, -1.0f);
    for (int i = 0; i < F.N; i += 3) F.mvuRight[i] = K.k[i].pt.x - 12.0f - (float)(i % 7);
    F.mvpMapPoints.assign(F.N, static_cast<MapPoint*>(NULL));
    F.mvbOutlier.assign(F.N, false);
    F.mvScaleFactors = scale_factors();
    F.mvInvScaleFactors.resize(NLEVELS);
    for (int l = 0; l < NLEVELS; l++) F.mvInvScaleFactors[l] = 1.0f / F.mvScaleFactors[l];
    F.mnScaleLevels = NLEVELS;
    F.mfLogScaleFactor = std::log(1.2f);
    F.mnMinX = 0; F.mnMaxX = (float)W; F.mnMinY = 0; F.mnMaxY = (float)H;
    F.mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(F.mnMaxX - F.mnMinX);
    F.mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(F.mnMaxY - F.mnMinY);
    F.fx = FX; F.fy = FY; F.cx = CX; F.cy = CY; F.mbf = BF; F.mb = BF / FX;
    F.mTcw = mat44(pose);
    F.AssignFeaturesToGrid();
}

static void setup_keyframe(KeyFrame& K, const Keys& keys, const Pose& pose) {
    Frame F;
    setup_frame(F, keys, pose);
    K.N = F.N;
    K.mvKeysUn = F.mvKeysUn; K.mvuRight = F.mvuRight; K.mDescriptors = F.mDescriptors;
    K.mvScaleFactors = F.mvScaleFactors;
    K.mvLevelSigma2.resize(NLEVELS); K.mvInvLevelSigma2.resize(NLEVELS);
    for (int l = 0; l < NLEVELS; l++) { K.mvLevelSigma2[l] = K.mvScaleFactors[l] * K.mvScaleFactors[l]; K.mvInvLevelSigma2[l] = 1.0f / K.mvLevelSigma2[l]; }
    K.mnScaleLevels = NLEVELS; K.mfLogScaleFactor = F.mfLogScaleFactor;
    K.fx = FX; K.fy = FY; K.cx = CX; K.cy = CY; K.mbf = BF;
    K.mnMinX = F.mnMinX; K.mnMaxX = F.mnMaxX; K.mnMinY = F.mnMinY; K.mnMaxY = F.mnMaxY;
    K.mfGridElementWidthInv = F.mfGridElementWidthInv; K.mfGridElementHeightInv = F.mfGridElementHeightInv;
    K.mnGridCols = FRAME_GRID_COLS; K.mnGridRows = FRAME_GRID_ROWS;
    K.mGrid.resize(K.mnGridCols);                                                     /* src/KeyFrame.cc:301-306 */
    for (int i = 0; i < K.mnGridCols; i++) {
        K.mGrid[i].resize(K.mnGridRows);
        for (int j = 0; j < K.mnGridRows; j++) K.mGrid[i][j] = F.mGrid[i][j];
    }
    K.Rcw = mat33(pose.R); K.tcw = mat31(pose.t); K.Ow = centre(pose);
    K.mapPoints.assign(K.N, static_cast<MapPoint*>(NULL));
}

/* a map point that the camera `pose` sees near keypoint k of `keys` (pixel noise `du`), at a random depth */
static void make_point(MapPoint& mp, const Keys& keys, int k, const Pose& pose, float du, float tilt) {
    const double z = 4.0 + 26.0 * rndf();
    const double xc[3] = {(keys.k[k].pt.x + du * (rndf() - 0.5) * 2 - CX) / FX * z, (keys.k[k].pt.y + du * (rndf() - 0.5) * 2 - CY) / FY * z, z};
    double w[3];
    for (int r = 0; r < 3; r++) w[r] = pose.R[r] * (xc[0] - pose.t[0]) + pose.R[3 + r] * (xc[1] - pose.t[1]) + pose.R[6 + r] * (xc[2] - pose.t[2]);
    mp.worldPos = mat31(w);
    const double dist = std::sqrt(xc[0] * xc[0] + xc[1] * xc[1] + xc[2] * xc[2]);
    /* mean viewing direction: towards the point, tilted by up to `tilt` radians (beyond 60 degrees fails the gate) */
    double dir[3];
    for (int r = 0; r < 3; r++) dir[r] = pose.R[r] * xc[0] + pose.R[3 + r] * xc[1] + pose.R[6 + r] * xc[2];
    const double a = tilt * rndf();
    double nrm[3] = {dir[0] / dist * std::cos(a) + std::sin(a) * 0.6, dir[1] / dist * std::cos(a) - std::sin(a) * 0.8, dir[2] / dist * std::cos(a)};
    const double nn = std::sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
    for (int r = 0; r < 3; r++) nrm[r] /= nn;
    mp.normal = mat31(nrm);
    mp.mfMaxDistance = (float)(dist * std::pow(1.2, keys.k[k].octave + 0.9 * rndf() - 0.2));
    mp.mfMinDistance = mp.mfMaxDistance / (float)std::pow(1.2, NLEVELS - 1);
    mp.descriptor = keys.d.row(k).clone();
    const int flips = (int)(rnd() % 36);
    for (int b = 0; b < flips; b++) mp.descriptor.data[rnd() % 32] ^= (uint8_t)(1u << (rnd() % 8));
    mp.nObs = 1 + (int)(rnd() % 5);
    mp.bad = (rnd() % 23) == 0;
}

#endif
