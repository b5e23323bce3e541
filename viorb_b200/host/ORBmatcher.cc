/* ORBmatcher.cc -- see ORBmatcher.h.  Host-side marshalling only; every distance is computed on the GPU. */
#include "ORBmatcher.h"

#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <string>

#include "ORBextractor.h"
#include "viorb_gpu.h"

namespace ORB_SLAM2 {

const int ORBmatcher::TH_HIGH = 100;      /* reference src/ORBmatcher.cc:37-39 */
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

namespace {

/* OpenCV's arithmetic for the 3x3 / 3x1 CV_32F expressions of src/ORBmatcher.cc.  Each convention is pinned on the real
 * cv::gemm / cv::norm (python cv2) by tests/test_ref_minicv.py and on the reference compiled against those conventions by
 * tests/test_cpp_shims.py:
 *   R*x + t         one gemm, the hand-unrolled block of core/matmul.cpp: float products summed left to right, then + t
 *                   (written inline below as  R[0]*X + R[1]*Y + R[2]*Z + t[0]  -- same operations in the same order);
 *   -R.t()*t        gemm with GEMM_1_T and alpha = -1: products and the running sum in double, rounded once;
 *   M/s             MatExpr operator/ scales by 1./s: every element is multiplied by (float)(1./s), not divided;
 *   a.dot(b)        double products summed in double;  cv::norm(a) = sqrt of that in double. */
inline void neg_transpose_mul(const float R[9], const float t[3], float out[3]) {
    for (int r = 0; r < 3; r++) {
        double s = 0;
        s += (double)R[r] * (double)t[0];
        s += (double)R[3 + r] * (double)t[1];
        s += (double)R[6 + r] * (double)t[2];
        out[r] = (float)(s * -1.0);
    }
}
inline float reciprocal_scale(float s) { return (float)(1. / (double)s); }
inline double dot3(float ax, float ay, float az, float bx, float by, float bz) {
    double r = 0;
    r += (double)ax * (double)bx;
    r += (double)ay * (double)by;
    r += (double)az * (double)bz;
    return r;
}
inline float norm3(float x, float y, float z) { return (float)std::sqrt(dot3(x, y, z, x, y, z)); }

/* ORBmatcher::DescriptorDistance, reference src/ORBmatcher.cc:1648-1664 (a SWAR bit count on eight 32-bit words).  The
 * host files are built for a generic x86-64, so the POPCNT version is a separately targeted function chosen once at load
 * time; the fallback is the reference's SWAR count on four 64-bit words.  Both give the reference's value. */
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("popcnt"))) int hamming256_popcnt(const unsigned char* a, const unsigned char* b) {
    unsigned long long x[4], y[4];
    memcpy(x, a, 32);
    memcpy(y, b, 32);
    return (int)(__builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) + __builtin_popcountll(x[2] ^ y[2]) +
                 __builtin_popcountll(x[3] ^ y[3]));
}
#endif
int hamming256_swar(const unsigned char* a, const unsigned char* b) {
    unsigned long long dist = 0;
    for (int i = 0; i < 4; i++) {
        unsigned long long x, y;
        memcpy(&x, a + 8 * i, 8);
        memcpy(&y, b + 8 * i, 8);
        unsigned long long v = x ^ y;
        v = v - ((v >> 1) & 0x5555555555555555ull);
        v = (v & 0x3333333333333333ull) + ((v >> 2) & 0x3333333333333333ull);
        dist += (((v + (v >> 4)) & 0x0f0f0f0f0f0f0f0full) * 0x0101010101010101ull) >> 56;
    }
    return (int)dist;
}
typedef int (*Hamming256Fn)(const unsigned char*, const unsigned char*);
Hamming256Fn pick_hamming256() {
#if defined(__x86_64__) && defined(__GNUC__)
    __builtin_cpu_init();
    if (__builtin_cpu_supports("popcnt")) return hamming256_popcnt;
#endif
    return hamming256_swar;
}
const Hamming256Fn hamming256 = pick_hamming256();

void check(int rc, const char* what) {
    if (rc != VIORB_OK) throw std::runtime_error(std::string(what) + ": " + viorb_last_error());
}

/* one context per calling thread: the matcher is entered concurrently by Tracking, LocalMapping and LoopClosing */
viorb_ctx* thread_ctx() {
    struct Holder {
        viorb_ctx* c = nullptr;
        ~Holder() { viorb_ctx_destroy(c); }
    };
    static thread_local Holder h;
    if (!h.c) check(viorb_ctx_create(0, nullptr, &h.c), "viorb_ctx_create");
    return h.c;
}

struct FrameIndexGuard {
    viorb_frame_index* fi = nullptr;
    ~FrameIndexGuard() { viorb_frame_index_destroy(fi); }
};

std::vector<uint8_t> pack_descriptors(const cv::Mat& d, int n) {
    std::vector<uint8_t> out((size_t)n * 32);
    for (int i = 0; i < n; i++) memcpy(&out[(size_t)i * 32], d.ptr<uint8_t>(i), 32);
    return out;
}

void make_index(Frame& F, FrameIndexGuard& g) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(viorb_keypoint), "KeyPoint layout");
    const std::vector<uint8_t> desc = pack_descriptors(F.mDescriptors, F.N);
    check(viorb_frame_index_create(thread_ctx(), reinterpret_cast<const viorb_keypoint*>(F.mvKeysUn.data()), desc.data(),
                                   F.mvuRight.empty() ? nullptr : F.mvuRight.data(), F.N, F.mnMinX, F.mnMaxX, F.mnMinY,
                                   F.mnMaxY, F.mvScaleFactors.data(), (int)F.mvScaleFactors.size(), &g.fi),
          "viorb_frame_index_create");
}

/* DBoW2::FeatureVector (std::map<NodeId, vector<unsigned>>) -> node ids, CSR offsets, keypoint indices */
void flatten(const DBoW2::FeatureVector& fv, std::vector<int32_t>& ids, std::vector<int32_t>& ptr, std::vector<int32_t>& idx) {
    ptr.push_back(0);
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
        ids.push_back((int32_t)it->first);
        for (size_t k = 0; k < it->second.size(); k++) idx.push_back((int32_t)it->second[k]);
        ptr.push_back((int32_t)idx.size());
    }
}

}  // namespace

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

/* One pair is 32 bytes: the reference calls this inside O(N^2) host loops (MapPoint::ComputeDistinctiveDescriptors,
 * src/MapPoint.cc:289), where a device round trip per pair would cost four orders of magnitude more than the popcounts.
 * Batches go to the GPU through DescriptorDistances / viorb_descriptor_distance / viorb_distinctive_descriptors. */
int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    return hamming256(a.ptr<unsigned char>(), b.ptr<unsigned char>());
}

std::vector<int> ORBmatcher::DescriptorDistances(const cv::Mat& a, const cv::Mat& b) {
    const int n = a.rows;
    const std::vector<uint8_t> pa = pack_descriptors(a, n), pb = pack_descriptors(b, n);
    std::vector<int> d(n);
    check(viorb_descriptor_distance(thread_ctx(), pa.data(), pb.data(), n, d.data()), "viorb_descriptor_distance");
    return d;
}

int ORBmatcher::SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    const int nmp = (int)vpMapPoints.size();
    FrameIndexGuard g;
    make_index(F, g);
    std::vector<float> px(nmp), py(nmp), pxr(nmp), vc(nmp);
    std::vector<int32_t> lvl(nmp), nobs(nmp), obs(F.N), match(F.N, -1);
    std::vector<uint8_t> valid(nmp), desc((size_t)nmp * 32);
    for (int i = 0; i < nmp; i++) {
        MapPoint* p = vpMapPoints[i];
        valid[i] = p->mbTrackInView && !p->isBad();              /* :53-57 */
        px[i] = p->mTrackProjX; py[i] = p->mTrackProjY; pxr[i] = p->mTrackProjXR;
        lvl[i] = p->mnTrackScaleLevel; vc[i] = p->mTrackViewCos; nobs[i] = p->Observations();
        if (valid[i]) memcpy(&desc[(size_t)i * 32], p->GetDescriptor().ptr<uint8_t>(), 32);
    }
    for (int k = 0; k < F.N; k++) obs[k] = F.mvpMapPoints[k] ? F.mvpMapPoints[k]->Observations() : 0;   /* :87-89 */
    int n = 0;
    check(viorb_search_by_projection_local(g.fi, obs.data(), px.data(), py.data(), pxr.data(), lvl.data(), vc.data(),
                                           valid.data(), nobs.data(), desc.data(), nmp, th, mfNNratio, match.data(), &n),
          "viorb_search_by_projection_local");
    for (int k = 0; k < F.N; k++)
        if (match[k] >= 0) F.mvpMapPoints[k] = vpMapPoints[match[k]];                                    /* :122 */
    return n;
}

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
    /* projection of the last frame's map points with the current pose, reference :1338-1376 */
    const cv::Mat& Tcw = CurrentFrame.mTcw;
    const cv::Mat& Tlw = LastFrame.mTcw;
    float Rcw[9], tcw[3], twc[3], tlc[3];
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) Rcw[3 * r + c] = Tcw.at<float>(r, c);
        tcw[r] = Tcw.at<float>(r, 3);
    }
    neg_transpose_mul(Rcw, tcw, twc);                                                    /* -Rcw.t()*tcw, :1341 */
    for (int r = 0; r < 3; r++)
        tlc[r] = Tlw.at<float>(r, 0) * twc[0] + Tlw.at<float>(r, 1) * twc[1] + Tlw.at<float>(r, 2) * twc[2] + Tlw.at<float>(r, 3);
    const bool bForward = tlc[2] > CurrentFrame.mb && !bMono;
    const bool bBackward = -tlc[2] > CurrentFrame.mb && !bMono;
    const int mode = bForward ? 1 : (bBackward ? 2 : 0);

    const int nl = LastFrame.N;
    std::vector<float> u(nl), v(nl), invz(nl), ang(nl);
    std::vector<int32_t> oct(nl), nobs(nl), obs(CurrentFrame.N), match(CurrentFrame.N, -1);
    std::vector<uint8_t> valid(nl, 0), desc((size_t)nl * 32);
    for (int i = 0; i < nl; i++) {
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;
        const cv::Mat x3Dw = pMP->GetWorldPos();
        const float X = x3Dw.at<float>(0), Y = x3Dw.at<float>(1), Z = x3Dw.at<float>(2);
        const float xc = Rcw[0] * X + Rcw[1] * Y + Rcw[2] * Z + tcw[0];
        const float yc = Rcw[3] * X + Rcw[4] * Y + Rcw[5] * Z + tcw[1];
        const float zc = Rcw[6] * X + Rcw[7] * Y + Rcw[8] * Z + tcw[2];
        const float invzc = 1.0 / zc;
        if (invzc < 0) continue;
        u[i] = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
        v[i] = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
        invz[i] = invzc;
        oct[i] = LastFrame.mvKeys[i].octave;
        ang[i] = LastFrame.mvKeysUn[i].angle;
        nobs[i] = pMP->Observations();
        memcpy(&desc[(size_t)i * 32], pMP->GetDescriptor().ptr<uint8_t>(), 32);
        valid[i] = 1;
    }
    for (int k = 0; k < CurrentFrame.N; k++)
        obs[k] = CurrentFrame.mvpMapPoints[k] ? CurrentFrame.mvpMapPoints[k]->Observations() : 0;
    FrameIndexGuard g;
    make_index(CurrentFrame, g);
    int n = 0;
    check(viorb_search_by_projection_frame(g.fi, obs.data(), u.data(), v.data(), invz.data(), oct.data(), ang.data(),
                                           valid.data(), nobs.data(), desc.data(), nl, th, CurrentFrame.mbf, mode,
                                           mbCheckOrientation, TH_HIGH, match.data(), &n),
          "viorb_search_by_projection_frame");
    for (int k = 0; k < CurrentFrame.N; k++) {
        if (match[k] >= 0) CurrentFrame.mvpMapPoints[k] = LastFrame.mvpMapPoints[match[k]];
        /* matched, then rejected by the rotation histogram: the reference nulls the slot (:1460), also when it held a map
         * point without observations before the call (such a point does not block the search, :1403-1405) */
        else if (match[k] == -2) CurrentFrame.mvpMapPoints[k] = static_cast<MapPoint*>(NULL);
    }
    return n;
}


int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th,
                                   const int ORBdist) {
    /* projection and scale prediction on the host (reference :1477-1530), the search on the GPU */
    const cv::Mat& Tcw = CurrentFrame.mTcw;
    float Rcw[9], tcw[3], Ow[3];
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) Rcw[3 * r + c] = Tcw.at<float>(r, c);
        tcw[r] = Tcw.at<float>(r, 3);
    }
    neg_transpose_mul(Rcw, tcw, Ow);                                                     /* Ow = -Rcw.t()*tcw */
    const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
    const int nq = (int)vpMPs.size();
    std::vector<float> u(nq), v(nq), invz(nq), ang(nq);
    std::vector<int32_t> lvl(nq), nobs(nq, 1), obs(CurrentFrame.N), match(CurrentFrame.N, -1);
    std::vector<uint8_t> valid(nq, 0), desc((size_t)std::max(nq, 1) * 32);
    for (int i = 0; i < nq; i++) {
        MapPoint* pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
        const cv::Mat x3Dw = pMP->GetWorldPos();
        const float X = x3Dw.at<float>(0), Y = x3Dw.at<float>(1), Z = x3Dw.at<float>(2);
        const float xc = Rcw[0] * X + Rcw[1] * Y + Rcw[2] * Z + tcw[0];
        const float yc = Rcw[3] * X + Rcw[4] * Y + Rcw[5] * Z + tcw[1];
        const float zc = Rcw[6] * X + Rcw[7] * Y + Rcw[8] * Z + tcw[2];
        const float invzc = 1.0 / zc;
        const float uu = CurrentFrame.fx * xc * invzc + CurrentFrame.cx, vv = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
        if (uu < CurrentFrame.mnMinX || uu > CurrentFrame.mnMaxX) continue;
        if (vv < CurrentFrame.mnMinY || vv > CurrentFrame.mnMaxY) continue;
        const float px = X - Ow[0], py = Y - Ow[1], pz = Z - Ow[2];
        const float dist3D = norm3(px, py, pz);     /* cv::norm */
        if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
        u[i] = uu; v[i] = vv; invz[i] = invzc;
        lvl[i] = pMP->PredictScale(dist3D, &CurrentFrame);
        ang[i] = pKF->mvKeysUn[i].angle;
        memcpy(&desc[(size_t)i * 32], pMP->GetDescriptor().ptr<uint8_t>(), 32);
        valid[i] = 1;
    }
    for (int k = 0; k < CurrentFrame.N; k++) obs[k] = CurrentFrame.mvpMapPoints[k] ? 1 : 0;             /* :1541-1542 */
    FrameIndexGuard g;
    make_index(CurrentFrame, g);
    int n = 0;
    check(viorb_search_by_projection_frame(g.fi, obs.data(), u.data(), v.data(), invz.data(), lvl.data(), ang.data(),
                                           valid.data(), nobs.data(), desc.data(), nq, th, CurrentFrame.mbf, 0 | 8,
                                           mbCheckOrientation, ORBdist, match.data(), &n),
          "viorb_search_by_projection_frame");
    for (int k = 0; k < CurrentFrame.N; k++) {
        if (match[k] >= 0) CurrentFrame.mvpMapPoints[k] = vpMPs[match[k]];
        else if (match[k] == -2) CurrentFrame.mvpMapPoints[k] = static_cast<MapPoint*>(NULL);    /* :1586 */
    }
    return n;
}

int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints,
                                   std::vector<MapPoint*>& vpMatched, int th) {
    /* Sim3 decomposition, projection and gates on the host (reference :292-366), the search on the GPU */
    float sR[9], Rcw[9], tcw[3], Ow[3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) sR[3 * r + c] = Scw.at<float>(r, c);
    const float scw = (float)std::sqrt(dot3(sR[0], sR[1], sR[2], sR[0], sR[1], sR[2]));
    const float iscw = reciprocal_scale(scw);                                            /* sRcw/scw scales by 1./scw */
    for (int i = 0; i < 9; i++) Rcw[i] = sR[i] * iscw;
    for (int r = 0; r < 3; r++) tcw[r] = Scw.at<float>(r, 3) * iscw;
    neg_transpose_mul(Rcw, tcw, Ow);                                                     /* Ow = -Rcw.t()*tcw */
    std::set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint*>(nullptr));
    const int nq = (int)vpPoints.size();
    std::vector<float> u(nq), v(nq), invz(nq, 0.f), ang(nq, 0.f);
    std::vector<int32_t> lvl(nq), nobs(nq, 1), obs(pKF->N), match(pKF->N, -1);
    std::vector<uint8_t> valid(nq, 0), desc((size_t)std::max(nq, 1) * 32);
    for (int i = 0; i < nq; i++) {
        MapPoint* pMP = vpPoints[i];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        const cv::Mat p3Dw = pMP->GetWorldPos();
        const float X = p3Dw.at<float>(0), Y = p3Dw.at<float>(1), Z = p3Dw.at<float>(2);
        const float xc = Rcw[0] * X + Rcw[1] * Y + Rcw[2] * Z + tcw[0];
        const float yc = Rcw[3] * X + Rcw[4] * Y + Rcw[5] * Z + tcw[1];
        const float zc = Rcw[6] * X + Rcw[7] * Y + Rcw[8] * Z + tcw[2];
        if (zc < 0.0) continue;
        const float iz = 1 / zc;
        const float uu = pKF->fx * (xc * iz) + pKF->cx, vv = pKF->fy * (yc * iz) + pKF->cy;
        if (!pKF->IsInImage(uu, vv)) continue;
        const float px = X - Ow[0], py = Y - Ow[1], pz = Z - Ow[2];
        const float dist = norm3(px, py, pz);
        if (dist < pMP->GetMinDistanceInvariance() || dist > pMP->GetMaxDistanceInvariance()) continue;
        const cv::Mat Pn = pMP->GetNormal();
        if (dot3(px, py, pz, Pn.at<float>(0), Pn.at<float>(1), Pn.at<float>(2)) < 0.5 * dist) continue;   /* < 60 deg */
        u[i] = uu; v[i] = vv;
        lvl[i] = pMP->PredictScale(dist, pKF);
        memcpy(&desc[(size_t)i * 32], pMP->GetDescriptor().ptr<uint8_t>(), 32);
        valid[i] = 1;
    }
    for (int k = 0; k < pKF->N; k++) obs[k] = vpMatched[k] ? 1 : 0;                                      /* :372-373 */
    const std::vector<uint8_t> kdesc = pack_descriptors(pKF->mDescriptors, pKF->N);
    FrameIndexGuard g;
    check(viorb_frame_index_create(thread_ctx(), reinterpret_cast<const viorb_keypoint*>(pKF->mvKeysUn.data()), kdesc.data(),
                                   nullptr, pKF->N, pKF->mnMinX, pKF->mnMaxX, pKF->mnMinY, pKF->mnMaxY,
                                   pKF->mvScaleFactors.data(), (int)pKF->mvScaleFactors.size(), &g.fi),
          "viorb_frame_index_create");
    int n = 0;
    check(viorb_search_by_projection_frame(g.fi, obs.data(), u.data(), v.data(), invz.data(), lvl.data(), ang.data(),
                                           valid.data(), nobs.data(), desc.data(), nq, (float)th, 0.f, 3 | 8, 0, TH_LOW,
                                           match.data(), &n),
          "viorb_search_by_projection_frame");
    for (int k = 0; k < pKF->N; k++)
        if (match[k] >= 0) vpMatched[k] = vpPoints[match[k]];                                            /* :396 */
    return n;
}

int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12,
                                       std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo) {
    /* epipole in the second image, reference :663-670 */
    const cv::Mat Cw = pKF1->GetCameraCenter(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
    float C2[3];
    for (int r = 0; r < 3; r++)
        C2[r] = R2w.at<float>(r, 0) * Cw.at<float>(0) + R2w.at<float>(r, 1) * Cw.at<float>(1) + R2w.at<float>(r, 2) * Cw.at<float>(2) +
                t2w.at<float>(r);
    const float invz = 1.0f / C2[2];
    const float ex = pKF2->fx * C2[0] * invz + pKF2->cx;
    const float ey = pKF2->fy * C2[1] * invz + pKF2->cy;

    std::vector<int32_t> id1, p1, i1, id2, p2, i2;
    flatten(pKF1->mFeatVec, id1, p1, i1);
    flatten(pKF2->mFeatVec, id2, p2, i2);
    std::vector<uint8_t> mp1(pKF1->N), mp2(pKF2->N);
    for (int i = 0; i < pKF1->N; i++) mp1[i] = pKF1->GetMapPoint(i) != nullptr;
    for (int i = 0; i < pKF2->N; i++) mp2[i] = pKF2->GetMapPoint(i) != nullptr;
    const std::vector<uint8_t> d1 = pack_descriptors(pKF1->mDescriptors, pKF1->N), d2 = pack_descriptors(pKF2->mDescriptors, pKF2->N);
    float F[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) F[3 * r + c] = F12.at<float>(r, c);
    std::vector<int32_t> m12(pKF1->N, -1);
    int n = 0;
    check(viorb_search_for_triangulation(thread_ctx(), reinterpret_cast<const viorb_keypoint*>(pKF1->mvKeysUn.data()), d1.data(),
                                         pKF1->mvuRight.data(), mp1.data(), pKF1->N,
                                         reinterpret_cast<const viorb_keypoint*>(pKF2->mvKeysUn.data()), d2.data(),
                                         pKF2->mvuRight.data(), mp2.data(), pKF2->N, id1.data(), p1.data(), i1.data(),
                                         (int)id1.size(), id2.data(), p2.data(), i2.data(), (int)id2.size(), F, ex, ey,
                                         pKF2->mvScaleFactors.data(), pKF2->mvLevelSigma2.data(), (int)pKF2->mvScaleFactors.size(),
                                         bOnlyStereo, mbCheckOrientation, m12.data(), &n),
          "viorb_search_for_triangulation");
    vMatchedPairs.clear();
    vMatchedPairs.reserve(n);
    for (size_t i = 0; i < m12.size(); i++)
        if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)m12[i]));                     /* :815-820 */
    return n;
}

int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));                      /* :163 */
    std::vector<int32_t> id1, p1, i1, id2, p2, i2;
    flatten(pKF->mFeatVec, id1, p1, i1);
    flatten(F.mFeatVec, id2, p2, i2);
    std::vector<uint8_t> v1(pKF->N);
    for (int i = 0; i < pKF->N; i++) v1[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();            /* :190-196 */
    const std::vector<uint8_t> d1 = pack_descriptors(pKF->mDescriptors, pKF->N), d2 = pack_descriptors(F.mDescriptors, F.N);
    std::vector<int32_t> match(F.N, -1);
    int n = 0;
    check(viorb_search_by_bow(thread_ctx(), 0, reinterpret_cast<const viorb_keypoint*>(pKF->mvKeysUn.data()), d1.data(), v1.data(),
                              pKF->N, reinterpret_cast<const viorb_keypoint*>(F.mvKeys.data()), d2.data(), nullptr, F.N, id1.data(),
                              p1.data(), i1.data(), (int)id1.size(), id2.data(), p2.data(), i2.data(), (int)id2.size(), mfNNratio,
                              mbCheckOrientation, match.data(), &n),
          "viorb_search_by_bow");
    for (int k = 0; k < F.N; k++)
        if (match[k] >= 0) vpMapPointMatches[k] = vpMapPointsKF[match[k]];                              /* :220 */
    return n;
}

int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    vpMatches12 = std::vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));            /* :534 */
    std::vector<int32_t> id1, p1, i1, id2, p2, i2;
    flatten(pKF1->mFeatVec, id1, p1, i1);
    flatten(pKF2->mFeatVec, id2, p2, i2);
    std::vector<uint8_t> v1(pKF1->N), v2(pKF2->N);
    for (int i = 0; i < pKF1->N; i++) v1[i] = vpMapPoints1[i] && !vpMapPoints1[i]->isBad();
    for (int i = 0; i < pKF2->N; i++) v2[i] = vpMapPoints2[i] && !vpMapPoints2[i]->isBad();
    const std::vector<uint8_t> d1 = pack_descriptors(pKF1->mDescriptors, pKF1->N), d2 = pack_descriptors(pKF2->mDescriptors, pKF2->N);
    std::vector<int32_t> match(pKF1->N, -1);
    int n = 0;
    check(viorb_search_by_bow(thread_ctx(), 1, reinterpret_cast<const viorb_keypoint*>(pKF1->mvKeysUn.data()), d1.data(), v1.data(),
                              pKF1->N, reinterpret_cast<const viorb_keypoint*>(pKF2->mvKeysUn.data()), d2.data(), v2.data(), pKF2->N,
                              id1.data(), p1.data(), i1.data(), (int)id1.size(), id2.data(), p2.data(), i2.data(), (int)id2.size(),
                              mfNNratio, mbCheckOrientation, match.data(), &n),
          "viorb_search_by_bow");
    for (int i = 0; i < pKF1->N; i++)
        if (match[i] >= 0) vpMatches12[i] = vpMapPoints2[match[i]];                                     /* :601 */
    return n;
}

int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched,
                                        std::vector<int>& vnMatches12, int windowSize) {
    FrameIndexGuard g;
    make_index(F2, g);
    const int n1 = (int)F1.mvKeysUn.size();
    const std::vector<uint8_t> d1 = pack_descriptors(F1.mDescriptors, n1);
    static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "Point2f layout");
    vnMatches12 = std::vector<int>(n1, -1);
    int n = 0;
    check(viorb_search_for_initialization(g.fi, reinterpret_cast<const viorb_keypoint*>(F1.mvKeysUn.data()), d1.data(), n1,
                                          reinterpret_cast<float*>(vbPrevMatched.data()), windowSize, mfNNratio, mbCheckOrientation,
                                          vnMatches12.data(), &n),
          "viorb_search_for_initialization");
    return n;
}

namespace {

void make_kf_index(KeyFrame* pKF, FrameIndexGuard& g) {
    const std::vector<uint8_t> kdesc = pack_descriptors(pKF->mDescriptors, pKF->N);
    check(viorb_frame_index_create(thread_ctx(), reinterpret_cast<const viorb_keypoint*>(pKF->mvKeysUn.data()), kdesc.data(),
                                   pKF->mvuRight.empty() ? nullptr : pKF->mvuRight.data(), pKF->N, pKF->mnMinX, pKF->mnMaxX,
                                   pKF->mnMinY, pKF->mnMaxY, pKF->mvScaleFactors.data(), (int)pKF->mvScaleFactors.size(), &g.fi),
          "viorb_frame_index_create");
}

/* query arrays of one projected map-point set */
struct Queries {
    std::vector<float> u, v, ur;
    std::vector<int32_t> lvl;
    std::vector<uint8_t> valid, desc;
    explicit Queries(int n) : u(n), v(n), ur(n), lvl(n), valid(n, 0), desc((size_t)std::max(n, 1) * 32) {}
};

/* one direction of SearchBySim3 (:1149-1191 / :1228-1270): map points of `from`, camera pose (Rw, tw) of their key frame,
 * similarity (sR, t) into the other key frame `to` */
void sim3_queries(const std::vector<MapPoint*>& pts, const std::vector<bool>& already, const float* Rw, const float* tw,
                  const float* sR, const float* t, KeyFrame* to, float fx, float fy, float cx, float cy, Queries& q) {
    for (size_t i = 0; i < pts.size(); i++) {
        MapPoint* pMP = pts[i];
        if (!pMP || already[i]) continue;
        if (pMP->isBad()) continue;
        const cv::Mat p3Dw = pMP->GetWorldPos();
        const float X = p3Dw.at<float>(0), Y = p3Dw.at<float>(1), Z = p3Dw.at<float>(2);
        float a[3], b[3];
        for (int r = 0; r < 3; r++) a[r] = Rw[3 * r] * X + Rw[3 * r + 1] * Y + Rw[3 * r + 2] * Z + tw[r];
        for (int r = 0; r < 3; r++) b[r] = sR[3 * r] * a[0] + sR[3 * r + 1] * a[1] + sR[3 * r + 2] * a[2] + t[r];
        if (b[2] < 0.0) continue;                                     /* depth must be positive */
        const float invz = 1.0 / b[2];
        const float x = b[0] * invz, y = b[1] * invz;
        const float uu = fx * x + cx, vv = fy * y + cy;
        if (!to->IsInImage(uu, vv)) continue;
        const float dist3D = norm3(b[0], b[1], b[2]);
        if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
        q.u[i] = uu; q.v[i] = vv;
        q.lvl[i] = pMP->PredictScale(dist3D, to);
        memcpy(&q.desc[i * 32], pMP->GetDescriptor().ptr<uint8_t>(), 32);
        q.valid[i] = 1;
    }
}

}  // namespace

int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12,
                             const cv::Mat& R12, const cv::Mat& t12, const float th) {
    const float fx = pKF1->fx, fy = pKF1->fy, cx = pKF1->cx, cy = pKF1->cy;
    float R1w[9], t1w[3], R2w[9], t2w[3], sR12[9], sR21[9], t12f[3], t21[3];
    const cv::Mat mR1w = pKF1->GetRotation(), mt1w = pKF1->GetTranslation(), mR2w = pKF2->GetRotation(), mt2w = pKF2->GetTranslation();
    for (int r = 0; r < 3; r++) {
        t1w[r] = mt1w.at<float>(r); t2w[r] = mt2w.at<float>(r); t12f[r] = t12.at<float>(r);
        for (int c = 0; c < 3; c++) {
            R1w[3 * r + c] = mR1w.at<float>(r, c); R2w[3 * r + c] = mR2w.at<float>(r, c);
            sR12[3 * r + c] = s12 * R12.at<float>(r, c);                       /* :1119 */
            sR21[3 * r + c] = R12.at<float>(c, r) * (float)(1.0 / s12);        /* :1120: transpose, then scale by (float)(1./s12) */
        }
    }
    for (int r = 0; r < 3; r++) t21[r] = -(sR21[3 * r] * t12f[0] + sR21[3 * r + 1] * t12f[1] + sR21[3 * r + 2] * t12f[2]);   /* :1121 */
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
    std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
    for (int i = 0; i < N1; i++) {                                             /* :1132-1143 */
        MapPoint* pMP = vpMatches12[i];
        if (pMP) {
            vbAlreadyMatched1[i] = true;
            const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
        }
    }
    Queries q12(N1), q21(N2);
    sim3_queries(vpMapPoints1, vbAlreadyMatched1, R1w, t1w, sR21, t21, pKF2, fx, fy, cx, cy, q12);
    sim3_queries(vpMapPoints2, vbAlreadyMatched2, R2w, t2w, sR12, t12f, pKF1, fx, fy, cx, cy, q21);
    FrameIndexGuard g1, g2;
    make_kf_index(pKF1, g1);
    make_kf_index(pKF2, g2);
    std::vector<int32_t> match(std::max(N1, 1), -1);
    int nFound = 0;
    check(viorb_search_by_sim3(g1.fi, g2.fi, q12.u.data(), q12.v.data(), q12.lvl.data(), q12.valid.data(), q12.desc.data(),
                               q21.u.data(), q21.v.data(), q21.lvl.data(), q21.valid.data(), q21.desc.data(), th, match.data(), &nFound),
          "viorb_search_by_sim3");
    for (int i1 = 0; i1 < N1; i1++)
        if (match[i1] >= 0) vpMatches12[i1] = vpMapPoints2[match[i1]];         /* :1315 */
    return nFound;
}

int ORBmatcher::Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    const cv::Mat mRcw = pKF->GetRotation(), mtcw = pKF->GetTranslation(), mOw = pKF->GetCameraCenter();
    float Rcw[9], tcw[3], Ow[3];
    for (int r = 0; r < 3; r++) {
        tcw[r] = mtcw.at<float>(r); Ow[r] = mOw.at<float>(r);
        for (int c = 0; c < 3; c++) Rcw[3 * r + c] = mRcw.at<float>(r, c);
    }
    const float fx = pKF->fx, fy = pKF->fy, cx = pKF->cx, cy = pKF->cy, bf = pKF->mbf;
    const int nMPs = (int)vpMapPoints.size();
    /* projection gates of :846-881 for every point that is not bad now; isBad / IsInKeyFrame are re-checked in order below */
    auto project = [&](MapPoint* pMP, Queries& q, int i) {
        q.valid[i] = 0;
        const cv::Mat p3Dw = pMP->GetWorldPos();
        const float X = p3Dw.at<float>(0), Y = p3Dw.at<float>(1), Z = p3Dw.at<float>(2);
        float p[3];
        for (int r = 0; r < 3; r++) p[r] = Rcw[3 * r] * X + Rcw[3 * r + 1] * Y + Rcw[3 * r + 2] * Z + tcw[r];
        if (p[2] < 0.0f) return;
        const float invz = 1 / p[2];
        const float x = p[0] * invz, y = p[1] * invz;
        const float u = fx * x + cx, v = fy * y + cy;
        if (!pKF->IsInImage(u, v)) return;
        const float ur = u - bf * invz;
        const float px = X - Ow[0], py = Y - Ow[1], pz = Z - Ow[2];
        const float dist3D = norm3(px, py, pz);
        if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) return;
        const cv::Mat Pn = pMP->GetNormal();
        if (dot3(px, py, pz, Pn.at<float>(0), Pn.at<float>(1), Pn.at<float>(2)) < 0.5 * dist3D) return;
        q.u[i] = u; q.v[i] = v; q.ur[i] = ur;
        q.lvl[i] = pMP->PredictScale(dist3D, pKF);
        memcpy(&q.desc[(size_t)i * 32], pMP->GetDescriptor().ptr<uint8_t>(), 32);
        q.valid[i] = 1;
    };
    Queries q(nMPs);
    for (int i = 0; i < nMPs; i++) {
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP || pMP->isBad()) continue;
        project(pMP, q, i);
    }
    FrameIndexGuard g;
    make_kf_index(pKF, g);
    std::vector<int32_t> best(std::max(nMPs, 1), -1);
    check(viorb_search_window_top1(g.fi, q.u.data(), q.v.data(), q.ur.data(), q.lvl.data(), q.valid.data(), q.desc.data(), nMPs, th,
                                   TH_LOW, pKF->mvInvLevelSigma2.data(), best.data(), nullptr),
          "viorb_search_window_top1");
    int nFused = 0;
    for (int i = 0; i < nMPs; i++) {
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP) continue;
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;                   /* :842-843, in loop order */
        int bestIdx = best[i];
        const cv::Mat dNow = pMP->GetDescriptor();
        if (q.valid[i] && memcmp(dNow.ptr<uint8_t>(), &q.desc[(size_t)i * 32], 32) != 0) {
            /* an earlier Replace recomputed this point's descriptor (MapPoint.cc:221): search it again on its own */
            Queries q1(1);
            project(pMP, q1, 0);
            int32_t b1 = -1;
            if (q1.valid[0])
                check(viorb_search_window_top1(g.fi, q1.u.data(), q1.v.data(), q1.ur.data(), q1.lvl.data(), q1.valid.data(), q1.desc.data(), 1,
                                               th, TH_LOW, pKF->mvInvLevelSigma2.data(), &b1, nullptr),
                      "viorb_search_window_top1");
            bestIdx = b1;
        }
        if (bestIdx < 0) continue;
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);                          /* :945-967 */
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, bestIdx);
            pKF->AddMapPoint(pMP, bestIdx);
        }
        nFused++;
    }
    return nFused;
}

int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint) {
    float sR[9], Rcw[9], tcw[3], Ow[3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) sR[3 * r + c] = Scw.at<float>(r, c);
    const float scw = (float)std::sqrt(dot3(sR[0], sR[1], sR[2], sR[0], sR[1], sR[2]));     /* :989 */
    const float iscw = reciprocal_scale(scw);                                            /* sRcw/scw scales by 1./scw */
    for (int i = 0; i < 9; i++) Rcw[i] = sR[i] * iscw;
    for (int r = 0; r < 3; r++) tcw[r] = Scw.at<float>(r, 3) * iscw;
    neg_transpose_mul(Rcw, tcw, Ow);                                                     /* Ow = -Rcw.t()*tcw */
    const std::set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();                                                /* :995 */
    const int nPoints = (int)vpPoints.size();
    Queries q(nPoints);
    for (int i = 0; i < nPoints; i++) {
        MapPoint* pMP = vpPoints[i];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        const cv::Mat p3Dw = pMP->GetWorldPos();
        const float X = p3Dw.at<float>(0), Y = p3Dw.at<float>(1), Z = p3Dw.at<float>(2);
        float p[3];
        for (int r = 0; r < 3; r++) p[r] = Rcw[3 * r] * X + Rcw[3 * r + 1] * Y + Rcw[3 * r + 2] * Z + tcw[r];
        if (p[2] < 0.0f) continue;
        const float invz = 1.0 / p[2];
        const float u = pKF->fx * (p[0] * invz) + pKF->cx, v = pKF->fy * (p[1] * invz) + pKF->cy;
        if (!pKF->IsInImage(u, v)) continue;
        const float px = X - Ow[0], py = Y - Ow[1], pz = Z - Ow[2];
        const float dist3D = norm3(px, py, pz);
        if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
        const cv::Mat Pn = pMP->GetNormal();
        if (dot3(px, py, pz, Pn.at<float>(0), Pn.at<float>(1), Pn.at<float>(2)) < 0.5 * dist3D) continue;
        q.u[i] = u; q.v[i] = v;
        q.lvl[i] = pMP->PredictScale(dist3D, pKF);
        memcpy(&q.desc[(size_t)i * 32], pMP->GetDescriptor().ptr<uint8_t>(), 32);
        q.valid[i] = 1;
    }
    FrameIndexGuard g;
    make_kf_index(pKF, g);
    std::vector<int32_t> best(std::max(nPoints, 1), -1);
    check(viorb_search_window_top1(g.fi, q.u.data(), q.v.data(), nullptr, q.lvl.data(), q.valid.data(), q.desc.data(), nPoints, th, TH_LOW,
                                   nullptr, best.data(), nullptr),
          "viorb_search_window_top1");
    int nFused = 0;
    for (int iMP = 0; iMP < nPoints; iMP++) {                                   /* :1076-1094: no state read by later searches */
        if (best[iMP] < 0) continue;
        MapPoint* pMP = vpPoints[iMP];
        MapPoint* pMPinKF = pKF->GetMapPoint(best[iMP]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) vpReplacePoint[iMP] = pMPinKF;
        } else {
            pMP->AddObservation(pKF, best[iMP]);
            pKF->AddMapPoint(pMP, best[iMP]);
        }
        nFused++;
    }
    return nFused;
}

#ifndef VIORB_USE_ORBSLAM_HEADERS
/* ---- the members of Frame on the ORB path (reference src/Frame.cc); inside the VIORB tree paste these bodies into
 * src/Frame.cc ---- */
void Frame::ExtractORB(int flag, const cv::Mat& im) {
    if (flag == 0) (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors);
    else (*mpORBextractorRight)(im, cv::Mat(), mvKeysRight, mDescriptorsRight);
}

namespace {
void distortion_of(const Frame& F, float k[12], int& nk) {
    nk = F.mDistCoef.rows * F.mDistCoef.cols;
    if (nk > 12) nk = 12;
    for (int i = 0; i < nk; i++) k[i] = F.mDistCoef.rows >= F.mDistCoef.cols ? F.mDistCoef.at<float>(i, 0) : F.mDistCoef.at<float>(0, i);
}
}  // namespace

void Frame::UndistortKeyPoints() {
    float k[12];
    int nk = 0;
    distortion_of(*this, k, nk);
    mvKeysUn.resize(N);
    if (N == 0) return;
    check(viorb_undistort_keypoints(thread_ctx(), reinterpret_cast<const viorb_keypoint*>(mvKeys.data()), N, mK.at<float>(0, 0),
                                    mK.at<float>(1, 1), mK.at<float>(0, 2), mK.at<float>(1, 2), k, nk,
                                    reinterpret_cast<viorb_keypoint*>(mvKeysUn.data())),
          "viorb_undistort_keypoints");
}

void Frame::ComputeImageBounds(const cv::Mat& imLeft) {
    float k[12], b[4];
    int nk = 0;
    distortion_of(*this, k, nk);
    check(viorb_compute_image_bounds(thread_ctx(), imLeft.cols, imLeft.rows, mK.at<float>(0, 0), mK.at<float>(1, 1), mK.at<float>(0, 2),
                                     mK.at<float>(1, 2), k, nk, b),
          "viorb_compute_image_bounds");
    mnMinX = b[0]; mnMaxX = b[1]; mnMinY = b[2]; mnMaxY = b[3];
    mfGridElementWidthInv = 64.0f / (mnMaxX - mnMinX);              /* :180-183 */
    mfGridElementHeightInv = 48.0f / (mnMaxY - mnMinY);
}

void Frame::AssignFeaturesToGrid() {
    for (int i = 0; i < 64; i++)
        for (int j = 0; j < 48; j++) mGrid[i][j].clear();
    if (N == 0) return;
    FrameIndexGuard g;
    make_index(*this, g);
    std::vector<int32_t> start(64 * 48 + 1), items(N);
    check(viorb_frame_index_grid(g.fi, start.data(), items.data()), "viorb_frame_index_grid");
    for (int c = 0; c < 64 * 48; c++)
        mGrid[c / 48][c % 48].assign(items.begin() + start[c], items.begin() + start[c + 1]);
}

std::vector<size_t> Frame::GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel, const int maxLevel) const {
    std::vector<size_t> out;
    if (N == 0) return out;
    FrameIndexGuard g;
    make_index(const_cast<Frame&>(*this), g);
    std::vector<int32_t> idx(N);
    int n = 0;
    check(viorb_frame_features_in_area(g.fi, x, y, r, minLevel, maxLevel, idx.data(), N, &n), "viorb_frame_features_in_area");
    out.assign(idx.begin(), idx.begin() + n);
    return out;
}

/* Frame::ComputeStereoMatches (reference src/Frame.cc:646-820) on the pyramids resident in the two extractors */
void Frame::ComputeStereoMatches() {
    mvuRight.assign(N, -1.0f);
    mvDepth.assign(N, -1.0f);
    if (N == 0) return;
    const std::vector<uint8_t> dl = pack_descriptors(mDescriptors, N), dr = pack_descriptors(mDescriptorsRight, (int)mvKeysRight.size());
    check(viorb_stereo_match(mpORBextractorLeft->Handle(), 0, mpORBextractorRight->Handle(), 0,
                             reinterpret_cast<const viorb_keypoint*>(mvKeys.data()), dl.data(), N,
                             reinterpret_cast<const viorb_keypoint*>(mvKeysRight.data()), dr.data(), (int)mvKeysRight.size(), mbf, mb,
                             mvuRight.data(), mvDepth.data()),
          "viorb_stereo_match");
}
#endif

}  // namespace ORB_SLAM2
