/*
 * minicv.cpp -- the cv:: functions behind oracle/refbuild/minicv/minicv.hpp.  TEST INFRASTRUCTURE ONLY.
 *
 * The image primitives forward to the cv2-pinned restatements of oracle/orb_oracle.cpp (orc_resize_linear_u8, orc_fast9_16,
 * orc_gaussian7_u8_variant, orc_fast_atan2); the small-matrix algebra restates OpenCV core/matmul.cpp as verified on
 * cv2.gemm / cv2.norm in tests/test_ref_minicv.py.
 */
#include "minicv/minicv.hpp"

#include "../orb_oracle.h"
#include "minicv_hooks.h"

namespace cv {

static int g_gaussian_variant = 0;      /* 0 = OpenCV >= 3.4 fixed point taps, 1 = OpenCV 2.4 taps */
void minicv_set_gaussian_variant(int v) { g_gaussian_variant = v; }

/* optional record of every cv::FAST call (image origin, threshold, result): lets the test adapters reconstruct the
 * candidate lists that ComputeKeyPointsOctTree keeps in a local variable (src/ORBextractor.cc:778, :822-825) */
static thread_local std::vector<MinicvFastCall>* g_fast_log = nullptr;
void minicv_set_fast_log(std::vector<MinicvFastCall>* log) { g_fast_log = log; }

/* cv::gemm for the CV_32F shapes the reference produces (core/matmul.cpp).
 *  - flags == 0 and inner length 2..4: the hand-unrolled block -- float products summed left to right,
 *    d = (float)(t*alpha + c*beta) evaluated in double;
 *  - otherwise GEMMSingleMul<float,double>: products and the running sum in double, d = (float)(s*alpha + c*beta).
 * Both verified against cv2.gemm (inner length 3). */
void gemm(InputArray A_, InputArray B_, double alpha, InputArray C_, double beta, OutputArray D_, int flags) {
    Mat A = A_.getMat(), B = B_.getMat(), C = C_.getMat();
    if (A.depth() != B.depth() || (A.depth() != CV_32F && A.depth() != CV_64F)) throw Exception("minicv gemm: CV_32F / CV_64F only");
    const bool f32 = A.depth() == CV_32F;
    const bool ta = (flags & GEMM_1_T) != 0, tb = (flags & GEMM_2_T) != 0, tc = (flags & GEMM_3_T) != 0;
    const int M = ta ? A.cols : A.rows, K = ta ? A.rows : A.cols;
    const int Kb = tb ? B.cols : B.rows, N = tb ? B.rows : B.cols;
    if (K != Kb) throw Exception("minicv gemm: inner dimensions differ");
    if (C.empty()) beta = 0;
    Mat D(M, N, A.type());
    auto a = [&](int i, int k) { return ta ? A.getElem(k, i) : A.getElem(i, k); };
    auto b = [&](int k, int j) { return tb ? B.getElem(j, k) : B.getElem(k, j); };
    auto c = [&](int i, int j) { return C.empty() ? 0.0 : (tc ? C.getElem(j, i) : C.getElem(i, j)); };
    const bool smallBlock = f32 && flags == 0 && K >= 2 && K <= 4 && (K == N || K == M) && (K == N || N <= 16);
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            if (smallBlock) {
                float t = (float)a(i, 0) * (float)b(0, j);
                for (int k = 1; k < K; k++) t = t + (float)a(i, k) * (float)b(k, j);
                D.at<float>(i, j) = (float)((double)t * alpha + c(i, j) * beta);
            } else {
                double s = 0;
                for (int k = 0; k < K; k++) s += a(i, k) * b(k, j);
                const double v = C.empty() ? s * alpha : s * alpha + c(i, j) * beta;
                if (f32) D.at<float>(i, j) = (float)v; else D.at<double>(i, j) = v;
            }
        }
    D_.create(M, N, A.type());
    Mat out = D_.getMat();
    D.copyTo(out);
}

void transpose(InputArray src_, OutputArray dst_) {
    Mat s = src_.getMat();
    Mat d(s.cols, s.rows, s.type());
    const size_t esz = s.elemSize();
    for (int r = 0; r < s.rows; r++)
        for (int c = 0; c < s.cols; c++) memcpy(d.ptr(c) + (size_t)r * esz, s.ptr(r) + (size_t)c * esz, esz);
    dst_.create(d.rows, d.cols, d.type());
    Mat out = dst_.getMat();
    d.copyTo(out);
}

/* not on the hot path (the matcher never inverts): Gauss-Jordan in double */
double invert(InputArray src_, OutputArray dst_, int) {
    Mat s = src_.getMat();
    const int n = s.rows;
    std::vector<double> m((size_t)n * 2 * n, 0.0);
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) m[(size_t)i * 2 * n + j] = s.getElem(i, j);
        m[(size_t)i * 2 * n + n + i] = 1;
    }
    for (int i = 0; i < n; i++) {
        int p = i;
        for (int r = i + 1; r < n; r++) if (std::fabs(m[(size_t)r * 2 * n + i]) > std::fabs(m[(size_t)p * 2 * n + i])) p = r;
        if (m[(size_t)p * 2 * n + i] == 0) { dst_.create(n, n, s.type()); Mat z = dst_.getMat(); z.setTo(Scalar::all(0)); return 0; }
        for (int j = 0; j < 2 * n; j++) std::swap(m[(size_t)i * 2 * n + j], m[(size_t)p * 2 * n + j]);
        const double d = m[(size_t)i * 2 * n + i];
        for (int j = 0; j < 2 * n; j++) m[(size_t)i * 2 * n + j] /= d;
        for (int r = 0; r < n; r++) if (r != i) {
            const double f = m[(size_t)r * 2 * n + i];
            for (int j = 0; j < 2 * n; j++) m[(size_t)r * 2 * n + j] -= f * m[(size_t)i * 2 * n + j];
        }
    }
    dst_.create(n, n, s.type());
    Mat d = dst_.getMat();
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) d.setElem(i, j, m[(size_t)i * 2 * n + n + j]);
    return 1;
}

/* cv::norm: CV_32F accumulates in double (stat.cpp normL2_<float,double>, normL1_<float,double>); L2 returns sqrt */
double norm(InputArray a_, int normType) {
    Mat a = a_.getMat();
    const int n = a.cols * a.channels();
    double s = 0;
    for (int r = 0; r < a.rows; r++)
        for (int j = 0; j < n; j++) {
            const double v = a.getElem(r, j);
            if (normType == NORM_L2) s += v * v;
            else if (normType == NORM_L1) s += std::fabs(v);
            else s = std::max(s, std::fabs(v));
        }
    return normType == NORM_L2 ? std::sqrt(s) : s;
}
double norm(InputArray a_, InputArray b_, int normType) {
    Mat a = a_.getMat(), b = b_.getMat();
    const int n = a.cols * a.channels();
    double s = 0;
    for (int r = 0; r < a.rows; r++)
        for (int j = 0; j < n; j++) {
            double v;
            if (a.depth() == CV_32F) v = (double)(((const float*)a.ptr(r))[j] - ((const float*)b.ptr(r))[j]);   /* float difference first */
            else v = a.getElem(r, j) - b.getElem(r, j);
            if (normType == NORM_L2) s += v * v;
            else if (normType == NORM_L1) s += std::fabs(v);
            else s = std::max(s, std::fabs(v));
        }
    return normType == NORM_L2 ? std::sqrt(s) : s;
}

void resize(InputArray src_, OutputArray dst_, Size dsize, double fx, double fy, int interpolation) {
    Mat src = src_.getMat();
    if (src.type() != CV_8UC1 || interpolation != INTER_LINEAR) throw Exception("minicv resize: CV_8UC1 INTER_LINEAR only");
    if (dsize.area() == 0) dsize = Size(saturate_cast<int>(src.cols * fx), saturate_cast<int>(src.rows * fy));
    dst_.create(dsize, src.type());
    Mat dst = dst_.getMat();
    orc_resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

static inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

/* cv::copyMakeBorder (imgproc/utils.cpp / core copy.cpp) for CV_8UC1 and BORDER_REFLECT_101: without BORDER_ISOLATED a
 * source that is a submatrix takes its border from the surrounding pixels of the parent as far as they exist. */
void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType, const Scalar&) {
    Mat src = src_.getMat();
    if (src.type() != CV_8UC1 || (borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101) throw Exception("minicv copyMakeBorder: CV_8UC1 REFLECT_101 only");
    if (src.isSubmatrix() && (borderType & BORDER_ISOLATED) == 0) {
        Size wholeSize;
        Point ofs;
        src.locateROI(wholeSize, ofs);
        const int dtop = std::min(ofs.y, top), dbottom = std::min(wholeSize.height - src.rows - ofs.y, bottom);
        const int dleft = std::min(ofs.x, left), dright = std::min(wholeSize.width - src.cols - ofs.x, right);
        Mat grown = src;
        grown.data -= (size_t)dtop * src.step + dleft;
        grown.rows += dtop + dbottom; grown.cols += dleft + dright;
        src = grown;
        top -= dtop; left -= dleft; bottom -= dbottom; right -= dright;
    }
    dst_.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = dst_.getMat();
    /* the source may be the interior of dst (ORBextractor.cc:1122): stage it first */
    Mat tmp = src.clone();
    for (int y = 0; y < dst.rows; y++) {
        const uchar* S = tmp.ptr(reflect101(y - top, tmp.rows));
        uchar* D = dst.ptr(y);
        for (int x = 0; x < dst.cols; x++) D[x] = S[reflect101(x - left, tmp.cols)];
    }
}

void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double sigmaY, int borderType) {
    Mat src = src_.getMat();
    if (src.type() != CV_8UC1 || ksize.width != 7 || ksize.height != 7 || sigmaX != 2 || (sigmaY != 2 && sigmaY != 0) ||
        (borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101)
        throw Exception("minicv GaussianBlur: only the call of ORBextractor.cc:1086 (CV_8UC1, 7x7, sigma 2, REFLECT_101)");
    /* the reference blurs a clone, i.e. a continuous non-sub matrix: the border reflects at the level's edge */
    Mat in = src.clone();
    dst_.create(src.rows, src.cols, src.type());
    Mat dst = dst_.getMat();
    orc_gaussian7_u8_variant(in.data, in.cols, in.rows, in.step, dst.data, dst.step, g_gaussian_variant);
}

/* cv::FAST(image, keypoints, threshold, nms): FAST-9/16, KeyPoint(x, y, 7.f, -1, score) in row-major order */
void FAST(InputArray image_, std::vector<KeyPoint>& keypoints, int threshold, bool nms) {
    Mat img = image_.getMat();
    keypoints.clear();
    if (img.type() != CV_8UC1) throw Exception("minicv FAST: CV_8UC1 only");
    if (img.cols < 7 || img.rows < 7) return;
    std::vector<orc_corner> c((size_t)img.cols * img.rows);
    const int n = orc_fast9_16(img.data, img.cols, img.rows, img.step, threshold, nms ? 1 : 0, c.data(), (int)c.size());
    keypoints.reserve(n);
    for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)c[i].x, (float)c[i].y, 7.f, -1, (float)c[i].score));
    if (g_fast_log) {
        std::vector<MinicvFastCall>* log = g_fast_log;
        g_fast_log = nullptr;                       /* the record itself must not be logged or come from a test arena */
        MinicvFastCall call;
        call.origin = img.data; call.step = img.step; call.threshold = threshold;
        call.corners.assign(c.begin(), c.begin() + n);
        log->push_back(call);
        g_fast_log = log;
    }
}

float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

/* cv::undistortPoints (imgproc undistort: cvUndistortPoints): normalise with the inverse intrinsics, five fixed iterations of
 * the distortion compensation in double, re-project with P*R.  src/dst: N x 1 CV_32FC2 (or 1 x N). */
void undistortPoints(InputArray src_, OutputArray dst_, InputArray K_, InputArray dist_, InputArray R_, InputArray P_) {
    Mat src = src_.getMat(), K = K_.getMat(), dist = dist_.getMat(), R = R_.getMat(), P = P_.getMat();
    if (src.type() != CV_32FC2) throw Exception("minicv undistortPoints: CV_32FC2 points only");
    double k[12] = {0};
    const int nd = dist.empty() ? 0 : (int)dist.total();
    for (int i = 0; i < nd && i < 12; i++) k[i] = dist.depth() == CV_32F ? (double)dist.at<float>(i) : dist.at<double>(i);
    double A[3][3], RR[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) A[i][j] = K.getElem(i, j);
    if (!R.empty())
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) RR[i][j] = R.getElem(i, j);
    if (!P.empty()) {                       /* RR = PP * RR */
        double PP[3][3], T[3][3];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) PP[i][j] = P.getElem(i, j);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                double s = 0;
                for (int q = 0; q < 3; q++) s += PP[i][q] * RR[q][j];
                T[i][j] = s;
            }
        memcpy(RR, T, sizeof(T));
    }
    const double fx = A[0][0], fy = A[1][1], ifx = 1. / fx, ify = 1. / fy, cx = A[0][2], cy = A[1][2];
    const int n = (int)src.total();
    Mat in = src.clone();                  /* dst may be src (Frame.cc:602) */
    dst_.create(src.rows, src.cols, src.type());
    Mat dst = dst_.getMat();
    for (int i = 0; i < n; i++) {
        const float* s = in.rows == 1 ? in.ptr<float>(0) + 2 * i : in.ptr<float>(i);
        float* d = dst.rows == 1 ? dst.ptr<float>(0) + 2 * i : dst.ptr<float>(i);
        double x = s[0], y = s[1];
        const double u = x, v = y;
        x = (x - cx) * ifx;
        y = (y - cy) * ify;
        if (nd > 0) {
            const double x0 = x, y0 = y;
            for (int j = 0; j < 5; j++) {
                const double r2 = x * x + y * y;
                const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
                if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }      /* OpenCV >= 3.4 */
                const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
                const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
                x = (x0 - deltaX) * icdist;
                y = (y0 - deltaY) * icdist;
            }
        }
        const double xx = RR[0][0] * x + RR[0][1] * y + RR[0][2];
        const double yy = RR[1][0] * x + RR[1][1] * y + RR[1][2];
        const double ww = 1. / (RR[2][0] * x + RR[2][1] * y + RR[2][2]);
        d[0] = (float)(xx * ww);
        d[1] = (float)(yy * ww);
    }
}

}  // namespace cv
