/*
 * ref_minicv_capi.cpp -- flat C entry points onto the stand-in cv:: functions (minicv), so tests/test_ref_minicv.py can pin
 * each of them on the real OpenCV that runs in this image (python cv2 4.13).  TEST INFRASTRUCTURE ONLY.
 */
#include <opencv2/core/core.hpp>

#include <cstdint>
#include <vector>

extern "C" {

void ref_cv_resize(const uint8_t* src, int sh, int sw, size_t sstep, uint8_t* dst, int dh, int dw, size_t dstep) {
    cv::Mat s(sh, sw, CV_8UC1, (void*)src, sstep), d(dh, dw, CV_8UC1, dst, dstep);
    cv::resize(s, d, cv::Size(dw, dh), 0, 0, cv::INTER_LINEAR);
}

/* src is the view [y0, y0+h) x [x0, x0+w) of a parent image: without BORDER_ISOLATED OpenCV borrows the parent's pixels */
void ref_cv_copy_make_border(const uint8_t* parent, int ph, int pw, size_t pstep, int x0, int y0, int w, int h, int border,
                             int isolated, uint8_t* dst, size_t dstep) {
    cv::Mat p(ph, pw, CV_8UC1, (void*)parent, pstep);
    cv::Mat view = p(cv::Rect(x0, y0, w, h));
    cv::Mat d(h + 2 * border, w + 2 * border, CV_8UC1, dst, dstep);
    cv::copyMakeBorder(view, d, border, border, border, border, cv::BORDER_REFLECT_101 + (isolated ? cv::BORDER_ISOLATED : 0));
}

/* the in-place form of src/ORBextractor.cc:1122: source = interior of the destination */
void ref_cv_copy_make_border_inplace(uint8_t* padded, int h, int w, size_t step, int border) {
    cv::Mat temp(h + 2 * border, w + 2 * border, CV_8UC1, padded, step);
    cv::Mat roi = temp(cv::Rect(border, border, w, h));
    cv::copyMakeBorder(roi, temp, border, border, border, border, cv::BORDER_REFLECT_101 + cv::BORDER_ISOLATED);
}

void ref_cv_gaussian7(const uint8_t* src, int h, int w, size_t sstep, uint8_t* dst, size_t dstep) {
    cv::Mat s(h, w, CV_8UC1, (void*)src, sstep), d(h, w, CV_8UC1, dst, dstep);
    cv::Mat work = s.clone();
    cv::GaussianBlur(work, work, cv::Size(7, 7), 2, 2, cv::BORDER_REFLECT_101);
    work.copyTo(d);
}

int ref_cv_fast(const uint8_t* img, int h, int w, size_t step, int threshold, int nms, float* xyr, int cap) {
    cv::Mat m(h, w, CV_8UC1, (void*)img, step);
    std::vector<cv::KeyPoint> k;
    cv::FAST(m, k, threshold, nms != 0);
    for (size_t i = 0; i < k.size() && (int)i < cap; i++) { xyr[3 * i] = k[i].pt.x; xyr[3 * i + 1] = k[i].pt.y; xyr[3 * i + 2] = k[i].response; }
    return (int)k.size();
}

float ref_cv_fast_atan2(float y, float x) { return cv::fastAtan2(y, x); }
int ref_cv_round(double v) { return cvRound(v); }

/* D = alpha*op(A)*op(B) + beta*op(C), CV_32F, row-major dense inputs */
void ref_cv_gemm(const float* A, int ar, int ac, const float* B, int br, int bc, double alpha, const float* C, int cr, int cc,
                 double beta, int flags, float* D) {
    cv::Mat a(ar, ac, CV_32F, (void*)A), b(br, bc, CV_32F, (void*)B), c, d;
    if (C) c = cv::Mat(cr, cc, CV_32F, (void*)C);
    cv::gemm(a, b, alpha, c, beta, d, flags);
    for (int i = 0; i < d.rows; i++)
        for (int j = 0; j < d.cols; j++) D[i * d.cols + j] = d.at<float>(i, j);
}

double ref_cv_norm(const float* a, int n, int type) { return cv::norm(cv::Mat(n, 1, CV_32F, (void*)a), type); }
double ref_cv_norm_diff(const float* a, const float* b, int rows, int cols, int type) {
    return cv::norm(cv::Mat(rows, cols, CV_32F, (void*)a), cv::Mat(rows, cols, CV_32F, (void*)b), type);
}
double ref_cv_dot(const float* a, const float* b, int n) { return cv::Mat(1, n, CV_32F, (void*)a).dot(cv::Mat(1, n, CV_32F, (void*)b)); }

void ref_cv_undistort(const float* pts, int n, const float* K, const float* dist, int nd, float* out) {
    cv::Mat mat(n, 2, CV_32F);
    for (int i = 0; i < n; i++) { mat.at<float>(i, 0) = pts[2 * i]; mat.at<float>(i, 1) = pts[2 * i + 1]; }
    cv::Mat k(3, 3, CV_32F, (void*)K), d(nd, 1, CV_32F, (void*)dist);
    mat = mat.reshape(2);
    cv::undistortPoints(mat, mat, k, d, cv::Mat(), k);           /* the call of src/Frame.cc:602 */
    mat = mat.reshape(1);
    for (int i = 0; i < n; i++) { out[2 * i] = mat.at<float>(i, 0); out[2 * i + 1] = mat.at<float>(i, 1); }
}

/* the expression forms of src/ORBmatcher.cc, evaluated through the stand-in MatExpr:
 *   form 0: R*x + t (:324,:1361)   form 1: -R.t()*t (:303,:1341)   form 2: M/s (:301)   form 3: s*R (:1119)
 *   form 4: (1.0/s)*R.t() (:1120)  form 5: -M*t (:1121)            form 6: x - y (:345) */
void ref_cv_expr(int form, const float* R, const float* x, const float* t, float s, float* out) {
    cv::Mat Rm(3, 3, CV_32F, (void*)R), xm(3, 1, CV_32F, (void*)x), tm(3, 1, CV_32F, (void*)t), r;
    switch (form) {
        case 0: r = Rm * xm + tm; break;
        case 1: r = -Rm.t() * tm; break;
        case 2: r = Rm / s; break;
        case 3: r = s * Rm; break;
        case 4: r = (1.0 / s) * Rm.t(); break;
        case 5: r = -Rm * tm; break;
        default: r = xm - tm; break;
    }
    for (int i = 0; i < r.rows; i++)
        for (int j = 0; j < r.cols; j++) out[i * r.cols + j] = r.at<float>(i, j);
}

}  // extern "C"
