/*
 * minicv.hpp -- stand-in for the OpenCV headers the reference includes, so that the reference's own source files
 * (/root/reference/src/ORBextractor.cc, src/ORBmatcher.cc, parts of src/Frame.cc, src/KeyFrame.cc, src/MapPoint.cc and
 * Thirdparty/DBoW2) compile UNMODIFIED in this image, which has no OpenCV C++ headers or libraries.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/): nothing under viorb_b200/ includes or links this.
 *
 * OpenCV is a third-party dependency of the reference that is not vendored (CMakeLists.txt:31 pins
 * find_package(OpenCV 2.4.3), README.md:58 "tested with 2.4.11").  What is restated here is the published behaviour of the
 * cv:: types and functions the hot path calls; every arithmetic primitive is pinned bit for bit on the OpenCV that is
 * runnable here (python cv2 4.13) by tests/test_ref_minicv.py:
 *   cv::resize INTER_LINEAR 8u, cv::copyMakeBorder REFLECT_101 (+ISOLATED), cv::FAST 9/16 + NMS, cv::GaussianBlur 7x7
 *   sigma 2 (4.x taps by default, the 2.4 taps selectable), cv::fastAtan2, cvRound, cv::gemm on small CV_32F matrices
 *   (the plain path sums float products left to right, the transposed path accumulates in double -- both verified on
 *   cv2.gemm), cv::norm, Mat::dot, cv::undistortPoints.
 * Expression semantics follow OpenCV's MatExpr: A*B+C is ONE gemm, A/s multiplies by (float)(1./s), -A.t()*B is a gemm with
 * GEMM_1_T and alpha=-1.
 */
#ifndef VIORB_MINICV_HPP
#define VIORB_MINICV_HPP

#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef signed char schar;
typedef int64_t int64;
typedef uint64_t uint64;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 63) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795
#define CV_Assert(expr)                                                                          \
    do {                                                                                         \
        if (!(expr)) throw cv::Exception(std::string("CV_Assert failed: ") + #expr);             \
    } while (0)

/* cvRound/cvFloor/cvCeil as OpenCV defines them on x86-64 (cvtsd2si = round half to even in the default mode) */
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

class Exception : public std::runtime_error {
public:
    explicit Exception(const std::string& m) : std::runtime_error(m) {}
};

using std::string;
typedef std::string String;

template <typename T> inline T saturate_cast(double v) { return (T)v; }
template <> inline uchar saturate_cast<uchar>(double v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> operator Point_<U>() const { return Point_<U>(saturate_cast<U>(x), saturate_cast<U>(y)); }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
/* OpenCV: a.x = saturate_cast<T>(a.x*b), one overload per scalar type (core/operations.hpp) */
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, int b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, float b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, double b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> inline bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }

template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
};
typedef Size_<int> Size;
template <typename T> inline bool operator==(const Size_<T>& a, const Size_<T>& b) { return a.width == b.width && a.height == b.height; }
template <typename T> inline bool operator!=(const Size_<T>& a, const Size_<T>& b) { return !(a == b); }

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    static Range all() { return Range(INT_MIN, INT_MAX); }
};

struct Scalar {
    double val[4];
    Scalar(double v0 = 0, double v1 = 0, double v2 = 0, double v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    double operator[](int i) const { return val[i]; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
};

class KeyPoint {
public:
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f _pt, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(_pt), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct KeyPointsFilter {
    /* features2d/keypoint.cpp: keep the n strongest by response, plus every keypoint tying with the n-th */
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints) {
        if (npoints >= 0 && keypoints.size() > (size_t)npoints) {
            if (npoints == 0) { keypoints.clear(); return; }
            std::nth_element(keypoints.begin(), keypoints.begin() + npoints, keypoints.end(),
                             [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
            const float ambiguous = keypoints[npoints - 1].response;
            std::vector<KeyPoint>::const_iterator newEnd =
                std::partition(keypoints.begin() + npoints, keypoints.end(), [ambiguous](const KeyPoint& k) { return k.response >= ambiguous; });
            keypoints.resize(newEnd - keypoints.begin());
        }
    }
};

enum { GEMM_1_T = 1, GEMM_2_T = 2, GEMM_3_T = 4 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };
enum { DECOMP_LU = 0 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_REFLECT101 = 4,
       BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };

class Mat;
class MatExpr;
template <typename T> class Mat_;

/* Mat::step: converts to size_t and indexes like OpenCV's MatStep */
struct MatStep {
    size_t buf[2];
    size_t* p;
    MatStep() : p(buf) { buf[0] = buf[1] = 0; }
    MatStep(const MatStep& o) : p(buf) { buf[0] = o.buf[0]; buf[1] = o.buf[1]; }
    MatStep& operator=(const MatStep& o) { buf[0] = o.buf[0]; buf[1] = o.buf[1]; return *this; }
    MatStep& operator=(size_t s) { buf[0] = s; return *this; }
    operator size_t() const { return buf[0]; }
    size_t operator[](int i) const { return buf[i]; }
    size_t& operator[](int i) { return buf[i]; }
};

class Mat {
public:
    int flags, dims, rows, cols;
    uchar* data;
    const uchar* datastart;
    const uchar* dataend;
    MatStep step;

    Mat() { init(); }
    Mat(int r, int c, int type) { init(); create(r, c, type); }
    Mat(Size sz, int type) { init(); create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, const Scalar& s) { init(); create(r, c, type); setTo(s); }
    Mat(int r, int c, int type, void* ext, size_t step_ = 0) {
        init();
        flags = type; dims = 2; rows = r; cols = c; data = (uchar*)ext;
        step.buf[1] = elemSize();
        step.buf[0] = step_ ? step_ : (size_t)c * elemSize();
        datastart = data; dataend = data + (size_t)r * step.buf[0];
    }
    Mat(const Mat& m) = default;
    Mat& operator=(const Mat& m) = default;
    inline Mat(const MatExpr& e);
    inline Mat& operator=(const MatExpr& e);
    template <typename T> explicit Mat(const std::vector<T>& v);        /* N x 1, header over the vector */

    void create(int r, int c, int type) {
        type &= 0xfff;
        if (data && rows == r && cols == c && this->type() == type) return;
        release();
        flags = type; dims = 2; rows = r; cols = c;
        step.buf[1] = elemSize();
        step.buf[0] = (size_t)c * elemSize();
        const size_t bytes = (size_t)r * step.buf[0];
        buf_ = std::shared_ptr<uchar>(new uchar[bytes + 64], std::default_delete<uchar[]>());
        data = buf_.get();
        datastart = data; dataend = data + bytes;
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { buf_.reset(); data = nullptr; datastart = dataend = nullptr; rows = cols = 0; step.buf[0] = 0; }

    int type() const { return flags & 0xfff; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize1() const { static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return sz[depth()]; }
    size_t elemSize() const { return elemSize1() * channels(); }
    size_t step1(int i = 0) const { return step.buf[i] / elemSize1(); }
    size_t total() const { return (size_t)rows * cols; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return rows <= 1 || step.buf[0] == (size_t)cols * elemSize(); }
    bool isSubmatrix() const { return data != datastart || (size_t)(dataend - datastart) != (size_t)rows * step.buf[0] || !isContinuous(); }
    void locateROI(Size& whole, Point& ofs) const {
        const size_t esz = elemSize(), st = step.buf[0];
        const ptrdiff_t d1 = data - datastart, d2 = dataend - datastart;
        if (d1 == 0) ofs.x = ofs.y = 0;
        else { ofs.y = (int)(d1 / st); ofs.x = (int)((d1 - (size_t)ofs.y * st) / esz); }
        whole.height = std::max((int)((d2 - (ptrdiff_t)((size_t)cols * esz)) / (ptrdiff_t)st + 1), ofs.y + rows);
        whole.width = std::max((int)((d2 - (ptrdiff_t)(st * (size_t)(whole.height - 1))) / (ptrdiff_t)esz), ofs.x + cols);
    }

    Mat rowRange(int s, int e) const { Mat m(*this); m.rows = e - s; m.data = data + (size_t)s * step.buf[0]; return m; }
    Mat colRange(int s, int e) const { Mat m(*this); m.cols = e - s; m.data = data + (size_t)s * elemSize(); return m; }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat col(int c) const { return colRange(c, c + 1); }
    Mat operator()(const Rect& r) const { Mat m(*this); m.rows = r.height; m.cols = r.width; m.data = data + (size_t)r.y * step.buf[0] + (size_t)r.x * elemSize(); return m; }
    Mat operator()(Range rr, Range cr) const {
        Mat m(*this);
        if (rr.start != INT_MIN) m = m.rowRange(rr.start, rr.end);
        if (cr.start != INT_MIN) m = m.colRange(cr.start, cr.end);
        return m;
    }

    uchar* ptr(int r = 0) { return data + (size_t)r * step.buf[0]; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step.buf[0]; }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step.buf[0]); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step.buf[0]); }
    template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
    /* OpenCV Mat::at(int i0): element i0 of a single-row or single-column matrix */
    template <typename T> T& at(int i) { return (isContinuous() || rows == 1) ? reinterpret_cast<T*>(data)[i] : ptr<T>(i)[0]; }
    template <typename T> const T& at(int i) const { return (isContinuous() || rows == 1) ? reinterpret_cast<const T*>(data)[i] : ptr<T>(i)[0]; }

    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat& m) const {
        if (empty()) { m.release(); return; }
        m.create(rows, cols, type());
        if (m.data == data) return;
        for (int r = 0; r < rows; r++) memcpy(m.ptr(r), ptr(r), (size_t)cols * elemSize());
    }
    inline void copyTo(const class _OutputArray& o) const;
    Mat& setTo(const Scalar& s) {
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols * channels(); c++) setElem(r, c, s.val[c % channels()]);
        return *this;
    }
    Mat& operator=(const Scalar& s) { return setTo(s); }
    /* MatExpr initialisers: assigned to an existing matrix of the same size and type they fill it in place */
    static inline MatExpr zeros(int r, int c, int type);
    static inline MatExpr zeros(Size s, int type);
    static inline MatExpr ones(int r, int c, int type);
    static inline MatExpr eye(int r, int c, int type);

    /* channels <-> columns, OpenCV Mat::reshape(cn, rows = 0) */
    Mat reshape(int cn, int new_rows = 0) const {
        Mat m(*this);
        const int total_w = cols * channels();
        if (new_rows != 0 && new_rows != rows) {
            assert(isContinuous());
            const size_t total_elems = (size_t)rows * total_w;
            m.rows = new_rows;
            m.step.buf[0] = total_elems / new_rows * elemSize1();
            m.cols = (int)(total_elems / new_rows / cn);
        } else {
            m.cols = total_w / cn;
        }
        m.flags = CV_MAKETYPE(depth(), cn);
        m.step.buf[1] = m.elemSize();
        return m;
    }
    /* convertTo(dst, rtype, alpha, beta): dst = saturate_cast<D>(src*alpha + beta); float destinations compute in float
     * with (float)alpha,(float)beta (OpenCV cvtScale_<T,float,float>) */
    void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const {
        if (rtype < 0) rtype = type();
        rtype = CV_MAKETYPE(CV_MAT_DEPTH(rtype), channels());
        Mat src = *this;                      /* dst may be *this */
        Mat out(rows, cols, rtype);
        const int n = cols * channels();
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < n; c++) {
                if (CV_MAT_DEPTH(rtype) == CV_32F) {
                    const float v = (float)src.getElem(r, c);
                    float w = v;
                    if (alpha != 1 || beta != 0) w = v * (float)alpha + (float)beta;
                    reinterpret_cast<float*>(out.ptr(r))[c] = w;
                } else {
                    out.setElem(r, c, src.getElem(r, c) * alpha + beta);
                }
            }
        dst = out;
    }

    inline MatExpr t() const;
    inline MatExpr inv(int method = DECOMP_LU) const;
    inline MatExpr mul(const Mat& m, double scale = 1) const;
    /* Mat::dot: CV_32F accumulates double products sequentially (core matmul.cpp dotProd_ scalar path, length < 4) */
    double dot(const Mat& m) const {
        double r = 0;
        const int n = cols * channels();
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < n; j++) r += getElem(i, j) * m.getElem(i, j);
        return r;
    }

    double getElem(int r, int c) const {
        const uchar* p = ptr(r);
        switch (depth()) {
            case CV_8U: return p[c];
            case CV_8S: return ((const schar*)p)[c];
            case CV_16U: return ((const ushort*)p)[c];
            case CV_16S: return ((const short*)p)[c];
            case CV_32S: return ((const int*)p)[c];
            case CV_32F: return ((const float*)p)[c];
            default: return ((const double*)p)[c];
        }
    }
    void setElem(int r, int c, double v) {
        uchar* p = ptr(r);
        switch (depth()) {
            case CV_8U: p[c] = saturate_cast<uchar>(v); break;
            case CV_8S: ((schar*)p)[c] = (schar)cvRound(v); break;
            case CV_16U: ((ushort*)p)[c] = (ushort)cvRound(v); break;
            case CV_16S: ((short*)p)[c] = (short)cvRound(v); break;
            case CV_32S: ((int*)p)[c] = cvRound(v); break;
            case CV_32F: ((float*)p)[c] = (float)v; break;
            default: ((double*)p)[c] = v; break;
        }
    }

protected:
    void init() { flags = 0; dims = 2; rows = cols = 0; data = nullptr; datastart = dataend = nullptr; }
    std::shared_ptr<uchar> buf_;
};

template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8UC1 }; };
template <> struct DataType<int> { enum { type = CV_32SC1 }; };
template <> struct DataType<float> { enum { type = CV_32FC1 }; };
template <> struct DataType<double> { enum { type = CV_64FC1 }; };
template <> struct DataType<Point2f> { enum { type = CV_32FC2 }; };

template <typename T> Mat::Mat(const std::vector<T>& v) {
    init();
    if (v.empty()) return;
    flags = DataType<T>::type; dims = 2; rows = (int)v.size(); cols = 1; data = (uchar*)&v[0];
    step.buf[1] = elemSize(); step.buf[0] = elemSize();
    datastart = data; dataend = data + (size_t)rows * step.buf[0];
}

/* cv::Mat_<T>(r,c) << a, b, c  (comma initialiser) */
template <typename T> class MatCommaInitializer_ {
public:
    MatCommaInitializer_(Mat_<T>* m) : m_(m), i_(0) {}
    template <typename U> MatCommaInitializer_& operator,(U v);
    operator Mat_<T>() const { return *m_; }
    operator Mat() const { return *m_; }
    Mat_<T>* m_;
    int i_;
};
template <typename T> class Mat_ : public Mat {
public:
    Mat_() : Mat() {}
    Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
    Mat_(const Mat& m) : Mat(m) {}
    T& operator()(int r, int c) { return this->template at<T>(r, c); }
    const T& operator()(int r, int c) const { return this->template at<T>(r, c); }
};
template <typename T> template <typename U> MatCommaInitializer_<T>& MatCommaInitializer_<T>::operator,(U v) {
    m_->template at<T>(i_ / m_->cols, i_ % m_->cols) = (T)v;
    i_++;
    return *this;
}
/* holds its own header copy so that a temporary Mat_ (the usual `(Mat_<float>(3,1) << x, y, z)`) stays alive */
template <typename T> class MatCommaHolder_ : public MatCommaInitializer_<T> {
public:
    MatCommaHolder_(const Mat_<T>& m) : MatCommaInitializer_<T>(&own_), own_(m) {}
    MatCommaHolder_(const MatCommaHolder_& o) : MatCommaInitializer_<T>(&own_), own_(o.own_) { this->i_ = o.i_; }
    Mat_<T> own_;
};
template <typename T, typename U> inline MatCommaHolder_<T> operator<<(const Mat_<T>& m, U v) {
    MatCommaHolder_<T> h(m);
    h, v;
    return h;
}

/* ---- InputArray / OutputArray (core/mat.hpp): proxies over Mat or std::vector ---- */
class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    inline _InputArray(const MatExpr& e);
    template <typename T> _InputArray(const std::vector<T>& v) : own_(v), m_(&own_) {}
    Mat getMat(int = -1) const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
    Size size() const { return m_ ? m_->size() : Size(); }
    int type() const { return m_ ? m_->type() : 0; }
    int depth() const { return CV_MAT_DEPTH(type()); }
    int channels() const { return CV_MAT_CN(type()); }
    size_t total() const { return m_ ? m_->total() : 0; }

protected:
    Mat own_;
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) { m_ = &m; }
    _OutputArray(const Mat& m) { own_ = m; m_ = &own_; }      /* temporaries (submatrix headers): fixed size and type */
    template <typename T> _OutputArray(std::vector<T>& v) { vecResize_ = [&v](size_t n) { v.resize(n); return (void*)(n ? &v[0] : nullptr); }; vecType_ = DataType<T>::type; }
    void create(int r, int c, int type) const {
        if (m_) { m_->create(r, c, type); return; }
        if (vecResize_) {
            void* p = vecResize_((size_t)r * c);
            const_cast<_OutputArray*>(this)->own_ = Mat(r * c, 1, vecType_, p);
            const_cast<_OutputArray*>(this)->m_ = &const_cast<_OutputArray*>(this)->own_;
        }
    }
    void create(Size sz, int type) const { create(sz.height, sz.width, type); }
    void release() const { if (m_) m_->release(); }
    Mat& getMatRef() const { return *m_; }
    bool needed() const { return m_ != nullptr || (bool)vecResize_; }

protected:
    std::function<void*(size_t)> vecResize_;
    int vecType_ = 0;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
inline InputArray noArray() { static _InputArray none; return none; }
inline void Mat::copyTo(const _OutputArray& o) const { o.create(rows, cols, type()); Mat d = o.getMat(); copyTo(d); }

/* ---- the arithmetic primitives (minicv.cpp) ---- */
void gemm(InputArray A, InputArray B, double alpha, InputArray C, double beta, OutputArray D, int flags = 0);
void transpose(InputArray src, OutputArray dst);
double invert(InputArray src, OutputArray dst, int method = DECOMP_LU);
double norm(InputArray a, int normType = NORM_L2);
double norm(InputArray a, InputArray b, int normType = NORM_L2);
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType, const Scalar& value = Scalar());
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
float fastAtan2(float y, float x);
void undistortPoints(InputArray src, OutputArray dst, InputArray cameraMatrix, InputArray distCoeffs, InputArray R = noArray(), InputArray P = noArray());

/* ---- MatExpr: the lazy forms the reference uses; evaluation follows OpenCV's matop.cpp ---- */
class MatExpr {
public:
    enum Kind { IDENT, T, GEMM, ADDEX, INV, MULEL };
    Kind kind;
    Mat a, b, c;
    double alpha, beta;
    int flags;
    Scalar s;
    MatExpr() : kind(IDENT), alpha(1), beta(0), flags(0) {}
    explicit MatExpr(const Mat& m) : kind(IDENT), a(m), alpha(1), beta(0), flags(0) {}
    operator Mat() const { Mat m; eval(m); return m; }
    void eval(Mat& m) const {
        switch (kind) {
            case IDENT: m = a; break;
            case T: {                       /* MatOp_T::assign: transpose, then convertTo(alpha) unless alpha == 1 */
                Mat tmp;
                cv::transpose(a, tmp);
                if (alpha == 1) m = tmp; else tmp.convertTo(m, -1, alpha);
                break;
            }
            case GEMM: {
                Mat d;
                cv::gemm(a, b, alpha, c, beta, d, flags);
                m = d;
                break;
            }
            case ADDEX: {                   /* MatOp_AddEx::assign */
                Mat d;
                if (b.empty()) { a.convertTo(d, -1, alpha, s.val[0]); m = d; break; }
                d.create(a.rows, a.cols, a.type());
                const int n = a.cols * a.channels();
                for (int r = 0; r < a.rows; r++)
                    for (int j = 0; j < n; j++) {
                        if (a.depth() == CV_32F) {
                            const float x = ((const float*)a.ptr(r))[j], y = ((const float*)b.ptr(r))[j];
                            float v;
                            if (alpha == 1 && beta == 1) v = x + y;                   /* cv::add */
                            else if (alpha == 1 && beta == -1) v = x - y;             /* cv::subtract */
                            else v = x * (float)alpha + y * (float)beta;              /* cv::addWeighted, float work type */
                            ((float*)d.ptr(r))[j] = v;
                        } else {
                            d.setElem(r, j, a.getElem(r, j) * alpha + b.getElem(r, j) * beta);
                        }
                    }
                m = d;
                break;
            }
            case INV: { Mat d; cv::invert(a, d, flags); m = d; break; }
            case MULEL: {
                Mat d(a.rows, a.cols, a.type());
                const int n = a.cols * a.channels();
                for (int r = 0; r < a.rows; r++)
                    for (int j = 0; j < n; j++) {
                        if (a.depth() == CV_32F) ((float*)d.ptr(r))[j] = alpha == 1 ? ((const float*)a.ptr(r))[j] * ((const float*)b.ptr(r))[j]
                                                                                    : (float)alpha * ((const float*)a.ptr(r))[j] * ((const float*)b.ptr(r))[j];
                        else d.setElem(r, j, alpha * a.getElem(r, j) * b.getElem(r, j));
                    }
                m = d;
                break;
            }
        }
    }
    /* the few Mat members the reference applies to expressions directly */
    template <typename U> U at(int i) const { return Mat(*this).at<U>(i); }
    template <typename U> U at(int r, int c) const { return Mat(*this).at<U>(r, c); }
    Mat row(int r) const { return Mat(*this).row(r); }
    Mat col(int c) const { return Mat(*this).col(c); }
    MatExpr t() const { return Mat(*this).t(); }
    double dot(const Mat& m) const { return Mat(*this).dot(m); }
    Mat clone() const { return Mat(*this).clone(); }
};
inline Mat::Mat(const MatExpr& e) { init(); e.eval(*this); }
/* OpenCV's MatOp::assign ends in m.create(size, type), a no-op for a matrix that already has them: the result lands in
 * the existing buffer (ORBextractor.cc:1037 relies on it: `descriptors = Mat::zeros(...)` clears the caller's rows) */
inline Mat& Mat::operator=(const MatExpr& e) {
    Mat m;
    e.eval(m);
    if (data && rows == m.rows && cols == m.cols && type() == m.type()) m.copyTo(*this);
    else *this = m;
    return *this;
}
inline MatExpr Mat::zeros(int r, int c, int type) { return MatExpr(Mat(r, c, type, Scalar::all(0))); }
inline MatExpr Mat::zeros(Size s, int type) { return zeros(s.height, s.width, type); }
inline MatExpr Mat::ones(int r, int c, int type) { return MatExpr(Mat(r, c, type, Scalar(1))); }
inline MatExpr Mat::eye(int r, int c, int type) {
    Mat m(r, c, type, Scalar::all(0));
    for (int i = 0; i < std::min(r, c); i++) m.setElem(i, i, 1);
    return MatExpr(m);
}
inline _InputArray::_InputArray(const MatExpr& e) : own_(e), m_(&own_) {}
inline MatExpr Mat::t() const { MatExpr e; e.kind = MatExpr::T; e.a = *this; return e; }
inline MatExpr Mat::inv(int method) const { MatExpr e; e.kind = MatExpr::INV; e.a = *this; e.flags = method; return e; }
inline MatExpr Mat::mul(const Mat& m, double scale) const { MatExpr e; e.kind = MatExpr::MULEL; e.a = *this; e.b = m; e.alpha = scale; return e; }

namespace detail {
inline MatExpr addex(const Mat& a, const Mat& b, double alpha, double beta, double s = 0) {
    MatExpr e; e.kind = MatExpr::ADDEX; e.a = a; e.b = b; e.alpha = alpha; e.beta = beta; e.s = Scalar(s); return e;
}
/* a scaled or transposed operand of a product, as MatOp_GEMM::makeExpr folds it */
inline void fold(const MatExpr& x, Mat& m, double& alpha, int& tflag) {
    if (x.kind == MatExpr::IDENT) { m = x.a; }
    else if (x.kind == MatExpr::T) { m = x.a; alpha *= x.alpha; tflag = 1; }
    else if (x.kind == MatExpr::ADDEX && x.b.empty() && x.s.val[0] == 0) { m = x.a; alpha *= x.alpha; }
    else { m = Mat(x); }
}
inline MatExpr product(const MatExpr& x, const MatExpr& y) {
    MatExpr e; e.kind = MatExpr::GEMM; e.alpha = 1; e.beta = 0; e.flags = 0;
    int ta = 0, tb = 0;
    fold(x, e.a, e.alpha, ta);
    fold(y, e.b, e.alpha, tb);
    e.flags = (ta ? GEMM_1_T : 0) | (tb ? GEMM_2_T : 0);
    return e;
}
inline MatExpr plus(const MatExpr& x, const MatExpr& y, double sign) {
    /* MatOp_GEMM::add / subtract: a product without a C operand absorbs the other term */
    if (x.kind == MatExpr::GEMM && x.c.empty() && y.kind == MatExpr::IDENT) { MatExpr e = x; e.c = y.a; e.beta = sign; return e; }
    if (y.kind == MatExpr::GEMM && y.c.empty() && x.kind == MatExpr::IDENT && sign == 1) { MatExpr e = y; e.c = x.a; e.beta = 1; return e; }
    Mat a, b;
    double al = 1, be = sign;
    if (x.kind == MatExpr::ADDEX && x.b.empty() && x.s.val[0] == 0) { a = x.a; al = x.alpha; } else a = Mat(x);
    if (y.kind == MatExpr::ADDEX && y.b.empty() && y.s.val[0] == 0) { b = y.a; be = sign * y.alpha; } else b = Mat(y);
    return addex(a, b, al, be);
}
}  // namespace detail

inline MatExpr operator*(const Mat& a, const Mat& b) { return detail::product(MatExpr(a), MatExpr(b)); }
inline MatExpr operator*(const MatExpr& a, const Mat& b) { return detail::product(a, MatExpr(b)); }
inline MatExpr operator*(const Mat& a, const MatExpr& b) { return detail::product(MatExpr(a), b); }
inline MatExpr operator*(const MatExpr& a, const MatExpr& b) { return detail::product(a, b); }
inline MatExpr scaled(const MatExpr& x, double s) {
    if (x.kind == MatExpr::T || x.kind == MatExpr::GEMM || x.kind == MatExpr::MULEL) { MatExpr e = x; e.alpha *= s; if (x.kind == MatExpr::GEMM) e.beta *= s; return e; }
    if (x.kind == MatExpr::ADDEX && x.b.empty()) { MatExpr e = x; e.alpha *= s; e.s = Scalar(x.s.val[0] * s); return e; }
    return detail::addex(Mat(x), Mat(), s, 0);
}
inline MatExpr operator*(const Mat& a, double s) { return detail::addex(a, Mat(), s, 0); }
inline MatExpr operator*(double s, const Mat& a) { return detail::addex(a, Mat(), s, 0); }
inline MatExpr operator*(const MatExpr& a, double s) { return scaled(a, s); }
inline MatExpr operator*(double s, const MatExpr& a) { return scaled(a, s); }
inline MatExpr operator/(const Mat& a, double s) { return detail::addex(a, Mat(), 1. / s, 0); }          /* matop.cpp: 1./s */
inline MatExpr operator/(const MatExpr& a, double s) { return scaled(a, 1. / s); }
inline MatExpr operator-(const Mat& a) { return detail::addex(a, Mat(), -1, 0); }
inline MatExpr operator-(const MatExpr& a) { return scaled(a, -1); }
inline MatExpr operator+(const Mat& a, const Mat& b) { return detail::addex(a, b, 1, 1); }
inline MatExpr operator-(const Mat& a, const Mat& b) { return detail::addex(a, b, 1, -1); }
inline MatExpr operator+(const MatExpr& a, const Mat& b) { return detail::plus(a, MatExpr(b), 1); }
inline MatExpr operator+(const Mat& a, const MatExpr& b) { return detail::plus(MatExpr(a), b, 1); }
inline MatExpr operator+(const MatExpr& a, const MatExpr& b) { return detail::plus(a, b, 1); }
inline MatExpr operator-(const MatExpr& a, const Mat& b) { return detail::plus(a, MatExpr(b), -1); }
inline MatExpr operator-(const Mat& a, const MatExpr& b) { return detail::plus(MatExpr(a), b, -1); }
inline MatExpr operator-(const MatExpr& a, const MatExpr& b) { return detail::plus(a, b, -1); }
inline MatExpr operator+(const Mat& a, const Scalar& s) { return detail::addex(a, Mat(), 1, 0, s.val[0]); }
inline MatExpr operator-(const Mat& a, const Scalar& s) { return detail::addex(a, Mat(), 1, 0, -s.val[0]); }
inline MatExpr abs(const Mat& a) {
    Mat d(a.rows, a.cols, a.type());
    for (int r = 0; r < a.rows; r++)
        for (int j = 0; j < a.cols * a.channels(); j++) d.setElem(r, j, std::fabs(a.getElem(r, j)));
    return MatExpr(d);
}
inline double determinant(InputArray m_) {
    Mat m = m_.getMat();
    if (m.rows == 2) return m.getElem(0, 0) * m.getElem(1, 1) - m.getElem(0, 1) * m.getElem(1, 0);
    double d = 0;
    for (int i = 0; i < 3; i++)
        d += m.getElem(0, i) * (m.getElem(1, (i + 1) % 3) * m.getElem(2, (i + 2) % 3) - m.getElem(1, (i + 2) % 3) * m.getElem(2, (i + 1) % 3));
    return d;
}
inline std::ostream& operator<<(std::ostream& os, const Mat& m) {
    os << "[";
    for (int r = 0; r < m.rows; r++) {
        for (int c = 0; c < m.cols * m.channels(); c++) os << (c ? ", " : "") << m.getElem(r, c);
        os << (r + 1 < m.rows ? ";\n" : "");
    }
    return os << "]";
}

/* cv::FileStorage: the DBoW2 save/load members mention it; never executed on the hot path */
class FileNode {
public:
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0; }
    operator std::string() const { return std::string(); }
    enum { SEQ = 5 };
    int type() const { return 0; }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
};
template <typename T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

}  // namespace cv
#endif
