/*
 * cv_compat.h -- the few OpenCV types the ORB front-end's call surface mentions (cv::Mat, cv::KeyPoint,
 * cv::Point2f, InputArray/OutputArray), for builds without OpenCV headers (this image has none).
 * Define VIORB_USE_OPENCV to compile the shims against the real <opencv2/core/core.hpp> instead; the
 * layouts used here (KeyPoint = 28-byte POD, Mat = rows/cols/step/data) are the ones OpenCV uses.
 */
#ifndef VIORB_CV_COMPAT_H
#define VIORB_CV_COMPAT_H

#ifdef VIORB_USE_OPENCV
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#else

#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

namespace cv {

struct Point2f {
    float x, y;
    Point2f() : x(0), y(0) {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};

struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

/* 2-D matrix of CV_8U or CV_32F with shared ownership and row views -- what the shims need, no more */
class Mat {
public:
    int rows, cols;
    size_t step;
    unsigned char* data;

    Mat() : rows(0), cols(0), step(0), data(nullptr), type_(CV_8U) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t step_ = 0) : rows(r), cols(c), data((unsigned char*)ext), type_(type) {
        step = step_ ? step_ : (size_t)c * elemSize();
    }
    void create(int r, int c, int type) {
        type_ = type; rows = r; cols = c;
        step = (size_t)c * elemSize();
        buf_ = std::make_shared<std::vector<unsigned char> >((size_t)r * step);
        data = buf_->data();
    }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); if (m.data) memset(m.data, 0, (size_t)r * m.step); return m; }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    size_t elemSize() const { return type_ == CV_32F ? 4 : 1; }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++) memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elemSize());
        return m;
    }
    Mat row(int r) const { Mat m(*this); m.rows = 1; m.data = data + (size_t)r * step; return m; }
    Mat operator()(const Rect& r) const { Mat m(*this); m.rows = r.height; m.cols = r.width; m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize(); return m; }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
    template <typename T> T& at(int i) { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }
    template <typename T> const T& at(int i) const { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }

private:
    int type_;
    std::shared_ptr<std::vector<unsigned char> > buf_;
};

/* the proxies of core/mat.hpp: same class names (hence the same mangled operator() as the reference's library) and the
 * members the shims call -- getMat / empty / type on inputs, create / release / getMat on outputs */
class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    Mat getMat(int = -1) const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
    int type() const { return m_ ? m_->type() : 0; }

protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) { m_ = &m; }
    void create(int r, int c, int type) const {
        if (m_ && !(m_->data && m_->rows == r && m_->cols == c && m_->type() == type)) m_->create(r, c, type);
    }
    void release() const { if (m_) m_->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

}  // namespace cv
#endif  /* VIORB_USE_OPENCV */
#endif
