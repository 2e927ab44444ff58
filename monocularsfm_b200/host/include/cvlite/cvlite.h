// cvlite — the handful of OpenCV value types the reference's hot-path API mentions (cv::Mat, cv::DMatch,
// cv::KeyPoint, cv::Vec*), so that the drop-in classes compile in an image without OpenCV C++ headers.
// Build with -DMSFM_WITH_OPENCV to use the real <opencv2/opencv.hpp> instead (then this file is a no-op and the
// classes link against the reference's own OpenCV).  Only what FeatureUtils / FeatureMatcher / BundleData /
// Database need is provided; semantics follow OpenCV's documentation (row-major, reference-counted buffer).
#ifndef MSFM_CVLITE_H_
#define MSFM_CVLITE_H_
#ifdef MSFM_WITH_OPENCV
#include <opencv2/opencv.hpp>
#else
#include <cassert>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6

namespace cv {
typedef unsigned char uchar;

template <class T>
using Ptr = std::shared_ptr<T>;

template <class T, int N>
struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
    Vec(T a, T b) { static_assert(N == 2, ""); val[0] = a; val[1] = b; }
    Vec(T a, T b, T c) { static_assert(N == 3, ""); val[0] = a; val[1] = b; val[2] = c; }
    T& operator()(int i) { return val[i]; }
    const T& operator()(int i) const { return val[i]; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<double, 2> Vec2d;
typedef Vec<double, 3> Vec3d;
typedef Vec<uchar, 3> Vec3b;

struct Point2f {
    float x = 0, y = 0;
    Point2f() {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};

struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};

struct DMatch {
    int queryIdx = -1, trainIdx = -1, imgIdx = -1;
    float distance = 3.402823466e+38f;
    DMatch() {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
};

class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        const size_t bytes = static_cast<size_t>(r) * c * elemSize();
        buf_ = std::shared_ptr<uchar>(new uchar[bytes ? bytes : 1], std::default_delete<uchar[]>());
        data = buf_.get();
        std::memset(data, 0, bytes);
    }
    int type() const { return type_; }
    bool empty() const { return rows == 0 || cols == 0 || !data; }
    bool isContinuous() const { return true; }
    size_t elemSize() const { return type_ == CV_8U ? 1 : (type_ == CV_64F ? 8 : 4); }
    size_t step() const { return static_cast<size_t>(cols) * elemSize(); }
    template <class T> T& at(int i, int j = 0) { return reinterpret_cast<T*>(data + i * step())[j]; }
    template <class T> const T& at(int i, int j = 0) const { return reinterpret_cast<const T*>(data + i * step())[j]; }
    template <class T> T* ptr(int i = 0) { return reinterpret_cast<T*>(data + i * step()); }
    template <class T> const T* ptr(int i = 0) const { return reinterpret_cast<const T*>(data + i * step()); }
    Mat clone() const {
        Mat m(rows, cols, type_);
        if (!empty()) std::memcpy(m.data, data, static_cast<size_t>(rows) * step());
        return m;
    }
private:
    int type_ = CV_8U;
    std::shared_ptr<uchar> buf_;
};
}  // namespace cv
#endif  // MSFM_WITH_OPENCV
#endif  // MSFM_CVLITE_H_
