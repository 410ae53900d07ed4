// cv_min.h -- the handful of OpenCV value types that cross the reference's stage boundary (SfMCommon.h:55-99), for
// building and testing the shim in an image WITHOUT OpenCV.  With real OpenCV available, compile the shim with
// -DSFMB200_WITH_OPENCV and this file is not used.  Only members the shim touches are provided.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <vector>

namespace cv {

enum { CV_8U = 0, CV_32F = 5, CV_8UC3 = 16 };
struct Vec3b { unsigned char val[3]; unsigned char operator()(int i) const { return val[i]; } };

template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f;
template <typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<float> Point3f;

// same members, order and size (28 bytes) as cv::KeyPoint
struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

struct DMatch {
    int queryIdx, trainIdx, imgIdx; float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(0) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
};

template <typename T, int M, int N> struct Matx {
    T val[M * N];
    Matx() { std::memset(val, 0, sizeof val); }
    T& operator()(int r, int c) { return val[r * N + c]; }
    const T& operator()(int r, int c) const { return val[r * N + c]; }
    static Matx eye() { Matx m; for (int i = 0; i < (M < N ? M : N); ++i) m(i, i) = 1; return m; }
};
typedef Matx<float, 3, 4> Matx34f;
typedef Matx<float, 3, 3> Matx33f;

// dense row-major matrix of uint8 or float32, reference-counted like cv::Mat
class Mat {
public:
    int rows = 0, cols = 0;
    uint8_t* data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type), buf_(new std::vector<uint8_t>((size_t)r * c * elem(type))) { data = buf_->data(); }
    int type() const { return type_; }
    bool empty() const { return rows == 0 || cols == 0; }
    bool isContinuous() const { return true; }
    Mat clone() const { Mat m(rows, cols, type_); if (data) std::memcpy(m.data, data, (size_t)rows * cols * elem(type_)); return m; }
    size_t elemSize() const { return elem(type_); }
    int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
    template <typename T> T& at(int r, int c) { return reinterpret_cast<T*>(data)[(size_t)r * cols + c]; }
    template <typename T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(data)[(size_t)r * cols + c]; }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data) + (size_t)r * cols; }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data) + (size_t)r * cols; }
private:
    static size_t elem(int t) { return t == CV_32F ? 4 : (t == CV_8UC3 ? 3 : 1); }
    int type_ = CV_8U;
    std::shared_ptr<std::vector<uint8_t>> buf_;
};

}  // namespace cv
