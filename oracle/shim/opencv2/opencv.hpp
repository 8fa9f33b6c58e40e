// opencv2/opencv.hpp -- MINIMAL OpenCV-2.4 API SHIM (test infrastructure, part of oracle/).
//
// Purpose: compile the reference's own pipeline sources (/root/reference/core/cnn_softam.h, cnn.h, maxloss.h,
// Hypothesis.cpp, types.h, properties.cpp, read_data.cpp, the four RANSAC drivers ...) UNMODIFIED into oracle/_ref,
// so that the oracle's restatement of the pipeline-level control flow (sampling loop, refinement stop rules, quirks,
// gradient assembly, conventions) can be pinned against the reference's own code.  OpenCV itself is not available in
// this image; this header provides exactly the slice of the cv:: API those sources use, written from scratch:
//
//   * containers and their semantics (cv::Mat header/data sharing, row/col/range views, write-through assignment of
//     matrix expressions into views, Mat_<T>, Vec, Point_, Point3_, Scalar, Size) are implemented here;
//   * the calib3d / core NUMERICS (solvePnP CV_P3P / CV_ITERATIVE, projectPoints, Rodrigues + Jacobian, SVD) are
//     forwarded to the oracle's restated routines (oracle/dsac_oracle.h: orc_solve_p3p, orc_solve_pnp_iterative,
//     orc_project_points, orc_rodrigues, orc_rodrigues_inv, orc_svd3), which are pinned to cv2 4.13 golden vectors
//     in tests/test_oracle_golden.py.  What oracle/_ref therefore pins is everything ABOVE that boundary.
//
// Not a general OpenCV replacement: single-channel float/double arithmetic only, no ROI bookkeeping beyond views,
// no bounds checks (like OpenCV release builds, which the reference relies on: Hypothesis.cpp:281-282 reads
// rv.at<double>(0,1) of a 3x1 matrix).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../dsac_oracle.h"

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 63) + 1)
#define CV_PI 3.1415926535897932384626433832795

enum { CV_ITERATIVE = 0, CV_EPNP = 1, CV_P3P = 2 };

// OpenCV's C headers put these in the global namespace (core/types_c.h); the reference relies on it (types.h:47)
typedef unsigned char uchar;
typedef unsigned short ushort;

namespace cv {

using ::uchar;
using ::ushort;

// ---------------------------------------------------------------- saturate_cast (cv::saturate_cast semantics: round to nearest even, clamp)
template <typename T> inline T saturate_cast(double v) { return (T)v; }
template <typename T> inline T saturate_cast(float v) { return saturate_cast<T>((double)v); }
template <typename T> inline T saturate_cast(int v) { return (T)v; }
inline int shim_round(double v) { return (int)std::nearbyint(v); }   // cvRound: round-half-to-even under the default FP mode
template <> inline uchar saturate_cast<uchar>(int v) { return (uchar)((unsigned)v <= 255 ? v : v > 0 ? 255 : 0); }
template <> inline uchar saturate_cast<uchar>(double v) { return saturate_cast<uchar>(shim_round(v)); }
template <> inline short saturate_cast<short>(int v) { return (short)((unsigned)(v + 32768) <= 65535 ? v : v > 0 ? 32767 : -32768); }
template <> inline short saturate_cast<short>(double v) {
    // cvRound of an out-of-int-range double is INT_MIN on x86 (cvtsd2si); the clamp then gives SHRT_MIN
    if (!(v > -2147483648.0 && v < 2147483648.0)) return saturate_cast<short>((int)0x80000000);
    return saturate_cast<short>(shim_round(v));
}
template <> inline ushort saturate_cast<ushort>(int v) { return (ushort)((unsigned)v <= 65535 ? v : v > 0 ? 65535 : 0); }
template <> inline ushort saturate_cast<ushort>(double v) { return saturate_cast<ushort>(shim_round(v)); }
template <> inline int saturate_cast<int>(double v) { return shim_round(v); }
template <> inline float saturate_cast<float>(double v) { return (float)v; }
template <> inline double saturate_cast<double>(double v) { return v; }
template <typename T, typename S> inline T shim_cast(S v) { return saturate_cast<T>((double)v); }

// ---------------------------------------------------------------- small value types
template <typename T, int n> class Vec {
public:
    T val[n];
    Vec() { for (int i = 0; i < n; i++) val[i] = T(); }
    Vec(T a, T b) { static_assert(n >= 2, ""); for (int i = 0; i < n; i++) val[i] = T(); val[0] = a; val[1] = b; }
    Vec(T a, T b, T c) { static_assert(n >= 3, ""); for (int i = 0; i < n; i++) val[i] = T(); val[0] = a; val[1] = b; val[2] = c; }
    Vec(T a, T b, T c, T d) { static_assert(n >= 4, ""); for (int i = 0; i < n; i++) val[i] = T(); val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    T& operator()(int i) { return val[i]; }
    const T& operator()(int i) const { return val[i]; }
    template <typename T2> operator Vec<T2, n>() const {
        Vec<T2, n> r;
        for (int i = 0; i < n; i++) r.val[i] = shim_cast<T2>(val[i]);
        return r;
    }
};
template <typename T, int n> inline Vec<T, n> operator*(const Vec<T, n>& a, double s) { Vec<T, n> r; for (int i = 0; i < n; i++) r.val[i] = shim_cast<T>(a.val[i] * s); return r; }
template <typename T, int n> inline Vec<T, n> operator*(const Vec<T, n>& a, int s) { Vec<T, n> r; for (int i = 0; i < n; i++) r.val[i] = shim_cast<T>(a.val[i] * s); return r; }
template <typename T, int n> inline Vec<T, n> operator*(const Vec<T, n>& a, float s) { Vec<T, n> r; for (int i = 0; i < n; i++) r.val[i] = shim_cast<T>(a.val[i] * s); return r; }
template <typename T, int n> inline Vec<T, n> operator+(const Vec<T, n>& a, const Vec<T, n>& b) { Vec<T, n> r; for (int i = 0; i < n; i++) r.val[i] = shim_cast<T>(a.val[i] + b.val[i]); return r; }
template <typename T, int n> inline Vec<T, n> operator-(const Vec<T, n>& a, const Vec<T, n>& b) { Vec<T, n> r; for (int i = 0; i < n; i++) r.val[i] = shim_cast<T>(a.val[i] - b.val[i]); return r; }
typedef Vec<uchar, 3> Vec3b;
typedef Vec<short, 3> Vec3s;
typedef Vec<float, 3> Vec3f;
typedef Vec<double, 3> Vec3d;

template <typename T> class Scalar_ : public Vec<T, 4> {
public:
    Scalar_() {}
    Scalar_(T v0) { this->val[0] = v0; }
    Scalar_(T v0, T v1, T v2 = 0, T v3 = 0) { this->val[0] = v0; this->val[1] = v1; this->val[2] = v2; this->val[3] = v3; }
};
typedef Scalar_<double> Scalar;

template <typename T> class Size_ {
public:
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
};
typedef Size_<int> Size;

template <typename T> class Point_ {
public:
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename T2> operator Point_<T2>() const { return Point_<T2>(shim_cast<T2>(x), shim_cast<T2>(y)); }
};
template <typename T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(shim_cast<T>(a.x - b.x), shim_cast<T>(a.y - b.y)); }
template <typename T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(shim_cast<T>(a.x + b.x), shim_cast<T>(a.y + b.y)); }
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> class Point3_ {
public:
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
    Point3_(const Vec<T, 3>& v) : x(v[0]), y(v[1]), z(v[2]) {}
    template <typename T2> operator Point3_<T2>() const { return Point3_<T2>(shim_cast<T2>(x), shim_cast<T2>(y), shim_cast<T2>(z)); }
    Point3_& operator+=(const Point3_& o) { x = shim_cast<T>(x + o.x); y = shim_cast<T>(y + o.y); z = shim_cast<T>(z + o.z); return *this; }
    Point3_& operator-=(const Point3_& o) { x = shim_cast<T>(x - o.x); y = shim_cast<T>(y - o.y); z = shim_cast<T>(z - o.z); return *this; }
    Point3_& operator*=(double s) { x = shim_cast<T>(x * s); y = shim_cast<T>(y * s); z = shim_cast<T>(z * s); return *this; }
};
template <typename T> inline Point3_<T> operator-(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(shim_cast<T>(a.x - b.x), shim_cast<T>(a.y - b.y), shim_cast<T>(a.z - b.z)); }
template <typename T> inline Point3_<T> operator+(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(shim_cast<T>(a.x + b.x), shim_cast<T>(a.y + b.y), shim_cast<T>(a.z + b.z)); }
template <typename T> inline Point3_<T> operator-(const Point3_<T>& a) { return Point3_<T>(shim_cast<T>(-a.x), shim_cast<T>(-a.y), shim_cast<T>(-a.z)); }
typedef Point3_<int> Point3i;
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

template <typename T> inline double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }
template <typename T> inline double norm(const Point3_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y + (double)p.z * p.z); }

// ---------------------------------------------------------------- element type traits
template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { depth = CV_8U, channels = 1, type = CV_MAKETYPE(CV_8U, 1) }; };
template <> struct DataType<signed char> { enum { depth = CV_8S, channels = 1, type = CV_MAKETYPE(CV_8S, 1) }; };
template <> struct DataType<ushort> { enum { depth = CV_16U, channels = 1, type = CV_MAKETYPE(CV_16U, 1) }; };
template <> struct DataType<short> { enum { depth = CV_16S, channels = 1, type = CV_MAKETYPE(CV_16S, 1) }; };
template <> struct DataType<int> { enum { depth = CV_32S, channels = 1, type = CV_MAKETYPE(CV_32S, 1) }; };
template <> struct DataType<float> { enum { depth = CV_32F, channels = 1, type = CV_MAKETYPE(CV_32F, 1) }; };
template <> struct DataType<double> { enum { depth = CV_64F, channels = 1, type = CV_MAKETYPE(CV_64F, 1) }; };
template <typename T, int n> struct DataType<Vec<T, n>> { enum { depth = DataType<T>::depth, channels = n, type = CV_MAKETYPE(DataType<T>::depth, n) }; };
template <typename T> struct DataType<Point_<T>> { enum { depth = DataType<T>::depth, channels = 2, type = CV_MAKETYPE(DataType<T>::depth, 2) }; };
template <typename T> struct DataType<Point3_<T>> { enum { depth = DataType<T>::depth, channels = 3, type = CV_MAKETYPE(DataType<T>::depth, 3) }; };

inline size_t shim_depth_size(int depth) {
    static const size_t sz[] = {1, 1, 2, 2, 4, 4, 8, 0};
    return sz[depth & 7];
}

class MatExpr;
template <typename T> class Mat_;

// ---------------------------------------------------------------- cv::Mat: a header (rows, cols, type, step, data pointer) over shared storage
class Mat {
public:
    int rows, cols;
    uchar* data;
    size_t step;   // bytes per row

    Mat() : rows(0), cols(0), data(nullptr), step(0), type_(CV_8U) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
    Mat(int r, int c, int type, const Scalar& s) : Mat() { create(r, c, type); setTo(s[0]); }
    template <typename T> explicit Mat(const Point3_<T>& p) : Mat() {
        create(3, 1, DataType<T>::type);
        at<T>(0, 0) = p.x; at<T>(1, 0) = p.y; at<T>(2, 0) = p.z;
    }
    template <typename T> explicit Mat(const Point_<T>& p) : Mat() {
        create(2, 1, DataType<T>::type);
        at<T>(0, 0) = p.x; at<T>(1, 0) = p.y;
    }
    template <typename T, int n> explicit Mat(const Vec<T, n>& v) : Mat() {
        create(n, 1, DataType<T>::type);
        for (int i = 0; i < n; i++) at<T>(i, 0) = v[i];
    }
    Mat(const MatExpr& e);
    // header copy: shares the data (OpenCV semantics)
    Mat(const Mat&) = default;
    Mat& operator=(const Mat&) = default;
    // a matrix expression assigned to a matrix of the same size and type is evaluated INTO its memory (so assigning
    // to a row/col/range view modifies the parent: types.h:142-143, cnn_softam.h:497-500); otherwise re-allocated
    Mat& operator=(const MatExpr& e);
    Mat& operator=(const Scalar& s) { setTo(s[0]); return *this; }

    void create(int r, int c, int type) {
        if (data && rows == r && cols == c && type_ == type) return;
        // cv::Mat::create raises cv::Exception (a std::exception) for impossible sizes; the reference relies on that
        // when it probes for ./sensorTrans.dat (properties.cpp:79-90 reads rows / cols from a stream that failed to open)
        if (r < 0 || c < 0 || (double)r * (double)c * (double)(shim_depth_size(CV_MAT_DEPTH(type)) * CV_MAT_CN(type)) > 1.0e9)
            throw std::runtime_error("shim: cv::Mat::create with an impossible size");
        rows = r; cols = c; type_ = type;
        const size_t es = elemSize();
        step = (size_t)c * es;
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * step + 16, (uchar)0);
        data = buf_->data();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    int type() const { return type_; }
    int depth() const { return CV_MAT_DEPTH(type_); }
    int channels() const { return CV_MAT_CN(type_); }
    size_t elemSize() const { return shim_depth_size(depth()) * (size_t)channels(); }
    size_t total() const { return (size_t)rows * cols; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }

    template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }

    // element access as double, single-channel numeric types only
    double getd(int r, int c) const {
        const uchar* p = data + (size_t)r * step + (size_t)c * elemSize();
        switch (depth()) {
            case CV_8U: return *p;
            case CV_8S: return *reinterpret_cast<const signed char*>(p);
            case CV_16U: return *reinterpret_cast<const ushort*>(p);
            case CV_16S: return *reinterpret_cast<const short*>(p);
            case CV_32S: return *reinterpret_cast<const int*>(p);
            case CV_32F: return *reinterpret_cast<const float*>(p);
            default: return *reinterpret_cast<const double*>(p);
        }
    }
    void setd(int r, int c, double v) {
        uchar* p = data + (size_t)r * step + (size_t)c * elemSize();
        switch (depth()) {
            case CV_8U: *p = saturate_cast<uchar>(v); break;
            case CV_8S: *reinterpret_cast<signed char*>(p) = (signed char)std::max(-128, std::min(127, shim_round(v))); break;
            case CV_16U: *reinterpret_cast<ushort*>(p) = saturate_cast<ushort>(v); break;
            case CV_16S: *reinterpret_cast<short*>(p) = saturate_cast<short>(v); break;
            case CV_32S: *reinterpret_cast<int*>(p) = saturate_cast<int>(v); break;
            case CV_32F: *reinterpret_cast<float*>(p) = (float)v; break;
            default: *reinterpret_cast<double*>(p) = v; break;
        }
    }
    void setTo(double v) {
        Mat flat = *this;   // channel-flattened view of the same data
        flat.type_ = CV_MAKETYPE(depth(), 1);
        flat.cols = cols * channels();
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < flat.cols; c++) flat.setd(r, c, v);
    }

    // views share the storage
    Mat sub(int r0, int r1, int c0, int c1) const {
        Mat m = *this;
        m.rows = r1 - r0; m.cols = c1 - c0;
        m.data = data + (size_t)r0 * step + (size_t)c0 * elemSize();
        return m;
    }
    Mat row(int r) const { return sub(r, r + 1, 0, cols); }
    Mat col(int c) const { return sub(0, rows, c, c + 1); }
    Mat rowRange(int a, int b) const { return sub(a, b, 0, cols); }
    Mat colRange(int a, int b) const { return sub(0, rows, a, b); }

    Mat clone() const {
        Mat m;
        copyTo(m);
        return m;
    }
    // copies the data; an output of matching size and type is written in place (also a temporary view header)
    void copyTo(Mat& m) const {
        if (!(m.data && m.rows == rows && m.cols == cols && m.type_ == type_)) { m = Mat(); m.create(rows, cols, type_); }
        copy_into(m);
    }
    void copyTo(const Mat& m) const {
        if (!(m.data && m.rows == rows && m.cols == cols && m.type_ == type_)) throw std::runtime_error("shim: copyTo into a temporary of different shape");
        copy_into(const_cast<Mat&>(m));
    }
    void convertTo(Mat& m, int rtype) const {
        const int t = CV_MAKETYPE(CV_MAT_DEPTH(rtype), channels());
        Mat src = *this;   // keeps the data alive if &m == this
        Mat dst;
        if (m.data && m.rows == rows && m.cols == cols && m.type_ == t && m.data != data) dst = m;
        else dst.create(rows, cols, t);
        Mat s1 = src, d1 = dst;
        s1.type_ = CV_MAKETYPE(src.depth(), 1); s1.cols = cols * channels();
        d1.type_ = CV_MAKETYPE(dst.depth(), 1); d1.cols = cols * channels();
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < s1.cols; c++) d1.setd(r, c, s1.getd(r, c));
        m = dst;
    }
    MatExpr t() const;
    MatExpr inv(int method = 0) const;
    static MatExpr eye(int r, int c, int type);
    static MatExpr zeros(int r, int c, int type);
    static MatExpr ones(int r, int c, int type);

    int type_;
    std::shared_ptr<std::vector<uchar>> buf_;

private:
    void copy_into(Mat& m) const {
        const size_t rb = (size_t)cols * elemSize();
        if (m.data == data) return;
        for (int r = 0; r < rows; r++) std::memmove(m.data + (size_t)r * m.step, data + (size_t)r * step, rb);
    }
};

// A matrix expression.  Evaluated eagerly (it simply wraps its value); the distinct TYPE is what matters: assigning a
// MatExpr writes through into an existing matrix of the same shape, assigning a Mat re-binds the header.
class MatExpr {
public:
    Mat m;
    MatExpr() {}
    MatExpr(const Mat& a) : m(a) {}   // implicit: lets one operator set serve Mat, Mat_<T> and MatExpr operands
    template <typename T> MatExpr(const Mat_<T>& a);
    MatExpr t() const { return m.t(); }
    MatExpr inv(int method = 0) const { return m.inv(method); }
    Size size() const { return m.size(); }
    template <typename T> operator Mat_<T>() const;
};

inline Mat::Mat(const MatExpr& e) : Mat(e.m) {}
inline Mat& Mat::operator=(const MatExpr& e) {
    if (data && rows == e.m.rows && cols == e.m.cols && type_ == e.m.type_) {
        if (e.m.data != data) {
            Mat tmp = e.m.clone();   // the expression may alias this matrix through another view
            tmp.copyTo(*this);
        }
    } else {
        *this = e.m;
    }
    return *this;
}

inline Mat shim_new_like(int r, int c, int depth) { return Mat(r, c, CV_MAKETYPE(depth, 1)); }
inline int shim_res_depth(const Mat& a, const Mat& b) { return std::max(a.depth(), b.depth()) >= CV_64F ? CV_64F : (std::max(a.depth(), b.depth()) == CV_32F ? CV_32F : std::max(a.depth(), b.depth())); }

inline MatExpr Mat::t() const {
    Mat r = shim_new_like(cols, rows, depth());
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++) r.setd(j, i, getd(i, j));
    return MatExpr(r);
}
inline MatExpr Mat::eye(int r, int c, int type) {
    Mat m(r, c, type);
    for (int i = 0; i < std::min(r, c); i++) m.setd(i, i, 1.0);
    return MatExpr(m);
}
inline MatExpr Mat::zeros(int r, int c, int type) { return MatExpr(Mat(r, c, type)); }
inline MatExpr Mat::ones(int r, int c, int type) { Mat m(r, c, type); m.setTo(1.0); return MatExpr(m); }

// determinant / inverse: Gaussian elimination with partial pivoting in double (cv::invert DECOMP_LU / cv::determinant)
inline double determinant(const MatExpr& e) {
    const Mat& a = e.m;
    const int n = a.rows;
    std::vector<double> w((size_t)n * n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) w[(size_t)i * n + j] = a.getd(i, j);
    if (n == 3)
        return w[0] * (w[4] * w[8] - w[5] * w[7]) - w[1] * (w[3] * w[8] - w[5] * w[6]) + w[2] * (w[3] * w[7] - w[4] * w[6]);
    double det = 1;
    for (int k = 0; k < n; k++) {
        int p = k;
        for (int i = k + 1; i < n; i++)
            if (std::fabs(w[(size_t)i * n + k]) > std::fabs(w[(size_t)p * n + k])) p = i;
        if (w[(size_t)p * n + k] == 0) return 0;
        if (p != k) { for (int j = 0; j < n; j++) std::swap(w[(size_t)p * n + j], w[(size_t)k * n + j]); det = -det; }
        det *= w[(size_t)k * n + k];
        for (int i = k + 1; i < n; i++) {
            const double f = w[(size_t)i * n + k] / w[(size_t)k * n + k];
            for (int j = k; j < n; j++) w[(size_t)i * n + j] -= f * w[(size_t)k * n + j];
        }
    }
    return det;
}
inline MatExpr Mat::inv(int) const {
    const int n = rows;
    std::vector<double> w((size_t)n * 2 * n, 0.0);
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) w[(size_t)i * 2 * n + j] = getd(i, j);
        w[(size_t)i * 2 * n + n + i] = 1.0;
    }
    bool ok = (rows == cols);
    for (int k = 0; k < n && ok; k++) {
        int p = k;
        for (int i = k + 1; i < n; i++)
            if (std::fabs(w[(size_t)i * 2 * n + k]) > std::fabs(w[(size_t)p * 2 * n + k])) p = i;
        if (w[(size_t)p * 2 * n + k] == 0) { ok = false; break; }
        if (p != k) for (int j = 0; j < 2 * n; j++) std::swap(w[(size_t)p * 2 * n + j], w[(size_t)k * 2 * n + j]);
        const double d = 1.0 / w[(size_t)k * 2 * n + k];
        for (int j = 0; j < 2 * n; j++) w[(size_t)k * 2 * n + j] *= d;
        for (int i = 0; i < n; i++) {
            if (i == k) continue;
            const double f = w[(size_t)i * 2 * n + k];
            if (f != 0) for (int j = 0; j < 2 * n; j++) w[(size_t)i * 2 * n + j] -= f * w[(size_t)k * 2 * n + j];
        }
    }
    Mat r = shim_new_like(n, n, depth() == CV_32F ? CV_32F : CV_64F);
    if (ok)
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) r.setd(i, j, w[(size_t)i * 2 * n + n + j]);
    return MatExpr(r);   // singular: zeros, as cv::invert
}

// ---------------------------------------------------------------- arithmetic
inline MatExpr shim_binary(const Mat& a, const Mat& b, int op) {
    if (a.rows != b.rows || a.cols != b.cols) throw std::runtime_error("shim: element-wise operands differ in size");
    Mat r = shim_new_like(a.rows, a.cols, shim_res_depth(a, b));
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) {
            const double x = a.getd(i, j), y = b.getd(i, j);
            r.setd(i, j, op == 0 ? x + y : x - y);
        }
    return MatExpr(r);
}
inline MatExpr shim_scale(const Mat& a, double s, bool divide) {
    Mat r = shim_new_like(a.rows, a.cols, a.depth());
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) r.setd(i, j, divide ? a.getd(i, j) / s : a.getd(i, j) * s);
    return MatExpr(r);
}
inline MatExpr operator+(const MatExpr& a, const MatExpr& b) { return shim_binary(a.m, b.m, 0); }
inline MatExpr operator-(const MatExpr& a, const MatExpr& b) { return shim_binary(a.m, b.m, 1); }
inline MatExpr operator-(const MatExpr& a) { return shim_scale(a.m, -1.0, false); }
inline MatExpr operator*(const MatExpr& a, const MatExpr& b) {   // matrix product (gemm), accumulated in double
    const Mat &x = a.m, &y = b.m;
    if (x.cols != y.rows) throw std::runtime_error("shim: matrix product of incompatible shapes");
    Mat r = shim_new_like(x.rows, y.cols, shim_res_depth(x, y));
    for (int i = 0; i < x.rows; i++)
        for (int j = 0; j < y.cols; j++) {
            double s = 0;
            for (int k = 0; k < x.cols; k++) s += x.getd(i, k) * y.getd(k, j);
            r.setd(i, j, s);
        }
    return MatExpr(r);
}
inline MatExpr operator*(const MatExpr& a, double s) { return shim_scale(a.m, s, false); }
inline MatExpr operator*(double s, const MatExpr& a) { return shim_scale(a.m, s, false); }
inline MatExpr operator/(const MatExpr& a, double s) { return shim_scale(a.m, s, true); }
inline MatExpr operator!=(const MatExpr& a, const MatExpr& b) {
    Mat r(a.m.rows, a.m.cols, CV_8U);
    for (int i = 0; i < a.m.rows; i++)
        for (int j = 0; j < a.m.cols; j++) r.at<uchar>(i, j) = (a.m.getd(i, j) != b.m.getd(i, j)) ? 255 : 0;
    return MatExpr(r);
}
// compound assignment takes const references on purpose: the reference applies them to temporary views
// (cnn_softam.h:641, train_ransac_softam.cpp:349,369), exactly as OpenCV's own operators allow
inline const Mat& operator+=(const Mat& a, const MatExpr& b) { Mat r = shim_binary(a, b.m, 0).m; Mat dst = a; Mat c; r.convertTo(c, a.depth()); c.copyTo(dst); return a; }
inline const Mat& operator-=(const Mat& a, const MatExpr& b) { Mat r = shim_binary(a, b.m, 1).m; Mat dst = a; Mat c; r.convertTo(c, a.depth()); c.copyTo(dst); return a; }
inline const Mat& operator*=(const Mat& a, double s) { Mat r = shim_scale(a, s, false).m; Mat dst = a; r.copyTo(dst); return a; }
inline const Mat& operator/=(const Mat& a, double s) { Mat r = shim_scale(a, s, true).m; Mat dst = a; r.copyTo(dst); return a; }

inline Scalar sum(const MatExpr& e) {
    double s = 0;
    for (int i = 0; i < e.m.rows; i++)
        for (int j = 0; j < e.m.cols; j++) s += e.m.getd(i, j);
    return Scalar(s);
}
inline Scalar trace(const MatExpr& e) {
    double s = 0;
    for (int i = 0; i < std::min(e.m.rows, e.m.cols); i++) s += e.m.getd(i, i);
    return Scalar(s);
}
inline double norm(const MatExpr& e) {
    double s = 0;
    for (int i = 0; i < e.m.rows; i++)
        for (int j = 0; j < e.m.cols; j++) s += e.m.getd(i, j) * e.m.getd(i, j);
    return std::sqrt(s);
}
inline std::ostream& operator<<(std::ostream& os, const Mat& m) {
    os << "[";
    for (int i = 0; i < m.rows; i++) {
        for (int j = 0; j < m.cols; j++) os << m.getd(i, j) << (j + 1 < m.cols ? ", " : "");
        os << (i + 1 < m.rows ? ";\n " : "");
    }
    return os << "]";
}

// ---------------------------------------------------------------- Mat_<T>
template <typename T> class Mat_ : public Mat {
public:
    Mat_() : Mat() { type_ = DataType<T>::type; }
    Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
    Mat_(int r, int c, const T& v) : Mat(r, c, DataType<T>::type) { for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) (*this)(i, j) = v; }
    explicit Mat_(Size s) : Mat(s.height, s.width, DataType<T>::type) {}
    Mat_(const Mat& m) : Mat() { type_ = DataType<T>::type; assign(m); }
    Mat_(const Mat_& m) = default;
    Mat_(const MatExpr& e) : Mat() { type_ = DataType<T>::type; assign(e.m); }
    Mat_& operator=(const Mat_& m) = default;
    Mat_& operator=(const Mat& m) { assign(m); return *this; }
    Mat_& operator=(const MatExpr& e) {
        if (e.m.type() == (int)DataType<T>::type) Mat::operator=(e);
        else { Mat c; e.m.convertTo(c, DataType<T>::depth); Mat::operator=(MatExpr(c)); }
        return *this;
    }
    T& operator()(int r, int c) { return this->template at<T>(r, c); }
    const T& operator()(int r, int c) const { return this->template at<T>(r, c); }
    T& operator()(int i) { return this->template at<T>(i); }
    const T& operator()(int i) const { return this->template at<T>(i); }
    Mat_ clone() const { return Mat_(Mat::clone()); }
    Mat_ row(int r) const { return Mat_(Mat::row(r)); }
    Mat_ col(int c) const { return Mat_(Mat::col(c)); }
    Mat_ rowRange(int a, int b) const { return Mat_(Mat::rowRange(a, b)); }
    Mat_ colRange(int a, int b) const { return Mat_(Mat::colRange(a, b)); }
    static MatExpr zeros(int r, int c) { return Mat::zeros(r, c, DataType<T>::type); }
    static MatExpr zeros(Size s) { return Mat::zeros(s.height, s.width, DataType<T>::type); }
    static MatExpr ones(int r, int c) { return Mat::ones(r, c, DataType<T>::type); }
    static MatExpr eye(int r, int c) { return Mat::eye(r, c, DataType<T>::type); }

private:
    void assign(const Mat& m) {   // same type: share the data; other type: convert
        if (m.empty() && m.type() != (int)DataType<T>::type) { Mat::operator=(Mat()); type_ = DataType<T>::type; return; }
        if (m.type() == (int)DataType<T>::type) Mat::operator=(m);
        else { Mat c; m.convertTo(c, DataType<T>::depth); Mat::operator=(c); }
    }
};
template <typename T> inline MatExpr::MatExpr(const Mat_<T>& a) : m(static_cast<const Mat&>(a)) {}
template <typename T> inline MatExpr::operator Mat_<T>() const { return Mat_<T>(m); }

// ---------------------------------------------------------------- statistics
inline void meanStdDev(const std::vector<double>& v, std::vector<double>& mean, std::vector<double>& stddev) {
    double s = 0, sq = 0;
    for (double x : v) { s += x; sq += x * x; }
    const double n = v.empty() ? 1.0 : (double)v.size();
    const double mu = s / n;
    mean.assign(1, mu);
    stddev.assign(1, std::sqrt(std::max(sq / n - mu * mu, 0.0)));
}

// ---------------------------------------------------------------- SVD (3x3 only: Hypothesis.cpp:178)
class SVD {
public:
    Mat u, w, vt;
    SVD() {}
    explicit SVD(const Mat& a) {
        if (a.rows != 3 || a.cols != 3) throw std::runtime_error("shim: cv::SVD is implemented for 3x3 matrices only");
        double A[9], U[9], W[3], Vt[9];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) A[i * 3 + j] = a.getd(i, j);
        orc_svd3(A, U, W, Vt);
        u = Mat(3, 3, CV_64F); vt = Mat(3, 3, CV_64F); w = Mat(3, 1, CV_64F);
        for (int i = 0; i < 3; i++) {
            w.at<double>(i, 0) = W[i];
            for (int j = 0; j < 3; j++) { u.at<double>(i, j) = U[i * 3 + j]; vt.at<double>(i, j) = Vt[i * 3 + j]; }
        }
    }
};

// ---------------------------------------------------------------- calib3d, forwarded to the oracle's restated routines
inline void shim_out(const Mat& dst_in, Mat& dst_ref, bool is_const, const Mat& value) {
    // cv::OutputArray::create semantics: an output of matching size and type is written in place
    if (dst_in.data && dst_in.rows == value.rows && dst_in.cols == value.cols && dst_in.type() == value.type()) value.copyTo(const_cast<Mat&>(dst_in));
    else if (!is_const) dst_ref = value.clone();
    else throw std::runtime_error("shim: output into a const matrix of different shape");
}

inline void shim_rodrigues(const Mat& src, Mat& dst, Mat* jac, const Mat* dst_const, const Mat* jac_const) {
    if (src.total() == 3) {   // vector -> matrix, Jacobian 3x9
        double r[3], R[9], J[27];
        for (int i = 0; i < 3; i++) r[i] = src.rows == 1 ? src.getd(0, i) : src.getd(i, 0);
        orc_rodrigues(r, R, J);
        Mat Rm(3, 3, CV_64F);
        for (int i = 0; i < 9; i++) Rm.at<double>(i / 3, i % 3) = R[i];
        shim_out(dst_const ? *dst_const : dst, dst, dst_const != nullptr, Rm);
        if (jac || jac_const) {
            Mat Jm(3, 9, CV_64F);
            for (int i = 0; i < 27; i++) Jm.at<double>(i / 9, i % 9) = J[i];
            shim_out(jac_const ? *jac_const : *jac, jac ? *jac : const_cast<Mat&>(*jac_const), jac == nullptr, Jm);
        }
    } else {                  // matrix -> vector (3x1)
        double R[9], r[3];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[i * 3 + j] = src.getd(i, j);
        orc_rodrigues_inv(R, r);
        Mat rv(3, 1, CV_64F);
        for (int i = 0; i < 3; i++) rv.at<double>(i, 0) = r[i];
        shim_out(dst_const ? *dst_const : dst, dst, dst_const != nullptr, rv);
        if (jac || jac_const) throw std::runtime_error("shim: Jacobian of the matrix -> vector Rodrigues is not implemented");
    }
}
inline void Rodrigues(const Mat& src, Mat& dst) { shim_rodrigues(src, dst, nullptr, nullptr, nullptr); }
inline void Rodrigues(const Mat& src, Mat& dst, Mat& jac) { shim_rodrigues(src, dst, &jac, nullptr, nullptr); }
// a const matrix handed to an OutputArray (cnn_softam.h:508: the const& rotation is re-created from its Rodrigues vector)
inline void Rodrigues(const Mat& src, const Mat& dst, Mat& jac) { Mat dummy; shim_rodrigues(src, dummy, &jac, &dst, nullptr); }

inline void shim_intrinsics(const Mat& K, double* f, double* cx, double* cy) {
    *f = K.getd(0, 0); *cx = K.getd(0, 2); *cy = K.getd(1, 2);
    if (K.getd(1, 1) != *f) throw std::runtime_error("shim: fx != fy is not supported");
}
inline void shim_vec3(const Mat& m, double v[3]) {
    for (int i = 0; i < 3; i++) v[i] = m.rows == 1 ? m.getd(0, i) : m.getd(i, 0);
}

inline void projectPoints(const std::vector<Point3f>& obj, const Mat& rvec, const Mat& tvec, const Mat& K, const Mat& /*dist*/,
                          std::vector<Point2f>& out) {
    double f, cx, cy, r[3], t[3];
    shim_intrinsics(K, &f, &cx, &cy);
    shim_vec3(rvec, r); shim_vec3(tvec, t);
    const int n = (int)obj.size();
    std::vector<double> X((size_t)n * 3), uv((size_t)n * 2);
    for (int i = 0; i < n; i++) { X[i * 3] = obj[i].x; X[i * 3 + 1] = obj[i].y; X[i * 3 + 2] = obj[i].z; }
    orc_project_points(n, X.data(), r, t, f, cx, cy, uv.data(), nullptr, nullptr);
    out.resize(n);
    for (int i = 0; i < n; i++) out[i] = Point2f((float)uv[i * 2], (float)uv[i * 2 + 1]);
}

inline bool solvePnP(const std::vector<Point3f>& obj, const std::vector<Point2f>& img, const Mat& K, const Mat& /*dist*/, Mat& rvec,
                     Mat& tvec, bool useExtrinsicGuess = false, int flags = CV_ITERATIVE) {
    double f, cx, cy, r[3] = {0, 0, 0}, t[3] = {0, 0, 0};
    shim_intrinsics(K, &f, &cx, &cy);
    const int n = (int)obj.size();
    std::vector<float> o((size_t)n * 3), im((size_t)n * 2);
    for (int i = 0; i < n; i++) {
        o[i * 3] = obj[i].x; o[i * 3 + 1] = obj[i].y; o[i * 3 + 2] = obj[i].z;
        im[i * 2] = img[i].x; im[i * 2 + 1] = img[i].y;
    }
    int ok;
    if (flags == CV_P3P) {
        if (n != 4) throw std::runtime_error("shim: CV_P3P needs exactly 4 correspondences");
        ok = orc_solve_p3p(o.data(), im.data(), f, cx, cy, r, t);
    } else if (flags == CV_ITERATIVE) {
        if (!useExtrinsicGuess) throw std::runtime_error("shim: CV_ITERATIVE without an extrinsic guess is not used by the reference's pipeline");
        shim_vec3(rvec, r); shim_vec3(tvec, t);
        int iters = 0;
        ok = orc_solve_pnp_iterative(n, o.data(), im.data(), f, cx, cy, r, t, &iters);
    } else {
        throw std::runtime_error("shim: solvePnP flag not implemented");
    }
    if (!ok && flags == CV_P3P) return false;
    Mat rv(3, 1, CV_64F), tv(3, 1, CV_64F);
    for (int i = 0; i < 3; i++) { rv.at<double>(i, 0) = r[i]; tv.at<double>(i, 0) = t[i]; }
    shim_out(rvec, rvec, false, rv);
    shim_out(tvec, tvec, false, tv);
    return ok != 0;
}

}  // namespace cv
