// types.h -- the value types of the reference's C++ surface (core/types.h) without OpenCV.
//
// Same names and conventions as /root/reference/core/types.h so that code written against the
// reference's headers reads the same: jp::coord3_t / img_coord_t (int16 mm, types.h:43-51),
// cv_trans_t = (rvec, tvec) and jp_trans_t = (R, t) (types.h:91-92), cv2our / our2cv (types.h:137-214).
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

#define EPS 0.00000001   // core/types.h:32
#define PI 3.1415926     // core/types.h:33

namespace cvlite {
// minimal stand-ins for the cv:: value types the reference's signatures use
struct Point2i { int x = 0, y = 0; Point2i() {} Point2i(int x_, int y_) : x(x_), y(y_) {} };
struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct Point3f { float x = 0, y = 0, z = 0; Point3f() {} Point3f(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {} };
struct Point3d {
    double x = 0, y = 0, z = 0;
    Point3d() {}
    Point3d(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
    Point3d operator+(const Point3d& o) const { return {x + o.x, y + o.y, z + o.z}; }
    Point3d operator-(const Point3d& o) const { return {x - o.x, y - o.y, z - o.z}; }
    Point3d operator-() const { return {-x, -y, -z}; }
    Point3d operator*(double s) const { return {x * s, y * s, z * s}; }
};
inline double norm(const Point3d& p) { return std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z); }

typedef std::array<double, 3> Vec3d;

// row-major dense matrix, enough of cv::Mat_<T> for this surface
template <typename T>
class Mat_ {
public:
    int rows = 0, cols = 0;
    Mat_() {}
    Mat_(int r, int c) : rows(r), cols(c), d_((size_t)r * c) {}
    static Mat_ zeros(int r, int c) { Mat_ m(r, c); for (auto& v : m.d_) v = T(); return m; }
    static Mat_ eye(int r, int c) { Mat_ m = zeros(r, c); for (int i = 0; i < (r < c ? r : c); i++) m(i, i) = T(1); return m; }
    T& operator()(int y, int x) { return d_[(size_t)y * cols + x]; }
    const T& operator()(int y, int x) const { return d_[(size_t)y * cols + x]; }
    T* data() { return d_.data(); }
    const T* data() const { return d_.data(); }
    bool empty() const { return d_.empty(); }
    Mat_ t() const { Mat_ m(cols, rows); for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) m(j, i) = (*this)(i, j); return m; }
    Mat_ operator*(const Mat_& o) const {
        Mat_ m = zeros(rows, o.cols);
        for (int i = 0; i < rows; i++) for (int k = 0; k < cols; k++) for (int j = 0; j < o.cols; j++) m(i, j) += (*this)(i, k) * o(k, j);
        return m;
    }
private:
    std::vector<T> d_;
};
typedef Mat_<double> Matd;

double determinant3(const Matd& m);
Matd inv(const Matd& m);                    // general n x n inverse (Gauss-Jordan, n <= 4 here)
void Rodrigues(const Vec3d& rvec, Matd& R);  // vector -> 3x3
void Rodrigues(const Matd& R, Vec3d& rvec);  // 3x3 -> vector
}  // namespace cvlite

namespace jp {
typedef unsigned char id_t;
typedef short coord1_t;                                   // one dimension, millimetres
struct coord3_t { coord1_t v[3]; coord1_t& operator()(int i) { return v[i]; } coord1_t operator()(int i) const { return v[i]; } };
typedef cvlite::Mat_<coord3_t> img_coord_t;               // object / scene coordinate images

struct info_t {                                           // ground truth per image (types.h:65-88)
    cvlite::Mat_<float> rotation;                         // 3x3
    float center[3];                                      // metres
    bool visible;
    info_t(bool v = true) : rotation(cvlite::Mat_<float>::eye(3, 3)), visible(v) { center[0] = 0; center[1] = 0; center[2] = -1; }
};

typedef std::pair<cvlite::Vec3d, cvlite::Vec3d> cv_trans_t;   // (rvec, tvec [mm]) as OpenCV expects it
typedef std::pair<cvlite::Matd, cvlite::Point3d> jp_trans_t;  // (R, t [mm]) in the reference's own convention

cv_trans_t our2cv(const jp_trans_t& trans);   // types.h:137-151
jp_trans_t cv2our(const cv_trans_t& trans);   // types.h:186-214
}  // namespace jp
