// Hypothesis.cpp -- see Hypothesis.h.  Behaviour follows /root/reference/core/Hypothesis.cpp
// (cited per function); arithmetic is plain C++ on cvlite types instead of cv::Mat.
#include "Hypothesis.h"

#include <cassert>
#include <cstring>

#define DSAC_HOST_ONLY 1
#include "../csrc/pose_math.cuh"

using namespace cvlite;

namespace cvlite {

double determinant3(const Matd& m) {
    return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) +
           m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
}

Matd inv(const Matd& m) {
    const int n = m.rows;
    Matd a = m, b = Matd::eye(n, n);
    for (int c = 0; c < n; c++) {
        int piv = c;
        for (int r = c + 1; r < n; r++)
            if (std::fabs(a(r, c)) > std::fabs(a(piv, c))) piv = r;
        for (int k = 0; k < n; k++) { std::swap(a(c, k), a(piv, k)); std::swap(b(c, k), b(piv, k)); }
        double d = 1.0 / a(c, c);
        for (int k = 0; k < n; k++) { a(c, k) *= d; b(c, k) *= d; }
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            double f = a(r, c);
            for (int k = 0; k < n; k++) { a(r, k) -= f * a(c, k); b(r, k) -= f * b(c, k); }
        }
    }
    return b;
}

void Rodrigues(const Vec3d& rvec, Matd& R) {
    double Rm[9];
    dsac::rodrigues_v2m(rvec.data(), Rm);
    R = Matd(3, 3);
    std::memcpy(R.data(), Rm, sizeof(Rm));
}

void Rodrigues(const Matd& R, Vec3d& rvec) { dsac::rodrigues_m2v(R.data(), rvec.data()); }

}  // namespace cvlite

namespace jp {
// types.h:137-151: rows 1,2 of R and y,z of t change sign (180 deg about x), then matrix -> Rodrigues vector
cv_trans_t our2cv(const jp_trans_t& trans) {
    Matd rmat = trans.first;
    for (int j = 0; j < 3; j++) { rmat(1, j) = -rmat(1, j); rmat(2, j) = -rmat(2, j); }
    Vec3d rvec;
    Rodrigues(rmat, rvec);
    return cv_trans_t(rvec, Vec3d{trans.second.x, -trans.second.y, -trans.second.z});
}
// types.h:186-214: the inverse map, with the det < 0 flip and NaN translation -> 0
jp_trans_t cv2our(const cv_trans_t& trans) {
    Matd rmat;
    Rodrigues(trans.first, rmat);
    Point3d tpt(trans.second[0], -trans.second[1], -trans.second[2]);
    for (int j = 0; j < 3; j++) { rmat(1, j) = -rmat(1, j); rmat(2, j) = -rmat(2, j); }
    if (determinant3(rmat) < 0) {
        tpt = -tpt;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rmat(i, j) = -rmat(i, j);
    }
    if (tpt.x != tpt.x || tpt.y != tpt.y || tpt.z != tpt.z) tpt = Point3d(0, 0, 0);
    return jp_trans_t(rmat, tpt);
}
}  // namespace jp

Hypothesis::Hypothesis() : rotation(Matd::eye(3, 3)), invRotation(Matd::eye(3, 3)), translation(0, 0, 0) {}

Hypothesis::Hypothesis(Matd rot, Point3d trans) : rotation(rot), invRotation(inv(rot)), translation(trans) {}

Hypothesis::Hypothesis(jp::info_t info) {
    rotation = Matd(3, 3);
    for (int y = 0; y < 3; y++) for (int x = 0; x < 3; x++) rotation(y, x) = info.rotation(y, x);
    translation = Point3d(info.center[0] * 1e3, info.center[1] * 1e3, info.center[2] * 1e3);  // m -> mm
    invRotation = inv(rotation);
}

Hypothesis::Hypothesis(Matd transform) : rotation(Matd::eye(3, 3)) {
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) rotation(a, b) = transform(a, b);
    translation = Point3d(transform(0, 3), transform(1, 3), transform(2, 3));
    invRotation = inv(rotation);
}

Hypothesis::Hypothesis(std::vector<std::pair<Point3d, Point3d>> pts) { refine(pts); }

Hypothesis::Hypothesis(std::vector<double> v) {
    assert(v.size() == 6);
    translation = Point3d(v[3], v[4], v[5]);
    double length = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (length > 1e-5) Rodrigues(Vec3d{v[0], v[1], v[2]}, rotation);   // Hypothesis.cpp:92-95
    else rotation = Matd::eye(3, 3);
    invRotation = inv(rotation);
}

void Hypothesis::setRotation(Matd rot) { rotation = rot; invRotation = inv(rot); }
void Hypothesis::setTranslation(Point3d trans) { translation = trans; }
Point3d Hypothesis::getTranslation() const { return translation; }
Matd Hypothesis::getRotation() const { return rotation; }
Matd Hypothesis::getInvRotation() const { return invRotation; }

static Point3d mul3(const Matd& R, const Point3d& p) {
    return Point3d(R(0, 0) * p.x + R(0, 1) * p.y + R(0, 2) * p.z, R(1, 0) * p.x + R(1, 1) * p.y + R(1, 2) * p.z,
                   R(2, 0) * p.x + R(2, 1) * p.y + R(2, 2) * p.z);
}

Point3d Hypothesis::transform(Point3d p, bool isNormal) {
    Point3d tp = mul3(rotation, p);
    return isNormal ? tp : tp + translation;
}

Point3d Hypothesis::invTransform(Point3d p) { return mul3(invRotation, p - translation); }

// Hypothesis.cpp:137-143
double Hypothesis::calcAngularDistance(const Hypothesis& h) const {
    Matd d = rotation * h.getInvRotation();
    double trace = d(0, 0) + d(1, 1) + d(2, 2);
    trace = std::min(3.0, std::max(-1.0, trace));
    return 180 * std::acos((trace - 1.0) / 2.0) / 3.14159265358979323846;
}

// symmetric 3x3 eigen-decomposition (cyclic Jacobi) -> SVD of the covariance for Kabsch
static void symEig3(const Matd& S, double w[3], Matd& V) {
    Matd A = S;
    V = Matd::eye(3, 3);
    for (int sweep = 0; sweep < 50; sweep++) {
        double off = std::fabs(A(0, 1)) + std::fabs(A(0, 2)) + std::fabs(A(1, 2));
        if (off == 0.0) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (A(p, q) == 0.0) continue;
                double theta = (A(q, q) - A(p, p)) / (2.0 * A(p, q));
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; k++) { double a = A(k, p), b = A(k, q); A(k, p) = c * a - s * b; A(k, q) = s * a + c * b; }
                for (int k = 0; k < 3; k++) { double a = A(p, k), b = A(q, k); A(p, k) = c * a - s * b; A(q, k) = s * a + c * b; }
                A(p, q) = A(q, p) = 0.0;
                for (int k = 0; k < 3; k++) { double a = V(k, p), b = V(k, q); V(k, p) = c * a - s * b; V(k, q) = s * a + c * b; }
            }
    }
    for (int i = 0; i < 3; i++) w[i] = A(i, i);
}

// Hypothesis.cpp:174-200: R = V diag(1,1,sign) U^T from the SVD of the covariance, t = -R cA + cB
std::pair<Matd, Point3d> Hypothesis::calcRigidBodyTransform(Matd& coV, Point3d cA, Point3d cB) {
    Matd S = coV.t() * coV, Vv;
    double w[3];
    symEig3(S, w, Vv);
    int o[3] = {0, 1, 2};
    if (w[o[0]] < w[o[1]]) std::swap(o[0], o[1]);
    if (w[o[1]] < w[o[2]]) std::swap(o[1], o[2]);
    if (w[o[0]] < w[o[1]]) std::swap(o[0], o[1]);
    Matd V(3, 3), U(3, 3);
    for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) V(k, j) = Vv(k, o[j]);
    for (int j = 0; j < 3; j++) {
        double u[3], nn = 0;
        for (int k = 0; k < 3; k++) { u[k] = coV(k, 0) * V(0, j) + coV(k, 1) * V(1, j) + coV(k, 2) * V(2, j); nn += u[k] * u[k]; }
        if (j == 2 && !(nn > 1e-24 * (w[o[0]] > 0 ? w[o[0]] : 1.0))) {   // rank-deficient: complete the frame
            U(0, 2) = U(1, 0) * U(2, 1) - U(2, 0) * U(1, 1);
            U(1, 2) = U(2, 0) * U(0, 1) - U(0, 0) * U(2, 1);
            U(2, 2) = U(0, 0) * U(1, 1) - U(1, 0) * U(0, 1);
            break;
        }
        nn = nn > 0 ? 1.0 / std::sqrt(nn) : 0.0;
        for (int k = 0; k < 3; k++) U(k, j) = u[k] * nn;
    }
    double sign = determinant3(V * U.t()) < 0 ? -1 : 1;
    Matd dm = Matd::eye(3, 3);
    dm(2, 2) = sign;
    Matd R = V * dm * U.t();
    Point3d RcA = mul3(R, cA);
    return std::make_pair(R, Point3d(-RcA.x + cB.x, -RcA.y + cB.y, -RcA.z + cB.z));
}

// Hypothesis.cpp:145-172
std::pair<Matd, Point3d> Hypothesis::calcRigidBodyTransform(std::vector<std::pair<Point3d, Point3d>> pts) {
    Point3d cA(0, 0, 0), cB(0, 0, 0);
    for (auto& p : pts) { cA = cA + p.first; cB = cB + p.second; }
    cA = cA * (1.0 / (double)pts.size());
    cB = cB * (1.0 / (double)pts.size());
    Matd a = Matd::zeros(3, 3);
    for (auto& p : pts) {
        double A[3] = {p.first.x - cA.x, p.first.y - cA.y, p.first.z - cA.z};
        double B[3] = {p.second.x - cB.x, p.second.y - cB.y, p.second.z - cB.z};
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) a(r, c) += A[r] * B[c];
    }
    return calcRigidBodyTransform(a, cA, cB);
}

void Hypothesis::refine(std::vector<std::pair<Point3d, Point3d>> pts) {
    points.insert(points.end(), pts.begin(), pts.end());
    auto est = calcRigidBodyTransform(pts);
    rotation = est.first;
    translation = est.second;
    invRotation = inv(rotation);
}

void Hypothesis::refine(Matd& coV, Point3d pointsA, Point3d pointsB) {
    auto est = calcRigidBodyTransform(coV, pointsA, pointsB);
    rotation = est.first;
    translation = est.second;
    invRotation = inv(rotation);
}

Matd Hypothesis::getTransformation() const {
    Matd r = Matd::zeros(4, 4);
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) r(a, b) = rotation(a, b);
    r(0, 3) = translation.x; r(1, 3) = translation.y; r(2, 3) = translation.z; r(3, 3) = 1.0;
    return r;
}

Hypothesis Hypothesis::getInv() { return Hypothesis(inv(getTransformation())); }
Hypothesis Hypothesis::operator*(const Hypothesis& o) const { return Hypothesis(getTransformation() * o.getTransformation()); }
Hypothesis Hypothesis::operator/(const Hypothesis& o) const { return Hypothesis(getTransformation() * inv(o.getTransformation())); }

Vec3d Hypothesis::getRodriguesVector() const {
    Vec3d r;
    Rodrigues(rotation, r);
    return r;
}

std::vector<double> Hypothesis::getRodVecAndTrans() const {   // Hypothesis.cpp:274-290
    Vec3d rv = getRodriguesVector();
    return {rv[0], rv[1], rv[2], translation.x, translation.y, translation.z};
}
