/*
 * dsac_oracle.cpp -- CPU ORACLE (test infrastructure; see dsac_oracle.h header).
 *
 * Every function cites the reference file:line (relative to /root/reference/) whose
 * behaviour it restates.  The OpenCV-2.4 routines the reference calls are not in
 * /root/reference; they are restated from the published algorithms:
 *   - Rodrigues + its 3x9 Jacobian (cv::Rodrigues),
 *   - pin-hole projection + Jacobians (cv::projectPoints, zero distortion),
 *   - P3P of Gao, Hou, Tang, Cheng (PAMI 2003) with Horn's quaternion absolute
 *     orientation and 4th-point disambiguation (cv::solvePnP CV_P3P),
 *   - Levenberg-Marquardt pose refinement, CvLevMarq schedule (cv::solvePnP CV_ITERATIVE
 *     with useExtrinsicGuess=true),
 *   - one-sided Jacobi SVD (cv::SVD).
 * and pinned against cv2 4.13 golden vectors in tests/.
 */
#include "dsac_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <random>
#include <vector>
#include <atomic>
#include <thread>

namespace {

constexpr double kEps = 1e-8;  // EPS, types.h:32
constexpr double kPi = 3.14159265358979323846;  // CV_PI

// ------------------------------------------------------------------ small linear algebra
inline void mat3_mul(const double* A, const double* B, double* C) {
    double T[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) T[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
    std::memcpy(C, T, sizeof(T));
}
inline void mat3_t(const double* A, double* T) {
    double B[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) B[i * 3 + j] = A[j * 3 + i];
    std::memcpy(T, B, sizeof(B));
}
inline double det3(const double* A) {
    return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}
inline void mat3_vec(const double* A, const double* x, double* y) {
    double t0 = A[0] * x[0] + A[1] * x[1] + A[2] * x[2];
    double t1 = A[3] * x[0] + A[4] * x[1] + A[5] * x[2];
    double t2 = A[6] * x[0] + A[7] * x[1] + A[8] * x[2];
    y[0] = t0; y[1] = t1; y[2] = t2;
}

// One-sided (Hestenes) Jacobi SVD of an n x n matrix, n <= 6: A = U diag(w) V^T,
// singular values sorted descending.  Stands in for cv::SVD (Hypothesis.cpp:178 and
// inside cv::Rodrigues / CvLevMarq::step).
void jacobi_svd(int n, const double* A, double* U, double* w, double* V) {
    double G[36], Vm[36];  // G columns are rotated until mutually orthogonal
    for (int i = 0; i < n * n; i++) G[i] = A[i];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) Vm[i * n + j] = (i == j) ? 1.0 : 0.0;
    const double eps = std::numeric_limits<double>::epsilon() * 2;
    for (int sweep = 0; sweep < 60; sweep++) {
        bool changed = false;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double a = 0, b = 0, c = 0;
                for (int k = 0; k < n; k++) {
                    a += G[k * n + p] * G[k * n + p];
                    b += G[k * n + q] * G[k * n + q];
                    c += G[k * n + p] * G[k * n + q];
                }
                if (std::fabs(c) <= eps * std::sqrt(a * b)) continue;
                // rotation that zeroes the (p,q) inner product
                double zeta = (b - a) / (2.0 * c);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
                for (int k = 0; k < n; k++) {
                    double gp = G[k * n + p], gq = G[k * n + q];
                    G[k * n + p] = cs * gp - sn * gq;
                    G[k * n + q] = sn * gp + cs * gq;
                    double vp = Vm[k * n + p], vq = Vm[k * n + q];
                    Vm[k * n + p] = cs * vp - sn * vq;
                    Vm[k * n + q] = sn * vp + cs * vq;
                }
                changed = true;
            }
        if (!changed) break;
    }
    int order[6];
    double sv[6];
    for (int j = 0; j < n; j++) {
        double s = 0;
        for (int k = 0; k < n; k++) s += G[k * n + j] * G[k * n + j];
        sv[j] = std::sqrt(s);
        order[j] = j;
    }
    std::stable_sort(order, order + n, [&](int x, int y) { return sv[x] > sv[y]; });
    for (int jj = 0; jj < n; jj++) {
        int j = order[jj];
        w[jj] = sv[j];
        for (int k = 0; k < n; k++) {
            V[k * n + jj] = Vm[k * n + j];
            U[k * n + jj] = sv[j] > 0 ? G[k * n + j] / sv[j] : 0.0;
        }
    }
    // complete U for (numerically) zero singular values so that it stays orthonormal
    for (int jj = 0; jj < n; jj++) {
        if (w[jj] > std::numeric_limits<double>::min() * 1e10) continue;
        for (int trial = 0; trial < n; trial++) {
            double v[6];
            for (int k = 0; k < n; k++) v[k] = (k == trial) ? 1.0 : 0.0;
            for (int pass = 0; pass < 2; pass++)
                for (int c2 = 0; c2 < n; c2++) {
                    if (c2 == jj) continue;
                    if (c2 > jj && w[c2] <= std::numeric_limits<double>::min() * 1e10) continue;
                    double d = 0;
                    for (int k = 0; k < n; k++) d += v[k] * U[k * n + c2];
                    for (int k = 0; k < n; k++) v[k] -= d * U[k * n + c2];
                }
            double nn = 0;
            for (int k = 0; k < n; k++) nn += v[k] * v[k];
            if (nn > 1e-6) {
                nn = std::sqrt(nn);
                for (int k = 0; k < n; k++) U[k * n + jj] = v[k] / nn;
                break;
            }
        }
    }
}

// Cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (n <= 6).
void jacobi_eig(int n, const double* Ain, double* evals, double* evecs /* columns */) {
    double A[36];
    for (int i = 0; i < n * n; i++) A[i] = Ain[i];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) evecs[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) off += std::fabs(A[p * n + q]);
        if (off == 0.0) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double apq = A[p * n + q];
                if (apq == 0.0) continue;
                double app = A[p * n + p], aqq = A[q * n + q];
                if (std::fabs(apq) < 1e-300 + 1e-20 * (std::fabs(app) + std::fabs(aqq)) && sweep > 3) {
                    A[p * n + q] = A[q * n + p] = 0.0;
                    continue;
                }
                double theta = (aqq - app) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) {  // A <- A * G
                    double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {  // A <- G^T * A
                    double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                A[p * n + q] = A[q * n + p] = 0.0;
                for (int k = 0; k < n; k++) {
                    double vkp = evecs[k * n + p], vkq = evecs[k * n + q];
                    evecs[k * n + p] = c * vkp - s * vkq;
                    evecs[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; i++) evals[i] = A[i * n + i];
}

// ------------------------------------------------------------------ Rodrigues
// cv::Rodrigues vector -> matrix with the 3x9 Jacobian (row i = d vec(R) / d r_i,
// R flattened row-major).  Call sites: types.h:190, maxloss.h:95-96, cnn_softam.h:507-508.
void rodrigues_v2m(const double r[3], double R[9], double* J) {
    double theta = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < std::numeric_limits<double>::epsilon()) {
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        if (J) {
            std::memset(J, 0, 27 * sizeof(double));
            J[5] = J[15] = J[19] = -1;
            J[7] = J[11] = J[21] = 1;
        }
        return;
    }
    double c = std::cos(theta), s = std::sin(theta), c1 = 1.0 - c, itheta = 1.0 / theta;
    double rx = r[0] * itheta, ry = r[1] * itheta, rz = r[2] * itheta;
    double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int k = 0; k < 9; k++) R[k] = c * I[k] + c1 * rrt[k] + s * r_x[k];
    if (J) {
        double drrt[27] = {rx + rx, ry, rz, ry, 0, 0, rz, 0, 0,
                           0, rx, 0, rx, ry + ry, rz, 0, rz, 0,
                           0, 0, rx, 0, 0, ry, rx, ry, rz + rz};
        const double d_r_x[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0,
                                  0, 0, 1, 0, 0, 0, -1, 0, 0,
                                  0, -1, 0, 1, 0, 0, 0, 0, 0};
        for (int i = 0; i < 3; i++) {
            double ri = (i == 0) ? rx : (i == 1) ? ry : rz;
            double a0 = -s * ri, a1 = (s - 2 * c1 * itheta) * ri, a2 = c1 * itheta;
            double a3 = (c - s * itheta) * ri, a4 = s * itheta;
            for (int k = 0; k < 9; k++)
                J[i * 9 + k] = a0 * I[k] + a1 * rrt[k] + a2 * drrt[i * 9 + k] + a3 * r_x[k] + a4 * d_r_x[i * 9 + k];
        }
    }
}

// cv::Rodrigues matrix -> vector (the matrix is first projected onto SO(3) by an SVD,
// as OpenCV does).  Call sites: Hypothesis.cpp:270, types.h:143, cnn_softam.h:507.
void rodrigues_m2v(const double Rin[9], double r[3]) {
    double U[9], w[3], V[9], R[9];
    jacobi_svd(3, Rin, U, w, V);
    double Vt[9];
    mat3_t(V, Vt);
    mat3_mul(U, Vt, R);
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    double s = std::sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = std::acos(c);
    if (s < 1e-5) {
        if (c > 0) {
            r[0] = r[1] = r[2] = 0;
        } else {
            double t;
            t = (R[0] + 1) * 0.5; rx = std::sqrt(std::max(t, 0.));
            t = (R[4] + 1) * 0.5; ry = std::sqrt(std::max(t, 0.)) * (R[1] < 0 ? -1. : 1.);
            t = (R[8] + 1) * 0.5; rz = std::sqrt(std::max(t, 0.)) * (R[2] < 0 ? -1. : 1.);
            if (std::fabs(rx) < std::fabs(ry) && std::fabs(rx) < std::fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
            theta /= std::sqrt(rx * rx + ry * ry + rz * rz);
            r[0] = rx * theta; r[1] = ry * theta; r[2] = rz * theta;
        }
    } else {
        double vth = 1 / (2 * s);
        vth *= theta;
        r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
    }
}

// ------------------------------------------------------------------ projection
// cv::projectPoints with zero distortion (cnn_softam.h:351, :1046): all arithmetic in
// double; Jacobians as cvProjectPoints2 produces them for CV_ITERATIVE.
void project_points(int n, const double* X, const double rvec[3], const double tvec[3], double f, double cx, double cy,
                    double* uv, double* dpdr, double* dpdt) {
    double R[9], dRdr[27];
    rodrigues_v2m(rvec, R, dpdr ? dRdr : nullptr);
    for (int i = 0; i < n; i++) {
        double Xw = X[i * 3], Yw = X[i * 3 + 1], Zw = X[i * 3 + 2];
        double x = R[0] * Xw + R[1] * Yw + R[2] * Zw + tvec[0];
        double y = R[3] * Xw + R[4] * Yw + R[5] * Zw + tvec[1];
        double z = R[6] * Xw + R[7] * Yw + R[8] * Zw + tvec[2];
        z = z ? 1. / z : 1;
        x *= z;
        y *= z;
        uv[i * 2] = x * f + cx;
        uv[i * 2 + 1] = y * f + cy;
        if (dpdt) {
            double* a = dpdt + (2 * i) * 3;
            a[0] = f * z; a[1] = 0; a[2] = f * (-x * z);
            a[3] = 0; a[4] = f * z; a[5] = f * (-y * z);
        }
        if (dpdr) {
            double* a = dpdr + (2 * i) * 3;
            for (int j = 0; j < 3; j++) {
                double dx0 = Xw * dRdr[j * 9 + 0] + Yw * dRdr[j * 9 + 1] + Zw * dRdr[j * 9 + 2];
                double dy0 = Xw * dRdr[j * 9 + 3] + Yw * dRdr[j * 9 + 4] + Zw * dRdr[j * 9 + 5];
                double dz0 = Xw * dRdr[j * 9 + 6] + Yw * dRdr[j * 9 + 7] + Zw * dRdr[j * 9 + 8];
                double dxdr = z * (dx0 - x * dz0), dydr = z * (dy0 - y * dz0);
                a[j] = f * dxdr;
                a[3 + j] = f * dydr;
            }
        }
    }
}

// ------------------------------------------------------------------ P3P (Gao et al.)
// Real roots of a x^4 + b x^3 + c x^2 + d x + e (resolvent-cubic / Ferrari scheme as in
// MathWorld "Quartic Equation", which is what OpenCV's p3p uses).
int solve_cubic_one(double a2, double a1, double a0, double roots[3]) {
    // x^3 + a2 x^2 + a1 x + a0 = 0 ; roots[0] is the one Ferrari's step uses
    double Q = (3 * a1 - a2 * a2) / 9;
    double R = (9 * a2 * a1 - 27 * a0 - 2 * a2 * a2 * a2) / 54;
    double Q3 = Q * Q * Q, D = Q3 + R * R, sh = a2 / 3;
    if (Q == 0) {
        if (R == 0) { roots[0] = roots[1] = roots[2] = -sh; return 3; }
        roots[0] = std::cbrt(2 * R) - sh;
        return 1;
    }
    if (D <= 0) {
        double theta = std::acos(R / std::sqrt(-Q3)), sq = std::sqrt(-Q);
        roots[0] = 2 * sq * std::cos(theta / 3.0) - sh;
        roots[1] = 2 * sq * std::cos((theta + 2 * kPi) / 3.0) - sh;
        roots[2] = 2 * sq * std::cos((theta + 4 * kPi) / 3.0) - sh;
        return 3;
    }
    double AD = std::cbrt(std::fabs(R) + std::sqrt(D)) * (R > 0 ? 1 : (R < 0 ? -1 : 0));
    double BD = (AD == 0) ? 0 : -Q / AD;
    roots[0] = AD + BD - sh;
    return 1;
}

int solve_quartic(double a, double b, double c, double d, double e, double x[4]) {
    if (a == 0) return 0;
    double ia = 1.0 / a;
    b *= ia; c *= ia; d *= ia; e *= ia;
    double cub[3];
    if (solve_cubic_one(-c, d * b - 4 * e, 4 * c * e - d * d - b * b * e, cub) == 0) return 0;
    double y1 = cub[0];
    double R2 = 0.25 * b * b - c + y1;
    if (R2 < 0) return 0;
    double R = std::sqrt(R2), D2, E2;
    if (R < 10e-12) {
        double t = y1 * y1 - 4 * e;
        if (t < 0) {
            D2 = E2 = -1;
        } else {
            double st = std::sqrt(t);
            D2 = 0.75 * b * b - 2 * c + 2 * st;
            E2 = D2 - 4 * st;
        }
    } else {
        double u = 0.75 * b * b - 2 * c - R2;
        double v = 0.25 * (4 * b * c - 8 * d - b * b * b) / R;
        D2 = u + v;
        E2 = u - v;
    }
    int n = 0;
    if (D2 >= 0) {
        double Dq = std::sqrt(D2);
        x[n++] = 0.5 * R + 0.5 * Dq - 0.25 * b;
        x[n] = x[n - 1] - Dq;
        n++;
    }
    if (E2 >= 0) {
        double Eq = std::sqrt(E2);
        x[n++] = -0.5 * R + 0.5 * Eq - 0.25 * b;
        x[n] = x[n - 1] - Eq;
        n++;
    }
    return n;
}

// Horn's closed-form absolute orientation (unit quaternion = dominant eigenvector of a
// 4x4 symmetric matrix) between the 3 world points and their camera-frame positions.
bool align3(const double Mc[3][3], const double Xw[3][3], double R[9], double t[3]) {
    double Cc[3], Cw[3];
    for (int j = 0; j < 3; j++) {
        Cc[j] = (Mc[0][j] + Mc[1][j] + Mc[2][j]) / 3;
        Cw[j] = (Xw[0][j] + Xw[1][j] + Xw[2][j]) / 3;
    }
    double s[9];
    for (int i = 0; i < 3; i++)      // world axis
        for (int j = 0; j < 3; j++)  // camera axis
            s[i * 3 + j] = (Xw[0][i] * Mc[0][j] + Xw[1][i] * Mc[1][j] + Xw[2][i] * Mc[2][j]) / 3 - Cc[j] * Cw[i];
    double Q[16];
    Q[0] = s[0] + s[4] + s[8];
    Q[5] = s[0] - s[4] - s[8];
    Q[10] = s[4] - s[8] - s[0];
    Q[15] = s[8] - s[0] - s[4];
    Q[4] = Q[1] = s[5] - s[7];
    Q[8] = Q[2] = s[6] - s[2];
    Q[12] = Q[3] = s[1] - s[3];
    Q[9] = Q[6] = s[3] + s[1];
    Q[13] = Q[7] = s[6] + s[2];
    Q[14] = Q[11] = s[7] + s[5];
    double ev[4], U[16];
    jacobi_eig(4, Q, ev, U);
    int im = 0;
    for (int i = 1; i < 4; i++)
        if (ev[i] > ev[im]) im = i;
    double q0 = U[0 * 4 + im], q1 = U[1 * 4 + im], q2 = U[2 * 4 + im], q3 = U[3 * 4 + im];
    R[0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3;
    R[1] = 2. * (q1 * q2 - q0 * q3);
    R[2] = 2. * (q1 * q3 + q0 * q2);
    R[3] = 2. * (q1 * q2 + q0 * q3);
    R[4] = q0 * q0 + q2 * q2 - q1 * q1 - q3 * q3;
    R[5] = 2. * (q2 * q3 - q0 * q1);
    R[6] = 2. * (q1 * q3 - q0 * q2);
    R[7] = 2. * (q2 * q3 + q0 * q1);
    R[8] = q0 * q0 + q3 * q3 - q1 * q1 - q2 * q2;
    for (int i = 0; i < 3; i++) t[i] = Cc[i] - (R[i * 3] * Cw[0] + R[i * 3 + 1] * Cw[1] + R[i * 3 + 2] * Cw[2]);
    return true;
}

// All P3P solutions for 3 correspondences (mu,mv in pixels).  Unknown ratios x = |PA|/|PC|,
// y = |PB|/|PC| satisfy the two quadrics obtained from the three law-of-cosines
// equations; eliminating y^2 gives y as a rational function of x and a quartic in x.
int p3p_lengths(const double dist[3] /* |BC|,|AC|,|AB| */, const double cosv[3] /* BPC, APC, APB */, double L[4][3]) {
    double p = 2 * cosv[0], q = 2 * cosv[1], r = 2 * cosv[2];
    double inv_c2 = 1.0 / (dist[2] * dist[2]);
    double a = inv_c2 * dist[0] * dist[0], b = inv_c2 * dist[1] * dist[1];
    if (p * p + q * q + r * r - p * q * r - 1 == 0) return 0;  // projection centre coplanar with A,B,C
    // y * Dn(x) = -Nn(x):   Nn = (1-a-b) x^2 + q(a-1) x + (1-a+b),  Dn = b (r x - p)
    double N2 = 1 - a - b, N1 = q * (a - 1), N0 = 1 - a + b;
    double D1 = b * r, D0 = -b * p;
    // quartic: ((1-b)x^2 - q x + 1) Dn^2 - b Nn^2 - b r x Nn Dn = 0
    double F2 = 1 - b, F1 = -q, F0 = 1;
    double DD2 = D1 * D1, DD1 = 2 * D1 * D0, DD0 = D0 * D0;
    double NN4 = N2 * N2, NN3 = 2 * N2 * N1, NN2 = 2 * N2 * N0 + N1 * N1, NN1 = 2 * N1 * N0, NN0 = N0 * N0;
    double ND3 = N2 * D1, ND2 = N2 * D0 + N1 * D1, ND1 = N1 * D0 + N0 * D1, ND0 = N0 * D0;
    double c4 = F2 * DD2 - b * NN4 - b * r * ND3;
    double c3 = F2 * DD1 + F1 * DD2 - b * NN3 - b * r * ND2;
    double c2 = F2 * DD0 + F1 * DD1 + F0 * DD2 - b * NN2 - b * r * ND1;
    double c1 = F1 * DD0 + F0 * DD1 - b * NN1 - b * r * ND0;
    double c0 = F0 * DD0 - b * NN0;
    if (c4 == 0) return 0;
    double xr[4];
    int n = solve_quartic(c4, c3, c2, c1, c0, xr);
    // The closed-form roots lose accuracy when two roots nearly coincide, and y = -Nn/Dn is
    // ill-conditioned where Dn ~ 0 (there the two quadrics become proportional in y and
    // BOTH roots of the first quadric are solutions).  Every (x, y) candidate is therefore
    // polished by Newton's method on the two quadrics
    //   f1 = (1-a) y^2 - a x^2 - p y + a r x y + 1,  f2 = (1-b) x^2 - b y^2 - q x + b r x y + 1
    // and kept only if it converges to a positive, not-yet-found solution.
    int ns = 0;
    double sx[4], sy[4];
    for (int i = 0; i < n; i++) {
        double x0 = xr[i];
        if (!(x0 == x0)) continue;
        double Dn = D1 * x0 + D0;
        double ycand[2];
        int nc = 0;
        if (std::fabs(Dn) > 1e-3 * (std::fabs(D1 * x0) + std::fabs(D0))) {
            ycand[nc++] = -((N2 * x0 + N1) * x0 + N0) / Dn;
        } else {
            double qa = 1 - a, qb = a * r * x0 - p, qc = 1 - a * x0 * x0;
            double disc = qb * qb - 4 * qa * qc;
            if (disc < 0) disc = 0;
            double sq = std::sqrt(disc);
            if (qa != 0) {
                ycand[nc++] = (-qb + sq) / (2 * qa);
                ycand[nc++] = (-qb - sq) / (2 * qa);
            } else if (qb != 0) {
                ycand[nc++] = -qc / qb;
            }
        }
        for (int c = 0; c < nc; c++) {
            double x = x0, y = ycand[c];
            bool good = false;
            for (int it = 0; it < 8; it++) {
                double f1 = (1 - a) * y * y - a * x * x - p * y + a * r * x * y + 1;
                double f2 = (1 - b) * x * x - b * y * y - q * x + b * r * x * y + 1;
                double j11 = -2 * a * x + a * r * y, j12 = 2 * (1 - a) * y - p + a * r * x;
                double j21 = 2 * (1 - b) * x - q + b * r * y, j22 = -2 * b * y + b * r * x;
                double det = j11 * j22 - j12 * j21;
                if (det == 0 || !(det == det)) break;
                double dx = (f1 * j22 - f2 * j12) / det, dy = (j11 * f2 - j21 * f1) / det;
                x -= dx;
                y -= dy;
                if (std::fabs(dx) + std::fabs(dy) <= 1e-15 * (std::fabs(x) + std::fabs(y))) {
                    good = true;
                    break;
                }
            }
            if (!good) {  // accept a slowly converging candidate only if its residual is tiny
                double f1 = (1 - a) * y * y - a * x * x - p * y + a * r * x * y + 1;
                double f2 = (1 - b) * x * x - b * y * y - q * x + b * r * x * y + 1;
                good = std::fabs(f1) + std::fabs(f2) < 1e-12 * (1 + x * x + y * y);
            }
            if (!good || !(x > 0) || !(y > 0)) continue;
            bool dup = false;
            for (int k = 0; k < ns; k++)
                if (std::fabs(sx[k] - x) + std::fabs(sy[k] - y) < 1e-9 * (std::fabs(x) + std::fabs(y))) dup = true;
            if (dup || ns >= 4) continue;
            sx[ns] = x;
            sy[ns] = y;
            ns++;
        }
    }
    int no = 0;
    for (int k = 0; k < ns; k++) {
        double x = sx[k], y = sy[k];
        double v = x * x + y * y - x * y * r;
        if (!(v > 0)) continue;
        double Z = dist[2] / std::sqrt(v);
        L[no][0] = x * Z;
        L[no][1] = y * Z;
        L[no][2] = Z;
        no++;
    }
    return no;
}

int p3p_solutions(const double mu_in[3], const double mv_in[3], const double Xw[3][3], double f, double cx, double cy,
                  double Rs[4][9], double ts[4][3]) {
    double inv_f = 1.0 / f, cx_f = cx / f, cy_f = cy / f;
    double mu[3], mv[3], mk[3];
    for (int i = 0; i < 3; i++) {
        mu[i] = inv_f * mu_in[i] - cx_f;
        mv[i] = inv_f * mv_in[i] - cy_f;
        double nrm = std::sqrt(mu[i] * mu[i] + mv[i] * mv[i] + 1);
        mk[i] = 1. / nrm;
        mu[i] *= mk[i];
        mv[i] *= mk[i];
    }
    auto d3 = [&](int i, int j) {
        double dx = Xw[i][0] - Xw[j][0], dy = Xw[i][1] - Xw[j][1], dz = Xw[i][2] - Xw[j][2];
        return std::sqrt(dx * dx + dy * dy + dz * dz);
    };
    double dist[3] = {d3(1, 2), d3(0, 2), d3(0, 1)};
    double cosv[3] = {mu[1] * mu[2] + mv[1] * mv[2] + mk[1] * mk[2], mu[0] * mu[2] + mv[0] * mv[2] + mk[0] * mk[2],
                      mu[0] * mu[1] + mv[0] * mv[1] + mk[0] * mk[1]};
    double L[4][3];
    int n = p3p_lengths(dist, cosv, L);
    int ns = 0;
    for (int i = 0; i < n; i++) {
        double Mc[3][3];
        for (int k = 0; k < 3; k++) {
            Mc[k][0] = L[i][k] * mu[k];
            Mc[k][1] = L[i][k] * mv[k];
            Mc[k][2] = L[i][k] * mk[k];
        }
        if (!align3(Mc, Xw, Rs[ns], ts[ns])) continue;
        ns++;
    }
    return ns;
}

// cv::solvePnP(..., CV_P3P) as called by safeSolvePnP (cnn_softam.h:56-73, :1042):
// 4 correspondences; undistortPoints rounds the normalised pixel to float, p3p then maps
// it back through K; the first three points give up to 4 poses, the 4th picks one.
bool solve_p3p(const float obj[12], const float img[8], double f, double cx, double cy, double rvec[3], double tvec[3]) {
    double mu[4], mv[4], Xw[4][3];
    double ifx = 1. / f;
    for (int i = 0; i < 4; i++) {
        float xn = (float)(((double)img[i * 2] - cx) * ifx);
        float yn = (float)(((double)img[i * 2 + 1] - cy) * ifx);
        mu[i] = xn * f + cx;
        mv[i] = yn * f + cy;
        for (int k = 0; k < 3; k++) Xw[i][k] = obj[i * 3 + k];
    }
    double Rs[4][9], ts[4][3];
    int n = p3p_solutions(mu, mv, Xw, f, cx, cy, Rs, ts);
    if (n == 0) return false;
    int best = 0;
    double min_reproj = 0;
    for (int i = 0; i < n; i++) {
        double Xc[3];
        mat3_vec(Rs[i], Xw[3], Xc);
        Xc[0] += ts[i][0]; Xc[1] += ts[i][1]; Xc[2] += ts[i][2];
        double u = cx + f * Xc[0] / Xc[2], v = cy + f * Xc[1] / Xc[2];
        double reproj = (u - mu[3]) * (u - mu[3]) + (v - mv[3]) * (v - mv[3]);
        if (i == 0 || min_reproj > reproj) {
            best = i;
            min_reproj = reproj;
        }
    }
    rodrigues_m2v(Rs[best], rvec);
    tvec[0] = ts[best][0]; tvec[1] = ts[best][1]; tvec[2] = ts[best][2];
    return true;
}

// ------------------------------------------------------------------ iterative PnP (LM)
// cv::solvePnP(..., useExtrinsicGuess=true, CV_ITERATIVE) (cnn_softam.h:708, :1144):
// cvFindExtrinsicCameraParams2 skips initialisation and runs CvLevMarq(6 params,
// max_iter 20, eps FLT_EPSILON) on the reprojection residuals.
void lm_step(const double JtJ[36], const double JtErr[6], const double prev[6], int lambdaLg10, double param[6]) {
    double lambda = std::exp(lambdaLg10 * std::log(10.));
    double A[36];
    for (int i = 0; i < 36; i++) A[i] = JtJ[i];
    for (int i = 0; i < 6; i++) A[i * 7] *= 1. + lambda;
    double U[36], w[6], V[36];
    jacobi_svd(6, A, U, w, V);
    double thr = 0;
    for (int i = 0; i < 6; i++) thr += w[i];
    thr *= std::numeric_limits<double>::epsilon() * 2;
    double x[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 6; k++) {
        if (!(w[k] > thr)) continue;
        double d = 0;
        for (int i = 0; i < 6; i++) d += U[i * 6 + k] * JtErr[i];
        d /= w[k];
        for (int i = 0; i < 6; i++) x[i] += d * V[i * 6 + k];
    }
    for (int i = 0; i < 6; i++) param[i] = prev[i] - x[i];
}

int solve_pnp_iterative(int n, const float* obj, const float* img, double f, double cx, double cy, double rvec[3],
                        double tvec[3]) {
    std::vector<double> M(n * 3), m(n * 2), proj(n * 2), dpdr(n * 6), dpdt(n * 6), err(n * 2);
    for (int i = 0; i < n * 3; i++) M[i] = obj[i];
    for (int i = 0; i < n * 2; i++) m[i] = img[i];
    double param[6] = {rvec[0], rvec[1], rvec[2], tvec[0], tvec[1], tvec[2]}, prev[6];
    double JtJ[36], JtErr[6];
    int lambdaLg10 = -3, iters = 0;
    double prevErrNorm = std::numeric_limits<double>::max(), errNorm = 0;
    const int max_iter = 20;
    const double eps = std::numeric_limits<float>::epsilon();
    auto residual = [&](bool jac) {
        project_points(n, M.data(), param, param + 3, f, cx, cy, proj.data(), jac ? dpdr.data() : nullptr,
                       jac ? dpdt.data() : nullptr);
        for (int i = 0; i < n * 2; i++) err[i] = proj[i] - m[i];
    };
    auto norm_err = [&]() {
        double s = 0;
        for (int i = 0; i < n * 2; i++) s += err[i] * err[i];
        return std::sqrt(s);
    };
    // state STARTED -> CALC_J
    residual(true);
    for (;;) {
        // CALC_J: normal equations at the current parameters
        for (int a = 0; a < 6; a++) {
            for (int b = a; b < 6; b++) {
                double s = 0;
                for (int i = 0; i < n * 2; i++) {
                    double ja = a < 3 ? dpdr[i * 3 + a] : dpdt[i * 3 + a - 3];
                    double jb = b < 3 ? dpdr[i * 3 + b] : dpdt[i * 3 + b - 3];
                    s += ja * jb;
                }
                JtJ[a * 6 + b] = JtJ[b * 6 + a] = s;
            }
            double s = 0;
            for (int i = 0; i < n * 2; i++) s += (a < 3 ? dpdr[i * 3 + a] : dpdt[i * 3 + a - 3]) * err[i];
            JtErr[a] = s;
        }
        std::memcpy(prev, param, sizeof(prev));
        lm_step(JtJ, JtErr, prev, lambdaLg10, param);
        if (iters == 0) prevErrNorm = norm_err();
        // CHECK_ERR loop
        bool done = false;
        for (;;) {
            residual(false);
            errNorm = norm_err();
            if (errNorm > prevErrNorm) {
                if (++lambdaLg10 <= 16) {
                    lm_step(JtJ, JtErr, prev, lambdaLg10, param);
                    continue;
                }
            }
            lambdaLg10 = std::max(lambdaLg10 - 1, -16);
            double dn = 0, pn = 0;
            for (int i = 0; i < 6; i++) {
                dn += (param[i] - prev[i]) * (param[i] - prev[i]);
                pn += prev[i] * prev[i];
            }
            double change = std::sqrt(dn) / std::sqrt(pn);
            if (++iters >= max_iter || change < eps) done = true;
            break;
        }
        if (done) break;
        prevErrNorm = errNorm;
        residual(true);
    }
    for (int i = 0; i < 3; i++) {
        rvec[i] = param[i];
        tvec[i] = param[3 + i];
    }
    return iters;
}

// ------------------------------------------------------------------ conventions
// jp::cv2our, types.h:186-214
void cv2our(const double rvec[3], const double tvec[3], double R[9], double t[3]) {
    rodrigues_v2m(rvec, R, nullptr);
    t[0] = tvec[0]; t[1] = tvec[1]; t[2] = tvec[2];
    for (int j = 0; j < 3; j++) {
        R[3 + j] = -R[3 + j];
        R[6 + j] = -R[6 + j];
    }
    t[1] = -t[1];
    t[2] = -t[2];
    if (det3(R) < 0) {
        for (int k = 0; k < 9; k++) R[k] = -R[k];
        for (int k = 0; k < 3; k++) t[k] = -t[k];
    }
    if (t[0] != t[0] || t[1] != t[1] || t[2] != t[2]) t[0] = t[1] = t[2] = 0;
}

// jp::our2cv, types.h:137-151
void our2cv(const double R[9], const double t[3], double rvec[3], double tvec[3]) {
    double Rm[9];
    std::memcpy(Rm, R, sizeof(Rm));
    for (int j = 0; j < 3; j++) {
        Rm[3 + j] = -Rm[3 + j];
        Rm[6 + j] = -Rm[6 + j];
    }
    rodrigues_m2v(Rm, rvec);
    tvec[0] = t[0]; tvec[1] = -t[1]; tvec[2] = -t[2];
}

// Hypothesis(jp).getRodVecAndTrans(), Hypothesis.cpp:274-290
void jp6_from_cv(const double rvec[3], const double tvec[3], double out6[6]) {
    double R[9], t[3];
    cv2our(rvec, tvec, R, t);
    rodrigues_m2v(R, out6);
    out6[3] = t[0]; out6[4] = t[1]; out6[5] = t[2];
}

// getInvHyp, maxloss.h:39-61 (inverse of the rigid 4x4)
void inv_pose(const double R[9], const double t[3], double Ri[9], double ti[3]) {
    // general 3x3 inverse (the reference inverts the 4x4 numerically)
    double d = det3(R), id = 1.0 / d;
    Ri[0] = (R[4] * R[8] - R[5] * R[7]) * id; Ri[1] = (R[2] * R[7] - R[1] * R[8]) * id; Ri[2] = (R[1] * R[5] - R[2] * R[4]) * id;
    Ri[3] = (R[5] * R[6] - R[3] * R[8]) * id; Ri[4] = (R[0] * R[8] - R[2] * R[6]) * id; Ri[5] = (R[2] * R[3] - R[0] * R[5]) * id;
    Ri[6] = (R[3] * R[7] - R[4] * R[6]) * id; Ri[7] = (R[1] * R[6] - R[0] * R[7]) * id; Ri[8] = (R[0] * R[4] - R[1] * R[3]) * id;
    double m[3];
    mat3_vec(Ri, t, m);
    ti[0] = -m[0]; ti[1] = -m[1]; ti[2] = -m[2];
}

// Hypothesis::calcAngularDistance, Hypothesis.cpp:137-143 (uses the other pose's inverse rotation)
double angular_distance(const double R1[9], const double R2inv[9]) {
    double D[9];
    mat3_mul(R1, R2inv, D);
    double tr = D[0] + D[4] + D[8];
    tr = std::min(3.0, std::max(-1.0, tr));
    return 180 * std::acos((tr - 1.0) / 2.0) / kPi;
}

void inv3(const double R[9], double Ri[9]) {
    double t0[3] = {0, 0, 0}, ti[3];
    inv_pose(R, t0, Ri, ti);
}

// maxLoss, maxloss.h:69-79 ; also returns the two errors as cnn_softam.h:1166-1170 measures them
double max_loss(const double R1[9], const double t1[3], const double R2[9], const double t2[3], double* rot_err, double* t_err) {
    double Ri1[9], ti1[3], Ri2[9], ti2[3];
    inv_pose(R1, t1, Ri1, ti1);
    inv_pose(R2, t2, Ri2, ti2);
    double Ri2inv[9];
    inv3(Ri2, Ri2inv);
    double re = angular_distance(Ri1, Ri2inv);
    double dx = ti1[0] - ti2[0], dy = ti1[1] - ti2[1], dz = ti1[2] - ti2[2];
    double te = std::sqrt(dx * dx + dy * dy + dz * dz);
    if (rot_err) *rot_err = re;
    if (t_err) *t_err = te;
    return std::min(std::max(re, te / 10), 10000000.0);
}

}  // namespace

// ====================================================================== exported: primitives
extern "C" {

void orc_default_config(orc_config* c) {
    c->f = 525.0;                   // properties.cpp:55
    c->cx = 320.0; c->cy = 240.0;   // properties.cpp:310-311 with iw 640, ih 480, xs ys 0
    c->n_hyps = 256;                // properties.cpp:45
    c->thr2d = 10;                  // properties.cpp:50, truncated at test_ransac_softam.cpp:51
    c->inlier_count = 100;          // properties.cpp:47
    c->ref_steps = 8;               // properties.cpp:46
    c->sub_sample = 0.01;           // properties.cpp:48
    c->alpha = 0.1; c->beta = 0.5;  // engine defaults, not reference values
    c->seed = 1305;                 // thread_rand.h:100
    c->n_streams = 1;
    c->stream_skip = 6400;
    c->max_candidates = 0;
    c->fix_q4 = 0;
    c->grad_clamp = 0.1;            // clampE2E, train_score_softam.lua:13
}

void orc_rodrigues(const double r[3], double R[9], double J[27]) { rodrigues_v2m(r, R, J); }
void orc_rodrigues_inv(const double R[9], double r[3]) { rodrigues_m2v(R, r); }
void orc_project_points(int n, const double* X, const double rvec[3], const double tvec[3], double f, double cx, double cy,
                        double* uv, double* dpdr, double* dpdt) {
    project_points(n, X, rvec, tvec, f, cx, cy, uv, dpdr, dpdt);
}
int orc_solve_p3p(const float obj[12], const float img[8], double f, double cx, double cy, double rvec[3], double tvec[3]) {
    if (!solve_p3p(obj, img, f, cx, cy, rvec, tvec)) {
        for (int i = 0; i < 3; i++) rvec[i] = tvec[i] = 0;  // safeSolvePnP, cnn_softam.h:66-71
        return 0;
    }
    return 1;
}
int orc_p3p_all(const float obj[9], const float img[6], double f, double cx, double cy, double Rs[36], double ts[12]) {
    double mu[3], mv[3], Xw[3][3], R4[4][9], t4[4][3];
    for (int i = 0; i < 3; i++) {
        mu[i] = img[i * 2];
        mv[i] = img[i * 2 + 1];
        for (int k = 0; k < 3; k++) Xw[i][k] = obj[i * 3 + k];
    }
    int n = p3p_solutions(mu, mv, Xw, f, cx, cy, R4, t4);
    for (int i = 0; i < n; i++) {
        std::memcpy(Rs + i * 9, R4[i], 9 * sizeof(double));
        std::memcpy(ts + i * 3, t4[i], 3 * sizeof(double));
    }
    return n;
}
int orc_solve_pnp_iterative(int n, const float* obj, const float* img, double f, double cx, double cy, double rvec[3],
                            double tvec[3], int* iters_out) {
    int it = solve_pnp_iterative(n, obj, img, f, cx, cy, rvec, tvec);
    if (iters_out) *iters_out = it;
    return 1;
}
void orc_svd3(const double A[9], double U[9], double w[3], double Vt[9]) {
    double V[9];
    jacobi_svd(3, A, U, w, V);
    mat3_t(V, Vt);
}

// Hypothesis::calcRigidBodyTransform, Hypothesis.cpp:145-200 (Kabsch: b ~ R a + t)
void orc_kabsch(int n, const double* a, const double* b, double R[9], double t[3]) {
    double cA[3] = {0, 0, 0}, cB[3] = {0, 0, 0};
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            cA[k] += a[i * 3 + k];
            cB[k] += b[i * 3 + k];
        }
    for (int k = 0; k < 3; k++) {
        cA[k] *= 1.0 / (double)n;
        cB[k] *= 1.0 / (double)n;
    }
    double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // pointsA * pointsB^T
    for (int i = 0; i < n; i++)
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) C[r * 3 + c] += (a[i * 3 + r] - cA[r]) * (b[i * 3 + c] - cB[c]);
    double U[9], w[3], V[9], Ut[9], VUt[9];
    jacobi_svd(3, C, U, w, V);
    mat3_t(U, Ut);
    mat3_mul(V, Ut, VUt);
    double sign = det3(VUt) < 0 ? -1 : 1;
    double Dm[9] = {1, 0, 0, 0, 1, 0, 0, 0, sign}, VD[9];
    mat3_mul(V, Dm, VD);
    mat3_mul(VD, Ut, R);
    double Ra[3];
    mat3_vec(R, cA, Ra);
    for (int k = 0; k < 3; k++) t[k] = -Ra[k] + cB[k];
}

// ====================================================================== RNG contract
// ThreadRand with one OMP thread: generators[0].seed(seed) (thread_rand.cpp:40-57);
// drand = fresh uniform_real_distribution<double> per call (thread_rand.cpp:71-81);
// irand(a,b) = fresh uniform_int_distribution<int>(a, b-1) (thread_rand.cpp:59-69,95-98).
// stochasticSubSample, cnn_softam.h:283-309 (targetSize 40, patchSize 42).
void orc_stochastic_subsample(uint32_t seed, int width, int height, int32_t* pix) {
    std::mt19937 gen;
    gen.seed(seed);
    const int targetSize = ORC_GRID, patchSize = 42;
    float xStride = (width - patchSize) / (float)targetSize;
    float yStride = (height - patchSize) / (float)targetSize;
    int sampleX = 0;
    for (float minX = patchSize / 2, x = xStride + patchSize / 2; x <= width - patchSize / 2 + 1; minX = x, x += xStride) {
        int sampleY = 0;
        for (float minY = patchSize / 2, y = yStride + patchSize / 2; y <= height - patchSize / 2 + 1; minY = y, y += yStride) {
            std::uniform_real_distribution<double> dx(minX, x);
            int curX = dx(gen);
            std::uniform_real_distribution<double> dy(minY, y);
            int curY = dy(gen);
            if (sampleX < targetSize && sampleY < targetSize) {
                pix[(sampleY * targetSize + sampleX) * 2] = curX;
                pix[(sampleY * targetSize + sampleX) * 2 + 1] = curY;
            }
            sampleY++;
        }
        sampleX++;
    }
}

void orc_mt19937_raw(uint32_t seed, int n, uint32_t* out) {
    std::mt19937 gen;
    gen.seed(seed);
    for (int i = 0; i < n; i++) out[i] = (uint32_t)gen();
}

namespace {
// One minimal-set candidate drawn exactly as cnn_softam.h:1015-1039 does.
inline unsigned draw_candidate(std::mt19937& gen, int32_t cells[8], unsigned char* chosen /* N zeros */) {
    unsigned draws = 0;
    int xs[4], ys[4];
    for (int j = 0; j < 4; j++) {
        std::uniform_int_distribution<int> dxx(0, ORC_GRID - 1);
        int x = dxx(gen);
        std::uniform_int_distribution<int> dyy(0, ORC_GRID - 1);
        int y = dyy(gen);
        draws += 2;
        if (chosen[y * ORC_GRID + x] > 0) {
            j--;
            continue;
        }
        chosen[y * ORC_GRID + x] = 1;
        xs[j] = x;
        ys[j] = y;
    }
    for (int j = 0; j < 4; j++) {
        chosen[ys[j] * ORC_GRID + xs[j]] = 0;
        cells[j * 2] = xs[j];
        cells[j * 2 + 1] = ys[j];
    }
    return draws;
}
}  // namespace

int orc_candidates(uint32_t seed, uint32_t skip, int n_cand, int32_t* cells, uint32_t* draws_used) {
    std::mt19937 gen;
    gen.seed(seed);
    gen.discard(skip);
    std::vector<unsigned char> chosen(ORC_N, 0);
    for (int k = 0; k < n_cand; k++) {
        unsigned d = draw_candidate(gen, cells + k * 8, chosen.data());
        if (draws_used) draws_used[k] = d;  // counts irand calls (Lemire rejections are extra raw words)
    }
    return n_cand;
}

// std::mt19937 randG default-seeded once per frame; iota + std::shuffle per refinement
// step with the continuing generator, cnn_softam.h:1104,1112-1114.
void orc_refine_permutations(int steps, int32_t* perm) {
    std::mt19937 randG;
    for (int s = 0; s < steps; s++) {
        std::vector<int> idx(ORC_N);
        for (int i = 0; i < ORC_N; i++) idx[i] = i;
        std::shuffle(idx.begin(), idx.end(), randG);
        for (int i = 0; i < ORC_N; i++) perm[s * ORC_N + i] = idx[i];
    }
}

// ====================================================================== pipeline stages
// getDiffMap, cnn_softam.h:319-362: projection in double, result rounded to float
// (Point2f), float subtraction, norm in double, min with 100.0, stored as float.
void orc_diff_map(const int16_t* coords, const int32_t* pix, const double rvec[3], const double tvec[3], double f, double cx,
                  double cy, float* diff) {
    double R[9];
    rodrigues_v2m(rvec, R, nullptr);
    for (int i = 0; i < ORC_N; i++) {
        double Xw = (float)coords[i * 3], Yw = (float)coords[i * 3 + 1], Zw = (float)coords[i * 3 + 2];
        double x = R[0] * Xw + R[1] * Yw + R[2] * Zw + tvec[0];
        double y = R[3] * Xw + R[4] * Yw + R[5] * Zw + tvec[1];
        double z = R[6] * Xw + R[7] * Yw + R[8] * Zw + tvec[2];
        z = z ? 1. / z : 1;
        x *= z;
        y *= z;
        float pu = (float)(x * f + cx), pv = (float)(y * f + cy);
        float du = (float)pix[i * 2] - pu, dv = (float)pix[i * 2 + 1] - pv;
        double nrm = std::sqrt((double)du * du + (double)dv * dv);
        diff[i] = (float)std::min(nrm, ORC_MAXINPUT);
    }
}

// Closed-form soft-inlier score (north_star): replaces forward(diffMaps, stateObj),
// cnn_softam.h:1072 -> lua_calls.h:284-300.
double orc_soft_inlier_score(const float* diff, int n, double tau, double alpha, double beta) {
    double s = 0;
    for (int i = 0; i < n; i++) s += 1.0 / (1.0 + std::exp(-beta * (tau - (double)diff[i])));
    return alpha * s;
}

// softMax, cnn_softam.h:535-553
void orc_softmax(const double* s, int n, double* p) {
    double mx = 0;
    for (int i = 0; i < n; i++)
        if (i == 0 || s[i] > mx) mx = s[i];
    double sum = 0;
    for (int i = 0; i < n; i++) {
        p[i] = std::exp(s[i] - mx);
        sum += p[i];
    }
    for (int i = 0; i < n; i++) p[i] /= sum;
}

// entropy, cnn_softam.h:80-88
double orc_entropy(const double* p, int n) {
    double e = 0;
    for (int i = 0; i < n; i++)
        if (p[i] > 0) e -= p[i] * std::log2(p[i]);
    return e;
}

void orc_cv2our(const double rvec[3], const double tvec[3], double R[9], double t[3]) { cv2our(rvec, tvec, R, t); }
void orc_our2cv(const double R[9], const double t[3], double rvec[3], double tvec[3]) { our2cv(R, t, rvec, tvec); }
void orc_jp6(const double rvec[3], const double tvec[3], double out6[6]) { jp6_from_cv(rvec, tvec, out6); }
double orc_max_loss(const double R1[9], const double t1[3], const double R2[9], const double t2[3], double* rot_err,
                    double* t_err) {
    return max_loss(R1, t1, R2, t2, rot_err, t_err);
}

// dLossMax, maxloss.h:87-198
void orc_dloss_max(const double est[6], const double gt[6], double jac[6]) {
    for (int i = 0; i < 6; i++) jac[i] = 0;
    double rot1[9], rot2[9], dRod[27];
    rodrigues_v2m(est, rot1, dRod);
    rodrigues_v2m(gt, rot2, nullptr);
    double invRot1[9], invRot2[9], diffRot[9];
    mat3_t(rot1, invRot1);
    mat3_t(rot2, invRot2);
    mat3_mul(rot1, invRot2, diffRot);
    double trace = diffRot[0] + diffRot[4] + diffRot[8];
    trace = std::min(3.0, std::max(-1.0, trace));
    double rotErr = 180 * std::acos((trace - 1.0) / 2.0) / kPi;
    double a1[3] = {-est[3] / 10, -est[4] / 10, -est[5] / 10}, a2[3] = {-gt[3] / 10, -gt[4] / 10, -gt[5] / 10};
    double invT1[3], invT2[3];
    mat3_vec(invRot1, a1, invT1);
    mat3_vec(invRot2, a2, invT2);
    double d[3] = {invT1[0] - invT2[0], invT1[1] - invT2[1], invT1[2] - invT2[2]};
    double tErr = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (std::max(rotErr, tErr) > 10000000.0) return;
    if ((tErr + rotErr) < kEps) return;
    if (tErr > rotErr) {
        double dDist[3] = {d[0] / tErr, d[1] / tErr, d[2] / tErr};
        // cols 3..5: dDist * (-invRot1)
        for (int j = 0; j < 3; j++)
            jac[3 + j] = -(dDist[0] * invRot1[0 * 3 + j] + dDist[1] * invRot1[1 * 3 + j] + dDist[2] * invRot1[2 * 3 + j]);
        // dInvT1_dInvRot1 (3x9), maxloss.h:146-158
        double D[27];
        std::memset(D, 0, sizeof(D));
        for (int r = 0; r < 3; r++) {
            D[r * 9 + r] = a1[0];
            D[r * 9 + 3 + r] = a1[1];
            D[r * 9 + 6 + r] = a1[2];
        }
        double v9[9];
        for (int k = 0; k < 9; k++) v9[k] = dDist[0] * D[k] + dDist[1] * D[9 + k] + dDist[2] * D[18 + k];
        // times dRod^T (9x3): result_j = sum_k v9[k] * dRod[j*9+k]
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 9; k++) s += v9[k] * dRod[j * 9 + k];
            jac[j] = s;
        }
    } else {
        // dRotDiff^T rows: maxloss.h:168-181 ; dTrace picks entries 0,4,8
        double M[81];
        std::memset(M, 0, sizeof(M));
        for (int blk = 0; blk < 3; blk++)
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) M[(blk * 3 + r) * 9 + blk * 3 + c] = invRot2[r * 3 + c];
        // dRotDiff = M^T ; dTrace * dRotDiff = rows 0,4,8 of M^T summed = columns 0,4,8 of M
        double v9[9];
        for (int k = 0; k < 9; k++) v9[k] = M[k * 9 + 0] + M[k * 9 + 4] + M[k * 9 + 8];
        double coef = 180 / kPi * -1 / std::sqrt(3 - trace * trace + 2 * trace);
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 9; k++) s += v9[k] * dRod[j * 9 + k];
            jac[j] = coef * s;
        }
    }
    for (int i = 0; i < 6; i++)
        if (jac[i] != jac[i]) {
            for (int k = 0; k < 6; k++) jac[k] = 0;
            return;
        }
}

}  // extern "C"

#include "dsac_oracle_pipeline.inc"
