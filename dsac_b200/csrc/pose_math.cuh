// pose_math.cuh -- fp64 pose arithmetic of the hypothesis engine (device + host-testable).
//
// Everything here is written from the published algorithms (Rodrigues formula; Gao et al.
// P3P, PAMI 2003; absolute orientation of congruent triangles) for one-thread-per-problem
// execution in registers.  It replaces, for the hot path, what the reference obtains from
// OpenCV: cv::solvePnP(CV_P3P) via safeSolvePnP (cnn_softam.h:56-73, call site :1042),
// cv::Rodrigues (types.h:190) and cv::projectPoints (cnn_softam.h:351,1046).
//
// The header compiles both under nvcc (device code) and under a plain C++ compiler with
// DSAC_HOST_ONLY defined; tests/test_host_math.py uses the latter to check the product
// arithmetic against the oracle on the CPU box (the product itself never runs on the CPU).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__) && !defined(DSAC_HOST_ONLY)
#define DSAC_HD __host__ __device__ __forceinline__
#define DSAC_HDN __host__ __device__
#else
#define DSAC_HD inline
#define DSAC_HDN inline
#endif

#define DSAC_GRID_CONST 40  /* CNN_OBJ_PATCHSIZE, lua_calls.h:33 */
#define DSAC_N_CONST 1600
#define DSAC_MAXINPUT_F 100.0f  /* CNN_OBJ_MAXINPUT, lua_calls.h:36 */

namespace dsac {

constexpr double kPi = 3.14159265358979323846;

// ----------------------------------------------------------------------------- Rodrigues
// rvec -> R (row-major).  Same formula as cv::Rodrigues: R = c I + (1-c) r r^T + s [r]x.
DSAC_HD void rodrigues_v2m(const double r[3], double R[9]) {
    double th2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    double theta = sqrt(th2);
    if (theta < 2.220446049250313e-16) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        return;
    }
    double s, c;
#if defined(__CUDA_ARCH__)
    sincos(theta, &s, &c);
#else
    s = sin(theta); c = cos(theta);
#endif
    double c1 = 1.0 - c, it = 1.0 / theta;
    double x = r[0] * it, y = r[1] * it, z = r[2] * it;
    R[0] = c + c1 * x * x;     R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
    R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y;     R[5] = c1 * y * z - s * x;
    R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}

// R -> rvec for an orthonormal R (cv::Rodrigues' matrix branch without its SVD projection,
// which is the identity for the rotation matrices produced here).
DSAC_HD void rodrigues_m2v(const double R[9], double r[3]) {
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) {
            r[0] = r[1] = r[2] = 0;
        } else {
            double t;
            t = (R[0] + 1) * 0.5; rx = sqrt(t > 0 ? t : 0.);
            t = (R[4] + 1) * 0.5; ry = sqrt(t > 0 ? t : 0.) * (R[1] < 0 ? -1. : 1.);
            t = (R[8] + 1) * 0.5; rz = sqrt(t > 0 ? t : 0.) * (R[2] < 0 ? -1. : 1.);
            if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
            theta /= sqrt(rx * rx + ry * ry + rz * rz);
            r[0] = rx * theta; r[1] = ry * theta; r[2] = rz * theta;
        }
    } else {
        double vth = theta / (2 * s);
        r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
    }
}

// cv::projectPoints for one point, zero distortion (double arithmetic, z==0 -> 1).
DSAC_HD void project_point(const double R[9], const double t[3], double X, double Y, double Z, double f, double cx,
                           double cy, double* u, double* v) {
    double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
    double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
    double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    z = z ? 1. / z : 1;
    x *= z;
    y *= z;
    *u = x * f + cx;
    *v = y * f + cy;
}

// ----------------------------------------------------------------------------- P3P
// Largest real root of y^3 + a2 y^2 + a1 y + a0 (the resolvent cubic of Ferrari's method).
DSAC_HD double cubic_first_root(double a2, double a1, double a0) {
    double Q = (3 * a1 - a2 * a2) / 9;
    double R = (9 * a2 * a1 - 27 * a0 - 2 * a2 * a2 * a2) / 54;
    double Q3 = Q * Q * Q, D = Q3 + R * R, sh = a2 / 3;
    if (Q == 0) return (R == 0) ? -sh : cbrt(2 * R) - sh;
    if (D <= 0) {
        double theta = acos(R / sqrt(-Q3));
        return 2 * sqrt(-Q) * cos(theta / 3.0) - sh;
    }
    double AD = cbrt(fabs(R) + sqrt(D)) * (R > 0 ? 1 : (R < 0 ? -1 : 0));
    double BD = (AD == 0) ? 0 : -Q / AD;
    return AD + BD - sh;
}

// Real roots of the quartic (closed form; accuracy is restored afterwards by Newton steps
// on the two-quadric system, see p3p_best).
DSAC_HD int quartic_roots(double a, double b, double c, double d, double e, double x[4]) {
    if (a == 0) return 0;
    double ia = 1.0 / a;
    b *= ia; c *= ia; d *= ia; e *= ia;
    double y1 = cubic_first_root(-c, d * b - 4 * e, 4 * c * e - d * d - b * b * e);
    double R2 = 0.25 * b * b - c + y1;
    if (!(R2 >= 0)) return 0;
    double R = sqrt(R2), D2, E2;
    if (R < 10e-12) {
        double t = y1 * y1 - 4 * e;
        if (t < 0) {
            D2 = E2 = -1;
        } else {
            double st = sqrt(t);
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
        double Dq = sqrt(D2);
        x[0] = 0.5 * R + 0.5 * Dq - 0.25 * b;
        x[1] = x[0] - Dq;
        n = 2;
    }
    if (E2 >= 0) {
        double Eq = sqrt(E2);
        x[n] = -0.5 * R + 0.5 * Eq - 0.25 * b;
        x[n + 1] = x[n] - Eq;
        n += 2;
    }
    return n;
}

struct P3PProblem {
    double mu[4], mv[4];  // pixel positions as P3P sees them (float-rounded normalised coords mapped back through K)
    double X[4][3];       // scene coordinates (mm)
};

// Orthonormal frame of a triangle: e1 along P1-P0, e3 along the normal, e2 = e3 x e1.
DSAC_HD void triangle_frame(const double P0[3], const double P1[3], const double P2[3], double E[9]) {
    double ax = P1[0] - P0[0], ay = P1[1] - P0[1], az = P1[2] - P0[2];
    double bx = P2[0] - P0[0], by = P2[1] - P0[1], bz = P2[2] - P0[2];
    double ia = 1.0 / sqrt(ax * ax + ay * ay + az * az);
    ax *= ia; ay *= ia; az *= ia;
    double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
    double in = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
    nx *= in; ny *= in; nz *= in;
    E[0] = ax; E[1] = ay; E[2] = az;
    E[3] = ny * az - nz * ay; E[4] = nz * ax - nx * az; E[5] = nx * ay - ny * ax;
    E[6] = nx; E[7] = ny; E[8] = nz;
}

// Pose from one P3P solution: camera-frame points L_i * bearing_i aligned with the world
// triangle (congruent after the Newton polish, so the frame-to-frame rotation equals the
// least-squares absolute orientation to rounding).
DSAC_HD void pose_from_lengths(const double Lx, const double Ly, const double Lz, const double bear[3][3],
                               const double Ew[9], const double Cw[3], double R[9], double t[3]) {
    double M0[3] = {Lx * bear[0][0], Lx * bear[0][1], Lx * bear[0][2]};
    double M1[3] = {Ly * bear[1][0], Ly * bear[1][1], Ly * bear[1][2]};
    double M2[3] = {Lz * bear[2][0], Lz * bear[2][1], Lz * bear[2][2]};
    double Ec[9];
    triangle_frame(M0, M1, M2, Ec);
    // R = Ec^T * Ew  (rows of E are the frame axes)
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i * 3 + j] = Ec[0 + i] * Ew[0 + j] + Ec[3 + i] * Ew[3 + j] + Ec[6 + i] * Ew[6 + j];
    double Cc[3] = {(M0[0] + M1[0] + M2[0]) / 3, (M0[1] + M1[1] + M2[1]) / 3, (M0[2] + M1[2] + M2[2]) / 3};
    for (int i = 0; i < 3; i++) t[i] = Cc[i] - (R[i * 3] * Cw[0] + R[i * 3 + 1] * Cw[1] + R[i * 3 + 2] * Cw[2]);
}

// All P3P solutions from points 0..2, the one with the smallest squared reprojection
// error of point 3 is returned (cv::solvePnP CV_P3P semantics).  Returns the number of
// solutions; *best_err2 is that smallest squared error (pixels^2).
DSAC_HDN int p3p_best(const P3PProblem& pr, double f, double cx, double cy, double Rbest[9], double tbest[3],
                      double* best_err2) {
    double inv_f = 1.0 / f, cx_f = cx / f, cy_f = cy / f;
    double bear[3][3];
    for (int i = 0; i < 3; i++) {
        double u = inv_f * pr.mu[i] - cx_f, v = inv_f * pr.mv[i] - cy_f;
        double k = 1. / sqrt(u * u + v * v + 1);
        bear[i][0] = u * k; bear[i][1] = v * k; bear[i][2] = k;
    }
    double d12, d02, d01;
    {
        double dx = pr.X[1][0] - pr.X[2][0], dy = pr.X[1][1] - pr.X[2][1], dz = pr.X[1][2] - pr.X[2][2];
        d12 = sqrt(dx * dx + dy * dy + dz * dz);
        dx = pr.X[0][0] - pr.X[2][0]; dy = pr.X[0][1] - pr.X[2][1]; dz = pr.X[0][2] - pr.X[2][2];
        d02 = sqrt(dx * dx + dy * dy + dz * dz);
        dx = pr.X[0][0] - pr.X[1][0]; dy = pr.X[0][1] - pr.X[1][1]; dz = pr.X[0][2] - pr.X[1][2];
        d01 = sqrt(dx * dx + dy * dy + dz * dz);
    }
    double p = 2 * (bear[1][0] * bear[2][0] + bear[1][1] * bear[2][1] + bear[1][2] * bear[2][2]);
    double q = 2 * (bear[0][0] * bear[2][0] + bear[0][1] * bear[2][1] + bear[0][2] * bear[2][2]);
    double r = 2 * (bear[0][0] * bear[1][0] + bear[0][1] * bear[1][1] + bear[0][2] * bear[1][2]);
    double inv_c2 = 1.0 / (d01 * d01);
    double a = inv_c2 * d12 * d12, b = inv_c2 * d02 * d02;
    if (p * p + q * q + r * r - p * q * r - 1 == 0) return 0;
    // x = |PA|/|PC|, y = |PB|/|PC|:  y * Dn(x) = -Nn(x)
    double N2 = 1 - a - b, N1 = q * (a - 1), N0 = 1 - a + b;
    double D1 = b * r, D0 = -b * p;
    double c4, c3, c2, c1, c0;
    {
        double F2 = 1 - b, F1 = -q;
        double DD2 = D1 * D1, DD1 = 2 * D1 * D0, DD0 = D0 * D0;
        double br = b * r;
        c4 = F2 * DD2 - b * (N2 * N2) - br * (N2 * D1);
        c3 = F2 * DD1 + F1 * DD2 - b * (2 * N2 * N1) - br * (N2 * D0 + N1 * D1);
        c2 = F2 * DD0 + F1 * DD1 + DD2 - b * (2 * N2 * N0 + N1 * N1) - br * (N1 * D0 + N0 * D1);
        c1 = F1 * DD0 + DD1 - b * (2 * N1 * N0) - br * (N0 * D0);
        c0 = DD0 - b * (N0 * N0);
    }
    if (c4 == 0) return 0;
    double xr[4];
    int nroots = quartic_roots(c4, c3, c2, c1, c0, xr);
    if (nroots == 0) return 0;

    double Ew[9], Cw[3];
    triangle_frame(pr.X[0], pr.X[1], pr.X[2], Ew);
    for (int k = 0; k < 3; k++) Cw[k] = (pr.X[0][k] + pr.X[1][k] + pr.X[2][k]) / 3;

    double sx[4], sy[4];
    int ns = 0, nsol = 0;
    double best = 0;
    for (int i = 0; i < nroots; i++) {
        double x0 = xr[i];
        if (!(x0 == x0)) continue;
        double Dn = D1 * x0 + D0;
        double y0a, y0b;
        int nc;
        if (fabs(Dn) > 1e-3 * (fabs(D1 * x0) + fabs(D0))) {
            y0a = -((N2 * x0 + N1) * x0 + N0) / Dn;
            y0b = y0a;
            nc = 1;
        } else {  // both roots of the first quadric are candidates
            double qa = 1 - a, qb = a * r * x0 - p, qc = 1 - a * x0 * x0;
            double disc = qb * qb - 4 * qa * qc;
            if (disc < 0) disc = 0;
            double sq = sqrt(disc);
            if (qa != 0) {
                y0a = (-qb + sq) / (2 * qa);
                y0b = (-qb - sq) / (2 * qa);
                nc = 2;
            } else if (qb != 0) {
                y0a = y0b = -qc / qb;
                nc = 1;
            } else {
                y0a = y0b = 0;
                nc = 0;
            }
        }
        for (int c = 0; c < nc; c++) {
            double x = x0, y = c ? y0b : y0a;
            bool good = false;
            for (int it = 0; it < 8; it++) {
                double f1 = (1 - a) * y * y - a * x * x - p * y + a * r * x * y + 1;
                double f2 = (1 - b) * x * x - b * y * y - q * x + b * r * x * y + 1;
                double j11 = -2 * a * x + a * r * y, j12 = 2 * (1 - a) * y - p + a * r * x;
                double j21 = 2 * (1 - b) * x - q + b * r * y, j22 = -2 * b * y + b * r * x;
                double det = j11 * j22 - j12 * j21;
                if (det == 0 || !(det == det)) break;
                double idet = 1.0 / det;
                double dx = (f1 * j22 - f2 * j12) * idet, dy = (j11 * f2 - j21 * f1) * idet;
                x -= dx;
                y -= dy;
                if (fabs(dx) + fabs(dy) <= 1e-15 * (fabs(x) + fabs(y))) {
                    good = true;
                    break;
                }
            }
            if (!good) {
                double f1 = (1 - a) * y * y - a * x * x - p * y + a * r * x * y + 1;
                double f2 = (1 - b) * x * x - b * y * y - q * x + b * r * x * y + 1;
                good = fabs(f1) + fabs(f2) < 1e-12 * (1 + x * x + y * y);
            }
            if (!good || !(x > 0) || !(y > 0)) continue;
            bool dup = false;
            for (int k = 0; k < 4; k++)
                if (k < ns && fabs(sx[k] - x) + fabs(sy[k] - y) < 1e-9 * (fabs(x) + fabs(y))) dup = true;
            if (dup || ns >= 4) continue;
            sx[ns] = x;
            sy[ns] = y;
            ns++;
            double v = x * x + y * y - x * y * r;
            if (!(v > 0)) continue;
            double Z = d01 / sqrt(v);
            double R[9], t[3];
            pose_from_lengths(x * Z, y * Z, Z, bear, Ew, Cw, R, t);
            double X3 = R[0] * pr.X[3][0] + R[1] * pr.X[3][1] + R[2] * pr.X[3][2] + t[0];
            double Y3 = R[3] * pr.X[3][0] + R[4] * pr.X[3][1] + R[5] * pr.X[3][2] + t[1];
            double Z3 = R[6] * pr.X[3][0] + R[7] * pr.X[3][1] + R[8] * pr.X[3][2] + t[2];
            double u3 = cx + f * X3 / Z3, v3 = cy + f * Y3 / Z3;
            double e2 = (u3 - pr.mu[3]) * (u3 - pr.mu[3]) + (v3 - pr.mv[3]) * (v3 - pr.mv[3]);
            if (nsol == 0 || best > e2) {
                best = e2;
                for (int k = 0; k < 9; k++) Rbest[k] = R[k];
                tbest[0] = t[0]; tbest[1] = t[1]; tbest[2] = t[2];
            }
            nsol++;
        }
    }
    *best_err2 = best;
    return nsol;
}

// The pixel a P3P solve actually sees: cv::undistortPoints rounds the normalised
// coordinate to float, p3p maps it back through K.
DSAC_HD double p3p_pixel(float pix, double c, double f) {
    float n = (float)(((double)pix - c) * (1. / f));
    return n * f + c;
}

// Minimal-set hypothesis exactly as the sampling loop evaluates it (cnn_softam.h:1041-1059):
// P3P on the 4 correspondences, then all 4 reprojection errors (projection rounded to
// float, float difference, double norm) must be below the integer threshold.
// Returns true if accepted; rvec/tvec are the cv pose.  *fragile is set when a decision
// was within 1e-6 px of the threshold.
DSAC_HDN bool minimal_set_hypothesis(const float obj[12], const float img[8], double f, double cx, double cy, int thr,
                                     double rvec[3], double tvec[3], bool* fragile) {
    P3PProblem pr;
    for (int i = 0; i < 4; i++) {
        pr.mu[i] = p3p_pixel(img[i * 2], cx, f);
        pr.mv[i] = p3p_pixel(img[i * 2 + 1], cy, f);
        pr.X[i][0] = obj[i * 3]; pr.X[i][1] = obj[i * 3 + 1]; pr.X[i][2] = obj[i * 3 + 2];
    }
    double R[9], t[3], e2;
    *fragile = false;
    if (p3p_best(pr, f, cx, cy, R, t, &e2) == 0) return false;
    // cheap exact-safe pre-check: the 4th point's error as P3P measured it differs from the
    // reference's float-rounded check by < 1e-4 px
    if (!(e2 < ((double)thr + 1e-3) * ((double)thr + 1e-3))) return false;
    rodrigues_m2v(R, rvec);
    tvec[0] = t[0]; tvec[1] = t[1]; tvec[2] = t[2];
    double Rp[9];
    rodrigues_v2m(rvec, Rp);  // cv::projectPoints rebuilds R from rvec
    bool ok = true;
    for (int j = 0; j < 4; j++) {
        double u, v;
        project_point(Rp, tvec, obj[j * 3], obj[j * 3 + 1], obj[j * 3 + 2], f, cx, cy, &u, &v);
        float du = img[j * 2] - (float)u, dv = img[j * 2 + 1] - (float)v;
        double nrm = sqrt((double)du * du + (double)dv * dv);
        if (fabs(nrm - thr) < 1e-6) *fragile = true;
        if (!(nrm < thr)) ok = false;
    }
    return ok;
}

}  // namespace dsac
