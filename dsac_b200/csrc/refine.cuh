// refine.cuh -- K4: batched refinement of pose hypotheses (one CTA per job).
//
// Restates the refinement loop of processImage (cnn_softam.h:1099-1154) and its replay
// refine() (cnn_softam.h:663-723): per step walk a fixed permutation of the 1600 cells,
// take the first <= inlierCount cells whose reprojection error is below the threshold, stop
// if fewer than 50, run cv::solvePnP(CV_ITERATIVE, useExtrinsicGuess) -- Levenberg-Marquardt
// on the 6 pose parameters with CvLevMarq's schedule -- and recompute the error map.
//
// A "job" is (frame, initial cv pose, optional +-delta on one int16 scene coordinate), so
// the same kernel serves the forward pass (1 job per frame) and the central differences of
// dRefineHyp / dRefineObj (cnn_softam.h:738-923; 12 + 6*k jobs per frame).
//
// All arithmetic is fp64 (the reference runs this in double); selection is a block-wide
// ordered compaction; the 6x6 normal equations are reduced with warp shuffles.
#pragma once
#include <cuda_runtime.h>

#include "pose_math.cuh"

namespace dsac {

constexpr int K4_THREADS = 128;
constexpr int K4_WARPS = K4_THREADS / 32;
constexpr int K4_MAX_INLIERS = 128;

struct RefineParams {
    const int16_t* coords;      // [n][N][3]
    const int32_t* pix;         // [n or 1][N][2]
    int pix_stride;
    const uint16_t* perm;       // [ref_steps][N]  (std::shuffle with default mt19937, cnn_softam.h:1104-1114)
    double f, cx, cy;
    int thr, inlier_count, ref_steps;
    // jobs
    int n_jobs;
    const int32_t* job_frame;   // [n_jobs] or null (job j -> frame j)
    const double* job_init;     // [n_jobs][6] cv pose
    const int32_t* job_pert;    // [n_jobs][2] {cell*3+chan or -1, delta} or null
    const int32_t* max_steps;   // [n_frames] replay limit (n_perm_steps of the forward) or null
    const int32_t* job_max_steps;   // [n_jobs] replay limit per job (DSAC variant: per hypothesis) or null
    // outputs
    double* out_pose;           // [n_jobs][6] refined cv pose
    double* out_jp6;            // [n_jobs][6] Hypothesis(cv2our(pose)).getRodVecAndTrans() or null
    int32_t* inlier_map;        // [n_jobs][N] or null
    int32_t* steps_done;        // [n_jobs] or null
    int32_t* n_perm_steps;      // [n_jobs] or null
    uint32_t* status;           // [n_frames] or null
    // evaluation (cnn_softam.h:1160-1179)
    const double* gt_jp;        // [n_frames][12] or null
    double* loss;
    double* rot_err;
    double* t_err;
    int32_t* correct;
};

__device__ __forceinline__ double dot3_plain(double a0, double b0, double a1, double b1, double a2, double b2, double c) {
    // left-to-right, no FMA contraction: matches the scalar double code of cv::projectPoints
    return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1)), __dmul_rn(a2, b2)), c);
}

// Rodrigues Jacobian, 3x9 (row i = d vec(R)/d r_i), as cv::Rodrigues returns it.
__device__ void rodrigues_jac(const double r[3], double R[9], double J[27]) {
    double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < 2.220446049250313e-16) {
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        for (int i = 0; i < 27; i++) J[i] = 0;
        J[5] = J[15] = J[19] = -1;
        J[7] = J[11] = J[21] = 1;
        return;
    }
    double s, c;
    sincos(theta, &s, &c);
    double c1 = 1.0 - c, itheta = 1.0 / theta;
    double rx = r[0] * itheta, ry = r[1] * itheta, rz = r[2] * itheta;
    double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    for (int k = 0; k < 9; k++) R[k] = c * ((k % 4 == 0) ? 1.0 : 0.0) + c1 * rrt[k] + s * r_x[k];
    const double drrt[27] = {rx + rx, ry, rz, ry, 0, 0, rz, 0, 0, 0, rx, 0, rx, ry + ry, rz, 0, rz, 0,
                             0, 0, rx, 0, 0, ry, rx, ry, rz + rz};
    const double d_r_x[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0};
    for (int i = 0; i < 3; i++) {
        double ri = (i == 0) ? rx : (i == 1) ? ry : rz;
        double a0 = -s * ri, a1 = (s - 2 * c1 * itheta) * ri, a2 = c1 * itheta;
        double a3 = (c - s * itheta) * ri, a4 = s * itheta;
        for (int k = 0; k < 9; k++)
            J[i * 9 + k] = a0 * ((k % 4 == 0) ? 1.0 : 0.0) + a1 * rrt[k] + a2 * drrt[i * 9 + k] + a3 * r_x[k] +
                           a4 * d_r_x[i * 9 + k];
    }
}

// Solve (JtJ with diagonal * (1+lambda)) x = JtErr.  Cholesky when positive definite (the
// normal case), else minimum-norm solution through a Jacobi eigen-decomposition with
// cv::SVBkSb's threshold (CvLevMarq::step solves with an SVD).
__device__ void lm_solve6(const double* JtJ, const double* JtErr, double lambda, double x[6]) {
    double A[36];
    for (int i = 0; i < 36; i++) A[i] = JtJ[i];
    for (int i = 0; i < 6; i++) A[i * 7] *= 1. + lambda;
    double L[36];
    bool pd = true;
    for (int i = 0; i < 6 && pd; i++) {
        for (int j = 0; j <= i; j++) {
            double s = A[i * 6 + j];
            for (int k = 0; k < j; k++) s -= L[i * 6 + k] * L[j * 6 + k];
            if (i == j) {
                if (!(s > 1e-300)) { pd = false; break; }
                L[i * 6 + i] = sqrt(s);
            } else {
                L[i * 6 + j] = s / L[j * 6 + j];
            }
        }
    }
    if (pd) {
        double y[6];
        for (int i = 0; i < 6; i++) {
            double s = JtErr[i];
            for (int k = 0; k < i; k++) s -= L[i * 6 + k] * y[k];
            y[i] = s / L[i * 6 + i];
        }
        for (int i = 5; i >= 0; i--) {
            double s = y[i];
            for (int k = i + 1; k < 6; k++) s -= L[k * 6 + i] * x[k];
            x[i] = s / L[i * 6 + i];
        }
        // guard against a numerically singular factorisation
        bool finite = true;
        for (int i = 0; i < 6; i++)
            if (!(fabs(x[i]) < 1.7e308)) finite = false;
        if (finite) return;
    }
    // symmetric eigen-decomposition (cyclic Jacobi), pseudo-inverse
    double V[36];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) V[i * 6 + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < 5; p++)
            for (int q = p + 1; q < 6; q++) off += fabs(A[p * 6 + q]);
        if (off == 0.0) break;
        for (int p = 0; p < 5; p++)
            for (int q = p + 1; q < 6; q++) {
                double apq = A[p * 6 + q];
                if (apq == 0.0) continue;
                double theta = (A[q * 6 + q] - A[p * 6 + p]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 6; k++) {
                    double akp = A[k * 6 + p], akq = A[k * 6 + q];
                    A[k * 6 + p] = c * akp - s * akq;
                    A[k * 6 + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 6; k++) {
                    double apk = A[p * 6 + k], aqk = A[q * 6 + k];
                    A[p * 6 + k] = c * apk - s * aqk;
                    A[q * 6 + k] = s * apk + c * aqk;
                }
                A[p * 6 + q] = A[q * 6 + p] = 0.0;
                for (int k = 0; k < 6; k++) {
                    double vkp = V[k * 6 + p], vkq = V[k * 6 + q];
                    V[k * 6 + p] = c * vkp - s * vkq;
                    V[k * 6 + q] = s * vkp + c * vkq;
                }
            }
    }
    double thr = 0;
    for (int i = 0; i < 6; i++) thr += fabs(A[i * 7]);
    thr *= 2.220446049250313e-16 * 2;
    for (int i = 0; i < 6; i++) x[i] = 0;
    for (int k = 0; k < 6; k++) {
        double w = A[k * 7];
        if (!(fabs(w) > thr)) continue;
        double d = 0;
        for (int i = 0; i < 6; i++) d += V[i * 6 + k] * JtErr[i];
        d /= w;
        for (int i = 0; i < 6; i++) x[i] += d * V[i * 6 + k];
    }
}

// jp::cv2our, types.h:186-214
__device__ void cv2our_dev(const double pose[6], double R[9], double t[3]) {
    rodrigues_v2m(pose, R);
    t[0] = pose[3]; t[1] = -pose[4]; t[2] = -pose[5];
    for (int j = 0; j < 3; j++) {
        R[3 + j] = -R[3 + j];
        R[6 + j] = -R[6 + j];
    }
    double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
    if (det < 0) {
        for (int k = 0; k < 9; k++) R[k] = -R[k];
        for (int k = 0; k < 3; k++) t[k] = -t[k];
    }
    if (t[0] != t[0] || t[1] != t[1] || t[2] != t[2]) t[0] = t[1] = t[2] = 0;
}

// getInvHyp, maxloss.h:39-61
__device__ void inv_pose_dev(const double R[9], const double t[3], double Ri[9], double ti[3]) {
    double d = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
    double id = 1.0 / d;
    Ri[0] = (R[4] * R[8] - R[5] * R[7]) * id; Ri[1] = (R[2] * R[7] - R[1] * R[8]) * id; Ri[2] = (R[1] * R[5] - R[2] * R[4]) * id;
    Ri[3] = (R[5] * R[6] - R[3] * R[8]) * id; Ri[4] = (R[0] * R[8] - R[2] * R[6]) * id; Ri[5] = (R[2] * R[3] - R[0] * R[5]) * id;
    Ri[6] = (R[3] * R[7] - R[4] * R[6]) * id; Ri[7] = (R[1] * R[6] - R[0] * R[7]) * id; Ri[8] = (R[0] * R[4] - R[1] * R[3]) * id;
    ti[0] = -(Ri[0] * t[0] + Ri[1] * t[1] + Ri[2] * t[2]);
    ti[1] = -(Ri[3] * t[0] + Ri[4] * t[1] + Ri[5] * t[2]);
    ti[2] = -(Ri[6] * t[0] + Ri[7] * t[1] + Ri[8] * t[2]);
}

// maxLoss (maxloss.h:69-79) + the errors of cnn_softam.h:1166-1170
__device__ double max_loss_dev(const double R1[9], const double t1[3], const double R2[9], const double t2[3],
                               double* rot_err, double* t_err) {
    double Ri1[9], ti1[3], Ri2[9], ti2[3], Ri2inv[9], z[3] = {0, 0, 0}, zz[3];
    inv_pose_dev(R1, t1, Ri1, ti1);
    inv_pose_dev(R2, t2, Ri2, ti2);
    inv_pose_dev(Ri2, z, Ri2inv, zz);
    double tr = 0;
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) tr += Ri1[i * 3 + k] * Ri2inv[k * 3 + i];
    tr = fmin(3.0, fmax(-1.0, tr));
    double re = 180 * acos((tr - 1.0) / 2.0) / kPi;
    double dx = ti1[0] - ti2[0], dy = ti1[1] - ti2[1], dz = ti1[2] - ti2[2];
    double te = sqrt(dx * dx + dy * dy + dz * dz);
    *rot_err = re;
    *t_err = te;
    return fmin(fmax(re, te / 10), 10000000.0);
}

__global__ void __launch_bounds__(K4_THREADS) k_refine(RefineParams p) {
    __shared__ float s_diff[DSAC_N_CONST];
    __shared__ int s_imap[DSAC_N_CONST];
    __shared__ int s_sel[K4_MAX_INLIERS];
    __shared__ double s_R[9], s_J[27], s_param[6], s_prev[6], s_pose[6];
    __shared__ double s_red[K4_WARPS][28];
    __shared__ double s_sum[28];
    __shared__ int s_scan[K4_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int job = blockIdx.x;
    const int frame = p.job_frame ? p.job_frame[job] : job;
    if (frame < 0) return;  // unused job slot
    const int16_t* coords = p.coords + (size_t)frame * DSAC_N_CONST * 3;
    const int32_t* pix = p.pix + (size_t)frame * p.pix_stride;
    int pert_idx = -1, pert_delta = 0;
    if (p.job_pert) {
        pert_idx = p.job_pert[job * 2];
        pert_delta = p.job_pert[job * 2 + 1];
    }
    const int step_limit = p.job_max_steps ? min(p.ref_steps, p.job_max_steps[job])
                                           : (p.max_steps ? min(p.ref_steps, p.max_steps[frame]) : p.ref_steps);

    auto coord = [&](int c, int k) -> double {
        int v = coords[c * 3 + k];
        if (c * 3 + k == pert_idx) v = (int)(short)(v + pert_delta);  // short += float eps (cnn_softam.h:887)
        return (double)(float)v;
    };

    if (tid < 6) s_pose[tid] = p.job_init[(size_t)job * 6 + tid];
    for (int i = tid; i < DSAC_N_CONST; i += K4_THREADS) s_imap[i] = 0;
    __syncthreads();

    // getDiffMap, cnn_softam.h:319-362, with the exact float/double roundings of the reference
    auto diffmap = [&]() {
        if (tid == 0) rodrigues_v2m(s_pose, s_R);
        __syncthreads();
        for (int c = tid; c < DSAC_N_CONST; c += K4_THREADS) {
            double X = coord(c, 0), Y = coord(c, 1), Z = coord(c, 2);
            double x = dot3_plain(s_R[0], X, s_R[1], Y, s_R[2], Z, s_pose[3]);
            double y = dot3_plain(s_R[3], X, s_R[4], Y, s_R[5], Z, s_pose[4]);
            double z = dot3_plain(s_R[6], X, s_R[7], Y, s_R[8], Z, s_pose[5]);
            z = z ? __ddiv_rn(1., z) : 1;
            x = __dmul_rn(x, z);
            y = __dmul_rn(y, z);
            float pu = (float)__dadd_rn(__dmul_rn(x, p.f), p.cx), pv = (float)__dadd_rn(__dmul_rn(y, p.f), p.cy);
            float du = __fsub_rn((float)pix[c * 2], pu), dv = __fsub_rn((float)pix[c * 2 + 1], pv);
            double nrm = sqrt(__dadd_rn(__dmul_rn((double)du, (double)du), __dmul_rn((double)dv, (double)dv)));
            s_diff[c] = (float)fmin(nrm, 100.0);
        }
        __syncthreads();
    };

    // block reduction of the per-point normal-equation contributions
    auto reduce28 = [&](double* v, int nvals) {
        for (int k = 0; k < nvals; k++) {
            double x = v[k];
#pragma unroll
            for (int off = 16; off; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
            if (lane == 0) s_red[warp][k] = x;
        }
        __syncthreads();
        if (tid < nvals) {
            double t = 0;
            for (int w = 0; w < K4_WARPS; w++) t += s_red[w][tid];
            s_sum[tid] = t;
        }
        __syncthreads();
    };

    // residuals (and Jacobian) of the selected points at s_param; fills s_sum:
    //   with jac:  [0..20] upper triangle of JtJ, [21..26] JtErr, [27] |err|^2 ; without: [27] only
    auto accumulate = [&](int n, bool jac) {
        if (tid == 0) {
            if (jac) rodrigues_jac(s_param, s_R, s_J);
            else rodrigues_v2m(s_param, s_R);
        }
        __syncthreads();
        double v[28];
#pragma unroll
        for (int k = 0; k < 28; k++) v[k] = 0;
        if (tid < n) {
            int c = s_sel[tid];
            double X = coord(c, 0), Y = coord(c, 1), Z = coord(c, 2);
            double x = s_R[0] * X + s_R[1] * Y + s_R[2] * Z + s_param[3];
            double y = s_R[3] * X + s_R[4] * Y + s_R[5] * Z + s_param[4];
            double z = s_R[6] * X + s_R[7] * Y + s_R[8] * Z + s_param[5];
            z = z ? 1. / z : 1;
            x *= z;
            y *= z;
            double eu = (x * p.f + p.cx) - (double)(float)pix[c * 2];
            double ev = (y * p.f + p.cy) - (double)(float)pix[c * 2 + 1];
            v[27] = eu * eu + ev * ev;
            if (jac) {
                double ju[6], jv[6];
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    double dx0 = X * s_J[j * 9 + 0] + Y * s_J[j * 9 + 1] + Z * s_J[j * 9 + 2];
                    double dy0 = X * s_J[j * 9 + 3] + Y * s_J[j * 9 + 4] + Z * s_J[j * 9 + 5];
                    double dz0 = X * s_J[j * 9 + 6] + Y * s_J[j * 9 + 7] + Z * s_J[j * 9 + 8];
                    ju[j] = p.f * (z * (dx0 - x * dz0));
                    jv[j] = p.f * (z * (dy0 - y * dz0));
                }
                ju[3] = p.f * z; ju[4] = 0; ju[5] = p.f * (-x * z);
                jv[3] = 0; jv[4] = p.f * z; jv[5] = p.f * (-y * z);
                int k = 0;
#pragma unroll
                for (int a = 0; a < 6; a++)
#pragma unroll
                    for (int b = a; b < 6; b++) v[k++] = ju[a] * ju[b] + jv[a] * jv[b];
#pragma unroll
                for (int a = 0; a < 6; a++) v[21 + a] = ju[a] * eu + jv[a] * ev;
            }
        }
        if (jac) reduce28(v, 28);
        else {
            double x = v[27];
#pragma unroll
            for (int off = 16; off; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
            if (lane == 0) s_red[warp][27] = x;
            __syncthreads();
            if (tid == 0) {
                double t = 0;
                for (int w = 0; w < K4_WARPS; w++) t += s_red[w][27];
                s_sum[27] = t;
            }
            __syncthreads();
        }
    };

    diffmap();

    int steps_done = 0, n_perm = 0;
    const float thrf = (float)p.thr;
    for (int rStep = 0; rStep < step_limit; rStep++) {
        n_perm++;
        // ---- ordered selection of the first inlier_count cells below the threshold
        int count = 0;
        const uint16_t* perm = p.perm + (size_t)rStep * DSAC_N_CONST;
        for (int base = 0; base < DSAC_N_CONST && count < p.inlier_count; base += K4_THREADS) {
            int k = base + tid;
            int c = (k < DSAC_N_CONST) ? perm[k] : 0;
            int flag = (k < DSAC_N_CONST) && (s_diff[c] < thrf);
            // block exclusive scan
            int incl = flag;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                int nb = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += nb;
            }
            __syncthreads();
            if (lane == 31) s_scan[warp] = incl;
            __syncthreads();
            int pre = 0, tot = 0;
            for (int w = 0; w < K4_WARPS; w++) {
                if (w < warp) pre += s_scan[w];
                tot += s_scan[w];
            }
            int rank = count + pre + incl - flag;
            if (flag && rank < p.inlier_count) {
                s_sel[rank] = c;
                s_imap[c] += 1;  // cnn_softam.h:1129 (each cell appears once per permutation)
            }
            count = min(count + tot, p.inlier_count);
        }
        __syncthreads();
        if (count < 50) break;  // cnn_softam.h:1136

        // ---- Levenberg-Marquardt, CvLevMarq(6, 2n, max_iter 20, eps FLT_EPSILON)
        if (tid < 6) s_param[tid] = s_pose[tid];
        __syncthreads();
        int lambdaLg10 = -3, iters = 0;
        double prevErrNorm = 1.7976931348623157e308, errNorm = 0;
        accumulate(count, true);
        for (;;) {
            // CALC_J -> step
            if (tid == 0) {
                double JtJ[36], JtErr[6], x[6];
                int k = 0;
                for (int a = 0; a < 6; a++)
                    for (int b = a; b < 6; b++) {
                        JtJ[a * 6 + b] = JtJ[b * 6 + a] = s_sum[k];
                        k++;
                    }
                for (int a = 0; a < 6; a++) JtErr[a] = s_sum[21 + a];
                for (int a = 0; a < 6; a++) s_prev[a] = s_param[a];
                lm_solve6(JtJ, JtErr, exp(lambdaLg10 * 2.302585092994046), x);
                for (int a = 0; a < 6; a++) s_param[a] = s_prev[a] - x[a];
            }
            if (iters == 0) prevErrNorm = sqrt(s_sum[27]);
            __syncthreads();
            bool done = false;
            for (;;) {  // CHECK_ERR
                accumulate(count, false);
                errNorm = sqrt(s_sum[27]);
                if (errNorm > prevErrNorm) {
                    if (++lambdaLg10 <= 16) {
                        if (tid == 0) {
                            // JtJ / JtErr of the last CALC_J are still in s_sum[0..26]
                            double JtJ[36], JtErr[6], x[6];
                            int k = 0;
                            for (int a = 0; a < 6; a++)
                                for (int b = a; b < 6; b++) {
                                    JtJ[a * 6 + b] = JtJ[b * 6 + a] = s_sum[k];
                                    k++;
                                }
                            for (int a = 0; a < 6; a++) JtErr[a] = s_sum[21 + a];
                            lm_solve6(JtJ, JtErr, exp(lambdaLg10 * 2.302585092994046), x);
                            for (int a = 0; a < 6; a++) s_param[a] = s_prev[a] - x[a];
                        }
                        __syncthreads();
                        continue;
                    }
                }
                lambdaLg10 = max(lambdaLg10 - 1, -16);
                double dn = 0, pn = 0;
                for (int a = 0; a < 6; a++) {
                    double d = s_param[a] - s_prev[a];
                    dn += d * d;
                    pn += s_prev[a] * s_prev[a];
                }
                double change = sqrt(dn) / sqrt(pn);
                if (++iters >= 20 || change < 1.1920928955078125e-07) done = true;
                break;
            }
            if (done) break;
            prevErrNorm = errNorm;
            accumulate(count, true);
        }
        // NaN -> abort without accepting (cnn_softam.h:1147)
        bool nan = false;
        for (int a = 0; a < 6; a++)
            if (s_param[a] != s_param[a]) nan = true;
        if (nan) break;
        __syncthreads();
        if (tid < 6) s_pose[tid] = s_param[tid];
        __syncthreads();
        steps_done++;
        diffmap();
    }

    // ---- outputs
    if (tid < 6) p.out_pose[(size_t)job * 6 + tid] = s_pose[tid];
    if (p.inlier_map)
        for (int i = tid; i < DSAC_N_CONST; i += K4_THREADS) p.inlier_map[(size_t)job * DSAC_N_CONST + i] = s_imap[i];
    if (tid == 0) {
        if (p.steps_done) p.steps_done[job] = steps_done;
        if (p.n_perm_steps) p.n_perm_steps[job] = n_perm;
        if (p.status && steps_done < p.ref_steps) atomicOr(p.status + frame, 2u /* DSAC_ST_REFINE_ABORTED */);
        double R[9], t[3];
        if (p.out_jp6 || p.gt_jp) cv2our_dev(s_pose, R, t);
        if (p.out_jp6) {
            double r[3];
            rodrigues_m2v(R, r);
            double* o = p.out_jp6 + (size_t)job * 6;
            o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = t[0]; o[4] = t[1]; o[5] = t[2];
        }
        if (p.gt_jp) {
            const double* g = p.gt_jp + (size_t)frame * 12;
            double re, te;
            double l = max_loss_dev(g, g + 9, R, t, &re, &te);
            p.loss[job] = l;
            p.rot_err[job] = re;
            p.t_err[job] = te;
            p.correct[job] = (re < 5 && te < 50) ? 1 : 0;
        }
    }
}

}  // namespace dsac
