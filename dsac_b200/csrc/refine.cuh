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
// The pose arithmetic is fp64 (the reference runs this in double).  The error maps are only ever compared with
// the inlier threshold, so they are kept as bit maps, decided in fp32 with a rigorous error bound and in the
// reference's exact arithmetic where the bound does not separate the error from the threshold.
#pragma once
#include <cuda_runtime.h>

#include "lm_math.cuh"
#include "pose_math.cuh"

namespace dsac {

#ifndef K4_THREADS_DEF
#define K4_THREADS_DEF 64
#endif
#ifndef K4_MIN_BLOCKS
#define K4_MIN_BLOCKS 8
#endif
constexpr int K4_THREADS = K4_THREADS_DEF;   // 2 warps per job: the work per job is small and mostly sequential
constexpr int K4_WARPS = K4_THREADS / 32;
constexpr int K4_MAX_INLIERS = 128;
static_assert(K4_THREADS % 32 == 0 && K4_THREADS >= 32 && K4_THREADS <= 256, "k_refine block size");

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

// One CTA per job.  Structure of a refinement step:
//   flags    all threads: "error < threshold" per cell as a bit map -- fp32 with a rigorous error bound, the exact
//            reference arithmetic (reproj_error_exact) only for cells within the bound of the threshold;
//   select   warp 0 walks the step's permutation with ballots (no block barrier) and takes the first
//            inlier_count flagged cells in order;
//   LM       CvLevMarq's state machine with ONE pass over the selected points per trial parameter vector: the pass
//            accumulates the error norm AND the normal equations at the trial point (which the reference would
//            compute in its next CALC_J if the trial is accepted -- the same numbers); warp 0 reduces, decides,
//            solves the damped 6x6 system in registers and builds the next R / dR/dr across its lanes.  Two block
//            barriers per trial.
__global__ void __launch_bounds__(K4_THREADS, K4_MIN_BLOCKS) k_refine(RefineParams p) {
    // the frame's scene coordinates (int16 mm, this job's perturbation applied) and pixel positions, staged once per job:
    // every refinement step reads all of them for the inlier flags and the selected ones in each LM pass
    __shared__ short s_cX[DSAC_N_CONST], s_cY[DSAC_N_CONST], s_cZ[DSAC_N_CONST];
    __shared__ float s_fu[DSAC_N_CONST], s_fv[DSAC_N_CONST];
    __shared__ uint32_t s_bits[DSAC_N_CONST / 32];
    __shared__ unsigned char s_imap[DSAC_N_CONST];
    __shared__ unsigned short s_sel[K4_MAX_INLIERS];
    __shared__ double s_R[9], s_J[27], s_param[6], s_prev[6], s_pose[6], s_Rd[9];
    __shared__ float s_Pf[12];
    __shared__ double s_red[K4_WARPS][32];
    __shared__ double s_sum[2][32];
    __shared__ int s_count, s_ctl;

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

    if (tid < 6) s_pose[tid] = p.job_init[(size_t)job * 6 + tid];
    for (int i = tid; i < DSAC_N_CONST; i += K4_THREADS) {
        s_imap[i] = 0;
        int v[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            v[k] = __ldg(coords + i * 3 + k);
            if (i * 3 + k == pert_idx) v[k] = (int)(short)(v[k] + pert_delta);  // short += float eps (cnn_softam.h:887)
        }
        s_cX[i] = (short)v[0]; s_cY[i] = (short)v[1]; s_cZ[i] = (short)v[2];
        s_fu[i] = (float)__ldg(pix + i * 2);
        s_fv[i] = (float)__ldg(pix + i * 2 + 1);
    }
    __syncthreads();

    const float thrf = (float)p.thr;
    const float cxf = (float)p.cx, cyf = (float)p.cy, c_abs = fabsf(cxf) + fabsf(cyf);

    // inlier flags of the current pose s_pose (getDiffMap, cnn_softam.h:319-362, reduced to "< threshold")
    auto inlier_flags = [&]() {
        if (warp == 0) {
            double pose[6], R[9];
#pragma unroll
            for (int k = 0; k < 6; k++) pose[k] = s_pose[k];
            rodrigues_v2m(pose, R);   // (every lane computes the same R; lane k stores element k)
#pragma unroll
            for (int k = 0; k < 9; k++)
                if (lane == k) s_Rd[k] = R[k];
#pragma unroll
            for (int k = 0; k < 12; k++) {
                const int row = k >> 2, col = k & 3;
                const double v = (col < 3) ? R[row * 3 + col] : pose[3 + row];
                if (lane == k) s_Pf[k] = (float)(row < 2 ? p.f * v : v);
            }
        }
        __syncthreads();
        float P[12];
#pragma unroll
        for (int k = 0; k < 12; k++) P[k] = s_Pf[k];
        for (int base = 0; base < DSAC_N_CONST; base += K4_THREADS) {
            const int c = base + tid;
            bool flag = false;
            if (c < DSAC_N_CONST) {
                const float X = (float)s_cX[c], Y = (float)s_cY[c], Z = (float)s_cZ[c];
                const float fu = s_fu[c], fv = s_fv[c];
                const int r = reproj_below_thr_fast(P, X, Y, Z, fu - cxf, fv - cyf, c_abs, thrf);
                if (r >= 0) {
                    flag = (r != 0);
                } else {
                    double Rd[9], t[3];
#pragma unroll
                    for (int k = 0; k < 9; k++) Rd[k] = s_Rd[k];
                    t[0] = s_pose[3]; t[1] = s_pose[4]; t[2] = s_pose[5];
                    flag = reproj_error_exact(Rd, t, (double)X, (double)Y, (double)Z, p.f, p.cx, p.cy, fu, fv) < thrf;
                }
            }
            const uint32_t m = __ballot_sync(0xffffffffu, flag);
            if (lane == 0 && base + warp * 32 < DSAC_N_CONST) s_bits[(base >> 5) + warp] = m;
        }
        __syncthreads();
    };

    // this thread's selected points of the current refinement step (constant over its LM loop), loaded once
    constexpr int K4_PPT = (K4_MAX_INLIERS + K4_THREADS - 1) / K4_THREADS;
    float selX[K4_PPT], selY[K4_PPT], selZ[K4_PPT], selU[K4_PPT], selV[K4_PPT];
    auto load_selected = [&](int n) {
#pragma unroll
        for (int k = 0; k < K4_PPT; k++) {
            const int i = tid + k * K4_THREADS;
            const int c = (i < n) ? s_sel[i] : 0;
            selX[k] = (float)s_cX[c]; selY[k] = (float)s_cY[c]; selZ[k] = (float)s_cZ[c];
            selU[k] = s_fu[c]; selV[k] = s_fv[c];
        }
    };

    // per-thread contribution of the selected points to [0..20] upper triangle of JtJ, [21..26] JtErr, [27] |err|^2
    // at (s_param, s_R, s_J), reduced over the block into s_red[warp][0..31] (lane L holds sum L)
    auto normal_equations_pass = [&](int n) {
        double v[32];
#pragma unroll
        for (int k = 0; k < 32; k++) v[k] = 0;
#pragma unroll
        for (int k = 0; k < K4_PPT; k++)
            if (tid + k * K4_THREADS < n)
                lm_point_contrib(s_R, s_J, s_param, (double)selX[k], (double)selY[k], (double)selZ[k], (double)selU[k], (double)selV[k],
                                 p.f, p.cx, p.cy, v);
        // transposed butterfly: after the step with offset `off` every lane keeps half of its values; lane L ends with sum L
#pragma unroll
        for (int half = 16, off = 16; half >= 1; half >>= 1, off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (i < half) {
                    const double send = up ? v[i] : v[i + half];
                    const double keep = up ? v[i + half] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
        }
        s_red[warp][lane] = v[0];
    };

    inlier_flags();

    int steps_done = 0, n_perm = 0;
    for (int rStep = 0; rStep < step_limit; rStep++) {
        n_perm++;
        // ---- ordered selection of the first inlier_count flagged cells of the step's permutation (cnn_softam.h:1117-1134)
        if (warp == 0) {
            const uint16_t* perm = p.perm + (size_t)rStep * DSAC_N_CONST;
            int count = 0;
            for (int k0 = 0; k0 < DSAC_N_CONST && count < p.inlier_count; k0 += 32) {
                const int c = __ldg(perm + k0 + lane);
                const bool flag = (s_bits[c >> 5] >> (c & 31)) & 1u;
                const uint32_t m = __ballot_sync(0xffffffffu, flag);
                const int rank = count + __popc(m & ((1u << lane) - 1u));
                if (flag && rank < p.inlier_count) {
                    s_sel[rank] = (unsigned short)c;
                    s_imap[c] += 1;  // cnn_softam.h:1129 (each cell appears once per permutation)
                }
                count = min(count + __popc(m), p.inlier_count);
            }
            if (lane == 0) s_count = count;
        }
        __syncthreads();
        const int count = s_count;
        if (count < 50) break;  // cnn_softam.h:1136

        // ---- Levenberg-Marquardt, CvLevMarq(6, 2n, max_iter 20, eps FLT_EPSILON)
        load_selected(count);
        // (state used by warp 0 only, uniform across its lanes; trial / base parameter vectors live in s_param / s_prev)
        LMState lm;
        if (warp == 0) {
            double r0[3] = {s_pose[0], s_pose[1], s_pose[2]};
            if (lane < 6) s_param[lane] = s_pose[lane];
            rodrigues_jac_warp(r0, lane, s_R, s_J);
            if (lane == 0) s_ctl = 0;
        }
        __syncthreads();
        for (;;) {
            if (s_ctl) break;
            normal_equations_pass(count);
            __syncthreads();
            if (warp == 0) {
                const int buf = lm_next_buf(lm);
                {
                    double t = 0;
#pragma unroll
                    for (int w = 0; w < K4_WARPS; w++) t += s_red[w][lane];
                    s_sum[buf][lane] = t;
                }
                __syncwarp();
                const int action = lm_advance(lm, buf, s_sum[buf][27], s_param, s_prev);
                if (action != LM_DONE) {
                    __syncwarp();             // every lane has read s_prev / s_param in lm_advance
                    if (action == LM_SOLVE_NEWBASE && lane < 6) s_prev[lane] = s_param[lane];
                    __syncwarp();
                    double x[6], trial[6];
                    lm_solve6_fast(s_sum[lm.cur], c_lm_lambda[lm.lambdaLg10 + 16], x);
#pragma unroll
                    for (int k = 0; k < 6; k++) trial[k] = s_prev[k] - x[k];
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < 6; k++)
                        if (lane == k) s_param[k] = trial[k];
                    rodrigues_jac_warp(trial, lane, s_R, s_J);
                }
                if (lane == 0) s_ctl = (action == LM_DONE) ? 1 : 0;
            }
            __syncthreads();
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
        inlier_flags();
    }

    // ---- outputs
    if (tid < 6) p.out_pose[(size_t)job * 6 + tid] = s_pose[tid];
    if (p.inlier_map)
        for (int i = tid; i < DSAC_N_CONST; i += K4_THREADS) p.inlier_map[(size_t)job * DSAC_N_CONST + i] = s_imap[i];
    if (tid == 0) {
        if (p.steps_done) p.steps_done[job] = steps_done;
        if (p.n_perm_steps) p.n_perm_steps[job] = n_perm;
        if (p.status && steps_done < p.ref_steps) atomicOr(p.status + frame, 2u /* DSAC_ST_REFINE_ABORTED */);
        double pose[6], R[9], t[3];
        for (int k = 0; k < 6; k++) pose[k] = s_pose[k];
        if (p.out_jp6 || p.gt_jp) cv2our_dev(pose, R, t);
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
