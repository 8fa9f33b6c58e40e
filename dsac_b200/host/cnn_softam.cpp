// cnn_softam.cpp -- see cnn_softam.h.
#include "cnn_softam.h"

#include <algorithm>
#include <cstring>
#include <numeric>
#include <random>

using namespace cvlite;

Mat_<Point2i> stochasticSubSample(int width, int height, unsigned seed) {
    std::vector<int32_t> pix(DSAC_N * 2);
    dsac_stochastic_subsample(seed, width, height, pix.data());
    Mat_<Point2i> s(CNN_OBJ_PATCHSIZE, CNN_OBJ_PATCHSIZE);
    for (int y = 0; y < CNN_OBJ_PATCHSIZE; y++)
        for (int x = 0; x < CNN_OBJ_PATCHSIZE; x++) s(y, x) = Point2i(pix[(y * CNN_OBJ_PATCHSIZE + x) * 2], pix[(y * CNN_OBJ_PATCHSIZE + x) * 2 + 1]);
    return s;
}

int processImages(dsac_engine* engine, int n, long long frame0, const short* coords, const int* pix, const double* gtJp,
                  std::vector<FrameResult>& results) {
    dsac_config cfg;
    int rc = dsac_engine_config(engine, &cfg);
    if (rc != DSAC_OK) return rc;
    const int H = cfg.n_hyps, N = DSAC_N;
    std::vector<double> hyp((size_t)n * H * 6), sf((size_t)n * H), ent(n), avg((size_t)n * 6), ref((size_t)n * 6), loss(n), rot(n), terr(n);
    std::vector<int32_t> idx((size_t)n * H * 4), imap((size_t)n * N), correct(n), nperm(n);
    std::vector<uint32_t> status(n);
    dsac_forward_out out;
    std::memset(&out, 0, sizeof(out));
    out.hyp_pose = hyp.data(); out.img_idx = idx.data(); out.sf = sf.data(); out.entropy = ent.data();
    out.avg_pose = avg.data(); out.ref_pose = ref.data(); out.inlier_map = imap.data(); out.n_perm_steps = nperm.data();
    out.loss = loss.data(); out.rot_err = rot.data(); out.t_err = terr.data(); out.correct = correct.data();
    out.status = status.data();
    rc = dsac_forward(engine, n, frame0, coords, pix, 0, gtJp, &out);
    if (rc != DSAC_OK) return rc;
    // the refinement permutations are the same for every frame: std::mt19937 default-seeded per frame,
    // iota + std::shuffle per step with the continuing generator (cnn_softam.h:1104-1114)
    std::vector<std::vector<int>> perms(cfg.ref_steps);
    {
        std::mt19937 randG;
        for (int s = 0; s < cfg.ref_steps; s++) {
            perms[s].resize(N);
            std::iota(perms[s].begin(), perms[s].end(), 0);
            std::shuffle(perms[s].begin(), perms[s].end(), randG);
        }
    }
    results.assign(n, FrameResult());
    for (int f = 0; f < n; f++) {
        FrameResult& r = results[f];
        r.loss = loss[f]; r.sfEntropy = ent[f]; r.tErr = terr[f]; r.rotErr = rot[f]; r.correct = correct[f] != 0;
        r.status = status[f];
        r.hyps.resize(H); r.imgPts.resize(H); r.objPts.resize(H); r.imgIdx.resize(H); r.sampledPoints.resize(H);
        r.sfScores.assign(sf.begin() + (size_t)f * H, sf.begin() + (size_t)(f + 1) * H);
        for (int h = 0; h < H; h++) {
            const double* p = &hyp[((size_t)f * H + h) * 6];
            r.hyps[h] = jp::cv_trans_t(Vec3d{p[0], p[1], p[2]}, Vec3d{p[3], p[4], p[5]});
            for (int j = 0; j < 4; j++) {
                int c = idx[((size_t)f * H + h) * 4 + j];
                r.imgIdx[h].push_back(c);
                if (c < 0) continue;
                int x = c % CNN_OBJ_PATCHSIZE, y = c / CNN_OBJ_PATCHSIZE;
                r.sampledPoints[h].push_back(Point2i(x, y));
                r.imgPts[h].push_back(Point2f((float)pix[((size_t)f * N + c) * 2], (float)pix[((size_t)f * N + c) * 2 + 1]));
                const short* q = coords + ((size_t)f * N + c) * 3;
                r.objPts[h].push_back(Point3f(q[0], q[1], q[2]));
            }
        }
        const double* a = &avg[(size_t)f * 6];
        const double* b = &ref[(size_t)f * 6];
        r.avgHyp = jp::cv_trans_t(Vec3d{a[0], a[1], a[2]}, Vec3d{a[3], a[4], a[5]});
        r.refAvgHyp = jp::cv_trans_t(Vec3d{b[0], b[1], b[2]}, Vec3d{b[3], b[4], b[5]});
        r.inlierMap = Mat_<int>(CNN_OBJ_PATCHSIZE, CNN_OBJ_PATCHSIZE);
        for (int i = 0; i < N; i++) r.inlierMap(i / CNN_OBJ_PATCHSIZE, i % CNN_OBJ_PATCHSIZE) = imap[(size_t)f * N + i];
        r.pixelIdxs.assign(cfg.ref_steps, std::vector<int>());
        for (int s = 0; s < nperm[f] && s < cfg.ref_steps; s++) r.pixelIdxs[s] = perms[s];
    }
    return DSAC_OK;
}

int processImage(dsac_engine* engine, long long frameIndex, const Hypothesis& poseGT, int objHyps, int ptCount,
                 const Mat_<float>& camMat, int inlierThreshold2D, int inlierCount, int refSteps, double& loss,
                 double& sfEntropy, bool& correct, std::vector<jp::cv_trans_t>& hyps, jp::cv_trans_t& refAvgHyp,
                 jp::cv_trans_t& avgHyp, std::vector<std::vector<Point2f>>& imgPts, std::vector<std::vector<Point3f>>& objPts,
                 std::vector<std::vector<int>>& imgIdx, std::vector<double>& sfScores, const jp::img_coord_t& estObj,
                 const Mat_<Point2i>& sampling, std::vector<std::vector<Point2i>>& sampledPoints, Mat_<int>& inlierMap,
                 std::vector<std::vector<int>>& pixelIdxs, double& tErr, double& rotErr) {
    dsac_config cfg;
    int rc = dsac_engine_config(engine, &cfg);
    if (rc != DSAC_OK) return rc;
    // the engine was created from these very values; a mismatch is a caller error, not something to paper over
    if (cfg.n_hyps != objHyps || ptCount != 4 || cfg.thr2d != inlierThreshold2D || cfg.inlier_count != inlierCount ||
        cfg.ref_steps != refSteps || (float)cfg.focal != camMat(0, 0) || (float)cfg.cx != camMat(0, 2) || (float)cfg.cy != camMat(1, 2))
        return DSAC_ERR_ARG;
    const int N = DSAC_N;
    std::vector<short> coords((size_t)N * 3);
    std::vector<int> pix((size_t)N * 2);
    for (int y = 0; y < CNN_OBJ_PATCHSIZE; y++)
        for (int x = 0; x < CNN_OBJ_PATCHSIZE; x++) {
            int c = y * CNN_OBJ_PATCHSIZE + x;
            for (int k = 0; k < 3; k++) coords[c * 3 + k] = estObj(y, x)(k);
            pix[c * 2] = sampling(y, x).x;
            pix[c * 2 + 1] = sampling(y, x).y;
        }
    double gt[12];
    Matd R = poseGT.getRotation();
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) gt[i * 3 + j] = R(i, j);
    gt[9] = poseGT.getTranslation().x; gt[10] = poseGT.getTranslation().y; gt[11] = poseGT.getTranslation().z;
    std::vector<FrameResult> res;
    rc = processImages(engine, 1, frameIndex, coords.data(), pix.data(), gt, res);
    if (rc != DSAC_OK) return rc;
    FrameResult& r = res[0];
    loss = r.loss; sfEntropy = r.sfEntropy; correct = r.correct; hyps = r.hyps; refAvgHyp = r.refAvgHyp; avgHyp = r.avgHyp;
    imgPts = r.imgPts; objPts = r.objPts; imgIdx = r.imgIdx; sfScores = r.sfScores; sampledPoints = r.sampledPoints;
    inlierMap = r.inlierMap; pixelIdxs = r.pixelIdxs; tErr = r.tErr; rotErr = r.rotErr;
    return DSAC_OK;
}
