// test_ransac -- the reference's DSAC / RANSAC test driver (/root/reference/core/test_ransac.cpp) on the CUDA engine
// in synthetic-input mode: processImage of core/cnn.h -- draw one hypothesis, refine ALL
// hypotheses, expected loss -- through dsac_forward_dsac.  Log files keep the reference's formats.
//
//   ./test_ransac [-frames 200] [-batch 50] [-rdraw 1] [-rI 256] ...
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>

#include "../dsac_b200/host/cnn_softam.h"

using namespace cvlite;

static void meanStdDev(const std::vector<double>& v, double& mean, double& sd) {
    mean = 0;
    for (double x : v) mean += x;
    mean /= std::max<size_t>(1, v.size());
    sd = 0;
    for (double x : v) sd += (x - mean) * (x - mean);
    sd = std::sqrt(sd / std::max<size_t>(1, v.size()));
}

int main(int argc, const char* argv[]) {
    GlobalProperties* gp = GlobalProperties::getInstance();
    gp->eP.frames = 200;
    gp->eP.batch = 50;
    gp->parseConfig();
    gp->parseCmdLine(argc, argv);
    const int H = gp->pP.ransacIterations, nFrames = gp->eP.frames, batch = std::max(1, gp->eP.batch), N = DSAC_N;
    dsac_config cfg = gp->engineConfig(batch, 0);
    cfg.write_diffmaps = 0;
    dsac_engine* eng = nullptr;
    if (dsac_engine_create(&cfg, &eng) != DSAC_OK) { std::cerr << dsac_last_error(nullptr) << std::endl; return 1; }
    const std::string tag = gp->dP.objModel + "_rdraw" + std::to_string((int)gp->pP.randomDraw) + ".txt";
    std::ofstream testFile("ransac_test_loss_" + tag), testErrFile("ransac_test_errors_" + tag);
    std::vector<short> coords((size_t)batch * N * 3);
    std::vector<int> pix((size_t)batch * N * 2);
    std::vector<double> gtJp((size_t)batch * 12), ref((size_t)batch * H * 6), losses((size_t)batch * H), ent(batch), expl(batch), rot(batch), terr(batch);
    std::vector<int32_t> hypIdx(batch), correct(batch);
    double avgCorrect = 0, busy = 0;
    std::vector<double> expLosses, sfEntropies, rotErrs, tErrs;
    for (int f0 = 0; f0 < nFrames; f0 += batch) {
        int n = std::min(batch, nFrames - f0);
        dsac_synth_frames(20170721u, gp->eP.seed, gp->eP.streams, f0, n, gp->eP.inlierRatio, gp->eP.noise, gp->eP.trajectory,
                          cfg.focal, cfg.cx, cfg.cy, coords.data(), pix.data(), nullptr, gtJp.data());
        dsac_dsac_out out;
        std::memset(&out, 0, sizeof(out));
        out.ref_pose = ref.data(); out.losses = losses.data(); out.entropy = ent.data(); out.expected_loss = expl.data();
        out.hyp_idx = hypIdx.data(); out.rot_err = rot.data(); out.t_err = terr.data(); out.correct = correct.data();
        auto t0 = std::chrono::high_resolution_clock::now();
        int rc = dsac_forward_dsac(eng, n, f0, coords.data(), pix.data(), 0, gtJp.data(), gp->pP.randomDraw ? 1 : 0, &out);
        busy += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
        if (rc != DSAC_OK) { std::cerr << dsac_last_error(eng) << std::endl; return 1; }
        for (int i = 0; i < n; i++) {
            avgCorrect += correct[i];
            const double* p = &ref[((size_t)i * H + hypIdx[i]) * 6];
            // back to the 7-Scenes norm (test_ransac.cpp:170-216)
            jp::jp_trans_t jpHyp = jp::cv2our(jp::cv_trans_t(Vec3d{p[0], p[1], p[2]}, Vec3d{p[3], p[4], p[5]}));
            Hypothesis hyp(jpHyp.first, jpHyp.second);
            Matd T = inv(hyp.getTransformation());
            Matd corr = Matd::eye(4, 4);
            corr(1, 1) = -1; corr(2, 2) = -1;
            T = T * corr;
            Matd R(3, 3);
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) R(a, b) = T(a, b);
            hyp.setTranslation(Point3d(T(0, 3), T(1, 3), T(2, 3)));
            hyp.setRotation(R);
            std::vector<double> v = hyp.getRodVecAndTrans();
            v[3] /= 1000; v[4] /= 1000; v[5] /= 1000;
            std::ifstream transFile("translation.txt");
            if (transFile.is_open()) {
                double a, b, c;
                if (transFile >> a >> b >> c) { v[3] += a; v[4] += b; v[5] += c; }
            }
            testErrFile << expl[i] << " " << ent[i] << " " << losses[(size_t)i * H + hypIdx[i]] << " " << terr[i] << " " << rot[i] << " "
                        << v[0] << " " << v[1] << " " << v[2] << " " << v[3] << " " << v[4] << " " << v[5] << " " << std::endl;
            expLosses.push_back(expl[i]); sfEntropies.push_back(ent[i]); tErrs.push_back(terr[i]); rotErrs.push_back(rot[i]);
        }
    }
    double lossMean, lossSd, entMean, entSd;
    meanStdDev(expLosses, lossMean, lossSd);
    meanStdDev(sfEntropies, entMean, entSd);
    avgCorrect /= nFrames;
    std::sort(rotErrs.begin(), rotErrs.end());
    std::sort(tErrs.begin(), tErrs.end());
    double medianRotErr = rotErrs[rotErrs.size() / 2], medianTErr = tErrs[tErrs.size() / 2];
    std::cout << "-----------------------------------------------------------" << std::endl;
    std::cout << "Avg. test loss: " << lossMean << ", accuracy: " << avgCorrect * 100 << "%" << std::endl;
    std::cout << "Median Rot. Error: " << medianRotErr << "deg, Median T. Error: " << medianTErr / 10 << "cm." << std::endl;
    std::cout << nFrames << " frames x " << H << " hypotheses (all refined) in " << busy << " s inside dsac_forward_dsac: "
              << (double)nFrames * H / busy << " hypotheses refined/s" << std::endl;
    testFile << avgCorrect << " " << lossMean << " " << lossSd << " " << entMean << " " << entSd << " " << medianRotErr << " "
             << medianTErr << std::endl;
    dsac_engine_destroy(eng);
    return 0;
}
