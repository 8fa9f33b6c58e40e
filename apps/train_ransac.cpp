// train_ransac -- the reference's DSAC / RANSAC training driver (/root/reference/core/train_ransac.cpp) on the CUDA
// engine in synthetic-input mode: per round a random frame goes through processImage of
// core/cnn.h (dsac_forward_dsac: draw, refinement of all hypotheses, expected loss) and the gradient block of
// train_ransac.cpp:304-381 (dsac_backward_dsac: sum_h sf_h dLossMax . dRefine_h over the likely hypotheses + dSMScore).
// The coordinate-CNN update itself (backward(), lua_calls.h:229, out of scope) is replaced by the gradient statistics
// the reference prints before it (train_ransac.cpp:383-396).
//
//   ./train_ransac [-frames ROUNDS] [-rdraw 1] [-rI 256] ...
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>

#include "../dsac_b200/host/cnn_softam.h"
#include "../dsac_b200/host/thread_rand.h"

int main(int argc, const char* argv[]) {
    GlobalProperties* gp = GlobalProperties::getInstance();
    gp->eP.frames = 20;   // rounds in synthetic mode (the reference runs 5000)
    gp->parseConfig();
    gp->parseCmdLine(argc, argv);
    const int trainingRounds = gp->eP.frames, N = DSAC_N, H = gp->pP.ransacIterations;
    dsac_config cfg = gp->engineConfig(1, 0);
    cfg.write_diffmaps = 0;
    dsac_engine* eng = nullptr;
    if (dsac_engine_create(&cfg, &eng) != DSAC_OK) { std::cerr << dsac_last_error(nullptr) << std::endl; return 1; }
    std::ofstream trainFile("ransac_training_loss_" + gp->dP.objScript + ".txt");
    std::vector<short> coords((size_t)N * 3);
    std::vector<int> pix((size_t)N * 2);
    std::vector<double> gtJp(12), grad((size_t)N * 3), sf(H);
    const int poolSize = 1000;   // "training set"
    ThreadRand::forceInit(gp->eP.seed);
    for (int round = 0; round <= trainingRounds; round++) {
        std::cout << "Round " << round << " of " << trainingRounds << "." << std::endl;
        int imgID = irand(0, poolSize);   // train_ransac.cpp:231
        dsac_synth_frames(20170721u, gp->eP.seed, gp->eP.streams, imgID, 1, gp->eP.inlierRatio, gp->eP.noise, 0, cfg.focal,
                          cfg.cx, cfg.cy, coords.data(), pix.data(), nullptr, gtJp.data());
        auto t0 = std::chrono::high_resolution_clock::now();
        double expectedLoss = 0, sfEntropy = 0, rotErr = 0, tErr = 0;
        int32_t hypIdx = 0, correct = 0, nSel = 0, nJobs = 0;
        dsac_dsac_out fo;
        std::memset(&fo, 0, sizeof(fo));
        fo.sf = sf.data(); fo.expected_loss = &expectedLoss; fo.entropy = &sfEntropy; fo.hyp_idx = &hypIdx;
        fo.rot_err = &rotErr; fo.t_err = &tErr; fo.correct = &correct;
        int rc = dsac_forward_dsac(eng, 1, imgID, coords.data(), pix.data(), 0, gtJp.data(), gp->pP.randomDraw ? 1 : 0, &fo);
        if (rc != DSAC_OK) { std::cerr << dsac_last_error(eng) << std::endl; return 1; }
        dsac_backward_dsac_out bo;
        std::memset(&bo, 0, sizeof(bo));
        bo.dloss_dobj = grad.data(); bo.n_selected = &nSel; bo.n_refine_jobs = &nJobs;
        rc = dsac_backward_dsac(eng, 1, &bo);
        if (rc != DSAC_OK) { std::cerr << dsac_last_error(eng) << std::endl; return 1; }
        double ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
        // gradient statistics (train_ransac.cpp:383-396)
        int zeroGrads = 0;
        double mx = -1, avg = 0;
        std::vector<double> vals;
        for (int p = 0; p < N; p++) {
            double nn = 0;
            for (int c = 0; c < 3; c++) {
                double v = std::fabs(grad[p * 3 + c]);
                nn += v * v; avg += v; vals.push_back(v);
                if (mx < 0 || v > mx) mx = v;
            }
            if (std::sqrt(nn) < EPS) zeroGrads++;
        }
        std::sort(vals.begin(), vals.end());
        std::cout << "Loss of winning hyp prob: " << sf[hypIdx] << ", expected loss: " << expectedLoss << std::endl;
        std::cout << "Rotation Err: " << rotErr << ", Translation Err: " << tErr << std::endl;
        std::cout << "Max gradient: " << mx << "\nAvg gradient: " << avg / (N * 3) << "\nMed gradient: " << vals[vals.size() / 2]
                  << "\nZero gradients: " << zeroGrads << "\n" << nSel << " hypotheses, " << nJobs << " refinements differentiated; forward + backward in "
                  << ms << " ms." << std::endl;
        trainFile << round << " " << expectedLoss << " " << sfEntropy << std::endl;   // train_ransac.cpp:404-408
    }
    dsac_engine_destroy(eng);
    return 0;
}
