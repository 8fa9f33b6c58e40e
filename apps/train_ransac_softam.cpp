// train_ransac_softam -- the reference's soft-argmax training driver (/root/reference/core/train_ransac_softam.cpp)
// on the CUDA engine in synthetic-input mode: per round a random frame goes through processImage and the
// backward pass of train_ransac_softam.cpp:288-394 (dsac_backward).  The coordinate-CNN update itself
// (lua_calls.h:229, out of scope) is replaced by printing the gradient statistics the reference prints.
//
//   ./train_ransac_softam [-rounds via -frames N] [-rI 256] ...
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
    const int trainingRounds = gp->eP.frames, N = DSAC_N;
    dsac_config cfg = gp->engineConfig(1, 0);
    dsac_engine* eng = nullptr;
    if (dsac_engine_create(&cfg, &eng) != DSAC_OK) { std::cerr << dsac_last_error(nullptr) << std::endl; return 1; }
    std::ofstream trainFile("ransac_training_loss_" + gp->dP.objScript + ".txt");
    std::vector<short> coords((size_t)N * 3);
    std::vector<int> pix((size_t)N * 2);
    std::vector<double> gtJp(12), grad((size_t)N * 3);
    const int poolSize = 1000;   // "training set"
    ThreadRand::forceInit(gp->eP.seed);
    for (int round = 0; round <= trainingRounds; round++) {
        std::cout << "Round " << round << " of " << trainingRounds << "." << std::endl;
        int imgID = irand(0, poolSize);   // train_ransac_softam.cpp:228
        dsac_synth_frames(20170721u, gp->eP.seed, gp->eP.streams, imgID, 1, gp->eP.inlierRatio, gp->eP.noise, 0, cfg.focal,
                          cfg.cx, cfg.cy, coords.data(), pix.data(), nullptr, gtJp.data());
        auto t0 = std::chrono::high_resolution_clock::now();
        std::vector<FrameResult> res;
        int rc = processImages(eng, 1, imgID, coords.data(), pix.data(), gtJp.data(), res);
        if (rc != DSAC_OK) { std::cerr << dsac_last_error(eng) << std::endl; return 1; }
        dsac_backward_out bo;
        std::memset(&bo, 0, sizeof(bo));
        bo.dloss_dobj = grad.data();
        rc = dsac_backward(eng, 1, nullptr, nullptr, 0, nullptr, &bo);
        if (rc != DSAC_OK) { std::cerr << dsac_last_error(eng) << std::endl; return 1; }
        double ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
        // gradient statistics (train_ransac_softam.cpp:397-408)
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
        std::cout << "Rotation Err: " << res[0].rotErr << ", Translation Err: " << res[0].tErr << std::endl;
        std::cout << "Max gradient: " << mx << "\nAvg gradient: " << avg / (N * 3) << "\nMed gradient: " << vals[vals.size() / 2]
                  << "\nZero gradients: " << zeroGrads << "\nForward + backward in " << ms << " ms." << std::endl;
        trainFile << round << " " << res[0].loss << " " << res[0].sfEntropy << std::endl;   // train_ransac_softam.cpp:416-420
    }
    dsac_engine_destroy(eng);
    return 0;
}
