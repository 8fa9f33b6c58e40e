// test_ransac_softam -- the reference's soft-argmax test driver (/root/reference/core/test_ransac_softam.cpp)
// on the CUDA engine, in synthetic-input mode: the dataset + coordinate CNN are replaced by the
// SURVEY.md section 8(d) generator (7Scenes-chess-shaped camera trajectory, BASELINE config 5); everything after
// the scene-coordinate grid is the reference's pipeline and its log files keep the reference's formats.
//
//   ./test_ransac_softam [-frames 1000] [-batch 128] [-gpus 1] [-rI 256] ... (flags of properties.cpp)
#include <algorithm>
#include <chrono>
#include <cmath>
#include <fstream>
#include <iostream>
#include <thread>

#include "../dsac_b200/host/cnn_softam.h"
#include "../dsac_b200/host/thread_rand.h"

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
    gp->parseConfig();
    gp->parseCmdLine(argc, argv);
    const int objHyps = gp->pP.ransacIterations;
    const int nFrames = gp->eP.frames, batch = std::max(1, gp->eP.batch), nGpus = std::max(1, gp->eP.gpus);
    const std::string modelFileRGB = gp->dP.objModel;
    const int N = DSAC_N;

    std::vector<FrameResult> all(nFrames);
    std::vector<int> rcs(nGpus, 0);
    std::vector<std::string> errs(nGpus);
    std::vector<double> busy(nGpus, 0.0);   // seconds inside processImages per GPU
    // frames are independent: contiguous shards, one engine + host thread per GPU, no collective
    std::vector<std::thread> workers;
    for (int g = 0; g < nGpus; g++)
        workers.emplace_back([&, g]() {
            int q = nFrames / nGpus, r = nFrames % nGpus;
            int lo = g * q + std::min(g, r), hi = lo + q + (g < r ? 1 : 0);
            dsac_config cfg = gp->engineConfig(batch, g);
            cfg.write_diffmaps = 0;   // the test driver only needs scores
            dsac_engine* eng = nullptr;
            if (dsac_engine_create(&cfg, &eng) != DSAC_OK) { rcs[g] = -1; errs[g] = dsac_last_error(nullptr); return; }
            std::vector<short> coords((size_t)batch * N * 3);
            std::vector<int> pix((size_t)batch * N * 2);
            std::vector<double> gtJp((size_t)batch * 12);
            for (int f0 = lo; f0 < hi; f0 += batch) {
                int n = std::min(batch, hi - f0);
                dsac_synth_frames(20170721u, gp->eP.seed, gp->eP.streams, f0, n, gp->eP.inlierRatio, gp->eP.noise,
                                  gp->eP.trajectory, cfg.focal, cfg.cx, cfg.cy, coords.data(), pix.data(), nullptr, gtJp.data());
                std::vector<FrameResult> res;
                auto t0 = std::chrono::high_resolution_clock::now();
                int rc = processImages(eng, n, f0, coords.data(), pix.data(), gtJp.data(), res);
                busy[g] += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
                if (rc != DSAC_OK) { rcs[g] = rc; errs[g] = dsac_last_error(eng); break; }
                for (int i = 0; i < n; i++) all[f0 + i] = std::move(res[i]);
            }
            dsac_engine_destroy(eng);
        });
    for (auto& w : workers) w.join();
    for (int g = 0; g < nGpus; g++)
        if (rcs[g] != 0) { std::cerr << "engine on GPU " << g << " failed: " << errs[g] << std::endl; return 1; }
    double secs = 0;
    for (double b : busy) secs = std::max(secs, b);

    const std::string tag = modelFileRGB + "_rdraw" + std::to_string((int)gp->pP.randomDraw) + "_softam.txt";
    std::ofstream testFile("ransac_test_loss_" + tag), testErrFile("ransac_test_errors_" + tag);
    double avgCorrect = 0;
    std::vector<double> losses, sfEntropies, rotErrs, tErrs;
    for (int i = 0; i < nFrames; i++) {
        const FrameResult& r = all[i];
        avgCorrect += r.correct;
        // back to the 7-Scenes norm (test_ransac_softam.cpp:161-210): invert, flip y/z axes, mm -> m
        jp::jp_trans_t jpHyp = jp::cv2our(r.refAvgHyp);
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
        std::ifstream transFile("translation.txt");   // optional per-scene offset (test_ransac_softam.cpp:196-210)
        if (transFile.is_open()) {
            double a, b, c;
            if (transFile >> a >> b >> c) { v[3] += a; v[4] += b; v[5] += c; }
        }
        testErrFile << r.loss << " " << r.sfEntropy << " " << r.tErr << " " << r.rotErr << " " << v[0] << " " << v[1] << " "
                    << v[2] << " " << v[3] << " " << v[4] << " " << v[5] << " " << std::endl;
        losses.push_back(r.loss); sfEntropies.push_back(r.sfEntropy); tErrs.push_back(r.tErr); rotErrs.push_back(r.rotErr);
    }
    double lossMean, lossSd, entMean, entSd;
    meanStdDev(losses, lossMean, lossSd);
    meanStdDev(sfEntropies, entMean, entSd);
    avgCorrect /= nFrames;
    std::sort(rotErrs.begin(), rotErrs.end());
    std::sort(tErrs.begin(), tErrs.end());
    double medianRotErr = rotErrs[rotErrs.size() / 2], medianTErr = tErrs[tErrs.size() / 2];
    std::cout << "-----------------------------------------------------------" << std::endl;
    std::cout << "Avg. test loss: " << lossMean << ", accuracy: " << avgCorrect * 100 << "%" << std::endl;
    std::cout << "Median Rot. Error: " << medianRotErr << "deg, Median T. Error: " << medianTErr / 10 << "cm." << std::endl;
    std::cout << nFrames << " frames x " << objHyps << " hypotheses on " << nGpus << " GPU(s) in " << secs << " s: "
              << nFrames / secs << " frames/s, " << (double)nFrames * objHyps / secs << " hypotheses/s (time inside processImages: H2D + kernels + D2H + unpacking; max over GPUs)" << std::endl;
    testFile << avgCorrect << " " << lossMean << " " << lossSd << " " << entMean << " " << entSd << " " << medianRotErr << " "
             << medianTErr << std::endl;
    return 0;
}
