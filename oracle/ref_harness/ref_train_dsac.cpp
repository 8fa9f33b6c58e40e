// ref_train_dsac.cpp -- oracle/_ref/libref_train_dsac.so: the reference's OWN DSAC training driver, core/train_ransac.cpp
// (forward + the expectation-of-loss gradient of lines 304-381), compiled UNMODIFIED with its main() renamed, run on a
// synthetic one-frame dataset.  dLoss_dObj is captured where the reference hands it to the coordinate CNN
// (train_ransac.cpp:399 -> lua_calls.h:229); the loop is then left by an exception.  TEST INFRASTRUCTURE ONLY.
#include <unistd.h>

#include <iostream>
#include <sstream>

#include "ref_env.h"

#define main ref_train_ransac_main
#include "train_ransac.cpp"
#undef main

extern "C" {

struct ref_config {
    double alpha, beta, grad_clamp;
    int32_t n_hyps, thr2d, inlier_count, ref_steps;
    float sub_sample;
    uint32_t seed;
    int32_t n_threads;
    int64_t frame;
};

int ref_train_dsac_round(const ref_config* c, const char* dir, const int16_t* coords, int n_args, const char** args, double* dloss,
                         double* loss) {
    std::streambuf* old_buf = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    g_env.coords = coords; g_env.n_frames = 1; g_env.frame0 = c->frame;
    g_env.alpha = c->alpha; g_env.beta = c->beta; g_env.grad_clamp = c->grad_clamp; g_env.thr = c->thr2d;
    g_env.seed = c->seed; g_env.T = c->n_threads;
    g_env.backward_calls = 0; g_env.stop_after = 1;
    omp_set_num_threads(c->n_threads);
    char old[4096];
    int rc = -1;
    if (getcwd(old, sizeof(old)) && chdir(dir) == 0) {
        std::vector<const char*> argv;
        argv.push_back("train_ransac");
        for (int i = 0; i < n_args; i++) argv.push_back(args[i]);
        argv.push_back(REF_ARGV_SENTINEL);
        try {
            ref_train_ransac_main((int)argv.size(), argv.data());
            rc = -5;
        } catch (const RefStop&) {
            rc = 0;
        } catch (const std::exception& e) {
            fprintf(stderr, "[ref] train_ransac main threw: %s\n--- its output so far ---\n%s\n", e.what(), sink.str().c_str());
            rc = -6;
        }
        if (chdir(old) != 0) rc = -3;
    }
    std::cout.rdbuf(old_buf);
    if (rc != 0) return rc;
    for (int i = 0; i < ORC_N * 3; i++) dloss[i] = g_env.dloss[i];
    *loss = g_env.loss;
    return 0;
}

}  // extern "C"
