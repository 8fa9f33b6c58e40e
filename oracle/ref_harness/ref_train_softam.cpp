// ref_train_softam.cpp -- oracle/_ref/libref_train_softam.so: the reference's OWN training driver,
// core/train_ransac_softam.cpp (forward + the gradient assembly of lines 288-394), compiled UNMODIFIED with its main()
// renamed, run on a synthetic one-frame dataset.  The final gradient dLoss_dObj is captured where the reference hands
// it to the coordinate CNN (train_ransac_softam.cpp:412 -> lua_calls.h:229), after which the 5000-round loop is left
// by an exception.  TEST INFRASTRUCTURE ONLY.
#include <unistd.h>

#include <iostream>
#include <sstream>

#include "ref_env.h"

#define main ref_train_ransac_softam_main
#include "train_ransac_softam.cpp"
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

// One round of the training loop on the frame in `dir` (./training/<scene>/... with exactly one frame).
// Returns 0 and fills dloss [1600*3], loss and the score output-gradients [H] (as handed to the score backward).
int ref_train_softam_round(const ref_config* c, const char* dir, const int16_t* coords, int n_args, const char** args, double* dloss,
                           double* loss, double* score_out_grads) {
    std::streambuf* old_buf = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    g_env.coords = coords;
    g_env.n_frames = 1;
    g_env.frame0 = c->frame;
    g_env.alpha = c->alpha; g_env.beta = c->beta; g_env.grad_clamp = c->grad_clamp;
    g_env.thr = c->thr2d;
    g_env.seed = c->seed;
    g_env.T = c->n_threads;
    g_env.backward_calls = 0;
    g_env.stop_after = 1;
    omp_set_num_threads(c->n_threads);
    char old[4096];
    int rc = -1;
    if (getcwd(old, sizeof(old)) && chdir(dir) == 0) {
        std::vector<const char*> argv;
        argv.push_back("train_ransac_softam");
        for (int i = 0; i < n_args; i++) argv.push_back(args[i]);
    argv.push_back(REF_ARGV_SENTINEL);
        argv.push_back(REF_ARGV_SENTINEL);
        try {
            ref_train_ransac_softam_main((int)argv.size(), argv.data());
            rc = -5;   // the loop ended without a backward call
        } catch (const RefStop&) {
            rc = 0;
        } catch (const std::exception& e) {
            fprintf(stderr, "[ref] train main threw: %s\n--- its output so far ---\n%s\n", e.what(), sink.str().c_str());
            rc = -6;
        }
        if (chdir(old) != 0) rc = -3;
    }
    std::cout.rdbuf(old_buf);
    if (rc != 0) return rc;
    for (int i = 0; i < ORC_N * 3; i++) dloss[i] = g_env.dloss[i];
    *loss = g_env.loss;
    if (score_out_grads)
        for (int h = 0; h < c->n_hyps && h < (int)g_env.score_out_grads.size(); h++) score_out_grads[h] = g_env.score_out_grads[h];
    return 0;
}

}  // extern "C"
