// ref_env.h -- the environment the reference's own sources run in when they are compiled into oracle/_ref
// (TEST INFRASTRUCTURE; see oracle/shim/opencv2/opencv.hpp for the why).
//
// The reference's pipeline talks to two things that do not exist here: its CNNs (through the Lua C API,
// core/lua_calls.h) and a dataset on disk (png++ frames + 7-Scenes pose files, core/dataset.h, core/read_data.cpp).
// This header is the other end of both, included by every oracle/ref_harness/*.cpp translation unit:
//
//   * coordinate CNN  forward(count, patches) -> the synthetic frame's int16 scene coordinates / 1000 (metres), so that
//                     cnn_softam.h:265 (prediction * 1000, saturating) reproduces the int16 grid exactly;
//                     backward(count, loss, patches, dLoss) -> captured: dLoss IS the pipeline's final gradient
//                     (train_ransac_softam.cpp:412);
//   * score CNN       forward / backward = the closed-form soft-inlier score of north_star and its derivative, with the
//                     Lua script's own conventions (output-gradient clamp train_score_softam.lua:97; gradients returned
//                     x-major, train_score_softam.lua:122-131);
//   * png::image      a black frame; loading frame g re-seeds the reference's ThreadRand with seed + g*T
//                     (ThreadRand::forceInit), which is the engine's stream contract.
#pragma once
#include <omp.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../dsac_oracle.h"
#include "thread_rand.h"   // the reference's own (core/thread_rand.h)
#include <lua.hpp>

// GlobalProperties::readArguments (properties.cpp:97-268, a bool function) has no return statement after its loop; GCC >= 8
// plants a trap there.  The drivers' command lines therefore always end with one unknown flag: the loop then leaves
// through its own "unkown argument ... return false" (properties.cpp:265-266), whose value parseCmdLine ignores.
#define REF_ARGV_SENTINEL "--end-of-arguments"

struct RefStop {};   // thrown out of the reference's training loop once enough rounds were captured

struct RefEnv {
    // ---- inputs
    const int16_t* coords = nullptr;   // [n][1600][3]
    int n_frames = 0;
    long long frame0 = 0;              // global index of frame 0 (stream key)
    int cur = 0;                       // frame being processed
    double alpha = 0.1, beta = 0.5, grad_clamp = 0.1;
    int thr = 10;
    uint32_t seed = 1305;
    int T = 1;
    int extra_draws = 0;               // thread-0 draws consumed after the re-seed and before stochasticSubSample
    // ---- captures
    std::vector<float> diffmaps;       // score forward: [H][1600]
    std::vector<double> scores;        // [H]
    std::vector<double> score_out_grads;   // score backward: [H] (before the clamp)
    std::vector<double> dloss;         // coordinate backward: [1600*3] row-major (rows = cells)
    double loss = 0;
    int backward_calls = 0, stop_after = 1;
    int loads = 0;
};
static RefEnv g_env;

static inline bool ref_is_score_state(const lua_State* L) { return L->script.find("score") != std::string::npos; }

static inline void ref_reseed_for_frame(int frame) {
    omp_set_num_threads(g_env.T);
    ThreadRand::forceInit(g_env.seed + (uint32_t)((g_env.frame0 + frame) * (long long)g_env.T));
}

extern "C" void shim_png_on_load(const char* path, int* width, int* height) {
    // frame index = the digits of "frame-%06d.color.png" (7-Scenes naming)
    const char* base = strrchr(path, '/');
    base = base ? base + 1 : path;
    int idx = 0;
    const char* d = base;
    while (*d && (*d < '0' || *d > '9')) d++;
    idx = atoi(d);
    g_env.cur = idx;
    g_env.loads++;
    *width = 640; *height = 480;
    ref_reseed_for_frame(idx);
}

void shim_lua_dispatch(lua_State* L, const std::string& fn, std::vector<shim_lua_value>& args, int /*nresults*/,
                       std::vector<shim_lua_value>& results) {
    if (getenv("REF_DEBUG")) fprintf(stderr, "[ref] lua %s(%zu args) on %s\n", fn.c_str(), args.size(), L->script.c_str());
    auto number = [](double v) { shim_lua_value x; x.kind = shim_lua_value::NUMBER; x.num = v; return x; };
    if (fn == "loadModel" || fn == "constructModel" || fn == "setEvaluate" || fn == "setTraining") return;
    const int N = ORC_N;
    if (ref_is_score_state(L)) {
        if (fn == "forward") {           // lua_calls.h:284-300: (count, maps pushed n -> y -> x) -> count numbers
            const int count = (int)args[0].num;
            const std::vector<double>& maps = *args[1].tab;
            g_env.diffmaps.resize((size_t)count * N);
            g_env.scores.resize(count);
            for (size_t i = 0; i < (size_t)count * N; i++) g_env.diffmaps[i] = (float)maps[i];
            for (int h = 0; h < count; h++) {
                g_env.scores[h] = orc_soft_inlier_score(&g_env.diffmaps[(size_t)h * N], N, (double)g_env.thr, g_env.alpha, g_env.beta);
                results.push_back(number(g_env.scores[h]));
            }
        } else if (fn == "backward") {   // lua_calls.h:312-341: (count, maps, outputGradients) -> one table
            const int count = (int)args[0].num;
            const std::vector<double>& maps = *args[1].tab;
            const std::vector<double>& og = *args[2].tab;
            g_env.score_out_grads.assign(og.begin(), og.begin() + count);
            shim_lua_value t;
            t.kind = shim_lua_value::TABLE;
            t.tab = std::make_shared<std::vector<double>>((size_t)count * N, 0.0);
            for (int c = 0; c < count; c++) {
                double go = og[c];
                if (g_env.grad_clamp > 0) go = std::max(-g_env.grad_clamp, std::min(g_env.grad_clamp, go));   // train_score_softam.lua:97
                for (int y = 0; y < ORC_GRID; y++)
                    for (int x = 0; x < ORC_GRID; x++) {
                        const double e = (double)(float)maps[(size_t)c * N + y * ORC_GRID + x];
                        const double sg = 1.0 / (1.0 + std::exp(-g_env.beta * ((double)g_env.thr - e)));
                        const double g = go * (-g_env.alpha * g_env.beta * sg * (1.0 - sg));
                        (*t.tab)[(size_t)c * N + x * ORC_GRID + y] = g;   // x-major, train_score_softam.lua:122-131
                    }
            }
            results.push_back(t);
        }
        return;
    }
    // coordinate CNN
    if (fn == "forward") {               // lua_calls.h:252-273: (count, patches) -> table of count*3 numbers (metres)
        const int count = (int)args[0].num;
        shim_lua_value t;
        t.kind = shim_lua_value::TABLE;
        t.tab = std::make_shared<std::vector<double>>((size_t)count * 3, 0.0);
        const int16_t* c = g_env.coords + (size_t)g_env.cur * N * 3;
        for (int i = 0; i < count * 3 && i < N * 3; i++) (*t.tab)[i] = (double)c[i] / 1000.0;
        results.push_back(t);
    } else if (fn == "backward") {       // lua_calls.h:229-241: (rows, loss, patches, dLoss pushed row-major)
        g_env.loss = args[1].num;
        g_env.dloss = *args[3].tab;
        g_env.backward_calls++;
        if (g_env.backward_calls >= g_env.stop_after) throw RefStop();
    }
}
