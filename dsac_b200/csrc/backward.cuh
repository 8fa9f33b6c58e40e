// backward.cuh -- K5/K6: backward of the soft-argmax pipeline (train_ransac_softam.cpp:288-394).
#pragma once
#include <cuda_runtime.h>

#include "pose_math.cuh"

struct dsac_engine;
struct dsac_backward_out;

namespace dsac {
struct BackwardScratch {
    void* buf = nullptr;
    size_t bytes = 0;
};
inline void backward_scratch_free(BackwardScratch* s) {
    if (s->buf) cudaFree(s->buf);
    s->buf = nullptr;
    s->bytes = 0;
}
}  // namespace dsac

int backward_run(dsac_engine* e, int32_t n, const int16_t* coords, const int32_t* pix, int32_t pix_shared,
                 const double* gt_jp, dsac_backward_out* out);
int kabsch_run(dsac_engine* e, int32_t n, int32_t m, const double* a, const double* b, double* R, double* t);
