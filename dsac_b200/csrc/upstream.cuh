// upstream.cuh -- the step immediately before the hypothesis engine, kept on the
// device so that frames never leave HBM:
//   * k_gather_patches: getCoordImg's patch assembly (cnn_softam.h:221-256) fused with the normalisation of the
//     coordinate CNN's forward() (lua/train_obj.lua:117-124) and with pushMaps' channel-major order
//     (lua_calls.h:65-82): patch(c, y, x) = frame(oy - 21 + y, ox - 21 + x)[c] - mean
//   * k_coords_from_prediction: modeImg(y, x) = prediction * 1000 (cnn_softam.h:262-268), Vec3f -> Vec<short,3>
//     with cv::saturate_cast<short> (round half to even, saturate)
// Both are pure streaming kernels: the HBM write of the patches (21 168 B each) is the roofline.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dsac {

constexpr int UP_PATCH = 42;                                   // CNN_RGB_PATCHSIZE, lua_calls.h:30
constexpr int UP_PLANE = UP_PATCH * UP_PATCH;                  // 1764
constexpr int UP_ELEMS = 3 * UP_PLANE;                         // 5292 floats per patch
constexpr int UP_ROWB = 3 * UP_PATCH;                          // 126 bytes of one patch row in the BGR frame
constexpr int UP_THREADS = 128;

struct GatherParams {
    const uint8_t* frames;    // [n][height][width][3] BGR (jp::img_bgr_t)
    int width, height;
    const int32_t* pix;       // [n or 1][N][2] sampling(y, x) = (origX, origY)
    int pix_stride;           // N*2 or 0 (shared grid)
    int n_cells;
    float mean;
    float* patches;           // [n][N][3][42][42]
    uint32_t* status;         // [n] or null: DSAC_ST_BORDER_PATCH
};

// One CTA per patch.  Phase 1: thread t < 126 owns byte column t = 3*x + c of the patch's 42 frame rows (a warp reads 32
// consecutive bytes per row), converts and writes the normalised value to its channel-major place in shared memory.
// Phase 2: the 21 168 B patch leaves as 1323 coalesced 16-byte streaming stores.  No index arithmetic in either loop.
__global__ void __launch_bounds__(UP_THREADS) k_gather_patches(GatherParams p) {
    __shared__ __align__(16) float sf[UP_ELEMS];
    const int cell = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x;
    const int32_t* px = p.pix + (size_t)frame * p.pix_stride + cell * 2;
    const int ox = px[0], oy = px[1];
    const int half = UP_PATCH / 2;
    // "skip border patches" (cnn_softam.h:236-240); stochasticSubSample never produces one
    const bool border = (ox < half) || (oy < half) || (ox > p.width - half) || (oy > p.height - half);
    float4* out = reinterpret_cast<float4*>(p.patches + ((size_t)frame * p.n_cells + cell) * UP_ELEMS);
    if (border) {
        if (tid == 0 && p.status) atomicOr(p.status + frame, 4u /* DSAC_ST_BORDER_PATCH */);
        for (int q = tid; q < UP_ELEMS / 4; q += UP_THREADS) __stcs(out + q, make_float4(0.f, 0.f, 0.f, 0.f));
        return;
    }
    if (tid < UP_ROWB) {
        const int x = tid / 3, c = tid - 3 * x;
        const size_t row_bytes = (size_t)p.width * 3;
        const uint8_t* src = p.frames + ((size_t)frame * p.height + (oy - half)) * row_bytes + (size_t)(ox - half) * 3 + tid;
        float* dst = sf + c * UP_PLANE + x;
#pragma unroll 14
        for (int r = 0; r < UP_PATCH; r++) dst[r * UP_PATCH] = (float)__ldg(src + r * row_bytes) - p.mean;
    }
    __syncthreads();
    const float4* sf4 = reinterpret_cast<const float4*>(sf);
#pragma unroll 4
    for (int q = tid; q < UP_ELEMS / 4; q += UP_THREADS) __stcs(out + q, sf4[q]);
}

// prediction (metres) -> int16 millimetres; cv::Vec3f * 1000 is a float product, saturate_cast<short>(float) = cvRound + clamp
__global__ void k_coords_from_prediction(const float* __restrict__ pred, int16_t* __restrict__ coords, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float v = __fmul_rn(pred[i], 1000.f);
    int r = __float2int_rn(v);            // round half to even, like cvRound (lrint)
    if (!(fabsf(v) < 2147483648.f)) r = -2147483647 - 1;   // cvRound = cvtss2si on x86-64: NaN / out of int range -> INT_MIN, i.e. -32768 after saturation
    r = max(-32768, min(32767, r));
    coords[i] = (int16_t)r;
}

}  // namespace dsac
