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

// One CTA (4 warps) per patch.
// Phase 1: a warp takes every 4th of the patch's 42 frame rows.  A row is 126 consecutive bytes of the BGR frame at an
// arbitrary byte offset: lane l loads the aligned 32-bit word l of the row (one coalesced 128-byte request per row; all of a
// warp's rows are requested before the first is used), takes its neighbour's word by shuffle and funnel-shifts the four
// bytes 4l .. 4l+3 of the row into place.  byte -> float without the conversion unit: 0x4B000000 | b is the float
// 2^23 + b, and (2^23 + b) - 2^23 = (float)b exactly; then "- mean" as the reference subtracts it.  The four values go to
// their channel-major places in shared memory (offsets fixed per lane).
// Phase 2: the finished 21 168-byte patch leaves shared memory as ONE bulk asynchronous copy (TMA, cp.async.bulk
// shared -> global; SASS UBLKCP): no per-thread load / store instructions for the 2.1 GB this kernel writes per 64 frames.
__device__ __forceinline__ void up_bulk_store(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(src_smem);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the generic-proxy writes of phase 1 (ordered by the barrier) before the async read
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory may be released (CTA exit) once it has been read
}

constexpr int UP_ROWS_PER_WARP = (UP_PATCH + UP_THREADS / 32 - 1) / (UP_THREADS / 32);   // 11

__global__ void __launch_bounds__(UP_THREADS) k_gather_patches(GatherParams p) {
    __shared__ __align__(128) float sf[UP_ELEMS];
    const int cell = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t* px = p.pix + (size_t)frame * p.pix_stride + cell * 2;
    const int ox = px[0], oy = px[1];
    const int half = UP_PATCH / 2;
    // "skip border patches" (cnn_softam.h:236-240); stochasticSubSample never produces one
    const bool border = (ox < half) || (oy < half) || (ox > p.width - half) || (oy > p.height - half);
    float* out = p.patches + ((size_t)frame * p.n_cells + cell) * UP_ELEMS;
    if (border) {
        if (tid == 0 && p.status) atomicOr(p.status + frame, 4u /* DSAC_ST_BORDER_PATCH */);
        float4* out4 = reinterpret_cast<float4*>(out);
        for (int q = tid; q < UP_ELEMS / 4; q += UP_THREADS) __stcs(out4 + q, make_float4(0.f, 0.f, 0.f, 0.f));
        return;
    }
    const size_t row_bytes = (size_t)p.width * 3;
    const uint8_t* row0 = p.frames + ((size_t)frame * p.height + (oy - half)) * row_bytes + (size_t)(ox - half) * 3;
    // this lane's four elements e = 4 lane + j of every row: column x = e / 3, channel c = e % 3 (lane 31: only e = 124, 125)
    int off[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int e = 4 * lane + j, x = e / 3, c = e - 3 * x;
        off[j] = c * UP_PLANE + x;
    }
    uint32_t w0[UP_ROWS_PER_WARP], w1x[UP_ROWS_PER_WARP];
    uint32_t sh8[UP_ROWS_PER_WARP];
#pragma unroll
    for (int k = 0; k < UP_ROWS_PER_WARP; k++) {
        const int r = warp + k * (UP_THREADS / 32);
        w0[k] = 0u; w1x[k] = 0u; sh8[k] = 0u;
        if (r < UP_PATCH) {
            const uint8_t* src = row0 + (size_t)r * row_bytes;
            const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
            const uint8_t* base = src - sh;
            w0[k] = __ldg(reinterpret_cast<const uint32_t*>(base) + lane);
            // the word after lane 31's: only its first byte can belong to the row (offset 3), fetched as a byte so that
            // nothing is read beyond the row's last pixel
            if (lane == 31 && sh == 3u) w1x[k] = (uint32_t)__ldg(base + 128);
            sh8[k] = sh * 8u;
        }
    }
    const float magic = 8388608.f;
#pragma unroll
    for (int k = 0; k < UP_ROWS_PER_WARP; k++) {
        const int r = warp + k * (UP_THREADS / 32);
        uint32_t nb = __shfl_down_sync(0xffffffffu, w0[k], 1);
        if (lane == 31) nb = w1x[k];
        if (r < UP_PATCH) {
            const uint32_t v = __funnelshift_r(w0[k], nb, sh8[k]);
            float* dst = sf + r * UP_PATCH;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (j < 2 || lane < 31) {
                    const float fb = __uint_as_float(0x4B000000u | ((v >> (8 * j)) & 0xffu)) - magic;   // == (float)byte
                    dst[off[j]] = fb - p.mean;
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) up_bulk_store(out, sf, (uint32_t)(UP_ELEMS * sizeof(float)));
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
