// kernels.cuh -- sm_100a kernels of the forward pass.
//
//   k_sample  (K1)  minimal-set sampling: device MT19937 streams -> candidates -> fp64 P3P ->
//                   reprojection check -> ordered compaction of the first H accepted per
//                   stream                           (cnn_softam.h:1010-1060)
//   k_score   (K2)  fused HxN reprojection-error matrix + soft-inlier score; the last CTA of a
//                   frame finishes softmax / entropy / soft-argmax (K3)
//                                                    (cnn_softam.h:1065-1094, :319-362, :535-553, :80-88)
//
// No tensor cores anywhere: the work is point-wise + reductions.
#pragma once
#include <cuda_runtime.h>

#include "pose_math.cuh"
#include "sampler.cuh"

namespace dsac {

// ------------------------------------------------------------------ block helpers
// Exclusive prefix sum over a 256-thread block (+ block total).  Two barriers.
__device__ __forceinline__ int block_excl_scan_256(int v, int* total, int* s_warp /* >= 8 ints */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += n;
    }
    __syncthreads();  // protect s_warp reuse
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        int x = s_warp[w];
        if (w < warp) base += x;
        tot += x;
    }
    *total = tot;
    return base + incl - v;
}

// ------------------------------------------------------------------ K1: sampling
struct SampleParams {
    const int16_t* coords;   // [n][N][3]
    const int32_t* pix;      // [n or 1][N][2]
    int pix_stride;          // N*2 or 0 (shared grid)
    double f, cx, cy;
    int H, T, thr;
    uint32_t seed, skip;
    int max_candidates;
    long long frame0;
    double* hyp_pose;        // [n][H][6]
    float* hyp_P;            // [n][H][12]  rows (f*R0|f*t0), (f*R1|f*t1), (R2|t2) in float for K2
    int32_t* img_idx;        // [n][H][4]
    int32_t* cand_idx;       // [n][H]
    long long* stream_ncand; // [n][T]
    uint32_t* status;        // [n]
    unsigned long long* n_fragile;  // [1]
};

constexpr int K1_THREADS = 256;
constexpr int K1_RING = 4096;          // words; >= 2048 + margin + 623
constexpr int K1_NEED = 2048 + 128;    // words that must be available past `pos` before a round

__device__ __forceinline__ void mt_twist_block(uint32_t* mt, uint32_t* ring, uint32_t out_base) {
    const int tid = threadIdx.x;
    uint32_t v = 0;
    // phase 1: k in [0,227) depends on old state only
    if (tid < MT_N - MT_M) v = mt_twist(mt[tid], mt[tid + 1], mt[tid + MT_M]);
    __syncthreads();
    if (tid < MT_N - MT_M) {
        mt[tid] = v;
        ring[(out_base + tid) & (K1_RING - 1)] = mt_temper(v);
    }
    __syncthreads();
    // phase 2: k in [227,454) uses the new words [0,227)
    int k = MT_N - MT_M + tid;
    if (tid < MT_N - MT_M) v = mt_twist(mt[k], mt[k + 1], mt[k - (MT_N - MT_M)]);
    __syncthreads();
    if (tid < MT_N - MT_M) {
        mt[k] = v;
        ring[(out_base + k) & (K1_RING - 1)] = mt_temper(v);
    }
    __syncthreads();
    // phase 3: k in [454,624) uses the new words [227,397); the last word wraps to new mt[0]
    k = 2 * (MT_N - MT_M) + tid;
    if (k < MT_N - 1) v = mt_twist(mt[k], mt[k + 1], mt[k - (MT_N - MT_M)]);
    else if (k == MT_N - 1) v = mt_twist(mt[MT_N - 1], mt[0], mt[MT_M - 1]);
    __syncthreads();
    if (k < MT_N) {
        mt[k] = v;
        ring[(out_base + k) & (K1_RING - 1)] = mt_temper(v);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(K1_THREADS) k_sample(SampleParams p) {
    __shared__ uint32_t s_mt[MT_N];
    __shared__ uint32_t s_ring[K1_RING];
    __shared__ int s_warp[8];
    __shared__ uint32_t s_newpos;

    const int tid = threadIdx.x;
    const int s = blockIdx.x, frame = blockIdx.y;
    int h0, quota;
    stream_chunk(p.H, p.T, s, &h0, &quota);
    if (quota == 0) {
        if (tid == 0) p.stream_ncand[(size_t)frame * p.T + s] = 0;
        return;
    }

    // stream s of global frame g: mt19937(seed + g*T + s)   (thread_rand.cpp:52 for g = 0)
    if (tid == 0) mt_seed(s_mt, p.seed + (uint32_t)((p.frame0 + frame) * (long long)p.T + s));
    __syncthreads();

    const int16_t* coords = p.coords + (size_t)frame * DSAC_N_CONST * 3;
    const int32_t* pix = p.pix + (size_t)frame * p.pix_stride;
    WordRing ring{s_ring, K1_RING - 1};

    uint32_t pos = (s == 0) ? p.skip : 0u;  // stream position of the next unread word
    uint32_t gen = 0;                       // words generated so far
    int acc = 0;                            // hypotheses accepted so far
    long long cand_base = 0;                // candidates consumed so far
    const long long cand_max = p.max_candidates > 0 ? (long long)p.max_candidates : (1ll << 40);

    while (acc < quota && cand_base < cand_max) {
        while ((int)(gen - pos) < K1_NEED) {
            mt_twist_block(s_mt, s_ring, gen);
            gen += MT_N;
        }
        // ---- parse up to 256 candidates starting at pos (fixed point over the extras)
        int extra = 0, cells[4];
        uint32_t start, consumed;
        for (;;) {
            int tot;
            int excl = block_excl_scan_256(extra, &tot, s_warp);
            start = pos + 8u * tid + (uint32_t)excl;
            consumed = parse_candidate(ring, start, gen, cells);
            int ne = consumed ? (int)consumed - 8 : 0;
            int changed = (ne != extra);
            extra = ne;
            if (!__syncthreads_or(changed)) break;
        }
        // candidates that did not fit in the generated window (practically never) wait for the next round
        int bad = (consumed == 0) ? tid : K1_THREADS;
#pragma unroll
        for (int off = 16; off; off >>= 1) bad = min(bad, __shfl_xor_sync(0xffffffffu, bad, off));
        __syncthreads();
        if ((tid & 31) == 0) s_warp[tid >> 5] = bad;
        __syncthreads();
        int n_valid = K1_THREADS;
#pragma unroll
        for (int w = 0; w < 8; w++) n_valid = min(n_valid, s_warp[w]);
        if (cand_max - cand_base < n_valid) n_valid = (int)(cand_max - cand_base);
        if (tid == n_valid - 1) s_newpos = start + consumed;
        // (n_valid == 0 cannot happen: K1_NEED words always hold at least one candidate)

        // ---- evaluate: fp64 P3P + reprojection check (cnn_softam.h:1041-1059)
        bool ok = false, fragile = false;
        double rvec[3], tvec[3];
        if (tid < n_valid) {
            float obj[12], img[8];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int c = cells[j];
                img[j * 2] = (float)__ldg(pix + c * 2);
                img[j * 2 + 1] = (float)__ldg(pix + c * 2 + 1);
                obj[j * 3] = (float)__ldg(coords + c * 3);
                obj[j * 3 + 1] = (float)__ldg(coords + c * 3 + 1);
                obj[j * 3 + 2] = (float)__ldg(coords + c * 3 + 2);
            }
            ok = minimal_set_hypothesis(obj, img, p.f, p.cx, p.cy, p.thr, rvec, tvec, &fragile);
        }
        if (fragile) atomicAdd(p.n_fragile, 1ull);

        // ---- ordered compaction: the first `quota` accepted candidates of the stream
        int tot;
        int rank = acc + block_excl_scan_256(ok ? 1 : 0, &tot, s_warp);
        if (ok && rank < quota) {
            size_t hi = (size_t)frame * p.H + h0 + rank;
            double* hp = p.hyp_pose + hi * 6;
            hp[0] = rvec[0]; hp[1] = rvec[1]; hp[2] = rvec[2];
            hp[3] = tvec[0]; hp[4] = tvec[1]; hp[5] = tvec[2];
            double R[9];
            rodrigues_v2m(rvec, R);  // getDiffMap -> cv::projectPoints rebuilds R from rvec
            float4* P = reinterpret_cast<float4*>(p.hyp_P + hi * 12);
            P[0] = make_float4((float)(p.f * R[0]), (float)(p.f * R[1]), (float)(p.f * R[2]), (float)(p.f * tvec[0]));
            P[1] = make_float4((float)(p.f * R[3]), (float)(p.f * R[4]), (float)(p.f * R[5]), (float)(p.f * tvec[1]));
            P[2] = make_float4((float)R[6], (float)R[7], (float)R[8], (float)tvec[2]);
            int4 ii = make_int4(cells[0], cells[1], cells[2], cells[3]);
            *reinterpret_cast<int4*>(p.img_idx + hi * 4) = ii;
            p.cand_idx[hi] = (int32_t)(cand_base + tid);
            if (rank == quota - 1) p.stream_ncand[(size_t)frame * p.T + s] = cand_base + tid + 1;
        }
        acc += tot;
        cand_base += n_valid;
        __syncthreads();
        pos = s_newpos;
    }

    if (acc < quota) {  // sampler exhausted: value-encode like a failed PnP (zero pose, cnn_softam.h:66-71)
        for (int r = acc + tid; r < quota; r += K1_THREADS) {
            size_t hi = (size_t)frame * p.H + h0 + r;
            for (int k = 0; k < 6; k++) p.hyp_pose[hi * 6 + k] = 0.0;
            float4* P = reinterpret_cast<float4*>(p.hyp_P + hi * 12);
            P[0] = make_float4((float)p.f, 0.f, 0.f, 0.f);
            P[1] = make_float4(0.f, (float)p.f, 0.f, 0.f);
            P[2] = make_float4(0.f, 0.f, 1.f, 0.f);
            for (int k = 0; k < 4; k++) p.img_idx[hi * 4 + k] = -1;
            p.cand_idx[hi] = -1;
        }
        if (tid == 0) {
            p.stream_ncand[(size_t)frame * p.T + s] = cand_base;
            atomicOr(p.status + frame, 1u /* DSAC_ST_SAMPLER_EXHAUSTED */);
        }
    }
}

// ------------------------------------------------------------------ K2: fused scoring
struct ScoreParams {
    const int16_t* coords;   // [n][N][3]
    const int32_t* pix;      // [n or 1][N][2]
    int pix_stride;
    const float* hyp_P;      // [n][H][12]
    const double* hyp_pose;  // [n][H][6]
    float* diffmaps;         // [n][H][N] or null
    double* scores;          // [n][H]
    double* sf;              // [n][H]
    double* entropy;         // [n]
    double* avg_pose;        // [n][6]
    unsigned int* frame_counter;  // [n] zero-initialised, self-resetting
    int H, tile, tiles_per_frame;
    float cxf, cyf;
    float thr, kbeta;        // kbeta = beta * log2(e)
    double alpha;
    int external_scores;     // 1: scores come from the score hook, only the softmax tail runs
};

constexpr int K2_THREADS = 320;  // 10 warps x 5 points per thread = 1600 scene coordinates
constexpr int K2_PTS = 5;
constexpr int K2_WARPS = K2_THREADS / 32;
constexpr int K2_MAX_TILE = 256;

__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// softmax / entropy / soft-argmax over the H hypotheses of one frame, in double
// (softMax cnn_softam.h:535-553, entropy :80-88, averaging :1082-1094).
__device__ void softargmax_tail(const ScoreParams& p, int frame, double* s_red /* >= 8*K2_WARPS doubles */) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* sc = p.scores + (size_t)frame * p.H;
    // max
    double m = -1.7976931348623157e308;
    for (int h = tid; h < p.H; h += K2_THREADS) m = fmax(m, __ldcg(sc + h));
#pragma unroll
    for (int off = 16; off; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    m = s_red[0];
    for (int w = 1; w < K2_WARPS; w++) m = fmax(m, s_red[w]);
    __syncthreads();
    // sum of exp and exp-weighted poses
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int h = tid; h < p.H; h += K2_THREADS) {
        double e = exp(__ldcg(sc + h) - m);
        acc[0] += e;
        const double* hp = p.hyp_pose + ((size_t)frame * p.H + h) * 6;
#pragma unroll
        for (int k = 0; k < 6; k++) acc[1 + k] += e * hp[k];
    }
#pragma unroll
    for (int k = 0; k < 7; k++) {
#pragma unroll
        for (int off = 16; off; off >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
        if (lane == 0) s_red[warp * 8 + k] = acc[k];
    }
    __syncthreads();
    double tot[7];
#pragma unroll
    for (int k = 0; k < 7; k++) {
        double t = 0;
        for (int w = 0; w < K2_WARPS; w++) t += s_red[w * 8 + k];
        tot[k] = t;
    }
    __syncthreads();
    // probabilities + entropy
    double ent = 0;
    for (int h = tid; h < p.H; h += K2_THREADS) {
        double pr = exp(__ldcg(sc + h) - m) / tot[0];
        p.sf[(size_t)frame * p.H + h] = pr;
        if (pr > 0) ent -= pr * log2(pr);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) ent += __shfl_xor_sync(0xffffffffu, ent, off);
    if (lane == 0) s_red[warp] = ent;
    __syncthreads();
    if (tid == 0) {
        double e = 0;
        for (int w = 0; w < K2_WARPS; w++) e += s_red[w];
        p.entropy[frame] = e;
        for (int k = 0; k < 6; k++) p.avg_pose[(size_t)frame * 6 + k] = tot[1 + k] / tot[0];
    }
}

template <bool WRITE_DM>
__global__ void __launch_bounds__(K2_THREADS) k_score(ScoreParams p) {
    __shared__ __align__(16) float s_P[K2_MAX_TILE * 12];
    __shared__ float s_part[K2_WARPS][K2_MAX_TILE];
    __shared__ double s_red[8 * K2_WARPS];
    __shared__ unsigned int s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.y, tile_idx = blockIdx.x;
    const int hbeg = tile_idx * p.tile;
    const int nh = min(p.tile, p.H - hbeg);

    if (!p.external_scores) {
        // this thread's 5 scene coordinates and their (pixel - principal point), as floats
        float X[K2_PTS], Y[K2_PTS], Z[K2_PTS], pu[K2_PTS], pv[K2_PTS];
        {
            const int16_t* c = p.coords + (size_t)frame * DSAC_N_CONST * 3;
            const int2* px = reinterpret_cast<const int2*>(p.pix + (size_t)frame * p.pix_stride);
#pragma unroll
            for (int j = 0; j < K2_PTS; j++) {
                int pt = tid + j * K2_THREADS;
                X[j] = (float)__ldg(c + pt * 3);
                Y[j] = (float)__ldg(c + pt * 3 + 1);
                Z[j] = (float)__ldg(c + pt * 3 + 2);
                int2 q = __ldg(px + pt);
                pu[j] = (float)q.x - p.cxf;
                pv[j] = (float)q.y - p.cyf;
            }
        }
        // stage the tile's 3x4 projection rows in shared memory
        {
            const float4* src = reinterpret_cast<const float4*>(p.hyp_P + ((size_t)frame * p.H + hbeg) * 12);
            float4* dst = reinterpret_cast<float4*>(s_P);
            for (int i = tid; i < nh * 3; i += K2_THREADS) dst[i] = __ldg(src + i);
        }
        __syncthreads();

        float* dm = WRITE_DM ? p.diffmaps + ((size_t)frame * p.H + hbeg) * DSAC_N_CONST : nullptr;
        const float tau_k = p.thr * p.kbeta;

        for (int hb = 0; hb < nh; hb += 8) {
            float acc[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int h = hb + u;
                float a = 0.f;
                if (h < nh) {
                    const float4 r0 = *reinterpret_cast<const float4*>(s_P + h * 12);
                    const float4 r1 = *reinterpret_cast<const float4*>(s_P + h * 12 + 4);
                    const float4 r2 = *reinterpret_cast<const float4*>(s_P + h * 12 + 8);
#pragma unroll
                    for (int j = 0; j < K2_PTS; j++) {
                        float xs = fmaf(r0.x, X[j], fmaf(r0.y, Y[j], fmaf(r0.z, Z[j], r0.w)));
                        float ys = fmaf(r1.x, X[j], fmaf(r1.y, Y[j], fmaf(r1.z, Z[j], r1.w)));
                        float zs = fmaf(r2.x, X[j], fmaf(r2.y, Y[j], fmaf(r2.z, Z[j], r2.w)));
                        float iz = (zs != 0.f) ? fast_rcp(zs) : 1.f;  // z ? 1/z : 1 (cv::projectPoints)
                        float du = fmaf(-xs, iz, pu[j]);
                        float dv = fmaf(-ys, iz, pv[j]);
                        float e = fast_sqrt(fmaf(du, du, dv * dv));
                        e = fminf(e, DSAC_MAXINPUT_F);  // min(norm, CNN_OBJ_MAXINPUT), cnn_softam.h:357
                        if (WRITE_DM) __stcs(dm + (size_t)h * DSAC_N_CONST + tid + j * K2_THREADS, e);
                        // sigmoid(beta*(tau-e)) = 1/(1+2^(kbeta*e - kbeta*tau))
                        a += fast_rcp(1.f + fast_ex2(fmaf(p.kbeta, e, -tau_k)));
                    }
                }
                acc[u] = a;
            }
            // transposed warp reduction: 8 partials -> lane (4*b4+2*b3+b2) group holds hypothesis sum
#pragma unroll
            for (int half = 4, off = 16; half >= 1; half >>= 1, off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (i < half) {
                        float send = up ? acc[i] : acc[i + half];
                        float keep = up ? acc[i + half] : acc[i];
                        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
            }
            float v = acc[0];
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if ((lane & 3) == 0) {
                int u = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                if (hb + u < nh) s_part[warp][hb + u] = v;
            }
        }
        __syncthreads();
        // fixed-order cross-warp sum -> score (double from here on, like the reference's vector<double>)
        for (int h = tid; h < nh; h += K2_THREADS) {
            float ssum = 0.f;
#pragma unroll
            for (int w = 0; w < K2_WARPS; w++) ssum += s_part[w][h];
            p.scores[(size_t)frame * p.H + hbeg + h] = p.alpha * (double)ssum;
        }
    }

    // ---- last CTA of the frame runs the softmax / soft-argmax tail
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned int prev = atomicAdd(p.frame_counter + frame, 1u);
        s_last = (prev == (unsigned int)p.tiles_per_frame - 1u);
        if (s_last) p.frame_counter[frame] = 0u;  // self-reset for the next launch
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        softargmax_tail(p, frame, s_red);
    }
}

}  // namespace dsac
