// Development aid: cycles per MT19937 block regeneration of a lone CTA under different arrangements.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../dsac_b200/csrc -o regen_bench regen_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define DSAC_BUILD 1
#include "sampler.cuh"
using namespace dsac;

template <int NT, int MODE>
__global__ void __launch_bounds__(NT) k(int iters, long long* out, uint32_t* sink) {
    __shared__ uint32_t st[2 * MT_N];
    __shared__ unsigned char vals[32768];
    const int tid = threadIdx.x;
    for (int k = tid; k < MT_N; k += NT) st[k] = 1812433253u * (uint32_t)k + 12345u;
    __syncthreads();
    uint32_t par = 0;
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < iters; it++) {
        const uint32_t* so = st + par * MT_N;
        uint32_t* sn = st + (par ^ 1u) * MT_N;
        par ^= 1u;
        if (MODE == 0) {          // fused: 227 threads twist 3 words and decode them
            if (tid < 227) {
                uint32_t x[3];
                const bool has3 = mt_regenerate_words(so, tid, x) == 3;
#pragma unroll
                for (int w = 0; w < 3; w++) if (w < 2 || has3) {
                    const int kk = tid + w * 227;
                    sn[kk] = x[w];
                    const uint32_t tv = mt_temper(x[w]);
                    vals[(it & 31) * 624 + kk] = (unsigned char)__umulhi(tv, 40u);
                }
            }
        } else if (MODE == 1) {   // twist only
            if (tid < 227) {
                uint32_t x[3];
                const bool has3 = mt_regenerate_words(so, tid, x) == 3;
                sn[tid] = x[0]; sn[tid + 227] = x[1]; if (has3) sn[tid + 454] = x[2];
            }
        } else if (MODE == 2) {   // twist by the first 256 threads, decode of the previous block by the others
            if (tid < 256) {
                if (tid < 227) {
                    uint32_t x[3];
                    const bool has3 = mt_regenerate_words(so, tid, x) == 3;
                    sn[tid] = x[0]; sn[tid + 227] = x[1]; if (has3) sn[tid + 454] = x[2];
                }
            } else {
                for (int kk = tid - 256; kk < MT_N; kk += NT - 256) {
                    const uint32_t tv = mt_temper(so[kk]);
                    vals[(it & 31) * 624 + kk] = (unsigned char)__umulhi(tv, 40u);
                }
            }
        } else if (MODE == 3) {   // barrier only
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (tid == 0) { out[0] = t1 - t0; }
    acc = st[tid % 624] + vals[tid];
    sink[tid] = acc;
}

template <int NT, int MODE>
void run(const char* name) {
    long long* d; uint32_t* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4096 * 4);
    const int iters = 2000;
    k<NT, MODE><<<1, NT>>>(iters, d, s); cudaDeviceSynchronize();
    k<NT, MODE><<<1, NT>>>(iters, d, s); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("NT=%4d %-40s %7.1f cycles / block  (%s)\n", NT, name, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d); cudaFree(s);
}

int main() {
    run<256, 0>("fused twist+decode (227 thr)");
    run<256, 1>("twist only");
    run<256, 3>("barrier only");
    run<512, 2>("twist 256 | decode 256");
    run<1024, 0>("fused twist+decode (227 thr)");
    run<1024, 1>("twist only");
    run<1024, 2>("twist 256 | decode 768");
    run<1024, 3>("barrier only");
    return 0;
}
