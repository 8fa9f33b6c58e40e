// Development aid: cycles per MT19937 twist-only block (skip-ahead) of a lone CTA, variants.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../dsac_b200/csrc -o twist_bench twist_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define DSAC_BUILD 1
#include "sampler.cuh"
using namespace dsac;

__device__ __forceinline__ uint32_t tw(uint32_t cur, uint32_t nxt, uint32_t far) {
    const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((0u - (nxt & 1u)) & 0x9908b0dfu);
}

// MODE 0: mt_regenerate_words (as shipped).  MODE 1: branch-free, all loads first (thread 169's extra word through a select).
// MODE 2: MODE 1 on 227 threads of a 256-thread CTA but the state kept in registers between blocks where possible (x[0..2] are
// the thread's own words of the new state: the next block's `cur` for stage s is the thread's own x[s]).
template <int NT, int MODE>
__global__ void __launch_bounds__(NT) k(int iters, long long* out, uint32_t* sink) {
    __shared__ uint32_t st[2 * 640];
    const int tid = threadIdx.x;
    for (int k = tid; k < 624; k += NT) st[k] = 1812433253u * (uint32_t)k + 12345u;
    __syncthreads();
    uint32_t par = 0;
    uint32_t own0 = 0, own1 = 0, own2 = 0;
    if (tid < 227) { own0 = st[tid]; own1 = st[tid + 227]; own2 = (tid < 170) ? st[tid + 454] : 0u; }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const uint32_t* so = st + par * 640;
        uint32_t* sn = st + (par ^ 1u) * 640;
        par ^= 1u;
        if (MODE == 0) {
            if (tid < 227) {
                uint32_t x[3];
                const bool has3 = mt_regenerate_words(so, tid, x) == 3;
                sn[tid] = x[0]; sn[tid + 227] = x[1]; if (has3) sn[tid + 454] = x[2];
            }
        } else if (MODE == 1) {
            if (tid < 227) {
                const int t = tid, k3 = min(t + 454, 623);
                const uint32_t a0 = so[t], a1 = so[t + 1], af = so[t + 397];
                const uint32_t b0 = so[t + 227], b1 = so[t + 228];
                const uint32_t c0 = so[k3], c1 = so[min(k3 + 1, 623)];
                const uint32_t z0 = so[0], z1 = so[1], zf = so[397];
                const uint32_t n0 = tw(z0, z1, zf);
                const uint32_t x0 = tw(a0, a1, af);
                const uint32_t x1 = tw(b0, b1, x0);
                const uint32_t x2 = tw(c0, (k3 == 623) ? n0 : c1, x1);
                sn[t] = x0; sn[t + 227] = x1; if (t < 170) sn[t + 454] = x2;
            }
        } else if (MODE == 2) {
            if (tid < 227) {
                const int t = tid, k3 = min(t + 454, 623);
                // own words of the old state are in registers; neighbours' from shared memory
                const uint32_t a1 = so[t + 1], af = so[t + 397];
                const uint32_t b1 = so[t + 228];
                const uint32_t c1 = so[min(k3 + 1, 623)];
                const uint32_t z0 = so[0], z1 = so[1], zf = so[397];
                const uint32_t n0 = tw(z0, z1, zf);
                const uint32_t x0 = tw(own0, a1, af);
                const uint32_t x1 = tw(own1, b1, x0);
                const uint32_t x2 = tw(own2, (k3 == 623) ? n0 : c1, x1);
                sn[t] = x0; sn[t + 227] = x1; if (t < 170) sn[t + 454] = x2;
                own0 = x0; own1 = x1; own2 = x2;
            }
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (tid == 0) out[0] = t1 - t0;
    sink[tid] = st[par * 640 + (tid % 624)] + own0;
}

template <int NT, int MODE>
uint32_t run(const char* name) {
    long long* d; uint32_t* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4096 * 4);
    const int iters = 2000;
    k<NT, MODE><<<1, NT>>>(iters, d, s); cudaDeviceSynchronize();
    k<NT, MODE><<<1, NT>>>(iters, d, s); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    uint32_t hs[256]; cudaMemcpy(hs, s, sizeof(hs), cudaMemcpyDeviceToHost);
    uint32_t chk = 0; for (int i = 0; i < 227; i++) chk = chk * 31u + hs[i];
    printf("NT=%4d %-46s %7.1f cycles / block  checksum %08x (%s)\n", NT, name, (double)h / iters, chk, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d); cudaFree(s);
    return chk;
}

int main() {
    run<256, 0>("mt_regenerate_words (shipped)");
    run<256, 1>("branch-free, loads first");
    run<256, 2>("own words in registers");
    return 0;
}
