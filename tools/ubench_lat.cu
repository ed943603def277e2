// latency micro-benchmarks (single warp, dependent chains): DFMA, SHFL+DADD, rsqrt, sqrt, div, LDS->DFMA
#include <cuda_runtime.h>
#include <cstdio>
__global__ void lat(double* out, long long* cyc, int iters) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = 1.0 + 1e-9 * i;
    __syncwarp();
    double x = 1.0 + threadIdx.x * 1e-6, y = 0.999999;
    long long t0, t1;
    t0 = clock64(); for (int i = 0; i < iters; ++i) x = fma(x, y, 1e-9); t1 = clock64(); if (!threadIdx.x) cyc[0] = t1 - t0;
    t0 = clock64(); for (int i = 0; i < iters; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1); t1 = clock64(); if (!threadIdx.x) cyc[1] = t1 - t0;
    x = fabs(x) * 1e-300 + 2.0;
    t0 = clock64(); for (int i = 0; i < iters; ++i) x = rsqrt(x) + 1.5; t1 = clock64(); if (!threadIdx.x) cyc[2] = t1 - t0;
    t0 = clock64(); for (int i = 0; i < iters; ++i) x = sqrt(x) + 1.5; t1 = clock64(); if (!threadIdx.x) cyc[3] = t1 - t0;
    t0 = clock64(); for (int i = 0; i < iters; ++i) x = 1.0 / x + 1.5; t1 = clock64(); if (!threadIdx.x) cyc[4] = t1 - t0;
    int idx = threadIdx.x;
    t0 = clock64(); for (int i = 0; i < iters; ++i) { x = fma(sm[idx], y, x); idx = (idx + 33) & 1023; } t1 = clock64(); if (!threadIdx.x) cyc[5] = t1 - t0;
    float f = (float)x;
    t0 = clock64(); for (int i = 0; i < iters; ++i) f = fmaf(f, 0.999f, 1e-3f); t1 = clock64(); if (!threadIdx.x) cyc[6] = t1 - t0;
    t0 = clock64(); for (int i = 0; i < iters; ++i) { __syncthreads(); } t1 = clock64(); if (!threadIdx.x) cyc[7] = t1 - t0;
    out[threadIdx.x] = x + f;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 32); cudaMalloc(&cyc, 8 * 8);
    const int iters = 4096;
    for (int rep = 0; rep < 2; ++rep) lat<<<1, 32>>>(out, cyc, iters);
    long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    const char* nm[8] = {"DFMA dependent", "SHFL+DADD dependent", "rsqrt(double)+add", "sqrt(double)+add", "1/x+add", "LDS->DFMA (idx chain)", "FFMA dependent", "__syncthreads (1 warp)"};
    for (int i = 0; i < 8; ++i) printf("%-26s %.1f cycles\n", nm[i], (double)h[i] / iters);
    return 0;
}
