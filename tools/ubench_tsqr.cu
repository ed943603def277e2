// ubench_tsqr.cu -- cycle-level look at the warp-synchronous TSQR blocks (qil_wqr.cuh):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I include -o tools/bin/ubench_tsqr tools/ubench_tsqr.cu
#define QIL_WQR_PROFILE 1
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__device__ long long g_clk[256];
#include "../qilaplace.jl_b200/csrc/qil_wqr.cuh"
using namespace qil;

template <int RPL, int THREADS>
__global__ void __launch_bounds__(THREADS) k_factor(double* out, int m, int n, int pitch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* blk = reinterpret_cast<double*>(smem_raw);
    double* beta = blk + (size_t)(32 * RPL) * pitch;
    double* tau = beta + n;
    for (int idx = threadIdx.x; idx < m * pitch; idx += blockDim.x) {
        const int i = idx / pitch, c = idx % pitch;
        blk[idx] = c < n ? sin(0.37 * i + 1.3 * c) + 0.01 * cos(0.11 * i * c) : 0.0;
    }
    __syncthreads();
    const long long t0 = clock64();
    wqr_factor<double, RPL, (THREADS > 256 ? 32 : (RPL >= 16 ? 64 : 16))>(blk, pitch, m, n, beta, tau, threadIdx.x >> 5, blockDim.x >> 5, 1);
    const long long t1 = clock64();
    if (threadIdx.x == 0) { g_clk[200] = t1 - t0; }
    __syncthreads();
    constexpr int CH = (THREADS > 256 && RPL >= 16) ? 2 : 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double reg[RPL][CH];
    for (int t = 0; t < RPL; ++t)
        for (int q = 0; q < CH; ++q) reg[t][q] = (lane + 32 * t == warp * CH + q) ? 1.0 : 0.0;
    const long long t2 = clock64();
    if (warp * CH < n) wqr_apply_chunk<double, RPL, CH>(blk, pitch, m, min(m, n), tau, reg);
    const long long t3 = clock64();
    if (threadIdx.x == 0) { g_clk[201] = t3 - t2; }
    double acc = 0;
    for (int t = 0; t < RPL; ++t)
        for (int q = 0; q < CH; ++q) acc += reg[t][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + beta[0];
}

template <int RPL, int THREADS = 256>
void run(int m, int n, int threads) {
    const int pitch = wqr_pitch(n);
    double* out;
    cudaMalloc(&out, 1 << 20);
    size_t smem = ((size_t)32 * RPL * pitch + 4 * n + 16) * 8;
    cudaFuncSetAttribute(k_factor<RPL, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_factor<RPL, THREADS><<<1, threads, smem>>>(out, m, n, pitch);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256];
    cudaMemcpyFromSymbol(h, g_clk, sizeof(h));
    printf("RPL=%d m=%d n=%d threads=%d: kernel %.1f us, factor %lld cyc (%.0f / step), apply-chunk %lld cyc (%.0f / step); err=%s\n", RPL, m, n,
           threads, ms * 1e3, h[200], (double)h[200] / n, h[201], (double)h[201] / n, cudaGetErrorString(cudaGetLastError()));
    printf("   per-step cycles of warp 0:");
    for (int j = 0; j < n; ++j) printf(" %lld", h[j]);
    printf("\n   breakdown step 0 (load u | dots | reduce | scalars | f | update | barrier): %lld %lld %lld %lld %lld %lld %lld\n", h[100], h[101],
           h[102], h[103], h[104], h[105], h[106]);
    cudaFree(out);
}

int main() {
    run<16>(512, 20, 256);
    run<16, 512>(512, 20, 512);
    run<8, 512>(256, 20, 512);
    run<4, 512>(128, 20, 512);
    run<16>(512, 20, 32);
    run<8>(256, 20, 256);
    run<8>(213, 20, 160);
    run<8>(256, 8, 256);
    run<8>(64, 20, 256);
    return 0;
}
