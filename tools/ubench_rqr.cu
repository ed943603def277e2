// ubench_rqr.cu -- where a column step of the register-resident Householder (rqr_factor, qil_wqr.cuh) spends its cycles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I include -o tools/bin/ubench_rqr tools/ubench_rqr.cu
#define QIL_RQR_PROFILE 1
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__device__ long long g_rqr_seg[8];
__device__ long long g_total[2];
#include "../qilaplace.jl_b200/csrc/qil_wqr.cuh"
using namespace qil;

template <int NC, int RT>
__global__ void __launch_bounds__(256) k_factor(double* out, int m, int n, int pitch, int NW) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* blk = reinterpret_cast<double*>(smem_raw);
    double* beta = blk + (size_t)m * pitch;
    double* tau = beta + 32;
    double* scr = tau + 32;
    for (int idx = threadIdx.x; idx < m * pitch; idx += blockDim.x) {
        const int i = idx / pitch, c = idx % pitch;
        blk[idx] = c < n ? sin(0.37 * i + 1.3 * c) + 0.01 * cos(0.11 * i * c) : 0.0;
    }
    __syncthreads();
    const long long t0 = clock64();
    if ((int)(threadIdx.x >> 5) < NW) rqr_factor<NC, RT>(blk, pitch, m, n, beta, tau, scr, NW, 1);
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) { g_total[0] = t1 - t0; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = blk[(threadIdx.x % m) * pitch] + beta[0];
}

template <int NC, int RT>
void run(int m, int n, int NW) {
    const int pitch = wqr_pitch(n);
    double* out;
    cudaMalloc(&out, 1 << 20);
    size_t smem = ((size_t)m * pitch + 64 + rqr_scratch_elems(NW, NC) + 8) * 8;
    cudaFuncSetAttribute(k_factor<NC, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 3; ++rep) {
        k_factor<NC, RT><<<1, 256, smem>>>(out, m, n, pitch, NW);
        cudaDeviceSynchronize();
    }
    long long h[8], tot[2];
    cudaMemcpyFromSymbol(h, g_rqr_seg, sizeof(h));
    cudaMemcpyFromSymbol(tot, g_total, sizeof(tot));
    printf("NC=%d RT=%d m=%d n=%d NW=%d: total %lld cyc (%.0f / step); per step: products+stage %.0f | colsum %.0f | barrier %.0f | totals+shfl %.0f | reflector %.0f | f+syncwarp %.0f | update %.0f ; err=%s\n",
           NC, RT, m, n, NW, tot[0], (double)tot[0] / n, (double)h[0] / n, (double)h[1] / n, (double)h[2] / n, (double)h[3] / n,
           (double)h[4] / n, (double)h[5] / n, (double)h[6] / n, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    run<24, 1>(128, 20, 4);
    run<24, 2>(128, 20, 2);
    run<24, 4>(128, 20, 1);
    run<24, 1>(200, 20, 7);
    run<24, 2>(200, 20, 4);
    run<24, 4>(200, 20, 2);
    run<24, 1>(256, 20, 8);
    run<24, 2>(256, 20, 4);
    run<24, 4>(256, 20, 2);
    run<24, 3>(384, 20, 4);
    run<24, 2>(512, 20, 8);
    run<8, 1>(128, 6, 4);
    run<8, 4>(128, 6, 1);
    run<32, 1>(128, 30, 4);
    run<32, 2>(128, 30, 2);
    run<24, 1>(32, 20, 1);
    run<32, 1>(128, 20, 4);
    run<32, 2>(128, 20, 2);
    run<32, 2>(256, 20, 4);
    return 0;
}
