// does access through a generic pointer to dynamic shared memory cost more than ld.shared with a 32-bit address?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ double lds(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v)); }

template <int MODE>
__device__ __noinline__ double dot_update(double* As, uint32_t base, int mpad, int j, int c, int mloc, int lane) {
    double w = 0.0;
    if (MODE == 0) {
        const double* col = As + j * mpad; double* cc = As + c * mpad;
        for (int i = j + lane; i < mloc; i += 32) w = fma(col[i], cc[i], w);
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        for (int i = j + lane; i < mloc; i += 32) cc[i] = fma(-w * 1e-3, col[i], cc[i]);
    } else {
        const uint32_t col = base + 8u * (j * mpad), cc = base + 8u * (c * mpad);
        for (int i = j + lane; i < mloc; i += 32) w = fma(lds(col + 8u * i), lds(cc + 8u * i), w);
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        for (int i = j + lane; i < mloc; i += 32) sts(cc + 8u * i, fma(-w * 1e-3, lds(col + 8u * i), lds(cc + 8u * i)));
    }
    return w;
}
template <int MODE>
__global__ void k(double* out, long long* cyc, int mloc, int n) {
    extern __shared__ __align__(16) unsigned char raw[];
    double* As = reinterpret_cast<double*>(raw);
    const int mpad = mloc | 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < n * mpad; i += blockDim.x) As[i] = 1.0 + 1e-6 * i;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(As);
    long long t0 = clock64();
    double acc = 0;
    for (int j = 0; j < n - 8; ++j) { acc += dot_update<MODE>(As, base, mpad, j, j + 1 + warp, mloc, lane); __syncthreads(); }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = acc;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 256); cudaMalloc(&cyc, 8);
    for (int mloc : {64, 256}) {
        int n = 20; size_t smem = (size_t)n * (mloc | 1) * 8;
        long long h;
        cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<0><<<1, 256, smem>>>(out, cyc, mloc, n); k<0><<<1, 256, smem>>>(out, cyc, mloc, n);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("mloc=%d generic-pointer smem : %lld cycles per column step\n", mloc, h / (n - 8));
        k<1><<<1, 256, smem>>>(out, cyc, mloc, n); k<1><<<1, 256, smem>>>(out, cyc, mloc, n);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("mloc=%d ld.shared 32-bit addr   : %lld cycles per column step\n", mloc, h / (n - 8));
    }
    return 0;
}
