// ubench_fp64.cu -- B200 micro-benchmarks that size the encode kernels (K1/K2):
//   * DFMA vs DMMA (mma.sync f64 m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16) peak issue rate
//   * DMMA fragment-layout check against a scalar product
//   * streaming read bandwidth of a 2 GiB buffer with plain 128-bit loads
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench_fp64 tools/ubench_fp64.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void dfma_kernel(double* out, int iters) {
    double a[8];
    double x = 1.0000001 + threadIdx.x * 1e-9, y = 0.999999;
    for (int i = 0; i < 8; ++i) a[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double* c, const double* a, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int KIND>
__global__ void dmma_kernel(double* out, int iters) {
    double c[4][4];
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    for (int i = 0; i < 4; ++i) b[i] = 1.0 - 1e-9 * (threadIdx.x + i);
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (KIND == 0) { mma884(c[j][0], c[j][1], a[0], b[0]); }
            if (KIND == 1) { mma1684(c[j], a, b[0]); }
            if (KIND == 2) { mma1688(c[j], a, b); }
            if (KIND == 3) { mma16816(c[j], a, b); }
        }
    }
    double s = 0; for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: every warp issues DMMA and independent DFMA chains -- do the two FP64 paths add up?
template <int NF>
__global__ void mixed_kernel(double* out, int iters) {
    double c[4][4];
    double a[8], b[4], f[NF > 0 ? NF : 1];
    double x = 1.0000001 + threadIdx.x * 1e-9, y = 0.999999;
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    for (int i = 0; i < 4; ++i) b[i] = 1.0 - 1e-9 * (threadIdx.x + i);
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0;
    for (int i = 0; i < NF; ++i) f[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            mma16816(c[j], a, b);
#pragma unroll
            for (int i = 0; i < NF; ++i) f[i] = fma(f[i], x, y);
        }
    }
    double s = 0; for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
    for (int i = 0; i < NF; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// layout check: C[16x8] = A[16x16] * B[16x8] with the assumed fragment layout
__global__ void layout_kernel(const double* A, const double* B, double* C) {
    int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    double a[8], b[4], c[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        a[2 * i + 0] = A[g * 16 + t + 4 * i];
        a[2 * i + 1] = A[(g + 8) * 16 + t + 4 * i];
        b[i] = B[(t + 4 * i) * 8 + g];
    }
    mma16816(c, a, b);
    C[g * 8 + 2 * t] = c[0]; C[g * 8 + 2 * t + 1] = c[1];
    C[(g + 8) * 8 + 2 * t] = c[2]; C[(g + 8) * 8 + 2 * t + 1] = c[3];
}

__global__ void read_kernel(const double2* __restrict__ p, size_t n, double* out) {
    double s = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        double2 v0 = p[i], v1 = p[i + stride], v2 = p[i + 2 * stride], v3 = p[i + 3 * stride];
        s += v0.x + v0.y + v1.x + v1.y + v2.x + v2.y + v3.x + v3.y;
    }
    for (; i < n; i += stride) { double2 v = p[i]; s += v.x + v.y; }
    if (s == 1.2345e300) out[0] = s;
}

template <typename F> float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s sms=%d\n", prop.name, prop.multiProcessorCount);
    int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int iters = 20000;
    for (int wpb : {4, 8, 16, 32}) {
        int blocks = sms * 2;
        float ms = time_ms([&] { dfma_kernel<<<blocks, wpb * 32>>>(out, iters); }, 3);
        double fl = 2.0 * 8 * iters * (double)blocks * wpb * 32;
        printf("DFMA   warps/blk=%2d blocks=%d : %.2f TFLOP/s\n", wpb, blocks, fl / ms / 1e9);
    }
    const char* names[4] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
    double fl_per[4] = {2.0 * 8 * 8 * 4, 2.0 * 16 * 8 * 4, 2.0 * 16 * 8 * 8, 2.0 * 16 * 8 * 16};
    for (int kind = 0; kind < 4; ++kind)
        for (int wpb : {4, 8, 16}) {
            int blocks = sms * 2;
            int it = iters / (kind + 1);
            float ms = time_ms([&] {
                if (kind == 0) dmma_kernel<0><<<blocks, wpb * 32>>>(out, it);
                if (kind == 1) dmma_kernel<1><<<blocks, wpb * 32>>>(out, it);
                if (kind == 2) dmma_kernel<2><<<blocks, wpb * 32>>>(out, it);
                if (kind == 3) dmma_kernel<3><<<blocks, wpb * 32>>>(out, it);
            }, 3);
            double fl = fl_per[kind] * 4 * it * (double)blocks * wpb;
            printf("DMMA %-9s warps/blk=%2d : %.2f TFLOP/s\n", names[kind], wpb, fl / ms / 1e9);
        }
    {
        int blocks = sms * 2, wpb = 8, it = 4000;
        auto report = [&](const char* nm, int nf, float ms) {
            double fl_mma = fl_per[3] * 4 * it * (double)blocks * wpb;
            double fl_fma = 2.0 * nf * 4 * it * (double)blocks * wpb * 32;
            printf("MIXED %-22s : DMMA %.2f + DFMA %.2f = %.2f TFLOP/s\n", nm, fl_mma / ms / 1e9, fl_fma / ms / 1e9, (fl_mma + fl_fma) / ms / 1e9);
        };
        report("dmma only", 0, time_ms([&] { mixed_kernel<0><<<blocks, wpb * 32>>>(out, it); }, 3));
        report("dmma + 8 dfma/mma", 8, time_ms([&] { mixed_kernel<8><<<blocks, wpb * 32>>>(out, it); }, 3));
        report("dmma + 16 dfma/mma", 16, time_ms([&] { mixed_kernel<16><<<blocks, wpb * 32>>>(out, it); }, 3));
        report("dmma + 32 dfma/mma", 32, time_ms([&] { mixed_kernel<32><<<blocks, wpb * 32>>>(out, it); }, 3));
        report("dmma + 64 dfma/mma", 64, time_ms([&] { mixed_kernel<64><<<blocks, wpb * 32>>>(out, it); }, 3));
    }
    // layout check
    {
        std::vector<double> A(256), B(128), Cw(128, 0), Cg(128);
        for (int i = 0; i < 256; ++i) A[i] = sin(0.37 * i) ;
        for (int i = 0; i < 128; ++i) B[i] = cos(0.11 * i);
        for (int m = 0; m < 16; ++m) for (int n = 0; n < 8; ++n) for (int k = 0; k < 16; ++k) Cw[m * 8 + n] += A[m * 16 + k] * B[k * 8 + n];
        double *dA, *dB, *dC; CK(cudaMalloc(&dA, 2048)); CK(cudaMalloc(&dB, 1024)); CK(cudaMalloc(&dC, 1024));
        CK(cudaMemcpy(dA, A.data(), 2048, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), 1024, cudaMemcpyHostToDevice));
        layout_kernel<<<1, 32>>>(dA, dB, dC); CK(cudaMemcpy(Cg.data(), dC, 1024, cudaMemcpyDeviceToHost));
        double err = 0; for (int i = 0; i < 128; ++i) err = fmax(err, fabs(Cg[i] - Cw[i]));
        printf("m16n8k16 fragment layout check: max err %.3e (%s)\n", err, err < 1e-12 ? "OK" : "MISMATCH");
    }
    // HBM read bandwidth, 2 GiB
    {
        size_t bytes = (size_t)2 << 30; double2* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
        for (int bpsm : {4, 8, 16}) {
            float ms = time_ms([&] { read_kernel<<<sms * bpsm, 256>>>(buf, bytes / 16, out); }, 5);
            printf("read 2GiB blocks/sm=%2d : %.1f GB/s (%.3f ms)\n", bpsm, bytes / ms / 1e6, ms);
        }
        CK(cudaFree(buf));
    }
    return 0;
}
