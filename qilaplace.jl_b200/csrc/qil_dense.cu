// qil_dense.cu -- generic strided contraction, tiled GEMM and elementwise helpers.
// These serve the small, latency-bound tensors of compress!/canonicalize!/the MPO builders and the
// lower levels of the divide-and-conquer encoder; the streaming top-level GEMMs live in qil_sketch.cu.
#include "qil_dense.cuh"

namespace qil {

// ------------------------------------------------------------------------------------------------
// generic contraction: one thread per output element, sequential loop over the contracted dims
// ------------------------------------------------------------------------------------------------
template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256) contract_kernel(const ContractDesc d, const TA* __restrict__ A,
                                                       const TB* __restrict__ B, TC* __restrict__ C,
                                                       long long total) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        long long rem = idx, oa = 0, ob = 0, oc = 0;
#pragma unroll
        for (int i = 5; i >= 0; --i) {
            if (i < d.nout) {
                const long long e = d.od[i];
                const long long q = rem / e, r = rem - q * e;
                rem = q;
                oa += r * d.sa_o[i];
                ob += r * d.sb_o[i];
                oc += r * d.sc_o[i];
            }
        }
        TC acc = Scalar<TC>::zero();
        const long long c0 = d.ncon > 0 ? d.cd[0] : 1, c1 = d.ncon > 1 ? d.cd[1] : 1, c2 = d.ncon > 2 ? d.cd[2] : 1;
        for (long long k0 = 0; k0 < c0; ++k0)
            for (long long k1 = 0; k1 < c1; ++k1) {
                const long long pa = oa + k0 * d.sa_c[0] + k1 * d.sa_c[1];
                const long long pb = ob + k0 * d.sb_c[0] + k1 * d.sb_c[1];
                for (long long k2 = 0; k2 < c2; ++k2) {
                    TC a = promote<TC, TA>(A[pa + k2 * d.sa_c[2]]);
                    if (d.conj_a) a = Scalar<TC>::conj(a);
                    const TC b = promote<TC, TB>(B[pb + k2 * d.sb_c[2]]);
                    acc = Scalar<TC>::fma(a, b, acc);
                }
            }
        C[oc] = acc;
    }
}

template <typename TA, typename TB, typename TC>
void contract(qil_ctx* ctx, const ContractDesc& din, const TA* A, const TB* B, TC* C) {
    ContractDesc d = din;
    long long total = 1;
    for (int i = 0; i < d.nout; ++i) total *= d.od[i];
    for (int i = d.ncon; i < 3; ++i) { d.cd[i] = 1; d.sa_c[i] = 0; d.sb_c[i] = 0; }
    if (total == 0) return;
    int grid = (int)std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 16);
    contract_kernel<TA, TB, TC><<<grid, 256, 0, ctx->stream>>>(d, A, B, C, total);
    QIL_LAUNCH_CHECK(ctx);
}

template void contract<double, double, double>(qil_ctx*, const ContractDesc&, const double*, const double*, double*);
template void contract<cplx, cplx, cplx>(qil_ctx*, const ContractDesc&, const cplx*, const cplx*, cplx*);
template void contract<double, cplx, cplx>(qil_ctx*, const ContractDesc&, const double*, const cplx*, cplx*);
template void contract<cplx, double, cplx>(qil_ctx*, const ContractDesc&, const cplx*, const double*, cplx*);

// ------------------------------------------------------------------------------------------------
// tiled GEMM (64x64 tile, 16x16 threads, 4x4 register tile); any op, any size
// ------------------------------------------------------------------------------------------------
constexpr int GT = 64, GK = 16;

template <typename T>
__device__ __forceinline__ T load_op(const T* __restrict__ A, int64_t ld, Op op, int64_t i, int64_t k, int64_t M,
                                     int64_t K) {
    // element (i,k) of op(A), where op(A) is M x K
    if (i >= M || k >= K) return Scalar<T>::zero();
    if (op == OP_N) return A[i * ld + k];
    T v = A[k * ld + i];
    return op == OP_C ? Scalar<T>::conj(v) : v;
}

template <typename T>
__global__ void __launch_bounds__(256) gemm_kernel(Op opa, Op opb, int64_t M, int64_t N, int64_t K, double alpha,
                                                   const T* __restrict__ A, int64_t lda, const T* __restrict__ B,
                                                   int64_t ldb, double beta, T* __restrict__ C, int64_t ldc) {
    __shared__ T As[GK][GT + 1];
    __shared__ T Bs[GK][GT + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    // tiles are numbered linearly on blockIdx.x (grid.y is limited to 65535: the 2^23-row steps of the sequential TT-SVD
    // at n = 24 have more row tiles than that)
    const int64_t tiles_n = (N + GT - 1) / GT;
    const int64_t i0 = ((int64_t)blockIdx.x / tiles_n) * GT, j0 = ((int64_t)blockIdx.x % tiles_n) * GT;
    T acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = Scalar<T>::zero();
    for (int64_t k0 = 0; k0 < K; k0 += GK) {
        for (int e = threadIdx.x; e < GK * GT; e += 256) {
            int kk, ii;
            if (opa == OP_N) { ii = e / GK; kk = e % GK; } else { kk = e / GT; ii = e % GT; }
            As[kk][ii] = load_op(A, lda, opa, i0 + ii, k0 + kk, M, K);
            int jj;
            if (opb == OP_N) { kk = e / GT; jj = e % GT; } else { jj = e / GK; kk = e % GK; }
            // op(B) is K x N: element (k, j)
            T v = Scalar<T>::zero();
            if (k0 + kk < K && j0 + jj < N) {
                if (opb == OP_N) v = B[(k0 + kk) * ldb + (j0 + jj)];
                else { v = B[(j0 + jj) * ldb + (k0 + kk)]; if (opb == OP_C) v = Scalar<T>::conj(v); }
            }
            Bs[kk][jj] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            T a[4], b[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) { a[x] = As[kk][ty * 4 + x]; b[x] = Bs[kk][tx * 4 + x]; }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = Scalar<T>::fma(a[x], b[y], acc[x][y]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int64_t i = i0 + ty * 4 + x, j = j0 + tx * 4 + y;
            if (i < M && j < N) {
                T v = Scalar<T>::scale(acc[x][y], alpha);
                if (beta != 0.0) v = Scalar<T>::add(v, Scalar<T>::scale(C[i * ldc + j], beta));
                C[i * ldc + j] = v;
            }
        }
}

template <typename T>
void gemm(qil_ctx* ctx, Op opa, Op opb, int64_t M, int64_t N, int64_t K, double alpha, const T* A, int64_t lda,
          const T* B, int64_t ldb, double beta, T* C, int64_t ldc) {
    if (M == 0 || N == 0) return;
    const int64_t tiles = ((N + GT - 1) / GT) * ((M + GT - 1) / GT);
    QIL_REQUIRE(tiles < ((int64_t)1 << 31), QIL_ERR_UNSUPPORTED, "gemm: %lld x %lld has too many tiles", (long long)M, (long long)N);
    gemm_kernel<T><<<(unsigned)tiles, 256, 0, ctx->stream>>>(opa, opb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    QIL_LAUNCH_CHECK(ctx);
}
template void gemm<double>(qil_ctx*, Op, Op, int64_t, int64_t, int64_t, double, const double*, int64_t,
                           const double*, int64_t, double, double*, int64_t);
template void gemm<cplx>(qil_ctx*, Op, Op, int64_t, int64_t, int64_t, double, const cplx*, int64_t, const cplx*,
                         int64_t, double, cplx*, int64_t);

// ------------------------------------------------------------------------------------------------
// elementwise helpers
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void scale_copy_kernel(long long n, double alpha, const T* __restrict__ x, T* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = Scalar<T>::scale(x[i], alpha);
}
template <typename T>
void scale_copy(qil_ctx* ctx, int64_t n, double alpha, const T* x, T* y) {
    if (n == 0) return;
    int grid = (int)std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16);
    scale_copy_kernel<T><<<grid, 256, 0, ctx->stream>>>(n, alpha, x, y);
    QIL_LAUNCH_CHECK(ctx);
}
template void scale_copy<double>(qil_ctx*, int64_t, double, const double*, double*);
template void scale_copy<cplx>(qil_ctx*, int64_t, double, const cplx*, cplx*);

template <typename T>
__global__ void scale_rc_kernel(long long m, long long n, const T* __restrict__ A, long long lda,
                                const double* __restrict__ s, int by_row, int inv, T* __restrict__ B, long long ldb) {
    const long long total = m * n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx / n, j = idx - i * n;
        double f = s[by_row ? i : j];
        if (inv) f = (f != 0.0) ? 1.0 / f : 0.0;
        B[i * ldb + j] = Scalar<T>::scale(A[i * lda + j], f);
    }
}
template <typename T>
void scale_rows_cols(qil_ctx* ctx, int64_t m, int64_t n, const T* A, int64_t lda, const double* s, bool by_row,
                     bool inv, T* B, int64_t ldb) {
    if (m * n == 0) return;
    int grid = (int)std::min<long long>((m * n + 255) / 256, (long long)ctx->sm_count * 16);
    scale_rc_kernel<T><<<grid, 256, 0, ctx->stream>>>(m, n, A, lda, s, by_row ? 1 : 0, inv ? 1 : 0, B, ldb);
    QIL_LAUNCH_CHECK(ctx);
}
template void scale_rows_cols<double>(qil_ctx*, int64_t, int64_t, const double*, int64_t, const double*, bool, bool,
                                      double*, int64_t);
template void scale_rows_cols<cplx>(qil_ctx*, int64_t, int64_t, const cplx*, int64_t, const double*, bool, bool,
                                    cplx*, int64_t);

template <typename T>
__global__ void transpose_conj_kernel(long long m, long long n, const T* __restrict__ A, long long lda,
                                      T* __restrict__ B, long long ldb, int conj) {
    __shared__ T tile[32][33];
    // tiles are numbered linearly on blockIdx.x (grid.y is limited to 65535: a 2 x 2^21 step of the sequential TT-SVD
    // at n = 22 has more row tiles than that)
    const long long tiles_x = (n + 31) / 32;
    const long long i0 = ((long long)blockIdx.x / tiles_x) * 32, j0 = ((long long)blockIdx.x % tiles_x) * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const long long i = i0 + r, j = j0 + threadIdx.x;
        if (i < m && j < n) tile[r][threadIdx.x] = conj ? Scalar<T>::conj(A[i * lda + j]) : A[i * lda + j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const long long j = j0 + r, i = i0 + threadIdx.x;
        if (i < m && j < n) B[j * ldb + i] = tile[threadIdx.x][r];
    }
}
template <typename T>
void transpose_conj(qil_ctx* ctx, int64_t m, int64_t n, const T* A, int64_t lda, T* B, int64_t ldb, bool conj) {
    if (m * n == 0) return;
    const long long tiles = ((n + 31) / 32) * ((m + 31) / 32);
    QIL_REQUIRE(tiles < ((long long)1 << 31), QIL_ERR_UNSUPPORTED, "transpose: matrix too large");
    dim3 grid((unsigned)tiles), block(32, 8);
    transpose_conj_kernel<T><<<grid, block, 0, ctx->stream>>>(m, n, A, lda, B, ldb, conj ? 1 : 0);
    QIL_LAUNCH_CHECK(ctx);
}
template void transpose_conj<double>(qil_ctx*, int64_t, int64_t, const double*, int64_t, double*, int64_t, bool);
template void transpose_conj<cplx>(qil_ctx*, int64_t, int64_t, const cplx*, int64_t, cplx*, int64_t, bool);

}  // namespace qil
