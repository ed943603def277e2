// qil_qr.cu -- K3 / K5: Householder thin QR in shared memory and TSQR for tall-skinny matrices.
//
// Replaces ITensors `qr` at rsvd.jl:83,90,94 (positive=true), dt_transformer.jl:73,131,190,237 and the
// QR branch of `factorize` (qft_transformer.jl:52).  Householder reflectors keep Q orthonormal to
// rounding even when the matrix is numerically rank deficient (the usual case for Y = A*Omega of a
// structured signal), which Gram/Cholesky shortcuts do not.
#include "qil_fast.cuh"
#include "qil_hh.cuh"

namespace qil {


constexpr int kQrThreads = 256;        // shared-memory variant
constexpr int kQrWarps = kQrThreads / 32;
constexpr int kQrThreadsGlobal = 1024; // single-CTA variant on an L2-resident scratch copy (large bond matrices)

template <typename T>
struct QrParams {
    const T* A;          // input, row-major
    long long lda;
    int nsum;            // number of partial matrices summed on load
    long long sum_stride;
    long long m;         // total rows
    int n;               // columns
    int nblk;            // row blocks (balanced split); block b owns rows [b*m/nblk, (b+1)*m/nblk)
    T* Q;                // m x kq, row-major (ldq); may be null
    long long ldq;
    T* R;                // block b writes rows [b*rrows, b*rrows + min(mloc,n)) of an (nblk*rrows) x n matrix
    int rrows;
    int positive;
    int mpad;            // smem column pitch (elements)
    T* gscratch;         // if non-null: the (n + nwarps) x mpad work area lives here instead of shared memory
};

template <typename T>
__device__ __forceinline__ T warp_sum(T v);
template <>
__device__ __forceinline__ double warp_sum<double>(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <>
__device__ __forceinline__ cplx warp_sum<cplx>(cplx v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}

// GLOBAL = false: the work area is shared memory (the compiler then emits LDS/STS with 32-bit addressing);
// GLOBAL = true : single CTA working on an L2-resident scratch (wide bond matrices that do not fit).
template <typename T, bool GLOBAL>
__global__ void __launch_bounds__(GLOBAL ? kQrThreadsGlobal : kQrThreads) hhqr_kernel(const QrParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n, mpad = p.mpad;
    const int kQrThreads = blockDim.x, kQrWarps = blockDim.x >> 5;
    T* As;                                                 // [n][mpad] column-major
    T* sbeta;                                              // [n]
    if (GLOBAL) {
        As = p.gscratch;
        sbeta = reinterpret_cast<T*>(smem_raw);
    } else {
        As = reinterpret_cast<T*>(smem_raw);
        sbeta = As + (n + kHhIlp * kQrWarps) * mpad;
    }
    T* qb = As + n * mpad;                                 // [kHhIlp * kQrWarps][mpad]
    T* su0 = sbeta + n;                                    // [n]
    double* ss = reinterpret_cast<double*>(su0 + n);       // [n]

    const int b = blockIdx.x;
    const long long r0 = ((long long)b * p.m) / p.nblk;
    const long long r1 = ((long long)(b + 1) * p.m) / p.nblk;
    const int mloc = (int)(r1 - r0);
    const int k = min(mloc, n);
    const int tid = threadIdx.x, lane = tid & 31;

    // ---- load (sum of partials), transposing into column-major smem
    for (int idx = tid; idx < mloc * n; idx += kQrThreads) {
        const int i = idx / n, j = idx - i * n;
        const T* src = p.A + (r0 + i) * p.lda + j;
        T v = src[0];
        for (int s = 1; s < p.nsum; ++s) v = Scalar<T>::add(v, src[(long long)s * p.sum_stride]);
        As[j * mpad + i] = v;
    }
    __syncthreads();

    // ---- factorisation: H_j = I - s_j u_j u_j^H,  H_j x = beta_j e_1  (qil_hh.cuh)
    hh_factor<T>(As, mpad, mloc, n, k, sbeta, ss);

    // ---- R (k x n), optional positive diagonal
    if (p.R) {
        for (int idx = tid; idx < k * n; idx += kQrThreads) {
            const int j = idx / n, c = idx - j * n;
            T v = Scalar<T>::zero();
            if (c == j) v = sbeta[j];
            else if (c > j) v = As[c * mpad + j];
            if (p.positive) {
                const T bj = sbeta[j];
                const double ab = sqrt(Scalar<T>::abs2(bj));
                if (ab > 0.0) v = Scalar<T>::mul(Scalar<T>::conj(Scalar<T>::scale(bj, 1.0 / ab)), v);
            }
            p.R[((long long)b * p.rrows + j) * n + c] = v;
        }
    }

    // ---- explicit Q (mloc x k)
    if (p.Q) {
        hh_form_q<T>(As, mpad, mloc, k, ss, qb, [&](int c, const T* q) {
            T ph = Scalar<T>::one();
            if (p.positive) {
                const T bj = sbeta[c];
                const double ab = sqrt(Scalar<T>::abs2(bj));
                if (ab > 0.0) ph = Scalar<T>::scale(bj, 1.0 / ab);
            }
            for (int i = lane; i < mloc; i += 32) p.Q[(r0 + i) * p.ldq + c] = Scalar<T>::mul(q[i], ph);
        });
    }
}

// out[b-th row block] = Q0[b] (mloc x n) * Q1[b*n : (b+1)*n, :] (n x n)
template <typename T>
__global__ void __launch_bounds__(256) tsqr_combine_kernel(const T* __restrict__ Q0, const T* __restrict__ Q1,
                                                           T* __restrict__ out, long long m, int n, int nblk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s1 = reinterpret_cast<T*>(smem_raw);  // [n][n]
    const int b = blockIdx.x;
    const long long r0 = ((long long)b * m) / nblk, r1 = ((long long)(b + 1) * m) / nblk;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) s1[idx] = Q1[(long long)b * n * n + idx];
    __syncthreads();
    const long long tot = (r1 - r0) * n;
    for (long long idx = (long long)blockIdx.y * blockDim.x + threadIdx.x; idx < tot;
         idx += (long long)gridDim.y * blockDim.x) {
        const long long i = idx / n;
        const int c = (int)(idx - i * n);
        const T* row = Q0 + (r0 + i) * n;
        T acc = Scalar<T>::zero();
        for (int j = 0; j < n; ++j) acc = Scalar<T>::fma(row[j], s1[j * n + c], acc);
        out[(r0 + i) * n + c] = acc;
    }
}

template <typename T>
static size_t qr_smem(int n, int mloc) {
    const int mpad = mloc | 1;
    return ((size_t)n * mpad + (size_t)kHhIlp * kQrWarps * mpad + 2 * (size_t)n) * sizeof(T) + (size_t)n * sizeof(double) + 32;
}

template <typename T>
static int qr_capacity(qil_ctx* ctx, int n) {
    const size_t budget = std::min<size_t>(ctx->smem_optin, 220 * 1024);
    // rows that fit: (n + warps) * mpad * sizeof(T) <= budget - small
    long long fixed = (2ll * n) * sizeof(T) + (long long)n * 8 + 64;
    long long rows = ((long long)budget - fixed) / ((long long)(n + kHhIlp * kQrWarps) * sizeof(T));
    rows -= 2;
    return (int)std::max<long long>(rows, 0);
}

template <typename T>
static void launch_hhqr(qil_ctx* ctx, const QrParams<T>& p, int max_mloc, bool use_global = false) {
    QrParams<T> q = p;
    q.mpad = max_mloc | 1;
    if (!use_global) {
        auto kern = hhqr_kernel<T, false>;
        const size_t smem = qr_smem<T>(p.n, max_mloc);
        q.gscratch = nullptr;
        ensure_dynamic_smem(kern, smem);
        kern<<<p.nblk, kQrThreads, smem, ctx->stream>>>(q);
        QIL_LAUNCH_CHECK(ctx);
        return;
    }
    // one CTA, 32 warps, work area in global memory (stays in L2): slow but shape-agnostic
    QIL_REQUIRE(p.nblk == 1, QIL_ERR_RUNTIME, "qr: global-scratch variant handles a single block");
    const int nw = kQrThreadsGlobal / 32;
    Mat<T> scratch(ctx, (int64_t)p.n + kHhIlp * nw, q.mpad);
    q.gscratch = scratch.p;
    const size_t smem = 2 * (size_t)p.n * sizeof(T) + (size_t)p.n * sizeof(double) + 32;
    auto kern = hhqr_kernel<T, true>;
    ensure_dynamic_smem(kern, smem);
    kern<<<1, kQrThreadsGlobal, smem, ctx->stream>>>(q);
    QIL_LAUNCH_CHECK(ctx);
}


// ---- blocked QR for tall panels wider than the TSQR's 32 columns (k = 100 sketches: l = 105, complex l = 55) ------------
// Block classical Gram-Schmidt with reorthogonalisation ("twice is enough") over panels of <= 32 (complex: 16) columns,
// each panel factored by the Householder TSQR:  P <- P - Q (Q^H P) twice, then P = Qp Rp.  The inner products Q^H P are
// m-long reductions with a handful of outputs: row chunks per CTA, fixed-order second stage (deterministic).
template <typename T>
__global__ void __launch_bounds__(256) tall_inner_kernel(const T* __restrict__ Q, long long ldq, int c0, const T* __restrict__ P,
                                                         long long ldp, int w, long long m, int rows_per, T* __restrict__ part) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sq = reinterpret_cast<T*>(smem_raw);          // [rows][c0]
    T* sp = sq + (size_t)rows_per * c0;              // [rows][w]
    const long long r0 = (long long)blockIdx.x * rows_per;
    const int rows = (int)min((long long)rows_per, m - r0);
    for (int idx = threadIdx.x; idx < rows * c0; idx += blockDim.x) {
        const int r = idx / c0, i = idx - r * c0;
        sq[idx] = Q[(r0 + r) * ldq + i];
    }
    for (int idx = threadIdx.x; idx < rows * w; idx += blockDim.x) {
        const int r = idx / w, j = idx - r * w;
        sp[idx] = P[(r0 + r) * ldp + j];
    }
    __syncthreads();
    T* out = part + (size_t)blockIdx.x * c0 * w;
    for (int idx = threadIdx.x; idx < c0 * w; idx += blockDim.x) {
        const int j = idx / c0, i = idx - j * c0;     // consecutive threads: consecutive i (contiguous in sq rows)
        T acc = Scalar<T>::zero();
        for (int r = 0; r < rows; ++r) acc = Scalar<T>::fma(Scalar<T>::conj(sq[r * c0 + i]), sp[r * w + j], acc);
        out[(size_t)i * w + j] = acc;
    }
}
// S (c0 x w, ld lds) = sum of the chunk partials (fixed order); R block (ld ldr) accumulates it
template <typename T>
__global__ void tall_inner_reduce_kernel(const T* __restrict__ part, int nchunks, int c0, int w, T* __restrict__ S,
                                         T* __restrict__ Rblk, long long ldr, int accumulate) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < c0 * w; idx += gridDim.x * blockDim.x) {
        T acc = Scalar<T>::zero();
        for (int c = 0; c < nchunks; ++c) acc = Scalar<T>::add(acc, part[(size_t)c * c0 * w + idx]);
        S[idx] = acc;
        const int i = idx / w, j = idx - i * w;
        T* r = Rblk + (long long)i * ldr + j;
        *r = accumulate ? Scalar<T>::add(*r, acc) : acc;
    }
}

template <typename T>
static void qr_blocked_tall(qil_ctx* ctx, int64_t m, int n, const T* A, int64_t lda, bool positive, Mat<T>& Q, Mat<T>& R,
                            int nsum, int64_t sum_stride) {
    const int bw = Scalar<T>::is_complex ? 16 : 32;
    Q = Mat<T>(ctx, m, n);
    R = Mat<T>(ctx, n, n);
    QIL_CUDA(cudaMemsetAsync(R.p, 0, (size_t)n * n * sizeof(T), ctx->stream));
    const size_t budget = std::min<size_t>(ctx->smem_optin, 200 * 1024);
    int rows_per = 128;
    while (rows_per > 16 && (size_t)rows_per * (n + bw) * sizeof(T) > budget) rows_per >>= 1;
    const int nchunks = (int)((m + rows_per - 1) / rows_per);
    Mat<T> part(ctx, (int64_t)nchunks * n, bw), S(ctx, n, bw), P(ctx, m, bw), Rp(ctx, bw, bw);
    for (int c0 = 0; c0 < n; c0 += bw) {
        const int w = std::min(bw, n - c0);
        // P = A[:, c0:c0+w]
        if (nsum == 1) {
            QIL_CUDA(cudaMemcpy2DAsync(P.p, (size_t)w * sizeof(T), A + c0, (size_t)lda * sizeof(T), (size_t)w * sizeof(T),
                                       (size_t)m, cudaMemcpyDeviceToDevice, ctx->stream));
        } else {
            QIL_THROW(QIL_ERR_UNSUPPORTED, "blocked qr: partial-sum input is not supported");
        }
        for (int rep = 0; rep < 2 && c0 > 0; ++rep) {
            const size_t smem = (size_t)rows_per * (c0 + w) * sizeof(T);
            auto kern = tall_inner_kernel<T>;
            ensure_dynamic_smem(kern, smem);
            kern<<<nchunks, 256, smem, ctx->stream>>>(Q.p, n, c0, P.p, w, w, m, rows_per, part.p);
            QIL_LAUNCH_CHECK(ctx);
            tall_inner_reduce_kernel<T><<<(c0 * w + 255) / 256, 256, 0, ctx->stream>>>(part.p, nchunks, c0, w, S.p, R.p + c0, n, rep);
            QIL_LAUNCH_CHECK(ctx);
            // P -= Q[:, :c0] S
            gemm<T>(ctx, OP_N, OP_N, m, w, c0, -1.0, Q.p, n, S.p, w, 1.0, P.p, w);
        }
        qr_fast<T>(ctx, m, w, P.p, w, 1, 0, positive, Q.p + c0, n, w, Rp.p);
        QIL_CUDA(cudaMemcpy2DAsync(R.p + (size_t)c0 * n + c0, (size_t)n * sizeof(T), Rp.p, (size_t)w * sizeof(T),
                                   (size_t)w * sizeof(T), (size_t)w, cudaMemcpyDeviceToDevice, ctx->stream));
    }
}

template <typename T>
void qr_thin(qil_ctx* ctx, int64_t m, int64_t n64, const T* A, int64_t lda, bool positive, Mat<T>& Q, Mat<T>& R,
             int nsum, int64_t sum_stride, bool want_q) {
    struct Region {   // profiler class 3 (outermost region only; TSQR recursion and SVD callers nest)
        qil_ctx* c;
        explicit Region(qil_ctx* cc) : c(cc) { c->prof_begin(PROF_QR); }
        ~Region() { c->prof_end(); }
    } region(ctx);
    QIL_REQUIRE(m >= 1 && n64 >= 1, QIL_ERR_ARGUMENT, "qr: empty matrix");
    QIL_REQUIRE(n64 < (1 << 20), QIL_ERR_UNSUPPORTED, "qr: too many columns");
    const int n = (int)n64;
    if (qr_fast_supported<T>(ctx, m, n)) {
        // skinny panels (n <= 32): warp-synchronous TSQR (qil_tsqr.cu), one launch (or three for tall ones)
        Mat<T> Qf(ctx, m, n);
        R = Mat<T>(ctx, n, n);
        qr_fast<T>(ctx, m, n, A, lda, nsum, sum_stride, positive, Qf.p, n, n, R.p);
        if (want_q) Q = std::move(Qf);
        return;
    }
    {
        // tall panels wider than the TSQR limit: block Gram-Schmidt over TSQR panels
        const int bw = Scalar<T>::is_complex ? 16 : 32;
        if (n > bw && n <= 512 && m >= 8 * (int64_t)n && nsum == 1 && qr_fast_supported<T>(ctx, m, bw) &&
            (size_t)16 * (n + bw) * sizeof(T) <= std::min<size_t>(ctx->smem_optin, 200 * 1024)) {
            Mat<T> Qb;
            qr_blocked_tall<T>(ctx, m, n, A, lda, positive, Qb, R, nsum, sum_stride);
            if (want_q) Q = std::move(Qb);
            return;
        }
    }
    const int cap = qr_capacity<T>(ctx, n);
    const int64_t k = std::min<int64_t>(m, n);
    if (m <= cap) {
        if (want_q) Q = Mat<T>(ctx, m, k);
        R = Mat<T>(ctx, k, n);
        QrParams<T> p{A, lda, nsum, sum_stride, m, n, 1, want_q ? Q.p : nullptr, k, R.p, (int)k, positive ? 1 : 0, 0, nullptr};
        launch_hhqr(ctx, p, (int)m);
        return;
    }
    if (cap < 2 * n) {
        // wide bond matrices of the zT builder / large compress!: single CTA on a global work area
        QIL_REQUIRE(m * (int64_t)n <= ((int64_t)1 << 26), QIL_ERR_UNSUPPORTED,
                    "qr: %lld x %d exceeds the global-scratch Householder path", (long long)m, n);
        if (want_q) Q = Mat<T>(ctx, m, k);
        R = Mat<T>(ctx, k, n);
        QrParams<T> p{A, lda, nsum, sum_stride, m, n, 1, want_q ? Q.p : nullptr, k, R.p, (int)k, positive ? 1 : 0, 0, nullptr};
        launch_hhqr(ctx, p, (int)m, true);
        return;
    }
    // TSQR: balanced row blocks of at most mb rows, each with at least n rows
    int mb = std::min(cap, std::max(2 * n, 256));
    int64_t nblk = (m + mb - 1) / mb;
    while (nblk > 1 && m / nblk < n) --nblk;
    const int max_mloc = (int)((m + nblk - 1) / nblk);
    QIL_REQUIRE(max_mloc <= cap, QIL_ERR_UNSUPPORTED, "qr: TSQR block of %d rows exceeds capacity %d", max_mloc, cap);
    Mat<T> Q0;
    if (want_q) Q0 = Mat<T>(ctx, m, n);
    Mat<T> Rst(ctx, nblk * n, n);
    QrParams<T> p{A, lda, nsum, sum_stride, m, n, (int)nblk, want_q ? Q0.p : nullptr, n, Rst.p, n, 0, 0, nullptr};
    launch_hhqr(ctx, p, max_mloc);
    Mat<T> Q1;
    qr_thin<T>(ctx, nblk * n, n, Rst.p, n, positive, Q1, R, 1, 0, want_q);
    if (want_q) {
        Q = Mat<T>(ctx, m, n);
        const size_t smem = (size_t)n * n * sizeof(T);
        auto kern = tsqr_combine_kernel<T>;
        ensure_dynamic_smem(kern, smem);
        int gy = (int)std::max<int64_t>(1, std::min<int64_t>(8, ((int64_t)max_mloc * n + 255) / 256));
        dim3 grid((unsigned)nblk, gy);
        kern<<<grid, 256, smem, ctx->stream>>>(Q0.p, Q1.p, Q.p, (long long)m, n, (int)nblk);
        QIL_LAUNCH_CHECK(ctx);
    }
}

template void qr_thin<double>(qil_ctx*, int64_t, int64_t, const double*, int64_t, bool, Mat<double>&, Mat<double>&,
                              int, int64_t, bool);
template void qr_thin<cplx>(qil_ctx*, int64_t, int64_t, const cplx*, int64_t, bool, Mat<cplx>&, Mat<cplx>&, int,
                            int64_t, bool);

}  // namespace qil
