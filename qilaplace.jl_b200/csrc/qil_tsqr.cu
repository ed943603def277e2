// qil_tsqr.cu -- fast thin QR of tall-skinny panels (n <= 32 columns) from the warp-synchronous blocks of
// qil_wqr.cuh.  Replaces the shared-memory CTA Householder (qil_qr.cu, ~1.7 us per column step) on the latency path
// of the randomized SVD (rsvd.jl:83,90,94 and the QR inside rsvd.jl:103's svd):
//
//   m small enough for one CTA's shared memory : ONE launch   (tsqr_cta_kernel: multi-level TSQR inside the CTA)
//   otherwise                                   : THREE        (leaf factor | tree on the stacked triangles | leaf apply)
//
// The leaf kernels work on 512-row blocks (one warp, 16 rows per lane); the explicit Q is produced by applying the
// leaf reflectors to the tree's n x n seed of that block (no separate form-Q + combine), 4-column chunks in
// registers, one warp per chunk.  Input may be given as `nsum` partial matrices summed on load (the split-K
// partials of the streaming GEMM), output may have a padded pitch with zero-filled columns (the operand layout of
// the next streaming pass), so the reduce / prepare launches between a pass and its QR disappear.  A batch of equal
// problems rides on blockIdx.y.
#include "qil_dense.cuh"
#include "qil_wqr.cuh"
#include "qil_fast.cuh"

namespace qil {

// leaf blocks: up to 512 rows (complex: 256) per CTA, 16 (8) row slots per lane
template <typename T> struct Leaf { static constexpr int RPL = 16; static constexpr int ROWS = 512; };
template <> struct Leaf<cplx> { static constexpr int RPL = 8; static constexpr int ROWS = 256; };
constexpr int kLeafThreads = 256;             // <= 256 threads: the Householder warp may use up to 255 registers
// single block of <= 256 rows: 8 factor warps and <= 8 apply chunks of 4 columns for real panels (255 registers, no
// spills); complex panels have up to 16 chunks of 2 columns and keep 16 warps
template <typename T> struct TreeThreads { static constexpr int N = 256; };
template <> struct TreeThreads<cplx> { static constexpr int N = 512; };
template <typename T> struct LeafChunk { static constexpr int CH = 4; };
template <> struct LeafChunk<cplx> { static constexpr int CH = 2; };

template <typename T>
struct TsqrParams {
    const T* A; long long lda; int nsum; long long sum_stride; long long a_bs;
    long long m; int n; int nblk;
    T* V; long long v_bs;              // leaf reflectors, m x n (ld n)
    double* tau; long long tau_bs;     // [nblk][n]
    T* beta;                           // [nblk][n]   (same batch stride as tau)
    T* Rst; long long rst_bs;          // stacked triangles [nblk * n][n]
    const T* M; long long m_bs;        // explicit Q of the tree, [nblk * n][n]
    T* Q; long long ldq; int qcols; long long q_bs;
    T* R; long long r_bs;              // n x n, ld n
    int positive;
    int pitch;                         // shared-memory pitch (odd)
};

template <typename T>
__device__ __forceinline__ T load_sum(const T* p, int nsum, long long stride) {
    T v = p[0];
    for (int s = 1; s < nsum; ++s) v = Scalar<T>::add(v, p[(long long)s * stride]);
    return v;
}

// ---- leaf factor: one 512-row block per CTA ----------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kLeafThreads) tsqr_leaf_factor_kernel(const TsqrParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* blk = reinterpret_cast<T*>(smem_raw);
    const int b = blockIdx.x;
    const long long bat = blockIdx.y;
    const long long r0 = ((long long)b * p.m) / p.nblk, r1 = ((long long)(b + 1) * p.m) / p.nblk;
    const int mloc = (int)(r1 - r0), n = p.n, pitch = p.pitch;
    T* beta = blk + (size_t)Leaf<T>::ROWS * pitch;
    double* tau = reinterpret_cast<double*>(beta + n);
    const T* A = p.A + bat * p.a_bs;
    {
        // one row per warp and iteration, lanes = columns (no index division); 8 rows in flight per warp
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
        for (int i0 = warp; i0 < mloc; i0 += 8 * nwarps) {
            T v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = i0 + e * nwarps;
                v[e] = Scalar<T>::zero();
                if (i < mloc && lane < n) v[e] = load_sum<T>(A + (r0 + i) * p.lda + lane, p.nsum, p.sum_stride);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = i0 + e * nwarps;
                if (i < mloc && lane < pitch) blk[i * pitch + lane] = v[e];     // padding columns (n .. pitch-1) zeroed
            }
        }
        if (pitch > 32) {
            for (int idx = threadIdx.x; idx < mloc * (pitch - 32); idx += blockDim.x) {
                const int i = idx / (pitch - 32), c = 32 + idx % (pitch - 32);
                blk[i * pitch + c] = Scalar<T>::zero();
            }
        }
    }
    __syncthreads();
    wqr_factor<T, Leaf<T>::RPL, 64>(blk, pitch, mloc, n, beta, tau, threadIdx.x >> 5, blockDim.x >> 5, 1);
    __syncthreads();
    T* V = p.V + bat * p.v_bs;
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
        if (lane < n)
            for (int i = warp; i < mloc; i += nwarps) V[(r0 + i) * n + lane] = blk[i * pitch + lane];
    }
    T* Rst = p.Rst + bat * p.rst_bs + (size_t)b * n * n;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int j = idx / n, c = idx - j * n;
        T v = Scalar<T>::zero();
        if (j < mloc) {
            if (c == j) v = beta[j];
            else if (c > j) v = blk[j * pitch + c];
        }
        Rst[idx] = v;
    }
    if (threadIdx.x < n) p.tau[bat * p.tau_bs + (size_t)b * n + threadIdx.x] = tau[threadIdx.x];
}

// ---- leaf apply: Q rows of one block = H_0 ... H_{n-1} [seed; 0] ----------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kLeafThreads) tsqr_leaf_apply_kernel(const TsqrParams<T> p) {
    constexpr int CH = LeafChunk<T>::CH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* blk = reinterpret_cast<T*>(smem_raw);
    const int b = blockIdx.x;
    const long long bat = blockIdx.y;
    const long long r0 = ((long long)b * p.m) / p.nblk, r1 = ((long long)(b + 1) * p.m) / p.nblk;
    const int mloc = (int)(r1 - r0), n = p.n, pitch = p.pitch;
    T* seed = blk + (size_t)Leaf<T>::ROWS * pitch;          // n x n, pitch n
    double* tau = reinterpret_cast<double*>(seed + n * n);
    const T* V = p.V + bat * p.v_bs;
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
        for (int i0 = warp; i0 < mloc; i0 += 8 * nwarps) {
            T v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = i0 + e * nwarps;
                v[e] = (i < mloc && lane < n) ? V[(r0 + i) * n + lane] : Scalar<T>::zero();
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = i0 + e * nwarps;
                if (i < mloc && lane < n) blk[i * pitch + lane] = v[e];
            }
        }
    }
    const T* M = p.M + bat * p.m_bs + (size_t)b * n * n;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) seed[idx] = M[idx];
    if (threadIdx.x < n) tau[threadIdx.x] = p.tau[bat * p.tau_bs + (size_t)b * n + threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    const int nch = (n + CH - 1) / CH;
    T* Q = p.Q + bat * p.q_bs;
    for (int ch = warp; ch < nch; ch += nwarps) {
        const int c0 = ch * CH;
        T reg[Leaf<T>::RPL][CH];
#pragma unroll
        for (int t = 0; t < Leaf<T>::RPL; ++t) {
            const int i = lane + 32 * t;
#pragma unroll
            for (int q = 0; q < CH; ++q)
                reg[t][q] = (i < n && i < mloc && c0 + q < n) ? seed[i * n + c0 + q] : Scalar<T>::zero();
        }
        wqr_apply_chunk<T, Leaf<T>::RPL, CH>(blk, pitch, mloc, min(mloc, n), tau, reg);
#pragma unroll
        for (int t = 0; t < Leaf<T>::RPL; ++t) {
            const int i = lane + 32 * t;
            if (i < mloc) {
#pragma unroll
                for (int q = 0; q < CH; ++q) {
                    const int c = c0 + q;
                    if (c < n) Q[(r0 + i) * p.ldq + c] = reg[t][q];
                    else if (c < p.qcols) Q[(r0 + i) * p.ldq + c] = Scalar<T>::zero();
                }
            }
        }
    }
    // zero fill of the remaining padding columns
    const int cz = nch * CH;
    if (p.qcols > cz) {
        const int wdt = p.qcols - cz;
        for (int idx = threadIdx.x; idx < mloc * wdt; idx += blockDim.x) {
            const int i = idx / wdt, c = cz + idx % wdt;
            Q[(r0 + i) * p.ldq + c] = Scalar<T>::zero();
        }
    }
}

// ---- whole QR inside one CTA (also the tree stage of the three-launch form) -----------------------------------
template <typename T>
__global__ void __launch_bounds__(TreeThreads<T>::N) tsqr_cta_kernel(const TsqrParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* panel = reinterpret_cast<T*>(smem_raw);
    const long long bat = blockIdx.x;
    const int m = (int)p.m, n = p.n, pitch = p.pitch;
    T* work = panel + (size_t)m * pitch;
    const T* A = p.A + bat * p.a_bs;
    for (int idx = threadIdx.x; idx < m * pitch; idx += blockDim.x) {
        const int i = idx / pitch, c = idx - i * pitch;
        panel[idx] = (c < n) ? load_sum<T>(A + (long long)i * p.lda + c, p.nsum, p.sum_stride) : Scalar<T>::zero();
    }
    __syncthreads();
    cta_qr<T>(panel, pitch, m, n, p.positive != 0, p.R ? p.R + bat * p.r_bs : nullptr, n,
              p.Q ? p.Q + bat * p.q_bs : nullptr, p.ldq, p.qcols, work);
}

template <typename T>
static size_t tsqr_cta_smem(int64_t m, int n) {
    const int pitch = wqr_pitch(n);
    return ((size_t)m * pitch + cta_qr_extra_elems<T>((int)m, n)) * sizeof(T) + 64;
}
template <typename T>
static size_t tsqr_leaf_smem(int n) {
    const int pitch = wqr_pitch(n);
    return ((size_t)Leaf<T>::ROWS * pitch + (size_t)n * n + 2 * n + 8) * sizeof(T) + 64;
}

template <typename T>
bool qr_fast_supported(qil_ctx* ctx, int64_t m, int64_t n) {
    if (n < 1 || n > kWqrMaxN || m < n) return false;
    if (m >= ((int64_t)1 << 30)) return false;
    return tsqr_leaf_smem<T>((int)n) <= std::min<size_t>(ctx->smem_optin, 225 * 1024);
}
template bool qr_fast_supported<double>(qil_ctx*, int64_t, int64_t);
template bool qr_fast_supported<cplx>(qil_ctx*, int64_t, int64_t);

template <typename T>
void qr_fast(qil_ctx* ctx, int64_t m, int n, const T* A, int64_t lda, int nsum, int64_t sum_stride, bool positive,
             T* Q, int64_t ldq, int qcols, T* R, int batch, int64_t a_bs, int64_t q_bs, int64_t r_bs) {
    QIL_REQUIRE(qr_fast_supported<T>(ctx, m, n), QIL_ERR_UNSUPPORTED, "qr_fast: %lld x %d not supported", (long long)m, n);
    QIL_REQUIRE(Q != nullptr, QIL_ERR_ARGUMENT, "qr_fast: Q is required");
    const size_t budget = std::min<size_t>(ctx->smem_optin, 225 * 1024);
    TsqrParams<T> p{};
    p.A = A; p.lda = lda; p.nsum = nsum; p.sum_stride = sum_stride; p.a_bs = a_bs;
    p.m = m; p.n = n; p.positive = positive ? 1 : 0; p.pitch = wqr_pitch(n);
    p.Q = Q; p.ldq = ldq; p.qcols = std::max(qcols, n); p.q_bs = q_bs;
    p.R = R; p.r_bs = r_bs;
    // one CTA only for a single block (<= 256 rows): the leaf kernels (256 threads, 255 registers, 8 warps per block,
    // trailing values kept in registers) are ~2x faster per column step than the multi-block levels of cta_qr
    if (m <= kWqrMaxRows && tsqr_cta_smem<T>(m, n) <= budget) {
        auto kern = tsqr_cta_kernel<T>;
        const size_t smem = tsqr_cta_smem<T>(m, n);
        ensure_dynamic_smem(kern, smem);
        kern<<<batch, TreeThreads<T>::N, smem, ctx->stream>>>(p);
        QIL_LAUNCH_CHECK(ctx);
        return;
    }
    const int nblk = (int)((m + Leaf<T>::ROWS - 1) / Leaf<T>::ROWS);
    const int64_t m2 = (int64_t)nblk * n;
    Mat<T> V(ctx, (int64_t)batch * m, n), Rst(ctx, (int64_t)batch * m2, n), Mq(ctx, (int64_t)batch * m2, n);
    Mat<double> tau(ctx, (int64_t)batch * nblk, n);
    p.nblk = nblk;
    p.V = V.p; p.v_bs = m * n;
    p.tau = tau.p; p.tau_bs = (int64_t)nblk * n;
    p.Rst = Rst.p; p.rst_bs = m2 * n;
    p.M = Mq.p; p.m_bs = m2 * n;
    const size_t smem = tsqr_leaf_smem<T>(n);
    {
        auto kern = tsqr_leaf_factor_kernel<T>;
        ensure_dynamic_smem(kern, smem);
        kern<<<dim3(nblk, batch), kLeafThreads, smem, ctx->stream>>>(p);
        QIL_LAUNCH_CHECK(ctx);
    }
    // tree: QR of the stacked triangles; its explicit Q seeds the leaf blocks
    qr_fast<T>(ctx, m2, n, Rst.p, n, 1, 0, positive, Mq.p, n, n, R, batch, m2 * n, m2 * n, r_bs);
    {
        auto kern = tsqr_leaf_apply_kernel<T>;
        ensure_dynamic_smem(kern, smem);
        kern<<<dim3(nblk, batch), kLeafThreads, smem, ctx->stream>>>(p);
        QIL_LAUNCH_CHECK(ctx);
    }
}
template void qr_fast<double>(qil_ctx*, int64_t, int, const double*, int64_t, int, int64_t, bool, double*, int64_t, int,
                              double*, int, int64_t, int64_t, int64_t);
template void qr_fast<cplx>(qil_ctx*, int64_t, int, const cplx*, int64_t, int, int64_t, bool, cplx*, int64_t, int, cplx*,
                            int, int64_t, int64_t, int64_t);

}  // namespace qil
