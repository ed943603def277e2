// qil_rsvd_fast.cu -- the small-matrix end of a randomized split without host round trips (rsvd.jl:98-121):
//   B = Q^H A is l x C; its adjoint Bh = A^H Q comes out of the streaming kernel, Bh = Qb Rb by the fast TSQR, then
//   svd_finish   : one warp runs the one-sided Jacobi on G = Rb^H (l <= 32), sorts, applies the NDTensors
//                  truncation rule on the device and writes the two small factors,
//   rsvd_outputs : U = Q Us (rows x r) and S*Vh = (Us^H G) Qb^H (r x C), rank read from device memory.
#include "qil_fast.cuh"
#include "qil_wqr.cuh"

namespace qil {

template <typename T>
struct FinishParams {
    const T* Rb;        // l x l, ld l (upper triangular)
    int l;
    const double* scale;
    double cutoff;
    long long maxdim, mindim;
    T* Us;              // l x r, ld r
    T* T2;              // r x l, ld l
    double* S;          // l
    int* rank;
    double* margin;
    // batch (blockIdx.x): strides between problems
    long long rb_bs, us_bs, s_bs;
    int scale_bs, rank_bs;
};

template <typename T>
__global__ void __launch_bounds__(128) svd_finish_kernel(const FinishParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int l = p.l, pg = l | 1;
    T* G0 = reinterpret_cast<T*>(smem_raw);            // [l][pg] column-major
    T* Gw = G0 + l * pg;
    T* Us = Gw + l * pg;                               // [l][pg] row-major (i, j)
    double* sig = reinterpret_cast<double*>(Us + l * pg);
    int* order = reinterpret_cast<int*>(sig + l);
    __shared__ int s_rank;
    __shared__ double s_nu;
    const int tid = threadIdx.x;
    const long long bat = blockIdx.x;
    const double sc = p.scale ? p.scale[bat * p.scale_bs] : 1.0;
    const T* Rb = p.Rb + bat * p.rb_bs;
    T* gUs = p.Us + bat * p.us_bs;
    T* gT2 = p.T2 + bat * p.us_bs;
    double* gS = p.S + bat * p.s_bs;
    // G = scale * Rb^H: column j of G is the conjugated row j of Rb
    for (int idx = tid; idx < l * l; idx += blockDim.x) {
        const int j = idx / l, i = idx - j * l;
        G0[j * pg + i] = Scalar<T>::scale(Scalar<T>::conj(Rb[(size_t)j * l + i]), sc);
    }
    __syncthreads();
    cta_jacobi_rank<T>(Gw, G0, pg, l, p.cutoff, p.maxdim, p.mindim, sig, order, &s_rank, p.margin, &s_nu);
    const int r = s_rank;
    if (tid == 0) p.rank[bat * p.rank_bs] = r;
    for (int j = tid; j < l; j += blockDim.x) gS[j] = sig[j];
    // Us = W Sigma^-1 (sorted columns)
    for (int idx = tid; idx < l * r; idx += blockDim.x) {
        const int i = idx / r, j = idx - i * r;
        const double sj = sig[j];
        const T v = Scalar<T>::scale(Gw[order[j] * pg + i], sj > 0.0 ? 1.0 / sj : 0.0);
        Us[i * pg + j] = v;
        gUs[(size_t)i * r + j] = v;
    }
    __syncthreads();
    // T2 = Us^H G  (r x l)
    for (int idx = tid; idx < r * l; idx += blockDim.x) {
        const int j = idx / l, c = idx - j * l;
        T acc = Scalar<T>::zero();
        for (int k = 0; k < l; ++k) acc = Scalar<T>::fma(Scalar<T>::conj(Us[k * pg + j]), G0[c * pg + k], acc);
        gT2[(size_t)j * l + c] = acc;
    }
}

template <typename T>
void svd_finish(qil_ctx* ctx, int l, const T* Rb, const double* d_scale, double cutoff, int64_t maxdim, int64_t mindim,
                T* Us, T* T2, double* S, int* d_rank, int batch, int scale_bs, int rank_bs) {
    QIL_REQUIRE(l >= 1 && l <= kWqrMaxN, QIL_ERR_UNSUPPORTED, "svd_finish: l = %d", l);
    FinishParams<T> p;
    p.Rb = Rb; p.l = l; p.scale = d_scale; p.cutoff = cutoff; p.maxdim = maxdim; p.mindim = mindim;
    p.Us = Us; p.T2 = T2; p.S = S; p.rank = d_rank; p.margin = ctx->d_margin;
    p.rb_bs = (long long)l * l; p.us_bs = (long long)l * l; p.s_bs = l; p.scale_bs = scale_bs; p.rank_bs = rank_bs;
    const int pg = l | 1;
    const size_t smem = (size_t)3 * l * pg * sizeof(T) + (size_t)l * (sizeof(double) + sizeof(int)) + 64;
    svd_finish_kernel<T><<<batch, 128, smem, ctx->stream>>>(p);
    QIL_LAUNCH_CHECK(ctx);
}
template void svd_finish<double>(qil_ctx*, int, const double*, const double*, double, int64_t, int64_t, double*, double*,
                                 double*, int*, int, int, int);
template void svd_finish<cplx>(qil_ctx*, int, const cplx*, const double*, double, int64_t, int64_t, cplx*, cplx*, double*,
                               int*, int, int, int);

// one thread per output element; Us / T2 staged in shared memory
template <typename T>
__global__ void __launch_bounds__(256) rsvd_outputs_kernel(long long R, long long C, int l, const T* __restrict__ Q,
                                                           long long ldq, const T* __restrict__ Qb, long long ldqb,
                                                           const T* __restrict__ Us, const T* __restrict__ T2,
                                                           const double* __restrict__ S, const int* __restrict__ d_rank,
                                                           T* __restrict__ U, T* __restrict__ SVh, T* __restrict__ Vh,
                                                           long long q_bs, long long qb_bs, long long u_bs,
                                                           long long sv_bs, int rank_bs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    {   // batch member blockIdx.y
        const long long bat = blockIdx.y;
        Q += bat * q_bs; Qb += bat * qb_bs; Us += bat * (long long)l * l; T2 += bat * (long long)l * l; S += bat * l;
        d_rank += bat * rank_bs;
        if (U) U += bat * u_bs;
        if (SVh) SVh += bat * sv_bs;
        if (Vh) Vh += bat * sv_bs;
    }
    const int r = *d_rank;
    T* sU = reinterpret_cast<T*>(smem_raw);     // [l][r]
    T* sT = sU + l * r;                         // [r][l]
    double* sS = reinterpret_cast<double*>(sT + r * l);
    for (int idx = threadIdx.x; idx < l * r; idx += blockDim.x) { sU[idx] = Us[idx]; sT[idx] = T2[idx]; }
    for (int j = threadIdx.x; j < r; j += blockDim.x) sS[j] = S[j];
    __syncthreads();
    const long long gstride = (long long)gridDim.x * blockDim.x;
    if (U) {
        for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < R; row += gstride) {
            const T* qr = Q + row * ldq;
            T q[kWqrMaxN];
#pragma unroll
            for (int i = 0; i < kWqrMaxN; ++i) q[i] = (i < l) ? qr[i] : Scalar<T>::zero();
            for (int j = 0; j < r; ++j) {
                T acc = Scalar<T>::zero();
#pragma unroll
                for (int i = 0; i < kWqrMaxN; ++i)
                    if (i < l) acc = Scalar<T>::fma(q[i], sU[i * r + j], acc);
                U[row * r + j] = acc;
            }
        }
    }
    if (SVh || Vh) {
        for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gstride) {
            const T* qr = Qb + c * ldqb;
            T q[kWqrMaxN];
#pragma unroll
            for (int i = 0; i < kWqrMaxN; ++i) q[i] = (i < l) ? Scalar<T>::conj(qr[i]) : Scalar<T>::zero();
            for (int j = 0; j < r; ++j) {
                T acc = Scalar<T>::zero();
#pragma unroll
                for (int i = 0; i < kWqrMaxN; ++i)
                    if (i < l) acc = Scalar<T>::fma(sT[j * l + i], q[i], acc);
                if (SVh) SVh[(long long)j * C + c] = acc;
                if (Vh) {
                    const double sj = sS[j];
                    Vh[(long long)j * C + c] = Scalar<T>::scale(acc, sj > 0.0 ? 1.0 / sj : 0.0);
                }
            }
        }
    }
}

template <typename T>
void rsvd_outputs(qil_ctx* ctx, int64_t R, int64_t C, int l, const T* Q, int64_t ldq, const T* Qb, int64_t ldqb,
                  const T* Us, const T* T2, const double* S, const int* d_rank, T* U, T* SVh, T* Vh, int batch,
                  int64_t q_bs, int64_t qb_bs, int64_t u_bs, int64_t sv_bs, int rank_bs) {
    const size_t smem = (size_t)2 * l * l * sizeof(T) + (size_t)l * sizeof(double) + 32;
    const int64_t work = std::max(R, C);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)ctx->sm_count * 4));
    rsvd_outputs_kernel<T><<<dim3(grid, batch), 256, smem, ctx->stream>>>(R, C, l, Q, ldq, Qb, ldqb, Us, T2, S, d_rank, U,
                                                                       SVh, Vh, q_bs, qb_bs, u_bs, sv_bs, rank_bs);
    QIL_LAUNCH_CHECK(ctx);
}
template void rsvd_outputs<double>(qil_ctx*, int64_t, int64_t, int, const double*, int64_t, const double*, int64_t,
                                   const double*, const double*, const double*, const int*, double*, double*, double*, int,
                                   int64_t, int64_t, int64_t, int64_t, int);
template void rsvd_outputs<cplx>(qil_ctx*, int64_t, int64_t, int, const cplx*, int64_t, const cplx*, int64_t, const cplx*,
                                 const cplx*, const double*, const int*, cplx*, cplx*, cplx*, int, int64_t, int64_t, int64_t,
                                 int64_t, int);

}  // namespace qil
