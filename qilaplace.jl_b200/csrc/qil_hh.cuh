// qil_hh.cuh -- Householder QR building blocks shared by hhqr_kernel (qil_qr.cu) and the fused small SVD
// (qil_svd_small.cu).  The block lives column-major in (shared or L2-resident) memory: column j at As + j*mpad.
//   H_j = I - s_j u_j u_j^H,  H_j x = beta_j e_1,  u_j stored in column j rows j.. (u_j[0] replaces the diagonal),
//   beta_j (the diagonal of R) in sbeta[j], s_j in ss[j].
// Latency matters more than flops here (a 256 x 20 block is ~40 dependent column steps), so
//   * every warp updates up to kHhIlp trailing columns at once (independent dot/shuffle chains overlap),
//   * the warp that owns column j+1 accumulates its norm while updating it and builds reflector j+1 right away:
//     one block barrier per column instead of two and no separate norm pass,
//   * explicit Q applies H_j to up to kHhIlp columns per warp at once.
#pragma once
#include "qil_common.cuh"

namespace qil {

constexpr int kHhIlp = 4;

template <typename T> __device__ __forceinline__ T hh_wsum(T v);
template <> __device__ __forceinline__ double hh_wsum<double>(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <> __device__ __forceinline__ cplx hh_wsum<cplx>(cplx v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}

// reflector from x0 = col[j] and xn2 = sum_{i>j} |col[i]|^2 ; executed by one lane
template <typename T>
__device__ __forceinline__ void hh_make_reflector(T* col, int j, double xn2, T* sbeta, double* ss) {
    const T x0 = col[j];
    const double a02 = Scalar<T>::abs2(x0);
    const double nx2 = a02 + xn2;
    if (nx2 == 0.0) {
        sbeta[j] = Scalar<T>::zero();
        ss[j] = 0.0;
        return;
    }
    // one rsqrt each for |x0| and |x| (sqrt(v) = v * rsqrt(v)), one division for s
    const double ra = (a02 > 0.0) ? rsqrt(a02) : 0.0;
    const double a0 = a02 * ra;
    const double nx = nx2 * rsqrt(nx2);
    const T ph = (a02 > 0.0) ? Scalar<T>::scale(x0, ra) : Scalar<T>::one();
    const T beta = Scalar<T>::scale(ph, -nx);
    sbeta[j] = beta;
    col[j] = Scalar<T>::sub(x0, beta);
    ss[j] = 1.0 / (nx * (nx + a0));
}

// In-place factorisation of the mloc x n block (k = min(mloc, n) reflectors).  All threads of the CTA call it.
template <typename T>
__device__ __forceinline__ void hh_factor(T* As, int mpad, int mloc, int n, int k, T* sbeta, double* ss) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nwarps = blockDim.x >> 5;
    if (warp == 0) {   // reflector 0
        double xn2 = 0.0;
        for (int i = 1 + lane; i < mloc; i += 32) xn2 += Scalar<T>::abs2(As[i]);
        xn2 = hh_wsum<double>(xn2);
        if (lane == 0) hh_make_reflector<T>(As, 0, xn2, sbeta, ss);
    }
    __syncthreads();
    for (int j = 0; j < k; ++j) {
        const T* col = As + j * mpad;
        const double s = ss[j];
        for (int cb = j + 1 + warp; cb < n; cb += kHhIlp * nwarps) {
            T w[kHhIlp];
#pragma unroll
            for (int q = 0; q < kHhIlp; ++q) w[q] = Scalar<T>::zero();
            if (s != 0.0) {
                int i = j + lane;
                for (; i + 96 < mloc; i += 128) {       // 4 row chunks in flight: loads first, then the FMAs
                    T uc[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) uc[r] = Scalar<T>::conj(col[i + 32 * r]);
#pragma unroll
                    for (int q = 0; q < kHhIlp; ++q) {
                        const int c = cb + q * nwarps;
                        if (c < n) {
                            T a[4];
#pragma unroll
                            for (int r = 0; r < 4; ++r) a[r] = As[c * mpad + i + 32 * r];
#pragma unroll
                            for (int r = 0; r < 4; ++r) w[q] = Scalar<T>::fma(uc[r], a[r], w[q]);
                        }
                    }
                }
                for (; i < mloc; i += 32) {
                    const T uc = Scalar<T>::conj(col[i]);
#pragma unroll
                    for (int q = 0; q < kHhIlp; ++q) {
                        const int c = cb + q * nwarps;
                        if (c < n) w[q] = Scalar<T>::fma(uc, As[c * mpad + i], w[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < kHhIlp; ++q) w[q] = Scalar<T>::scale(hh_wsum<T>(w[q]), -s);
            }
            // update; the owner of column j+1 (q == 0, cb == j+1) also accumulates its sub-diagonal norm
            double xn2 = 0.0;
            const bool owner = (cb == j + 1) && (j + 1 < k);
            int i2 = j + lane;
            for (; i2 + 96 < mloc; i2 += 128) {
                T u[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) u[r] = col[i2 + 32 * r];
#pragma unroll
                for (int q = 0; q < kHhIlp; ++q) {
                    const int c = cb + q * nwarps;
                    if (c < n) {
                        T* dst = As + c * mpad + i2;
                        T v[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) v[r] = dst[32 * r];
                        if (s != 0.0) {
#pragma unroll
                            for (int r = 0; r < 4; ++r) { v[r] = Scalar<T>::fma(w[q], u[r], v[r]); dst[32 * r] = v[r]; }
                        }
                        if (q == 0 && owner) {
#pragma unroll
                            for (int r = 0; r < 4; ++r) if (i2 + 32 * r > j + 1) xn2 += Scalar<T>::abs2(v[r]);
                        }
                    }
                }
            }
            for (; i2 < mloc; i2 += 32) {
                const T u = col[i2];
#pragma unroll
                for (int q = 0; q < kHhIlp; ++q) {
                    const int c = cb + q * nwarps;
                    if (c < n) {
                        T* dst = As + c * mpad + i2;
                        T v = *dst;
                        if (s != 0.0) { v = Scalar<T>::fma(w[q], u, v); *dst = v; }
                        if (q == 0 && owner && i2 > j + 1) xn2 += Scalar<T>::abs2(v);
                    }
                }
            }
            if (owner) {
                xn2 = hh_wsum<double>(xn2);
                __syncwarp();
                if (lane == 0) hh_make_reflector<T>(As + (j + 1) * mpad, j + 1, xn2, sbeta, ss);
            }
        }
        __syncthreads();
    }
}

// Explicit Q (mloc x k) from the reflectors.  Column c = H_0 ... H_c e_c.  Each warp builds up to kHhIlp columns at
// once in its private buffer qb (kHhIlp * mpad elements per warp), then `store(c, q)` is called by the whole warp
// for every finished column after a block barrier (so that storing may overwrite reflector storage in place,
// provided groups are processed from the highest columns down).
template <typename T, typename Store>
__device__ __forceinline__ void hh_form_q(const T* As, int mpad, int mloc, int k, const double* ss, T* qb_all,
                                          Store store) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nwarps = blockDim.x >> 5;
    T* qb = qb_all + warp * kHhIlp * mpad;
    const int per_group = kHhIlp * nwarps;
    const int ngroups = (k + per_group - 1) / per_group;
    for (int gi = ngroups - 1; gi >= 0; --gi) {
        const int c0 = gi * per_group + warp;            // this warp's columns: c0 + q * nwarps
        int cmax = -1;
#pragma unroll
        for (int q = 0; q < kHhIlp; ++q) {
            const int c = c0 + q * nwarps;
            if (c < k) {
                cmax = c;
                T* qq = qb + q * mpad;
                for (int i = lane; i < mloc; i += 32) qq[i] = (i == c) ? Scalar<T>::one() : Scalar<T>::zero();
            }
        }
        __syncwarp();
        for (int j = cmax; j >= 0; --j) {
            const double s = ss[j];
            if (s == 0.0) continue;
            const T* col = As + j * mpad;
            T w[kHhIlp];
#pragma unroll
            for (int q = 0; q < kHhIlp; ++q) w[q] = Scalar<T>::zero();
            bool act[kHhIlp];
#pragma unroll
            for (int q = 0; q < kHhIlp; ++q) act[q] = (c0 + q * nwarps >= j) && (c0 + q * nwarps < k);
            int i = j + lane;
            for (; i + 96 < mloc; i += 128) {
                T uc[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) uc[r] = Scalar<T>::conj(col[i + 32 * r]);
#pragma unroll
                for (int q = 0; q < kHhIlp; ++q)
                    if (act[q]) {
                        T a[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) a[r] = qb[q * mpad + i + 32 * r];
#pragma unroll
                        for (int r = 0; r < 4; ++r) w[q] = Scalar<T>::fma(uc[r], a[r], w[q]);
                    }
            }
            for (; i < mloc; i += 32) {
                const T uc = Scalar<T>::conj(col[i]);
#pragma unroll
                for (int q = 0; q < kHhIlp; ++q)
                    if (act[q]) w[q] = Scalar<T>::fma(uc, qb[q * mpad + i], w[q]);
            }
#pragma unroll
            for (int q = 0; q < kHhIlp; ++q) w[q] = Scalar<T>::scale(hh_wsum<T>(w[q]), -s);
            i = j + lane;
            for (; i + 96 < mloc; i += 128) {
                T u[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) u[r] = col[i + 32 * r];
#pragma unroll
                for (int q = 0; q < kHhIlp; ++q)
                    if (act[q]) {
                        T* dst = qb + q * mpad + i;
                        T v[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) v[r] = dst[32 * r];
#pragma unroll
                        for (int r = 0; r < 4; ++r) dst[32 * r] = Scalar<T>::fma(w[q], u[r], v[r]);
                    }
            }
            for (; i < mloc; i += 32) {
                const T u = col[i];
#pragma unroll
                for (int q = 0; q < kHhIlp; ++q)
                    if (act[q]) qb[q * mpad + i] = Scalar<T>::fma(w[q], u, qb[q * mpad + i]);
            }
            __syncwarp();
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kHhIlp; ++q) {
            const int c = c0 + q * nwarps;
            if (c < k) store(c, qb + q * mpad);
        }
        __syncthreads();
    }
}

}  // namespace qil
