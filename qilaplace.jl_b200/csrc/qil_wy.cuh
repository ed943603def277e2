// qil_wy.cuh -- compact-WY form of a block of Householder reflectors held in shared memory (LAPACK larft / larfb):
//   H_0 ... H_{n-1} = I - V T V^H,   rows of  H [S; 0]  =  [S; 0] - V (T V1^H) S,   V1 = V[0:n, :]
// Used by the one-launch TSQR (qil_tsqr.cu) for its way down and by the one-CTA QR of the tree-node kernel (qil_node.cu)
// instead of n dependent reflector applications.  All functions are called by every thread of a 256-thread CTA and all
// pointers are shared memory unless stated otherwise.
#pragma once
#include "qil_wqr.cuh"

namespace qil {

// debug clocks of the fused TSQR (QIL_TSQR_CLK=1): %globaltimer / clock64 of CTA 0 and of the last CTA per phase
__device__ __forceinline__ void fused_clk(long long* clk, int slot) {
    if (clk && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory");
        clk[(blockIdx.x == 0 ? 0 : 24) + slot] = t;
        clk[48 + (blockIdx.x == 0 ? 0 : 24) + slot] = clock64();
    }
}

// T of the compact WY form from G (strict upper triangle of Tm, pitch pt), in place:  T = (striu(G) + diag(1 / tau))^-1,
// i.e. T[j][j] = tau_j, T[0:j, j] = -tau_j T[0:j, 0:j] G[0:j, j].  Column by column that is a chain of n dependent dot
// products (4.5 us for n = 20 with one warp, whatever the scheduling of its loads); blocked it is short:
//   8 x 8 diagonal blocks by one warp each (lane i keeps row i of its block in registers, no cross-lane traffic), then
//   T12 = -T11 (G12 T22) for neighbouring blocks, doubling the block size per level: two tiny products by all threads.
// X: scratch of >= 256 elements.  Called by all threads of the CTA.
template <typename T>
__device__ __forceinline__ void wy_t_blocked(T* Tm, int pt, int n, const double* tau, T* X) {
    constexpr int BS = 8;
    const int tid = threadIdx.x, nth = blockDim.x, warp = tid >> 5, lane = tid & 31;
    const int nblk = (n + BS - 1) / BS;
    if (warp < nblk) {
        const int lo = warp * BS, w = min(BS, n - lo);
        T row[BS];
#pragma unroll
        for (int j = 0; j < BS; ++j) {
            row[j] = Scalar<T>::zero();
            if (j < w) {
                const double tj = tau[lo + j];
                T acc = Scalar<T>::zero();
#pragma unroll
                for (int k = 0; k < j; ++k) acc = Scalar<T>::fma(row[k], Tm[(lo + k) * pt + lo + j], acc);
                row[j] = (lane < j) ? Scalar<T>::scale(acc, -tj) : (lane == j ? Scalar<T>::from_real(tj) : Scalar<T>::zero());
            }
        }
        __syncwarp();
        if (lane < w) {
#pragma unroll
            for (int j = 0; j < BS; ++j)
                if (j < w && j >= lane) Tm[(lo + lane) * pt + lo + j] = row[j];
        }
    }
    __syncthreads();
    for (int span = BS; span < n; span *= 2) {
        for (int lo = 0; lo + span < n; lo += 2 * span) {
            const int mid = lo + span, hi = min(lo + 2 * span, n);
            const int h = mid - lo, wd = hi - mid;
            for (int idx = tid; idx < h * wd; idx += nth) {            // X = G12 T22  (T22 upper triangular)
                const int i = lo + idx / wd, j = mid + idx % wd;
                T acc = Scalar<T>::zero();
                for (int k = mid; k <= j; ++k) acc = Scalar<T>::fma(Tm[i * pt + k], Tm[k * pt + j], acc);
                X[idx] = acc;
            }
            __syncthreads();
            for (int idx = tid; idx < h * wd; idx += nth) {            // T12 = -T11 X  (T11 upper triangular)
                const int ii = idx / wd, jj = idx % wd;
                T acc = Scalar<T>::zero();
                for (int k = ii; k < h; ++k) acc = Scalar<T>::fma(Tm[(lo + ii) * pt + lo + k], X[k * wd + jj], acc);
                Tm[(lo + ii) * pt + mid + jj] = Scalar<T>::scale(acc, -1.0);
            }
            __syncthreads();
        }
    }
}

// G = V^T V (real; n x n, pitch pt) on the FP64 tensor pipe with the m rows split over the warps: warp w takes row tile
// w % mtiles of G and the k-segment w / mtiles of the m rows, partial tiles go to `part` ([nseg][32][pt]) and are summed
// in segment order (deterministic).  cta_gemm alone would leave this to the one or two warps that own a row tile of G.
__device__ __forceinline__ void wy_gram(const double* V, int pitch, int m, int n, double* G, int pt, double* part) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int mtiles = (n + 15) >> 4, ntl = (n + 7) >> 3;
    const int nseg = nwarps / mtiles;
    const int mt = warp % mtiles, seg = warp / mtiles;
    const int ksteps = (m + 15) >> 4;
    if (seg < nseg) {
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0; }
        const int r0 = mt * 16 + g, r1 = r0 + 8;
        const int ks0 = (seg * ksteps) / nseg, ks1 = ((seg + 1) * ksteps) / nseg;
        for (int ks = ks0; ks < ks1; ++ks) {
            const int k0 = ks * 16;
            double af[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = k0 + 4 * t + i;
                af[2 * i] = (k < m && r0 < n) ? V[k * pitch + r0] : 0.0;
                af[2 * i + 1] = (k < m && r1 < n) ? V[k * pitch + r1] : 0.0;
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                if (nt < ntl) {
                    double bf[4];
                    const int col = nt * 8 + g;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int k = k0 + 4 * t + i;
                        bf[i] = (k < m && col < n) ? V[k * pitch + col] : 0.0;
                    }
                    wq_dmma(acc[nt], af, bf);
                }
            }
        }
        double* P = part + (size_t)seg * 32 * pt;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            if (nt < ntl) {
                const int c0 = nt * 8 + 2 * t;
                if (c0 < n) { P[r0 * pt + c0] = acc[nt][0]; P[r1 * pt + c0] = acc[nt][2]; }
                if (c0 + 1 < n) { P[r0 * pt + c0 + 1] = acc[nt][1]; P[r1 * pt + c0 + 1] = acc[nt][3]; }
            }
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int i = idx / n, j = idx - i * n;
        double a = 0.0;
        for (int sg = 0; sg < nseg; ++sg) a += part[((size_t)sg * 32 + i) * pt + j];
        G[i * pt + j] = a;
    }
}

// Compact WY of a factored block (m >= n rows; reflectors below the diagonal, heads on it):  TV = T V1^H  (n x n, ld n).
// The strict upper triangle of the block (the R entries, already stored elsewhere) is zeroed so that blk IS V.
//   G[i][j] = v_i^H v_j (i < j);  T[j][j] = tau_j,  T[0:j, j] = -tau_j T[0:j, 0:j] G[0:j, j]  (in place, one warp)
// Tm: n x (n | 1) scratch (odd pitch: lane <-> row of T is conflict free).  All pointers are shared memory.
template <typename T>
__device__ __noinline__ void wy_build(T* blk, int pitch, int m, int n, const double* tau, T* Tm, T* TV, T* gpart,
                                      long long* clk = nullptr) {
    __builtin_assume(__isShared(blk));
    __builtin_assume(__isShared(tau));
    __builtin_assume(__isShared(Tm));
    __builtin_assume(__isShared(TV));
    __builtin_assume(__isShared(gpart));
    const int tid = threadIdx.x, nth = blockDim.x;
    const int pt = n | 1;
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += nth) {
        const int i = idx / n, c = idx - i * n;
        if (c > i) blk[i * pitch + c] = Scalar<T>::zero();
    }
    __syncthreads();
    // Gram G = V^H V (only the strict upper triangle is used below)
    if (!Scalar<T>::is_complex) {
        // real: FP64 tensor pipe, fragments straight from the panel
        wy_gram(reinterpret_cast<const double*>(blk), pitch, m, n, reinterpret_cast<double*>(Tm), pt,
                reinterpret_cast<double*>(gpart));
    } else {
        // the same row r for all lanes of a warp (column index = lane-contiguous: conflict free); rows above the
        // diagonal contribute zeros
        for (int idx = tid; idx < n * n; idx += nth) {
            const int i = idx / n, j = idx - i * n;
            T a0 = Scalar<T>::zero(), a1 = a0, a2 = a0, a3 = a0;
            if (i < j) {
                const T* vi = blk + i;
                const T* vj = blk + j;
                int r = (i + 1) & ~3;                     // v_j is zero above row j > i; keep r aligned for all lanes
                for (; r + 3 < m; r += 4) {
                    const T x0 = vi[(r + 0) * pitch], x1 = vi[(r + 1) * pitch], x2 = vi[(r + 2) * pitch], x3 = vi[(r + 3) * pitch];
                    const T y0 = vj[(r + 0) * pitch], y1 = vj[(r + 1) * pitch], y2 = vj[(r + 2) * pitch], y3 = vj[(r + 3) * pitch];
                    a0 = Scalar<T>::fma(Scalar<T>::conj(x0), y0, a0);
                    a1 = Scalar<T>::fma(Scalar<T>::conj(x1), y1, a1);
                    a2 = Scalar<T>::fma(Scalar<T>::conj(x2), y2, a2);
                    a3 = Scalar<T>::fma(Scalar<T>::conj(x3), y3, a3);
                }
                for (; r < m; ++r) a0 = Scalar<T>::fma(Scalar<T>::conj(vi[r * pitch]), vj[r * pitch], a0);
            }
            Tm[i * pt + j] = Scalar<T>::add(Scalar<T>::add(a0, a1), Scalar<T>::add(a2, a3));
        }
    }
    __syncthreads();
    fused_clk(clk, 12);
    wy_t_blocked<T>(Tm, pt, n, tau, gpart);
    __syncthreads();
    fused_clk(clk, 13);
    for (int idx = tid; idx < n * n; idx += nth) {
        const int i = idx / n, c = idx - i * n;
        T acc = Scalar<T>::zero();
        for (int k = i; k <= c; ++k) acc = Scalar<T>::fma(Tm[i * pt + k], Scalar<T>::conj(blk[c * pitch + k]), acc);
        TV[idx] = acc;
    }
    __syncthreads();
}

// rows [0, m) of out (ld ldo) = [S; 0] - V (TV S);  S: n x n (shared, ld n), W2: n x n scratch.  Columns n .. ocols-1
// of out are zero filled.  One thread per (row, chunk of CW columns): no reductions across threads.
template <typename T>
__device__ __noinline__ void wy_apply(const T* blk, int pitch, int m, int n, const T* TV, const T* S, T* W2, T* out,
                                      long long ldo, int ocols, long long* clk = nullptr) {
    __builtin_assume(__isShared(blk));
    __builtin_assume(__isShared(TV));
    __builtin_assume(__isShared(S));
    __builtin_assume(__isShared(W2));
    constexpr int CW = 4;
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int idx = tid; idx < n * n; idx += nth) {
        const int i = idx / n, c = idx - i * n;
        T a0 = Scalar<T>::zero(), a1 = a0, a2 = a0, a3 = a0;
        int k = 0;
        for (; k + 3 < n; k += 4) {
            a0 = Scalar<T>::fma(TV[i * n + k], S[k * n + c], a0);
            a1 = Scalar<T>::fma(TV[i * n + k + 1], S[(k + 1) * n + c], a1);
            a2 = Scalar<T>::fma(TV[i * n + k + 2], S[(k + 2) * n + c], a2);
            a3 = Scalar<T>::fma(TV[i * n + k + 3], S[(k + 3) * n + c], a3);
        }
        for (; k < n; ++k) a0 = Scalar<T>::fma(TV[i * n + k], S[k * n + c], a0);
        W2[idx] = Scalar<T>::add(Scalar<T>::add(a0, a1), Scalar<T>::add(a2, a3));
    }
    __syncthreads();
    fused_clk(clk, 15);
    if (!Scalar<T>::is_complex) {
        // out = -V W2 on the FP64 tensor pipe (the zeroed upper triangle of V takes care of k > row), then + S on top
        cta_gemm(false, reinterpret_cast<const double*>(blk), pitch, m, n, reinterpret_cast<const double*>(W2), n, n,
                 reinterpret_cast<double*>(out), (int)ldo, -1.0);
        __syncthreads();
        fused_clk(clk, 16);
        for (int idx = tid; idx < n * n; idx += nth) {
            const int i = idx / n, c = idx - i * n;
            out[(long long)i * ldo + c] = Scalar<T>::add(out[(long long)i * ldo + c], S[idx]);
        }
    } else {
        const int nch = (n + CW - 1) / CW;
        for (int item = tid; item < m * nch; item += nth) {
            const int row = item % m, c0 = (item / m) * CW;
            T acc[CW];
#pragma unroll
            for (int q = 0; q < CW; ++q) acc[q] = Scalar<T>::zero();
            const T* v = blk + (size_t)row * pitch;
            const T* w = W2 + min(c0, n - CW < 0 ? 0 : n - CW);   // last chunk shifted left when n % CW != 0 ...
            const int sh = c0 - (int)(w - W2);                     // ... its first `sh` columns are duplicates
            const int kend = min(row + 1, n);                      // V[row][k] == 0 for k > row
            int k = 0;
            for (; k + 3 < kend; k += 4) {
                const T v0 = v[k], v1 = v[k + 1], v2 = v[k + 2], v3 = v[k + 3];
#pragma unroll
                for (int q = 0; q < CW; ++q) {
                    acc[q] = Scalar<T>::fma(v0, w[k * n + q], acc[q]);
                    acc[q] = Scalar<T>::fma(v1, w[(k + 1) * n + q], acc[q]);
                    acc[q] = Scalar<T>::fma(v2, w[(k + 2) * n + q], acc[q]);
                    acc[q] = Scalar<T>::fma(v3, w[(k + 3) * n + q], acc[q]);
                }
            }
            for (; k < kend; ++k) {
                const T vk = v[k];
#pragma unroll
                for (int q = 0; q < CW; ++q) acc[q] = Scalar<T>::fma(vk, w[k * n + q], acc[q]);
            }
#pragma unroll
            for (int q = 0; q < CW; ++q) {
                const int c = c0 - sh + q;
                if (q >= sh && c < n) {
                    const T s = (row < n) ? S[row * n + c] : Scalar<T>::zero();
                    out[(long long)row * ldo + c] = Scalar<T>::sub(s, acc[q]);
                }
            }
        }
    }
    if (ocols > n) {
        const int wdt = ocols - n;
        for (int idx = tid; idx < m * wdt; idx += nth) {
            const int i = idx / wdt, c = n + idx % wdt;
            out[(long long)i * ldo + c] = Scalar<T>::zero();
        }
    }
}


// the register-resident factor behind one call site (its own register allocation: up to 255, no spills)
static __device__ __noinline__ void rqr_factor_call(double* blk, int pitch, int m, int n, double* beta, double* tau, double* scr,
                                             int bar) {
    __builtin_assume(__isShared(blk));
    __builtin_assume(__isShared(beta));
    __builtin_assume(__isShared(tau));
    __builtin_assume(__isShared(scr));
    rqr_factor_any(blk, pitch, m, n, beta, tau, scr, bar);
}

// Thin QR of one real panel inside a CTA (256 threads): register-resident Householder (rqr_factor_any) + compact-WY
// formation of the explicit Q IN PLACE (panel: m x n, pitch, padding columns zero; m <= rqr_max_rows(n)).
//   Rout (n x n, ld ldr) upper triangular with `positive` phases, or nullptr.
//   work: cta_qr_fast_elems(n) doubles, 16-byte aligned.
__host__ __device__ inline size_t cta_qr_fast_elems(int n) {
    const int mtiles = (n + 15) / 16;
    size_t small = 5 * (size_t)n * (n + 1) + 8 + 256 + (size_t)(8 / mtiles) * 32 * (n | 1);
    const size_t scr = rqr_scratch_elems(8, rqr_nc_for(n)) + 8;
    if (small < scr) small = scr;
    return small + 2 * (size_t)n + 8;
}
static __device__ __noinline__ void cta_qr_fast(double* panel, int pitch, int m, int n, bool positive, double* Rout, int ldr,
                                         double* work) {
    __builtin_assume(__isShared(panel));
    __builtin_assume(__isShared(work));
    double* beta = work;
    double* tau = beta + n;
    double* Tm = tau + n + (n & 1);                           // 16-byte aligned when work is
    double* TV = Tm + n * (n | 1) + (n & 1);
    double* S = TV + 2 * n * n;                               // (TV1 slot unused)
    double* W2 = S + n * n;
    double* Gp = W2 + n * n;
    rqr_factor_call(panel, pitch, m, n, beta, tau, Tm, 1);
    __syncthreads();
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int j = idx / n, c = idx - j * n;
        const double ph = positive ? wqr_phase<double>(beta[j]) : 1.0;
        if (Rout) {
            double v = 0.0;
            if (c == j) v = beta[j];
            else if (c > j) v = panel[j * pitch + c];
            Rout[j * ldr + c] = ph * v;
        }
        S[idx] = (c == j) ? ph : 0.0;
    }
    wy_build<double>(panel, pitch, m, n, tau, Tm, TV, Gp);
    wy_apply<double>(panel, pitch, m, n, TV, S, W2, panel, pitch, n);
    __syncthreads();
}

}  // namespace qil
