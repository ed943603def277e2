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
#include "qil_wy.cuh"
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

// ---- whole tall QR in ONE cooperative launch -------------------------------------------------------------------
// Three (two) TSQR levels separated by a software grid barrier; the reflectors of every level stay in the shared
// memory of the CTA that produced them, only the n x n triangles (up) and the n x n seeds of the explicit Q (down)
// travel through L2:
//   A  CTA b < nb0 factors its block of <= R0 rows (8 warps, column parallel)         -> triangle to Rst0
//   B  CTA b < nb1 factors a block of <= R1 rows of the stacked level-0 triangles       -> triangle to Rst1
//   C  CTA nb0 (it holds no other block) factors the remaining <= R1 rows, writes R and its explicit Q -> M1
//   D  CTA b < nb1: rows of the level-1 Q = its reflectors applied to its seed M1[b]     -> M0
//   E  CTA b < nb0: rows of Q = level-0 reflectors applied to the seed M0[b]
// nb1 == 0: the level-0 triangles fit one block and B / D are skipped.
// The apply-down (C, D, E) uses the compact WY form  H_0 ... H_{n-1} = I - V T V^H  (LAPACK larft/larfb) instead of n
// dependent reflector applications: rows of Q = [S; 0] - V (T V1^H) S with V1 = V[0:n, :].  The n x n matrix T V1^H of
// a block does not depend on the seed S, so a CTA builds it between ARRIVING at the grid barrier after its factor and
// WAITING for it, i.e. while the levels above are being factored; what is left on the critical path of the way down
// is one n x n x n product and one row-parallel m x n x n product per level (no cross-lane reductions).
constexpr int kFusedThreads = 256;

struct TsqrFusedPlan {
    bool ok = false;
    int R0 = 0, R1 = 0, nb0 = 0, nb1 = 0;
    size_t smem = 0;
};

template <typename T>
struct TsqrFusedParams {
    const T* A; long long lda; int nsum; long long sum_stride;
    long long m; int n;
    int nb0, nb1, R0, R1;
    T* Rst0; T* Rst1;       // stacked triangles of level 0 / 1, [nb * n][n]
    T* M1; T* M0;           // explicit Q of the top block / of level 1 (the seeds of the level below), ld n
    T* Q; long long ldq; int qcols;
    T* R;
    int positive; int pitch;
    unsigned int* sync;     // {arrival counter, generation}, zero before the first use
    int warm;               // instruction-cache warm-up passes in idle CTAs (QIL_TSQR_WARM=0 disables)
    long long* clk;         // debug (QIL_TSQR_CLK=1): %globaltimer of CTA 0 / the last CTA at the phase boundaries
};
// split grid barrier: everything a CTA does between arrive and wait overlaps the other CTAs' way to the barrier
__device__ __forceinline__ unsigned int wq_grid_arrive(unsigned int* sync) {
    unsigned int g = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        g = *((volatile unsigned int*)(sync + 1));
        if (atomicAdd(sync, 1u) == gridDim.x - 1) {
            *((volatile unsigned int*)sync) = 0u;
            __threadfence();
            atomicAdd(sync + 1, 1u);
        }
    }
    return g;
}
__device__ __forceinline__ void wq_grid_wait(unsigned int* sync, unsigned int g) {
    if (threadIdx.x == 0) {
        while (*((volatile unsigned int*)(sync + 1)) == g) { }
        __threadfence();
    }
    __syncthreads();
}

// one block factored by the CTA.  Real panels: row-parallel register-resident Householder (rqr_factor, thread <-> row,
// the first ceil(m / 32) warps); complex panels: the column-parallel shared-memory form (all 8 warps, named barrier 1).
// `scr`: rqr_scratch_elems(8, NC) doubles (real only).
template <typename T>
__device__ __noinline__ void fused_factor(T* blk, int pitch, int m, int n, T* beta, double* tau, T* scr) {
    wqr_factor_any<T>(blk, pitch, m, n, beta, tau, threadIdx.x >> 5, blockDim.x >> 5, 1);
}
template <>
__device__ __noinline__ void fused_factor<double>(double* blk, int pitch, int m, int n, double* beta, double* tau,
                                                  double* scr) {
    rqr_factor_call(blk, pitch, m, n, beta, tau, scr, 1);     // the caller's __syncthreads() follows
}

// n x n triangle of a factored block (diag in beta) -> dst (ld n)
template <typename T>
__device__ __forceinline__ void fused_store_triangle(const T* blk, int pitch, int m, int n, const T* beta, T* dst) {
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int j = idx / n, c = idx - j * n;
        T v = Scalar<T>::zero();
        if (j < m) {
            if (c == j) v = beta[j];
            else if (c > j) v = blk[j * pitch + c];
        }
        dst[idx] = v;
    }
}

// rows [s0, s1) of a dense (ld n) matrix written by OTHER CTAs of this launch -> shared panel, padding columns zero
template <typename T>
__device__ __forceinline__ void fused_load_stack(const T* src, int s0, int rows, int n, T* blk, int pitch) {
    // one row per warp and slot, lanes = columns (no index division), 16 rows of every warp in flight (L2 round trips,
    // not bytes, are what this load costs)
    constexpr int kInFlight = Scalar<T>::is_complex ? 8 : 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const T* base = src + (size_t)s0 * n;
    for (int i0 = warp; i0 < rows; i0 += kInFlight * nwarps) {
        T v[kInFlight];
#pragma unroll
        for (int e = 0; e < kInFlight; ++e) {
            const int i = i0 + e * nwarps;
            v[e] = (i < rows && lane < n) ? __ldcg(base + (size_t)i * n + lane) : Scalar<T>::zero();
        }
#pragma unroll
        for (int e = 0; e < kInFlight; ++e) {
            const int i = i0 + e * nwarps;
            if (i < rows && lane < pitch) blk[i * pitch + lane] = v[e];
        }
    }
    if (pitch > 32) {
        for (int idx = threadIdx.x; idx < rows * (pitch - 32); idx += blockDim.x) {
            const int i = idx / (pitch - 32), c = 32 + idx % (pitch - 32);
            blk[i * pitch + c] = Scalar<T>::zero();
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kFusedThreads) tsqr_fused_kernel(const TsqrFusedParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n, pitch = p.pitch;
    T* blk0 = reinterpret_cast<T*>(smem_raw);
    T* blk1 = blk0 + (size_t)p.R0 * pitch;
    T* blkT = p.nb1 > 0 ? blk1 : blk0;          // the top CTA owns no level-0 block: two-level plans reuse that region
    T* beta0 = blk1 + (size_t)p.R1 * pitch;
    T* beta1 = beta0 + n;
    double* tau0 = reinterpret_cast<double*>(beta1 + n);
    double* tau1 = tau0 + n;
    // n x (n | 1) + four n x n matrices of the WY phases; the same region is the scratch of the real factor (none of the
    // matrices is live while a block is being factored)
    T* Tm = reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(tau1 + n) + 15) & ~(uintptr_t)15);   // 16-byte aligned
    T* TV0 = Tm + n * (n | 1) + (n & 1);
    T* TV1 = TV0 + n * n;
    T* S = TV1 + n * n;
    T* W2 = S + n * n;
    T* Gp = W2 + n * n;                                       // partial Gram tiles (real): the tail of the factor scratch
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    // CTA nb0 (the last one) owns no level-0 block: it factors the top block, so that no block's WY build sits between
    // the top factor and the way down
    const bool is_top = (b == p.nb0);
    const bool is_l1 = (p.nb1 > 0 && b < p.nb1);
    const long long r0 = is_top ? 0 : ((long long)b * p.m) / p.nb0, r1 = is_top ? 0 : ((long long)(b + 1) * p.m) / p.nb0;
    const int mloc = (int)(r1 - r0);
    fused_clk(p.clk, 0);
    // ---- A: level-0 block
    if (!is_top) {
    constexpr int kInFlight = Scalar<T>::is_complex ? 8 : 16;
    for (int i0 = warp; i0 < mloc; i0 += kInFlight * nwarps) {
        T v[kInFlight];
#pragma unroll
        for (int e = 0; e < kInFlight; ++e) {
            const int i = i0 + e * nwarps;
            v[e] = Scalar<T>::zero();
            if (i < mloc && lane < n) v[e] = load_sum<T>(p.A + (r0 + i) * p.lda + lane, p.nsum, p.sum_stride);
        }
#pragma unroll
        for (int e = 0; e < kInFlight; ++e) {
            const int i = i0 + e * nwarps;
            if (i < mloc && lane < pitch) blk0[i * pitch + lane] = v[e];
        }
    }
    if (pitch > 32) {
        for (int idx = threadIdx.x; idx < mloc * (pitch - 32); idx += blockDim.x) {
            const int i = idx / (pitch - 32), c = 32 + idx % (pitch - 32);
            blk0[i * pitch + c] = Scalar<T>::zero();
        }
    }
    __syncthreads();
    fused_clk(p.clk, 1);
    fused_factor<T>(blk0, pitch, mloc, n, beta0, tau0, Tm);
    __syncthreads();
    fused_clk(p.clk, 2);
    fused_store_triangle<T>(blk0, pitch, mloc, n, beta0, p.Rst0 + (size_t)b * n * n);
    }
    unsigned int tok = wq_grid_arrive(p.sync);
    if (!is_top && !is_l1) {
        wy_build<T>(blk0, pitch, mloc, n, tau0, Tm, TV0, Gp);
        // warm-up of the way down (see below): the rows of Q written here are written again in phase E
        if (p.warm) wy_apply<T>(blk0, pitch, mloc, n, TV0, S, W2, p.Q + r0 * p.ldq, p.ldq, p.qcols);
    }
    if (is_top && p.warm) {
        // Instruction-cache warm-up.  Every phase of this kernel is code its SM has not run yet (and the streaming passes
        // between two QRs flush it from L2), so a phase costs more in instruction fetch than in arithmetic.  The CTA of
        // the top block is idle until the levels below have been factored: it runs its whole phase (factor, WY build,
        // WY apply) once on a synthetic block of the same shape, so that the real pass finds the code in its caches.
        // Everything it writes (shared memory, M1 / M0) is overwritten by the real pass before anybody reads it.
        const int rowsW = (p.nb1 > 0 ? p.nb1 : p.nb0) * n;
        for (int idx = threadIdx.x; idx < rowsW * pitch; idx += blockDim.x) {
            const int i = idx / pitch, c = idx - i * pitch;
            blkT[idx] = (c < n) ? Scalar<T>::from_real(((i % n) == c ? 1.0 : 0.0) + 1e-3 * (double)((i * 7 + c * 3) % 11)) : Scalar<T>::zero();
        }
        __syncthreads();
        fused_factor<T>(blkT, pitch, rowsW, n, beta1, tau1, Tm);
        __syncthreads();
        for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
            const int j = idx / n, c = idx - j * n;
            S[idx] = (c == j) ? wqr_phase<T>(beta1[j]) : Scalar<T>::zero();
        }
        wy_build<T>(blkT, pitch, rowsW, n, tau1, Tm, TV1, Gp);
        wy_apply<T>(blkT, pitch, rowsW, n, TV1, S, W2, p.nb1 > 0 ? p.M1 : p.M0, n, n);
        __syncthreads();
    }
    wq_grid_wait(p.sync, tok);
    fused_clk(p.clk, 3);
    // ---- B: level-1 block
    const int rows1 = p.nb0 * n;
    int s0 = 0, m1 = 0;
    if (p.nb1 > 0) {
        if (is_l1) {
            s0 = (int)(((long long)b * rows1) / p.nb1);
            m1 = (int)(((long long)(b + 1) * rows1) / p.nb1) - s0;
            fused_load_stack<T>(p.Rst0, s0, m1, n, blk1, pitch);
            __syncthreads();
            fused_factor<T>(blk1, pitch, m1, n, beta1, tau1, Tm);
            __syncthreads();
            fused_store_triangle<T>(blk1, pitch, m1, n, beta1, p.Rst1 + (size_t)b * n * n);
        }
        tok = wq_grid_arrive(p.sync);
        if (is_l1) {
            wy_build<T>(blk1, pitch, m1, n, tau1, Tm, TV1, Gp);
            wy_build<T>(blk0, pitch, mloc, n, tau0, Tm, TV0, Gp);
            if (p.warm) {       // warm-up of phases D and E; both outputs are written again there
                wy_apply<T>(blk1, pitch, m1, n, TV1, S, W2, p.M0 + (size_t)s0 * n, n, n);
                wy_apply<T>(blk0, pitch, mloc, n, TV0, S, W2, p.Q + r0 * p.ldq, p.ldq, p.qcols);
                __syncthreads();
            }
        }
        wq_grid_wait(p.sync, tok);
    }
    fused_clk(p.clk, 4);
    // ---- C: top block, on the last CTA
    const int rowsT = (p.nb1 > 0 ? p.nb1 : p.nb0) * n;
    if (is_top) {
        fused_load_stack<T>(p.nb1 > 0 ? p.Rst1 : p.Rst0, 0, rowsT, n, blkT, pitch);
        __syncthreads();
        fused_clk(p.clk, 5);
        fused_factor<T>(blkT, pitch, rowsT, n, beta1, tau1, Tm);
        __syncthreads();
        fused_clk(p.clk, 6);
        for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
            const int j = idx / n, c = idx - j * n;
            const T ph = p.positive ? wqr_phase<T>(beta1[j]) : Scalar<T>::one();
            if (p.R) {
                T v = Scalar<T>::zero();
                if (c == j) v = beta1[j];
                else if (c > j) v = blkT[j * pitch + c];
                p.R[idx] = Scalar<T>::mul(Scalar<T>::conj(ph), v);
            }
            S[idx] = (c == j) ? ph : Scalar<T>::zero();            // seed of the top block: diag(phases)
        }
        fused_clk(p.clk, 14);
        wy_build<T>(blkT, pitch, rowsT, n, tau1, Tm, TV1, Gp, p.clk);
        fused_clk(p.clk, 11);
        wy_apply<T>(blkT, pitch, rowsT, n, TV1, S, W2, p.nb1 > 0 ? p.M1 : p.M0, n, n, p.clk);
    }
    fused_clk(p.clk, 7);
    tok = wq_grid_arrive(p.sync);
    wq_grid_wait(p.sync, tok);
    fused_clk(p.clk, 8);
    // ---- D: level-1 rows of Q
    if (p.nb1 > 0) {
        if (is_l1) {
            for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) S[idx] = __ldcg(p.M1 + (size_t)b * n * n + idx);
            __syncthreads();
            wy_apply<T>(blk1, pitch, m1, n, TV1, S, W2, p.M0 + (size_t)s0 * n, n, n);
        }
        tok = wq_grid_arrive(p.sync);
        wq_grid_wait(p.sync, tok);
    }
    // ---- E: rows of Q
    fused_clk(p.clk, 9);
    if (is_top) return;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) S[idx] = __ldcg(p.M0 + (size_t)b * n * n + idx);
    __syncthreads();
    wy_apply<T>(blk0, pitch, mloc, n, TV0, S, W2, p.Q + r0 * p.ldq, p.ldq, p.qcols);
    __syncthreads();
    fused_clk(p.clk, 10);
}

template <typename T>
static TsqrFusedPlan tsqr_fused_plan(qil_ctx* ctx, int64_t m, int n) {
    TsqrFusedPlan best;
    static const bool off = [] { const char* e = getenv("QIL_TSQR_FUSED"); return e && e[0] == '0'; }();
    if (off || !ctx->tsqr_fused_ok) return best;
    const size_t budget = std::min<size_t>(ctx->smem_optin, 225 * 1024);
    const int pitch = wqr_pitch(n);
    // rows one block may have: the register-resident real factor takes 768 (512 beyond 24 columns), the complex
    // shared-memory form 256
    const int maxrows = Scalar<T>::is_complex ? kWqrMaxRows : rqr_max_rows(n);
    auto small_elems = [&]() {
        size_t small = 5 * (size_t)n * (n + 1) + 8 + 256;         // elements of T: WY matrices + scratch of wy_t_blocked
        if (!Scalar<T>::is_complex) {
            const int mtiles = (n + 15) / 16;
            small += (size_t)(8 / mtiles) * 32 * (n | 1);         // partial Gram tiles of wy_gram behind the matrices
            small = std::max(small, rqr_scratch_elems(8, rqr_nc_for(n)) + 8);
        }
        return small;
    };
    // two levels: nb0 blocks of <= R0 rows, their triangles go straight to the top block (which sits in the level-0
    // region of its own CTA).  The smallest R0 that fits: most CTAs, shortest column steps.  Blocks stay <= 256 rows
    // although the factor could take 768: one SM has 64 FP64 FMA / clock, and a 640-row top block (two levels for a
    // 2^14-row panel) was measured SLOWER than three levels of small blocks (127 vs 116 us; per column step 3300 cycles
    // for 640 rows against 1840 for 128, and every product of the WY phase scales the same way).
    const int maxrows2 = std::min(maxrows, kWqrMaxRows);
    for (int R0 = 128; R0 <= maxrows2 && !best.ok; R0 *= 2) {
        const int64_t nb0 = (m + R0 - 1) / R0;
        if (nb0 < 2 || nb0 + 1 > ctx->sm_count) continue;
        const int64_t rowsT = nb0 * n;
        if (rowsT > maxrows2) continue;
        const int RA = (int)std::max<int64_t>(R0, rowsT);
        const size_t smem = ((size_t)RA * pitch + small_elems() + 4 * (size_t)n + 8) * sizeof(T) + 64;
        if (smem > budget) continue;
        best.ok = true;
        best.R0 = RA; best.R1 = 0; best.nb0 = (int)nb0; best.nb1 = 0; best.smem = smem;
    }
    // three levels
    const int r0s[2] = {128, 256}, r1s[2] = {256, 128};
    for (int a = 0; a < 2 && !best.ok; ++a) {
        for (int c = 0; c < 2 && !best.ok; ++c) {
            const int R0 = r0s[a], R1 = r1s[c];
            const int64_t nb0 = (m + R0 - 1) / R0;
            if (nb0 < 2 || nb0 + 1 > ctx->sm_count) continue;       // one CTA per SM, plus the CTA of the top block
            int64_t nb1 = 0;
            if (nb0 * n > R1) {
                nb1 = (nb0 * n + R1 - 1) / R1;
                if (nb1 * n > R1) continue;
            } else {
                continue;                                           // would have been a two-level plan
            }
            const size_t smem = ((size_t)(R0 + R1) * pitch + small_elems() + 4 * (size_t)n + 8) * sizeof(T) + 64;
            if (smem > budget) continue;
            best.ok = true;
            best.R0 = R0; best.R1 = R1; best.nb0 = (int)nb0; best.nb1 = (int)nb1; best.smem = smem;
        }
    }
    return best;
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
    if (batch == 1) {
        const TsqrFusedPlan fp = tsqr_fused_plan<T>(ctx, m, n);
        if (fp.ok) {
            TsqrFusedParams<T> f{};
            f.A = A; f.lda = lda; f.nsum = nsum; f.sum_stride = sum_stride;
            f.m = m; f.n = n; f.nb0 = fp.nb0; f.nb1 = fp.nb1; f.R0 = fp.R0; f.R1 = fp.R1;
            const int64_t rows1 = (int64_t)fp.nb0 * n, rows2 = (int64_t)std::max(fp.nb1, 1) * n;
            Mat<T> Rst0(ctx, rows1, n), M0(ctx, rows1, n), Rst1(ctx, rows2, n), M1(ctx, rows2, n);
            f.Rst0 = Rst0.p; f.M0 = M0.p; f.Rst1 = Rst1.p; f.M1 = M1.p;
            f.Q = Q; f.ldq = ldq; f.qcols = p.qcols; f.R = R; f.positive = p.positive; f.pitch = p.pitch;
            f.sync = ctx->get_grid_sync();
            static const bool warm = [] { const char* e = getenv("QIL_TSQR_WARM"); return !(e && e[0] == '0'); }();
            f.warm = warm ? 1 : 0;
            static const bool dbg_clk = [] { const char* e = getenv("QIL_TSQR_CLK"); return e && e[0] == '1'; }();
            Mat<long long> clk;
            if (dbg_clk) {
                clk = Mat<long long>(ctx, 96, 1);
                QIL_CUDA(cudaMemsetAsync(clk.p, 0, 96 * sizeof(long long), ctx->stream));
                f.clk = clk.p;
            }
            auto kern = tsqr_fused_kernel<T>;
            ensure_dynamic_smem(kern, fp.smem);
            void* args[] = {(void*)&f};
            QIL_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(fp.nb0 + 1), dim3(kFusedThreads), args, fp.smem,
                                                 ctx->stream));
            QIL_LAUNCH_CHECK(ctx);
            static const bool dbg_twice = [] { const char* e = getenv("QIL_TSQR_TWICE"); return e && e[0] == '1'; }();
            if (dbg_twice) {      // debug: the same launch again with warm instruction caches (same inputs, same outputs)
                QIL_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(fp.nb0 + 1), dim3(kFusedThreads), args, fp.smem,
                                                     ctx->stream));
            }
            if (dbg_clk) {
                long long h[96];
                QIL_CUDA(cudaMemcpyAsync(h, clk.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
                ctx->sync();
                fprintf(stderr, "[tsqr_fused] cta0 factor A: %lld ns = %lld cycles\n", h[2] - h[1], h[50] - h[49]);
                fprintf(stderr, "[tsqr_fused %lld x %d nb0 %d nb1 %d] cta0 ns:", (long long)m, n, fp.nb0, fp.nb1);
                for (int i = 1; i <= 10; ++i) fprintf(stderr, " %lld", h[i] ? h[i] - h[0] : -1);
                fprintf(stderr, " | last:");
                for (int i = 1; i <= 18; ++i) fprintf(stderr, " %lld", h[24 + i] ? h[24 + i] - h[24] : -1);
                fprintf(stderr, "\n");
            }
            return;
        }
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
