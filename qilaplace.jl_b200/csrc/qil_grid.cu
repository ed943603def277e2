// qil_grid.cu -- dense coefficient grids: every combination of a set of "free" sites at once.
//
// Replaces the callers that loop `coefficient` over a regular grid of bitstrings
//   docs/src/tutorials/zt.jl:152-157, 283-287, 330-411   (k, l) pole scans on the zT output
//   docs/src/tutorials/signal.jl:149-150, 227-228          all 2^n coefficients
// and `mps_to_vector` (src/mps.jl:716-743), which is the all-sites-free case.
//
// Instead of 2^F independent chains the grid is evaluated by meet in the middle.  With a cut after site m
//   L[cL][:]  = prod_{i<m}  A_i[:, b_i, :]      one row per combination cL of the free sites left of the cut
//   R[:][cR]  = prod_{i>=m} A_i[:, b_i, :]      one column per combination cR of the free sites right of it
//   out[cL * 2^fR + cR] = amplitude * L[cL] . R[:, cR]
// A free site doubles the row (column) count; because cores are stored [l][s][r] row-major the doubling step
// is one plain GEMM on the untouched core:  L' (2c x chi_r) == L (c x chi_l) * core viewed as chi_l x (2 chi_r),
// R' (chi_l x 2c) == core viewed as (2 chi_l) x chi_r  *  R (chi_r x c).  The cut is chosen by a flop model.
// All GEMMs run on the FP64 tensor path (DMMA m16n8k16); complex ones through the real 2x2 embedding
// [Ar Ai] * [[Br Bi], [-Bi Br]], operands staged with cp.async into an XOR-swizzled 3-stage ring.
#include "qil_dense.cuh"
#include "qil_mpsops.cuh"

#include <algorithm>

namespace qil {

constexpr int kTcThreads = 256;             // 8 warps: (BM / 16) row tiles x (128 / BM) column groups
constexpr int kTcStages = 3;
constexpr int kTcBBytes = 32 * 128 * 8;     // real: 32 k x 128 cols; complex: 16 k x 64 complex cols (half used)
constexpr int tc_a_bytes(int bm) { return bm * 32 * 8; }   // BM rows x 32 doubles (= 16 complex) per k-chunk

struct TcParams {
    const void* A;
    const void* B;
    void* C;
    long long M, N, K;       // in elements of the scalar type
    long long lda, ldb, ldc;
    double alpha;
};

__device__ __forceinline__ void tc_cp_async(void* dst, const void* src, int bytes, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = valid ? bytes : 0;
    if (bytes == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void tc_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tc_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tc_dmma(double* c, const double* a, const double* b) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
        "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
          "d"(b[2]), "d"(b[3]));
}

// C (M x N) = alpha * A (M x K) * B (K x N), row-major with leading dimensions, T = double or cplx.
// CTA tile BM x (128 real | 64 complex) columns, BM in {16, 32, 64, 128}: the 8 warps form BM/16 row tiles times
// 128/BM column groups, so a skinny left operand (few grid rows) neither pads to 128 rows nor leaves SMs idle.
template <bool CPLX, int BM>
__global__ void __launch_bounds__(kTcThreads, 1) gemm_tc_kernel(const TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr int kTcABytes = tc_a_bytes(BM);
    constexpr int WM = BM / 16;            // warps along the rows
    constexpr int WN = 8 / WM;             // warps along the columns
    constexpr int NT = 16 / WN;            // n8 tiles per warp
    unsigned char* sA = smem_raw;
    unsigned char* sB = sA + kTcStages * kTcABytes;
    constexpr int KC = CPLX ? 16 : 32;     // elements along k per chunk
    constexpr int NC = CPLX ? 64 : 128;    // elements along n per tile
    constexpr int ES = CPLX ? 16 : 8;      // element size in bytes

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const long long m0 = (long long)blockIdx.y * BM, n0 = (long long)blockIdx.x * NC;
    const int wm = warp % WM, wn = warp / WM;
    const unsigned char* Ab = reinterpret_cast<const unsigned char*>(p.A);
    const unsigned char* Bb = reinterpret_cast<const unsigned char*>(p.B);
    const int iters = (int)((p.K + KC - 1) / KC);

    auto issue = [&](int it) {
        const int stage = it % kTcStages;
        const long long k0 = (long long)it * KC;
        unsigned char* a = sA + stage * kTcABytes;
        unsigned char* b = sB + stage * kTcBBytes;
#pragma unroll
        for (int e = 0; e < (BM * KC) / kTcThreads; ++e) {
            const int idx = e * kTcThreads + tid;
            const int row = idx / KC, j = idx % KC;
            const bool ok = (m0 + row < p.M) && (k0 + j < p.K);
            const unsigned char* src = ok ? Ab + ((m0 + row) * p.lda + k0 + j) * ES : Ab;
            if (CPLX) tc_cp_async(a + row * 256 + ((j ^ (row & 7)) << 4), src, 16, ok);
            else      tc_cp_async(a + row * 256 + (((j >> 1) ^ (row & 7)) << 4) + (j & 1) * 8, src, 8, ok);
        }
#pragma unroll
        for (int e = 0; e < (KC * NC) / kTcThreads; ++e) {
            const int idx = e * kTcThreads + tid;
            const int r = idx / NC, c = idx % NC;
            const bool ok = (k0 + r < p.K) && (n0 + c < p.N);
            const unsigned char* src = ok ? Bb + ((k0 + r) * p.ldb + n0 + c) * ES : Bb;
            if (CPLX) tc_cp_async(b + r * 1024 + ((c ^ ((r >> 1) & 7)) << 4), src, 16, ok);
            else      tc_cp_async(b + r * 1024 + ((c ^ (((r >> 2) & 3) << 2)) << 3), src, 8, ok);
        }
    };

    for (int it = 0; it < kTcStages - 1; ++it) {
        if (it < iters) issue(it);
        tc_commit();
    }
    double acc[NT][4];
#pragma unroll
    for (int x = 0; x < NT; ++x) { acc[x][0] = acc[x][1] = acc[x][2] = acc[x][3] = 0.0; }
    const int R0 = wm * 16 + g, R1 = R0 + 8;
    const int nt0 = wn * NT;               // first n8 tile of this warp

    for (int it = 0; it < iters; ++it) {
        tc_wait<kTcStages - 2>();
        __syncthreads();
        if (it + kTcStages - 1 < iters) issue(it + kTcStages - 1);
        tc_commit();
        const int stage = it % kTcStages;
        const unsigned char* a = sA + stage * kTcABytes;
        const unsigned char* b = sB + stage * kTcBBytes;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
            double af[8];
            const int j0 = kb * 8 + 2 * t;
            const double2 x00 = *reinterpret_cast<const double2*>(a + R0 * 256 + (((j0) ^ (R0 & 7)) << 4));
            const double2 x01 = *reinterpret_cast<const double2*>(a + R0 * 256 + (((j0 + 1) ^ (R0 & 7)) << 4));
            const double2 x10 = *reinterpret_cast<const double2*>(a + R1 * 256 + (((j0) ^ (R1 & 7)) << 4));
            const double2 x11 = *reinterpret_cast<const double2*>(a + R1 * 256 + (((j0 + 1) ^ (R1 & 7)) << 4));
            af[0] = x00.x; af[2] = x00.y; af[4] = x01.x; af[6] = x01.y;
            af[1] = x10.x; af[3] = x10.y; af[5] = x11.x; af[7] = x11.y;
            if (CPLX) {
                const int lr0 = kb * 8 + 2 * t, lr1 = lr0 + 1;
                const unsigned char* b0p = b + lr0 * 1024;
                const unsigned char* b1p = b + lr1 * 1024;
                const int sw = (lr0 >> 1) & 7;      // == (lr1 >> 1) & 7
#pragma unroll
                for (int x = 0; x < NT; ++x) {
                    const int nt = nt0 + x;
                    const int c = nt * 4 + (g >> 1);
                    const double2 e0 = *reinterpret_cast<const double2*>(b0p + ((c ^ sw) << 4));
                    const double2 e1 = *reinterpret_cast<const double2*>(b1p + ((c ^ sw) << 4));
                    double bf[4];
                    if (g & 1) { bf[0] = e0.y; bf[1] = e0.x; bf[2] = e1.y; bf[3] = e1.x; }
                    else       { bf[0] = e0.x; bf[1] = -e0.y; bf[2] = e1.x; bf[3] = -e1.y; }
                    tc_dmma(acc[x], af, bf);
                }
            } else {
                const int r0 = kb * 16 + 4 * t;
                const int sw = ((r0 >> 2) & 3) << 2;   // same for r0 .. r0+3
                const unsigned char* bp = b + r0 * 1024;
#pragma unroll
                for (int x = 0; x < NT; ++x) {
                    const int nt = nt0 + x;
                    const int c = ((nt * 8 + g) ^ sw) << 3;
                    double bf[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) bf[i] = *reinterpret_cast<const double*>(bp + i * 1024 + c);
                    tc_dmma(acc[x], af, bf);
                }
            }
        }
    }
    tc_wait<0>();

    // epilogue: acc[nt][0..1] -> row R0, real columns nt*8 + 2t, +1 ; acc[nt][2..3] -> row R1
    const long long gr0 = m0 + R0, gr1 = m0 + R1;
    if (CPLX) {
        cplx* C = reinterpret_cast<cplx*>(p.C);
#pragma unroll
        for (int x = 0; x < NT; ++x) {
            const long long c = n0 + (nt0 + x) * 4 + t;
            if (c < p.N) {
                if (gr0 < p.M) C[gr0 * p.ldc + c] = make_double2(acc[x][0] * p.alpha, acc[x][1] * p.alpha);
                if (gr1 < p.M) C[gr1 * p.ldc + c] = make_double2(acc[x][2] * p.alpha, acc[x][3] * p.alpha);
            }
        }
    } else {
        double* C = reinterpret_cast<double*>(p.C);
#pragma unroll
        for (int x = 0; x < NT; ++x) {
            const long long c = n0 + (nt0 + x) * 8 + 2 * t;
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (c + h < p.N) {
                    if (gr0 < p.M) C[gr0 * p.ldc + c + h] = acc[x][h] * p.alpha;
                    if (gr1 < p.M) C[gr1 * p.ldc + c + h] = acc[x][2 + h] * p.alpha;
                }
        }
    }
}

template <bool CP, int BM>
static void launch_gemm_tc(qil_ctx* ctx, const TcParams& p) {
    constexpr int NC = CP ? 64 : 128;
    const size_t smem = (size_t)kTcStages * (tc_a_bytes(BM) + kTcBBytes);
    auto kern = gemm_tc_kernel<CP, BM>;
    ensure_dynamic_smem(kern, smem);
    const long long gx = (p.N + NC - 1) / NC, gy = (p.M + BM - 1) / BM;
    QIL_REQUIRE(gy <= 65535, QIL_ERR_UNSUPPORTED, "gemm_tc: %lld row tiles exceed the grid limit", gy);
    dim3 grid((unsigned)gx, (unsigned)gy);
    kern<<<grid, kTcThreads, smem, ctx->stream>>>(p);
    QIL_LAUNCH_CHECK(ctx);
}

template <typename T>
void gemm_tc(qil_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha, const T* A, int64_t lda, const T* B,
             int64_t ldb, T* C, int64_t ldc) {
    if (M == 0 || N == 0) return;
    constexpr bool CP = Scalar<T>::is_complex;
    constexpr int NC = CP ? 64 : 128;
    TcParams p{A, B, C, M, N, K, lda, ldb, ldc, alpha};
    // tallest row tile that still gives every SM a CTA (B operand re-reads grow as the tile shrinks)
    const long long gx = (N + NC - 1) / NC;
    int bm = 128;
    while (bm > 16 && (bm / 2 >= M || gx * ((M + bm - 1) / bm) < ctx->sm_count)) bm >>= 1;
    switch (bm) {
        case 128: launch_gemm_tc<CP, 128>(ctx, p); break;
        case 64: launch_gemm_tc<CP, 64>(ctx, p); break;
        case 32: launch_gemm_tc<CP, 32>(ctx, p); break;
        default: launch_gemm_tc<CP, 16>(ctx, p); break;
    }
}
template void gemm_tc<double>(qil_ctx*, int64_t, int64_t, int64_t, double, const double*, int64_t, const double*,
                              int64_t, double*, int64_t);
template void gemm_tc<cplx>(qil_ctx*, int64_t, int64_t, int64_t, double, const cplx*, int64_t, const cplx*, int64_t,
                            cplx*, int64_t);

// ------------------------------------------------------------------------------------------------
// single-vector steps of a fixed-site chain (no grid row / column yet): bandwidth-bound GEMVs
// ------------------------------------------------------------------------------------------------
// y[i] = sum_j A[i * lda + j] * x[j],  i < rows, j < cols : one warp per row, lanes along the row
template <typename T>
__global__ void __launch_bounds__(256) gemv_rows_kernel(const T* __restrict__ A, long long lda, int rows, int cols,
                                                        const T* __restrict__ x, T* __restrict__ y) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    if (i >= rows) return;
    const T* a = A + (long long)i * lda;
    T acc = Scalar<T>::zero();
    for (int j = lane; j < cols; j += 32) acc = Scalar<T>::fma(a[j], x[j], acc);
    double* av = reinterpret_cast<double*>(&acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int c = 0; c < (int)(sizeof(T) / sizeof(double)); ++c) av[c] += __shfl_xor_sync(0xffffffffu, av[c], off);
    if (lane == 0) y[i] = acc;
}
// y[j] = sum_i x[i] * A[i * lda + j],  i < rows, j < cols : 32 columns per CTA, the 8 warps split the rows
template <typename T>
__global__ void __launch_bounds__(256) gemv_cols_kernel(const T* __restrict__ A, long long lda, int rows, int cols,
                                                        const T* __restrict__ x, T* __restrict__ y) {
    __shared__ T part[8][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 32 + lane;
    T acc = Scalar<T>::zero();
    if (j < cols)
        for (int i = warp; i < rows; i += 8) acc = Scalar<T>::fma(x[i], A[(long long)i * lda + j], acc);
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && j < cols) {
        T s = part[0][lane];
#pragma unroll
        for (int w = 1; w < 8; ++w) s = Scalar<T>::add(s, part[w][lane]);
        y[j] = s;
    }
}
template <typename T>
static void gemv_rows(qil_ctx* ctx, const T* A, int64_t lda, int64_t rows, int64_t cols, const T* x, T* y) {
    gemv_rows_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, ctx->stream>>>(A, lda, (int)rows, (int)cols, x, y);
    QIL_LAUNCH_CHECK(ctx);
}
template <typename T>
static void gemv_cols(qil_ctx* ctx, const T* A, int64_t lda, int64_t rows, int64_t cols, const T* x, T* y) {
    gemv_cols_kernel<T><<<(unsigned)((cols + 31) / 32), 256, 0, ctx->stream>>>(A, lda, (int)rows, (int)cols, x, y);
    QIL_LAUNCH_CHECK(ctx);
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void fill_one_kernel(T* p) {
    if (threadIdx.x == 0 && blockIdx.x == 0) p[0] = Scalar<T>::one();
}

struct BitPerm {
    int nbits;
    int dst[64];   // bit j of the source index (j = 0 is the LEAST significant) goes to bit dst[j] of the destination
};
template <typename T>
__global__ void bit_permute_kernel(const BitPerm bp, long long total, const T* __restrict__ in, T* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long o = 0;
        for (int j = 0; j < bp.nbits; ++j) o |= ((unsigned long long)((i >> j) & 1)) << bp.dst[j];
        out[o] = in[i];
    }
}

// ------------------------------------------------------------------------------------------------
// grid evaluation
// ------------------------------------------------------------------------------------------------
template <typename T>
static void coefficient_grid_impl(qil_ctx* ctx, const qil_mps* psi, const uint8_t* mode, const int32_t* out_bit,
                                  T* d_out) {
    const int n = psi->n;
    const std::vector<int64_t>& bond = psi->bond;
    int F = 0;
    for (int i = 0; i < n; ++i) F += (mode[i] == 2);
    QIL_REQUIRE(F <= 40, QIL_ERR_UNSUPPORTED, "coefficient grid: %d free sites is more than this build addresses", F);

    // ---- choose the cut by a flop model (+ a per-launch constant so that tiny chains are not over-weighted)
    std::vector<int> freeL(n + 1, 0);
    for (int i = 0; i < n; ++i) freeL[i + 1] = freeL[i] + (mode[i] == 2);
    const double launch_cost = 5e7;   // ~ 5 us at 10 TFLOP/s
    double best = -1.0;
    int m = n;
    for (int cut = 0; cut <= n; ++cut) {
        double cost = 0.0;
        for (int i = 0; i < cut; ++i)
            cost += launch_cost + std::ldexp(1.0, freeL[i]) * bond[i] * bond[i + 1] * (mode[i] == 2 ? 2.0 : 1.0);
        for (int i = cut; i < n; ++i)
            cost += launch_cost + std::ldexp(1.0, F - freeL[i + 1]) * bond[i] * bond[i + 1] * (mode[i] == 2 ? 2.0 : 1.0);
        cost += launch_cost + std::ldexp(1.0, F) * bond[cut];
        // memory of the two panels
        const double mem = (std::ldexp(1.0, freeL[cut]) + std::ldexp(1.0, F - freeL[cut])) * bond[cut] * sizeof(T);
        if (mem > 8e9) continue;
        if (best < 0.0 || cost < best) { best = cost; m = cut; }
    }
    QIL_REQUIRE(best >= 0.0, QIL_ERR_UNSUPPORTED, "coefficient grid: no cut fits the device memory budget");

    // ---- left panel: L (count x chi_m)
    Mat<T> L(ctx, 1, 1);
    fill_one_kernel<T><<<1, 32, 0, ctx->stream>>>(L.p);
    QIL_LAUNCH_CHECK(ctx);
    int64_t cntL = 1;
    for (int i = 0; i < m; ++i) {
        const int64_t cl = bond[i], cr = bond[i + 1];
        const T* core = reinterpret_cast<const T*>(psi->core[i]);
        if (mode[i] == 2) {
            Mat<T> Ln(ctx, cntL, 2 * cr);
            if (cntL == 1) gemv_cols<T>(ctx, core, 2 * cr, cl, 2 * cr, L.p, Ln.p);
            else gemm_tc<T>(ctx, cntL, 2 * cr, cl, 1.0, L.p, cl, core, 2 * cr, Ln.p, 2 * cr);
            L = std::move(Ln);
            cntL *= 2;
        } else {
            Mat<T> Ln(ctx, cntL, cr);
            if (cntL == 1) gemv_cols<T>(ctx, core + (size_t)mode[i] * cr, 2 * cr, cl, cr, L.p, Ln.p);
            else gemm_tc<T>(ctx, cntL, cr, cl, 1.0, L.p, cl, core + (size_t)mode[i] * cr, 2 * cr, Ln.p, cr);
            L = std::move(Ln);
        }
    }
    // ---- right panel: R (chi_m x count)
    Mat<T> R(ctx, 1, 1);
    fill_one_kernel<T><<<1, 32, 0, ctx->stream>>>(R.p);
    QIL_LAUNCH_CHECK(ctx);
    int64_t cntR = 1;
    for (int i = n - 1; i >= m; --i) {
        const int64_t cl = bond[i], cr = bond[i + 1];
        const T* core = reinterpret_cast<const T*>(psi->core[i]);
        if (mode[i] == 2) {
            Mat<T> Rn(ctx, 2 * cl, cntR);
            if (cntR == 1) gemv_rows<T>(ctx, core, cr, 2 * cl, cr, R.p, Rn.p);
            else gemm_tc<T>(ctx, 2 * cl, cntR, cr, 1.0, core, cr, R.p, cntR, Rn.p, cntR);
            R = std::move(Rn);
            cntR *= 2;
        } else {
            Mat<T> Rn(ctx, cl, cntR);
            if (cntR == 1) gemv_rows<T>(ctx, core + (size_t)mode[i] * cr, 2 * cr, cl, cr, R.p, Rn.p);
            else gemm_tc<T>(ctx, cl, cntR, cr, 1.0, core + (size_t)mode[i] * cr, 2 * cr, R.p, cntR, Rn.p, cntR);
            R = std::move(Rn);
        }
    }
    // ---- out[cL][cR] = amplitude * L . R ; the natural order is big-endian over the free sites
    bool identity = true;
    BitPerm bp;
    bp.nbits = F;
    if (out_bit) {
        unsigned long long seen = 0;
        for (int j = 0; j < F; ++j) {
            const int ob = out_bit[j];   // j-th free site in site order
            QIL_REQUIRE(ob >= 0 && ob < F && !((seen >> ob) & 1), QIL_ERR_ARGUMENT,
                        "coefficient grid: out_bit must be a permutation of 0..%d", F - 1);
            seen |= 1ull << ob;
            bp.dst[F - 1 - j] = ob;     // free site j sits at bit F-1-j of the natural index
            if (ob != F - 1 - j) identity = false;
        }
    }
    const int64_t total = (int64_t)1 << F;
    if (identity) {
        gemm_tc<T>(ctx, cntL, cntR, bond[m], psi->amplitude, L.p, bond[m], R.p, cntR, d_out, cntR);
    } else {
        Mat<T> tmp(ctx, cntL, cntR);
        gemm_tc<T>(ctx, cntL, cntR, bond[m], psi->amplitude, L.p, bond[m], R.p, cntR, tmp.p, cntR);
        const int grid = (int)std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 16);
        bit_permute_kernel<T><<<grid, 256, 0, ctx->stream>>>(bp, total, tmp.p, d_out);
        QIL_LAUNCH_CHECK(ctx);
    }
}

void coefficient_grid_dev(qil_ctx* ctx, const qil_mps* psi, const uint8_t* mode, const int32_t* out_bit, void* d_out) {
    { qil_prof_region prof_guard_(ctx, PROF_COEFF);
    if (psi->is_complex) coefficient_grid_impl<cplx>(ctx, psi, mode, out_bit, reinterpret_cast<cplx*>(d_out));
    else coefficient_grid_impl<double>(ctx, psi, mode, out_bit, reinterpret_cast<double*>(d_out));
    }
}

}  // namespace qil
