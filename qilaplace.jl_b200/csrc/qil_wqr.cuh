// qil_wqr.cuh -- warp-synchronous building blocks for the latency path of the encoder (round 2).
//
// The divide-and-conquer encoder (SignalConverters.jl:107-196 + rsvd.jl:38-121) is, apart from the
// 2+2q streaming passes of the top split, a chain of ~n small dependent factorizations: Householder QRs
// of l-column panels (rsvd.jl:83,90,94), SVDs of l x l triangles (rsvd.jl:103).  What bounds them is
// the number of dependent reductions, not flops, so the blocks below avoid CTA barriers altogether:
//
//   * wqr_factor      one WARP factors a block of <= 256 rows x <= 32 columns held in shared memory
//                     (row-major, odd pitch: lane <-> row is conflict free); ONE shuffle all-reduce per
//                     column step yields the column norm and all reflector dot products at once
//                     (x^H a_c for c >= j; u^H a_c = x^H a_c - conj(beta) a_jc follows algebraically),
//   * wqr_apply_chunk one warp applies the reflectors of a block to an 8-column (complex: 4) chunk
//                     held in REGISTERS (explicit Q / apply-down of a TSQR tree); the chunks of one
//                     block are independent, so ceil(n/8) warps work on a block at once,
//   * cta_qr          multi-level TSQR of an m x n panel inside one CTA from these two,
//   * wjacobi         one warp runs the one-sided Jacobi SVD of an ns x ns (ns <= 32) triangle; pairs of
//                     a round-robin round are spread over sub-warp lane groups, no block barrier,
//   * cta_gemm        panel GEMMs on DMMA m16n8k16 with fragments read straight from global / shared.
//
// Same reflector convention as qil_hh.cuh: H_j = I - tau_j u_j u_j^H, H_j x = beta_j e_1, u_j stored in
// column j rows j.. (head u_j[j] = x_j - beta_j), R strictly above the diagonal, diag(R) in beta.
#pragma once
#include "qil_common.cuh"

namespace qil {

constexpr int kWqrMaxN = 32;   // columns of a fast-path panel
constexpr int kWqrRpl = 8;     // default row slots per lane
constexpr int kWqrMaxRpl = 8;  // a block inside a CTA has at most 256 rows (measured: 16 warps on one 512-row block are
constexpr int kWqrMaxRows = 32 * kWqrMaxRpl;   // slower per column step than 8 warps on 256 rows -- the redundant scalar work and the
                                               // shuffles of all warps share the SM's FP64 / shuffle pipes)

template <typename T> struct WqrChunk { static constexpr int CH = 4; };      // register chunk of the apply-down (cta_qr)
template <> struct WqrChunk<cplx> { static constexpr int CH = 2; };

template <typename T> __device__ __forceinline__ T wq_sum(T v);
template <> __device__ __forceinline__ double wq_sum<double>(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <> __device__ __forceinline__ cplx wq_sum<cplx>(cplx v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}

template <typename T> __device__ __forceinline__ T wq_shfl_xor(T v, int o);
template <> __device__ __forceinline__ double wq_shfl_xor<double>(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
template <> __device__ __forceinline__ cplx wq_shfl_xor<cplx>(cplx v, int o) {
    return make_double2(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o));
}

// ---- Householder factorisation of one block by one warp -----------------------------------------------
// blk: m x n, row-major, pitch (elements, odd) in shared memory; m <= 256, n <= 32.  k = min(m, n) reflectors.
// Pitch of a panel of n columns: odd (lane <-> row is conflict free) and >= n + 3: column n is a ZERO padding column
// that the factorisation uses as the target of "no column" slots (read as zero, never written).
__host__ __device__ inline int wqr_pitch(int n) { return (n + 3) | 1; }

__device__ __forceinline__ void wq_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// W warps factor one block together, column parallel: at step j warp wi owns the trailing columns
// j+1+wi, j+1+wi+W, ... (at most NI per pass), computes x^H a_c for them (and |x|^2 redundantly), one shuffle
// all-reduce, the reflector scalars (identical in every warp), updates its columns; one named barrier per step.
// Per warp and step that is ~RPL*(1+NI) loads/FMAs and NI+1 reduced values instead of the whole trailing matrix:
// the step is a short latency chain (~600 cycles) rather than ~1000 instructions issued by one warp.
//   blk: m x n row-major, pitch = wqr_pitch(n), padding columns zero; m <= 32*RPL rows; all W*32 threads call with
//   the same arguments except wi; `bar` is a named-barrier id (1..15) private to this group of warps (W == 1: unused).
template <typename T> struct WqrNI { static constexpr int NI = 4; };      // trailing columns per warp and pass
template <> struct WqrNI<cplx> { static constexpr int NI = 2; };

// reflector scalars from |x|^2 (rows >= j) and the diagonal element: beta = -ph |x|, tau = 1 / (|x| (|x| + |x0|)).
// rsqrt-based (one MUFU + Newton each) -- this sits on the critical path of every column step.
template <typename T>
__device__ __forceinline__ void wqr_reflector(double s0, T x0, T& bj, double& tj, T& head) {
    bj = Scalar<T>::zero();
    tj = 0.0;
    if (s0 > 0.0) {
        const double a02 = Scalar<T>::abs2(x0);
        const double ra = (a02 > 0.0) ? rsqrt(a02) : 0.0;
        const double a0 = a02 * ra;
        const double nx = s0 * rsqrt(s0);
        const T ph = (a02 > 0.0) ? Scalar<T>::scale(x0, ra) : Scalar<T>::one();
        bj = Scalar<T>::scale(ph, -nx);
        // tau = 1 / (|x| (|x| + |x0|)): hardware reciprocal seed (2^-23) + three Newton steps instead of the IEEE
        // division sequence (this scalar chain is executed once per column step, on the critical path)
        const double d = s0 + nx * a0;
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
        r = fma(r, fma(-d, r, 1.0), r);
        r = fma(r, fma(-d, r, 1.0), r);
        r = fma(r, fma(-d, r, 1.0), r);
        tj = (d > 1e-290 && d < 1e290) ? r : 1.0 / d;
    }
    head = Scalar<T>::sub(x0, bj);       // u_j[j]
}

#ifdef QIL_WQR_PROFILE
#define WQR_CLK(slot) do { if (j == 0 && threadIdx.x == 0) { const long long c_ = clock64(); g_clk[slot] = c_ - clk_; clk_ = c_; } } while (0)
#else
#define WQR_CLK(slot) do { } while (0)
#endif

// One pass of a column step for NV (compile-time) columns c0, c0 + W, ... of this warp: everything inside is
// unconditional straight-line code (rows clamped, no per-slot branches), so the loads of the pass are all in flight
// together; the trailing values stay in registers between the dot products and the update.
template <typename T, int RPL, int NV, bool FIRST, int KEEPMAX>
__device__ __forceinline__ void wqr_pass(T* blk, int pitch, const int (&roff)[RPL], const T (&u)[RPL], int j, int m,
                                         int lane, int c0, int W, T x0, T& bj, double& tj, T& head, long long& clk_) {
    constexpr int NVV = NV > 0 ? NV : 1;
    constexpr bool KEEP = RPL * NV <= KEEPMAX;      // trailing values stay in registers between dots and update
    T ajc[NVV], s[NVV], a[KEEP ? RPL : 1][NVV];
    double s0 = 0.0;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        ajc[q] = blk[j * pitch + c0 + W * q];
        s[q] = Scalar<T>::zero();
    }
    if (KEEP) {
#pragma unroll
        for (int t = 0; t < RPL; ++t) {
#pragma unroll
            for (int q = 0; q < NV; ++q) a[KEEP ? t : 0][q] = blk[roff[t] + c0 + W * q];
        }
    }
#pragma unroll
    for (int t = 0; t < RPL; ++t) {
        const T uc = Scalar<T>::conj(u[t]);
        if (FIRST) s0 += Scalar<T>::abs2(u[t]);
#pragma unroll
        for (int q = 0; q < NV; ++q) s[q] = Scalar<T>::fma(uc, KEEP ? a[KEEP ? t : 0][q] : blk[roff[t] + c0 + W * q], s[q]);
    }
    WQR_CLK(101);
    if (FIRST) s0 = wq_sum<double>(s0);
#pragma unroll
    for (int q = 0; q < NV; ++q) s[q] = wq_sum<T>(s[q]);
    WQR_CLK(102);
    if (FIRST) wqr_reflector<T>(s0, x0, bj, tj, head);
    WQR_CLK(103);
    if (tj == 0.0 || NV == 0) return;
    // f = -tau * (x^H a_c - conj(beta) a_{jc})  (= -tau u^H a_c)
    const T cb = Scalar<T>::conj(bj);
#pragma unroll
    for (int q = 0; q < NV; ++q) s[q] = Scalar<T>::scale(Scalar<T>::sub(s[q], Scalar<T>::mul(cb, ajc[q])), -tj);
    WQR_CLK(104);
#pragma unroll
    for (int t = 0; t < RPL; ++t) {
        const int i = lane + 32 * t;
        const T uu = (i == j) ? head : u[t];
        T v[NVV];
#pragma unroll
        for (int q = 0; q < NV; ++q) v[q] = KEEP ? a[KEEP ? t : 0][q] : blk[roff[t] + c0 + W * q];
        if (i >= j && i < m) {
#pragma unroll
            for (int q = 0; q < NV; ++q) blk[roff[t] + c0 + W * q] = Scalar<T>::fma(s[q], uu, v[q]);
        }
    }
    __syncwarp();
    WQR_CLK(105);
}

template <typename T, int RPL = kWqrRpl, int KEEPMAX = 16>
__device__ __forceinline__ void wqr_factor(T* blk, int pitch, int m, int n, T* beta, double* tau, int wi = 0, int W = 1,
                                           int bar = 0) {
    constexpr int NI = WqrNI<T>::NI;
    const int lane = threadIdx.x & 31;
    const int k = min(m, n);
    const int lgW = 31 - __clz(W);
    int roff[RPL];
#pragma unroll
    for (int t = 0; t < RPL; ++t) roff[t] = min(lane + 32 * t, m - 1) * pitch;
    long long clk_ = 0;
#ifdef QIL_WQR_PROFILE
    clk_ = clock64();
    long long clk_step_ = clk_;
#endif
    for (int j = 0; j < k; ++j) {
        // columns of this warp at this step: j+1+wi, j+1+wi+W, ...  (round robin from the pivot: balanced as j grows)
        const int ncols_w = max(0, (n - (j + 1) - wi + W - 1) >> lgW);     // W is a power of two
        T bj = Scalar<T>::zero();
        T head = Scalar<T>::zero();
        double tj = 0.0;
        if (wi == 0 || ncols_w > 0) {
            T u[RPL];
#pragma unroll
            for (int t = 0; t < RPL; ++t) {
                const int i = lane + 32 * t;
                const T v = blk[roff[t] + j];
                u[t] = (i >= j && i < m) ? v : Scalar<T>::zero();
            }
            const T x0 = blk[j * pitch + j];
            WQR_CLK(100);
            bool first = true;
            for (int base = 0; first || base < ncols_w; base += NI) {
                const int nv = min(NI, ncols_w - base);
                const int c0 = j + 1 + wi + W * base;
#define WQR_PASS(NVC, FST) wqr_pass<T, RPL, (NVC <= NI ? NVC : NI), FST, KEEPMAX>(blk, pitch, roff, u, j, m, lane, c0, W, x0, bj, tj, head, clk_)
                if (first) {
                    if (nv <= 0) WQR_PASS(0, true);
                    else if (nv == 1) WQR_PASS(1, true);
                    else if (nv == 2) WQR_PASS(2, true);
                    else if (nv == 3) WQR_PASS(3, true);
                    else WQR_PASS(4, true);
                    first = false;
                } else {
                    if (nv == 1) WQR_PASS(1, false);
                    else if (nv == 2) WQR_PASS(2, false);
                    else if (nv == 3) WQR_PASS(3, false);
                    else WQR_PASS(4, false);
                }
#undef WQR_PASS
                if (tj == 0.0) break;
            }
        }
        if (W > 1) wq_bar(bar, W * 32);      // column j+1 is final before anybody reads it as the next x
        if (wi == 0 && lane == 0) {
            blk[j * pitch + j] = head;       // after the barrier: nobody reads the diagonal of column j any more
            beta[j] = bj;
            tau[j] = tj;
        }
        __syncwarp();
        WQR_CLK(106);
#ifdef QIL_WQR_PROFILE
        if (threadIdx.x == 0) { const long long c_ = clock64(); g_clk[j] = c_ - clk_step_; clk_step_ = c_; clk_ = c_; }
#endif
    }
    if (W > 1) wq_bar(bar, W * 32);
}

// ---- row-parallel, register-resident Householder (real panels) ---------------------------------------------
// wqr_factor above keeps the trailing matrix in shared memory and re-reads it every column step through a different
// code path per (column count, first pass) case: ~1900 cycles per step, much of it instruction fetch (the kernel is
// ~380 KB of SASS) and per-warp redundant scalar work.  Here thread r of the first NW warps holds row r of the block
// in REGISTERS for the whole factorisation; after every step the trailing columns move one register down
// (a[c-1] <- a[c] + f_c u: the update writes the shifted position), so the pivot column is always a[0] and ONE short
// loop body serves every step.  Per step:
//   products x_r a_rc (x = pivot column below the diagonal)            -> this warp's [32][NC+2] staging tile (STS.128)
//   lane c sums column c of the tile (32 rows), writes one partial     -> partial[warp][c]
//   ONE named barrier over the NW warps
//   lane c adds the NW partials: x^H a_c (c = 0: |x|^2), reads the pivot row element a_jc;  the reflector scalars
//   come from lane 0 by shuffle;  f_c = -tau (x^H a_c - beta a_jc) goes to a per-warp buffer, every lane reads all f
//   (LDS.128 broadcasts) and updates its row.
// Output convention is wqr_factor's: reflectors below the diagonal (head on it), R strictly above, diag(R) in beta.
//   blk: m x n row-major, pitch; m <= 32 * NW; n <= NC;  scr: rqr_scratch_elems(NW, NC) doubles, 16-byte aligned.
// Called by the first NW warps of the CTA (all 32 lanes each); `bar` is a named-barrier id private to them.
__host__ __device__ inline size_t rqr_scratch_elems(int nw, int nc) {
    return (size_t)nw * 32 * (nc + 2) + 3 * (size_t)nw * nc + 2 * (size_t)nc;
}
// RT rows per thread: thread tid of the NW participating warps holds rows tid, tid + 32 NW, ... (RT of them), so a block
// has up to 32 * NW * RT rows; the products of a thread's rows are summed before they are staged, i.e. the shared-memory
// traffic per step does not grow with RT, only the FMAs do.
#ifdef QIL_RQR_PROFILE
#define RQR_CLK(slot) do { const long long c_ = clock64(); rqr_seg[slot] += c_ - rqr_t; rqr_t = c_; } while (0)
#else
#define RQR_CLK(slot) do { } while (0)
#endif
template <int NC, int RT>
__device__ __forceinline__ void rqr_factor(double* blk, int pitch, int m, int n, double* beta, double* tau, double* scr,
                                           int NW, int bar) {
    constexpr int PP = NC + 2;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nthr = NW * 32;
    const int cl = min(lane, NC - 1);
    double* P = scr + (size_t)w * 32 * PP;                  // this warp's staging tile
    double* part = scr + (size_t)NW * 32 * PP;              // [2][NW][NC]
    double* piv = part + 2 * (size_t)NW * NC;               // [2][NC]
    double* fbuf = piv + 2 * NC + (size_t)w * NC;           // [NC] per warp
    double a[RT][NC];
    int row[RT];
#pragma unroll
    for (int t = 0; t < RT; ++t) {
        row[t] = tid + t * nthr;
#pragma unroll
        for (int c = 0; c < NC; ++c) a[t][c] = (row[t] < m && c < n) ? blk[row[t] * pitch + c] : 0.0;
    }
#ifdef QIL_RQR_PROFILE
    long long rqr_seg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long rqr_t = clock64();
#endif
    for (int j = 0; j < n; ++j) {
        const int par = j & 1;
        double x[RT];
#pragma unroll
        for (int t = 0; t < RT; ++t) x[t] = (row[t] >= j) ? a[t][0] : 0.0;   // rows above the pivot are finished
#pragma unroll
        for (int c = 0; c < NC; c += 2) {
            double p0 = x[0] * a[0][c], p1 = x[0] * a[0][c + 1];
#pragma unroll
            for (int t = 1; t < RT; ++t) { p0 = fma(x[t], a[t][c], p0); p1 = fma(x[t], a[t][c + 1], p1); }
            *reinterpret_cast<double2*>(P + lane * PP + c) = make_double2(p0, p1);
        }
#pragma unroll
        for (int t = 0; t < RT; ++t) {
            if (row[t] == j) {
#pragma unroll
                for (int c = 0; c < NC; c += 2)
                    *reinterpret_cast<double2*>(piv + par * NC + c) = make_double2(a[t][c], a[t][c + 1]);
            }
        }
        __syncwarp();
        RQR_CLK(0);
        {   // lanes >= NC repeat column NC - 1 (same values, same addresses): no divergent region in the step
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int r = 0; r < 32; r += 4) {
                s0 += P[(r + 0) * PP + cl];
                s1 += P[(r + 1) * PP + cl];
                s2 += P[(r + 2) * PP + cl];
                s3 += P[(r + 3) * PP + cl];
            }
            part[((size_t)par * NW + w) * NC + cl] = (s0 + s1) + (s2 + s3);
        }
        RQR_CLK(1);
        if (NW > 1) wq_bar(bar, nthr); else __syncwarp();
        RQR_CLK(2);
        double tot = 0.0;
        for (int ww = 0; ww < NW; ++ww) tot += part[((size_t)par * NW + ww) * NC + cl];
        const double pv = piv[par * NC + cl];
        const double s0 = __shfl_sync(0xffffffffu, tot, 0), x0 = __shfl_sync(0xffffffffu, pv, 0);
        RQR_CLK(3);
        double bj, tj, head;
        wqr_reflector<double>(s0, x0, bj, tj, head);
        RQR_CLK(4);
        fbuf[cl] = (tj != 0.0) ? -tj * (tot - bj * pv) : 0.0;
        __syncwarp();
        RQR_CLK(5);
        double u[RT];
#pragma unroll
        for (int t = 0; t < RT; ++t) {
            u[t] = (row[t] > j) ? a[t][0] : (row[t] == j ? head : 0.0);
            if (row[t] < m) blk[row[t] * pitch + j] = (row[t] == j) ? head : a[t][0];
        }
        {
            const double f1 = fbuf[1];                      // increasing c: position c is read before it is overwritten
#pragma unroll
            for (int t = 0; t < RT; ++t) a[t][0] = fma(f1, u[t], a[t][1]);
        }
#pragma unroll
        for (int c = 2; c < NC; c += 2) {
            const double2 f = *reinterpret_cast<const double2*>(fbuf + c);
#pragma unroll
            for (int t = 0; t < RT; ++t) {
                a[t][c - 1] = fma(f.x, u[t], a[t][c]);
                a[t][c] = fma(f.y, u[t], a[t][c + 1 < NC ? c + 1 : c]);
            }
        }
#pragma unroll
        for (int t = 0; t < RT; ++t) a[t][NC - 1] = 0.0;
        if (tid == 0) { beta[j] = bj; tau[j] = tj; }
        RQR_CLK(6);
    }
#ifdef QIL_RQR_PROFILE
    if (tid == 0) for (int i = 0; i < 8; ++i) g_rqr_seg[i] = rqr_seg[i];
#endif
    if (NW > 1) wq_bar(bar, nthr); else __syncwarp();
}
// run-time column capacity / rows per thread -> compile-time.  m <= 32 * NW * RT with NW <= 8 warps; RT = 3 only up to 24
// columns (register budget)
constexpr int kRqrMaxRows24 = 768, kRqrMaxRows32 = 512;
__host__ __device__ inline int rqr_max_rows(int n) { return n <= 24 ? kRqrMaxRows24 : kRqrMaxRows32; }
__host__ __device__ inline int rqr_nc_for(int n) { return (n + 7) & ~7; }
template <int NC>
__device__ __forceinline__ void rqr_factor_rt(double* blk, int pitch, int m, int n, double* beta, double* tau, double* scr,
                                              int NW, int bar) {
    const int rt = (m + NW * 32 - 1) / (NW * 32);
    if (rt <= 1) rqr_factor<NC, 1>(blk, pitch, m, n, beta, tau, scr, NW, bar);
    else if (rt == 2) rqr_factor<NC, 2>(blk, pitch, m, n, beta, tau, scr, NW, bar);
    else if (NC <= 24) rqr_factor<(NC <= 24 ? NC : 8), 3>(blk, pitch, m, n, beta, tau, scr, NW, bar);
}
// all threads of the CTA call; the first NW warps work.  No trailing CTA barrier.  Few warps with two or three rows per
// thread beat many warps with one (tools/ubench_rqr.cu, cycles per column step at n = 20: 128 rows 1840 with 4 warps,
// 1570 with 2; 256 rows 2364 with 8 warps, 1651 with 4; 384 rows 1823 with 4): fewer partial sums to combine, less
// barrier skew, and the FMAs of the extra rows are cheap next to the latency chain of a step.
__device__ __forceinline__ void rqr_factor_any(double* blk, int pitch, int m, int n, double* beta, double* tau, double* scr,
                                               int bar) {
    const int NW = (m <= 128) ? min(2, (m + 31) >> 5) : ((m <= (n <= 24 ? 384 : 256)) ? 4 : 8);
    if ((int)(threadIdx.x >> 5) >= NW) return;
    if (n <= 8) rqr_factor_rt<8>(blk, pitch, m, n, beta, tau, scr, NW, bar);
    else if (n <= 16) rqr_factor_rt<16>(blk, pitch, m, n, beta, tau, scr, NW, bar);
    else if (n <= 24) rqr_factor_rt<24>(blk, pitch, m, n, beta, tau, scr, NW, bar);
    else rqr_factor_rt<32>(blk, pitch, m, n, beta, tau, scr, NW, bar);
}

// run-time row-slot count -> compile-time RPL (1, 2, 4, 8): a block of m rows costs ceil(m/32) slots, not 8
template <typename T>
__device__ __forceinline__ void wqr_factor_any(T* blk, int pitch, int m, int n, T* beta, double* tau, int wi, int W,
                                               int bar) {
    const int rpl = (m + 31) >> 5;
    if (rpl <= 1) wqr_factor<T, 1>(blk, pitch, m, n, beta, tau, wi, W, bar);
    else if (rpl <= 2) wqr_factor<T, 2>(blk, pitch, m, n, beta, tau, wi, W, bar);
    else if (rpl <= 4) wqr_factor<T, 4>(blk, pitch, m, n, beta, tau, wi, W, bar);
    else wqr_factor<T, 8>(blk, pitch, m, n, beta, tau, wi, W, bar);
}

// ---- reflectors of one block applied to a register chunk -----------------------------------------------
// b[t][q] holds element (row lane + 32 t, chunk column q).  b <- H_0 H_1 ... H_{k-1} b.  The next reflector column is
// loaded while the current reduction is in flight.
template <typename T, int RPL, int CH>
__device__ __forceinline__ void wqr_apply_chunk(const T* V, int pitch, int m, int k, const double* tau,
                                                T (&b)[RPL][CH]) {
    const int lane = threadIdx.x & 31;
    int roff[RPL];
#pragma unroll
    for (int t = 0; t < RPL; ++t) roff[t] = min(lane + 32 * t, m - 1) * pitch;
    T un[RPL];
#pragma unroll
    for (int t = 0; t < RPL; ++t) un[t] = (k > 0) ? V[roff[t] + k - 1] : Scalar<T>::zero();
    for (int j = k - 1; j >= 0; --j) {
        const double tj = tau[j];
        T u[RPL];
#pragma unroll
        for (int t = 0; t < RPL; ++t) {
            const int i = lane + 32 * t;
            u[t] = (i >= j && i < m) ? un[t] : Scalar<T>::zero();
        }
        const int jn = max(j - 1, 0);
#pragma unroll
        for (int t = 0; t < RPL; ++t) un[t] = V[roff[t] + jn];       // next reflector, in flight during the reduction
        if (tj == 0.0) continue;
        T w[CH];
#pragma unroll
        for (int q = 0; q < CH; ++q) w[q] = Scalar<T>::zero();
#pragma unroll
        for (int t = 0; t < RPL; ++t) {
            const T uc = Scalar<T>::conj(u[t]);
#pragma unroll
            for (int q = 0; q < CH; ++q) w[q] = Scalar<T>::fma(uc, b[t][q], w[q]);
        }
#pragma unroll
        for (int q = 0; q < CH; ++q) w[q] = Scalar<T>::scale(wq_sum<T>(w[q]), -tj);
#pragma unroll
        for (int t = 0; t < RPL; ++t) {
#pragma unroll
            for (int q = 0; q < CH; ++q) b[t][q] = Scalar<T>::fma(w[q], u[t], b[t][q]);
        }
    }
}

// unit phase of a diagonal entry of R (ITensors qr(...; positive=true): Q <- Q diag(ph), R <- diag(conj ph) R)
template <typename T>
__device__ __forceinline__ T wqr_phase(T b) {
    const double a2 = Scalar<T>::abs2(b);
    return a2 > 0.0 ? Scalar<T>::scale(b, rsqrt(a2)) : Scalar<T>::one();
}

// ---- multi-level TSQR of an m x n panel inside one CTA ---------------------------------------------------
// Level 0 is the caller's panel (row-major, `pitch`), split into nb[0] balanced blocks of <= 256 rows, one warp
// each; the n x n triangles are stacked into the level-1 panel, and so on until one block remains.
struct CtaQrPlan {
    int nlev;
    int rows[4];
    int nb[4];
    int stack_rows;      // rows of all panels above level 0
    int blocks;          // blocks over all levels (beta / tau slots)
};

__host__ __device__ inline int cta_qr_blocks_for(int rows) { return (rows + kWqrMaxRows - 1) / kWqrMaxRows; }

__host__ __device__ inline CtaQrPlan cta_qr_plan(int m, int n) {
    CtaQrPlan p;
    p.nlev = 0;
    p.stack_rows = 0;
    p.blocks = 0;
    int rows = m;
    for (;;) {
        const int nb = cta_qr_blocks_for(rows);
        p.rows[p.nlev] = rows;
        p.nb[p.nlev] = nb;
        p.blocks += nb;
        ++p.nlev;
        if (nb == 1 || p.nlev == 4) break;
        rows = nb * n;          // every block of a multi-block level has >= n rows (rows / nb >= 128 >= n)
        p.stack_rows += rows;
    }
    return p;
}

// shared memory (in elements of T) a cta_qr call needs beyond the level-0 panel
template <typename T>
__host__ __device__ inline size_t cta_qr_extra_elems(int m, int n) {
    const CtaQrPlan p = cta_qr_plan(m, n);
    const int pitch = wqr_pitch(n);
    // stack panels + beta[blocks][n] + tau[blocks][n] (tau as doubles: at most one T each)
    return (size_t)p.stack_rows * pitch + 2 * (size_t)p.blocks * n + 8;
}

__device__ __forceinline__ void blk_range(int rows, int nb, int b, int& r0, int& r1) {
    r0 = (int)(((long long)b * rows) / nb);
    r1 = (int)(((long long)(b + 1) * rows) / nb);
}

// one level of the apply-down: every block's seed (n x n, rows of the level above; identity * phases at the top)
// is extended by zero rows and multiplied by the block's reflectors, CH columns per warp in registers.
template <typename T>
struct CtaQrLevel {
    T* P; int pp; int rows; int nb; int n;
    const T* seed; int spitch;
    const T* btop; bool positive;
    const double* tau;
    T* Qout; long long ldq; int qcols;
};
template <typename T, int RPL, int CH>
__device__ __forceinline__ void cta_qr_apply_level(const CtaQrLevel<T>& lv) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    const int n = lv.n;
    const int nch = (n + CH - 1) / CH;
    // rounds of whole blocks: all chunks of a block run in the same round, results are stored after the barrier.
    // REQUIRES nch <= nwarps (results of a block are stored in place after its round): callers launch 8 warps for real
    // panels (<= 8 chunks of 4 columns) and 16 for complex ones (<= 16 chunks of 2).
    const int blocks_per_round = max(1, nwarps / nch);
    for (int b0 = 0; b0 < lv.nb; b0 += blocks_per_round) {
        const int b = b0 + warp / nch;
        const int ch = warp % nch;
        const bool active = (warp < blocks_per_round * nch) && (b < lv.nb);
        T reg[RPL][CH];
        int r0 = 0, r1 = 0;
        if (active) {
            blk_range(lv.rows, lv.nb, b, r0, r1);
            const int mloc = r1 - r0;
            const int c0 = ch * CH;
#pragma unroll
            for (int t = 0; t < RPL; ++t) {
                const int i = lane + 32 * t;
#pragma unroll
                for (int q = 0; q < CH; ++q) {
                    const int c = c0 + q;
                    T v = Scalar<T>::zero();
                    if (i < n && i < mloc && c < n) {
                        if (lv.seed == nullptr) {
                            if (i == c) v = lv.positive ? wqr_phase<T>(lv.btop[c]) : Scalar<T>::one();
                        } else {
                            v = lv.seed[(size_t)(b * n + i) * lv.spitch + c];
                        }
                    }
                    reg[t][q] = v;
                }
            }
            wqr_apply_chunk<T, RPL, CH>(lv.P + (size_t)r0 * lv.pp, lv.pp, mloc, min(mloc, n), lv.tau + (size_t)b * n, reg);
        }
        __syncthreads();
        if (active) {
            const int mloc = r1 - r0;
            const int c0 = ch * CH;
#pragma unroll
            for (int t = 0; t < RPL; ++t) {
                const int i = lane + 32 * t;
                if (i < mloc) {
#pragma unroll
                    for (int q = 0; q < CH; ++q) {
                        const int c = c0 + q;
                        if (lv.Qout) {
                            if (c < n) lv.Qout[(long long)(r0 + i) * lv.ldq + c] = reg[t][q];
                            else if (c < lv.qcols) lv.Qout[(long long)(r0 + i) * lv.ldq + c] = Scalar<T>::zero();
                        } else if (c < n) {
                            lv.P[(size_t)(r0 + i) * lv.pp + c] = reg[t][q];
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

// Factor + explicit Q.  All threads of the CTA call it (blockDim.x a multiple of 32, >= 32 * ceil(n / CH)).
//   panel : m x n row-major (pitch), m >= n; overwritten by Q (m x n) when `Qout` == nullptr, else left holding the
//           reflectors and Q is written to Qout (row-major, ldq; GLOBAL or shared), columns n..qcols-1 zero filled
//   Rout  : n x n row-major (ldr) upper triangular, or nullptr
//   work  : cta_qr_extra_elems<T>(m, n) elements of shared memory
template <typename T>
__device__ __forceinline__ void cta_qr(T* panel, int pitch, int m, int n, bool positive, T* Rout, int ldr, T* Qout,
                                       long long ldq, int qcols, T* work) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    const CtaQrPlan pl = cta_qr_plan(m, n);
    const int spitch = wqr_pitch(n);
    T* lev_panel[4];
    int lev_pitch[4];
    lev_panel[0] = panel;
    lev_pitch[0] = pitch;
    {
        T* w = work;
        for (int L = 1; L < pl.nlev; ++L) {
            lev_panel[L] = w;
            lev_pitch[L] = spitch;
            w += (size_t)pl.rows[L] * spitch;
        }
    }
    T* beta = work + (size_t)pl.stack_rows * spitch;                 // [blocks][n]
    double* tau = reinterpret_cast<double*>(beta + (size_t)pl.blocks * n);   // [blocks][n]
    int lev_slot[4];
    {
        int s = 0;
        for (int L = 0; L < pl.nlev; ++L) { lev_slot[L] = s; s += pl.nb[L]; }
    }

    // ---- factor, level by level; each block's triangle goes to the next level's panel.  The warps of the CTA are
    // split into groups of W, one group per block (column-parallel wqr_factor, named barrier = group + 1).
    for (int L = 0; L < pl.nlev; ++L) {
        T* P = lev_panel[L];
        const int pp = lev_pitch[L];
        int W = max(1, min(nwarps / pl.nb[L], 8));
        W = 1 << (31 - __clz(W));                         // power of two (wqr_factor shifts by log2 W)
        const int bpr = max(1, min(nwarps / W, W > 1 ? 15 : nwarps));   // blocks per round
        const int group = warp / W, wi = warp % W;
        for (int b0 = 0; b0 < pl.nb[L]; b0 += bpr) {
            const int b = b0 + group;
            if (group < bpr && b < pl.nb[L]) {
                int r0, r1;
                blk_range(pl.rows[L], pl.nb[L], b, r0, r1);
                T* bb = beta + (size_t)(lev_slot[L] + b) * n;
                double* tt = tau + (size_t)(lev_slot[L] + b) * n;
                wqr_factor_any<T>(P + (size_t)r0 * pp, pp, r1 - r0, n, bb, tt, wi, W, group + 1);
                if (L + 1 < pl.nlev) {
                    T* S = lev_panel[L + 1] + (size_t)b * n * spitch;
                    for (int idx = wi * 32 + lane; idx < n * spitch; idx += W * 32) {   // padding columns zeroed too
                        const int j = idx / spitch, c = idx - j * spitch;
                        T v = Scalar<T>::zero();
                        if (c == j) v = bb[j];
                        else if (c > j && c < n) v = P[(size_t)(r0 + j) * pp + c];
                        S[j * spitch + c] = v;
                    }
                }
            }
        }
        __syncthreads();
    }
    // ---- R of the top level (k = min(rows_top, n) rows; rows_top >= n whenever m >= n)
    const int Lt = pl.nlev - 1;
    const T* btop = beta + (size_t)lev_slot[Lt] * n;
    if (Rout) {
        const T* P = lev_panel[Lt];
        const int pp = lev_pitch[Lt];
        const int kk = min(pl.rows[Lt], n);
        for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
            const int j = idx / n, c = idx - j * n;
            T v = Scalar<T>::zero();
            if (j < kk) {
                if (c == j) v = btop[j];
                else if (c > j) v = P[(size_t)j * pp + c];
                if (positive) v = Scalar<T>::mul(Scalar<T>::conj(wqr_phase<T>(btop[j])), v);
            }
            Rout[(size_t)j * ldr + c] = v;
        }
    }
    __syncthreads();
    // ---- apply down: level L's explicit Q rows are the n x n seeds of level L-1's blocks
    int nch_last = 1, ch_last = 1;
    for (int L = Lt; L >= 0; --L) {
        CtaQrLevel<T> lv;
        lv.P = lev_panel[L]; lv.pp = lev_pitch[L]; lv.rows = pl.rows[L]; lv.nb = pl.nb[L]; lv.n = n;
        lv.seed = (L == Lt) ? nullptr : lev_panel[L + 1]; lv.spitch = spitch;
        lv.btop = btop; lv.positive = positive;
        lv.tau = tau + (size_t)lev_slot[L] * n;
        lv.Qout = (L == 0) ? Qout : nullptr; lv.ldq = ldq; lv.qcols = qcols;
        const int rpl = ((pl.rows[L] + pl.nb[L] - 1) / pl.nb[L] + 31) >> 5;
        constexpr int C8 = Scalar<T>::is_complex ? 2 : 4;
        if (rpl <= 2) cta_qr_apply_level<T, 2, C8>(lv);
        else if (rpl <= 4) cta_qr_apply_level<T, 4, C8>(lv);
        else cta_qr_apply_level<T, 8, C8>(lv);
        ch_last = C8;
        nch_last = (n + ch_last - 1) / ch_last;
    }
    // zero fill of the padding columns beyond the last chunk (Qout only)
    if (Qout && qcols > nch_last * ch_last) {
        const int c0 = nch_last * ch_last;
        const int wdt = qcols - c0;
        for (int idx = threadIdx.x; idx < m * wdt; idx += blockDim.x) {
            const int i = idx / wdt, c = c0 + idx % wdt;
            Qout[(long long)i * ldq + c] = Scalar<T>::zero();
        }
    }
    __syncthreads();
}

// ---- one-sided Jacobi by one warp -------------------------------------------------------------------------
// G: ns x ns, COLUMN-major (column j at G + j * pg, pg odd), ns <= 32.  Rotates the columns of G until they are
// mutually orthogonal: G V = W.  Returns with W in G, sig[pos] = descending column norms, order[pos] = column index.
// Pairs whose columns are both below `nu` are not rotated against each other (qil_common.cuh).
template <typename T>
__device__ __forceinline__ void wjacobi(T* G, int pg, int ns, double nu, double* sig, int* order) {
    const int lane = threadIdx.x & 31;
    const int ne = ns + (ns & 1);
    const int npairs = ne / 2;
    int gl = 32;
    while (gl > 2 && (32 / gl) < npairs) gl >>= 1;       // lanes per pair: 32 / gl >= npairs (gl >= 2 since ns <= 32)
    const int grp = lane / gl, gln = lane % gl;
    const double tol = sqrt((double)ns) * 2.220446049250313e-16;
    for (int sweep = 0; sweep < 60 && ns > 1; ++sweep) {
        int rotated = 0;
        for (int r = 0; r < ne - 1; ++r) {
            int a, b;
            if (grp == 0) { a = ne - 1; b = r; }
            else { a = (r + grp) % (ne - 1); b = (r - grp + (ne - 1)) % (ne - 1); }
            const bool act = (grp < npairs) && a < ns && b < ns;
            const int cp = min(a, b), cq = max(a, b);
            T* gp = G + cp * pg;
            T* gq = G + cq * pg;
            double al = 0.0, be = 0.0;
            T ga = Scalar<T>::zero();
            if (act) {
                for (int i = gln; i < ns; i += gl) {
                    const T x = gp[i], y = gq[i];
                    al += Scalar<T>::abs2(x);
                    be += Scalar<T>::abs2(y);
                    ga = Scalar<T>::fma(Scalar<T>::conj(x), y, ga);
                }
            }
            for (int o = gl >> 1; o > 0; o >>= 1) {
                al += __shfl_xor_sync(0xffffffffu, al, o);
                be += __shfl_xor_sync(0xffffffffu, be, o);
                ga = Scalar<T>::add(ga, wq_shfl_xor<T>(ga, o));
            }
            const double g2 = Scalar<T>::abs2(ga);
            if (act && g2 > tol * tol * al * be && g2 > 0.0 && !(al < nu && be < nu)) {
                const double rg = rsqrt(g2);
                const double ag = g2 * rg;
                const T ph = Scalar<T>::scale(ga, rg);
                const double dd = be - al;
                const double hh = dd * dd + 4.0 * g2;
                const double sq = hh * rsqrt(hh);
                const double t = (dd >= 0.0 ? 2.0 : -2.0) * ag / (fabs(dd) + sq);
                const double c = rsqrt(1.0 + t * t);
                const double s = c * t;
                const T sp = Scalar<T>::scale(ph, s);
                const T spc = Scalar<T>::conj(sp);
                for (int i = gln; i < ns; i += gl) {
                    const T x = gp[i], y = gq[i];
                    gp[i] = Scalar<T>::sub(Scalar<T>::scale(x, c), Scalar<T>::mul(spc, y));
                    gq[i] = Scalar<T>::add(Scalar<T>::mul(sp, x), Scalar<T>::scale(y, c));
                }
                rotated = 1;
            }
            __syncwarp();
        }
        if (!__any_sync(0xffffffffu, rotated)) break;
    }
    // column norms, rank sort (descending, ties by index)
    double mysig = 0.0;
    if (lane < ns) {
        const T* g = G + lane * pg;
        double a = 0.0;
        for (int i = 0; i < ns; ++i) a += Scalar<T>::abs2(g[i]);
        mysig = sqrt(a);
        sig[lane] = mysig;          // unsorted, for the ranking below
    }
    __syncwarp();
    int pos = 0;
    if (lane < ns) {
        for (int i = 0; i < ns; ++i) {
            const double si = sig[i];
            pos += (si > mysig || (si == mysig && i < lane)) ? 1 : 0;
        }
    }
    __syncwarp();
    if (lane < ns) { sig[pos] = mysig; order[pos] = lane; }
    __syncwarp();
}

// shared-memory image of one finish: G (column-major, pitch pg), G0 copy, sig, order.  Called by all threads of a
// CTA; `G0` must already hold the (scaled) matrix whose columns are to be orthogonalised, column-major with pitch pg.
// On return (after a barrier): Gw = W, sig sorted descending, order, *s_rank.
template <typename T>
__device__ __forceinline__ void cta_jacobi_rank(T* Gw, const T* G0, int pg, int ns, double cutoff, long long maxdim,
                                                long long mindim, double* sig, int* order, int* s_rank, double* margin,
                                                double* s_nu) {
    const int tid = threadIdx.x;
    for (int idx = tid; idx < ns * pg; idx += blockDim.x) Gw[idx] = G0[idx];
    // ||G||_F^2 in a fixed order (skip threshold, qil_common.cuh)
    for (int j = tid; j < ns; j += blockDim.x) {
        double a = 0.0;
        for (int i = 0; i < ns; ++i) a += Scalar<T>::abs2(G0[j * pg + i]);
        sig[j] = a;
    }
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int j = 0; j < ns; ++j) tot += sig[j];
        *s_nu = jacobi_skip_threshold(tot, ns, cutoff, mindim);
    }
    __syncthreads();
    if (tid < 32) {
        wjacobi<T>(Gw, pg, ns, *s_nu, sig, order);
        if (tid == 0) *s_rank = truncate_rank_dev(sig, ns, cutoff, maxdim < 1 ? 1 : maxdim, mindim < 1 ? 1 : mindim, margin);
    }
    __syncthreads();
}

// ---- panel GEMM on the FP64 tensor pipe, fragments straight from memory -----------------------------------
// Cs[M x ncol] (shared, row-major, pc)  = scale * op(A)[M x K] * B[K x ncol] (shared, row-major, pb; rows >= K need
// not exist: the k range is guarded, columns >= ncol are never read).  A is GLOBAL (or shared) row-major with lda:
//   TRANS = false: op(A)[m][k] = A[m * lda + k]          TRANS = true: op(A)[m][k] = A[k * lda + m]   (real: A^T)
// (TRANS is a run-time flag: one instance of the code per kernel)
// Warps take 16-row tiles round robin; at most 4 column tiles of 8 (ncol <= 32).
__device__ __forceinline__ void wq_dmma(double* c, const double* a, const double* b) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
        "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
          "d"(b[2]), "d"(b[3]));
}

__device__ __forceinline__ void cta_gemm(const bool TRANS, const double* __restrict__ A, long long lda, int M, int K,
                                         const double* B, int pb, int ncol, double* Cs, int pc, double scale) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ntl = (ncol + 7) >> 3;
    const int mtiles = (M + 15) >> 4;
    for (int mt = warp; mt < mtiles; mt += nwarps) {
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0; }
        const int r0 = mt * 16 + g, r1 = r0 + 8;
        auto load_a = [&](int k0, double (&af)[8]) {
            // A fragment: rows r0 / r1, k = k0 + 4t + {0..3}  (same k permutation for the B fragment below)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = k0 + 4 * t + i;
                double v0 = 0.0, v1 = 0.0;
                if (k < K) {
                    if (!TRANS) {
                        if (r0 < M) v0 = A[(long long)r0 * lda + k];
                        if (r1 < M) v1 = A[(long long)r1 * lda + k];
                    } else {
                        if (r0 < M) v0 = A[(long long)k * lda + r0];
                        if (r1 < M) v1 = A[(long long)k * lda + r1];
                    }
                }
                af[2 * i] = v0;
                af[2 * i + 1] = v1;
            }
        };
        double anext[8];
        load_a(0, anext);
        for (int k0 = 0; k0 < K; k0 += 16) {
            double af[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = anext[i];
            if (k0 + 16 < K) load_a(k0 + 16, anext);     // in flight while the tensor pipe works on this step
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                if (nt < ntl) {
                    double bf[4];
                    const int col = nt * 8 + g;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int k = k0 + 4 * t + i;
                        bf[i] = (k < K && col < ncol) ? B[(size_t)k * pb + col] : 0.0;
                    }
                    wq_dmma(acc[nt], af, bf);
                }
            }
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            if (nt < ntl) {
                const int c0 = nt * 8 + 2 * t;
                if (r0 < M) {
                    if (c0 < ncol) Cs[(size_t)r0 * pc + c0] = scale * acc[nt][0];
                    if (c0 + 1 < ncol) Cs[(size_t)r0 * pc + c0 + 1] = scale * acc[nt][1];
                }
                if (r1 < M) {
                    if (c0 < ncol) Cs[(size_t)r1 * pc + c0] = scale * acc[nt][2];
                    if (c0 + 1 < ncol) Cs[(size_t)r1 * pc + c0 + 1] = scale * acc[nt][3];
                }
            }
        }
    }
}

}  // namespace qil
