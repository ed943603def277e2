// qil_encode.cu -- signal_mps(x; method=:rsvd): divide-and-conquer TT with randomized SVD.
//
// Reference: src/signals/SignalConverters.jl:107-196 (compress_tt!, mid = (first+last-1) / 2) calling
// src/linalg/rsvd.jl:38-121 at every split with the SAME seed.  In the MSB-first row-major layout used
// here every split matrix A = T[lb * 2^nl, 2^nr * rb] is a free contiguous view, so the top split
// streams the raw signal in place: no normalisation pass, no permute, no copy.
//
// Gaussian test matrix: the reference draws `random_itensor(eltype, cR, alpha)` after
// `Random.seed!(seed)`, i.e. the first C*l numbers of the seeded normal stream laid out column-major
// (cR fastest).  We keep exactly that structure -- Omega[c][j] = stream[c + C*j] with one stream per
// encode -- and let the host pass the stream (the Julia shim passes `randn` after `Random.seed!`), or
// generate it on the device with a counter-based generator.
#include "qil_fast.cuh"
#include "qil_rng.cuh"

#include <atomic>
#include <cstdlib>
#include <thread>

namespace qil {

// dense Omega (C x l, row-major) for the generic (small) path
template <typename T>
__global__ void omega_dense_kernel(const T* stream, unsigned long long seed, long long C, int l, T* out) {
    const long long total = C * l;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long c = idx / l;
        const int j = (int)(idx - c * l);
        out[idx] = stream_at<T>(stream, seed, c + C * j);
    }
}

// ---- operand preparation for the streaming kernels -------------------------------------------------
// K1 operand X (rows = reduction index of the REAL view, pitch lpp, zero padded):
//   real   : X[c][j]            = S(c, j)
//   complex: X[2c][2j] = Re S, X[2c][2j+1] = Im S, X[2c+1][2j] = -Im S, X[2c+1][2j+1] = Re S
// where S is Omega (src == nullptr) or a dense C x l matrix (src, ld = l).
template <typename T>
__global__ void prep_x_k1_kernel(const T* src, const T* stream, unsigned long long seed, long long C, int l,
                                 long long rows_pad, int lpp, double* X, int j0 = 0, int ldsrc = 0) {
    constexpr int F = Scalar<T>::is_complex ? 2 : 1;
    const long long total = rows_pad * lpp;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / lpp;
        const int col = (int)(idx - row * lpp);
        const long long c = row / F;
        const int j = col / F;
        double v = 0.0;
        if (c < C && j < l) {
            const T s = src ? src[c * (long long)(ldsrc ? ldsrc : l) + j0 + j] : stream_at<T>(stream, seed, c + C * (long long)(j0 + j));
            if (F == 1) {
                v = Scalar<T>::real(s);
            } else {
                const double re = Scalar<T>::real(s);
                const double im = reinterpret_cast<const double*>(&s)[F - 1];
                const int a = (int)(row & 1), b = col & 1;
                v = (a == b) ? re : (a == 0 ? im : -im);
            }
        }
        X[idx] = v;
    }
}

// K2 operand: X[r][.] = Q[r][.] viewed as real (R x F*l), zero padded to rows_pad x lpp
template <typename T>
__global__ void prep_x_k2_kernel(const T* Q, long long R, int l, long long rows_pad, int lpp, double* X, int j0 = 0,
                                 int ldq = 0) {
    constexpr int F = Scalar<T>::is_complex ? 2 : 1;
    const double* q = reinterpret_cast<const double*>(Q);
    const long long total = rows_pad * lpp;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / lpp;
        const int col = (int)(idx - r * lpp);
        X[idx] = (r < R && col < F * l) ? q[r * (long long)(F * (ldq ? ldq : l)) + F * j0 + col] : 0.0;
    }
}

// Y (R x l, dense T) = sum_ks part[ks][r][.]  (K1: complex output is already interleaved)
template <typename T>
__global__ void reduce_k1_kernel(const double* part, int ksplit, long long R, int ldo, int l, T* Y, int j0 = 0, int ldy = 0) {
    constexpr int F = Scalar<T>::is_complex ? 2 : 1;
    double* y = reinterpret_cast<double*>(Y);
    const long long total = R * l * F;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / (l * F);
        const int col = (int)(idx - r * (l * F));
        double s = 0.0;
        for (int ks = 0; ks < ksplit; ++ks) s += part[((long long)ks * R + r) * ldo + col];
        y[r * (long long)(F * (ldy ? ldy : l)) + F * j0 + col] = s;
    }
}

// Z (C x l, dense T) = scale * sum_ks part  with the complex recombination
//   Zr[c][j] = Z'[2c][2j] + Z'[2c+1][2j+1],  Zi[c][j] = Z'[2c][2j+1] - Z'[2c+1][2j]
template <typename T>
__global__ void reduce_k2_kernel(const double* part, int ksplit, long long C, int ldo, int l, const double* scale,
                                 T* Z, int j0 = 0, int ldz = 0) {
    constexpr int F = Scalar<T>::is_complex ? 2 : 1;
    const double sc = scale ? scale[0] : 1.0;
    const long long Mtot = C * F;
    const long long total = C * l;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long c = idx / l;
        const int j = (int)(idx - c * l);
        if (F == 1) {
            double s = 0.0;
            for (int ks = 0; ks < ksplit; ++ks) s += part[((long long)ks * Mtot + c) * ldo + j];
            reinterpret_cast<double*>(Z)[c * (long long)(ldz ? ldz : l) + j0 + j] = s * sc;
        } else {
            double re = 0.0, im = 0.0;
            for (int ks = 0; ks < ksplit; ++ks) {
                const double* p0 = part + ((long long)ks * Mtot + 2 * c) * ldo + 2 * j;
                const double* p1 = p0 + ldo;
                re += p0[0] + p1[1];
                im += p0[1] - p1[0];
            }
            const long long zo = c * (long long)(ldz ? ldz : l) + j0 + j;
            reinterpret_cast<double*>(Z)[2 * zo] = re * sc;
            reinterpret_cast<double*>(Z)[2 * zo + 1] = im * sc;
        }
    }
}

// nrm[0] = sqrt(sum partials), nrm[1] = 1 / nrm[0]
__global__ void finalize_norm_kernel(const double* partials, int n, double* nrm) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += partials[i];
        const double c = sqrt(s);
        nrm[0] = c;
        nrm[1] = 1.0 / c;
    }
}
__global__ void set_norm_kernel(double c, double* nrm) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { nrm[0] = c; nrm[1] = 1.0 / c; }
}
template <typename T>
__global__ void scale_by_dev_kernel(long long n, const double* scale, T* x) {
    const double s = scale[0];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[i] = Scalar<T>::scale(x[i], s);
}

static inline int grid_for(qil_ctx* ctx, long long total) {
    return (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 16));
}

// ---- one rsvd split ----------------------------------------------------------------------------------
template <typename T>
struct SplitCtx {
    qil_ctx* ctx;
    const RsvdOpts* o;
    const T* stream;        // device copy of the host-supplied normal stream, or nullptr
    int64_t stream_len;
    double* d_nrm;          // [0] = ||x||, [1] = 1/||x||  (device)
    bool nrm_ready;
    const qil_comm* comm = nullptr;   // row-sharded top split: collectives of the host (nullptr = single device)
};

// ---- collectives of a row-sharded encode (callbacks supplied by the host, ordered on ctx->stream) ----
static void comm_failed(const char* which) {
    const std::string note = callback_error_note();
    callback_error_note().clear();
    QIL_THROW(QIL_ERR_RUNTIME, "sharded encode: the %s callback failed%s%s", which, note.empty() ? "" : ": ", note.c_str());
}
static void comm_allreduce(const qil_comm* c, void* d_buf, int64_t count_f64) {
    callback_error_note().clear();
    if (c->allreduce_sum_f64(c->user, d_buf, count_f64) != 0) comm_failed("all-reduce");
}
static void comm_allgather(const qil_comm* c, const void* d_send, void* d_recv, int64_t count_f64) {
    callback_error_note().clear();
    if (c->allgather_f64(c->user, d_send, d_recv, count_f64) != 0) comm_failed("all-gather");
}
// nrm[0] = sum of the per-CTA partial sums of squares (before the cross-rank reduction)
__global__ void sum_partials_kernel(const double* partials, int n, double* nrm) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += partials[i];
        nrm[0] = s;
    }
}
__global__ void norm_from_sumsq_kernel(double* nrm) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double c = sqrt(nrm[0]);
        nrm[0] = c;
        nrm[1] = 1.0 / c;
    }
}

// A*X for X = Omega (Xsrc == nullptr) or a dense C x l matrix.  Returns Y (R x l).
template <typename T>
static Mat<T> mul_A(SplitCtx<T>& sc, const T* A, int64_t R, int64_t C, int l, const T* Xsrc, bool want_sumsq) {
    qil_ctx* ctx = sc.ctx;
    constexpr int F = Scalar<T>::is_complex ? 2 : 1;
    Mat<T> Y(ctx, R, l);
    const int lw_max = 112 / F;                         // sketch columns per streaming launch (14 column tiles of 8)
    if (stream_supported(R, C * F, C * F, std::min(l, lw_max) * F)) {
        // wider sketches (k = 100: l = 105 real, 210 real columns when complex) go through in column panels
        for (int j0 = 0; j0 < l; j0 += lw_max) {
            const int lw = std::min(lw_max, l - j0);
            const bool sumsq_now = want_sumsq && j0 == 0;
            const int nt = stream_nt_for(lw * F);
            const int lpp = stream_lpp(nt);
            const int64_t rows_pad = (C * F + 31) / 32 * 32;
            Mat<double> X(ctx, rows_pad, lpp);
            prep_x_k1_kernel<T><<<grid_for(ctx, rows_pad * lpp), 256, 0, ctx->stream>>>(
                Xsrc, sc.stream, (unsigned long long)sc.o->seed, C, lw, rows_pad, lpp, X.p, j0, l);
            QIL_LAUNCH_CHECK(ctx);
            int ks; long long kc;
            stream_plan(ctx, R, C * F, &ks, &kc, nt);
            Mat<double> part(ctx, (int64_t)ks * R, nt * 8);
            Mat<double> ssq;
            const int grid = stream_grid(ctx, R, ks, nt);
            if (sumsq_now) ssq = Mat<double>(ctx, grid, 1);
            stream_gemm(ctx, false, reinterpret_cast<const double*>(A), R, C * F, C * F, X.p, lpp, nt, part.p, ks, kc,
                        sumsq_now ? ssq.p : nullptr, lw * F);
            reduce_k1_kernel<T><<<grid_for(ctx, R * lw * F), 256, 0, ctx->stream>>>(part.p, ks, R, nt * 8, lw, Y.p, j0, l);
            QIL_LAUNCH_CHECK(ctx);
            if (sumsq_now && sc.comm) {
                // ||x||^2 = sum over ranks of the local sums (8-byte all-reduce)
                sum_partials_kernel<<<1, 32, 0, ctx->stream>>>(ssq.p, grid, sc.d_nrm);
                QIL_LAUNCH_CHECK(ctx);
                comm_allreduce(sc.comm, sc.d_nrm, 1);
                norm_from_sumsq_kernel<<<1, 32, 0, ctx->stream>>>(sc.d_nrm);
                QIL_LAUNCH_CHECK(ctx);
                sc.nrm_ready = true;
            } else if (sumsq_now) {
                finalize_norm_kernel<<<1, 32, 0, ctx->stream>>>(ssq.p, grid, sc.d_nrm);
                QIL_LAUNCH_CHECK(ctx);
                sc.nrm_ready = true;
            }
        }
    } else {
        Mat<T> Om;
        const T* Xp = Xsrc;
        if (!Xsrc) {
            Om = Mat<T>(ctx, C, l);
            omega_dense_kernel<T><<<grid_for(ctx, C * l), 256, 0, ctx->stream>>>(sc.stream, (unsigned long long)sc.o->seed,
                                                                              C, l, Om.p);
            QIL_LAUNCH_CHECK(ctx);
            Xp = Om.p;
        }
        gemm<T>(ctx, OP_N, OP_N, R, l, C, 1.0, A, C, Xp, l, 0.0, Y.p, l);
    }
    return Y;
}

// A^H * Q (C x l), optionally scaled by the device scalar `scale`
template <typename T>
static Mat<T> mul_AH(SplitCtx<T>& sc, const T* A, int64_t R, int64_t C, int l, const T* Q, const double* scale) {
    qil_ctx* ctx = sc.ctx;
    constexpr int F = Scalar<T>::is_complex ? 2 : 1;
    Mat<T> Z(ctx, C, l);
    const int lw_max = 112 / F;
    if (stream_supported(R, C * F, C * F, std::min(l, lw_max) * F)) {
        for (int j0 = 0; j0 < l; j0 += lw_max) {
            const int lw = std::min(lw_max, l - j0);
            const int nt = stream_nt_for(lw * F);
            const int lpp = stream_lpp(nt, true);
            const int64_t rows_pad = (R + 31) / 32 * 32;
            Mat<double> X(ctx, rows_pad, lpp);
            prep_x_k2_kernel<T><<<grid_for(ctx, rows_pad * lpp), 256, 0, ctx->stream>>>(Q, R, lw, rows_pad, lpp, X.p, j0, l);
            QIL_LAUNCH_CHECK(ctx);
            int ks; long long kc;
            stream_plan(ctx, C * F, R, &ks, &kc, nt);
            Mat<double> part(ctx, (int64_t)ks * C * F, nt * 8);
            stream_gemm(ctx, true, reinterpret_cast<const double*>(A), R, C * F, C * F, X.p, lpp, nt, part.p, ks, kc, nullptr, lw * F);
            reduce_k2_kernel<T><<<grid_for(ctx, C * lw), 256, 0, ctx->stream>>>(part.p, ks, C, nt * 8, lw, scale, Z.p, j0, l);
            QIL_LAUNCH_CHECK(ctx);
        }
    } else {
        gemm<T>(ctx, OP_C, OP_N, C, l, R, 1.0, A, C, Q, l, 0.0, Z.p, l);
        if (scale) {
            scale_by_dev_kernel<T><<<grid_for(ctx, C * l), 256, 0, ctx->stream>>>(C * l, scale, Z.p);
            QIL_LAUNCH_CHECK(ctx);
        }
    }
    return Z;
}

// ---- rank-adaptive sketch width ------------------------------------------------------------------------
// After the first QR of Y = A Omega the diagonal of R tells how many sketch directions carry anything: for a
// Gaussian Omega every column beyond the numerical rank rho of A lies in the span of the first rho up to rounding,
// so |R_jj| <= ~eps * max|R_ii| there.  Such directions give rows of B = Q^H A at the rounding level, i.e. singular
// values whose squares are <= ~1e-26 of the total and are removed by ANY positive cutoff; dropping them right away
// leaves the kept singular triplets unchanged to rounding while every later pass over A, QR and SVD works on
// l' = (last significant column) + 1 columns instead of k + p.  Only done when the truncation rule is guaranteed to
// discard them anyway (cutoff >= 1e-18, and never below mindim); full-rank inputs keep all k + p columns.
constexpr double kSketchNoise = 1e-13;
template <typename T>
__global__ void sketch_width_kernel(const T* __restrict__ R, int l, long long ld, double thr, int* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double mx = 0.0;
        for (int j = 0; j < l; ++j) mx = fmax(mx, sqrt(Scalar<T>::abs2(R[(long long)j * ld + j])));
        int keep = 1;
        for (int j = 0; j < l; ++j)
            if (sqrt(Scalar<T>::abs2(R[(long long)j * ld + j])) > thr * mx) keep = j + 1;
        out[0] = keep;
    }
}
// l' for the R factor of the first sketch (host read-back)
template <typename T>
static int shrink_sketch_width(qil_ctx* ctx, const RsvdOpts& o, int l, const Mat<T>& Rr) {
    // opt-in (RsvdOpts::adaptive, bit 0 of the `flags` argument of qil_encode_rsvd*): the reference keeps all k+p columns
    // q == 0: without a power iteration the basis of the narrowed sketch is visibly noisier than that of the full
    // one (C2 family at cutoff 1e-14: bonds of 6-7 instead of 4 three levels down the tree), so it keeps all columns
    if (!o.adaptive || o.q < 1 || !(o.cutoff >= 1e-18) || l <= 1) return l;
    int* d_keep = (int*)ctx->alloc(sizeof(int));
    sketch_width_kernel<T><<<1, 32, 0, ctx->stream>>>(Rr.p, l, Rr.cols, kSketchNoise, d_keep);
    QIL_LAUNCH_CHECK(ctx);
    int keep = l;
    QIL_CUDA(cudaMemcpyAsync(&keep, d_keep, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    ctx->free(d_keep);
    // a few of the noise-level columns stay as oversampling: the conditioning of the first QR (kappa(Y) ~ sigma_1 /
    // sigma_rho) otherwise shows up at the 1e-10 level when there is no power iteration to clean the basis
    static const int margin = [] { const char* e = getenv("QIL_RSVD_MARGIN"); return e ? atoi(e) : 3; }();
    keep = std::min(l, keep + margin);
    keep = (int)std::max<int64_t>(keep, std::min<int64_t>(o.mindim, l));
    return std::min(keep, l);
}
// Q (rows x l) is compacted to its first l' columns
template <typename T>
static int shrink_sketch(qil_ctx* ctx, const RsvdOpts& o, int l, int64_t rows, Mat<T>& Q, const Mat<T>& Rr) {
    const int keep = shrink_sketch_width<T>(ctx, o, l, Rr);
    if (keep >= l) return l;
    Mat<T> Qc(ctx, rows, keep);
    QIL_CUDA(cudaMemcpy2DAsync(Qc.p, (size_t)keep * sizeof(T), Q.p, (size_t)l * sizeof(T), (size_t)keep * sizeof(T),
                               (size_t)rows, cudaMemcpyDeviceToDevice, ctx->stream));
    Q = std::move(Qc);
    return keep;
}

// ---- fast tail of a randomized split (round 2): Bh = A^H Q given (possibly as split-K partials) ------------------
// Bh = Qb Rb by the warp-synchronous TSQR, Jacobi + truncation on the device (svd_finish), then U = Q Us and
// S Vh = (Us^H G) Qb^H in one launch.  One host read-back (the rank) instead of ~10 launches and 2 read-backs.
// preallocated worst-case outputs of a split whose rank stays on the device (no host read-back)
template <typename T>
struct DevOut {
    T* U;          // R x r, ld r
    T* SVh;        // r x C, ld C
    int* rank;     // device
};
template <typename T>
static int rsvd_tail(SplitCtx<T>& sc, int64_t R, int64_t C, int l, const T* Q, int64_t ldq, const T* Bh, int64_t ldbh,
                     int nsum, int64_t sum_stride, const double* d_scale, Mat<T>& U, Mat<T>* SVh, Mat<T>* Vh,
                     Mat<double>* S, const DevOut<T>* dev = nullptr) {
    qil_ctx* ctx = sc.ctx;
    const RsvdOpts& o = *sc.o;
    struct Region {
        qil_ctx* c;
        explicit Region(qil_ctx* cc) : c(cc) { c->prof_begin(PROF_SVD); }
        ~Region() { c->prof_end(); }
    } region(ctx);
    Mat<T> Qb(ctx, C, l), Rb(ctx, l, l), Us(ctx, l, l), T2(ctx, l, l);
    Mat<double> Sv(ctx, l, 1);
    qr_fast<T>(ctx, C, l, Bh, ldbh, nsum, sum_stride, false, Qb.p, l, l, Rb.p);
    if (dev) {
        svd_finish<T>(ctx, l, Rb.p, d_scale, o.cutoff, o.maxdim, o.mindim, Us.p, T2.p, Sv.p, dev->rank);
        rsvd_outputs<T>(ctx, R, C, l, Q, ldq, Qb.p, l, Us.p, T2.p, Sv.p, dev->rank, dev->U, dev->SVh, nullptr);
        return -1;
    }
    int* d_rank = (int*)ctx->alloc(sizeof(int));
    svd_finish<T>(ctx, l, Rb.p, d_scale, o.cutoff, o.maxdim, o.mindim, Us.p, T2.p, Sv.p, d_rank);
    // worst-case sized outputs, written compactly with leading dimension r
    Mat<T> Uo(ctx, R, l), SVo, Vo;
    if (SVh) SVo = Mat<T>(ctx, l, C);
    if (Vh) Vo = Mat<T>(ctx, l, C);
    rsvd_outputs<T>(ctx, R, C, l, Q, ldq, Qb.p, l, Us.p, T2.p, Sv.p, d_rank, Uo.p, SVh ? SVo.p : nullptr,
                    Vh ? Vo.p : nullptr);
    int r = 0;
    QIL_CUDA(cudaMemcpyAsync(&r, d_rank, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    ctx->free(d_rank);
    Uo.cols = r;
    U = std::move(Uo);
    if (SVh) { SVo.rows = r; *SVh = std::move(SVo); }
    if (Vh) { Vo.rows = r; *Vh = std::move(Vo); }
    if (S) { Sv.rows = r; *S = std::move(Sv); }
    return r;
}

// ---- fused real split: streaming pass -> TSQR straight from the split-K partials -> Q written in the operand
// layout of the next pass.  Between two passes over A there are the launches of one QR and nothing else.
static int rsvd_split_fused(SplitCtx<double>& sc, const double* A, int64_t R, int64_t C, bool top, int l0,
                            Mat<double>& U, Mat<double>* SVh, Mat<double>* Vh, Mat<double>* S,
                            const DevOut<double>* dev = nullptr) {
    qil_ctx* ctx = sc.ctx;
    const RsvdOpts& o = *sc.o;
    int nt = stream_nt_for(l0);
    int lppC = stream_lpp(nt, false), lppR = stream_lpp(nt, true);   // XC is the K1 operand, XR the K2 operand
    const int64_t rowsR = (R + 31) / 32 * 32, rowsC = (C + 31) / 32 * 32;
    Mat<double> XR(ctx, rowsR, lppR), XC(ctx, rowsC, lppC);
    if (rowsR > R) QIL_CUDA(cudaMemsetAsync(XR.p, 0, XR.elems() * sizeof(double), ctx->stream));
    if (rowsC > C) QIL_CUDA(cudaMemsetAsync(XC.p, 0, XC.elems() * sizeof(double), ctx->stream));
    int ks1, ks2; long long kc1, kc2;
    stream_plan(ctx, R, C, &ks1, &kc1, nt);
    stream_plan(ctx, C, R, &ks2, &kc2, nt);
    Mat<double> partR(ctx, (int64_t)ks1 * R, nt * 8), partC(ctx, (int64_t)ks2 * C, nt * 8);
    // ---- Y = A Omega
    prep_x_k1_kernel<double><<<grid_for(ctx, rowsC * lppC), 256, 0, ctx->stream>>>(
        nullptr, sc.stream, (unsigned long long)o.seed, C, l0, rowsC, lppC, XC.p);
    QIL_LAUNCH_CHECK(ctx);
    const bool want_sumsq = top && !sc.nrm_ready;
    Mat<double> ssq;
    const int grid1 = stream_grid(ctx, R, ks1, nt);
    if (want_sumsq) ssq = Mat<double>(ctx, grid1, 1);
    stream_gemm(ctx, false, A, R, C, C, XC.p, lppC, nt, partR.p, ks1, kc1, want_sumsq ? ssq.p : nullptr, l0);
    if (want_sumsq) {
        finalize_norm_kernel<<<1, 32, 0, ctx->stream>>>(ssq.p, grid1, sc.d_nrm);
        QIL_LAUNCH_CHECK(ctx);
        sc.nrm_ready = true;
    }
    // split-K partials -> dense panel (one full-grid launch: 22 MB read by 148 SMs; a leaf CTA of the TSQR summing
    // its own 7 partials needs ~10 dependent L2 round trips instead), then the TSQR writes Q in operand layout
    Mat<double> YR(ctx, R, l0), ZC(ctx, C, l0);
    auto sum_R = [&](int l_) {
        reduce_k1_kernel<double><<<grid_for(ctx, R * l_), 256, 0, ctx->stream>>>(partR.p, ks1, R, nt * 8, l_, YR.p);
        QIL_LAUNCH_CHECK(ctx);
    };
    auto sum_C = [&](int l_) {
        reduce_k2_kernel<double><<<grid_for(ctx, C * l_), 256, 0, ctx->stream>>>(partC.p, ks2, C, nt * 8, l_, nullptr, ZC.p);
        QIL_LAUNCH_CHECK(ctx);
    };
    int l = l0;
    {
        Mat<double> Rtop;
        const bool adapt = top && o.adaptive;
        if (adapt) Rtop = Mat<double>(ctx, l0, l0);
        { qil_prof_region prof_guard_(ctx, PROF_QR);
        sum_R(l0);
        qr_fast<double>(ctx, R, l0, YR.p, l0, 1, 0, true, XR.p, lppR, lppR, adapt ? Rtop.p : nullptr);
        }
        if (adapt) {
            l = shrink_sketch_width<double>(ctx, o, l0, Rtop);
            if (l < l0) {
                const int nt2 = stream_nt_for(l), lppC2 = stream_lpp(nt2, false), lppR2 = stream_lpp(nt2, true);
                Mat<double> XR2(ctx, rowsR, lppR2), XC2(ctx, rowsC, lppC2);
                QIL_CUDA(cudaMemsetAsync(XR2.p, 0, XR2.elems() * sizeof(double), ctx->stream));
                if (rowsC > C) QIL_CUDA(cudaMemsetAsync(XC2.p, 0, XC2.elems() * sizeof(double), ctx->stream));
                QIL_CUDA(cudaMemcpy2DAsync(XR2.p, (size_t)lppR2 * sizeof(double), XR.p, (size_t)lppR * sizeof(double),
                                           (size_t)l * sizeof(double), (size_t)R, cudaMemcpyDeviceToDevice, ctx->stream));
                XR = std::move(XR2);
                XC = std::move(XC2);
                nt = nt2;
                lppC = lppC2;
                lppR = lppR2;
            }
        }
    }
    for (int it = 0; it < o.q; ++it) {
        stream_gemm(ctx, true, A, R, C, C, XR.p, lppR, nt, partC.p, ks2, kc2, nullptr, l);
        { qil_prof_region prof_guard_(ctx, PROF_QR);
        sum_C(l);
        qr_fast<double>(ctx, C, l, ZC.p, l, 1, 0, true, XC.p, lppC, lppC, nullptr);
        }
        stream_gemm(ctx, false, A, R, C, C, XC.p, lppC, nt, partR.p, ks1, kc1, nullptr, l);
        { qil_prof_region prof_guard_(ctx, PROF_QR);
        sum_R(l);
        qr_fast<double>(ctx, R, l, YR.p, l, 1, 0, true, XR.p, lppR, lppR, nullptr);
        }
    }
    // ---- B^H = A^H Q, then the device-side tail
    stream_gemm(ctx, true, A, R, C, C, XR.p, lppR, nt, partC.p, ks2, kc2, nullptr, l);
    sum_C(l);
    return rsvd_tail<double>(sc, R, C, l, XR.p, lppR, ZC.p, l, 1, 0, top ? sc.d_nrm + 1 : nullptr, U, SVh, Vh, S, dev);
}
template <typename T>
static int rsvd_split_fused_dispatch(SplitCtx<T>&, const T*, int64_t, int64_t, bool, int, Mat<T>&, Mat<T>*, Mat<T>*,
                                     Mat<double>*) {
    QIL_THROW(QIL_ERR_RUNTIME, "fused split: real signals only");
}
template <>
int rsvd_split_fused_dispatch<double>(SplitCtx<double>& sc, const double* A, int64_t R, int64_t C, bool top, int l0,
                                      Mat<double>& U, Mat<double>* SVh, Mat<double>* Vh, Mat<double>* S) {
    return rsvd_split_fused(sc, A, R, C, top, l0, U, SVh, Vh, S);
}

// rsvd of A (R x C contiguous): U (R x r), SVh (r x C).  `top` => A is the raw signal (scale by 1/||x||,
// sum of squares fused into the first pass when the streaming kernel is used).
template <typename T>
static int rsvd_split(SplitCtx<T>& sc, const T* A, int64_t R, int64_t C, bool top, Mat<T>& U, Mat<T>* SVh,
                      Mat<T>* Vh = nullptr, Mat<double>* S = nullptr) {
    qil_ctx* ctx = sc.ctx;
    const RsvdOpts& o = *sc.o;
    const int l0 = (int)std::min<int64_t>((int64_t)o.k + o.p, std::min(R, C));
    QIL_REQUIRE(l0 >= 1, QIL_ERR_RUNTIME, "In `rsvd`, left or right index set is empty.");
    if (std::min(R, C) <= (int64_t)o.k + o.p) {
        // l == min(R, C): the sketch spans the whole row/column space, so the randomized SVD IS the truncated
        // SVD of A (to rounding, for any Omega).  Skip the sketch/QR/projection chain.
        if (top && !sc.nrm_ready) {
            const double c = device_norm2<T>(ctx, A, R * C);
            set_norm_kernel<<<1, 32, 0, ctx->stream>>>(c, sc.d_nrm);
            QIL_LAUNCH_CHECK(ctx);
            sc.nrm_ready = true;
        }
        const int r = svd_trunc<T>(ctx, R, C, A, C, o.cutoff, o.maxdim, o.mindim, &U, nullptr, Vh, SVh, S);
        if (top) {
            if (SVh) { scale_by_dev_kernel<T><<<grid_for(ctx, r * C), 256, 0, ctx->stream>>>(r * C, sc.d_nrm + 1, SVh->p); QIL_LAUNCH_CHECK(ctx); }
            if (S) { scale_by_dev_kernel<double><<<1, 256, 0, ctx->stream>>>(r, sc.d_nrm + 1, S->p); QIL_LAUNCH_CHECK(ctx); }
        }
        return r;
    }
    if (sc.stream) QIL_REQUIRE(C * (int64_t)l0 <= sc.stream_len, QIL_ERR_ARGUMENT,
                               "rsvd: the supplied normal stream is shorter than %lld", (long long)(C * l0));
    constexpr int F0 = Scalar<T>::is_complex ? 2 : 1;
    const bool fast_qr = !sc.comm && qr_fast_supported<T>(ctx, R, l0) && qr_fast_supported<T>(ctx, C, l0);
    if (fast_qr && !Scalar<T>::is_complex && stream_supported(R, C * F0, C * F0, l0 * F0))
        return rsvd_split_fused_dispatch<T>(sc, A, R, C, top, l0, U, SVh, Vh, S);
    Mat<T> Y = mul_A<T>(sc, A, R, C, l0, nullptr, top && !sc.nrm_ready);
    if (top && !sc.nrm_ready) {
        // generic path: the norm needs its own pass
        const double c = device_norm2<T>(ctx, A, R * C);
        set_norm_kernel<<<1, 32, 0, ctx->stream>>>(c, sc.d_nrm);
        QIL_LAUNCH_CHECK(ctx);
        sc.nrm_ready = true;
    }
    Mat<T> Q, Rr;
    qr_thin<T>(ctx, R, l0, Y.p, l0, true, Q, Rr);
    // only at the top split: the raw signal's rounding floor (~1e-16) is three decades below the threshold, whereas
    // the factors handed down the tree carry the parent's accuracy floor (eps * sigma_1 / sigma_r ~ 1e-13 .. 1e-12),
    // i.e. sit right at it, and a rounding-dependent width there was seen to move a bond by one on 4 GPUs
    const int l = top ? shrink_sketch<T>(ctx, o, l0, R, Q, Rr) : l0;
    for (int it = 0; it < o.q; ++it) {
        Mat<T> Z = mul_AH<T>(sc, A, R, C, l, Q.p, nullptr);
        Mat<T> Qz, Rz;
        qr_thin<T>(ctx, C, l, Z.p, l, true, Qz, Rz);
        Mat<T> Y2 = mul_A<T>(sc, A, R, C, l, Qz.p, false);
        qr_thin<T>(ctx, R, l, Y2.p, l, true, Q, Rr);
    }
    // B^H = A^H Q (C x l), scaled by 1/||x|| at the top
    if (fast_qr) {
        // B^H unscaled; the 1/||x|| of the top split is applied to the l x l triangle inside svd_finish
        Mat<T> Bh = mul_AH<T>(sc, A, R, C, l, Q.p, nullptr);
        return rsvd_tail<T>(sc, R, C, l, Q.p, l, Bh.p, l, 1, 0, top ? sc.d_nrm + 1 : nullptr, U, SVh, Vh, S);
    }
    Mat<T> Bh = mul_AH<T>(sc, A, R, C, l, Q.p, top ? sc.d_nrm + 1 : nullptr);
    Mat<T> Us;
    const int r = svd_trunc_adj<T>(ctx, l, C, Bh.p, l, o.cutoff, o.maxdim, o.mindim, &Us, nullptr, Vh, SVh, S);
    U = Mat<T>(ctx, R, r);
    gemm<T>(ctx, OP_N, OP_N, R, r, l, 1.0, Q.p, l, Us.p, r, 0.0, U.p, r);
    return r;
}

// ---- row-sharded top split -------------------------------------------------------------
// The signal matrix X (R x C, row-major) is split into G contiguous row blocks = leading-qubit blocks, one per
// rank.  Y = X Omega, U = Q U_s are row-local; the QR of the tall sketch is a TSQR whose per-rank l x l R factors
// are all-gathered; the l x C projections B = Q^H X (and Z = X^H Q of a power iteration) are sums of per-rank
// partials (all-reduce).  Everything below the top split is small and runs replicated on every rank.
template <typename T>
static void tsqr_sharded(SplitCtx<T>& sc, const Mat<T>& Y, int64_t Rg, int l, Mat<T>& Q, Mat<T>* Rout = nullptr) {
    qil_ctx* ctx = sc.ctx;
    constexpr int F = Scalar<T>::is_complex ? 2 : 1;
    const int G = sc.comm->world, g = sc.comm->rank;
    Mat<T> Qg, Rloc;
    {
        // The per-rank QR keeps the three-launch TSQR.  The cut-off decisions deep in the tree are knife-edge for the
        // bench family ("Rank parity and its limits"): any two correct QRs differ in the rounding-level sketch
        // directions beyond the numerical rank, and that is enough to move a sigma ~ 1e-6 by percents.  With this form the
        // sharded encode reproduces the bonds of the single-device encode at 2 / 4 / 8 ranks on every signal tested
        // (n = 20 .. 28); with the one-launch kernel (3x faster on a 4096-row block) the n = 28 bench signal came out with
        // bond 4 instead of 5 at site 4 on 4 ranks (margin 0.076 against 0.151).  Identity of bonds is the contract.
        struct Guard { qil_ctx* c; bool old; Guard(qil_ctx* cc) : c(cc), old(cc->tsqr_fused_ok) { c->tsqr_fused_ok = false; }
                       ~Guard() { c->tsqr_fused_ok = old; } } guard(ctx);
        qr_thin<T>(ctx, Rg, l, Y.p, l, true, Qg, Rloc);
    }
    Mat<T> Rall(ctx, (int64_t)G * l, l);
    comm_allgather(sc.comm, Rloc.p, Rall.p, (int64_t)l * l * F);
    Mat<T> Q2, R2;
    qr_thin<T>(ctx, (int64_t)G * l, l, Rall.p, l, true, Q2, R2);
    Q = Mat<T>(ctx, Rg, l);
    gemm<T>(ctx, OP_N, OP_N, Rg, l, l, 1.0, Qg.p, l, Q2.p + (size_t)g * l * l, l, 0.0, Q.p, l);
    if (Rout) *Rout = std::move(R2);
}

// A = this rank's Rg x C row block.  Returns the rank r; U is the FULL (G*Rg) x r factor, SVh is r x C (both replicated).
template <typename T>
static int rsvd_split_sharded(SplitCtx<T>& sc, const T* A, int64_t Rg, int64_t C, Mat<T>& U, Mat<T>& SVh) {
    qil_ctx* ctx = sc.ctx;
    const RsvdOpts& o = *sc.o;
    constexpr int F = Scalar<T>::is_complex ? 2 : 1;
    const int G = sc.comm->world;
    const int64_t R = Rg * G;
    const int l0 = (int)std::min<int64_t>((int64_t)o.k + o.p, std::min(R, C));
    QIL_REQUIRE(Rg >= l0 && stream_supported(Rg, C * F, C * F, l0 * F), QIL_ERR_UNSUPPORTED,
                "sharded encode: a %lld x %lld row block per rank is too small for k+p = %d; encode on one device",
                (long long)Rg, (long long)C, l0);
    if (sc.stream) QIL_REQUIRE(C * (int64_t)l0 <= sc.stream_len, QIL_ERR_ARGUMENT,
                               "rsvd: the supplied normal stream is shorter than %lld", (long long)(C * l0));
    Mat<T> Y = mul_A<T>(sc, A, Rg, C, l0, nullptr, true);
    Mat<T> Q, Rtop;
    tsqr_sharded<T>(sc, Y, Rg, l0, Q, &Rtop);
    const int l = shrink_sketch<T>(ctx, o, l0, Rg, Q, Rtop);      // Rtop is replicated: every rank picks the same l'
    for (int it = 0; it < o.q; ++it) {
        Mat<T> Z = mul_AH<T>(sc, A, Rg, C, l, Q.p, nullptr);
        comm_allreduce(sc.comm, Z.p, C * l * F);
        Mat<T> Qz, Rz;
        qr_thin<T>(ctx, C, l, Z.p, l, true, Qz, Rz);
        Mat<T> Y2 = mul_A<T>(sc, A, Rg, C, l, Qz.p, false);
        tsqr_sharded<T>(sc, Y2, Rg, l, Q);
    }
    int r;
    Mat<T> Ug;
    if (qr_fast_supported<T>(ctx, C, l)) {
        // device-side tail (TSQR of B^H, one-warp Jacobi, both products in one launch); replicated on every rank
        Mat<T> Bh = mul_AH<T>(sc, A, Rg, C, l, Q.p, nullptr);
        comm_allreduce(sc.comm, Bh.p, C * l * F);
        r = rsvd_tail<T>(sc, Rg, C, l, Q.p, l, Bh.p, l, 1, 0, sc.d_nrm + 1, Ug, &SVh, nullptr, nullptr);
    } else {
        Mat<T> Bh = mul_AH<T>(sc, A, Rg, C, l, Q.p, sc.d_nrm + 1);
        comm_allreduce(sc.comm, Bh.p, C * l * F);
        Mat<T> Us;
        r = svd_trunc_adj<T>(ctx, l, C, Bh.p, l, o.cutoff, o.maxdim, o.mindim, &Us, nullptr, nullptr, &SVh, nullptr);
        Ug = Mat<T>(ctx, Rg, r);
        gemm<T>(ctx, OP_N, OP_N, Rg, r, l, 1.0, Q.p, l, Us.p, r, 0.0, Ug.p, r);
    }
    U = Mat<T>(ctx, R, r);
    comm_allgather(sc.comm, Ug.p, U.p, Rg * r * F);
    return r;
}

// rsvd(A, Linds...; k, p, q, ...) on a dense device matrix (rsvd.jl:38-121) -> U (m x r), S (r), Vh (r x n)
template <typename T>
int rsvd_matrix(qil_ctx* ctx, const T* d_A, int64_t m, int64_t n, const RsvdOpts& o, Mat<T>& U, Mat<double>& S, Mat<T>& Vh) {
    QIL_REQUIRE(m >= 1 && n >= 1, QIL_ERR_RUNTIME, "In `rsvd`, left or right index set is empty.");
    Mat<double> nrm(ctx, 2, 1);
    SplitCtx<T> sc{ctx, &o, (const T*)o.omega, o.omega_rows * std::max<int64_t>(o.omega_cols, 1), nrm.p, true};
    return rsvd_split<T>(sc, d_A, m, n, false, U, nullptr, &Vh, &S);
}
template int rsvd_matrix<double>(qil_ctx*, const double*, int64_t, int64_t, const RsvdOpts&, Mat<double>&, Mat<double>&, Mat<double>&);
template int rsvd_matrix<cplx>(qil_ctx*, const cplx*, int64_t, int64_t, const RsvdOpts&, Mat<cplx>&, Mat<double>&, Mat<cplx>&);

// Level-synchronous divide and conquer (SignalConverters.jl:145-186).  The nodes of one level are independent:
// those whose split is an exact small SVD (min(R,C) <= k+p and the matrix fits one CTA) are decomposed by ONE
// batched launch; the rest (the streaming top split, mid-size randomized splits) go one by one.
template <typename T>
struct DcNode {
    const T* ptr;
    Mat<T> owner;
    int64_t lb, rb;
    int first, last;
    bool top;
};

template <typename T>
static void dc_encode(SplitCtx<T>& sc, std::vector<DcNode<T>> level, std::vector<void*>& cores,
                      std::vector<int64_t>& bond) {
    qil_ctx* ctx = sc.ctx;
    const RsvdOpts& o = *sc.o;
    while (!level.empty()) {
        std::vector<DcNode<T>> next;
        std::vector<SmallSvdItem<T>> batch;
        std::vector<size_t> batch_node;
        std::vector<Mat<T>> Us(level.size()), SVs(level.size());
        std::vector<int> ranks(level.size(), 0);
        std::vector<size_t> heavy;
        for (size_t i = 0; i < level.size(); ++i) {
            DcNode<T>& nd = level[i];
            if (nd.first == nd.last) {
                if (nd.owner.p == nd.ptr && nd.ptr != nullptr) {
                    cores[nd.first] = nd.owner.take();
                } else {
                    T* c = (T*)ctx->alloc((size_t)nd.lb * 2 * nd.rb * sizeof(T));
                    QIL_CUDA(cudaMemcpyAsync(c, nd.ptr, (size_t)nd.lb * 2 * nd.rb * sizeof(T), cudaMemcpyDeviceToDevice,
                                             ctx->stream));
                    cores[nd.first] = c;
                }
                continue;
            }
            const int mid = (nd.first + nd.last + 1) / 2 - 1;   // 0-based (first+last-1) / 2 (SignalConverters.jl:161)
            const int nl = mid - nd.first + 1, nr = nd.last - mid;
            const int64_t R = nd.lb << nl, C = ((int64_t)1 << nr) * nd.rb;
            if (!nd.top && std::min(R, C) <= (int64_t)o.k + o.p && svd_small_fits<T>(ctx, R, C)) {
                SmallSvdItem<T> it;
                it.A = nd.ptr; it.m = R; it.n = C; it.lda = C;
                it.want_U = true; it.want_SVh = true;
                batch.push_back(std::move(it));
                batch_node.push_back(i);
            } else {
                heavy.push_back(i);
            }
        }
        // randomized splits of one level are independent: run them on worker threads, one auxiliary stream each
        // (each split is a chain of small dependent launches with a host read-back of its rank)
        auto split_node = [&](SplitCtx<T>& s, size_t i) {
            DcNode<T>& nd = level[i];
            const int mid = (nd.first + nd.last + 1) / 2 - 1;
            const int nl = mid - nd.first + 1, nr = nd.last - mid;
            const int64_t R = nd.lb << nl, C = ((int64_t)1 << nr) * nd.rb;
            ranks[i] = rsvd_split<T>(s, nd.ptr, R, C, nd.top, Us[i], &SVs[i]);
        };
        const int nworkers = (int)std::min<size_t>(heavy.size(), 4);
        if (nworkers >= 2 && !ctx->is_aux && !sc.comm) {
            scoped_event ready;
            QIL_CUDA(cudaEventRecord(ready, ctx->stream));
            std::vector<int> code(nworkers, QIL_OK);
            std::vector<std::string> err(nworkers);
            std::vector<qil_ctx*> sub(nworkers);
            for (int w = 0; w < nworkers; ++w) {
                sub[w] = ctx->aux_ctx(w);
                QIL_CUDA(cudaStreamWaitEvent(sub[w]->stream, ready, 0));
            }
            auto body = [&](int w) {
                try {
                    QIL_CUDA(cudaSetDevice(ctx->device));
                    SplitCtx<T> s = sc;
                    s.ctx = sub[w];
                    for (size_t h = w; h < heavy.size(); h += nworkers) split_node(s, heavy[h]);
                } catch (const Error& e) {
                    code[w] = e.code; err[w] = e.msg;
                } catch (const std::exception& e) {
                    code[w] = QIL_ERR_RUNTIME; err[w] = e.what();
                }
            };
            std::vector<std::thread> th;
            for (int w = 1; w < nworkers; ++w) th.emplace_back(body, w);
            body(0);
            for (auto& t : th) t.join();
            for (int w = 0; w < nworkers; ++w) {
                // the parent stream continues after everything the worker enqueued
                QIL_CUDA(cudaEventRecord(ready, sub[w]->stream));
                QIL_CUDA(cudaStreamWaitEvent(ctx->stream, ready, 0));
                ctx->launches += sub[w]->launches;
                sub[w]->launches = 0;
            }
            for (size_t h : heavy) { Us[h].ctx = ctx; SVs[h].ctx = ctx; }     // results live on the parent context
            for (int w = 0; w < nworkers; ++w)
                if (code[w] != QIL_OK) throw Error(code[w], err[w]);
        } else {
            for (size_t h : heavy) split_node(sc, h);
        }
        svd_small_batch<T>(ctx, batch, o.cutoff, o.maxdim, o.mindim);
        for (size_t b = 0; b < batch.size(); ++b) {
            const size_t i = batch_node[b];
            ranks[i] = batch[b].rank;
            Us[i] = std::move(batch[b].U);
            SVs[i] = std::move(batch[b].SVh);
        }
        for (size_t i = 0; i < level.size(); ++i) {
            DcNode<T>& nd = level[i];
            if (nd.first == nd.last) continue;
            const int mid = (nd.first + nd.last + 1) / 2 - 1;
            const int r = ranks[i];
            bond[mid + 1] = r;
            nd.owner.release();
            DcNode<T> a, b;
            a.ptr = Us[i].p; a.owner = std::move(Us[i]); a.lb = nd.lb; a.rb = r; a.first = nd.first; a.last = mid; a.top = false;
            b.ptr = SVs[i].p; b.owner = std::move(SVs[i]); b.lb = r; b.rb = nd.rb; b.first = mid + 1; b.last = nd.last; b.top = false;
            next.push_back(std::move(a));
            next.push_back(std::move(b));
        }
        level = std::move(next);
    }
}

// ---- sync-free encoder (round 2): streaming top split with its rank left on the device, then ONE launch per tree
// level (qil_node.cu: one CTA per node, bond dimensions read from the device-side bond array).  The launch sequence
// depends on (n, k, p, q) only; the single host synchronisation is the final read-back of the bonds.  Real signals
// with k + p <= 32; everything else (and any node that overflows a CTA's shared memory) takes the general path.
__global__ void init_tree_state_kernel(int* state, int nbonds) {      // bonds = 1, overflow flag = 0
    for (int i = threadIdx.x; i <= nbonds; i += blockDim.x) state[i] = (i < nbonds) ? 1 : 0;
}
static int64_t bond_cap(int n, int pos, int kp) {
    const int a = std::min(pos, n - pos);
    return a >= 30 ? kp : std::min<int64_t>(kp, (int64_t)1 << a);
}
__global__ void set_int_kernel(int* p, int v) { if (threadIdx.x == 0 && blockIdx.x == 0) *p = v; }

// Ugiven / SVgiven / rgiven: the top split was already done by the caller (row-sharded encode): U (R x r) and S Vh
// (r x C), compact, stay owned by the caller; only the levels below run here.
static qil_mps* encode_rsvd_tree(SplitCtx<double>& sc, const double* x, int n, double* Ugiven = nullptr,
                                 double* SVgiven = nullptr, int rgiven = 0) {
    qil_ctx* ctx = sc.ctx;
    const RsvdOpts& o = *sc.o;
    const int kp = o.k + o.p;
    const bool given = Ugiven != nullptr;
    static const bool disabled = [] { const char* e = getenv("QIL_ENCODE_TREE"); return e && e[0] == '0'; }();
    if (disabled || kp > 32 || n < 4 || (sc.comm && !given)) return nullptr;
    const int mid = n / 2 - 1;                                  // 0-based (first+last-1)/2 of the root
    const int64_t R = (int64_t)1 << (mid + 1), C = (int64_t)1 << (n - 1 - mid);
    const int l0 = (int)std::min<int64_t>(kp, std::min(R, C));
    if (!given) {
        if (std::min(R, C) <= kp || !stream_supported(R, C, C, l0) || !qr_fast_supported<double>(ctx, R, l0) ||
            !qr_fast_supported<double>(ctx, C, l0))
            return nullptr;
        if (sc.stream && C * (int64_t)l0 > sc.stream_len) return nullptr;   // the general path reports the short stream
    }

    std::vector<int> h_bonds(n + 1, 1);
    int* d_state = (int*)ctx->alloc(sizeof(int) * (n + 2));     // bonds[0..n], overflow flag
    int* d_bonds = d_state;
    int* d_over = d_state + n + 1;
    init_tree_state_kernel<<<1, 128, 0, ctx->stream>>>(d_state, n + 1);
    QIL_LAUNCH_CHECK(ctx);
    struct Pending { int first, last; double* A; };
    std::vector<void*> cores(n, nullptr);
    std::vector<double*> temps;                                  // buffers that are not cores
    std::vector<std::vector<NodeDesc>> levels;
    auto fail_cleanup = [&]() {
        for (double* t : temps) ctx->free(t);
        for (void* c : cores) if (c) ctx->free(c);
        ctx->free(d_state);
    };
    try {
        // top split -> device buffers
        double* Utop = given ? Ugiven : (double*)ctx->alloc((size_t)R * l0 * sizeof(double));
        double* SVtop = given ? SVgiven : (double*)ctx->alloc((size_t)l0 * C * sizeof(double));
        std::vector<Pending> cur{{0, mid, Utop}, {mid + 1, n - 1, SVtop}}, next;
        while (!cur.empty()) {
            std::vector<NodeDesc> lvl;
            next.clear();
            for (const Pending& pd : cur) {
                if (pd.first == pd.last) { cores[pd.first] = pd.A; continue; }
                if (!(given && (pd.A == Ugiven || pd.A == SVgiven))) temps.push_back(pd.A);
                const int m2 = (pd.first + pd.last + 1) / 2 - 1;
                NodeDesc d;
                d.A = pd.A; d.bonds_off = 0;
                d.lb_pos = pd.first; d.rb_pos = pd.last + 1; d.out_pos = m2 + 1;
                d.nl = m2 - pd.first + 1; d.nr = pd.last - m2;
                const int64_t Rmax = bond_cap(n, pd.first, kp) << d.nl, Cmax = bond_cap(n, pd.last + 1, kp) << d.nr;
                const int64_t rmax = std::min<int64_t>(bond_cap(n, m2 + 1, kp), std::min(Rmax, Cmax));
                d.U = (double*)ctx->alloc((size_t)(Rmax * rmax) * sizeof(double));
                d.SVh = (double*)ctx->alloc((size_t)(rmax * Cmax) * sizeof(double));
                lvl.push_back(d);
                next.push_back({pd.first, m2, d.U});
                next.push_back({m2 + 1, pd.last, d.SVh});
            }
            if (!lvl.empty()) levels.push_back(std::move(lvl));
            cur = next;
        }
        size_t total = 0;
        for (auto& lv : levels) total += lv.size();
        std::vector<NodeDesc> flat;
        flat.reserve(total);
        for (auto& lv : levels) flat.insert(flat.end(), lv.begin(), lv.end());
        NodeDesc* d_nodes = nullptr;
        if (total) {
            d_nodes = (NodeDesc*)ctx->alloc(sizeof(NodeDesc) * total);
            temps.push_back(reinterpret_cast<double*>(d_nodes));
            QIL_CUDA(cudaMemcpyAsync(d_nodes, flat.data(), sizeof(NodeDesc) * total, cudaMemcpyHostToDevice, ctx->stream));
        }
        if (given) {
            set_int_kernel<<<1, 32, 0, ctx->stream>>>(d_bonds + mid + 1, rgiven);
            QIL_LAUNCH_CHECK(ctx);
        } else {
            Mat<double> Udummy;
            DevOut<double> dev{Utop, SVtop, d_bonds + mid + 1};
            rsvd_split_fused(sc, x, R, C, true, l0, Udummy, nullptr, nullptr, nullptr, &dev);
        }
        size_t off = 0;
        for (auto& lv : levels) {
            node_level_launch(ctx, d_nodes + off, (int)lv.size(), d_bonds, d_over, o, sc.stream, sc.stream_len);
            off += lv.size();
        }
        std::vector<int> h_state(n + 2);
        double h_nrm[2];
        QIL_CUDA(cudaMemcpyAsync(h_state.data(), d_state, sizeof(int) * (n + 2), cudaMemcpyDeviceToHost, ctx->stream));
        QIL_CUDA(cudaMemcpyAsync(h_nrm, sc.d_nrm, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->sync();                                             // flat / h_state stay alive until here
        if (h_state[n + 1] != 0) {                               // a node overflowed: general path
            fail_cleanup();
            return nullptr;
        }
        for (double* t : temps) ctx->free(t);
        ctx->free(d_state);
        std::vector<int64_t> bond(n + 1);
        for (int i = 0; i <= n; ++i) bond[i] = h_state[i];
        qil_mps* m = new_mps(ctx, n, 0, bond.data(), false);
        m->core = cores;
        m->amplitude = h_nrm[0];
        return m;
    } catch (...) {
        fail_cleanup();
        throw;
    }
}
// ---- the same for a batch of `count` equal-length signals, all levels in lock step (BASELINE configs[1]: 256 signals
// of n = 20).  The signals stored back to back ARE one tall matrix (count*R x C), so the sketch Y = X Omega of all of them
// is a single streaming launch (same Omega for every signal: same seed, same shape), and Z_s = X_s^H Q_s of all of them
// is a single split-K launch whose K chunks are the signals (no reduction needed); the panels go through the batched
// TSQR / Jacobi finish, then every tree level is one launch over (nodes of the level) x count CTAs.  All cores of the
// batch live in one pooled allocation.
__global__ void batch_sumsq_kernel(const double* __restrict__ x, long long N, int parts, double* __restrict__ partial) {
    // grid (parts, count): fixed-order block reduction of one slice of one signal
    __shared__ double red[256];
    const long long s = blockIdx.y;
    const long long per = (N + parts - 1) / parts;
    const long long b0 = blockIdx.x * per, b1 = min(b0 + per, N);
    double a = 0.0;
    for (long long i = b0 + threadIdx.x; i < b1; i += blockDim.x) { const double v = x[s * N + i]; a = fma(v, v, a); }
    red[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[s * parts + blockIdx.x] = red[0];
}
__global__ void batch_norm_finalize_kernel(const double* partial, int parts, long long count, double* nrm) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= count) return;
    double a = 0.0;
    for (int i = 0; i < parts; ++i) a += partial[s * parts + i];
    const double c = sqrt(a);
    nrm[2 * s] = c;
    nrm[2 * s + 1] = 1.0 / c;
}

static bool encode_rsvd_tree_batch(qil_ctx* ctx, const RsvdOpts& o, const double* x_all, int n, int64_t count,
                                   qil_mps** out) {
    const int kp = o.k + o.p;
    static const bool disabled = [] { const char* e = getenv("QIL_ENCODE_TREE"); return e && e[0] == '0'; }();
    if (disabled || kp > 32 || n < 14 || count < 2 || o.adaptive) return false;
    const int mid = n / 2 - 1;
    const int64_t R = (int64_t)1 << (mid + 1), C = (int64_t)1 << (n - 1 - mid), N = (int64_t)1 << n;
    const int64_t Rall = R * count;
    const int l0 = (int)std::min<int64_t>(kp, std::min(R, C));
    if (std::min(R, C) <= kp || R % 128 != 0 || C % 32 != 0 || !stream_supported(Rall, C, C, l0) ||
        !qr_fast_supported<double>(ctx, R, l0) || !qr_fast_supported<double>(ctx, C, l0) || Rall >= ((int64_t)1 << 31))
        return false;
    const double* nstream = (const double*)o.omega;
    const int64_t nstream_len = o.omega_rows * std::max<int64_t>(o.omega_cols, 1);
    if (nstream && C * (int64_t)l0 > nstream_len) return false;

    // ---- tree shape (the same for every signal) and buffer sizes per signal
    struct TNode { int first, last, m2, nl, nr; int64_t usz, ssz; int in_buf; int u_buf, s_buf; };
    struct Buf { int64_t size; bool is_core; int site; int64_t off; };
    std::vector<Buf> bufs;
    std::vector<std::vector<TNode>> levels;
    auto new_buf = [&](int64_t size) { bufs.push_back({size, false, -1, 0}); return (int)bufs.size() - 1; };
    const int b_utop = new_buf(R * l0), b_svtop = new_buf((int64_t)l0 * C);
    struct Pending { int first, last, buf; };
    std::vector<Pending> cur{{0, mid, b_utop}, {mid + 1, n - 1, b_svtop}}, next;
    while (!cur.empty()) {
        std::vector<TNode> lvl;
        next.clear();
        for (const Pending& pd : cur) {
            if (pd.first == pd.last) { bufs[pd.buf].is_core = true; bufs[pd.buf].site = pd.first; continue; }
            TNode t;
            t.first = pd.first; t.last = pd.last; t.in_buf = pd.buf;
            t.m2 = (pd.first + pd.last + 1) / 2 - 1;
            t.nl = t.m2 - pd.first + 1; t.nr = pd.last - t.m2;
            const int64_t Rmax = bond_cap(n, pd.first, kp) << t.nl, Cmax = bond_cap(n, pd.last + 1, kp) << t.nr;
            const int64_t rmax = std::min<int64_t>(bond_cap(n, t.m2 + 1, kp), std::min(Rmax, Cmax));
            t.usz = Rmax * rmax; t.ssz = rmax * Cmax;
            t.u_buf = new_buf(t.usz); t.s_buf = new_buf(t.ssz);
            lvl.push_back(t);
            next.push_back({pd.first, t.m2, t.u_buf});
            next.push_back({t.m2 + 1, pd.last, t.s_buf});
        }
        if (!lvl.empty()) levels.push_back(std::move(lvl));
        cur = next;
    }
    int64_t core_elems = 0, tmp_elems = 0;
    for (Buf& b : bufs) {
        b.size = (b.size + 1) & ~(int64_t)1;                       // 16-byte aligned regions
        int64_t& tot = b.is_core ? core_elems : tmp_elems;
        b.off = tot;
        tot += b.size * count;
    }
    const int nb1 = n + 1;
    double* core_pool = (double*)ctx->alloc((size_t)std::max<int64_t>(core_elems, 1) * sizeof(double));
    Mat<double> tmp_pool(ctx, std::max<int64_t>(tmp_elems, 1), 1);
    Mat<double> nrm(ctx, 2 * count, 1);
    int* d_state = (int*)ctx->alloc(sizeof(int) * (count * nb1 + 1));
    int* d_over = d_state + count * nb1;
    double* const core_base = core_pool;
    auto buf_ptr = [&](int b, int64_t s) { return (bufs[b].is_core ? core_base : tmp_pool.p) + bufs[b].off + s * bufs[b].size; };
    bool ok = false;
    try {
        init_tree_state_kernel<<<(int)std::min<int64_t>((count * nb1 + 256) / 256, 1024), 256, 0, ctx->stream>>>(
            d_state, (int)(count * nb1));
        QIL_LAUNCH_CHECK(ctx);
        // node table: level-major, node-major, signal-minor
        size_t total = 0;
        for (auto& lv : levels) total += lv.size() * (size_t)count;
        std::vector<NodeDesc> flat;
        flat.reserve(total);
        for (auto& lv : levels)
            for (const TNode& t : lv)
                for (int64_t s = 0; s < count; ++s) {
                    NodeDesc d;
                    d.A = buf_ptr(t.in_buf, s); d.U = buf_ptr(t.u_buf, s); d.SVh = buf_ptr(t.s_buf, s);
                    d.bonds_off = (int)(s * nb1);
                    d.lb_pos = t.first; d.rb_pos = t.last + 1; d.out_pos = t.m2 + 1; d.nl = t.nl; d.nr = t.nr;
                    flat.push_back(d);
                }
        Mat<double> node_tab;                                      // raw bytes of the table
        NodeDesc* d_nodes = nullptr;
        if (total) {
            node_tab = Mat<double>(ctx, (int64_t)((sizeof(NodeDesc) * total + 7) / 8), 1);
            d_nodes = reinterpret_cast<NodeDesc*>(node_tab.p);
            QIL_CUDA(cudaMemcpyAsync(d_nodes, flat.data(), sizeof(NodeDesc) * total, cudaMemcpyHostToDevice, ctx->stream));
        }
        // ---- norms
        {
            // enough CTAs to stream at HBM rate whatever the batch shape (4 x 2 GiB or 256 x 8 MiB)
            const int parts = (int)std::max<int64_t>(8, (4 * (int64_t)ctx->sm_count + count - 1) / count);
            Mat<double> partial(ctx, count * parts, 1);
            batch_sumsq_kernel<<<dim3(parts, (unsigned)count), 256, 0, ctx->stream>>>(x_all, N, parts, partial.p);
            QIL_LAUNCH_CHECK(ctx);
            batch_norm_finalize_kernel<<<(int)((count + 127) / 128), 128, 0, ctx->stream>>>(partial.p, parts, count, nrm.p);
            QIL_LAUNCH_CHECK(ctx);
        }
        // ---- top split of every signal
        const int nt = stream_nt_for(l0), lppC = stream_lpp(nt, false), lppR = stream_lpp(nt, true), ldo = nt * 8;
        Mat<double> XC(ctx, C, lppC), XCall, XRall(ctx, Rall, lppR);
        prep_x_k1_kernel<double><<<grid_for(ctx, C * lppC), 256, 0, ctx->stream>>>(nullptr, nstream, (unsigned long long)o.seed, C,
                                                                                 l0, C, lppC, XC.p);
        QIL_LAUNCH_CHECK(ctx);
        int ks1; long long kc1;
        stream_plan(ctx, Rall, C, &ks1, &kc1, nt);
        Mat<double> partR(ctx, (int64_t)ks1 * Rall, ldo), partC(ctx, count * C, ldo), Yall(ctx, Rall, l0);
        auto sum_R = [&]() {
            reduce_k1_kernel<double><<<grid_for(ctx, Rall * l0), 256, 0, ctx->stream>>>(partR.p, ks1, Rall, ldo, l0, Yall.p);
            QIL_LAUNCH_CHECK(ctx);
        };
        stream_gemm(ctx, false, x_all, Rall, C, C, XC.p, lppC, nt, partR.p, ks1, kc1, nullptr, l0);
        sum_R();
        qr_fast<double>(ctx, R, l0, Yall.p, l0, 1, 0, true, XRall.p, lppR, lppR, nullptr, (int)count, R * l0, R * lppR, 0);
        if (o.q > 0) XCall = Mat<double>(ctx, count * C, lppC);
        for (int it = 0; it < o.q; ++it) {
            // Z_s = X_s^H Q_s: split-K with one chunk per signal -> partial s IS Z_s
            stream_gemm(ctx, true, x_all, Rall, C, C, XRall.p, lppR, nt, partC.p, (int)count, R, nullptr, l0);
            qr_fast<double>(ctx, C, l0, partC.p, ldo, 1, 0, true, XCall.p, lppC, lppC, nullptr, (int)count, C * ldo, C * lppC, 0);
            stream_gemm(ctx, false, x_all, Rall, C, C, XCall.p, lppC, nt, partR.p, ks1, kc1, nullptr, l0, R, C * lppC);
            sum_R();
            qr_fast<double>(ctx, R, l0, Yall.p, l0, 1, 0, true, XRall.p, lppR, lppR, nullptr, (int)count, R * l0, R * lppR, 0);
        }
        stream_gemm(ctx, true, x_all, Rall, C, C, XRall.p, lppR, nt, partC.p, (int)count, R, nullptr, l0);   // B_s^H
        {
            Mat<double> Qb(ctx, count * C, l0), Rb(ctx, count * l0, l0), Us(ctx, count * l0, l0), T2(ctx, count * l0, l0);
            Mat<double> Sv(ctx, count * l0, 1);
            qr_fast<double>(ctx, C, l0, partC.p, ldo, 1, 0, false, Qb.p, l0, l0, Rb.p, (int)count, C * ldo, C * l0,
                            (int64_t)l0 * l0);
            svd_finish<double>(ctx, l0, Rb.p, nrm.p + 1, o.cutoff, o.maxdim, o.mindim, Us.p, T2.p, Sv.p, d_state + mid + 1,
                               (int)count, 2, nb1);
            rsvd_outputs<double>(ctx, R, C, l0, XRall.p, lppR, Qb.p, l0, Us.p, T2.p, Sv.p, d_state + mid + 1,
                                 buf_ptr(b_utop, 0), buf_ptr(b_svtop, 0), nullptr, (int)count, R * lppR, C * l0,
                                 bufs[b_utop].size, bufs[b_svtop].size, nb1);
        }
        // ---- the tree levels
        size_t off = 0;
        for (auto& lv : levels) {
            const size_t cnt = lv.size() * (size_t)count;
            node_level_launch(ctx, d_nodes + off, (int)cnt, d_state, d_over, o, nstream, nstream_len);
            off += cnt;
        }
        std::vector<int> h_state(count * nb1 + 1);
        std::vector<double> h_nrm(2 * count);
        QIL_CUDA(cudaMemcpyAsync(h_state.data(), d_state, sizeof(int) * h_state.size(), cudaMemcpyDeviceToHost, ctx->stream));
        QIL_CUDA(cudaMemcpyAsync(h_nrm.data(), nrm.p, sizeof(double) * 2 * count, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->sync();
        if (h_state[count * nb1] == 0) {
            std::shared_ptr<void> pool(core_pool, [ctx](void* p) { ctx->free(p); });
            core_pool = nullptr;
            for (int64_t s = 0; s < count; ++s) {
                std::vector<int64_t> bond(nb1);
                for (int i = 0; i < nb1; ++i) bond[i] = h_state[s * nb1 + i];
                qil_mps* m = new_mps(ctx, n, 0, bond.data(), false);
                for (size_t b = 0; b < bufs.size(); ++b)
                    if (bufs[b].is_core) m->core[bufs[b].site] = buf_ptr((int)b, s);
                m->pool = pool;
                m->amplitude = h_nrm[2 * s];
                out[s] = m;
            }
            ok = true;
        }
    } catch (...) {
        if (core_pool) ctx->free(core_pool);
        ctx->free(d_state);
        throw;
    }
    if (core_pool) ctx->free(core_pool);
    ctx->free(d_state);
    return ok;
}

template <typename T> static qil_mps* encode_rsvd_tree_dispatch(SplitCtx<T>&, const T*, int) { return nullptr; }
template <typename T>
static bool encode_rsvd_batch_tree_dispatch(qil_ctx*, const T*, int64_t, int64_t, const RsvdOpts&, qil_mps**) { return false; }
template <>
bool encode_rsvd_batch_tree_dispatch<double>(qil_ctx* ctx, const double* d_x, int64_t N, int64_t count, const RsvdOpts& o,
                                             qil_mps** out) {
    const int n = ilog2_round(N);
    if (n < 1 || N != ((int64_t)1 << n)) return false;              // padded signals: one by one
    // chunks of at most 2^31 stacked rows / a few GB of scratch
    const int64_t R = (int64_t)1 << (n / 2);
    const int64_t max_chunk = std::max<int64_t>(2, std::min<int64_t>(((int64_t)1 << 30) / R, ((int64_t)1 << 29) / std::max<int64_t>(N >> 6, 1)));
    for (int64_t s0 = 0; s0 < count; s0 += max_chunk) {
        const int64_t c = std::min(max_chunk, count - s0);
        if (c < 2 || !encode_rsvd_tree_batch(ctx, o, d_x + s0 * N, n, c, out + s0)) {
            for (int64_t i = 0; i < s0; ++i) { destroy(out[i]); out[i] = nullptr; }
            return false;
        }
    }
    return true;
}

template <> qil_mps* encode_rsvd_tree_dispatch<double>(SplitCtx<double>& sc, const double* x, int n) {
    return encode_rsvd_tree(sc, x, n);
}
template <typename T> static qil_mps* encode_tree_lower_dispatch(SplitCtx<T>&, int, T*, T*, int) { return nullptr; }
template <> qil_mps* encode_tree_lower_dispatch<double>(SplitCtx<double>& sc, int n, double* U, double* SVh, int r) {
    return encode_rsvd_tree(sc, nullptr, n, U, SVh, r);
}

template <typename T>
qil_mps* encode_rsvd(qil_ctx* ctx, const T* d_x, int64_t N, const RsvdOpts& o) {
    QIL_REQUIRE(N >= 1, QIL_ERR_ARGUMENT, "signal_mps: empty signal");
    QIL_REQUIRE(o.k >= 1 && o.p >= 0 && o.q >= 0, QIL_ERR_ARGUMENT, "rsvd: k >= 1, p >= 0, q >= 0 required");
    const int n = ilog2_round(N);
    QIL_REQUIRE(n >= 1, QIL_ERR_ARGUMENT, "_tensor_to_mps_rsvd: Need at least one site in the tensor to convert to MPS.");
    const int64_t Np = (int64_t)1 << n;
    QIL_REQUIRE(N <= Np, QIL_ERR_ASSERT, "_array_to_tensor: Length of signal vector must be a power of 2");
    QIL_REQUIRE(n <= kMaxSites, QIL_ERR_UNSUPPORTED, "signal too long");
    Mat<T> padded;
    const T* x = d_x;
    if (N < Np) {
        padded = Mat<T>(ctx, 1, Np);
        QIL_CUDA(cudaMemsetAsync(padded.p, 0, Np * sizeof(T), ctx->stream));
        QIL_CUDA(cudaMemcpyAsync(padded.p, d_x, N * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
        x = padded.p;
    }
    Mat<double> nrm(ctx, 2, 1);
    SplitCtx<T> sc{ctx, &o, (const T*)o.omega, o.omega_rows * std::max<int64_t>(o.omega_cols, 1), nrm.p, false};
    std::vector<int64_t> bond(n + 1, 1);
    std::vector<void*> cores(n, nullptr);
    if (n == 1) {
        const double c = device_norm2<T>(ctx, x, Np);
        T* core = (T*)ctx->alloc(2 * sizeof(T));
        scale_copy<T>(ctx, 2, 1.0 / c, x, core);
        cores[0] = core;
        qil_mps* m = new_mps(ctx, 1, Scalar<T>::is_complex ? 1 : 0, bond.data(), false);
        m->core = cores;
        m->amplitude = c;
        return m;
    }
    if (qil_mps* fastm = encode_rsvd_tree_dispatch<T>(sc, x, n)) return fastm;
    sc.nrm_ready = false;                                    // (a tree attempt that overflowed is repeated from scratch)
    {
        std::vector<DcNode<T>> level(1);
        DcNode<T>& root = level[0];
        root.ptr = x; root.lb = 1; root.rb = 1; root.first = 0; root.last = n - 1; root.top = true;
        dc_encode<T>(sc, std::move(level), cores, bond);
    }
    double h_nrm[2];
    QIL_CUDA(cudaMemcpyAsync(h_nrm, nrm.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    qil_mps* m = new_mps(ctx, n, Scalar<T>::is_complex ? 1 : 0, bond.data(), false);
    m->core = cores;
    m->amplitude = h_nrm[0];
    return m;
}
// signal_mps(x; method=:rsvd) with x split over comm->world ranks in contiguous chunks (rank order); every rank
// returns the same MPS.  N is the TOTAL length and must be a power of two divisible by the world size.
template <typename T>
qil_mps* encode_rsvd_sharded(qil_ctx* ctx, const qil_comm* comm, const T* d_x_local, int64_t N, const RsvdOpts& o) {
    QIL_REQUIRE(comm->world >= 1 && comm->rank >= 0 && comm->rank < comm->world, QIL_ERR_ARGUMENT,
                "sharded encode: invalid rank %d of %d", comm->rank, comm->world);
    QIL_REQUIRE(o.k >= 1 && o.p >= 0 && o.q >= 0, QIL_ERR_ARGUMENT, "rsvd: k >= 1, p >= 0, q >= 0 required");
    const int n = ilog2_round(N);
    QIL_REQUIRE(n >= 2 && N == ((int64_t)1 << n), QIL_ERR_ARGUMENT,
                "sharded encode: the total length must be a power of two (pad on the host)");
    QIL_REQUIRE(n <= kMaxSites, QIL_ERR_UNSUPPORTED, "signal too long");
    const int G = comm->world;
    QIL_REQUIRE((G & (G - 1)) == 0, QIL_ERR_ARGUMENT, "sharded encode: the world size must be a power of two");
    const int mid = n / 2 - 1;                               // top split, 0-based (SignalConverters.jl:161)
    const int64_t R = (int64_t)1 << (mid + 1), C = (int64_t)1 << (n - 1 - mid);
    QIL_REQUIRE(R % G == 0, QIL_ERR_UNSUPPORTED, "sharded encode: %d ranks do not divide the %lld rows", G, (long long)R);
    Mat<double> nrm(ctx, 2, 1);
    SplitCtx<T> sc{ctx, &o, (const T*)o.omega, o.omega_rows * std::max<int64_t>(o.omega_cols, 1), nrm.p, false, comm};
    std::vector<int64_t> bond(n + 1, 1);
    std::vector<void*> cores(n, nullptr);
    Mat<T> U, SVh;
    const int r = rsvd_split_sharded<T>(sc, d_x_local, R / G, C, U, SVh);
    bond[mid + 1] = r;
    sc.comm = nullptr;                                       // the lower levels are replicated, no collectives
    if (qil_mps* fastm = encode_tree_lower_dispatch<T>(sc, n, U.p, SVh.p, r)) return fastm;   // one launch per level
    {
        std::vector<DcNode<T>> level(2);
        DcNode<T>& a = level[0];
        DcNode<T>& b = level[1];
        a.ptr = U.p; a.owner = std::move(U); a.lb = 1; a.rb = r; a.first = 0; a.last = mid; a.top = false;
        b.ptr = SVh.p; b.owner = std::move(SVh); b.lb = r; b.rb = 1; b.first = mid + 1; b.last = n - 1; b.top = false;
        dc_encode<T>(sc, std::move(level), cores, bond);
    }
    double h_nrm[2];
    QIL_CUDA(cudaMemcpyAsync(h_nrm, nrm.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    qil_mps* m = new_mps(ctx, n, Scalar<T>::is_complex ? 1 : 0, bond.data(), false);
    m->core = cores;
    m->amplitude = h_nrm[0];
    return m;
}
template qil_mps* encode_rsvd_sharded<double>(qil_ctx*, const qil_comm*, const double*, int64_t, const RsvdOpts&);
template qil_mps* encode_rsvd_sharded<cplx>(qil_ctx*, const qil_comm*, const cplx*, int64_t, const RsvdOpts&);

// ---- batch of independent signals (BASELINE configs[1]: 256 signals of n = 20) ---------------------------------
// One encode is a chain of ~150 small dependent launches with a handful of host read-backs (the data-dependent
// ranks), i.e. latency bound; independent signals overlap perfectly.  `workers` host threads, each with its own
// stream forked from the context's stream, pull signals from a shared counter and run the ordinary encoder.
template <typename T>
void encode_rsvd_batch(qil_ctx* ctx, const T* d_x, int64_t N, int64_t count, const RsvdOpts& o, int workers,
                       qil_mps** out) {
    QIL_REQUIRE(count >= 0, QIL_ERR_ARGUMENT, "signal batch: negative count");
    if (count == 0) return;
    workers = (int)std::max<int64_t>(1, std::min<int64_t>(workers <= 0 ? 16 : workers, count));
    for (int64_t i = 0; i < count; ++i) out[i] = nullptr;
    if (encode_rsvd_batch_tree_dispatch<T>(ctx, d_x, N, count, o, out)) return;      // level-synchronous path
    scoped_event ready;
    QIL_CUDA(cudaEventRecord(ready, ctx->stream));
    std::vector<qil_ctx> sub(workers);
    std::vector<std::string> err(workers);
    std::vector<int> code(workers, QIL_OK);
    std::atomic<int64_t> next(0);
    for (int w = 0; w < workers; ++w) {
        sub[w].device = ctx->device;
        sub[w].sm_count = ctx->sm_count;
        sub[w].smem_optin = ctx->smem_optin;
        sub[w].own_stream = true;
        sub[w].d_margin = ctx->d_margin;
        sub[w].is_aux = true;                 // no nested worker threads below a batch worker
        QIL_CUDA(cudaStreamCreateWithFlags(&sub[w].stream, cudaStreamNonBlocking));
        QIL_CUDA(cudaStreamWaitEvent(sub[w].stream, ready, 0));
    }
    auto body = [&](int w) {
        try {
            QIL_CUDA(cudaSetDevice(ctx->device));
            for (;;) {
                const int64_t i = next.fetch_add(1);
                if (i >= count) break;
                out[i] = encode_rsvd<T>(&sub[w], d_x + (size_t)i * N, N, o);   // ends synchronised on its stream
                out[i]->ctx = ctx;                                             // the handle lives on the parent context
            }
        } catch (const Error& e) {
            code[w] = e.code; err[w] = e.msg;
        } catch (const std::exception& e) {
            code[w] = QIL_ERR_RUNTIME; err[w] = e.what();
        }
    };
    std::vector<std::thread> th;
    for (int w = 1; w < workers; ++w) th.emplace_back(body, w);
    body(0);
    for (auto& t : th) t.join();
    for (int w = 0; w < workers; ++w) {
        cudaStreamSynchronize(sub[w].stream);
        if (sub[w].scratch) cudaFreeAsync(sub[w].scratch, sub[w].stream);
        cudaStreamSynchronize(sub[w].stream);
        cudaStreamDestroy(sub[w].stream);
        ctx->launches += sub[w].launches;
    }
    for (int w = 0; w < workers; ++w)
        if (code[w] != QIL_OK) {
            for (int64_t i = 0; i < count; ++i) { if (out[i]) { destroy(out[i]); out[i] = nullptr; } }
            throw Error(code[w], err[w]);
        }
}
template void encode_rsvd_batch<double>(qil_ctx*, const double*, int64_t, int64_t, const RsvdOpts&, int, qil_mps**);
template void encode_rsvd_batch<cplx>(qil_ctx*, const cplx*, int64_t, int64_t, const RsvdOpts&, int, qil_mps**);

template qil_mps* encode_rsvd<double>(qil_ctx*, const double*, int64_t, const RsvdOpts&);
template qil_mps* encode_rsvd<cplx>(qil_ctx*, const cplx*, int64_t, const RsvdOpts&);

}  // namespace qil
