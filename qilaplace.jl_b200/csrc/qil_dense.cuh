// qil_dense.cuh -- small dense toolkit used by the encode / compress / builder orchestration:
// generic strided contraction, tiled GEMM, Householder QR (single CTA + TSQR), one-sided Jacobi SVD
// with the ITensors truncation rule.  Everything is stream-ordered on ctx->stream.
#pragma once
#include "qil_common.cuh"

namespace qil {

enum Op { OP_N = 0, OP_T = 1, OP_C = 2 };

// ---- owning device matrix (row-major, leading dimension == cols) ---------------------------
template <typename T>
struct Mat {
    qil_ctx* ctx = nullptr;
    T* p = nullptr;
    int64_t rows = 0, cols = 0;
    Mat() {}
    Mat(qil_ctx* c, int64_t r, int64_t cc) : ctx(c), rows(r), cols(cc) {
        p = (T*)c->alloc((size_t)std::max<int64_t>(r * cc, 1) * sizeof(T));
    }
    Mat(const Mat&) = delete;
    Mat& operator=(const Mat&) = delete;
    Mat(Mat&& o) noexcept { *this = std::move(o); }
    Mat& operator=(Mat&& o) noexcept {
        if (this != &o) {
            release();
            ctx = o.ctx; p = o.p; rows = o.rows; cols = o.cols;
            o.p = nullptr;
        }
        return *this;
    }
    ~Mat() { release(); }
    void release() {
        if (p && ctx) ctx->free(p);
        p = nullptr;
    }
    T* take() { T* q = p; p = nullptr; return q; }
    size_t elems() const { return (size_t)rows * cols; }
};

// ---- generic contraction ----------------------------------------------------------------------
// C[o0..o5] = alpha * sum_{k0..k2} opA(A[...]) * B[...]   with explicit element strides.
struct ContractDesc {
    int nout;                 // number of output dims (<= 6)
    int ncon;                 // number of contracted dims (<= 3)
    long long od[6];          // output extents
    long long cd[3];          // contracted extents
    long long sa_o[6], sb_o[6], sc_o[6];  // strides of A, B, C along each output dim (0 if absent)
    long long sa_c[3], sb_c[3];           // strides of A, B along each contracted dim
    int conj_a;               // conjugate A's elements
};
template <typename TA, typename TB, typename TC>
void contract(qil_ctx* ctx, const ContractDesc& d, const TA* A, const TB* B, TC* C);

// C(MxN) = alpha * op(A) * op(B) + beta * C, row-major with leading dimensions
template <typename T>
void gemm(qil_ctx* ctx, Op opa, Op opb, int64_t M, int64_t N, int64_t K, double alpha, const T* A, int64_t lda,
          const T* B, int64_t ldb, double beta, T* C, int64_t ldc);

// C(MxN) = alpha * A * B on the FP64 tensor path (DMMA m16n8k16, cp.async ring; qil_grid.cu); no transposes
template <typename T>
void gemm_tc(qil_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha, const T* A, int64_t lda, const T* B,
             int64_t ldb, T* C, int64_t ldc);

// y[i] = alpha * x[i]  (n elements); row/col scaling helpers
template <typename T> void scale_copy(qil_ctx* ctx, int64_t n, double alpha, const T* x, T* y);
// B[i][j] = A[i][j] * s[j] (col) or * s[i] (row); inv => divide (0 -> 0)
template <typename T>
void scale_rows_cols(qil_ctx* ctx, int64_t m, int64_t n, const T* A, int64_t lda, const double* s, bool by_row,
                     bool inv, T* B, int64_t ldb);
// B = A^H (conjugate transpose), A is m x n with lda
template <typename T>
void transpose_conj(qil_ctx* ctx, int64_t m, int64_t n, const T* A, int64_t lda, T* B, int64_t ldb, bool conj = true);

// ---- QR ------------------------------------------------------------------------------------------
// Thin QR of A (m x n, row-major, lda) -> Q (m x k), R (k x n), k = min(m,n).  Householder based
// (single CTA when the block fits shared memory, TSQR for tall matrices), so Q is orthonormal to
// rounding even for rank-deficient A.  positive => diag(R) real and >= 0 (ITensors qr(...; positive=true)).
// nsum > 1: A is given as nsum partial matrices, `sum_stride` elements apart, that are summed on load.
template <typename T>
void qr_thin(qil_ctx* ctx, int64_t m, int64_t n, const T* A, int64_t lda, bool positive, Mat<T>& Q, Mat<T>& R,
             int nsum = 1, int64_t sum_stride = 0, bool want_q = true);

// ---- truncated SVD ---------------------------------------------------------------------------------
struct SvdOut {
    int rank = 0;
};
// A (m x n, row-major, lda) ~ U diag(S) Vh with the NDTensors truncation rule on S^2
// (relative cumulative cutoff, maxdim, mindim).  Any of U (m x r), US (m x r), Vh (r x n), SVh (r x n),
// S (r doubles, device) may be requested (pass nullptr to skip).  Returns the kept rank (host sync).
template <typename T>
int svd_trunc(qil_ctx* ctx, int64_t m, int64_t n, const T* A, int64_t lda, double cutoff, int64_t maxdim,
              int64_t mindim, Mat<T>* U, Mat<T>* US, Mat<T>* Vh, Mat<T>* SVh, Mat<double>* S,
              int nsum = 1, int64_t sum_stride = 0);


// Same for a wide A (m <= n) whose adjoint At = A^H (n x m, row-major, pitch ldat) is what the caller holds
// (the randomized SVD produces B^H = A^H Q directly).
template <typename T>
int svd_trunc_adj(qil_ctx* ctx, int64_t m, int64_t n, const T* At, int64_t ldat, double cutoff, int64_t maxdim,
                  int64_t mindim, Mat<T>* U, Mat<T>* US, Mat<T>* Vh, Mat<T>* SVh, Mat<double>* S);


// Fused single-launch variant for matrices that fit one CTA's shared memory (qil_svd_small.cu)
template <typename T> bool svd_small_fits(qil_ctx* ctx, int64_t m, int64_t n);
template <typename T>
int svd_small(qil_ctx* ctx, int64_t m, int64_t n, const T* A, int64_t lda, double cutoff, int64_t maxdim,
              int64_t mindim, Mat<T>* U, Mat<T>* US, Mat<T>* Vh, Mat<T>* SVh, Mat<double>* S);

// Batch of independent small SVDs (one launch, one sync).  `copy_tensor`: A points at an MPS core [cl][2][cr]
// and the decomposed matrix is T[(l,s),(s',r)] = delta(s,s') core[l,s,r] (m = 2 cl, n = 2 cr).
template <typename T>
struct SmallSvdItem {
    const T* A = nullptr;
    int64_t lda = 0, m = 0, n = 0;
    bool copy_tensor = false;
    int cl = 0, cr = 0;
    bool want_U = false, want_US = false, want_Vh = false, want_SVh = false;
    Mat<T> U, US, Vh, SVh;
    int rank = 0;
};
template <typename T>
// pool_out != nullptr: all outputs are non-owning views into ONE allocation, handed back through *pool_out
void svd_small_batch(qil_ctx* ctx, std::vector<SmallSvdItem<T>>& items, double cutoff, int64_t maxdim, int64_t mindim,
                     std::shared_ptr<void>* pool_out = nullptr);

}  // namespace qil
