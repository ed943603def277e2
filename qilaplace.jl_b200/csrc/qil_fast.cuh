// qil_fast.cuh -- host entry points of the round-2 latency path (qil_tsqr.cu, qil_rsvd_fast.cu).
#pragma once
#include "qil_mpsops.cuh"

namespace qil {

// Thin QR of an m x n panel (m >= n, n <= 32) with the warp-synchronous TSQR.  A may be the sum of `nsum`
// partial matrices `sum_stride` elements apart.  Q (m x n) is written with leading dimension ldq, columns
// n..qcols-1 are zero filled; R (n x n, ld n) may be null.  `batch` equal problems a_bs / q_bs / r_bs apart.
template <typename T> bool qr_fast_supported(qil_ctx* ctx, int64_t m, int64_t n);
template <typename T>
void qr_fast(qil_ctx* ctx, int64_t m, int n, const T* A, int64_t lda, int nsum, int64_t sum_stride, bool positive,
             T* Q, int64_t ldq, int qcols, T* R, int batch = 1, int64_t a_bs = 0, int64_t q_bs = 0, int64_t r_bs = 0);

// SVD of the l x C matrix B = Rb^H Qb^H given by the QR of its adjoint (Bh = Qb Rb, Qb: C x l, Rb: l x l):
// one-sided Jacobi on G = scale * Rb^H by one warp, NDTensors truncation on the device.
//   Us (l x r, ld r) = left singular vectors, T2 (r x l, ld l) = Us^H G = S Vh' (so that S Vh = T2 Qb^H),
//   S (l, first r meaningful), rank (device int).
template <typename T>
void svd_finish(qil_ctx* ctx, int l, const T* Rb, const double* d_scale, double cutoff, int64_t maxdim, int64_t mindim,
                T* Us, T* T2, double* S, int* d_rank, int batch = 1, int scale_bs = 0, int rank_bs = 0);

// U (R x r, ld r) = Q[:, :l] Us ;  SVh (r x C, ld C) = T2 Qb^H  (or Vh = diag(1/S) T2 Qb^H when vh_only)
template <typename T>
void rsvd_outputs(qil_ctx* ctx, int64_t R, int64_t C, int l, const T* Q, int64_t ldq, const T* Qb, int64_t ldqb,
                  const T* Us, const T* T2, const double* S, const int* d_rank, T* U, T* SVh, T* Vh, int batch = 1,
                  int64_t q_bs = 0, int64_t qb_bs = 0, int64_t u_bs = 0, int64_t sv_bs = 0, int rank_bs = 0);

// ---- one CTA per tree node (qil_node.cu) -------------------------------------------------------------------
// Node matrix A = T[lb * 2^nl, 2^nr * rb] (compact, row-major) with lb = bonds[bonds_off + lb_pos], rb likewise;
// outputs U (R x r, ld r) and SVh (r x C, ld C) and the new bond bonds[bonds_off + out_pos] = r.
struct NodeDesc {
    const double* A;
    double* U;
    double* SVh;
    int bonds_off;
    int lb_pos, rb_pos, out_pos;
    int nl, nr;
};
void node_level_launch(qil_ctx* ctx, const NodeDesc* d_nodes, int count, int* d_bonds, int* d_overflow, const RsvdOpts& o,
                       const double* d_stream, int64_t stream_len);

}  // namespace qil
