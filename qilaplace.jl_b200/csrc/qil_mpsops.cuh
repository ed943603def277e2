// qil_mpsops.cuh -- MPS-level algorithms (see qil_mpsops.cu, qil_encode.cu, qil_builders.cu).
#pragma once
#include "qil_dense.cuh"

namespace qil {

int ilog2_round(int64_t N);
template <typename T> double device_norm2(qil_ctx* ctx, const T* x, int64_t n);

// signal_mps(x; method=:svd) -- d_x is a device pointer to N scalars
template <typename T> qil_mps* encode_svd(qil_ctx* ctx, const T* d_x, int64_t N, double cutoff, int64_t maxdim);

struct RsvdOpts {
    int k = 20, p = 10, q = 0;
    long long seed = 1234;
    double cutoff = 1e-15;
    int64_t maxdim = (int64_t)1 << 62;
    int64_t mindim = 1;
    const void* omega = nullptr;   // optional host-supplied test matrix (cols x l, row-major), device pointer
    int64_t omega_rows = 0, omega_cols = 0;
    bool adaptive = false;         // rank-adaptive sketch width at the top split (flags bit 0; off = reference behaviour)
};
// signal_mps(x; method=:rsvd, ...)
template <typename T> qil_mps* encode_rsvd(qil_ctx* ctx, const T* d_x, int64_t N, const RsvdOpts& o);

// `count` independent signals of N samples each, stored back to back; `workers` host threads / streams
template <typename T>
void encode_rsvd_batch(qil_ctx* ctx, const T* d_x, int64_t N, int64_t count, const RsvdOpts& o, int workers,
                       qil_mps** out);

// same with the signal row-sharded over the ranks of `comm` (host-supplied collectives, include/qilcuda.h)
template <typename T>
qil_mps* encode_rsvd_sharded(qil_ctx* ctx, const qil_comm* comm, const T* d_x_local, int64_t N, const RsvdOpts& o);

template <typename T>
int rsvd_matrix(qil_ctx* ctx, const T* d_A, int64_t m, int64_t n, const RsvdOpts& o, Mat<T>& U, Mat<double>& S, Mat<T>& Vh);

qil_mps* ztmps_split(qil_ctx* ctx, const qil_mps* psi, double cutoff, int64_t maxdim);
void canonicalize(qil_ctx* ctx, qil_mps* psi, int dir_right, int center, double cutoff, int64_t maxdim);
void compress(qil_ctx* ctx, qil_mps* psi, int64_t maxdim, double tol, int sweeps);
double mps_norm(qil_ctx* ctx, const qil_mps* psi);
qil_mps* apply_mpo_mps_zipup(qil_ctx* ctx, const qil_mpo* W, const qil_mps* psi, double cutoff, int64_t maxdim);
// read-out reductions (qil_scan.cu)
template <typename T> void argmax_abs(qil_ctx* ctx, const T* d_v, int64_t count, int64_t* index, double* absval, T* value);
qil_mps* mps_sum_sites(qil_ctx* ctx, const qil_mps* psi, const uint8_t* mask);

qil_mpo* build_qft_mpo(qil_ctx* ctx, int n, double cutoff, int64_t maxdim);
qil_mpo* build_dt_mpo(qil_ctx* ctx, int n, double wr, double cutoff, int64_t maxdim);
qil_mpo* build_zt_mpo(qil_ctx* ctx, int n, double wr, double cutoff, int64_t maxdim);

}  // namespace qil

namespace qil {
// streaming DMMA/TMA GEMMs (qil_sketch.cu)
int stream_nt_for(int cols);
// pitch (doubles) of the X operand: 8*nt + 1 -- odd, so that the B-fragment loads of the DMMA consumers (rows 4t+i,
// column g) hit 16 distinct bank pairs per half warp (8*nt + 2 gave a 2-way conflict on every load)
// pitch (doubles) of the small operand X of the streaming GEMM, chosen so that the B-fragment loads of a half-warp fall into
// distinct banks: K1 reads k-rows 4t+i (pitch = 1 mod 8), K2 reads k-rows 2t + (i&1) + 8(i>>1) (pitch = 2 mod 8), the order
// that also makes its A-fragment loads from the 128B-swizzled transposed boxes conflict-free (qil_sketch.cu)
inline int stream_lpp(int nt, bool trans = false) { return nt * 8 + (trans ? 2 : 1); }
bool stream_supported(long long R, long long C, long long ld, int cols);
void stream_plan(qil_ctx* ctx, long long Mtot, long long Kdim, int* ksplit, long long* kchunk, int nt = 1);
int stream_grid(qil_ctx* ctx, long long Mtot, int ksplit, int nt);
void stream_gemm(qil_ctx* ctx, bool trans, const double* A, long long R, long long C, long long ld, const double* X,
                 int lpp, int nt, double* out, int ksplit, long long kchunk, double* sumsq_partials,
                 int ncols /* real sketch columns, for the profiler's algorithmic flops */,
                 long long x_rows = 0 /* K1 over a stack of equal matrices: rows per matrix (multiple of 128) ... */,
                 long long x_bs = 0 /* ... each with its own X operand, x_bs doubles apart */);
}  // namespace qil
