/* qilcuda.h -- C ABI of libqilcuda.so, the B200 (sm_100a) implementation of the QILaplace.jl hot path.
 *
 * The reference (SUTD-MDQS/QILaplace.jl) has no FFI of its own: every function below replaces the
 * BODY of one Julia function, whose Index bookkeeping stays in Julia (see INTEGRATION.md for the
 * `ccall` stubs).  Citations are file:line inside the reference repository.
 *
 * Conventions
 *   - All entry points return an int status (QIL_OK == 0).  On failure `qil_last_error()` returns a
 *     thread-local message.  Status codes map onto the Julia exception types the reference's tests
 *     assert (ArgumentError, DomainError, ErrorException, AssertionError).
 *   - Buffers are plain host pointers unless the name ends in `_dev` (device pointers on the
 *     context's device).  The caller owns every buffer it passes; the library owns everything
 *     behind a handle.  Sizes of data-dependent results come from a `*_dims` query first.
 *   - Tensors are C-order (row-major): MPS core [l][s][r], MPO core [l][p][s][r] with p the
 *     primed/input leg and s the unprimed/output leg.  For a Julia caller that is exactly
 *     `Array(T, r, s, l)` / `Array(T, r, s, p, l)` of the ITensor (column-major, reversed order).
 *   - Scalars are float64 (`is_complex == 0`) or interleaved complex128 (`is_complex == 1`).
 *   - Signals are read MSB-first: sample j of a length-2^n vector has site-1 bit = top bit of j
 *     (src/signals/SignalConverters.jl:39-41, docs/src/core_concepts.md:34-41).
 *   - One host thread per context; every call is synchronous with respect to the host unless it
 *     ends in `_dev` (those are stream-ordered on the context's stream).
 */
#ifndef QILCUDA_H
#define QILCUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QIL_MAX_SITES 128

/* status codes */
#define QIL_OK 0
#define QIL_ERR_ARGUMENT 1    /* Julia ArgumentError  */
#define QIL_ERR_DOMAIN 2      /* Julia DomainError    */
#define QIL_ERR_RUNTIME 3     /* Julia ErrorException */
#define QIL_ERR_ASSERT 4      /* Julia AssertionError */
#define QIL_ERR_CUDA 5        /* CUDA runtime / driver failure */
#define QIL_ERR_UNSUPPORTED 6 /* shape outside what this build implements (never a silent fallback) */

typedef struct qil_ctx qil_ctx;
typedef struct qil_mps qil_mps;
typedef struct qil_mpo qil_mpo;
typedef struct qil_peer qil_peer;
typedef struct qil_uploader qil_uploader;

/* ---- context ------------------------------------------------------------------------------- */
const char* qil_last_error(void);
const char* qil_version(void);
int qil_create(int device, qil_ctx** out);
/* Same, but all work is ordered on an existing CUDA stream (e.g. torch's current stream). */
int qil_create_on_stream(int device, void* cuda_stream, qil_ctx** out);
int qil_destroy(qil_ctx* ctx);
int qil_sync(qil_ctx* ctx);
/* Number of kernels this library has launched on the context since creation. */
int qil_launch_count(qil_ctx* ctx, uint64_t* out);

/* Per-kernel-class device timing with CUDA events on the context's stream (used by bench.py for the
 * roofline numbers).  Classes: 0 = streaming sketch/projection GEMM (K1/K2), 1 = coefficient chain (K8),
 * 2 = apply (K6/K7), 3 = QR/TSQR (K3/K5), 4 = Jacobi SVD (K4).  qil_profile_read synchronises, returns the
 * summed milliseconds and launch count of one class since the last reset, and keeps the records. */
int qil_profile_enable(qil_ctx* ctx, int on);
int qil_profile_reset(qil_ctx* ctx);
int qil_profile_read(qil_ctx* ctx, int kernel_class, double* total_ms, int64_t* launches);
/* Same, plus the summed ALGORITHMIC bytes and flops the launches of that class declared (class 0: one read of the
 * streamed matrix view and 2 flops per element and sketch column; 0 for classes that do not state their work). */
int qil_profile_read_work(qil_ctx* ctx, int kernel_class, double* total_ms, int64_t* launches, double* bytes,
                          double* flops);

/* Closest truncation decision since the last reset, over every cutoff-ruled SVD the context has run (encoders, split,
 * compress, builders): min |w / (cutoff * sum sigma^2) - 1| where w is the discarded weight with and without the last
 * kept value.  Two correct implementations whose sigma^2 agree to better than this relative amount choose the same
 * bond dimensions; a value near rounding level flags a knife-edge rank ("rank parity").
 * 1e300 = no decision taken.  reset != 0 clears it after reading. */
int qil_truncation_margin(qil_ctx* ctx, int reset, double* out);

/* ---- MPS / MPO containers (replace Vector{ITensor} storage; src/mps.jl:70-130, src/mpo.jl:26-99) -- */
/* bond has n+1 entries with bond[0] == bond[n] == 1; cores[i] points at bond[i]*2*bond[i+1] scalars */
int qil_mps_from_host(qil_ctx* ctx, int n, int is_complex, const int64_t* bond,
                      const void* const* cores, double amplitude, qil_mps** out);
int qil_mps_info(const qil_mps* m, int* n, int* is_complex, double* amplitude);
int qil_mps_dims(const qil_mps* m, int64_t* bond /* n+1 */);
int qil_mps_get_core(const qil_mps* m, int site, void* host_buf);
/* every core, packed back to back in site order, with one synchronisation (pinned host memory makes it one DMA burst) */
int qil_mps_get_cores(const qil_mps* m, void* host_buf, int64_t host_bytes);
int qil_mps_set_amplitude(qil_mps* m, double amplitude);
int qil_mps_clone(const qil_mps* m, qil_mps** out);
int qil_mps_free(qil_mps* m);

int qil_mpo_from_host(qil_ctx* ctx, int n, int is_complex, const int64_t* bond,
                      const void* const* cores, qil_mpo** out);
int qil_mpo_info(const qil_mpo* m, int* n, int* is_complex);
int qil_mpo_dims(const qil_mpo* m, int64_t* bond /* n+1 */);
int qil_mpo_get_core(const qil_mpo* m, int site, void* host_buf);
int qil_mpo_free(qil_mpo* m);

/* ---- coefficient (src/mps.jl:669-693) --------------------------------------------------------
 * bits is uint8[B][n] with entries in {0,1}; out receives B scalars of the MPS element type,
 * already multiplied by the stored amplitude.  A ZTMPS is passed as its 2n-site chain
 * (main1, copy1, main2, ...; src/mps.jl:421-445). */
int qil_coefficient_batch(qil_ctx* ctx, const qil_mps* psi, const uint8_t* bits, int64_t B, void* out);
int qil_coefficient_batch_dev(qil_ctx* ctx, const qil_mps* psi, const uint8_t* d_bits, int64_t B,
                              void* d_out);

/* Dense grid of coefficients: every combination of the "free" sites at once.  Replaces the loops over
 * `coefficient` of the pole scans (docs/src/tutorials/zt.jl:152-157, 283-287, 330-411), of the full read-out
 * (docs/src/tutorials/signal.jl:149-150) and `mps_to_vector` (src/mps.jl:716-743; all sites free).
 * site_mode[i] (HOST array, n entries): 0 / 1 = site i fixed to that bit, 2 = free.  With F free sites the
 * result has 2^F scalars of the MPS element type, multiplied by the stored amplitude.  out_bit (HOST array,
 * F entries, or NULL): bit position, inside the linear output index, of the j-th free site in site order;
 * NULL = big-endian (first free site most significant -- the reference's Integer convention, mps.jl:633-645;
 * `mps_to_vector(reverse=false)`); out_bit[j] = j is `mps_to_vector(reverse=true)`. */
int qil_coefficient_grid(qil_ctx* ctx, const qil_mps* psi, const uint8_t* site_mode, const int32_t* out_bit, void* out);
int qil_coefficient_grid_dev(qil_ctx* ctx, const qil_mps* psi, const uint8_t* site_mode, const int32_t* out_bit,
                             void* d_out);

/* ---- pipelined host -> device staging -----------------------------------------------------------------------
 * A ring of `depth` device buffers of `bytes` each, filled on a private copy stream: the upload of the next signal
 * overlaps the work on the current one (end to end the path is PCIe bound: 2 GiB per n=28 signal vs a ~5 ms encode).
 *   qil_uploader_submit(u, host, bytes)  enqueue the H2D copy of one signal (pinned host memory makes it asynchronous;
 *                                        see qil_host_register) into the next free buffer; error if all are in flight
 *   qil_uploader_acquire(u, &d_ptr)      device pointer of the oldest submitted signal; the context's stream waits for
 *                                        its upload (no host synchronisation)
 *   ... qil_encode_rsvd_dev(ctx, ..., d_ptr, ...) or any other *_dev call on the context ...
 *   qil_uploader_release(u)              the buffer may be refilled once the work enqueued so far has finished
 * qil_host_register / qil_host_unregister page-lock an existing host buffer (cudaHostRegister) for callers whose arrays
 * are pageable (a Julia Vector): do it once per buffer, not per call. */
int qil_uploader_create(qil_ctx* ctx, int64_t bytes, int depth, qil_uploader** out);
int qil_uploader_submit(qil_uploader* u, const void* host, int64_t bytes);
int qil_uploader_acquire(qil_uploader* u, void** d_ptr);
int qil_uploader_release(qil_uploader* u);
int qil_uploader_destroy(qil_uploader* u);
int qil_host_register(void* host, int64_t bytes);
int qil_host_unregister(void* host);

/* ---- on-disk container (dims + raw cores; the reference itself persists only JLD2 benchmark dicts,
 * scripts/benchmark/common.jl:193-203).  Layout, little endian: "QILTN001" | u32 kind (0 MPS, 1 MPO) | u32 is_complex |
 * u32 n | u32 0 | f64 amplitude | i64 bond[n+1] | cores back to back in C order ([l][s][r] / [l][p][s][r]).  Written and
 * read by this library, by oracle/qil_container.py and by the Julia reader in julia/QILContainer.jl. */
int qil_mps_save(const qil_mps* m, const char* path);
int qil_mps_load(qil_ctx* ctx, const char* path, qil_mps** out);
int qil_mpo_save(const qil_mpo* m, const char* path);
int qil_mpo_load(qil_ctx* ctx, const char* path, qil_mpo** out);

/* ---- read-out reductions on the device ---------------------------------------------------------------------
 * arg-max of |chi| (first maximum on ties) -- the `argmax(abs.(chi))` that ends every stage of the coarse / fine /
 * superfine pole scan (docs/src/tutorials/zt.jl:324-326, 372-375, 412-415); only the index, |value| and value (one
 * scalar of the element type, may be NULL) come back to the host. */
int qil_argmax_abs_dev(qil_ctx* ctx, int is_complex, const void* d_values, int64_t count, int64_t* index, double* absval,
                       void* value);
/* qil_coefficient_grid / qil_coefficient_batch followed by the arg-max, without the D2H copy of the grid */
int qil_coefficient_grid_argmax(qil_ctx* ctx, const qil_mps* psi, const uint8_t* site_mode, const int32_t* out_bit,
                                int64_t* index, double* absval, void* value);
int qil_coefficient_batch_argmax(qil_ctx* ctx, const qil_mps* psi, const uint8_t* bits, int64_t B, int64_t* index,
                                 double* absval, void* value);
/* Sum over every configuration of the sites with sum_mask[i] != 0 (HOST array, n entries): they are contracted with
 * the all-ones vector and absorbed into a neighbouring kept site; the result is an MPS over the kept sites.  With the
 * copy register of a ZTMPS masked this is the inner sum of `laplace_coefficient` (docs/src/tutorials/dt.jl:187-197)
 * for every k at once: read it out with qil_coefficient_grid (all sites free). */
int qil_mps_sum_sites(qil_ctx* ctx, const qil_mps* psi, const uint8_t* sum_mask, qil_mps** out);
/* An MPS with uninitialised cores of the given bonds and the device address of a core: lets a multi-process host move
 * cores between devices with its own transport (e.g. an NCCL broadcast before a sharded pole scan). */
int qil_mps_alloc(qil_ctx* ctx, int n, int is_complex, const int64_t* bond, double amplitude, qil_mps** out);
int qil_mps_core_ptr(const qil_mps* m, int site, void** d_ptr, int64_t* elems);

/* ---- apply (src/linalg/apply.jl:75-122, 124-199, 201-236) ----------------------------------
 * Exact MPO x MPS: out core = [D_l*chi_l][2][D_r*chi_r] with the MPO bond fastest; never truncates;
 * amplitude is copied.  Paired operands are passed as 2n-site chains. */
int qil_apply_mpo_mps(qil_ctx* ctx, const qil_mpo* W, const qil_mps* psi, qil_mps** out);
/* Truncating MPO x MPS (zip-up): psi is brought to right-canonical form, then one sweep contracts carry x psi_i x W_i and
 * splits it with the truncated SVD of the path (cutoff relative and cumulative on sigma^2, maxdim <= 0 = no limit).  The
 * result has single (unfused) bonds.  Not a reference function: the remedy SURVEY.md 8f-3 names for the D*chi bonds of
 * the exact apply on high-rank inputs (docs/src/benchmarking.md:309). */
int qil_apply_mpo_mps_zipup(qil_ctx* ctx, const qil_mpo* W, const qil_mps* psi, double cutoff, int64_t maxdim, qil_mps** out);
/* `[W * psi for psi in psis]` (BASELINE configs[1]: one QFT MPO applied to a batch of encoded signals) in one launch;
 * the `count` results share one pooled device allocation. */
int qil_apply_mpo_mps_batch(qil_ctx* ctx, const qil_mpo* W, const qil_mps* const* psis, int64_t count, qil_mps** outs);
/* MPO o MPO over the matching window; W1 acts first; start1/start2 = first matching site (0-based). */
int qil_apply_mpo_mpo(qil_ctx* ctx, const qil_mpo* W1, const qil_mpo* W2, int start1, int start2,
                      qil_mpo** out);

/* ---- signal -> MPS (src/signals/SignalConverters.jl:16-104, 228-283) ------------------------
 * x holds N scalars (zero-padded to 2^round(log2 N) like the reference; N > 2^n is an AssertionError).
 * The result carries amplitude = ||x||_2.  maxdim <= 0 means "no limit" (typemax(Int)). */
int qil_encode_svd(qil_ctx* ctx, int is_complex, const void* x, int64_t N, double cutoff, int64_t maxdim,
                   qil_mps** out);
int qil_encode_svd_dev(qil_ctx* ctx, int is_complex, const void* d_x, int64_t N, double cutoff,
                       int64_t maxdim, qil_mps** out);
/* signal_mps(x; method=:rsvd, k, p, q, random_seed, cutoff, maxdim, mindim) -- divide-and-conquer TT with
 * randomized SVD (SignalConverters.jl:107-196, rsvd.jl:38-121).  Every split uses the same normal stream:
 * Omega[c][j] = stream[c + C*j] (the column-major `random_itensor` of the reference).  `normal_stream` may
 * be NULL (device generator seeded with `seed`) or point at `stream_len` scalars of the signal's type drawn
 * by the host (`Random.seed!(seed); randn(T, len)` in the Julia shim), len >= max over splits of C*l. */
/* flags: QIL_RSVD_ADAPTIVE = rank-adaptive sketch width at the top split (after the first QR of Y = A*Omega the
 * sketch columns whose |R_jj| is at rounding level are dropped for the remaining 2q+1 passes; needs q >= 1 and
 * cutoff >= 1e-18).  0 = the reference's fixed width k+p everywhere. */
#define QIL_RSVD_ADAPTIVE 1
int qil_encode_rsvd(qil_ctx* ctx, int is_complex, const void* x, int64_t N, int k, int p, int q, int64_t seed,
                    double cutoff, int64_t maxdim, int64_t mindim, const void* normal_stream, int64_t stream_len,
                    int64_t flags, qil_mps** out);
int qil_encode_rsvd_dev(qil_ctx* ctx, int is_complex, const void* d_x, int64_t N, int k, int p, int q, int64_t seed,
                        double cutoff, int64_t maxdim, int64_t mindim, const void* d_normal_stream,
                        int64_t stream_len, int64_t flags, qil_mps** out);
/* A batch of `count` independent signals of N samples each, stored back to back on the device (BASELINE configs[1]:
 * 256 signals of n = 20).  Same result per signal as qil_encode_rsvd_dev (every signal uses the same normal stream, like
 * the reference, which reseeds at every rsvd call).  Real power-of-two signals with k + p <= 32 are encoded in lock
 * step: the stacked signals are sketched by one streaming launch, projected by one split-K launch (one chunk per
 * signal), factored by batched TSQR / Jacobi kernels, and every tree level is one launch for all signals; the cores of
 * the batch share one pooled allocation.  Anything else falls back to `workers` host threads, each encoding whole
 * signals on its own stream (<= 0: default 16).  out[count] receives the handles.  Synchronous. */
int qil_encode_rsvd_batch_dev(qil_ctx* ctx, int is_complex, const void* d_x, int64_t N, int64_t count, int k, int p,
                              int q, int64_t seed, double cutoff, int64_t maxdim, int64_t mindim, int workers,
                              const void* d_normal_stream, int64_t stream_len, int64_t flags, qil_mps** out);

/* ---- one signal row-sharded over several devices (one process per device) ---------------------
 * The length-N signal is split in rank order into world contiguous chunks of N/world samples, i.e. into
 * leading-qubit (row) blocks of the top-level matrix of the divide and conquer (SignalConverters.jl:161).  The
 * sketch and projection GEMMs stream each rank's own block; the exchange steps are: an 8-byte all-reduce of the
 * sum of squares, an all-gather of the per-rank l x l R factors of the tall-skinny QR (once per QR), an all-reduce
 * of the partial l x 2^ceil(n/2) projections (once per power iteration and for B = Q^H A), and an all-gather of the
 * row blocks of U.  The library stays free of any communication dependency: the host supplies the two
 * collectives (torch.distributed / NCCL in the Python host, NCCL.jl or MPI.jl in a Julia host).  Both callbacks
 * receive DEVICE pointers to float64 data (complex = interleaved pairs), must be ordered on the context's stream
 * (see qil_get_stream) and return 0 on success.  Every rank must make the same call; every rank receives the
 * same MPS. */
typedef struct qil_comm {
    int rank;
    int world;
    void* user;
    int (*allreduce_sum_f64)(void* user, void* d_buf, int64_t count);                          /* in place */
    int (*allgather_f64)(void* user, const void* d_send, void* d_recv, int64_t count_per_rank); /* rank order */
} qil_comm;
int qil_get_stream(qil_ctx* ctx, void** cuda_stream);
int qil_encode_rsvd_sharded_dev(qil_ctx* ctx, const qil_comm* comm, int is_complex, const void* d_x_local,
                                int64_t N_total, int k, int p, int q, int64_t seed, double cutoff, int64_t maxdim,
                                int64_t mindim, const void* d_normal_stream, int64_t stream_len, int64_t flags,
                                qil_mps** out);

/* Native implementation of the two collectives over NVLink peer memory (qil_peer.cu): every rank allocates a
 * symmetric exchange buffer of `bytes` payload (>= the largest message: 16 * l * 2^ceil(n/2) bytes covers every
 * exchange of an encode) and publishes its 64-byte CUDA-IPC handle; the host gathers the handles of all ranks in
 * rank order (any transport) and passes them to qil_peer_connect.  qil_peer_comm then fills a qil_comm whose
 * callbacks launch the library's own exchange kernels on the context's stream: the reduction runs inside the
 * exchange kernel over the peers' memory, with flag-based arrival through the same mapping.  No NCCL involved. */
int qil_peer_create(qil_ctx* ctx, int rank, int world, int64_t bytes, qil_peer** out, unsigned char* handle64);
int qil_peer_connect(qil_peer* peer, const unsigned char* all_handles /* world * 64 bytes, rank order */);
int qil_peer_comm(qil_peer* peer, qil_comm* out);
int qil_peer_destroy(qil_peer* peer);

/* Per-site copy-tensor split of signal_ztmps (SignalConverters.jl:258-277): n-site MPS -> 2n-site chain. */
int qil_ztmps_split(qil_ctx* ctx, const qil_mps* psi, double cutoff, int64_t maxdim, qil_mps** out);

/* ---- gauge / compression (src/mps.jl:754-999), in place on the handle ------------------------ */
/* direction: 0 = :left (sweep N..center+1), 1 = :right (sweep 1..center-1); center 1-based, 0 = default */
int qil_canonicalize(qil_ctx* ctx, qil_mps* psi, int direction_right, int center, double cutoff, int64_t maxdim);
int qil_compress(qil_ctx* ctx, qil_mps* psi, int64_t maxdim, double tol, int sweeps);
int qil_norm(qil_ctx* ctx, const qil_mps* psi, double* out);

/* ---- transform MPO builders (src/transforms/{qft,dt,zt}_transformer.jl) ------------------------
 * dt/zt return the 2n-site chain of the PairedSiteMPO (main1, copy1, main2, ...). */
int qil_build_qft_mpo(qil_ctx* ctx, int n, double cutoff, int64_t maxdim, qil_mpo** out);
int qil_build_dt_mpo(qil_ctx* ctx, int n, double omega_r, double cutoff, int64_t maxdim, qil_mpo** out);
int qil_build_zt_mpo(qil_ctx* ctx, int n, double omega_r, double cutoff, int64_t maxdim, qil_mpo** out);

/* ---- dense factorizations on host matrices (row-major), the ITensors calls of the path ----------
 * qil_qr: thin QR, k = min(m,n); Q is m x k, R is k x n; positive != 0 => diag(R) >= 0 (rsvd.jl:83).
 * qil_svd_trunc: truncated SVD with the NDTensors rule (relative cumulative cutoff on sigma^2, maxdim,
 * mindim).  U / S / Vh must have room for min(m,n) columns / values / rows; they are written
 * compactly as m x r, r, r x n and *rank receives r. */
int qil_qr(qil_ctx* ctx, int is_complex, int64_t m, int64_t n, const void* A, int positive, void* Q, void* R);
int qil_svd_trunc(qil_ctx* ctx, int is_complex, int64_t m, int64_t n, const void* A, double cutoff,
                  int64_t maxdim, int64_t mindim, int64_t* rank, void* U, double* S, void* Vh);
/* rsvd(A, Linds...; k=20, p=10, q=0, random_seed, cutoff=1e-15, maxdim=k, mindim=1) (src/linalg/rsvd.jl:38-121)
 * on a host matrix; outputs as for qil_svd_trunc (room for min(k+p, m, n) columns/rows).  An empty index set
 * (m == 0 or n == 0) is an ErrorException like the reference (rsvd.jl:56-60). */
int qil_rsvd(qil_ctx* ctx, int is_complex, int64_t m, int64_t n, const void* A, int k, int p, int q, int64_t seed,
             double cutoff, int64_t maxdim, int64_t mindim, const void* normal_stream, int64_t stream_len,
             int64_t* rank, void* U, double* S, void* Vh);

#ifdef __cplusplus
}
#endif
#endif /* QILCUDA_H */
