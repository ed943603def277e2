// qil_common.cuh -- shared declarations for libqilcuda (sm_100a only).
//
// Layout conventions (row-major / C order everywhere on the device):
//   MPS core i : M[l][s][r]      dims bond[i] x 2 x bond[i+1]          (reference: ITensor (l,s,r))
//   MPO core i : W[l][p][s][r]   dims bond[i] x 2 x 2 x bond[i+1], p = primed/input, s = output
//   signal     : x[j], j MSB-first == site 1 is the most significant bit
// Scalars are double or interleaved complex double (cuDoubleComplex-compatible double2).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <exception>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/qilcuda.h"

namespace qil {

struct Error : public std::exception {
    int code;
    std::string msg;
    Error(int c, std::string m) : code(c), msg(std::move(m)) {}
    const char* what() const noexcept override { return msg.c_str(); }
};

#define QIL_THROW(code, ...)                                   \
    do {                                                       \
        char _b[512];                                          \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                 \
        throw ::qil::Error((code), std::string(_b));           \
    } while (0)

#define QIL_CUDA(expr)                                                                    \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            QIL_THROW(QIL_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                      __FILE__, __LINE__);                                                \
    } while (0)

#define QIL_REQUIRE(cond, code, ...) \
    do {                             \
        if (!(cond)) QIL_THROW(code, __VA_ARGS__); \
    } while (0)

constexpr int kMaxSites = QIL_MAX_SITES;

// A collective callback can only return a status; the library's own callbacks leave the reason here (per host thread)
// and the caller of the callback appends it to the error it raises.
inline std::string& callback_error_note() {
    static thread_local std::string note;
    return note;
}

// ---- complex helpers (double2 == interleaved complex) -------------------------------------
typedef double2 cplx;

template <typename T> struct Scalar;
template <> struct Scalar<double> {
    static constexpr bool is_complex = false;
    __host__ __device__ static inline double zero() { return 0.0; }
    __host__ __device__ static inline double one() { return 1.0; }
    __host__ __device__ static inline double conj(double a) { return a; }
    __host__ __device__ static inline double mul(double a, double b) { return a * b; }
    __host__ __device__ static inline double fma(double a, double b, double c) { return a * b + c; }
    __host__ __device__ static inline double add(double a, double b) { return a + b; }
    __host__ __device__ static inline double sub(double a, double b) { return a - b; }
    __host__ __device__ static inline double scale(double a, double s) { return a * s; }
    __host__ __device__ static inline double abs2(double a) { return a * a; }
    __host__ __device__ static inline double real(double a) { return a; }
    __host__ __device__ static inline double from_real(double a) { return a; }
};
template <> struct Scalar<cplx> {
    static constexpr bool is_complex = true;
    __host__ __device__ static inline cplx zero() { return make_double2(0.0, 0.0); }
    __host__ __device__ static inline cplx one() { return make_double2(1.0, 0.0); }
    __host__ __device__ static inline cplx conj(cplx a) { return make_double2(a.x, -a.y); }
    __host__ __device__ static inline cplx mul(cplx a, cplx b) {
        return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
    }
    __host__ __device__ static inline cplx fma(cplx a, cplx b, cplx c) {
        c.x = ::fma(a.x, b.x, c.x);
        c.x = ::fma(-a.y, b.y, c.x);
        c.y = ::fma(a.x, b.y, c.y);
        c.y = ::fma(a.y, b.x, c.y);
        return c;
    }
    __host__ __device__ static inline cplx add(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
    __host__ __device__ static inline cplx sub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
    __host__ __device__ static inline cplx scale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
    __host__ __device__ static inline double abs2(cplx a) { return a.x * a.x + a.y * a.y; }
    __host__ __device__ static inline double real(cplx a) { return a.x; }
    __host__ __device__ static inline cplx from_real(double a) { return make_double2(a, 0.0); }
};

// promote real -> complex on load
template <typename TO, typename TI> __host__ __device__ inline TO promote(TI v);
template <> __host__ __device__ inline double promote<double, double>(double v) { return v; }
template <> __host__ __device__ inline cplx promote<cplx, double>(double v) { return make_double2(v, 0.0); }
template <> __host__ __device__ inline cplx promote<cplx, cplx>(cplx v) { return v; }

inline size_t elem_size(int is_complex) { return is_complex ? 16 : 8; }

}  // namespace qil

// ---- opaque handles --------------------------------------------------------------------------
struct qil_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    size_t smem_optin = 0;
    unsigned long long launches = 0;  // kernels launched by this library on this context
    // scratch that lives as long as the context
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // two zero-initialised words (arrival counter, generation) for the software grid barrier of cooperative kernels
    // launched on `stream` (one such kernel at a time per context: launches on one stream are serialised)
    unsigned int* grid_sync = nullptr;
    bool tsqr_fused_ok = true;        // false: tall QRs take the three-launch TSQR (see tsqr_sharded in qil_encode.cu)
    unsigned int* get_grid_sync();
    qil_ctx() = default;
    qil_ctx(const qil_ctx&) = delete;
    qil_ctx& operator=(const qil_ctx&) = delete;
    ~qil_ctx();

    // optional per-kernel-class timing (CUDA events on `stream`), enabled by qil_profile_enable
    struct ProfRegion { int id; cudaEvent_t e0, e1; double bytes, flops; };
    bool prof_on = false;
    int prof_depth = 0;               // regions nest: only the outermost one is recorded
    std::vector<ProfRegion> prof;
    // bytes / flops: ALGORITHMIC work of the region (what the roofline is computed from), 0 if not stated
    void prof_begin(int id, double bytes = 0.0, double flops = 0.0);
    void prof_end();

    // closest truncation decision since the last reset (device scalar, shared with the auxiliary contexts):
    // min over all cutoff decisions of |discarded-or-kept weight / (cutoff * total) - 1|  -- see truncation_margin
    double* d_margin = nullptr;

    // auxiliary contexts (own non-blocking stream each) for independent sub-problems run by worker threads
    bool is_aux = false;
    std::vector<qil_ctx*> aux;
    qil_ctx* aux_ctx(int w);             // created on first use, destroyed with the context
    void release_aux();

    void* alloc(size_t bytes);           // stream-ordered device allocation
    void free(void* p);                  // stream-ordered free
    void* get_scratch(size_t bytes);     // grow-only scratch (valid until next get_scratch)
    void sync();
};

// profiler region that closes on every exit path (an exception thrown inside would otherwise leave prof_depth > 0 and
// silently disable later profiling)
struct qil_prof_region {
    qil_ctx* c;
    qil_prof_region(qil_ctx* ctx, int id, double bytes = 0.0, double flops = 0.0) : c(ctx) { c->prof_begin(id, bytes, flops); }
    ~qil_prof_region() { c->prof_end(); }
    qil_prof_region(const qil_prof_region&) = delete;
    qil_prof_region& operator=(const qil_prof_region&) = delete;
};

struct qil_mps {
    qil_ctx* ctx = nullptr;
    int n = 0;
    int is_complex = 0;
    std::vector<int64_t> bond;  // n+1 entries, bond[0] = bond[n] = 1
    std::vector<void*> core;    // device pointers, core[i] is [bond[i]][2][bond[i+1]]
    // non-null: the cores are sub-allocations of one pooled buffer shared by the MPS of a batch (freed with the last
    // of them); in-place operations call qil::unpool() first
    std::shared_ptr<void> pool;
    double amplitude = 1.0;
    size_t core_elems(int i) const { return (size_t)bond[i] * 2 * (size_t)bond[i + 1]; }
};

struct qil_mpo {
    qil_ctx* ctx = nullptr;
    int n = 0;
    int is_complex = 0;
    std::vector<int64_t> bond;  // n+1 entries
    std::vector<void*> core;    // core[i] is [bond[i]][2][2][bond[i+1]]
    size_t core_elems(int i) const { return (size_t)bond[i] * 4 * (size_t)bond[i + 1]; }
};

namespace qil {

// chain descriptor passed by value to kernels that walk all sites in one launch
struct ChainDesc {
    int n;
    int bond[kMaxSites + 1];
    const void* core[kMaxSites];
};

ChainDesc make_desc(const qil_mps* m);
ChainDesc make_desc(const qil_mpo* m);

qil_mps* new_mps(qil_ctx* ctx, int n, int is_complex, const int64_t* bond /* n+1 */, bool allocate);
qil_mpo* new_mpo(qil_ctx* ctx, int n, int is_complex, const int64_t* bond /* n+1 */, bool allocate);
void destroy(qil_mps* m);
void unpool(qil_mps* m);          // give a pooled MPS its own core allocations
void destroy(qil_mpo* m);

// owner of a chain handle until it is handed to the caller: an exception on the way releases it
template <typename H>
struct chain_owner {
    H* p;
    explicit chain_owner(H* h = nullptr) : p(h) {}
    chain_owner(const chain_owner&) = delete;
    chain_owner& operator=(const chain_owner&) = delete;
    ~chain_owner() { reset(); }
    void reset(H* h = nullptr) { if (p) destroy(p); p = h; }
    H* release() { H* h = p; p = nullptr; return h; }
    H* get() const { return p; }
    H* operator->() const { return p; }
};

// ordering-only event that is destroyed on every way out of its scope
struct scoped_event {
    cudaEvent_t e = nullptr;
    scoped_event() { QIL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); }
    scoped_event(const scoped_event&) = delete;
    scoped_event& operator=(const scoped_event&) = delete;
    ~scoped_event() { if (e) cudaEventDestroy(e); }
    operator cudaEvent_t() const { return e; }
};

// host -> device staging ring (qil_upload.cu)
qil_uploader* uploader_create(qil_ctx* ctx, int64_t bytes, int depth);
void uploader_submit(qil_uploader* u, const void* host, int64_t bytes);
void* uploader_acquire(qil_uploader* u);
void uploader_release(qil_uploader* u);
void uploader_destroy(qil_uploader* u);

// native collectives over peer memory (qil_peer.cu)
qil_peer* peer_create(qil_ctx* ctx, int rank, int world, int64_t bytes, unsigned char* handle64);
void peer_connect(qil_peer* p, const unsigned char* all_handles);
void peer_fill_comm(qil_peer* p, qil_comm* out);
void peer_destroy(qil_peer* p);

enum ProfId { PROF_STREAM_GEMM = 0, PROF_COEFF = 1, PROF_APPLY = 2, PROF_QR = 3, PROF_SVD = 4, PROF_COUNT = 5 };

// NDTensors truncate!! on P = sigma^2 (descending): drop while n > maxdim, then while the discarded weight stays
// <= cutoff * sum(P) and n > mindim (relative, cumulative).  `margin` (optional device scalar) receives, by atomic
// min, how close the cutoff rule came to deciding otherwise: min(|w_keep / t - 1|, |w_drop / t - 1|) with
// t = cutoff * sum(P), w_keep = discarded weight if the last kept value were dropped too, w_drop = discarded weight.
// Two implementations whose sigma^2 differ by less than that relative amount pick the same rank.
__device__ inline int truncate_rank_dev(const double* sig, int n, double cutoff, long long maxdim, long long mindim,
                                        double* margin = nullptr) {
    if (n <= 1) return n;
    int r = n;
    double err = 0.0;
    while ((long long)r > maxdim) { err += sig[r - 1] * sig[r - 1]; --r; }
    double scale = 0.0;
    for (int i = 0; i < n; ++i) scale += sig[i] * sig[i];
    if (scale == 0.0) scale = 1.0;
    const double t = cutoff * scale;
    bool dropped = false;
    while ((long long)r > mindim && err + sig[r - 1] * sig[r - 1] <= t) {
        err += sig[r - 1] * sig[r - 1];
        --r;
        dropped = true;
    }
    if (r < 1) r = 1;
    if (margin && t > 0.0) {
        double m = 1e300;
        if ((long long)r > mindim) m = fmin(m, fabs((err + sig[r - 1] * sig[r - 1]) / t - 1.0));   // kept by the cutoff
        if (dropped) m = fmin(m, fabs(err / t - 1.0));                                            // dropped by it
        if (m < 1e300) atomicMin(reinterpret_cast<unsigned long long*>(margin), (unsigned long long)__double_as_longlong(m));
    }
    return r;
}

// One-sided Jacobi: threshold below which two columns need not be orthogonalised AGAINST EACH OTHER.  Columns whose
// squared norm is below nu = 1e-3 * cutoff * ||G||_F^2 / ns carry, all together, less than 1e-3 of the weight the
// truncation rule may discard, so every one of them is dropped whatever their mutual angles are; their summed weight
// (what the rule accumulates) is invariant under the skipped rotations, and each of them is still rotated against
// every significant column, so the kept singular triplets are exact.  Only when the rule cannot be forced to keep
// such a column (mindim <= 1) and cutoff > 0; otherwise nu = 0 and nothing is skipped.  For the low-rank bond
// matrices of this path (a handful of significant columns out of 20 .. 500) it removes most sweeps.
__host__ __device__ inline double jacobi_skip_threshold(double total, int ns, double cutoff, long long mindim) {
    return (cutoff > 0.0 && mindim <= 1) ? 1e-3 * cutoff * total / (double)ns : 0.0;
}

// Raise (never lower) a kernel's dynamic shared-memory limit.  The attribute is process-wide per function, so
// concurrent host threads (batched encode) must not shrink what another thread is about to launch with; keeping
// the running maximum also removes a driver call from every launch after the first.
void ensure_dynamic_smem_impl(const void* func, size_t bytes);
template <typename F>
inline void ensure_dynamic_smem(F kern, size_t bytes) { ensure_dynamic_smem_impl(reinterpret_cast<const void*>(kern), bytes); }

#define QIL_LAUNCH_CHECK(ctx)            \
    do {                                 \
        (ctx)->launches++;               \
        QIL_CUDA(cudaGetLastError());    \
    } while (0)

// ---- kernels / device-side ops implemented across the .cu files -----------------------------
// K8: batched coefficient extraction (mps.jl:669-678)
void coefficient_batch_dev(qil_ctx* ctx, const qil_mps* psi, const uint8_t* d_bits, int64_t B,
                           void* d_out /* B scalars of psi's type */);
// Dense grid of coefficients over the free sites (mode[i] == 2; 0/1 = fixed bit), meet-in-the-middle GEMMs
// (qil_grid.cu).  mode / out_bit are HOST arrays; d_out receives 2^F scalars of psi's type.
void coefficient_grid_dev(qil_ctx* ctx, const qil_mps* psi, const uint8_t* mode, const int32_t* out_bit, void* d_out);
// K6: exact MPO x MPS (apply.jl:75-122)
qil_mps* apply_mpo_mps(qil_ctx* ctx, const qil_mpo* W, const qil_mps* psi);
// the same MPO applied to `count` MPS in one launch; outputs share one pooled allocation
void apply_mpo_mps_many(qil_ctx* ctx, const qil_mpo* W, const qil_mps* const* psis, int64_t count, qil_mps** outs);
// K7: MPO o MPO (apply.jl:124-199), W1 acts first, equal lengths or windowed
qil_mpo* apply_mpo_mpo(qil_ctx* ctx, const qil_mpo* W1, const qil_mpo* W2, int start1, int start2);

}  // namespace qil
