// qil_svd.cu -- K4: truncated SVD = Householder QR + one-sided (Hestenes) Jacobi in shared memory with
// the ITensors/NDTensors truncation rule evaluated on the device.
//
// Replaces every `ITensors.svd(...; cutoff, maxdim[, mindim])` on the path: rsvd.jl:103,
// SignalConverters.jl:84,266, mps.jl:929,946 (+ factorize -> svd at :804,:824), qft_transformer.jl:82,
// dt_transformer.jl:213,261.  One-sided Jacobi on the triangular factor delivers small singular
// values to high relative accuracy, which the cutoffs used here (down to ~1e-25 on sigma^2) need.
//
// With A = Q R (m >= n) and G = R^H, the kernel finds V with G V = W (orthogonal columns):
//   A = (Q V) W^H,  sigma_j = |W_j|,  U = Q V,  S*Vh = W^H      -- no division by sigma anywhere.
#include "qil_dense.cuh"

#include <cstdlib>

namespace qil {

constexpr int kJacThreads = 256;        // shared-memory variant
constexpr int kJacThreadsGlobal = 1024; // single-CTA variant working on an L2-resident scratch copy
constexpr int kMaxSweeps = 60;

template <typename T>
struct JacParams {
    const T* R;      // ns x ns row-major (ld); the kernel works on G = R^H
    long long ld;
    int ns;
    int pad;         // smem column pitch
    T* V;            // out: ns x ns row-major, columns sorted by descending sigma
    T* W;            // out: ns x ns row-major, W = G V (same column order)
    double* S;       // out: ns singular values, descending
    int* rank;       // out: kept rank
    double* margin;  // closest truncation decision (atomic min), may be null
    double cutoff;
    long long maxdim, mindim;
    int gl;          // lanes per column pair (8, 16 or 32)
    T* gscratch;     // if non-null: G and V live here (2 * ns * pad elements) instead of shared memory
};

// reductions inside an aligned group of `gl` lanes; `mask` names exactly that group so that groups of
// one warp may diverge (different pairs, dummy pairs) without deadlocking the shuffle
template <typename T>
__device__ __forceinline__ T group_sum(T v, int gl, unsigned mask);
template <>
__device__ __forceinline__ double group_sum<double>(double v, int gl, unsigned mask) {
    for (int o = gl >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
template <>
__device__ __forceinline__ cplx group_sum<cplx>(cplx v, int gl, unsigned mask) {
    for (int o = gl >> 1; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(mask, v.x, o);
        v.y += __shfl_xor_sync(mask, v.y, o);
    }
    return v;
}

template <typename T, bool GLOBAL>
__global__ void __launch_bounds__(kJacThreadsGlobal) jacobi_kernel(const JacParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ns = p.ns, pad = p.pad;
    const int kJacThreads = blockDim.x;
    T* G;                                         // [ns][pad] column-major
    double* sig;                                  // [ns]
    if (GLOBAL) {
        G = p.gscratch;
        sig = reinterpret_cast<double*>(smem_raw);
    } else {
        G = reinterpret_cast<T*>(smem_raw);
        sig = reinterpret_cast<double*>(G + 2 * ns * pad);
    }
    T* V = G + ns * pad;                          // [ns][pad]
    int* order = reinterpret_cast<int*>(sig + ns);                  // [ns]
    __shared__ int s_rot;

    const int tid = threadIdx.x;
    for (int idx = tid; idx < ns * ns; idx += kJacThreads) {
        const int j = idx / ns, i = idx - j * ns;
        G[j * pad + i] = Scalar<T>::conj(p.R[(long long)j * p.ld + i]);
        V[j * pad + i] = (i == j) ? Scalar<T>::one() : Scalar<T>::zero();
    }
    __syncthreads();

    const int gl = p.gl;
    const int grp = tid / gl, gln = tid % gl;
    const unsigned gmask = (gl == 32) ? 0xffffffffu : (((1u << gl) - 1u) << ((tid & 31) / gl * gl));
    const int ngroups = kJacThreads / gl;
    const int ne = ns + (ns & 1);
    const int npairs = ne / 2;
    const double tol = sqrt((double)ns) * 2.220446049250313e-16;
    // ||G||_F^2 in a fixed summation order (column norms, then thread 0), for the skip threshold
    __shared__ double s_nu;
    for (int j = tid; j < ns; j += kJacThreads) {
        double a = 0.0;
        for (int i = 0; i < ns; ++i) a += Scalar<T>::abs2(G[j * pad + i]);
        sig[j] = a;
    }
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int j = 0; j < ns; ++j) tot += sig[j];
        s_nu = jacobi_skip_threshold(tot, ns, p.cutoff, p.mindim);
    }
    __syncthreads();
    const double nu = s_nu;

    for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
        if (tid == 0) s_rot = 0;
        __syncthreads();
        for (int r = 0; r < ne - 1; ++r) {
            for (int pi = grp; pi < npairs; pi += ngroups) {
                int a, b;
                if (pi == 0) { a = ne - 1; b = r; }
                else { a = (r + pi) % (ne - 1); b = (r - pi + (ne - 1)) % (ne - 1); }
                if (a >= ns || b >= ns) continue;   // group-uniform
                const int cp = min(a, b), cq = max(a, b);
                T* gp = G + cp * pad;
                T* gq = G + cq * pad;
                double al = 0.0, be = 0.0;
                T ga = Scalar<T>::zero();
                for (int i = gln; i < ns; i += gl) {
                    const T x = gp[i], y = gq[i];
                    al += Scalar<T>::abs2(x);
                    be += Scalar<T>::abs2(y);
                    ga = Scalar<T>::fma(Scalar<T>::conj(x), y, ga);
                }
                al = group_sum<double>(al, gl, gmask);
                be = group_sum<double>(be, gl, gmask);
                ga = group_sum<T>(ga, gl, gmask);
                const double g2 = Scalar<T>::abs2(ga);
                if (g2 > tol * tol * al * be && g2 > 0.0 && !(al < nu && be < nu)) {
                    // t = 2|g| sgn(d) / (|d| + sqrt(d^2 + 4|g|^2)), d = beta - alpha; c = rsqrt(1 + t^2), s = c t.
                    // rsqrt-based: the rotation only has to be orthogonal to rounding (c^2 + s^2 = 1), its angle
                    // may carry a few ulps of error (fixed by the next sweep) -- 3 sqrt + 4 div become 3 rsqrt + 1 div.
                    const double rg = rsqrt(g2);
                    const double ag = g2 * rg;
                    const T ph = Scalar<T>::scale(ga, rg);                // e^{i phi}
                    const double dd = be - al;
                    const double hh = dd * dd + 4.0 * g2;
                    const double sq = hh * rsqrt(hh);
                    const double t = (dd >= 0.0 ? 2.0 : -2.0) * ag / (fabs(dd) + sq);
                    const double c = rsqrt(1.0 + t * t);
                    const double s = c * t;
                    const T sp = Scalar<T>::scale(ph, s);                 // s e^{i phi}
                    const T spc = Scalar<T>::conj(sp);                    // s e^{-i phi}
                    for (int i = gln; i < ns; i += gl) {
                        const T x = gp[i], y = gq[i];
                        gp[i] = Scalar<T>::sub(Scalar<T>::scale(x, c), Scalar<T>::mul(spc, y));
                        gq[i] = Scalar<T>::add(Scalar<T>::mul(sp, x), Scalar<T>::scale(y, c));
                    }
                    T* vp = V + cp * pad;
                    T* vq = V + cq * pad;
                    for (int i = gln; i < ns; i += gl) {
                        const T x = vp[i], y = vq[i];
                        vp[i] = Scalar<T>::sub(Scalar<T>::scale(x, c), Scalar<T>::mul(spc, y));
                        vq[i] = Scalar<T>::add(Scalar<T>::mul(sp, x), Scalar<T>::scale(y, c));
                    }
                    if (gln == 0) s_rot = 1;
                }
            }
            __syncthreads();
        }
        const int rot = s_rot;
        __syncthreads();
        if (!rot) break;
    }

    // singular values, descending order by rank sort
    for (int j = tid; j < ns; j += kJacThreads) {
        double a = 0.0;
        const T* g = G + j * pad;
        for (int i = 0; i < ns; ++i) a += Scalar<T>::abs2(g[i]);
        sig[j] = sqrt(a);
    }
    __syncthreads();
    for (int j = tid; j < ns; j += kJacThreads) {
        const double sj = sig[j];
        int pos = 0;
        for (int i = 0; i < ns; ++i) {
            const double si = sig[i];
            pos += (si > sj || (si == sj && i < j)) ? 1 : 0;
        }
        order[pos] = j;
    }
    __syncthreads();
    // sorted sigma (permute through registers), then the sequential truncation rule
    double mine[8];      // ns <= 8 * blockDim.x is checked by the launcher (blockDim.x >= 256)
    int cnt = 0;
    for (int j = tid; j < ns; j += kJacThreads) mine[cnt++] = sig[order[j]];
    __syncthreads();
    cnt = 0;
    for (int j = tid; j < ns; j += kJacThreads) {
        sig[j] = mine[cnt];
        p.S[j] = mine[cnt++];
    }
    __syncthreads();
    if (tid == 0) *p.rank = truncate_rank_dev(sig, ns, p.cutoff, p.maxdim, p.mindim, p.margin);
    // outputs: row-major with sorted columns
    for (int idx = tid; idx < ns * ns; idx += kJacThreads) {
        const int i = idx / ns, j = idx - i * ns;
        const int src = order[j];
        p.V[(long long)i * ns + j] = V[src * pad + i];
        p.W[(long long)i * ns + j] = G[src * pad + i];
    }
}

// ------------------------------------------------------------------------------------------------
// Multi-CTA variant for bond matrices that do not fit one CTA's shared memory (zT builder, large compress!):
// G and V live in an L2-resident scratch; the ceil(ns/2) independent pairs of a round are spread over the lane
// groups of ALL CTAs, rounds are separated by a software grid barrier (the kernel is launched cooperatively, so
// every CTA is resident).  Data written by other CTAs is read with ld.cg (L2), never through L1.
// ------------------------------------------------------------------------------------------------
constexpr int kJacMultiThreads = 256;      // 8 lane groups of 32 per CTA

struct GridSync {
    unsigned int* counter;   // arrivals of the current barrier
    unsigned int* gen;       // barrier generation
    int* rot;                // [3] "some pair was rotated" flags, indexed by sweep % 3
};

__device__ __forceinline__ void grid_barrier(const GridSync& gs) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int g = *((volatile unsigned int*)gs.gen);
        if (atomicAdd(gs.counter, 1u) == gridDim.x - 1) {
            *((volatile unsigned int*)gs.counter) = 0u;
            __threadfence();
            atomicAdd(gs.gen, 1u);
        } else {
            while (*((volatile unsigned int*)gs.gen) == g) { }
        }
        __threadfence();
    }
    __syncthreads();
}

template <typename T> __device__ __forceinline__ T ld_l2(const T* p);
template <> __device__ __forceinline__ double ld_l2<double>(const double* p) { return __ldcg(p); }
template <> __device__ __forceinline__ cplx ld_l2<cplx>(const cplx* p) { return __ldcg(p); }

template <typename T>
__global__ void __launch_bounds__(kJacMultiThreads) jacobi_multi_kernel(const JacParams<T> p, const GridSync gs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ns = p.ns, pad = p.pad;
    T* G = p.gscratch;                            // [ns][pad] column-major, global
    T* V = G + (size_t)ns * pad;
    double* sig = reinterpret_cast<double*>(smem_raw);              // [ns]   (used by CTA 0 at the end)
    int* order = reinterpret_cast<int*>(sig + ns);                  // [ns]

    const int tid = threadIdx.x;
    const long long gtid = (long long)blockIdx.x * blockDim.x + tid;
    const long long gthreads = (long long)gridDim.x * blockDim.x;
    for (long long idx = gtid; idx < (long long)ns * ns; idx += gthreads) {
        const int j = (int)(idx / ns), i = (int)(idx - (long long)j * ns);
        G[(size_t)j * pad + i] = Scalar<T>::conj(p.R[(long long)j * p.ld + i]);
        V[(size_t)j * pad + i] = (i == j) ? Scalar<T>::one() : Scalar<T>::zero();
    }
    if (gtid < 3) gs.rot[gtid] = 0;
    grid_barrier(gs);

    constexpr int gl = 32;
    const int grp = (int)(gtid / gl), gln = tid % gl;
    const int ngroups = (int)(gthreads / gl);
    const int ne = ns + (ns & 1);
    const int npairs = ne / 2;
    const double tol = sqrt((double)ns) * 2.220446049250313e-16;
    // ||G||_F^2, computed redundantly by every CTA in the same fixed order (skip threshold, see qil_common.cuh)
    __shared__ double s_nu;
    for (int j = tid; j < ns; j += kJacMultiThreads) {
        double a = 0.0;
        const T* g = G + (size_t)j * pad;
        for (int i = 0; i < ns; ++i) a += Scalar<T>::abs2(ld_l2<T>(g + i));
        sig[j] = a;
    }
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int j = 0; j < ns; ++j) tot += sig[j];
        s_nu = jacobi_skip_threshold(tot, ns, p.cutoff, p.mindim);
    }
    __syncthreads();
    const double nu = s_nu;

    for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
        int* rot = gs.rot + sweep % 3;
        if (gtid == 0) gs.rot[(sweep + 1) % 3] = 0;          // the flag of the NEXT sweep (nobody touches it now)
        for (int r = 0; r < ne - 1; ++r) {
            for (int pi = grp; pi < npairs; pi += ngroups) {
                int a, b;
                if (pi == 0) { a = ne - 1; b = r; }
                else { a = (r + pi) % (ne - 1); b = (r - pi + (ne - 1)) % (ne - 1); }
                if (a >= ns || b >= ns) continue;   // warp-uniform (one pair per warp)
                const int cp = min(a, b), cq = max(a, b);
                T* gp = G + (size_t)cp * pad;
                T* gq = G + (size_t)cq * pad;
                double al = 0.0, be = 0.0;
                T ga = Scalar<T>::zero();
                for (int i = gln; i < ns; i += gl) {
                    const T x = ld_l2<T>(gp + i), y = ld_l2<T>(gq + i);
                    al += Scalar<T>::abs2(x);
                    be += Scalar<T>::abs2(y);
                    ga = Scalar<T>::fma(Scalar<T>::conj(x), y, ga);
                }
                al = group_sum<double>(al, gl, 0xffffffffu);
                be = group_sum<double>(be, gl, 0xffffffffu);
                ga = group_sum<T>(ga, gl, 0xffffffffu);
                const double g2 = Scalar<T>::abs2(ga);
                if (g2 > tol * tol * al * be && g2 > 0.0 && !(al < nu && be < nu)) {
                    const double rg = rsqrt(g2);
                    const double ag = g2 * rg;
                    const T ph = Scalar<T>::scale(ga, rg);
                    const double dd = be - al;
                    const double hh = dd * dd + 4.0 * g2;
                    const double sq = hh * rsqrt(hh);
                    const double t = (dd >= 0.0 ? 2.0 : -2.0) * ag / (fabs(dd) + sq);
                    const double c = rsqrt(1.0 + t * t);
                    const double sn = c * t;
                    const T sp = Scalar<T>::scale(ph, sn);
                    const T spc = Scalar<T>::conj(sp);
                    for (int i = gln; i < ns; i += gl) {
                        const T x = ld_l2<T>(gp + i), y = ld_l2<T>(gq + i);
                        gp[i] = Scalar<T>::sub(Scalar<T>::scale(x, c), Scalar<T>::mul(spc, y));
                        gq[i] = Scalar<T>::add(Scalar<T>::mul(sp, x), Scalar<T>::scale(y, c));
                    }
                    T* vp = V + (size_t)cp * pad;
                    T* vq = V + (size_t)cq * pad;
                    for (int i = gln; i < ns; i += gl) {
                        const T x = ld_l2<T>(vp + i), y = ld_l2<T>(vq + i);
                        vp[i] = Scalar<T>::sub(Scalar<T>::scale(x, c), Scalar<T>::mul(spc, y));
                        vq[i] = Scalar<T>::add(Scalar<T>::mul(sp, x), Scalar<T>::scale(y, c));
                    }
                    if (gln == 0) *((volatile int*)rot) = 1;
                }
            }
            grid_barrier(gs);
        }
        if (*((volatile int*)rot) == 0) break;      // same value in every CTA: written before the last barrier
    }
    if (blockIdx.x != 0) return;

    // ---- CTA 0: singular values, order, rank, outputs (all reads through L2)
    for (int j = tid; j < ns; j += kJacMultiThreads) {
        double a = 0.0;
        const T* g = G + (size_t)j * pad;
        for (int i = 0; i < ns; ++i) a += Scalar<T>::abs2(ld_l2<T>(g + i));
        sig[j] = sqrt(a);
    }
    __syncthreads();
    for (int j = tid; j < ns; j += kJacMultiThreads) {
        const double sj = sig[j];
        int pos = 0;
        for (int i = 0; i < ns; ++i) {
            const double si = sig[i];
            pos += (si > sj || (si == sj && i < j)) ? 1 : 0;
        }
        order[pos] = j;
    }
    __syncthreads();
    double mine[8];
    int cnt = 0;
    for (int j = tid; j < ns; j += kJacMultiThreads) mine[cnt++] = sig[order[j]];
    __syncthreads();
    cnt = 0;
    for (int j = tid; j < ns; j += kJacMultiThreads) {
        sig[j] = mine[cnt];
        p.S[j] = mine[cnt++];
    }
    __syncthreads();
    if (tid == 0) *p.rank = truncate_rank_dev(sig, ns, p.cutoff, p.maxdim, p.mindim, p.margin);
    for (int idx = tid; idx < ns * ns; idx += kJacMultiThreads) {
        const int i = idx / ns, j = idx - i * ns;
        const int src = order[j];
        p.V[(long long)i * ns + j] = ld_l2<T>(V + (size_t)src * pad + i);
        p.W[(long long)i * ns + j] = ld_l2<T>(G + (size_t)src * pad + i);
    }
}

template <typename T>
__global__ void sum_partials_kernel(long long count, int nsum, long long stride, const T* __restrict__ in,
                                    T* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (long long)gridDim.x * blockDim.x) {
        T v = in[i];
        for (int s = 1; s < nsum; ++s) v = Scalar<T>::add(v, in[i + (long long)s * stride]);
        out[i] = v;
    }
}

template <typename T>
static size_t jac_smem(int ns) {
    const int pad = ns | 1;
    return (size_t)2 * ns * pad * sizeof(T) + (size_t)ns * (sizeof(double) + sizeof(int)) + 64;
}

// Jacobi on G = R^H for a square ns x ns R; returns V, W (ns x ns, sorted columns), S and the rank.
template <typename T>
static int jacobi_square(qil_ctx* ctx, int ns, const T* R, int64_t ld, double cutoff, int64_t maxdim, int64_t mindim,
                         Mat<T>& V, Mat<T>& W, Mat<double>& S) {
    QIL_REQUIRE(ns <= 8 * kJacThreads, QIL_ERR_UNSUPPORTED, "svd: %d columns exceed the Jacobi kernel", ns);
    size_t smem = jac_smem<T>(ns);
    const size_t budget = std::min<size_t>(ctx->smem_optin, 225 * 1024);
    const bool use_global = smem > budget;
    Mat<T> gscratch;
    if (use_global) {
        gscratch = Mat<T>(ctx, 2 * (int64_t)ns, ns | 1);
        smem = (size_t)ns * (sizeof(double) + sizeof(int)) + 64;
    }
    V = Mat<T>(ctx, ns, ns);
    W = Mat<T>(ctx, ns, ns);
    S = Mat<double>(ctx, ns, 1);
    int* d_rank = (int*)ctx->alloc(sizeof(int));
    JacParams<T> p;
    p.R = R; p.ld = ld; p.ns = ns; p.pad = ns | 1; p.V = V.p; p.W = W.p; p.S = S.p; p.rank = d_rank;
    p.cutoff = cutoff; p.maxdim = maxdim; p.mindim = std::max<int64_t>(mindim, 1);
    p.margin = ctx->d_margin;
    // lanes per column pair; one round of the round-robin ordering has ceil(ns/2) independent pairs, and every
    // pass over them costs about the same dependent-latency chain, so the CTA gets enough threads for all pairs of
    // a round at once whenever 1024 threads allow it
    p.gscratch = use_global ? gscratch.p : nullptr;
    const int npairs = (ns + 1) / 2;
    const int threads = use_global ? kJacThreadsGlobal
                                   : std::max(kJacThreads, std::min(((npairs * 8 + 31) / 32) * 32, kJacThreadsGlobal));
    p.gl = use_global ? 32 : ((npairs * 32 <= threads) ? 32 : (npairs * 16 <= threads ? 16 : 8));   // L2-resident: keep loads wide
    if (use_global) {
        // multi-CTA kernel (cooperative launch: every CTA must be resident for the software grid barrier)
        const int want = std::min((npairs + 7) / 8, ctx->sm_count);
        auto mk = jacobi_multi_kernel<T>;
        ensure_dynamic_smem(mk, smem);
        int occ = 0;
        QIL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mk, kJacMultiThreads, smem));
        static const bool single = [] { const char* e = getenv("QIL_JACOBI_SINGLE_CTA"); return e && e[0] == '1'; }();
        if (occ >= 1 && want >= 2 && !single) {
            unsigned int* sync_words = (unsigned int*)ctx->alloc(8 * sizeof(unsigned int));
            QIL_CUDA(cudaMemsetAsync(sync_words, 0, 8 * sizeof(unsigned int), ctx->stream));
            GridSync gs{sync_words, sync_words + 1, reinterpret_cast<int*>(sync_words + 2)};
            void* args[] = {(void*)&p, (void*)&gs};
            QIL_CUDA(cudaLaunchCooperativeKernel((const void*)mk, dim3(want), dim3(kJacMultiThreads), args, smem,
                                                 ctx->stream));
            QIL_LAUNCH_CHECK(ctx);
            int rank = 0;
            QIL_CUDA(cudaMemcpyAsync(&rank, d_rank, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            ctx->sync();
            ctx->free(d_rank);
            ctx->free(sync_words);
            return rank;
        }
        auto kern = jacobi_kernel<T, true>;
        ensure_dynamic_smem(kern, smem);
        kern<<<1, kJacThreadsGlobal, smem, ctx->stream>>>(p);
    } else {
        auto kern = jacobi_kernel<T, false>;
        ensure_dynamic_smem(kern, smem);
        kern<<<1, threads, smem, ctx->stream>>>(p);
    }
    QIL_LAUNCH_CHECK(ctx);
    int rank = 0;
    QIL_CUDA(cudaMemcpyAsync(&rank, d_rank, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    ctx->free(d_rank);
    return rank;
}

// adj != nullptr: the caller already holds A^H (n x m row-major, pitch ldadj) for a wide A (m <= n)
template <typename T>
static int svd_core(qil_ctx* ctx, int64_t m, int64_t n, const T* A, int64_t lda, const T* adj, int64_t ldadj,
                    double cutoff, int64_t maxdim, int64_t mindim, Mat<T>* U, Mat<T>* US, Mat<T>* Vh, Mat<T>* SVh,
                    Mat<double>* S, int nsum, int64_t sum_stride) {
    QIL_REQUIRE(m >= 1 && n >= 1, QIL_ERR_ARGUMENT, "svd: empty matrix");
    struct Region {   // profiler class 4 (contains its QR)
        qil_ctx* c;
        explicit Region(qil_ctx* cc) : c(cc) { c->prof_begin(PROF_SVD); }
        ~Region() { c->prof_end(); }
    } region(ctx);
    if (maxdim < 1) maxdim = 1;
    if (adj == nullptr && nsum == 1 && svd_small_fits<T>(ctx, m, n))
        return svd_small<T>(ctx, m, n, A, lda, cutoff, maxdim, mindim, U, US, Vh, SVh, S);
    Mat<T> Q, R, V, W;
    Mat<double> Sv;
    const bool tall = (adj == nullptr) && m >= n;
    const int ns = (int)std::min(m, n);
    bool have_q = true;
    Mat<T> At;
    if (tall) {
        if (m == n && nsum == 1) {
            have_q = false;
        } else {
            qr_thin<T>(ctx, m, n, A, lda, false, Q, R, nsum, sum_stride, true);
        }
    } else if (adj) {
        QIL_REQUIRE(m <= n, QIL_ERR_ARGUMENT, "svd: adjoint input requires a wide matrix");
        qr_thin<T>(ctx, n, m, adj, ldadj, false, Q, R, 1, 0, true);
    } else {
        Mat<T> Asum;
        const T* src = A;
        int64_t ld = lda;
        if (nsum > 1) {
            QIL_REQUIRE(lda == n, QIL_ERR_ARGUMENT, "svd: partial sums need a dense leading dimension");
            Asum = Mat<T>(ctx, m, n);
            int grid = (int)std::min<long long>((m * n + 255) / 256, (long long)ctx->sm_count * 16);
            sum_partials_kernel<T><<<grid, 256, 0, ctx->stream>>>(m * n, nsum, sum_stride, A, Asum.p);
            QIL_LAUNCH_CHECK(ctx);
            src = Asum.p;
            ld = n;
        }
        At = Mat<T>(ctx, n, m);
        transpose_conj<T>(ctx, m, n, src, ld, At.p, m);
        qr_thin<T>(ctx, n, m, At.p, m, false, Q, R, 1, 0, true);
    }
    const T* Rp = have_q ? R.p : A;
    const int64_t ldr = have_q ? ns : lda;
    const int r = jacobi_square<T>(ctx, ns, Rp, ldr, cutoff, maxdim, mindim, V, W, Sv);

    if (tall) {
        // U = Q V[:, :r] ; SVh = W[:, :r]^H
        Mat<T> Uloc;
        if (U || US) {
            Uloc = Mat<T>(ctx, m, r);
            if (have_q) gemm<T>(ctx, OP_N, OP_N, m, r, ns, 1.0, Q.p, ns, V.p, ns, 0.0, Uloc.p, r);
            else QIL_CUDA(cudaMemcpy2DAsync(Uloc.p, r * sizeof(T), V.p, ns * sizeof(T), r * sizeof(T), m,
                                            cudaMemcpyDeviceToDevice, ctx->stream));
        }
        if (US) {
            *US = Mat<T>(ctx, m, r);
            scale_rows_cols<T>(ctx, m, r, Uloc.p, r, Sv.p, false, false, US->p, r);
        }
        if (U) *U = std::move(Uloc);
        if (SVh || Vh) {
            Mat<T> sv(ctx, r, ns);
            transpose_conj<T>(ctx, ns, r, W.p, ns, sv.p, ns);
            if (Vh) {
                *Vh = Mat<T>(ctx, r, ns);
                scale_rows_cols<T>(ctx, r, ns, sv.p, ns, Sv.p, true, true, Vh->p, ns);
            }
            if (SVh) *SVh = std::move(sv);
        }
    } else {
        // A = W (Q V)^H : US = W[:, :r], Vh = (Q V[:, :r])^H
        if (US || U) {
            Mat<T> us(ctx, m, r);
            QIL_CUDA(cudaMemcpy2DAsync(us.p, r * sizeof(T), W.p, ns * sizeof(T), r * sizeof(T), m,
                                       cudaMemcpyDeviceToDevice, ctx->stream));
            if (U) {
                *U = Mat<T>(ctx, m, r);
                scale_rows_cols<T>(ctx, m, r, us.p, r, Sv.p, false, true, U->p, r);
            }
            if (US) *US = std::move(us);
        }
        if (Vh || SVh) {
            Mat<T> qv(ctx, n, r);
            gemm<T>(ctx, OP_N, OP_N, n, r, ns, 1.0, Q.p, ns, V.p, ns, 0.0, qv.p, r);
            Mat<T> vh(ctx, r, n);
            transpose_conj<T>(ctx, n, r, qv.p, r, vh.p, n);
            if (SVh) {
                *SVh = Mat<T>(ctx, r, n);
                scale_rows_cols<T>(ctx, r, n, vh.p, n, Sv.p, true, false, SVh->p, n);
            }
            if (Vh) *Vh = std::move(vh);
        }
    }
    if (S) {
        *S = Mat<double>(ctx, r, 1);
        QIL_CUDA(cudaMemcpyAsync(S->p, Sv.p, r * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return r;
}

template <typename T>
int svd_trunc(qil_ctx* ctx, int64_t m, int64_t n, const T* A, int64_t lda, double cutoff, int64_t maxdim,
              int64_t mindim, Mat<T>* U, Mat<T>* US, Mat<T>* Vh, Mat<T>* SVh, Mat<double>* S, int nsum,
              int64_t sum_stride) {
    return svd_core<T>(ctx, m, n, A, lda, nullptr, 0, cutoff, maxdim, mindim, U, US, Vh, SVh, S, nsum, sum_stride);
}

template <typename T>
int svd_trunc_adj(qil_ctx* ctx, int64_t m, int64_t n, const T* At, int64_t ldat, double cutoff, int64_t maxdim,
                  int64_t mindim, Mat<T>* U, Mat<T>* US, Mat<T>* Vh, Mat<T>* SVh, Mat<double>* S) {
    return svd_core<T>(ctx, m, n, nullptr, 0, At, ldat, cutoff, maxdim, mindim, U, US, Vh, SVh, S, 1, 0);
}
template int svd_trunc_adj<double>(qil_ctx*, int64_t, int64_t, const double*, int64_t, double, int64_t, int64_t,
                                   Mat<double>*, Mat<double>*, Mat<double>*, Mat<double>*, Mat<double>*);
template int svd_trunc_adj<cplx>(qil_ctx*, int64_t, int64_t, const cplx*, int64_t, double, int64_t, int64_t,
                                 Mat<cplx>*, Mat<cplx>*, Mat<cplx>*, Mat<cplx>*, Mat<double>*);

template int svd_trunc<double>(qil_ctx*, int64_t, int64_t, const double*, int64_t, double, int64_t, int64_t,
                               Mat<double>*, Mat<double>*, Mat<double>*, Mat<double>*, Mat<double>*, int, int64_t);
template int svd_trunc<cplx>(qil_ctx*, int64_t, int64_t, const cplx*, int64_t, double, int64_t, int64_t, Mat<cplx>*,
                             Mat<cplx>*, Mat<cplx>*, Mat<cplx>*, Mat<double>*, int, int64_t);

}  // namespace qil
