// qil_svd_small.cu -- fused truncated SVD for bond matrices that fit one CTA's shared memory.
//
// One launch does what qil_svd.cu does in ~8: load (optionally transposed), Householder QR, in-place
// explicit Q, one-sided Jacobi on G = R^H, sort, NDTensors truncation and the requested factors written
// compactly (leading dimension = kept rank).  This is the latency path of the divide-and-conquer encoder's
// lower levels, signal_ztmps, canonicalize!/compress! and the QFT/DT builders, where the matrices are a few
// dozen rows/columns and launch + sync overhead used to dominate.
#include "qil_dense.cuh"
#include "qil_hh.cuh"

namespace qil {

constexpr int kSsThreads = 256;
constexpr int kSsWarps = kSsThreads / 32;

template <typename T>
struct SmallSvdParams {
    const T* A;      // m x n row-major (lda)
    long long lda;
    int m, n;
    int mt, nt;      // tall orientation: mt >= nt; M = A (m >= n) or A^H (m < n)
    int mpad, npad;
    double cutoff;
    long long maxdim, mindim;
    T* U;            // m x r   (ld r) or null
    T* US;           // m x r   or null
    T* Vh;           // r x n   or null
    T* SVh;          // r x n   or null
    double* S;       // min(m,n) values (first r meaningful) or null
    int* rank;
    double* margin;  // closest truncation decision (atomic min), may be null
    int mode;        // 0: A is the matrix; 1: A is an MPS core [cl][2][cr] and the matrix is the copy tensor
    int cl, cr;      //    T[(l,s),(s',r)] = delta(s,s') core[l,s,r]  (signal_ztmps, SignalConverters.jl:263)
};

template <typename T> __device__ __forceinline__ T wsum(T v);
template <> __device__ __forceinline__ double wsum<double>(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <> __device__ __forceinline__ cplx wsum<cplx>(cplx v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}
template <typename T> __device__ __forceinline__ T gsum(T v, int gl, unsigned mask);
template <> __device__ __forceinline__ double gsum<double>(double v, int gl, unsigned mask) {
    for (int o = gl >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
template <> __device__ __forceinline__ cplx gsum<cplx>(cplx v, int gl, unsigned mask) {
    for (int o = gl >> 1; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(mask, v.x, o);
        v.y += __shfl_xor_sync(mask, v.y, o);
    }
    return v;
}

template <typename T>
__device__ __forceinline__ void svd_small_body(const SmallSvdParams<T>& p, unsigned char* smem_raw) {
    const int mt = p.mt, nt = p.nt, mpad = p.mpad, npad = p.npad;
    T* As = reinterpret_cast<T*>(smem_raw);              // [nt][mpad]  reflectors, later Q (column-major)
    T* qb = As + nt * mpad;                      // [kHhIlp * kSsWarps][mpad]
    T* G = qb + kHhIlp * kSsWarps * mpad;        // [nt][npad]  G = R^H, later W = G V
    T* V = G + nt * npad;                        // [nt][npad]
    T* sbeta = V + nt * npad;                    // [nt]
    double* ss = reinterpret_cast<double*>(sbeta + nt);  // [nt]
    double* sig = ss + nt;                               // [nt]
    int* order = reinterpret_cast<int*>(sig + nt);       // [nt]
    __shared__ int s_rot, s_rank;

    const int tid = threadIdx.x, lane = tid & 31;
    const bool tall = p.m >= p.n;

    // ---- load M (tall orientation) column-major: M[i][j] = A[i][j] or conj(A[j][i])
    for (int idx = tid; idx < p.m * p.n; idx += kSsThreads) {
        const int ia = idx / p.n, ja = idx - ia * p.n;
        T v;
        if (p.mode == 0) {
            v = p.A[(long long)ia * p.lda + ja];
        } else {
            const int ll = ia >> 1, s1 = ia & 1, s2 = ja / p.cr, rr = ja - s2 * p.cr;
            v = (s1 == s2) ? p.A[((long long)ll * 2 + s1) * p.cr + rr] : Scalar<T>::zero();
        }
        if (tall) As[ja * mpad + ia] = v;
        else As[ia * mpad + ja] = Scalar<T>::conj(v);
    }
    __syncthreads();

    // ---- Householder factorisation (qil_hh.cuh)
    hh_factor<T>(As, mpad, mt, nt, nt, sbeta, ss);

    // ---- G = R^H (column-major G[col j][row i] = conj(R[j][i])), V = I
    for (int idx = tid; idx < nt * nt; idx += kSsThreads) {
        const int j = idx / nt, i = idx - j * nt;
        T r = Scalar<T>::zero();
        if (i == j) r = sbeta[j];
        else if (i > j) r = As[i * mpad + j];     // R[j][i], stored in column i, row j
        G[j * npad + i] = Scalar<T>::conj(r);
        V[j * npad + i] = (i == j) ? Scalar<T>::one() : Scalar<T>::zero();
    }
    __syncthreads();

    // ---- explicit Q in place (groups from the highest columns down, so reflectors still needed stay intact)
    hh_form_q<T>(As, mpad, mt, nt, ss, qb, [&](int c, const T* q) {
        T* dst = As + c * mpad;
        for (int i = lane; i < mt; i += 32) dst[i] = q[i];
    });

    // ---- one-sided Jacobi on the columns of G (see qil_svd.cu)
    {
        const int ns = nt;
        // widest lane group that still lets every pair of a round run at once (a pass costs one latency chain)
        const int npr = (ns + 1) / 2;
        const int gl = (npr * 32 <= kSsThreads) ? 32 : (npr * 16 <= kSsThreads ? 16 : 8);
        const int grp = tid / gl, gln = tid % gl;
        const unsigned gmask = (gl == 32) ? 0xffffffffu : (((1u << gl) - 1u) << ((tid & 31) / gl * gl));
        const int ngr = kSsThreads / gl;
        const int ne = ns + (ns & 1);
        const int npairs = ne / 2;
        const double tol = sqrt((double)ns) * 2.220446049250313e-16;
        // ||G||_F^2 in a fixed order for the skip threshold (qil_common.cuh: jacobi_skip_threshold)
        __shared__ double s_nu;
        for (int j = tid; j < ns; j += kSsThreads) {
            double a = 0.0;
            for (int i = 0; i < ns; ++i) a += Scalar<T>::abs2(G[j * npad + i]);
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
        for (int sweep = 0; sweep < 60 && ns > 1; ++sweep) {
            if (tid == 0) s_rot = 0;
            __syncthreads();
            for (int r = 0; r < ne - 1; ++r) {
                for (int pi = grp; pi < npairs; pi += ngr) {
                    int a, b;
                    if (pi == 0) { a = ne - 1; b = r; }
                    else { a = (r + pi) % (ne - 1); b = (r - pi + (ne - 1)) % (ne - 1); }
                    if (a >= ns || b >= ns) continue;
                    const int cp = min(a, b), cq = max(a, b);
                    T* gp = G + cp * npad;
                    T* gq = G + cq * npad;
                    double al = 0.0, be = 0.0;
                    T ga = Scalar<T>::zero();
                    for (int i = gln; i < ns; i += gl) {
                        const T x = gp[i], y = gq[i];
                        al += Scalar<T>::abs2(x);
                        be += Scalar<T>::abs2(y);
                        ga = Scalar<T>::fma(Scalar<T>::conj(x), y, ga);
                    }
                    al = gsum<double>(al, gl, gmask);
                    be = gsum<double>(be, gl, gmask);
                    ga = gsum<T>(ga, gl, gmask);
                    const double g2 = Scalar<T>::abs2(ga);
                    if (g2 > tol * tol * al * be && g2 > 0.0 && !(al < nu && be < nu)) {
                        // rsqrt-based rotation (see qil_svd.cu)
                        const double rg = rsqrt(g2);
                        const double ag = g2 * rg;
                        const T ph = Scalar<T>::scale(ga, rg);
                        const double dd = be - al;
                        const double hh = dd * dd + 4.0 * g2;
                        const double sq = hh * rsqrt(hh);
                        const double tt = (dd >= 0.0 ? 2.0 : -2.0) * ag / (fabs(dd) + sq);
                        const double c = rsqrt(1.0 + tt * tt);
                        const double s = c * tt;
                        const T sp = Scalar<T>::scale(ph, s);
                        const T spc = Scalar<T>::conj(sp);
                        T* vp = V + cp * npad;
                        T* vq = V + cq * npad;
                        for (int i = gln; i < ns; i += gl) {
                            const T x = gp[i], y = gq[i];
                            gp[i] = Scalar<T>::sub(Scalar<T>::scale(x, c), Scalar<T>::mul(spc, y));
                            gq[i] = Scalar<T>::add(Scalar<T>::mul(sp, x), Scalar<T>::scale(y, c));
                            const T vx = vp[i], vy = vq[i];
                            vp[i] = Scalar<T>::sub(Scalar<T>::scale(vx, c), Scalar<T>::mul(spc, vy));
                            vq[i] = Scalar<T>::add(Scalar<T>::mul(sp, vx), Scalar<T>::scale(vy, c));
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
    }

    // ---- singular values, order, rank
    for (int j = tid; j < nt; j += kSsThreads) {
        double a = 0.0;
        const T* g = G + j * npad;
        for (int i = 0; i < nt; ++i) a += Scalar<T>::abs2(g[i]);
        sig[j] = sqrt(a);
    }
    __syncthreads();
    double mysig = 0.0;
    int mypos = 0;
    if (tid < nt) {
        mysig = sig[tid];
        for (int i = 0; i < nt; ++i) {
            const double si = sig[i];
            mypos += (si > mysig || (si == mysig && i < tid)) ? 1 : 0;
        }
    }
    __syncthreads();
    if (tid < nt) { order[mypos] = tid; sig[mypos] = mysig; }
    __syncthreads();
    if (tid == 0) {
        s_rank = truncate_rank_dev(sig, nt, p.cutoff, p.maxdim, p.mindim, p.margin);
        *p.rank = s_rank;
    }
    __syncthreads();
    const int r = s_rank;
    if (p.S) for (int j = tid; j < nt; j += kSsThreads) p.S[j] = sig[j];

    // ---- outputs.  M = Q R, G = R^H, G V = W  =>  M = (Q V) W^H
    //   tall (A = M):    U = Q V_r,            S Vh = W_r^H
    //   wide (A = M^H):  U S = W_r,            Vh   = (Q V_r)^H
    // QV[i][j] = sum_c Q[i][c] V[c][j]  with Q = As (column-major), V column-major, j -> order[j]
    const bool need_qv = tall ? (p.U || p.US) : (p.Vh || p.SVh);
    if (need_qv) {
        for (int idx = tid; idx < mt * r; idx += kSsThreads) {
            const int i = idx / r, j = idx - i * r;
            const T* vc = V + order[j] * npad;
            T acc = Scalar<T>::zero();
            for (int c = 0; c < nt; ++c) acc = Scalar<T>::fma(As[c * mpad + i], vc[c], acc);
            const double sj = sig[j];
            if (tall) {
                if (p.U) p.U[(long long)i * r + j] = acc;
                if (p.US) p.US[(long long)i * r + j] = Scalar<T>::scale(acc, sj);
            } else {
                const T cj = Scalar<T>::conj(acc);          // Vh[j][i]
                if (p.Vh) p.Vh[(long long)j * mt + i] = cj;
                if (p.SVh) p.SVh[(long long)j * mt + i] = Scalar<T>::scale(cj, sj);
            }
        }
    }
    const bool need_w = tall ? (p.Vh || p.SVh) : (p.U || p.US);
    if (need_w) {
        for (int idx = tid; idx < nt * r; idx += kSsThreads) {
            const int i = idx / r, j = idx - i * r;         // W[i][j], i < nt
            const T w = G[order[j] * npad + i];
            const double sj = sig[j];
            const double inv = sj != 0.0 ? 1.0 / sj : 0.0;
            if (tall) {
                const T wc = Scalar<T>::conj(w);            // (W^H)[j][i]
                if (p.SVh) p.SVh[(long long)j * nt + i] = wc;
                if (p.Vh) p.Vh[(long long)j * nt + i] = Scalar<T>::scale(wc, inv);
            } else {
                if (p.US) p.US[(long long)i * r + j] = w;
                if (p.U) p.U[(long long)i * r + j] = Scalar<T>::scale(w, inv);
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kSsThreads) svd_small_kernel(const SmallSvdParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    svd_small_body<T>(p, smem_dyn);
}

// one CTA per problem; descriptors live in device memory
template <typename T>
__global__ void __launch_bounds__(kSsThreads) svd_small_batched_kernel(const SmallSvdParams<T>* __restrict__ probs) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    __shared__ SmallSvdParams<T> sp;
    if (threadIdx.x == 0) sp = probs[blockIdx.x];
    __syncthreads();
    svd_small_body<T>(sp, smem_dyn);
}

template <typename T>
static size_t ss_smem(int mt, int nt) {
    const int mpad = mt | 1, npad = nt | 1;
    return ((size_t)(nt + kHhIlp * kSsWarps) * mpad + 2 * (size_t)nt * npad + nt) * sizeof(T) + (size_t)nt * (2 * sizeof(double) + sizeof(int)) + 64;
}

template <typename T>
bool svd_small_fits(qil_ctx* ctx, int64_t m, int64_t n) {
    const int64_t mt = std::max(m, n), nt = std::min(m, n);
    if (nt > 96 || mt > 4096) return false;
    return ss_smem<T>((int)mt, (int)nt) <= std::min<size_t>(ctx->smem_optin, 200 * 1024);
}

// Fused path of svd_trunc for small matrices; same contract (returns the rank after a host sync).
template <typename T>
int svd_small(qil_ctx* ctx, int64_t m, int64_t n, const T* A, int64_t lda, double cutoff, int64_t maxdim,
              int64_t mindim, Mat<T>* U, Mat<T>* US, Mat<T>* Vh, Mat<T>* SVh, Mat<double>* S) {
    const int k = (int)std::min(m, n);
    SmallSvdParams<T> p;
    p.A = A; p.lda = lda; p.m = (int)m; p.n = (int)n;
    p.mt = (int)std::max(m, n); p.nt = k;
    p.mpad = p.mt | 1; p.npad = p.nt | 1;
    p.cutoff = cutoff; p.maxdim = maxdim < 1 ? 1 : maxdim; p.mindim = std::max<int64_t>(mindim, 1);
    p.margin = ctx->d_margin;
    Mat<T> bu, bus, bvh, bsvh;
    Mat<double> bs;
    if (U) bu = Mat<T>(ctx, m, k);
    if (US) bus = Mat<T>(ctx, m, k);
    if (Vh) bvh = Mat<T>(ctx, k, n);
    if (SVh) bsvh = Mat<T>(ctx, k, n);
    if (S) bs = Mat<double>(ctx, k, 1);
    int* d_rank = (int*)ctx->alloc(sizeof(int));
    p.U = U ? bu.p : nullptr; p.US = US ? bus.p : nullptr; p.Vh = Vh ? bvh.p : nullptr;
    p.SVh = SVh ? bsvh.p : nullptr; p.S = S ? bs.p : nullptr; p.rank = d_rank;
    p.mode = 0; p.cl = 0; p.cr = 0;
    const size_t smem = ss_smem<T>(p.mt, p.nt);
    auto kern = svd_small_kernel<T>;
    ensure_dynamic_smem(kern, smem);
    kern<<<1, kSsThreads, smem, ctx->stream>>>(p);
    QIL_LAUNCH_CHECK(ctx);
    int r = 0;
    QIL_CUDA(cudaMemcpyAsync(&r, d_rank, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    ctx->free(d_rank);
    // the kernel wrote compactly with leading dimension r: shrink the logical shapes
    if (U) { bu.cols = r; *U = std::move(bu); }
    if (US) { bus.cols = r; *US = std::move(bus); }
    if (Vh) { bvh.rows = r; *Vh = std::move(bvh); }
    if (SVh) { bsvh.rows = r; *SVh = std::move(bsvh); }
    if (S) { bs.rows = r; *S = std::move(bs); }
    return r;
}

// Batched variant: every item is an independent small SVD (one CTA each, one launch, one host sync).
template <typename T>
void svd_small_batch(qil_ctx* ctx, std::vector<SmallSvdItem<T>>& items, double cutoff, int64_t maxdim, int64_t mindim,
                     std::shared_ptr<void>* pool_out) {
    struct Region {
        qil_ctx* c;
        explicit Region(qil_ctx* cc) : c(cc) { c->prof_begin(PROF_SVD); }
        ~Region() { c->prof_end(); }
    } region(ctx);
    const int nb = (int)items.size();
    if (nb == 0) return;
    std::vector<SmallSvdParams<T>> h(nb);
    int* d_rank = (int*)ctx->alloc(sizeof(int) * nb);
    size_t smem = 0;
    T* pool_base = nullptr;
    size_t pool_used = 0;
    if (pool_out) {
        size_t total = 0;
        for (int i = 0; i < nb; ++i) {
            const SmallSvdItem<T>& it = items[i];
            const int64_t k = std::min(it.m, it.n);
            const size_t mk = (size_t)((it.m * k + 1) & ~(int64_t)1), kn = (size_t)((k * it.n + 1) & ~(int64_t)1);
            total += (it.want_U ? mk : 0) + (it.want_US ? mk : 0) + (it.want_Vh ? kn : 0) + (it.want_SVh ? kn : 0);
        }
        pool_base = (T*)ctx->alloc(std::max<size_t>(total, 1) * sizeof(T));
        *pool_out = std::shared_ptr<void>(pool_base, [ctx](void* q) { ctx->free(q); });
    }
    for (int i = 0; i < nb; ++i) {
        SmallSvdItem<T>& it = items[i];
        const int64_t m = it.m, n = it.n;
        const int k = (int)std::min(m, n);
        SmallSvdParams<T>& p = h[i];
        p.A = it.A; p.lda = it.lda; p.m = (int)m; p.n = (int)n;
        p.mt = (int)std::max(m, n); p.nt = k; p.mpad = p.mt | 1; p.npad = p.nt | 1;
        p.cutoff = cutoff; p.maxdim = maxdim < 1 ? 1 : maxdim; p.mindim = std::max<int64_t>(mindim, 1);
        p.margin = ctx->d_margin;
        if (pool_out) {
            // all outputs of the batch in ONE allocation (the ~2 stream-ordered allocations per item were most of the host
            // time of a ZTMPS split: 56 of them at n = 28); the matrices are non-owning views into it
            auto view = [&](int64_t rr, int64_t cc) {
                Mat<T> v;
                v.ctx = nullptr; v.p = pool_base + pool_used; v.rows = rr; v.cols = cc;
                pool_used += (size_t)((rr * cc + 1) & ~(int64_t)1);
                return v;
            };
            if (it.want_U) it.U = view(m, k);
            if (it.want_US) it.US = view(m, k);
            if (it.want_Vh) it.Vh = view(k, n);
            if (it.want_SVh) it.SVh = view(k, n);
        } else {
            if (it.want_U) it.U = Mat<T>(ctx, m, k);
            if (it.want_US) it.US = Mat<T>(ctx, m, k);
            if (it.want_Vh) it.Vh = Mat<T>(ctx, k, n);
            if (it.want_SVh) it.SVh = Mat<T>(ctx, k, n);
        }
        p.U = it.want_U ? it.U.p : nullptr; p.US = it.want_US ? it.US.p : nullptr;
        p.Vh = it.want_Vh ? it.Vh.p : nullptr; p.SVh = it.want_SVh ? it.SVh.p : nullptr;
        p.S = nullptr; p.rank = d_rank + i;
        p.mode = it.copy_tensor ? 1 : 0; p.cl = it.cl; p.cr = it.cr;
        smem = std::max(smem, ss_smem<T>(p.mt, p.nt));
    }
    SmallSvdParams<T>* d_p = (SmallSvdParams<T>*)ctx->alloc(sizeof(SmallSvdParams<T>) * nb);
    QIL_CUDA(cudaMemcpyAsync(d_p, h.data(), sizeof(SmallSvdParams<T>) * nb, cudaMemcpyHostToDevice, ctx->stream));
    auto kern = svd_small_batched_kernel<T>;
    ensure_dynamic_smem(kern, smem);
    kern<<<nb, kSsThreads, smem, ctx->stream>>>(d_p);
    QIL_LAUNCH_CHECK(ctx);
    std::vector<int> ranks(nb);
    QIL_CUDA(cudaMemcpyAsync(ranks.data(), d_rank, sizeof(int) * nb, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    ctx->free(d_rank);
    ctx->free(d_p);
    for (int i = 0; i < nb; ++i) {
        SmallSvdItem<T>& it = items[i];
        const int r = ranks[i];
        it.rank = r;
        if (it.want_U) it.U.cols = r;
        if (it.want_US) it.US.cols = r;
        if (it.want_Vh) it.Vh.rows = r;
        if (it.want_SVh) it.SVh.rows = r;
    }
}
template void svd_small_batch<double>(qil_ctx*, std::vector<SmallSvdItem<double>>&, double, int64_t, int64_t,
                                      std::shared_ptr<void>*);
template void svd_small_batch<cplx>(qil_ctx*, std::vector<SmallSvdItem<cplx>>&, double, int64_t, int64_t,
                                    std::shared_ptr<void>*);

template bool svd_small_fits<double>(qil_ctx*, int64_t, int64_t);
template bool svd_small_fits<cplx>(qil_ctx*, int64_t, int64_t);
template int svd_small<double>(qil_ctx*, int64_t, int64_t, const double*, int64_t, double, int64_t, int64_t,
                               Mat<double>*, Mat<double>*, Mat<double>*, Mat<double>*, Mat<double>*);
template int svd_small<cplx>(qil_ctx*, int64_t, int64_t, const cplx*, int64_t, double, int64_t, int64_t, Mat<cplx>*,
                             Mat<cplx>*, Mat<cplx>*, Mat<cplx>*, Mat<double>*);

}  // namespace qil
