// qil_mpsops.cu -- MPS algorithms orchestrated on the device toolkit:
//   encode_svd   : sequential TT-SVD            (src/signals/SignalConverters.jl:49-104)
//   ztmps_split  : per-site copy-tensor split   (src/signals/SignalConverters.jl:258-277)
//   canonicalize : one-directional gauge sweep  (src/mps.jl:787-840)
//   compress     : two-site truncated sweeps    (src/mps.jl:913-973)
//   mps_norm     : transfer-matrix chain        (src/mps.jl:754-765)
// Host code only sequences kernels and reads back the data-dependent ranks.
#include "qil_mpsops.cuh"

namespace qil {

// ------------------------------------------------------------------------------------------------
// sum of squares (norm of the signal, SignalConverters.jl:36)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) sumsq_kernel(const T* __restrict__ x, long long n, double* __restrict__ part) {
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s += Scalar<T>::abs2(x[i]);
    __shared__ double sh[256];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

template <typename T>
double device_norm2(qil_ctx* ctx, const T* x, int64_t n) {
    const int grid = (int)std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 8);
    double* d_part = (double*)ctx->alloc(sizeof(double) * grid);
    sumsq_kernel<T><<<grid, 256, 0, ctx->stream>>>(x, n, d_part);
    QIL_LAUNCH_CHECK(ctx);
    std::vector<double> h(grid);
    QIL_CUDA(cudaMemcpyAsync(h.data(), d_part, sizeof(double) * grid, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    ctx->free(d_part);
    // pairwise-ish fixed-order host sum: deterministic
    double s = 0.0;
    for (double v : h) s += v;
    return sqrt(s);
}
template double device_norm2<double>(qil_ctx*, const double*, int64_t);
template double device_norm2<cplx>(qil_ctx*, const cplx*, int64_t);

int ilog2_round(int64_t N) {
    // n = round(Int, log2(N))  (SignalConverters.jl:18)
    return (int)llround(log2((double)N));
}

// ------------------------------------------------------------------------------------------------
// sequential TT-SVD
// ------------------------------------------------------------------------------------------------
template <typename T>
qil_mps* encode_svd(qil_ctx* ctx, const T* d_x, int64_t N, double cutoff, int64_t maxdim) {
    QIL_REQUIRE(N >= 1, QIL_ERR_ARGUMENT, "signal_mps: empty signal");
    const int n = ilog2_round(N);
    QIL_REQUIRE(n >= 1, QIL_ERR_ARGUMENT, "_tensor_to_mps_svd: Need at least one site in the tensor to convert to MPS.");
    const int64_t Np = (int64_t)1 << n;
    QIL_REQUIRE(N <= Np, QIL_ERR_ASSERT, "_array_to_tensor: Length of signal vector must be a power of 2");
    QIL_REQUIRE(n <= kMaxSites, QIL_ERR_UNSUPPORTED, "signal too long");
    // zero-pad (SignalConverters.jl:23-28), normalise (:36-37)
    Mat<T> cur(ctx, 1, Np);
    if (N < Np) QIL_CUDA(cudaMemsetAsync(cur.p, 0, Np * sizeof(T), ctx->stream));
    const double c = device_norm2<T>(ctx, d_x, N);
    scale_copy<T>(ctx, N, 1.0 / c, d_x, cur.p);

    std::vector<int64_t> bond(n + 1, 1);
    std::vector<void*> cores(n, nullptr);
    int64_t chi = 1;
    for (int i = 0; i < n - 1; ++i) {
        const int64_t rows = chi * 2, cols = Np >> (i + 1);
        Mat<T> U, SVh;
        const int r = svd_trunc<T>(ctx, rows, cols, cur.p, cols, cutoff, maxdim, 1, &U, nullptr, nullptr, &SVh, nullptr);
        cores[i] = U.take();
        bond[i + 1] = r;
        cur = std::move(SVh);
        chi = r;
    }
    cores[n - 1] = cur.take();
    qil_mps* m = new_mps(ctx, n, Scalar<T>::is_complex ? 1 : 0, bond.data(), false);
    m->core = cores;
    m->amplitude = c;
    return m;
}
template qil_mps* encode_svd<double>(qil_ctx*, const double*, int64_t, double, int64_t);
template qil_mps* encode_svd<cplx>(qil_ctx*, const cplx*, int64_t, double, int64_t);

// ------------------------------------------------------------------------------------------------
// ZTMPS split: T[(l,s),(s',r)] = delta(s,s') M[l,s,r]; svd -> Amain = U, Acopy = S*V
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void copy_tensor_kernel(const T* __restrict__ M, int l, int r, T* __restrict__ out) {
    const long long total = (long long)l * 2 * 2 * r;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int rr = (int)(idx % r);
        const int s2 = (int)((idx / r) % 2);
        const int s1 = (int)((idx / (2 * r)) % 2);
        const int ll = (int)(idx / (4ll * r));
        out[idx] = (s1 == s2) ? M[((long long)ll * 2 + s1) * r + rr] : Scalar<T>::zero();
    }
}

template <typename T>
static qil_mps* ztmps_split_t(qil_ctx* ctx, const qil_mps* psi, double cutoff, int64_t maxdim) {
    const int n = psi->n;
    QIL_REQUIRE(2 * n <= kMaxSites, QIL_ERR_UNSUPPORTED, "signal_ztmps: too many sites");
    std::vector<int64_t> bond(2 * n + 1, 1);
    std::vector<void*> cores(2 * n, nullptr);
    // sites whose (2 chi_l) x (2 chi_r) copy tensor fits one CTA go into ONE batched launch (the sites are
    // independent); larger bonds take the general path
    std::vector<SmallSvdItem<T>> batch;
    std::vector<int> batch_site;
    for (int i = 0; i < n; ++i) {
        const int l = (int)psi->bond[i], r = (int)psi->bond[i + 1];
        bond[2 * i] = l;
        bond[2 * i + 2] = r;
        if (svd_small_fits<T>(ctx, 2 * l, 2 * r)) {
            SmallSvdItem<T> it;
            it.A = (const T*)psi->core[i];
            it.m = 2 * l; it.n = 2 * r; it.lda = 2 * r;
            it.copy_tensor = true; it.cl = l; it.cr = r;
            it.want_U = true; it.want_SVh = true;
            batch.push_back(std::move(it));
            batch_site.push_back(i);
        } else {
            Mat<T> Tm(ctx, 2 * l, 2 * r);
            const long long total = 4ll * l * r;
            copy_tensor_kernel<T><<<(int)std::min<long long>((total + 255) / 256, 1024), 256, 0, ctx->stream>>>(
                (const T*)psi->core[i], l, r, Tm.p);
            QIL_LAUNCH_CHECK(ctx);
            Mat<T> U, SVh;
            const int c = svd_trunc<T>(ctx, 2 * l, 2 * r, Tm.p, 2 * r, cutoff, maxdim, 1, &U, nullptr, nullptr, &SVh, nullptr);
            bond[2 * i + 1] = c;
            cores[2 * i] = U.take();
            cores[2 * i + 1] = SVh.take();
        }
    }
    // every site in the batch (the usual case): the 2n cores share one pooled allocation
    const bool pooled = (int)batch.size() == n;
    std::shared_ptr<void> pool;
    svd_small_batch<T>(ctx, batch, cutoff, maxdim, 1, pooled ? &pool : nullptr);
    for (size_t b = 0; b < batch.size(); ++b) {
        const int i = batch_site[b];
        bond[2 * i + 1] = batch[b].rank;
        cores[2 * i] = batch[b].U.take();
        cores[2 * i + 1] = batch[b].SVh.take();
    }
    qil_mps* m = new_mps(ctx, 2 * n, psi->is_complex, bond.data(), false);
    m->core = cores;
    m->amplitude = psi->amplitude;
    if (pooled) m->pool = pool;
    return m;
}

qil_mps* ztmps_split(qil_ctx* ctx, const qil_mps* psi, double cutoff, int64_t maxdim) {
    return psi->is_complex ? ztmps_split_t<cplx>(ctx, psi, cutoff, maxdim)
                           : ztmps_split_t<double>(ctx, psi, cutoff, maxdim);
}

// ------------------------------------------------------------------------------------------------
// canonicalize! (in place on the handle)
// ------------------------------------------------------------------------------------------------
template <typename T>
static void canonicalize_t(qil_ctx* ctx, qil_mps* psi, int dir_right, int center, double cutoff, int64_t maxdim) {
    const int N = psi->n;
    if (dir_right) {
        for (int i = 0; i < center - 1; ++i) {
            const int64_t l = psi->bond[i], r = psi->bond[i + 1], r2 = psi->bond[i + 2];
            Mat<T> U, SVh;
            const int k = svd_trunc<T>(ctx, l * 2, r, (const T*)psi->core[i], r, cutoff, maxdim, 1, &U, nullptr, nullptr,
                                       &SVh, nullptr);
            Mat<T> nxt(ctx, k, 2 * r2);
            gemm<T>(ctx, OP_N, OP_N, k, 2 * r2, r, 1.0, SVh.p, r, (const T*)psi->core[i + 1], 2 * r2, 0.0, nxt.p, 2 * r2);
            ctx->free(psi->core[i]);
            ctx->free(psi->core[i + 1]);
            psi->core[i] = U.take();
            psi->core[i + 1] = nxt.take();
            psi->bond[i + 1] = k;
        }
    } else {
        for (int i = N - 1; i >= center; --i) {
            const int64_t l = psi->bond[i], r = psi->bond[i + 1], l0 = psi->bond[i - 1];
            Mat<T> US, Vh;
            const int k = svd_trunc<T>(ctx, l, 2 * r, (const T*)psi->core[i], 2 * r, cutoff, maxdim, 1, nullptr, &US, &Vh,
                                       nullptr, nullptr);
            Mat<T> prv(ctx, l0 * 2, k);
            gemm<T>(ctx, OP_N, OP_N, l0 * 2, k, l, 1.0, (const T*)psi->core[i - 1], l, US.p, k, 0.0, prv.p, k);
            ctx->free(psi->core[i]);
            ctx->free(psi->core[i - 1]);
            psi->core[i] = Vh.take();
            psi->core[i - 1] = prv.take();
            psi->bond[i] = k;
        }
    }
}

void canonicalize(qil_ctx* ctx, qil_mps* psi, int dir_right, int center, double cutoff, int64_t maxdim) {
    unpool(psi);
    const int N = psi->n;
    if (center <= 0) center = dir_right ? N : 1;
    QIL_REQUIRE(center >= 1 && center <= N, QIL_ERR_DOMAIN, "Center out of range [1,%d]", N);
    if (psi->is_complex) canonicalize_t<cplx>(ctx, psi, dir_right, center, cutoff, maxdim);
    else canonicalize_t<double>(ctx, psi, dir_right, center, cutoff, maxdim);
}

// ------------------------------------------------------------------------------------------------
// norm: E <- sum_{a,b,s} E[a,b] M[a,s,c] conj(M[b,s,d])
// ------------------------------------------------------------------------------------------------
template <typename T>
static double mps_norm_t(qil_ctx* ctx, const qil_mps* psi) {
    Mat<T> E(ctx, 1, 1);
    const T one = Scalar<T>::one();
    QIL_CUDA(cudaMemcpyAsync(E.p, &one, sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    for (int i = 0; i < psi->n; ++i) {
        const int64_t l = psi->bond[i], r = psi->bond[i + 1];
        const T* M = (const T*)psi->core[i];
        // X[b][s][c] = sum_a E[a][b] M[a][s][c]
        Mat<T> X(ctx, l, 2 * r);
        gemm<T>(ctx, OP_T, OP_N, l, 2 * r, l, 1.0, E.p, l, M, 2 * r, 0.0, X.p, 2 * r);
        // E'[c][d] = sum_{(b,s)} X[(b,s)][c] conj(M[(b,s)][d])  ==  (M^H X)^T ; compute G = X^T conj(M)
        Mat<T> E2(ctx, r, r);
        // conj(M) as a matrix: use OP_C on M^T trick: E2 = X^T * conj(M) = (M^H X)^T.  Compute H = M^H X (r x r)
        Mat<T> H(ctx, r, r);
        gemm<T>(ctx, OP_C, OP_N, r, r, l * 2, 1.0, M, r, X.p, r, 0.0, H.p, r);
        transpose_conj<T>(ctx, r, r, H.p, r, E2.p, r, false);
        E = std::move(E2);
    }
    T v;
    QIL_CUDA(cudaMemcpyAsync(&v, E.p, sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    return sqrt(sqrt(Scalar<T>::abs2(v)));
}

double mps_norm(qil_ctx* ctx, const qil_mps* psi) {
    return psi->is_complex ? mps_norm_t<cplx>(ctx, psi) : mps_norm_t<double>(ctx, psi);
}

// ------------------------------------------------------------------------------------------------
// compress!
// ------------------------------------------------------------------------------------------------
template <typename T>
static void compress_t(qil_ctx* ctx, qil_mps* psi, int64_t maxdim, double tol, int sweeps) {
    const int N = psi->n;
    const double cutoff = tol * tol / ((double)(N - 1) * sweeps);
    canonicalize(ctx, psi, 0, 0, 1e-12, (int64_t)1 << 62);
    for (int sw = 0; sw < sweeps; ++sw) {
        for (int dir = 0; dir < 2; ++dir) {
            for (int jj = 0; jj < N - 1; ++jj) {
                const int j = dir == 0 ? jj : N - 2 - jj;
                const int64_t l = psi->bond[j], k = psi->bond[j + 1], r = psi->bond[j + 2];
                Mat<T> th(ctx, 2 * l, 2 * r);
                gemm<T>(ctx, OP_N, OP_N, 2 * l, 2 * r, k, 1.0, (const T*)psi->core[j], k, (const T*)psi->core[j + 1],
                        2 * r, 0.0, th.p, 2 * r);
                Mat<T> A, B;
                int kk;
                if (dir == 0)
                    kk = svd_trunc<T>(ctx, 2 * l, 2 * r, th.p, 2 * r, cutoff, maxdim, 1, &A, nullptr, nullptr, &B, nullptr);
                else
                    kk = svd_trunc<T>(ctx, 2 * l, 2 * r, th.p, 2 * r, cutoff, maxdim, 1, nullptr, &A, &B, nullptr, nullptr);
                ctx->free(psi->core[j]);
                ctx->free(psi->core[j + 1]);
                psi->core[j] = A.take();
                psi->core[j + 1] = B.take();
                psi->bond[j + 1] = kk;
            }
        }
    }
    canonicalize(ctx, psi, 0, 0, 1e-12, (int64_t)1 << 62);
    const double nrm = mps_norm(ctx, psi);
    if (nrm != 0.0) {
        psi->amplitude *= nrm;
        scale_copy<T>(ctx, (int64_t)psi->core_elems(0), 1.0 / nrm, (const T*)psi->core[0], (T*)psi->core[0]);
    }
}

void compress(qil_ctx* ctx, qil_mps* psi, int64_t maxdim, double tol, int sweeps) {
    unpool(psi);
    QIL_REQUIRE(psi->n >= 2, QIL_ERR_DOMAIN, "SignalMPS must have at least 2 sites.");
    QIL_REQUIRE(sweeps >= 1, QIL_ERR_ARGUMENT, "compress!: sweeps must be >= 1");
    if (psi->is_complex) compress_t<cplx>(ctx, psi, maxdim, tol, sweeps);
    else compress_t<double>(ctx, psi, maxdim, tol, sweeps);
}


// ------------------------------------------------------------------------------------------------
// Truncating MPO x MPS ("zip-up").  Not in the reference, whose exact apply (apply.jl:75-122) fuses the
// bonds to D*chi -- 200 s / 35 GB on a :random n=16 input (docs/src/benchmarking.md:309).  Here the MPS is brought to
// right-canonical form, then one left-to-right sweep contracts carry x psi_i x W_i and splits it by the same truncated
// SVD (NDTensors rule: relative cumulative cutoff on sigma^2, maxdim) the rest of the path uses:
//   T[b,s,d,c] = sum_{a,p,l} C[b,a,l] psi_i[l,p,c] W_i[a,p,s,d];  T as (b s) x (d c) = U S Vh;  core_i = U,  C <- S Vh.
// ------------------------------------------------------------------------------------------------
template <typename TO, typename TP, typename TW>
static qil_mps* apply_zipup_t(qil_ctx* ctx, const qil_mpo* W, const qil_mps* psi_in, double cutoff, int64_t maxdim) {
    const int n = psi_in->n;
    // right-canonical copy of psi (orthogonality centre on site 1), untruncated (cutoff 0 drops exact zeros only)
    qil_mps* psi = new_mps(ctx, n, psi_in->is_complex, psi_in->bond.data(), true);
    for (int i = 0; i < n; ++i)
        QIL_CUDA(cudaMemcpyAsync(psi->core[i], psi_in->core[i], psi_in->core_elems(i) * elem_size(psi_in->is_complex),
                                 cudaMemcpyDeviceToDevice, ctx->stream));
    psi->amplitude = psi_in->amplitude;
    std::vector<int64_t> ob(n + 1, 1);
    std::vector<void*> cores(n, nullptr);
    try {
        canonicalize(ctx, psi, 0, 1, 0.0, (int64_t)1 << 62);
        Mat<TO> carry(ctx, 1, 1);
        {
            const TO one = Scalar<TO>::one();
            QIL_CUDA(cudaMemcpyAsync(carry.p, &one, sizeof(TO), cudaMemcpyHostToDevice, ctx->stream));
            ctx->sync();
        }
        int64_t b = 1;
        for (int i = 0; i < n; ++i) {
            const int64_t Da = W->bond[i], Dd = W->bond[i + 1], cl = psi->bond[i], cr = psi->bond[i + 1];
            // X[b,a,p,c] = sum_l C[(b,a), l] psi[l, (p,c)]
            Mat<TO> X(ctx, b * Da, 2 * cr);
            {
                ContractDesc d{};
                d.nout = 2; d.ncon = 1;
                d.od[0] = b * Da; d.od[1] = 2 * cr; d.cd[0] = cl;
                d.sa_o[0] = cl; d.sa_o[1] = 0; d.sb_o[0] = 0; d.sb_o[1] = 1; d.sc_o[0] = 2 * cr; d.sc_o[1] = 1;
                d.sa_c[0] = 1; d.sb_c[0] = 2 * cr; d.conj_a = 0;
                contract<TO, TP, TO>(ctx, d, carry.p, (const TP*)psi->core[i], X.p);
            }
            // T[b,s,d,c] = sum_{a,p} X[b,a,p,c] W[a,p,s,d]
            Mat<TO> T(ctx, b * 2, Dd * cr);
            {
                ContractDesc d{};
                d.nout = 4; d.ncon = 2;
                d.od[0] = b; d.od[1] = 2; d.od[2] = Dd; d.od[3] = cr;
                d.cd[0] = Da; d.cd[1] = 2;
                d.sa_o[0] = Da * 2 * cr; d.sa_o[1] = 0; d.sa_o[2] = 0; d.sa_o[3] = 1;
                d.sb_o[0] = 0; d.sb_o[1] = Dd; d.sb_o[2] = 1; d.sb_o[3] = 0;
                d.sc_o[0] = 2 * Dd * cr; d.sc_o[1] = Dd * cr; d.sc_o[2] = cr; d.sc_o[3] = 1;
                d.sa_c[0] = 2 * cr; d.sa_c[1] = cr;
                d.sb_c[0] = 4 * Dd; d.sb_c[1] = 2 * Dd;
                d.conj_a = 0;
                contract<TO, TW, TO>(ctx, d, X.p, (const TW*)W->core[i], T.p);
            }
            if (i == n - 1) {
                QIL_REQUIRE(Dd == 1 && cr == 1, QIL_ERR_ARGUMENT, "apply: boundary bonds must have dimension 1");
                cores[i] = T.take();
                ob[i + 1] = 1;
            } else {
                Mat<TO> U, SVh;
                const int r = svd_trunc<TO>(ctx, b * 2, Dd * cr, T.p, Dd * cr, cutoff, maxdim, 1, &U, nullptr, nullptr, &SVh, nullptr);
                cores[i] = U.take();
                carry = std::move(SVh);
                ob[i + 1] = r;
                b = r;
            }
        }
    } catch (...) {
        for (void* c : cores) ctx->free(c);
        destroy(psi);
        throw;
    }
    qil_mps* out = new_mps(ctx, n, Scalar<TO>::is_complex ? 1 : 0, ob.data(), false);
    out->core = cores;
    out->amplitude = psi_in->amplitude;
    destroy(psi);
    ctx->sync();
    return out;
}

qil_mps* apply_mpo_mps_zipup(qil_ctx* ctx, const qil_mpo* W, const qil_mps* psi, double cutoff, int64_t maxdim) {
    QIL_REQUIRE(W->n == psi->n, QIL_ERR_ARGUMENT,
                "apply: MPO and MPS must have the same number of sites. Found length(W)=%d, length(psi)=%d", W->n, psi->n);
    if (maxdim < 1) maxdim = (int64_t)1 << 62;
    if (W->is_complex && psi->is_complex) return apply_zipup_t<cplx, cplx, cplx>(ctx, W, psi, cutoff, maxdim);
    if (W->is_complex) return apply_zipup_t<cplx, double, cplx>(ctx, W, psi, cutoff, maxdim);
    if (psi->is_complex) return apply_zipup_t<cplx, cplx, double>(ctx, W, psi, cutoff, maxdim);
    return apply_zipup_t<double, double, double>(ctx, W, psi, cutoff, maxdim);
}

}  // namespace qil
