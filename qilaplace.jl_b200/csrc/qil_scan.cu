// qil_scan.cu -- read-out reductions that the reference's tutorials do with host loops over `coefficient`:
//   * arg-max of |chi| over a grid / batch of coefficients (docs/src/tutorials/zt.jl:296-411: coarse, fine and
//     superfine pole scans end in `argmax(abs.(chi))`) -- on the device, only 32 bytes come back;
//   * sums over a whole register (docs/src/tutorials/dt.jl:187-197: `laplace_coefficient` adds N coefficients, one per
//     copy-register value j) -- the summed sites are contracted with the all-ones vector and absorbed into their
//     neighbours, which turns N^2 chain evaluations into one dense read-out of an n-site MPS.
#include "qil_mpsops.cuh"

namespace qil {

struct ArgMax {
    double v;
    long long i;
};
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {      // larger value; ties -> lower index
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

template <typename T>
__global__ void __launch_bounds__(256) argmax_abs_kernel(const T* __restrict__ v, long long count, ArgMax* __restrict__ part) {
    __shared__ ArgMax sm[256];
    ArgMax best{-1.0, 0x7fffffffffffffffll};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const ArgMax c{Scalar<T>::abs2(v[i]), i};
        best = better(best, c);
    }
    sm[threadIdx.x] = best;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] = better(sm[threadIdx.x], sm[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sm[0];
}
__global__ void argmax_final_kernel(const ArgMax* __restrict__ part, int n, ArgMax* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        ArgMax best = part[0];
        for (int i = 1; i < n; ++i) best = better(best, part[i]);
        out[0] = best;
    }
}

// index and value of the entry of largest modulus (first one on ties); synchronises
template <typename T>
void argmax_abs(qil_ctx* ctx, const T* d_v, int64_t count, int64_t* index, double* absval, T* value) {
    QIL_REQUIRE(count >= 1, QIL_ERR_ARGUMENT, "argmax: empty collection");
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((count + 255) / 256, (int64_t)ctx->sm_count * 8));
    ArgMax* part = (ArgMax*)ctx->alloc(sizeof(ArgMax) * (grid + 1));
    argmax_abs_kernel<T><<<grid, 256, 0, ctx->stream>>>(d_v, count, part);
    QIL_LAUNCH_CHECK(ctx);
    argmax_final_kernel<<<1, 32, 0, ctx->stream>>>(part, grid, part + grid);
    QIL_LAUNCH_CHECK(ctx);
    ArgMax h;
    QIL_CUDA(cudaMemcpyAsync(&h, part + grid, sizeof(ArgMax), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    T hv;
    QIL_CUDA(cudaMemcpyAsync(&hv, d_v + h.i, sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    ctx->free(part);
    *index = h.i;
    *absval = sqrt(h.v);
    if (value) *value = hv;
}
template void argmax_abs<double>(qil_ctx*, const double*, int64_t, int64_t*, double*, double*);
template void argmax_abs<cplx>(qil_ctx*, const cplx*, int64_t, int64_t*, double*, cplx*);

// S[l][r] = M[l][0][r] + M[l][1][r]
template <typename T>
__global__ void sum_physical_kernel(const T* __restrict__ M, int cl, int cr, T* __restrict__ S) {
    const int total = cl * cr;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int l = idx / cr, r = idx - l * cr;
        S[idx] = Scalar<T>::add(M[((size_t)l * 2 + 0) * cr + r], M[((size_t)l * 2 + 1) * cr + r]);
    }
}

// Contract every site with mask != 0 with the all-ones vector and absorb the resulting bond matrix into a neighbouring
// kept site.  Returns an MPS over the kept sites (same amplitude).
template <typename T>
static qil_mps* sum_sites_t(qil_ctx* ctx, const qil_mps* psi, const uint8_t* mask) {
    const int n = psi->n;
    int kept = 0;
    for (int i = 0; i < n; ++i) kept += mask[i] ? 0 : 1;
    QIL_REQUIRE(kept >= 1, QIL_ERR_ARGUMENT, "sum over sites: at least one site must be kept");
    // walk left to right carrying a pending bond matrix P (pl x pr) that multiplies the next kept core from the left
    std::vector<int64_t> bond;
    std::vector<void*> cores;
    bond.push_back(1);
    Mat<T> P;                       // empty = identity
    int64_t pl = 1;
    for (int i = 0; i < n; ++i) {
        const int64_t cl = psi->bond[i], cr = psi->bond[i + 1];
        const T* M = (const T*)psi->core[i];
        if (mask[i]) {
            Mat<T> S(ctx, cl, cr);
            sum_physical_kernel<T><<<(int)std::min<int64_t>((cl * cr + 255) / 256, 1024), 256, 0, ctx->stream>>>(M, (int)cl, (int)cr, S.p);
            QIL_LAUNCH_CHECK(ctx);
            if (!cores.empty()) {
                // absorb into the previous kept core: C'[l,s,r'] = sum_r C[l,s,r] S[r,r']   (and any pending P first)
                Mat<T> S2;
                const T* Sp = S.p;
                if (P.p) {
                    S2 = Mat<T>(ctx, pl, cr);
                    gemm<T>(ctx, OP_N, OP_N, pl, cr, cl, 1.0, P.p, cl, S.p, cr, 0.0, S2.p, cr);
                    Sp = S2.p;
                    P = Mat<T>();
                }
                const int64_t rows = bond[bond.size() - 2] * 2, k = bond.back();
                T* old = (T*)cores.back();
                T* nw = (T*)ctx->alloc((size_t)rows * cr * sizeof(T));
                gemm<T>(ctx, OP_N, OP_N, rows, cr, k, 1.0, old, k, Sp, cr, 0.0, nw, cr);
                ctx->free(old);
                cores.back() = nw;
                bond.back() = cr;
            } else {
                // nothing kept yet: accumulate into the pending left matrix
                if (P.p) {
                    Mat<T> P2(ctx, pl, cr);
                    gemm<T>(ctx, OP_N, OP_N, pl, cr, cl, 1.0, P.p, cl, S.p, cr, 0.0, P2.p, cr);
                    P = std::move(P2);
                } else {
                    P = std::move(S);
                    pl = cl;
                }
            }
        } else {
            T* nw;
            if (P.p) {
                // C'[l',s,r] = sum_l P[l',l] C[l,s,r]
                nw = (T*)ctx->alloc((size_t)pl * 2 * cr * sizeof(T));
                gemm<T>(ctx, OP_N, OP_N, pl, 2 * cr, cl, 1.0, P.p, cl, M, 2 * cr, 0.0, nw, 2 * cr);
                bond.back() = pl;
                P = Mat<T>();
            } else {
                nw = (T*)ctx->alloc((size_t)cl * 2 * cr * sizeof(T));
                QIL_CUDA(cudaMemcpyAsync(nw, M, (size_t)cl * 2 * cr * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
            }
            cores.push_back(nw);
            bond.push_back(cr);
        }
    }
    qil_mps* m = new_mps(ctx, kept, Scalar<T>::is_complex ? 1 : 0, bond.data(), false);
    m->core = cores;
    m->amplitude = psi->amplitude;
    ctx->sync();
    return m;
}

qil_mps* mps_sum_sites(qil_ctx* ctx, const qil_mps* psi, const uint8_t* mask) {
    return psi->is_complex ? sum_sites_t<cplx>(ctx, psi, mask) : sum_sites_t<double>(ctx, psi, mask);
}

}  // namespace qil
