// qil_apply.cu -- K6 / K7: exact MPO x MPS and MPO o MPO (reference: src/linalg/apply.jl:75-199).
//
// out[(l,a), s, (r,b)] = sum_p W[a,p,s,b] * psi[l,p,r]      (MPO bond fastest in the fused bonds,
// `combiner(W.bonds[i], psi.bonds[i])`, apply.jl:105-119).  Contraction depth is 2, so this is a
// write-bound Kronecker-style expansion: all sites go in ONE launch, one thread per pair of output
// elements (s = 0,1), consecutive threads along the fused right bond for coalesced stores.
#include "qil_common.cuh"

namespace qil {

struct ApplyDesc {
    int n;
    int wb[kMaxSites + 1];   // MPO bonds
    int pb[kMaxSites + 1];   // MPS (or second MPO) bonds
    const void* w[kMaxSites];
    const void* p[kMaxSites];
    void* o[kMaxSites];
};

template <typename TW, typename TP, typename TO>
__global__ void __launch_bounds__(256) apply_mpo_mps_kernel(const ApplyDesc d) {
    const int i = blockIdx.y;
    const int Da = d.wb[i], Db = d.wb[i + 1], cl = d.pb[i], cr = d.pb[i + 1];
    const TW* __restrict__ W = reinterpret_cast<const TW*>(d.w[i]);
    const TP* __restrict__ P = reinterpret_cast<const TP*>(d.p[i]);
    TO* __restrict__ O = reinterpret_cast<TO*>(d.o[i]);
    const long long R = (long long)cr * Db;       // fused right bond
    const long long L = (long long)cl * Da;       // fused left bond
    const long long total = L * R;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long LA = idx / R, RB = idx - LA * R;
        const int l = (int)(LA / Da), a = (int)(LA - (long long)l * Da);
        const int r = (int)(RB / Db), b = (int)(RB - (long long)r * Db);
        const TO p0 = promote<TO, TP>(P[((size_t)l * 2 + 0) * cr + r]);
        const TO p1 = promote<TO, TP>(P[((size_t)l * 2 + 1) * cr + r]);
        // W[a][p][s][b]
        const size_t wbase = (size_t)a * 4 * Db + b;
        const TO w00 = promote<TO, TW>(W[wbase + 0 * Db]);  // p=0,s=0
        const TO w01 = promote<TO, TW>(W[wbase + 1 * Db]);  // p=0,s=1
        const TO w10 = promote<TO, TW>(W[wbase + 2 * Db]);  // p=1,s=0
        const TO w11 = promote<TO, TW>(W[wbase + 3 * Db]);  // p=1,s=1
        TO o0 = Scalar<TO>::mul(w00, p0);
        o0 = Scalar<TO>::fma(w10, p1, o0);
        TO o1 = Scalar<TO>::mul(w01, p0);
        o1 = Scalar<TO>::fma(w11, p1, o1);
        O[((size_t)LA * 2 + 0) * R + RB] = o0;
        O[((size_t)LA * 2 + 1) * R + RB] = o1;
    }
}

// out[(c,a), p, s, (d,b)] = sum_m W1[a,p,m,b] * W2[c,m,s,d]   (W1 acts first; W1 bond fastest)
template <typename T1, typename T2, typename TO>
__global__ void __launch_bounds__(256) apply_mpo_mpo_kernel(const ApplyDesc d) {
    const int i = blockIdx.y;
    const int Da = d.wb[i], Db = d.wb[i + 1], Dc = d.pb[i], Dd = d.pb[i + 1];
    const T1* __restrict__ W1 = reinterpret_cast<const T1*>(d.w[i]);
    const T2* __restrict__ W2 = reinterpret_cast<const T2*>(d.p[i]);
    TO* __restrict__ O = reinterpret_cast<TO*>(d.o[i]);
    const long long R = (long long)Dd * Db;
    const long long L = (long long)Dc * Da;
    const long long total = L * R;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long CA = idx / R, DB = idx - CA * R;
        const int c = (int)(CA / Da), a = (int)(CA - (long long)c * Da);
        const int dd = (int)(DB / Db), b = (int)(DB - (long long)dd * Db);
        TO w1[2][2], w2[2][2];
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
            for (int y = 0; y < 2; ++y) {
                w1[x][y] = promote<TO, T1>(W1[((size_t)a * 4 + x * 2 + y) * Db + b]);
                w2[x][y] = promote<TO, T2>(W2[((size_t)c * 4 + x * 2 + y) * Dd + dd]);
            }
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                TO v = Scalar<TO>::mul(w1[p][0], w2[0][s]);
                v = Scalar<TO>::fma(w1[p][1], w2[1][s], v);
                O[((size_t)CA * 4 + p * 2 + s) * R + DB] = v;
            }
    }
}

template <typename K>
static void launch_sites(qil_ctx* ctx, K kern, const ApplyDesc& d, long long max_total) {
    int bx = (int)std::min<long long>((max_total + 255) / 256, (long long)ctx->sm_count * 8);
    if (bx < 1) bx = 1;
    dim3 grid(bx, d.n);
    { qil_prof_region prof_guard_(ctx, PROF_APPLY);
    kern<<<grid, 256, 0, ctx->stream>>>(d);
    QIL_LAUNCH_CHECK(ctx);
    }
}

// ---- MPO x MPS for one or many MPS in ONE launch -------------------------------------------------------------
// A job is a block of fused-left-bond rows of one site of one MPS.  2-D indexing: the (l, a) split is done once per row,
// the (r, b) split with 32-bit arithmetic per element; consecutive threads write consecutive elements of a row
// (16-byte stores for complex).  All output cores of the call live in one pooled allocation (one cudaMallocAsync
// instead of one per site), shared by the returned handles.
struct ApplyJob {
    const void* w;
    const void* p;
    void* o;
    int Da, Db, cl, cr;
    int row0, rows;
};

template <typename TW, typename TP, typename TO>
__global__ void __launch_bounds__(256) apply_jobs_kernel(const ApplyJob* __restrict__ jobs) {
    const ApplyJob j = jobs[blockIdx.x];
    const TW* __restrict__ W = reinterpret_cast<const TW*>(j.w);
    const TP* __restrict__ P = reinterpret_cast<const TP*>(j.p);
    TO* __restrict__ O = reinterpret_cast<TO*>(j.o);
    const unsigned Da = j.Da, Db = j.Db, cr = j.cr;
    const unsigned R = cr * Db;
    // 32 x 8 threads: x along the fused right bond (coalesced stores), y over the rows of the job
    const unsigned tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (unsigned row = ty; row < (unsigned)j.rows; row += 8) {
        const unsigned LA = j.row0 + row;
        const unsigned l = LA / Da, a = LA - l * Da;
        const TW* __restrict__ wa = W + (size_t)a * 4 * Db;
        const TP* __restrict__ p0r = P + (size_t)l * 2 * cr;
        TO* __restrict__ o0 = O + (size_t)LA * 2 * R;
        // four independent elements per thread and iteration: the kernel is latency bound on the operand loads
        for (unsigned RB0 = tx; RB0 < R; RB0 += 128) {
            TO v0[4], v1[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const unsigned RB = min(RB0 + 32 * e, R - 1);
                const unsigned r = RB / Db, b = RB - r * Db;
                const TO p0 = promote<TO, TP>(p0r[r]);
                const TO p1 = promote<TO, TP>(p0r[cr + r]);
                const TO w00 = promote<TO, TW>(wa[b]);            // p=0,s=0
                const TO w01 = promote<TO, TW>(wa[Db + b]);       // p=0,s=1
                const TO w10 = promote<TO, TW>(wa[2 * Db + b]);   // p=1,s=0
                const TO w11 = promote<TO, TW>(wa[3 * Db + b]);   // p=1,s=1
                v0[e] = Scalar<TO>::fma(w10, p1, Scalar<TO>::mul(w00, p0));
                v1[e] = Scalar<TO>::fma(w11, p1, Scalar<TO>::mul(w01, p0));
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const unsigned RB = RB0 + 32 * e;
                if (RB < R) {
                    o0[RB] = v0[e];
                    o0[R + RB] = v1[e];
                }
            }
        }
    }
}

void apply_mpo_mps_many(qil_ctx* ctx, const qil_mpo* W, const qil_mps* const* psis, int64_t count, qil_mps** outs) {
    if (count <= 0) return;
    const int n = W->n;
    const int pc = psis[0]->is_complex;
    for (int64_t s = 0; s < count; ++s) {
        QIL_REQUIRE(psis[s] != nullptr, QIL_ERR_ARGUMENT, "apply: null MPS handle in the batch");
        QIL_REQUIRE(W->n == psis[s]->n, QIL_ERR_ARGUMENT,
                    "apply: MPO and MPS must have the same number of sites. Found length(W)=%d, length(psi)=%d", W->n,
                    psis[s]->n);
        QIL_REQUIRE(psis[s]->is_complex == pc, QIL_ERR_ARGUMENT, "apply: the MPS of a batch must share their element type");
    }
    const int oc = (W->is_complex || pc) ? 1 : 0;
    const size_t es = elem_size(oc);
    // output layout
    std::vector<std::vector<int64_t>> ob(count, std::vector<int64_t>(n + 1));
    size_t total = 0;
    std::vector<size_t> core_off((size_t)count * n);
    for (int64_t s = 0; s < count; ++s) {
        for (int i = 0; i <= n; ++i) {
            ob[s][i] = W->bond[i] * psis[s]->bond[i];
            QIL_REQUIRE(ob[s][i] < ((int64_t)1 << 28), QIL_ERR_UNSUPPORTED, "apply: fused bond too large");
        }
        for (int i = 0; i < n; ++i) {
            core_off[s * n + i] = total;
            total += ((size_t)ob[s][i] * 2 * ob[s][i + 1] + 1) & ~(size_t)1;
        }
    }
    void* pool_raw = ctx->alloc(std::max<size_t>(total, 1) * es);
    std::shared_ptr<void> pool(pool_raw, [ctx](void* p) { ctx->free(p); });
    // jobs: ~128 KB of output each (at least 8 rows: one per thread row)
    std::vector<ApplyJob> jobs;
    for (int64_t s = 0; s < count; ++s)
        for (int i = 0; i < n; ++i) {
            const int64_t L = ob[s][i], R = ob[s][i + 1];
            int64_t rows_per = std::max<int64_t>(8, (int64_t)(131072 / es) / std::max<int64_t>(2 * R, 1));
            rows_per = std::min(rows_per, L);
            for (int64_t r0 = 0; r0 < L; r0 += rows_per) {
                ApplyJob j;
                j.w = W->core[i]; j.p = psis[s]->core[i];
                j.o = (char*)pool_raw + core_off[s * n + i] * es;
                j.Da = (int)W->bond[i]; j.Db = (int)W->bond[i + 1];
                j.cl = (int)psis[s]->bond[i]; j.cr = (int)psis[s]->bond[i + 1];
                j.row0 = (int)r0; j.rows = (int)std::min(rows_per, L - r0);
                jobs.push_back(j);
            }
        }
    ApplyJob* d_jobs = (ApplyJob*)ctx->alloc(sizeof(ApplyJob) * jobs.size());
    QIL_CUDA(cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(ApplyJob) * jobs.size(), cudaMemcpyHostToDevice, ctx->stream));
    {
        qil_prof_region prof_guard_(ctx, PROF_APPLY);
        const unsigned grid = (unsigned)jobs.size();
        if (W->is_complex && pc) apply_jobs_kernel<cplx, cplx, cplx><<<grid, 256, 0, ctx->stream>>>(d_jobs);
        else if (W->is_complex) apply_jobs_kernel<cplx, double, cplx><<<grid, 256, 0, ctx->stream>>>(d_jobs);
        else if (pc) apply_jobs_kernel<double, cplx, cplx><<<grid, 256, 0, ctx->stream>>>(d_jobs);
        else apply_jobs_kernel<double, double, double><<<grid, 256, 0, ctx->stream>>>(d_jobs);
        QIL_LAUNCH_CHECK(ctx);
    }
    ctx->free(d_jobs);
    for (int64_t s = 0; s < count; ++s) {
        qil_mps* out = new_mps(ctx, n, oc, ob[s].data(), false);
        for (int i = 0; i < n; ++i) out->core[i] = (char*)pool_raw + core_off[s * n + i] * es;
        out->pool = pool;
        out->amplitude = psis[s]->amplitude;
        outs[s] = out;
    }
}

qil_mps* apply_mpo_mps(qil_ctx* ctx, const qil_mpo* W, const qil_mps* psi) {
    QIL_REQUIRE(W->n == psi->n, QIL_ERR_ARGUMENT,
                "apply: MPO and MPS must have the same number of sites. Found length(W)=%d, length(psi)=%d",
                W->n, psi->n);
    qil_mps* out = nullptr;
    apply_mpo_mps_many(ctx, W, &psi, 1, &out);
    return out;
}

qil_mpo* apply_mpo_mpo(qil_ctx* ctx, const qil_mpo* W1, const qil_mpo* W2, int start1, int start2) {
    QIL_REQUIRE(start1 >= 0 && start1 < W1->n && start2 >= 0 && start2 < W2->n, QIL_ERR_ARGUMENT,
                "apply: No matching sites found");
    const int n1 = W1->n, n2 = W2->n;
    const int match = std::min(n1 - start1, n2 - start2);
    const bool base1 = n1 >= n2;
    const qil_mpo* base = base1 ? W1 : W2;
    const int bstart = base1 ? start1 : start2;
    const int oc = (W1->is_complex || W2->is_complex) ? 1 : 0;
    // the operand that is not the base must lie inside the window with closed boundary bonds
    const qil_mpo* other = base1 ? W2 : W1;
    const int ostart = base1 ? start2 : start1;
    QIL_REQUIRE(ostart == 0 && match == other->n, QIL_ERR_UNSUPPORTED,
                "apply(MPO,MPO): the shorter operator must lie entirely inside the longer one");
    std::vector<int64_t> ob(base->bond);
    for (int i = 0; i <= match; ++i) ob[bstart + i] = W1->bond[start1 + i] * W2->bond[start2 + i];
    qil_mpo* out = new_mpo(ctx, base->n, oc, ob.data(), true);
    // copy the cores outside the window (promoting to complex when needed)
    ApplyDesc d;
    d.n = match;
    long long mx = 1;
    for (int i = 0; i <= match; ++i) {
        d.wb[i] = (int)W1->bond[start1 + i];
        d.pb[i] = (int)W2->bond[start2 + i];
    }
    for (int i = 0; i < match; ++i) {
        d.w[i] = W1->core[start1 + i];
        d.p[i] = W2->core[start2 + i];
        d.o[i] = out->core[bstart + i];
        mx = std::max<long long>(mx, ob[bstart + i] * ob[bstart + i + 1]);
    }
    if (W1->is_complex && W2->is_complex)
        launch_sites(ctx, apply_mpo_mpo_kernel<cplx, cplx, cplx>, d, mx);
    else if (W1->is_complex)
        launch_sites(ctx, apply_mpo_mpo_kernel<cplx, double, cplx>, d, mx);
    else if (W2->is_complex)
        launch_sites(ctx, apply_mpo_mpo_kernel<double, cplx, cplx>, d, mx);
    else
        launch_sites(ctx, apply_mpo_mpo_kernel<double, double, double>, d, mx);
    // sites of the base outside the window are kept as they are
    for (int i = 0; i < base->n; ++i) {
        if (i >= bstart && i < bstart + match) continue;
        QIL_REQUIRE(base->is_complex == oc, QIL_ERR_UNSUPPORTED,
                    "apply(MPO,MPO): mixed real/complex operands with a partial window");
        QIL_CUDA(cudaMemcpyAsync(out->core[i], base->core[i], base->core_elems(i) * elem_size(oc),
                                 cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return out;
}

}  // namespace qil
