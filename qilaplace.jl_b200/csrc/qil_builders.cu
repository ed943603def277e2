// qil_builders.cu -- transform MPO builders orchestrated on the device toolkit (setup path):
//   build_qft_mpo : zip-up (QR) + zip-down (truncated SVD)       src/transforms/qft_transformer.jl:13-165
//   build_dt_mpo  : zip-to-combine (QR) + zip-to-compress        src/transforms/dt_transformer.jl:20-412
//   build_zt_mpo  : compress(apply(W_dt, W_qft_paired))          src/transforms/zt_transformer.jl:41-112
// Gate blocks follow src/circuits/{qft,dt,zt}_gates.jl.  MPO cores are [l][p][s][r] row-major with
// p = primed/input leg, s = output leg; paired operators are 2n-site chains (main1, copy1, ...).
#include "qil_mpsops.cuh"

#include <cmath>
#include <complex>

namespace qil {

template <typename T>
struct Core {
    Mat<T> m;
    int64_t l = 1, r = 1;
};
template <typename T>
using Chain = std::vector<Core<T>>;

// ---- host-side gate blocks -------------------------------------------------------------------
template <typename T> struct G2 { T v[2][2]; };

template <typename T> static T mk(double re, double im);
template <> double mk<double>(double re, double) { return re; }
template <> cplx mk<cplx>(double re, double im) { return make_double2(re, im); }

template <typename T> static G2<T> gate_I() { return {{{mk<T>(1, 0), mk<T>(0, 0)}, {mk<T>(0, 0), mk<T>(1, 0)}}}; }
template <typename T> static G2<T> gate_H() {
    const double h = 1.0 / std::sqrt(2.0);
    return {{{mk<T>(h, 0), mk<T>(h, 0)}, {mk<T>(h, 0), mk<T>(-h, 0)}}};
}
static G2<cplx> gate_P(double theta) {  // diag(1, exp(-i theta))  (qft_gates.jl:24-30)
    return {{{mk<cplx>(1, 0), mk<cplx>(0, 0)}, {mk<cplx>(0, 0), mk<cplx>(std::cos(theta), -std::sin(theta))}}};
}
static G2<double> gate_R(double f) {    // diag(1, exp(-f))        (dt_gates.jl:19-25)
    return {{{1.0, 0.0}, {0.0, std::exp(-f)}}};
}
static G2<double> gate_Hd(double wr) {  // damped Hadamard          (dt_gates.jl:11-17)
    const double h = 1.0 / std::sqrt(2.0);
    return {{{h, h}, {h, h * std::exp(-wr / 2.0)}}};
}

template <typename T>
static Core<T> upload(qil_ctx* ctx, const std::vector<T>& h, int64_t l, int64_t r) {
    Core<T> c;
    c.l = l;
    c.r = r;
    c.m = Mat<T>(ctx, l * 4, r);
    QIL_CUDA(cudaMemcpyAsync(c.m.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    QIL_CUDA(cudaStreamSynchronize(ctx->stream));  // h is a temporary
    return c;
}

// bond-diagonal core: g0 on bond value 0, g1 on bond value 1; open (dim 1) sides sum/broadcast
template <typename T>
static Core<T> ctl(qil_ctx* ctx, const G2<T>& g0, const G2<T>& g1, bool left, bool right) {
    const int64_t l = left ? 2 : 1, r = right ? 2 : 1;
    std::vector<T> h((size_t)l * 4 * r, Scalar<T>::zero());
    for (int b = 0; b < 2; ++b) {
        const G2<T>& g = b ? g1 : g0;
        const int lb = left ? b : 0, rb = right ? b : 0;
        for (int p = 0; p < 2; ++p)
            for (int s = 0; s < 2; ++s) {
                T& dst = h[(((size_t)lb * 2 + p) * 2 + s) * r + rb];
                dst = Scalar<T>::add(dst, g.v[p][s]);
            }
    }
    return upload<T>(ctx, h, l, r);
}

template <typename T>
static Core<T> single(qil_ctx* ctx, const G2<T>& g) {
    std::vector<T> h(4);
    for (int p = 0; p < 2; ++p)
        for (int s = 0; s < 2; ++s) h[p * 2 + s] = g.v[p][s];
    return upload<T>(ctx, h, 1, 1);
}

// control_Hphase_mpo (qft_gates.jl:43-97)
static Chain<cplx> control_hphase(qil_ctx* ctx, int k) {
    Chain<cplx> c;
    if (k == 1) {
        c.push_back(single<cplx>(ctx, gate_H<cplx>()));
        return c;
    }
    {   // site 1: W[p][s][b] = H[p][s] * [s == b]
        const G2<cplx> H = gate_H<cplx>();
        std::vector<cplx> h(8, Scalar<cplx>::zero());
        for (int p = 0; p < 2; ++p)
            for (int s = 0; s < 2; ++s) h[(p * 2 + s) * 2 + s] = H.v[p][s];
        c.push_back(upload<cplx>(ctx, h, 1, 2));
    }
    for (int l = 2; l <= k; ++l)
        c.push_back(ctl<cplx>(ctx, gate_I<cplx>(), gate_P(2.0 * M_PI / std::pow(2.0, l)), true, l < k));
    return c;
}

// control_damping_mpo (dt_gates.jl:30-130)
static Chain<double> control_damping(qil_ctx* ctx, int k, double wr) {
    Chain<double> c;
    if (k == 1) {
        c.push_back(single<double>(ctx, gate_Hd(wr)));
        c.push_back(single<double>(ctx, gate_I<double>()));
        return c;
    }
    for (int l = 1; l < k; ++l) {
        c.push_back(ctl<double>(ctx, gate_I<double>(), gate_R(wr * std::pow(2.0, l - k - 1)), l > 1, true));
        c.push_back(ctl<double>(ctx, gate_I<double>(), gate_I<double>(), true, true));
    }
    {   // main k: W[b][p][s][b] = [p == b] * Hd[p][s]
        const G2<double> Hd = gate_Hd(wr);
        std::vector<double> h(16, 0.0);
        for (int b = 0; b < 2; ++b)
            for (int s = 0; s < 2; ++s) h[((b * 2 + b) * 2 + s) * 2 + b] = Hd.v[b][s];
        c.push_back(upload<double>(ctx, h, 2, 2));
    }
    c.push_back(ctl<double>(ctx, gate_I<double>(), gate_I<double>(), true, false));
    return c;
}

// control_damping_copy_mpo (dt_gates.jl:133-229), L = n-k+1 pairs
static Chain<double> control_damping_copy(qil_ctx* ctx, int n, int k, double wr) {
    const int L = n - k + 1;
    Chain<double> c;
    if (L == 1) {
        c.push_back(single<double>(ctx, gate_I<double>()));
        c.push_back(single<double>(ctx, gate_I<double>()));
        return c;
    }
    {   // main[1]: identity, right bond value 0 only
        std::vector<double> h(8, 0.0);
        for (int p = 0; p < 2; ++p) h[(p * 2 + p) * 2 + 0] = 1.0;
        c.push_back(upload<double>(ctx, h, 1, 2));
    }
    {   // copy[1]: projector |b><b|, left bond value 0, right bond value b
        std::vector<double> h(16, 0.0);
        for (int b = 0; b < 2; ++b) h[((0 * 2 + b) * 2 + b) * 2 + b] = 1.0;
        c.push_back(upload<double>(ctx, h, 2, 2));
    }
    for (int j = 2; j <= L; ++j) {
        c.push_back(ctl<double>(ctx, gate_I<double>(), gate_R(wr * std::pow(2.0, j - 2)), true, true));
        c.push_back(ctl<double>(ctx, gate_I<double>(), gate_I<double>(), true, j < L));
    }
    return c;
}

// control_Hphase_ztmps_mpo (zt_gates.jl:12-114)
static Chain<cplx> control_hphase_zt(qil_ctx* ctx, int k) {
    Chain<cplx> c;
    if (k == 1) {
        c.push_back(single<cplx>(ctx, gate_I<cplx>()));
        c.push_back(single<cplx>(ctx, gate_H<cplx>()));
        return c;
    }
    c.push_back(ctl<cplx>(ctx, gate_I<cplx>(), gate_I<cplx>(), false, true));  // main 1 opens both branches
    c.push_back(ctl<cplx>(ctx, gate_I<cplx>(), gate_P(2.0 * M_PI / std::pow(2.0, k)), true, true));
    for (int j = 2; j < k; ++j) {
        c.push_back(ctl<cplx>(ctx, gate_I<cplx>(), gate_I<cplx>(), true, true));
        c.push_back(ctl<cplx>(ctx, gate_I<cplx>(), gate_P(2.0 * M_PI / std::pow(2.0, k - j + 1)), true, true));
    }
    c.push_back(ctl<cplx>(ctx, gate_I<cplx>(), gate_I<cplx>(), true, true));
    {   // copy k: W[b][p][s] = [p == b] * H[p][s]
        const G2<cplx> H = gate_H<cplx>();
        std::vector<cplx> h(8, Scalar<cplx>::zero());
        for (int b = 0; b < 2; ++b)
            for (int s = 0; s < 2; ++s) h[(b * 2 + b) * 2 + s] = H.v[b][s];
        c.push_back(upload<cplx>(ctx, h, 2, 1));
    }
    return c;
}

// ---- contraction helpers -----------------------------------------------------------------------
static ContractDesc cdesc(int nout, int ncon) {
    ContractDesc d;
    memset(&d, 0, sizeof(d));
    d.nout = nout;
    d.ncon = ncon;
    return d;
}

// One step of an upward zip (qft_transformer.jl:34-57, dt_transformer.jl:99-141):
//   core[a,c,p,s,q] = sum_{m,b,d} M1[a,p,m,b] M2[c,m,s,d] T[b,d,q]   (M1 acts first)
//   QR with rows (p,s,q): site <- Q^T as [kk,p,s,q], T <- R^T as [a,c,kk]
template <typename T>
static void zip_up_step(qil_ctx* ctx, const Core<T>& M1, const Core<T>& M2, Mat<T>& Tm, int64_t& tq, Core<T>& out) {
    const int64_t a = M1.l, b = M1.r, c = M2.l, d = M2.r, q = tq;
    Mat<T> X(ctx, c * 4, b * q);  // [c][m][s][b][q]
    {
        ContractDesc e = cdesc(5, 1);
        const long long od[5] = {c, 2, 2, b, q};
        const long long sa[5] = {4 * d, 2 * d, d, 0, 0};
        const long long sb[5] = {0, 0, 0, d * q, 1};
        const long long sc[5] = {4 * b * q, 2 * b * q, b * q, q, 1};
        for (int i = 0; i < 5; ++i) { e.od[i] = od[i]; e.sa_o[i] = sa[i]; e.sb_o[i] = sb[i]; e.sc_o[i] = sc[i]; }
        e.cd[0] = d; e.sa_c[0] = 1; e.sb_c[0] = q;
        contract<T, T, T>(ctx, e, M2.m.p, Tm.p, X.p);
    }
    Mat<T> Mt(ctx, 4 * q, a * c);  // [(p,s,q)][(a,c)]
    {
        ContractDesc e = cdesc(5, 2);
        const long long od[5] = {2, 2, q, a, c};
        const long long sa[5] = {2 * b, 0, 0, 4 * b, 0};
        const long long sb[5] = {0, b * q, 1, 0, 4 * b * q};
        const long long sc[5] = {2 * q * a * c, q * a * c, a * c, c, 1};
        for (int i = 0; i < 5; ++i) { e.od[i] = od[i]; e.sa_o[i] = sa[i]; e.sb_o[i] = sb[i]; e.sc_o[i] = sc[i]; }
        e.cd[0] = 2; e.sa_c[0] = b; e.sb_c[0] = 2 * b * q;
        e.cd[1] = b; e.sa_c[1] = 1; e.sb_c[1] = q;
        contract<T, T, T>(ctx, e, M1.m.p, X.p, Mt.p);
    }
    Mat<T> Qm, R;
    qr_thin<T>(ctx, 4 * q, a * c, Mt.p, a * c, false, Qm, R);
    const int64_t kk = Qm.cols;
    out.m = Mat<T>(ctx, kk * 4, q);
    transpose_conj<T>(ctx, 4 * q, kk, Qm.p, kk, out.m.p, 4 * q, false);
    out.l = kk;
    out.r = q;
    Mat<T> Tn(ctx, a * c, kk);
    transpose_conj<T>(ctx, kk, a * c, R.p, a * c, Tn.p, kk, false);
    Tm = std::move(Tn);
    tq = kk;
}

// site[l,p,s,q'] = sum_a site[l,p,s,a] * T[a,0,q']   (T as (a x q'), the second operand had no left bond)
template <typename T>
static void absorb_right(qil_ctx* ctx, Core<T>& site, const Mat<T>& Tm, int64_t kk) {
    const int64_t a = site.r;
    Mat<T> n(ctx, site.l * 4, kk);
    gemm<T>(ctx, OP_N, OP_N, site.l * 4, kk, a, 1.0, site.m.p, a, Tm.p, kk, 0.0, n.p, kk);
    site.m = std::move(n);
    site.r = kk;
}

// site[q,p,s,r] = sum_a T[q,a] site[a,p,s,r]
template <typename T>
static void absorb_left(qil_ctx* ctx, Core<T>& site, const Mat<T>& Tm, int64_t q) {
    const int64_t a = site.l;
    Mat<T> n(ctx, q * 4, site.r);
    gemm<T>(ctx, OP_N, OP_N, q, 4 * site.r, a, 1.0, Tm.p, a, site.m.p, 4 * site.r, 0.0, n.p, 4 * site.r);
    site.m = std::move(n);
    site.l = q;
}

// zip_to_combine_mpos "down" (dt_transformer.jl:38-95); M1 acts first, then M2
template <typename T>
static void combine_down(qil_ctx* ctx, Chain<T>& M1, const Chain<T>& M2) {
    const size_t n1 = M1.size(), n2 = M2.size();
    Mat<T> Tm(ctx, 1, 1);
    {
        const T one = Scalar<T>::one();
        QIL_CUDA(cudaMemcpyAsync(Tm.p, &one, sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        QIL_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    int64_t tq = 1;  // T is [q][a][c]
    for (size_t kx = 0; kx < n2; ++kx) {
        const Core<T>& A = M1[kx];
        const Core<T>& B = M2[kx];
        const int64_t a = A.l, b = A.r, c = B.l, d = B.r, q = tq;
        Mat<T> X(ctx, q * c * 4, b);  // [q][c][p][m][b]
        {
            ContractDesc e = cdesc(5, 1);
            const long long od[5] = {q, c, 2, 2, b};
            const long long sa[5] = {a * c, 1, 0, 0, 0};
            const long long sb[5] = {0, 0, 2 * b, b, 1};
            const long long sc[5] = {c * 4 * b, 4 * b, 2 * b, b, 1};
            for (int i = 0; i < 5; ++i) { e.od[i] = od[i]; e.sa_o[i] = sa[i]; e.sb_o[i] = sb[i]; e.sc_o[i] = sc[i]; }
            e.cd[0] = a; e.sa_c[0] = c; e.sb_c[0] = 4 * b;
            contract<T, T, T>(ctx, e, Tm.p, A.m.p, X.p);
        }
        Mat<T> Cm(ctx, q * 4, b * d);  // [(q,p,s)][(b,d)]
        {
            ContractDesc e = cdesc(5, 2);
            const long long od[5] = {q, 2, 2, b, d};
            const long long sa[5] = {c * 4 * b, 2 * b, 0, 1, 0};
            const long long sb[5] = {0, 0, d, 0, 1};
            const long long sc[5] = {4 * b * d, 2 * b * d, b * d, d, 1};
            for (int i = 0; i < 5; ++i) { e.od[i] = od[i]; e.sa_o[i] = sa[i]; e.sb_o[i] = sb[i]; e.sc_o[i] = sc[i]; }
            e.cd[0] = c; e.sa_c[0] = 4 * b; e.sb_c[0] = 4 * d;
            e.cd[1] = 2; e.sa_c[1] = b; e.sb_c[1] = 2 * d;
            contract<T, T, T>(ctx, e, X.p, B.m.p, Cm.p);
        }
        if (kx == n2 - 1 && n1 == n2) {
            // empty right index set: Q*R is multiplied back together (dt_transformer.jl:73,90-94)
            M1[kx].m = std::move(Cm);
            M1[kx].l = q;
            M1[kx].r = 1;
            return;
        }
        Mat<T> Qm, R;
        qr_thin<T>(ctx, q * 4, b * d, Cm.p, b * d, false, Qm, R);
        const int64_t kk = Qm.cols;
        M1[kx].m = std::move(Qm);
        M1[kx].l = q;
        M1[kx].r = kk;
        Tm = std::move(R);  // [kk][b][d]
        tq = kk;
    }
    // remainder (kk x b, d == 1) goes into the next core of M1
    absorb_left<T>(ctx, M1[n2], Tm, tq);
}

// zip_to_combine_mpos "up" (dt_transformer.jl:97-153) == zip_up_mpos (qft_transformer.jl:13-66)
template <typename T>
static void combine_up(qil_ctx* ctx, Chain<T>& M1, const Chain<T>& M2) {
    const size_t n1 = M1.size(), n2 = M2.size();
    QIL_REQUIRE(n1 > n2, QIL_ERR_ARGUMENT, "zip_up_mpos: mpo1 must be longer than mpo2");
    Mat<T> Tm(ctx, 1, 1);
    {
        const T one = Scalar<T>::one();
        QIL_CUDA(cudaMemcpyAsync(Tm.p, &one, sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        QIL_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    int64_t tq = 1;
    for (size_t kx = 0; kx < n2; ++kx) {
        const size_t i1 = n1 - 1 - kx, i2 = n2 - 1 - kx;
        Core<T> out;
        zip_up_step<T>(ctx, M1[i1], M2[i2], Tm, tq, out);
        M1[i1] = std::move(out);
    }
    // T is [a][c=1][kk] -> (a x kk)
    absorb_right<T>(ctx, M1[n1 - n2 - 1], Tm, tq);
}

// zip_down_mpos (qft_transformer.jl:69-101): truncated SVD sweep oc .. L-1
template <typename T>
static void zip_down(qil_ctx* ctx, Chain<T>& M, size_t oc, double cutoff, int64_t maxdim) {
    for (size_t k = oc; k + 1 < M.size(); ++k) {
        Core<T>& A = M[k];
        Core<T>& B = M[k + 1];
        Mat<T> U, SVh;
        const int kk = svd_trunc<T>(ctx, A.l * 4, A.r, A.m.p, A.r, cutoff, maxdim, 1, &U, nullptr, nullptr, &SVh, nullptr);
        Mat<T> nb(ctx, (int64_t)kk * 4, B.r);
        gemm<T>(ctx, OP_N, OP_N, kk, 4 * B.r, A.r, 1.0, SVh.p, A.r, B.m.p, 4 * B.r, 0.0, nb.p, 4 * B.r);
        A.m = std::move(U);
        A.r = kk;
        B.m = std::move(nb);
        B.l = kk;
    }
}

// zip_to_compress_mpo over the full chain (dt_transformer.jl:167-288)
template <typename T>
static void compress_mpo(qil_ctx* ctx, Chain<T>& M, bool down, double cutoff, int64_t maxdim) {
    const int L = (int)M.size();
    if (L < 2) return;
    if (down) {
        for (int i = 0; i < L - 1; ++i) {  // QR gauge sweep 1 -> L
            Core<T>& A = M[i];
            Core<T>& B = M[i + 1];
            Mat<T> Qm, R;
            qr_thin<T>(ctx, A.l * 4, A.r, A.m.p, A.r, false, Qm, R);
            const int64_t kk = Qm.cols;
            Mat<T> nb(ctx, kk * 4, B.r);
            gemm<T>(ctx, OP_N, OP_N, kk, 4 * B.r, A.r, 1.0, R.p, A.r, B.m.p, 4 * B.r, 0.0, nb.p, 4 * B.r);
            A.m = std::move(Qm); A.r = kk;
            B.m = std::move(nb); B.l = kk;
        }
        for (int i = L - 1; i >= 1; --i) {  // two-site SVD sweep L -> 2: U*S -> site i-1, V -> site i
            Core<T>& A = M[i - 1];
            Core<T>& B = M[i];
            Mat<T> th(ctx, A.l * 4, 4 * B.r);
            gemm<T>(ctx, OP_N, OP_N, A.l * 4, 4 * B.r, A.r, 1.0, A.m.p, A.r, B.m.p, 4 * B.r, 0.0, th.p, 4 * B.r);
            Mat<T> US, Vh;
            const int kk = svd_trunc<T>(ctx, A.l * 4, 4 * B.r, th.p, 4 * B.r, cutoff, maxdim, 1, nullptr, &US, &Vh, nullptr, nullptr);
            A.m = std::move(US); A.r = kk;
            B.m = std::move(Vh); B.l = kk;
        }
    } else {
        for (int i = L - 1; i >= 1; --i) {  // QR gauge sweep L -> 1 (rows = site legs + right bond)
            Core<T>& A = M[i];
            Core<T>& P = M[i - 1];
            Mat<T> At(ctx, 4 * A.r, A.l);
            transpose_conj<T>(ctx, A.l, 4 * A.r, A.m.p, 4 * A.r, At.p, A.l, false);
            Mat<T> Qm, R;
            qr_thin<T>(ctx, 4 * A.r, A.l, At.p, A.l, false, Qm, R);
            const int64_t kk = Qm.cols;
            Mat<T> na(ctx, kk * 4, A.r);
            transpose_conj<T>(ctx, 4 * A.r, kk, Qm.p, kk, na.p, 4 * A.r, false);
            // prev[., l] * R^T (l x kk)
            Mat<T> np(ctx, P.l * 4, kk);
            gemm<T>(ctx, OP_N, OP_T, P.l * 4, kk, A.l, 1.0, P.m.p, A.l, R.p, A.l, 0.0, np.p, kk);
            A.m = std::move(na); A.l = kk;
            P.m = std::move(np); P.r = kk;
        }
        for (int i = 0; i < L - 1; ++i) {  // two-site SVD sweep 1 -> L-1: U -> site i, S*V -> site i+1
            Core<T>& A = M[i];
            Core<T>& B = M[i + 1];
            Mat<T> th(ctx, A.l * 4, 4 * B.r);
            gemm<T>(ctx, OP_N, OP_N, A.l * 4, 4 * B.r, A.r, 1.0, A.m.p, A.r, B.m.p, 4 * B.r, 0.0, th.p, 4 * B.r);
            Mat<T> U, SVh;
            const int kk = svd_trunc<T>(ctx, A.l * 4, 4 * B.r, th.p, 4 * B.r, cutoff, maxdim, 1, &U, nullptr, nullptr, &SVh, nullptr);
            A.m = std::move(U); A.r = kk;
            B.m = std::move(SVh); B.l = kk;
        }
    }
}

template <typename T>
static void extend_identity_pair(qil_ctx* ctx, Chain<T>& M) {
    M.push_back(single<T>(ctx, gate_I<T>()));
    M.push_back(single<T>(ctx, gate_I<T>()));
}

template <typename T>
static qil_mpo* to_handle(qil_ctx* ctx, Chain<T>& M) {
    const int n = (int)M.size();
    std::vector<int64_t> bond(n + 1, 1);
    for (int i = 0; i < n; ++i) {
        QIL_REQUIRE(M[i].l == bond[i], QIL_ERR_RUNTIME, "builder: inconsistent bond at site %d", i);
        bond[i + 1] = M[i].r;
    }
    qil_mpo* h = new_mpo(ctx, n, Scalar<T>::is_complex ? 1 : 0, bond.data(), false);
    for (int i = 0; i < n; ++i) h->core[i] = M[i].m.take();
    return h;
}

template <typename T>
static Chain<T> from_handle(qil_ctx* ctx, const qil_mpo* h) {
    Chain<T> M(h->n);
    for (int i = 0; i < h->n; ++i) {
        M[i].l = h->bond[i];
        M[i].r = h->bond[i + 1];
        M[i].m = Mat<T>(ctx, M[i].l * 4, M[i].r);
        QIL_CUDA(cudaMemcpyAsync(M[i].m.p, h->core[i], h->core_elems(i) * sizeof(T), cudaMemcpyDeviceToDevice,
                                 ctx->stream));
    }
    return M;
}

// ---- builders --------------------------------------------------------------------------------------
qil_mpo* build_qft_mpo(qil_ctx* ctx, int n, double cutoff, int64_t maxdim) {
    QIL_REQUIRE(n >= 1, QIL_ERR_ARGUMENT, "build_qft_mpo: Number of qubits 'n' must be at least 1. Found n=%d", n);
    QIL_REQUIRE(n <= kMaxSites, QIL_ERR_UNSUPPORTED, "build_qft_mpo: n=%d exceeds %d sites", n, kMaxSites);
    Chain<cplx> qft = control_hphase(ctx, n);
    for (int it = 1; it < n; ++it) {
        Chain<cplx> m2 = control_hphase(ctx, n - it);
        combine_up<cplx>(ctx, qft, m2);                 // zip_up_mpos, oc -> it
        zip_down<cplx>(ctx, qft, (size_t)(it - 1), cutoff, maxdim);
    }
    return to_handle<cplx>(ctx, qft);
}

static Chain<double> build_dt_chain(qil_ctx* ctx, int n, double wr, double cutoff, int64_t maxdim) {
    Chain<double> M = control_damping(ctx, 1, wr);
    if (n == 1) return M;
    for (int k = 2; k <= n; ++k) {  // Part 1 (dt_transformer.jl:348-390)
        extend_identity_pair<double>(ctx, M);
        Chain<double> blk = control_damping(ctx, k, wr);
        combine_down<double>(ctx, M, blk);
        compress_mpo<double>(ctx, M, true, cutoff, maxdim);
    }
    for (int k = 1; k < n; ++k) {   // Part 2 (dt_transformer.jl:396-405)
        Chain<double> blk = control_damping_copy(ctx, n, k, wr);
        if (blk.size() == M.size()) combine_down<double>(ctx, M, blk);
        else combine_up<double>(ctx, M, blk);
        compress_mpo<double>(ctx, M, false, cutoff, maxdim);
    }
    return M;
}

qil_mpo* build_dt_mpo(qil_ctx* ctx, int n, double wr, double cutoff, int64_t maxdim) {
    QIL_REQUIRE(n >= 1, QIL_ERR_ARGUMENT, "build_dt_mpo: n must be >= 1. Found n=%d", n);
    QIL_REQUIRE(2 * n <= kMaxSites, QIL_ERR_UNSUPPORTED, "build_dt_mpo: n=%d exceeds %d sites", n, kMaxSites / 2);
    Chain<double> M = build_dt_chain(ctx, n, wr, cutoff, maxdim);
    return to_handle<double>(ctx, M);
}

qil_mpo* build_zt_mpo(qil_ctx* ctx, int n, double wr, double cutoff, int64_t maxdim) {
    QIL_REQUIRE(n >= 1, QIL_ERR_ARGUMENT, "build_zt_mpo: n must be >= 1. Found n=%d", n);
    QIL_REQUIRE(2 * n <= kMaxSites, QIL_ERR_UNSUPPORTED, "build_zt_mpo: n=%d exceeds %d sites", n, kMaxSites / 2);
    Chain<double> Wdt = build_dt_chain(ctx, n, wr, cutoff, maxdim);
    Chain<cplx> Wq = control_hphase_zt(ctx, 1);
    for (int k = 2; k <= n; ++k) {  // paired QFT (zt_transformer.jl:78-99)
        extend_identity_pair<cplx>(ctx, Wq);
        Chain<cplx> blk = control_hphase_zt(ctx, k);
        combine_down<cplx>(ctx, Wq, blk);
        compress_mpo<cplx>(ctx, Wq, true, cutoff, maxdim);
    }
    chain_owner<qil_mpo> hdt(to_handle<double>(ctx, Wdt));
    chain_owner<qil_mpo> hq(to_handle<cplx>(ctx, Wq));
    chain_owner<qil_mpo> fused(apply_mpo_mpo(ctx, hdt.get(), hq.get(), 0, 0));   // apply(W_dt, mpo_qft) (zt_transformer.jl:103)
    hdt.reset();
    hq.reset();
    if (n == 1) return fused.release();
    Chain<cplx> Wzt = from_handle<cplx>(ctx, fused.get());
    fused.reset();
    compress_mpo<cplx>(ctx, Wzt, true, cutoff, maxdim);
    return to_handle<cplx>(ctx, Wzt);
}

}  // namespace qil
