// qil_node.cu -- the levels of the divide-and-conquer encoder below the top split, one CTA per tree node
// (SignalConverters.jl:145-186 calling rsvd.jl:38-121 at every node).
//
// Below the top split every node matrix A = T[lb * 2^nl, 2^nr * rb] is small (its bonds are <= k+p), but its shape
// depends on the ranks found one level up.  Reading those ranks back to the host, as round 1 did, costs a pipeline
// drain per node (and ~15 launches per randomized split).  Here the host launches ONE kernel per tree level with a
// static node table; each CTA reads its node's bond dimensions from the device-side bond array, decides itself
// whether the node is an exact SVD (min(R, C) <= k+p) or a randomized one, and runs the whole split in shared memory:
//   Omega (counter-based normal stream or the host's), Y = A Omega on DMMA (cta_gemm), Householder TSQR (cta_qr),
//   q power iterations, B^H = A^H Q, QR of B^H, one-warp Jacobi with the NDTensors truncation rule, and the two
//   products U = Q Us (-> left child) and S Vh = (Us^H G) Qb^H (-> right child), written compactly with the new rank
//   as leading dimension.  The new rank goes into the bond array for the level below.
// A node that does not fit the CTA's shared memory raises the overflow flag; the host then repeats the encode on the
// general multi-launch path (qil_encode.cu) -- never a silent wrong answer.
#include "qil_fast.cuh"
#include "qil_rng.cuh"
#include "qil_wqr.cuh"
#include "qil_wy.cuh"

namespace qil {

constexpr int kNodeThreads = 256;   // 8 warps, up to 255 registers: at 512 threads the 128-register cap spilled (24 % of the
                                    // stall samples of the level-2 launch sat on local-memory stores)

struct NodeParams {
    const NodeDesc* nodes;      // [gridDim.x]
    int* bonds;                 // bond array(s): node.bonds_off + position
    int* overflow;              // set to 1 when a node does not fit
    double* margin;
    int kp;                     // k + p
    int q;
    double cutoff;
    long long maxdim, mindim;
    const double* stream;       // host-supplied normal stream or nullptr
    long long stream_len;
    unsigned long long seed;
    int smem_elems;             // doubles of dynamic shared memory available
    int fast;                   // register-resident factor + compact-WY explicit Q (cta_qr_fast) where the panel allows
};

__global__ void __launch_bounds__(kNodeThreads) node_split_kernel(const NodeParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    __shared__ int s_rank;
    __shared__ double s_nu;
    const NodeDesc nd = p.nodes[blockIdx.x];
    const int tid = threadIdx.x, nth = blockDim.x;
    __shared__ int s_skip;
    if (tid == 0) s_skip = *((volatile int*)p.overflow);       // an earlier level already gave up (uniform decision)
    __syncthreads();
    if (s_skip) return;
    int* bonds = p.bonds + nd.bonds_off;
    const int lb = bonds[nd.lb_pos], rb = bonds[nd.rb_pos];
    const int R = lb << nd.nl, C = rb << nd.nr;
    const double* A = nd.A;
    const bool exact = min(R, C) <= p.kp;
    const bool tall = exact && R >= C;                         // exact node factored as A = Q Rm; else via the adjoint
    const int l = min(p.kp, min(R, C));                        // sketch width == number of QR columns in every case
    const int pitch = wqr_pitch(l), pg = l | 1;
    const int mt = max(R, C);
    // ---- shared-memory carve-up
    size_t need;
    double *panR = nullptr, *panC = nullptr;
    {
        size_t off = 0;
        if (!exact) { panR = sm + off; off += (size_t)R * pitch; panC = sm + off; off += (size_t)C * pitch; }
        else if (tall) { panR = sm + off; off += (size_t)R * pitch; }
        else { panC = sm + off; off += (size_t)C * pitch; }
        need = off;
    }
    need = (need + 1) & ~(size_t)1;                            // 16-byte aligned work area
    double* work = sm + need;
    // panels of up to 768 rows (512 beyond 24 columns) are factored in ONE level by the register-resident Householder and
    // their explicit Q is formed in compact-WY form; larger ones keep the multi-level reflector form
    bool fastqr = p.fast && mt <= rqr_max_rows(l);
    {
        const size_t slow = cta_qr_extra_elems<double>(mt, l), fast = cta_qr_fast_elems(l);
        // the fast form needs the larger work area: a node that only fits with the reflector form keeps it (an overflow
        // would send the whole encode to the general multi-launch path)
        const size_t rest = (size_t)l * l + 3 * (size_t)l * pg + (size_t)l * l + l + (l + 1) / 2 + 8;
        if (fastqr && need + std::max(slow, fast) + rest > (size_t)p.smem_elems) fastqr = false;
        need += (fastqr ? std::max(slow, fast) : slow) + 1;
        need = (need + 1) & ~(size_t)1;
    }
    double* Rm = sm + need;  need += (size_t)l * l;             // triangle of the last QR
    double* G0 = sm + need;  need += (size_t)l * pg;            // column-major
    double* Gw = sm + need;  need += (size_t)l * pg;
    double* Us = sm + need;  need += (size_t)l * pg;            // row-major [i][j]
    double* T2 = sm + need;  need += (size_t)l * l;             // row-major [j][c]
    double* sig = sm + need; need += l;
    int* order = reinterpret_cast<int*>(sm + need); need += (l + 1) / 2 + 1;
    if (need > (size_t)p.smem_elems) {
        if (tid == 0) *p.overflow = 1;
        return;
    }
    if (!exact && p.stream && (long long)C * l > p.stream_len) {
        if (tid == 0) *p.overflow = 2;                         // supplied normal stream too short: the host reports it
        return;
    }

    // ---- load: Omega (randomized) or the matrix itself in tall orientation (exact); padding columns zero
    if (!exact) {
        for (int idx = tid; idx < C * pitch; idx += nth) {
            const int c = idx / pitch, j = idx - c * pitch;
            panC[idx] = (j < l) ? stream_at<double>(p.stream, p.seed, (long long)c + (long long)C * j) : 0.0;
        }
        for (int idx = tid; idx < R * pitch; idx += nth) panR[idx] = 0.0;
    } else if (tall) {
        for (int idx = tid; idx < R * pitch; idx += nth) {
            const int i = idx / pitch, c = idx - i * pitch;
            panR[idx] = (c < C) ? A[(size_t)i * C + c] : 0.0;
        }
    } else {
        for (int idx = tid; idx < C * pitch; idx += nth) {
            const int c = idx / pitch, i = idx - c * pitch;
            panC[idx] = (i < R) ? A[(size_t)i * C + c] : 0.0;
        }
    }
    __syncthreads();
    // ---- the chain of QRs: randomized  Y=A*Omega | (Z=A^H Q | Y=A Qz) x q | B^H=A^H Q ; exact: the matrix itself.
    // One call site each for the GEMM and the QR (the loop body) keeps the kernel small.
    const int nqr = exact ? 1 : 2 + 2 * p.q;
    for (int s = 0; s < nqr; ++s) {
        const bool onR = exact ? tall : ((s & 1) == 0);          // which panel is factored in this step
        const bool last = (s == nqr - 1);
        double* pan = onR ? panR : panC;
        const int rows = onR ? R : C;
        if (!exact) {
            // onR: Y (R x l) = A * panC ;  else: Z (C x l) = A^T * panR
            cta_gemm(!onR, A, C, rows, onR ? C : R, onR ? panC : panR, pitch, l, pan, pitch, 1.0);
            __syncthreads();
        }
        if (fastqr) cta_qr_fast(pan, pitch, rows, l, !last, last ? Rm : nullptr, l, work);
        else cta_qr<double>(pan, pitch, rows, l, !last, last ? Rm : nullptr, l, nullptr, 0, 0, work);
    }
    // ---- G: the l x l matrix whose left singular vectors are wanted.  tall: G = Rm; otherwise G = Rm^H.
    for (int idx = tid; idx < l * pg; idx += nth) {
        const int j = idx / pg, i = idx - j * pg;               // column j, row i
        double v = 0.0;
        if (i < l) v = tall ? Rm[i * l + j] : Rm[j * l + i];
        G0[idx] = v;
    }
    __syncthreads();
    cta_jacobi_rank<double>(Gw, G0, pg, l, p.cutoff, p.maxdim, p.mindim, sig, order, &s_rank, p.margin, &s_nu);
    const int r = s_rank;
    if (tid == 0) bonds[nd.out_pos] = r;
    for (int idx = tid; idx < l * r; idx += nth) {
        const int i = idx / r, j = idx - i * r;
        const double sj = sig[j];
        Us[i * pg + j] = Gw[order[j] * pg + i] * (sj > 0.0 ? 1.0 / sj : 0.0);
    }
    __syncthreads();
    for (int idx = tid; idx < r * l; idx += nth) {              // T2 = Us^T G
        const int j = idx / l, c = idx - j * l;
        double acc = 0.0;
        for (int k = 0; k < l; ++k) acc = fma(Us[k * pg + j], G0[c * pg + k], acc);
        T2[j * l + c] = acc;
    }
    __syncthreads();
    // ---- outputs, compact with leading dimension r (U) / C (SVh)
    double* Uo = nd.U;
    double* So = nd.SVh;
    if (exact && !tall) {
        for (int idx = tid; idx < R * r; idx += nth) {          // U = Us (R == l)
            const int i = idx / r, j = idx - i * r;
            Uo[idx] = Us[i * pg + j];
        }
    } else {
        for (int idx = tid; idx < R * r; idx += nth) {          // U = Q Us
            const int i = idx / r, j = idx - i * r;
            const double* qrow = panR + (size_t)i * pitch;
            double acc = 0.0;
            for (int k = 0; k < l; ++k) acc = fma(qrow[k], Us[k * pg + j], acc);
            Uo[idx] = acc;
        }
    }
    if (tall) {
        for (int idx = tid; idx < r * C; idx += nth) So[idx] = T2[idx];        // S Vh = Us^T Rm  (l == C)
    } else {
        for (int idx = tid; idx < r * C; idx += nth) {          // S Vh = T2 Qb^T
            const int j = idx / C, c = idx - j * C;
            const double* qrow = panC + (size_t)c * pitch;
            double acc = 0.0;
            for (int k = 0; k < l; ++k) acc = fma(T2[j * l + k], qrow[k], acc);
            So[idx] = acc;
        }
    }
}

void node_level_launch(qil_ctx* ctx, const NodeDesc* d_nodes, int count, int* d_bonds, int* d_overflow, const RsvdOpts& o,
                       const double* d_stream, int64_t stream_len) {
    if (count <= 0) return;
    NodeParams p;
    p.nodes = d_nodes; p.bonds = d_bonds; p.overflow = d_overflow; p.margin = ctx->d_margin;
    p.kp = o.k + o.p; p.q = o.q; p.cutoff = o.cutoff; p.maxdim = o.maxdim; p.mindim = o.mindim;
    p.stream = d_stream; p.stream_len = stream_len; p.seed = (unsigned long long)o.seed;
    static const bool fast = [] { const char* e = getenv("QIL_NODE_FAST"); return !(e && e[0] == '0'); }();
    p.fast = fast ? 1 : 0;
    const size_t smem = std::min<size_t>(ctx->smem_optin, 227 * 1024) - 1024;
    p.smem_elems = (int)(smem / sizeof(double));
    ensure_dynamic_smem(node_split_kernel, smem);
    node_split_kernel<<<count, kNodeThreads, smem, ctx->stream>>>(p);
    QIL_LAUNCH_CHECK(ctx);
}

}  // namespace qil
