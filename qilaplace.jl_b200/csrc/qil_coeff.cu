// qil_coeff.cu -- K8: batched `coefficient` (reference: src/mps.jl:669-678).
//
// amp <- amp * A_i[:, b_i, :] for i = 1..n, result scaled by the stored amplitude.
// One CTA walks the whole chain for a tile of S bitstrings.  The running vectors V[s][chi] live in
// shared memory (ping-pong); the selected core slices are streamed from L2/HBM with coalesced loads
// along the right bond and are reused by every bitstring of the tile (register tile of ST strings
// per thread), which is what lets the kernel run above the one-slice-per-bitstring HBM roofline.
#include "qil_common.cuh"

namespace qil {

constexpr int kCoeffThreads = 256;
constexpr int kCoeffWarps = kCoeffThreads / 32;
constexpr int kST = 16;  // max bitstrings per thread (register tile)

__host__ __device__ inline int pow2_ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

template <typename T>
__global__ void __launch_bounds__(kCoeffThreads)
coeff_chain_kernel(const ChainDesc d, const uint8_t* __restrict__ bits, long long B, T* __restrict__ out,
                   double amplitude, int S, int chi_pad) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];   // one declaration (and alignment) per translation unit
    T* V0 = reinterpret_cast<T*>(smem_raw);
    T* V1 = V0 + (size_t)S * chi_pad;
    uint8_t* sb = reinterpret_cast<uint8_t*>(V1 + (size_t)S * chi_pad);  // [S][n]

    const int n = d.n;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const long long ntiles = (B + S - 1) / S;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile * S;
        const int ns = (int)min((long long)S, B - s0);
        __syncthreads();
        // bits of this tile -> smem (contiguous block of ns*n bytes)
        for (int idx = tid; idx < S * n; idx += kCoeffThreads)
            sb[idx] = (idx < ns * n) ? bits[s0 * n + idx] : (uint8_t)0;
        for (int s = tid; s < S; s += kCoeffThreads) V0[(size_t)s * chi_pad] = Scalar<T>::one();
        __syncthreads();

        T* Vc = V0;
        T* Vn = V1;
        for (int i = 0; i < n; ++i) {
            const int cl = d.bond[i], cr = d.bond[i + 1];
            const T* __restrict__ M = reinterpret_cast<const T*>(d.core[i]);
            // thread roles for this site
            const int rl = min(32, pow2_ceil(cr));          // lanes across the right bond
            const int rw = (rl == 32) ? min(kCoeffWarps, pow2_ceil((cr + 31) / 32)) : 1;  // warps across r
            const int sub = 32 / rl;                         // lane sub-groups, each with its own strings
            const int sgroups = sub * (kCoeffWarps / rw);
            const int st = (S + sgroups - 1) / sgroups;      // strings per thread (<= kST by construction)
            const int g = (warp / rw) * sub + lane / rl;
            const int rbase = (warp % rw) * 32 + (lane % rl);
            const int rstride = rw * rl;
            const int sfirst = g * st;

            unsigned bm = 0;  // bit of string (sfirst + k) at site i
#pragma unroll
            for (int k = 0; k < kST; ++k)
                if (k < st && sfirst + k < S) bm |= (unsigned)(sb[(sfirst + k) * n + i] & 1) << k;

            for (int r = rbase; r < cr; r += rstride) {
                T acc[kST];
#pragma unroll
                for (int k = 0; k < kST; ++k) acc[k] = Scalar<T>::zero();
#pragma unroll 2
                for (int l = 0; l < cl; ++l) {
                    const T m0 = __ldg(M + ((size_t)l * 2 + 0) * cr + r);
                    const T m1 = __ldg(M + ((size_t)l * 2 + 1) * cr + r);
#pragma unroll
                    for (int k = 0; k < kST; ++k) {
                        if (k < st) {
                            const int s = min(sfirst + k, S - 1);
                            const T v = Vc[(size_t)s * chi_pad + l];
                            const T m = ((bm >> k) & 1u) ? m1 : m0;
                            acc[k] = Scalar<T>::fma(v, m, acc[k]);
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < kST; ++k)
                    if (k < st && sfirst + k < S) Vn[(size_t)(sfirst + k) * chi_pad + r] = acc[k];
            }
            __syncthreads();
            T* t = Vc;
            Vc = Vn;
            Vn = t;
        }
        for (int s = tid; s < ns; s += kCoeffThreads)
            out[s0 + s] = Scalar<T>::scale(Vc[(size_t)s * chi_pad], amplitude);
    }
}

template <typename T>
static void launch_coeff(qil_ctx* ctx, const qil_mps* psi, const uint8_t* d_bits, int64_t B, void* d_out) {
    int chimax = 1;
    for (int i = 0; i <= psi->n; ++i) chimax = max(chimax, (int)psi->bond[i]);
    // odd multiple of 16 bytes per row to spread rows over banks
    int chi_pad = chimax | 1;
    // smallest number of string groups over all sites bounds the tile size
    int sg_min = 1 << 30;
    for (int i = 0; i < psi->n; ++i) {
        int cr = (int)psi->bond[i + 1];
        int rl = std::min(32, pow2_ceil(cr));
        int rw = (rl == 32) ? std::min(kCoeffWarps, pow2_ceil((cr + 31) / 32)) : 1;
        sg_min = std::min(sg_min, (32 / rl) * (kCoeffWarps / rw));
    }
    long long S = (long long)kST * sg_min;
    S = std::min<long long>(S, 512);
    const size_t budget = std::min<size_t>(ctx->smem_optin, 200 * 1024);
    auto smem_for = [&](long long s) { return (size_t)2 * s * chi_pad * sizeof(T) + (size_t)s * psi->n + 16; };
    while (S > 1 && smem_for(S) > budget) S >>= 1;
    QIL_REQUIRE(smem_for(S) <= budget, QIL_ERR_UNSUPPORTED,
                "coefficient: bond dimension %d does not fit the shared-memory chain kernel", chimax);
    S = std::min<long long>(S, std::max<long long>(1, B));
    const size_t smem = smem_for(S);
    auto kern = coeff_chain_kernel<T>;
    ensure_dynamic_smem(kern, smem);
    int occ = 1;
    QIL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kCoeffThreads, smem));
    occ = std::max(occ, 1);
    long long ntiles = (B + S - 1) / S;
    int grid = (int)std::min<long long>(ntiles, (long long)ctx->sm_count * occ);
    if (grid < 1) grid = 1;
    kern<<<grid, kCoeffThreads, smem, ctx->stream>>>(make_desc(psi), d_bits, (long long)B,
                                                     reinterpret_cast<T*>(d_out), psi->amplitude, (int)S,
                                                     chi_pad);
    QIL_LAUNCH_CHECK(ctx);
}

// ------------------------------------------------------------------------------------------------
// Large-bond variant (complex MPS, e.g. the output of a zT apply with chi ~ 100..600): per site the
// update of a tile of S strings is a GEMM  V'[S x chi_r] = V[S x chi_l] * A_i[:, b, :]  whose B operand
// depends on the bit of each string.  The strings of a CTA tile are regrouped by bit at every site
// (row lists, 16-row DMMA tiles never mix bits), V lives in an L2-resident scratch, and the complex
// product runs on the FP64 tensor path through the real 2x2 embedding
//     [Vr Vi] * [[Br Bi], [-Bi Br]]     (8*S*chi_l*chi_r flop, same as the complex MAC count).
// Operand tiles are staged with cp.async into an XOR-swizzled 3-stage ring.
// ------------------------------------------------------------------------------------------------
constexpr int kCgS = 112;        // strings per CTA tile: ceil(n0/16)+ceil(n1/16) <= 8 for any split
constexpr int kCgRows = 128;     // 8 DMMA m-tiles
constexpr int kCgKc = 16;        // complex k per chunk (32 real)
constexpr int kCgNc = 64;        // complex n per pass (128 real, 16 n8 tiles)
constexpr int kCgStages = 3;
constexpr int kCgABytes = kCgRows * kCgKc * 16;        // 32 KB
constexpr int kCgBBytes = 2 * kCgKc * kCgNc * 16;      // 32 KB (both bit slices)

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// mbarrier helpers for the full / empty ring of coeff_gemm_kernel<NW, true>
__device__ __forceinline__ unsigned cg_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cg_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cg_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void cg_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cg_smem_u32(bar)) : "memory");
}
// arrival that fires when all cp.async issued so far by this thread have landed (counts as one expected arrival)
__device__ __forceinline__ void cg_cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(cg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cg_mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CG_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CG_DONE;\n"
        "bra CG_WAIT;\n"
        "CG_DONE:\n"
        "}\n" ::"r"(cg_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void dmma16816c(double* c, const double* a, const double* b) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
        "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
          "d"(b[2]), "d"(b[3]));
}

// NW warps: warp w works on DMMA row tile w % 8 and on column groups [(w / 8) * NG, (w / 8 + 1) * NG) of every pass,
// NG = 64 / NW (8 warps: all 8 groups, 255 registers; 16 warps: 4 groups each, 128 registers, twice the warps per
// scheduler to cover the load / barrier bubbles of the tensor pipe)
// MB: the stages of the ring are handed over through mbarriers (every thread's cp.async arrive on full[stage], every warp
// arrives on empty[stage] after its last read) instead of cp.async.wait_group + __syncthreads() per k-chunk, so warps
// drift against each other by up to one chunk and no warp waits for the slowest one at every chunk.
template <int NW, bool MB>
__global__ void __launch_bounds__(NW * 32, 1)
coeff_gemm_kernel(const ChainDesc d, const uint8_t* __restrict__ bits, long long B, cplx* __restrict__ out,
                  double amplitude, cplx* __restrict__ vscratch, int chi_pad) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sA = smem_raw;                                        // stages x 32 KB
    unsigned char* sB = sA + kCgStages * kCgABytes;                      // stages x 32 KB
    uint8_t* sbits = sB + kCgStages * kCgBBytes;                         // [kCgS][n]
    int* rowmap = reinterpret_cast<int*>(sbits + ((kCgS * d.n + 15) & ~15));  // [128] string of each tile row
    int* tilebit = rowmap + kCgRows;                                     // [8]
    __shared__ int s_cnt[2];
    __shared__ uint64_t s_full[kCgStages], s_empty[kCgStages];

    constexpr int THREADS = NW * 32;
    constexpr int NG = 64 / NW;                 // column groups per warp and pass
    const int n = d.n;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mtile = warp & 7, gq0 = (warp >> 3) * NG;
    const int g = lane >> 2, t = lane & 3;
    cplx* V0 = vscratch + (size_t)blockIdx.x * 2 * kCgS * chi_pad;
    cplx* V1 = V0 + (size_t)kCgS * chi_pad;
    const long long ntiles = (B + kCgS - 1) / kCgS;
    if (MB) {
        if (threadIdx.x == 0) {
            for (int st = 0; st < kCgStages; ++st) { cg_mbar_init(&s_full[st], NW * 32); cg_mbar_init(&s_empty[st], NW); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    unsigned gprod = 0, gcons = 0;        // ring positions (chunks issued / consumed so far), identical in every thread

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile * kCgS;
        const int ns = (int)min((long long)kCgS, B - s0);
        __syncthreads();
        for (int idx = tid; idx < kCgS * n; idx += THREADS)
            sbits[idx] = (idx < ns * n) ? bits[s0 * n + idx] : (uint8_t)0;
        for (int s = tid; s < kCgS; s += THREADS) V0[(size_t)s * chi_pad] = make_double2(1.0, 0.0);
        __syncthreads();

        cplx* Vc = V0;
        cplx* Vn = V1;
        for (int i = 0; i < n; ++i) {
            const int cl = d.bond[i], cr = d.bond[i + 1];
            const cplx* __restrict__ M = reinterpret_cast<const cplx*>(d.core[i]);
            // ---- regroup the strings of the tile by their bit at this site (tiles never mix bits)
            if (tid < kCgRows) rowmap[tid] = -1;
            if (tid == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
            __syncthreads();
            int enc = -1;   // position inside the bit group | bit << 16, for string `tid`
            if (warp < 4) {
                const int bit = (tid < kCgS) ? (sbits[tid * n + i] & 1) : -1;
                const unsigned m0 = __ballot_sync(0xffffffffu, bit == 0);
                const unsigned m1 = __ballot_sync(0xffffffffu, bit == 1);
                int base0 = 0, base1 = 0;
                if (lane == 0) {
                    base0 = atomicAdd(&s_cnt[0], __popc(m0));
                    base1 = atomicAdd(&s_cnt[1], __popc(m1));
                }
                base0 = __shfl_sync(0xffffffffu, base0, 0);
                base1 = __shfl_sync(0xffffffffu, base1, 0);
                const unsigned below = (1u << lane) - 1u;
                if (bit == 0) enc = base0 + __popc(m0 & below);
                if (bit == 1) enc = (base1 + __popc(m1 & below)) | (1 << 16);
            }
            __syncthreads();
            {
                const int t0 = (s_cnt[0] + 15) >> 4;           // DMMA tiles holding bit-0 strings
                if (enc >= 0) {
                    const int bit = enc >> 16, pos = enc & 0xffff;
                    rowmap[bit == 0 ? pos : t0 * 16 + pos] = tid;
                }
                if (tid < 8) tilebit[tid] = (tid < t0) ? 0 : 1;
            }
            __syncthreads();

            const int kchunks = (cl + kCgKc - 1) / kCgKc;
            const int npass = (cr + kCgNc - 1) / kCgNc;
            const int iters = kchunks * npass;
            const int mybit = tilebit[mtile];
            const int R0 = mtile * 16 + g, R1 = R0 + 8;
            const int str0 = rowmap[R0], str1 = rowmap[R1];

            // one of the 16 cp.async of a stage per thread: pieces 0..7 = A (128 rows x 16 complex), 8..15 = B ((bit, 16 l) x 64)
            auto issue_piece = [&](int it, int e) {
                const int stage = MB ? (int)(gprod % kCgStages) : it % kCgStages;
                const int pass = it / kchunks, kc = it - pass * kchunks;
                if (e < 8) {
                    unsigned char* a = sA + stage * kCgABytes;
                    const int idx = e * THREADS + tid;
                    const int row = idx >> 4, j = idx & 15;
                    const int sidx = rowmap[row];
                    const int l = kc * kCgKc + j;
                    const bool ok = (sidx >= 0) && (l < cl);
                    const cplx* src = ok ? (Vc + (size_t)sidx * chi_pad + l) : Vc;
                    cp_async16(a + row * 256 + ((j ^ (row & 7)) << 4), src, ok);
                } else {
                    unsigned char* b = sB + stage * kCgBBytes;
                    const int idx = (e - 8) * THREADS + tid;
                    const int rowb = idx >> 6, c = idx & 63;        // rowb = bit*16 + lr
                    const int bit = rowb >> 4, lr = rowb & 15;
                    const int l = kc * kCgKc + lr, r = pass * kCgNc + c;
                    const bool ok = (l < cl) && (r < cr);
                    const cplx* src = ok ? (M + ((size_t)l * 2 + bit) * cr + r) : M;
                    cp_async16(b + rowb * 1024 + ((c ^ (((lr >> 1) & 3) << 1)) << 4), src, ok);
                }
            };
            auto issue = [&](int it) {
                if (MB) cg_mbar_wait(&s_empty[gprod % kCgStages], ((gprod / kCgStages) & 1) ^ 1);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (e * THREADS < kCgRows * kCgKc) issue_piece(it, e);
                    if (e * THREADS < 2 * kCgKc * kCgNc) issue_piece(it, 8 + e);
                }
                if (MB) { cg_cp_async_arrive(&s_full[gprod % kCgStages]); ++gprod; }
            };

            // prologue
            for (int it = 0; it < kCgStages - 1; ++it) {
                if (it < iters) issue(it);
                if (!MB) cp_async_commit();
            }
            // One pass = 64 complex output columns = 8 groups of 8; a group is TWO n8 DMMA tiles, the real parts and
            // the imaginary parts of its 8 columns:
            //   re = sum_l Vr Br - Vi Bi = (Vr, -Vi) . (Br, Bi)     A fragment with negated imaginary entries, B as loaded
            //   im = sum_l Vr Bi + Vi Br = (Vr,  Vi) . (Bi, Br)     plain A fragment, B with re / im swapped (register names)
            // so one pair of 16-byte loads (B[l0][c], B[l1][c], c = group column of this lane) feeds two DMMAs, and the
            // only sign flips are the four imaginary A entries per k-block (not one select + negation per B element).
            double accr[NG][4], acci[NG][4];
            for (int it = 0; it < iters; ++it) {
                const int pass = it / kchunks, kc = it - pass * kchunks;
                if (kc == 0) {
#pragma unroll
                    for (int x = 0; x < NG; ++x) {
                        accr[x][0] = accr[x][1] = accr[x][2] = accr[x][3] = 0.0;
                        acci[x][0] = acci[x][1] = acci[x][2] = acci[x][3] = 0.0;
                    }
                }
                const bool more = (it + kCgStages - 1 < iters);
                int stage;
                if (MB) {
                    stage = (int)(gcons % kCgStages);
                    cg_mbar_wait(&s_full[stage], (gcons / kCgStages) & 1);
                } else {
                    cp_async_wait<kCgStages - 2>();
                    __syncthreads();
                    if (more) issue(it + kCgStages - 1);
                    stage = it % kCgStages;
                }
                const unsigned char* a = sA + stage * kCgABytes;
                const unsigned char* b = sB + stage * kCgBBytes + mybit * (kCgKc * 1024);
                const int ngroups = min(8, (cr - pass * kCgNc + 7) >> 3);       // column groups that exist in this pass
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    double af[8], afn[8];
                    const int j0 = kb * 8 + 2 * t;
                    const double2 x00 = *reinterpret_cast<const double2*>(a + R0 * 256 + (((j0) ^ (R0 & 7)) << 4));
                    const double2 x01 = *reinterpret_cast<const double2*>(a + R0 * 256 + (((j0 + 1) ^ (R0 & 7)) << 4));
                    const double2 x10 = *reinterpret_cast<const double2*>(a + R1 * 256 + (((j0) ^ (R1 & 7)) << 4));
                    const double2 x11 = *reinterpret_cast<const double2*>(a + R1 * 256 + (((j0 + 1) ^ (R1 & 7)) << 4));
                    af[0] = x00.x; af[2] = x00.y; af[4] = x01.x; af[6] = x01.y;
                    af[1] = x10.x; af[3] = x10.y; af[5] = x11.x; af[7] = x11.y;
                    afn[0] = af[0]; afn[1] = af[1]; afn[4] = af[4]; afn[5] = af[5];
                    afn[2] = -af[2]; afn[3] = -af[3]; afn[6] = -af[6]; afn[7] = -af[7];
                    const int lr0 = kb * 8 + 2 * t, lr1 = lr0 + 1;
                    const unsigned char* b0p = b + lr0 * 1024;
                    const unsigned char* b1p = b + lr1 * 1024;
                    const int sw0 = ((lr0 >> 1) & 3) << 1, sw1 = ((lr1 >> 1) & 3) << 1;
                    // B elements of the next group are loaded before the DMMAs of this one are issued
                    double2 e0n = *reinterpret_cast<const double2*>(b0p + (((gq0 * 8 + g) ^ sw0) << 4));
                    double2 e1n = *reinterpret_cast<const double2*>(b1p + (((gq0 * 8 + g) ^ sw1) << 4));
#pragma unroll
                    for (int gq = 0; gq < NG; ++gq) {
                        const double2 e0 = e0n, e1 = e1n;
                        if (gq + 1 < NG) {
                            const int cn = (gq0 + gq + 1) * 8 + g;
                            e0n = *reinterpret_cast<const double2*>(b0p + ((cn ^ sw0) << 4));
                            e1n = *reinterpret_cast<const double2*>(b1p + ((cn ^ sw1) << 4));
                        }
                        if (gq0 + gq < ngroups) {
                            const double br[4] = {e0.x, e0.y, e1.x, e1.y};
                            const double bi[4] = {e0.y, e0.x, e1.y, e1.x};
                            dmma16816c(accr[gq], afn, br);
                            dmma16816c(acci[gq], af, bi);
                        }
                    }
                }
                if (MB) {
                    __syncwarp();
                    if (lane == 0) cg_mbar_arrive(&s_empty[stage]);     // this warp is done reading the stage
                    ++gcons;
                    if (more) issue(it + kCgStages - 1);                 // the stage of chunk it-1: long released by everybody
                } else {
                    cp_async_commit();
                }
                if (kc == kchunks - 1) {
                    // epilogue of this pass: complex columns pass*64 + gq*8 + {2t, 2t+1} of rows R0 / R1
#pragma unroll
                    for (int gq = 0; gq < NG; ++gq) {
                        const int r = pass * kCgNc + (gq0 + gq) * 8 + 2 * t;
                        if (r < cr) {
                            if (str0 >= 0) Vn[(size_t)str0 * chi_pad + r] = make_double2(accr[gq][0], acci[gq][0]);
                            if (str1 >= 0) Vn[(size_t)str1 * chi_pad + r] = make_double2(accr[gq][2], acci[gq][2]);
                        }
                        if (r + 1 < cr) {
                            if (str0 >= 0) Vn[(size_t)str0 * chi_pad + r + 1] = make_double2(accr[gq][1], acci[gq][1]);
                            if (str1 >= 0) Vn[(size_t)str1 * chi_pad + r + 1] = make_double2(accr[gq][3], acci[gq][3]);
                        }
                    }
                }
            }
            cp_async_wait<0>();
            __syncthreads();
            cplx* tmp = Vc; Vc = Vn; Vn = tmp;
        }
        for (int s = tid; s < ns; s += THREADS) {
            const cplx v = Vc[(size_t)s * chi_pad];
            out[s0 + s] = make_double2(v.x * amplitude, v.y * amplitude);
        }
    }
}

static void launch_coeff_gemm(qil_ctx* ctx, const qil_mps* psi, const uint8_t* d_bits, int64_t B, void* d_out) {
    int chimax = 1;
    for (int i = 0; i <= psi->n; ++i) chimax = max(chimax, (int)psi->bond[i]);
    const int chi_pad = (chimax + 3) & ~3;
    const long long ntiles = (B + kCgS - 1) / kCgS;
    const int grid = (int)std::min<long long>(ntiles, ctx->sm_count);
    const size_t smem = (size_t)kCgStages * (kCgABytes + kCgBBytes) + (((size_t)kCgS * psi->n + 15) & ~(size_t)15) +
                        (kCgRows + 8) * sizeof(int) + 64;
    QIL_REQUIRE(smem <= ctx->smem_optin, QIL_ERR_UNSUPPORTED, "coefficient: chain of %d sites does not fit", psi->n);
    cplx* scratch = (cplx*)ctx->alloc((size_t)grid * 2 * kCgS * chi_pad * sizeof(cplx));
    static const int nw = [] { const char* e = getenv("QIL_COEFF_WARPS"); return (e && atoi(e) == 8) ? 8 : 16; }();
    static const bool mb = [] { const char* e = getenv("QIL_COEFF_MBAR"); return !(e && e[0] == '0'); }();
    auto kern = (nw == 16) ? (mb ? coeff_gemm_kernel<16, true> : coeff_gemm_kernel<16, false>)
                           : (mb ? coeff_gemm_kernel<8, true> : coeff_gemm_kernel<8, false>);
    ensure_dynamic_smem(kern, smem);
    kern<<<grid, nw * 32, smem, ctx->stream>>>(make_desc(psi), d_bits, (long long)B, reinterpret_cast<cplx*>(d_out),
                                                  psi->amplitude, scratch, chi_pad);
    QIL_LAUNCH_CHECK(ctx);
    ctx->free(scratch);
}

void coefficient_batch_dev(qil_ctx* ctx, const qil_mps* psi, const uint8_t* d_bits, int64_t B, void* d_out) {
    if (B <= 0) return;
    int chimax = 1;
    for (int i = 0; i <= psi->n; ++i) chimax = max(chimax, (int)psi->bond[i]);
    { qil_prof_region prof_guard_(ctx, PROF_COEFF);
    if (psi->is_complex && chimax >= 48 && B >= 64)
        launch_coeff_gemm(ctx, psi, d_bits, B, d_out);
    else if (psi->is_complex)
        launch_coeff<cplx>(ctx, psi, d_bits, B, d_out);
    else
        launch_coeff<double>(ctx, psi, d_bits, B, d_out);
    }
}

}  // namespace qil
