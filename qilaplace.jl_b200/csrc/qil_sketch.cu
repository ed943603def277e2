// qil_sketch.cu -- K1 / K2: the streaming FP64 GEMMs of the randomized-SVD encoder (rsvd.jl:79,89,92,98).
//
//   K1 (TRANS = false):  Y[R x L] = A[R x C] * X[C x L]        sketch  Y = A*Omega, power step Y = A*Q_Z
//   K2 (TRANS = true ):  Z[C x L] = A[R x C]^T * X[R x L]      Z = A^H*Q and B^H = A^H*Q
//
// A is the signal itself viewed as a row-major matrix (MSB-first layout makes every divide-and-conquer
// reshape a free contiguous view), so these kernels are the only full passes over the 2^n samples.
// L = k+p is narrow (20..112): ~l/4 flop per byte, i.e. right at the B200 FP64 ridge
// (37 TFLOP/s DMMA == DFMA measured, profiles/r01_ubench_fp64.txt; ~7 TB/s read), so the kernel has to
// stream at HBM rate and keep the FP64 pipe busy at the same time:
//   * persistent CTAs, one producer warp issuing TMA (cp.async.bulk.tensor, 128B swizzle) for the A tile
//     and a bulk copy for the X tile into a 3-4 stage mbarrier ring,
//   * 8 consumer warps, each owning 16 output rows x all L columns in registers, issuing
//     mma.sync.m16n8k16.f64 (DMMA) straight from the swizzled tile,
//   * split-K over the long dimension so that the tile count is >> 148 SMs; partials are summed by a small
//     deterministic reduce kernel (no atomics -> bit-reproducible ranks),
//   * sum(x^2) for the amplitude (SignalConverters.jl:36) is accumulated from the A fragments of the first
//     pass, so normalisation costs no extra pass (the 1/c scale is folded into the small matrix B).
// Complex signals reuse the same real kernels on the interleaved view (R x 2C) with an expanded X.
#include "qil_mpsops.cuh"

#include <cstdlib>
#include <cuda.h>
#include <cudaTypedefs.h>

namespace qil {

constexpr int kBM = 128;          // output rows per tile (8 consumer warps x 16)
constexpr int kBK = 32;           // granularity of the K chunks and of the zero padding of X (a stage holds BK = 32 or 16)
constexpr int kConsumerWarps = 8;
constexpr int kStreamThreads = (kConsumerWarps + 1) * 32;

// ---- PTX helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
        "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
          "d"(b[2]), "d"(b[3]));
}

struct StreamParams {
    long long Mtot;      // output rows: R (K1) or C (K2)
    long long Kdim;      // reduction length: C (K1) or R (K2)
    int tilesM;          // ceil(Mtot / 128)
    int ksplit;          // number of K chunks
    long long kchunk;    // chunk length (multiple of 32)
    int lpp;             // pitch of X in doubles (>= 8*NT)
    int ldo;             // pitch of the output in doubles (8*NT)
    double* out;         // [ksplit][Mtot][ldo]
    const double* X;     // [ceil32(Kdim)][lpp]
    double* sumsq;       // optional [gridDim.x] partial sums of A^2 (K1 only)
    long long x_rows;    // K1 on a stack of equal matrices (batch of signals): output rows per matrix, 0 = one shared X
    long long x_bs;      // ... and the distance (doubles) between their X operands
};

// BK: reduction depth of one pipeline stage (A tile of 128 x BK doubles); CTAS: resident CTAs per SM the launch is sized for
template <int NT, bool TRANS, int STAGES, int BK, int CTAS>
__global__ void __launch_bounds__(kStreamThreads, CTAS)
stream_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const StreamParams p) {
    constexpr int kStageABytes = kBM * BK * 8;   // 32 KB / 16 KB
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 128B-swizzled TMA boxes need a 1024-byte aligned base; do not rely on the toolchain for that
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int xbytes = BK * p.lpp * 8;
    const int xstride = (xbytes + 127) & ~127;
    unsigned char* sA = smem;                                   // STAGES x 32 KB (1024-aligned)
    unsigned char* sX = smem + (size_t)STAGES * kStageABytes;   // STAGES x xstride
    uint64_t* full = reinterpret_cast<uint64_t*>(sX + (size_t)STAGES * xstride);
    uint64_t* empty = full + STAGES;
    double* red = reinterpret_cast<double*>(empty + STAGES);    // [kConsumerWarps]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long ntiles = (long long)p.tilesM * p.ksplit;
    if (warp == kConsumerWarps) {
        // ================= producer warp: one elected lane drives TMA =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int tm = (int)(tile % p.tilesM);
                const int ks = (int)(tile / p.tilesM);
                const long long k0 = (long long)ks * p.kchunk;
                const long long k1 = min(k0 + p.kchunk, p.Kdim);
                for (long long k = k0; k < k1; k += BK) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], kStageABytes + xbytes);
                    unsigned char* a = sA + (size_t)stage * kStageABytes;
                    if (!TRANS) {
                        // BK/16 boxes [128 rows][16 cols]
#pragma unroll
                        for (int b = 0; b < BK / 16; ++b)
                            tma_load_2d(a + b * 16384, &tmA, &full[stage], (int)k + 16 * b, tm * kBM);
                    } else {
                        // eight boxes [BK rows][16 cols], one per consumer warp
#pragma unroll
                        for (int w = 0; w < 8; ++w)
                            tma_load_2d(a + w * (BK * 128), &tmA, &full[stage], tm * kBM + w * 16, (int)k);
                    }
                    const double* xsrc = p.X + k * p.lpp;
                    if (!TRANS && p.x_rows) xsrc += ((long long)tm * kBM / p.x_rows) * p.x_bs;
                    bulk_load(sX + (size_t)stage * xstride, xsrc, (uint32_t)xbytes, &full[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ================= consumer warps: DMMA from the swizzled tiles =================
        const int g = lane >> 2, t = lane & 3;
        int stage = 0;
        uint32_t phase = 0;
        double ssq = 0.0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int tm = (int)(tile % p.tilesM);
            const int ks = (int)(tile / p.tilesM);
            const long long k0 = (long long)ks * p.kchunk;
            const long long k1 = min(k0 + p.kchunk, p.Kdim);
            double acc[NT][4];
#pragma unroll
            for (int i = 0; i < NT; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0; }
            for (long long k = k0; k < k1; k += BK) {
                mbar_wait(&full[stage], phase);
                const unsigned char* a = sA + (size_t)stage * kStageABytes;
                const double* xs = reinterpret_cast<const double*>(sX + (size_t)stage * xstride);
#pragma unroll
                for (int kb = 0; kb < BK / 16; ++kb) {
                    double af[8];
                    if (!TRANS) {
                        // box kb: rows = output rows, 16 k-columns; this thread owns k = 4t..4t+3 of rows g, g+8
                        const unsigned char* box = a + kb * 16384;
                        const int r0 = warp * 16 + g, r1 = r0 + 8;
                        const double2 v00 = *reinterpret_cast<const double2*>(box + r0 * 128 + (((2 * t) ^ (r0 & 7)) << 4));
                        const double2 v01 = *reinterpret_cast<const double2*>(box + r0 * 128 + (((2 * t + 1) ^ (r0 & 7)) << 4));
                        const double2 v10 = *reinterpret_cast<const double2*>(box + r1 * 128 + (((2 * t) ^ (r1 & 7)) << 4));
                        const double2 v11 = *reinterpret_cast<const double2*>(box + r1 * 128 + (((2 * t + 1) ^ (r1 & 7)) << 4));
                        af[0] = v00.x; af[2] = v00.y; af[4] = v01.x; af[6] = v01.y;
                        af[1] = v10.x; af[3] = v10.y; af[5] = v11.x; af[7] = v11.y;
                    } else {
                        // box `warp`: BK k-rows x 16 output columns, columns g and g+8.  The sum over k does not care which
                        // k-row a thread takes as long as A and B agree: element i of thread t is k-row 2t + (i&1) + 8(i>>1)
                        // instead of the fragment's natural 4t + i, so that the four rows one load touches in a half-warp
                        // (t = 0..3) sit in four different phases of the 128B swizzle -> no bank conflicts (the natural order
                        // put t and t+2 on the same banks: 29 % of this pass's shared-memory wavefronts were replays)
                        const unsigned char* box = a + warp * (BK * 128);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int row = kb * 16 + 2 * t + (i & 1) + 8 * (i >> 1);
                            const unsigned char* rp = box + row * 128;
                            af[2 * i] = *reinterpret_cast<const double*>(rp + ((((g) >> 1) ^ (row & 7)) << 4) + ((g & 1) << 3));
                            af[2 * i + 1] = *reinterpret_cast<const double*>(rp + ((((g + 8) >> 1) ^ (row & 7)) << 4) + ((g & 1) << 3));
                        }
                    }
                    if (p.sumsq) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) ssq = fma(af[i], af[i], ssq);
                    }
                    // B fragment: k-rows 4t + i (K1) or the permuted rows of the A fragment above (K2); the pitch of X
                    // (stream_lpp) spreads either set over distinct banks
                    const double* xr = xs + (size_t)(kb * 16 + (TRANS ? 2 : 4) * t) * p.lpp + g;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        double bf[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) bf[i] = xr[(size_t)(TRANS ? ((i & 1) + 8 * (i >> 1)) : i) * p.lpp + nt * 8];
                        dmma16816(acc[nt], af, bf);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            // epilogue: rows tm*128 + warp*16 + {g, g+8}, columns nt*8 + 2t, 2t+1
            const long long row0 = (long long)tm * kBM + warp * 16 + g;
            double* o = p.out + ((long long)ks * p.Mtot) * p.ldo;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                if (row0 < p.Mtot)
                    *reinterpret_cast<double2*>(o + row0 * p.ldo + nt * 8 + 2 * t) = make_double2(acc[nt][0], acc[nt][1]);
                if (row0 + 8 < p.Mtot)
                    *reinterpret_cast<double2*>(o + (row0 + 8) * p.ldo + nt * 8 + 2 * t) = make_double2(acc[nt][2], acc[nt][3]);
            }
        }
        if (p.sumsq) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
            if (lane == 0) red[warp] = ssq;
        }
    }
    __syncthreads();
    if (p.sumsq && threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < kConsumerWarps; ++w) s += red[w];
        p.sumsq[blockIdx.x] = s;
    }
}

// ---- tensor map creation through the driver entry point (no link-time libcuda dependency) ------------
static PFN_cuTensorMapEncodeTiled get_encode() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        QIL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        QIL_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, QIL_ERR_CUDA,
                    "cuTensorMapEncodeTiled is not available from the driver");
        fn = (PFN_cuTensorMapEncodeTiled)p;
    }
    return fn;
}

static CUtensorMap make_tmap(const double* A, long long R, long long C, long long ld, int box_cols, int box_rows) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)A, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    QIL_REQUIRE(r == CUDA_SUCCESS, QIL_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    return tm;
}

// ---- launch configuration ------------------------------------------------------------------------------
// Narrow sketches (NT <= 4) run two CTAs per SM (one CTA's epilogue / barrier bubbles are covered by the other's DMMA
// work; measured 2.77 vs 2.83 ms for the six n=28 passes against one CTA with a 4-stage ring).  QIL_STREAM_VARIANT picks
// the ring of that mode:  2 (default at three column tiles) five 16-deep stages | 1: four 16-deep stages | 0 (default
// otherwise): two 32-deep stages | 3: one CTA per SM with five 32-deep stages | 4 (or QIL_STREAM_STAGES=4): one CTA per SM with a
// 4-stage ring, which wide sketches (NT > 4) always use (3 stages beyond NT = 8).  With 16-deep stages a released stage is
// needed again only 4 stages later instead of 1, for about the same shared memory: six n=28 passes at l=20 take
// 2.463 ms (variant 2), 2.477 (1), 2.517 (0), 2.512 (3 and 4) -- profiles/r02_stream_ring_variants.txt.  The pass is bound by the
// FP64 pipe (profiles/r02_stream_gemm_ncu_full.txt), so the ring depth moves it by 2 % only.
static int stream_variant(int nt) {
    static const int v = [] {
        const char* e = getenv("QIL_STREAM_VARIANT");
        if (e) return atoi(e);
        const char* st = getenv("QIL_STREAM_STAGES");
        return (st && atoi(st) == 4) ? 4 : -1;
    }();
    // default: the deep ring only where it was measured to win -- three column tiles (l = 17..24, the co-bound case of the
    // reference's k+p = 20).  One tile (HBM-bound, 2 KB TMA boxes of the transposed pass cost more than the ring gains) and
    // four tiles on the short matrices of the C2 batch were 3-5 % slower with it (adaptive leg 3.45 -> 3.55 ms, C2 4.92 -> 5.10 ms).
    return v >= 0 ? v : (nt == 3 ? 2 : 0);
}
static int stream_ctas_per_sm(int nt) {
    const int v = stream_variant(nt);
    return (nt <= 4 && v >= 0 && v <= 2) ? 2 : 1;
}
static int stream_bk(int nt) {
    const int v = stream_variant(nt);
    return (nt <= 4 && (v == 1 || v == 2)) ? 16 : 32;
}

template <int NT, bool TRANS, int STAGES, int BK, int CTAS>
static void launch_stream_s(qil_ctx* ctx, const CUtensorMap& tm, const StreamParams& p) {
    const int xbytes = BK * p.lpp * 8;
    const int xstride = (xbytes + 127) & ~127;
    const size_t smem = (size_t)STAGES * (kBM * BK * 8 + xstride) + 2 * STAGES * 8 + kConsumerWarps * 8 + 1024;
    auto kern = stream_gemm_kernel<NT, TRANS, STAGES, BK, CTAS>;
    ensure_dynamic_smem(kern, smem);
    const long long ntiles = (long long)p.tilesM * p.ksplit;
    const int grid = (int)std::min<long long>(ntiles, (long long)ctx->sm_count * CTAS);
    kern<<<grid, kStreamThreads, smem, ctx->stream>>>(tm, p);
    QIL_LAUNCH_CHECK(ctx);
}

template <int NT, bool TRANS>
static void launch_stream(qil_ctx* ctx, const CUtensorMap& tm, const StreamParams& p) {
    constexpr int NN = (NT <= 4 ? NT : 1);            // the narrow-sketch variants are instantiated for NT <= 4 only
    if (NT <= 4) {
        switch (stream_variant(NT)) {
            case 0: launch_stream_s<NN, TRANS, 2, 32, 2>(ctx, tm, p); return;
            case 1: launch_stream_s<NN, TRANS, 4, 16, 2>(ctx, tm, p); return;
            case 2: launch_stream_s<NN, TRANS, 5, 16, 2>(ctx, tm, p); return;
            case 3: launch_stream_s<NN, TRANS, 5, 32, 1>(ctx, tm, p); return;
            default: break;
        }
    }
    launch_stream_s<NT, TRANS, ((NT <= 8) ? 4 : 3), 32, 1>(ctx, tm, p);
}

template <bool TRANS>
static void dispatch_stream(qil_ctx* ctx, int nt, const CUtensorMap& tm, const StreamParams& p) {
    switch (nt) {
        case 1: launch_stream<1, TRANS>(ctx, tm, p); break;
        case 2: launch_stream<2, TRANS>(ctx, tm, p); break;
        case 3: launch_stream<3, TRANS>(ctx, tm, p); break;
        case 4: launch_stream<4, TRANS>(ctx, tm, p); break;
        case 5: launch_stream<5, TRANS>(ctx, tm, p); break;
        case 6: launch_stream<6, TRANS>(ctx, tm, p); break;
        case 7: launch_stream<7, TRANS>(ctx, tm, p); break;
        case 8: launch_stream<8, TRANS>(ctx, tm, p); break;
        case 10: launch_stream<10, TRANS>(ctx, tm, p); break;
        case 12: launch_stream<12, TRANS>(ctx, tm, p); break;
        case 14: launch_stream<14, TRANS>(ctx, tm, p); break;
        default: QIL_THROW(QIL_ERR_UNSUPPORTED, "stream gemm: %d column tiles not instantiated", nt);
    }
}

int stream_nt_for(int cols) {
    int nt = (cols + 7) / 8;
    if (nt > 8) nt = (nt + 1) & ~1;   // 10, 12, 14
    return nt;
}

bool stream_supported(long long R, long long C, long long ld, int cols) {
    // TMA needs a 16-byte aligned row pitch; small problems go to the generic GEMM
    return (ld % 2 == 0) && R >= 64 && C >= 64 && (R * C) >= (1ll << 13) && stream_nt_for(cols) <= 14;
}

// out[ksplit][Mtot][8*nt] partials of A*X (trans=false) or A^T*X (trans=true); A is a REAL R x C view.
void stream_gemm(qil_ctx* ctx, bool trans, const double* A, long long R, long long C, long long ld, const double* X,
                 int lpp, int nt, double* out, int ksplit, long long kchunk, double* sumsq_partials, int ncols,
                 long long x_rows, long long x_bs) {
    StreamParams p;
    p.x_rows = x_rows;
    p.x_bs = x_bs;
    p.Mtot = trans ? C : R;
    p.Kdim = trans ? R : C;
    p.tilesM = (int)((p.Mtot + kBM - 1) / kBM);
    p.ksplit = ksplit;
    p.kchunk = kchunk;
    p.lpp = lpp;
    p.ldo = nt * 8;
    p.out = out;
    p.X = X;
    p.sumsq = sumsq_partials;
    const CUtensorMap tm = trans ? make_tmap(A, R, C, ld, 16, stream_bk(nt)) : make_tmap(A, R, C, ld, 16, kBM);
    // algorithmic work: one read of the R x C view, 2 flops per element and sketch column
    { qil_prof_region prof_guard_(ctx, PROF_STREAM_GEMM, 8.0 * (double)R * (double)C, 2.0 * (double)R * (double)C * (double)ncols);
    if (!trans) dispatch_stream<false>(ctx, nt, tm, p);
    else dispatch_stream<true>(ctx, nt, tm, p);
    }
}

void stream_plan(qil_ctx* ctx, long long Mtot, long long Kdim, int* ksplit, long long* kchunk, int nt) {
    // Tiles (row tile x K chunk) are dealt round robin to the resident persistent CTAs, so the launch takes
    // ceil(tiles / CTAs) waves: 896 tiles on 296 CTAs are 3.03 waves, i.e. a fourth, almost empty wave (25 % lost).  Among
    // the split factors around ~6 tiles per CTA (chunks of at least 256 along K) pick the one whose last wave is fullest.
    const long long tilesM = (Mtot + kBM - 1) / kBM;
    const long long ctas = (long long)ctx->sm_count * stream_ctas_per_sm(nt);
    const long long maxsplit = std::max<long long>(1, Kdim / 256);
    const long long want = std::max<long long>(1, std::min((ctas * 6 + tilesM - 1) / tilesM, maxsplit));
    long long best = want;
    double best_eff = -1.0;
    for (long long ks = std::max<long long>(1, want / 2); ks <= std::min(maxsplit, 2 * want + 1); ++ks) {
        const long long chunk = ((Kdim + ks - 1) / ks + kBK - 1) / kBK * kBK;
        const long long ks_eff = (Kdim + chunk - 1) / chunk;
        const long long tiles = tilesM * ks_eff;
        const long long waves = (tiles + ctas - 1) / ctas;
        double eff = (double)tiles / (double)(waves * ctas);
        eff -= 0.002 * (double)ks_eff;                       // mild preference for fewer partials at equal balance
        if (tiles < ctas) eff = (double)tiles / (double)ctas - 0.002 * (double)ks_eff;
        if (eff > best_eff) { best_eff = eff; best = ks_eff; }
    }
    long long chunk = ((Kdim + best - 1) / best + kBK - 1) / kBK * kBK;
    *ksplit = (int)((Kdim + chunk - 1) / chunk);
    *kchunk = chunk;
}

int stream_grid(qil_ctx* ctx, long long Mtot, int ksplit, int nt) {
    const long long ntiles = ((Mtot + kBM - 1) / kBM) * ksplit;
    return (int)std::min<long long>(ntiles, (long long)ctx->sm_count * stream_ctas_per_sm(nt));
}

}  // namespace qil
