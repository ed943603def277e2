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
    extern __shared__ __align__(16) unsigned char smem_raw[];
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
    QIL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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

void coefficient_batch_dev(qil_ctx* ctx, const qil_mps* psi, const uint8_t* d_bits, int64_t B, void* d_out) {
    if (B <= 0) return;
    if (psi->is_complex)
        launch_coeff<cplx>(ctx, psi, d_bits, B, d_out);
    else
        launch_coeff<double>(ctx, psi, d_bits, B, d_out);
}

}  // namespace qil
