// qil_rng.cuh -- counter-based normal stream standing in for `Random.seed!(seed); randn(...)` of rsvd.jl:74-76 when the
// host does not supply the stream (the Julia Xoshiro stream cannot be generated here; see DESIGN.md "Oracle").
#pragma once
#include "qil_common.cuh"

namespace qil {

// ---- counter-based N(0,1) stream -------------------------------------------------------------------
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double gauss_at(unsigned long long seed, unsigned long long idx) {
    // Box-Muller on two 53-bit uniforms derived from (seed, idx)
    const unsigned long long h1 = splitmix64(seed * 0xD1342543DE82EF95ull + 2 * idx);
    const unsigned long long h2 = splitmix64(seed * 0xD1342543DE82EF95ull + 2 * idx + 1);
    const double u1 = ((double)(h1 >> 11) + 1.0) * (1.0 / 9007199254740992.0);  // (0,1]
    const double u2 = (double)(h2 >> 11) * (1.0 / 9007199254740992.0);          // [0,1)
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}
template <typename T> __device__ __forceinline__ T stream_at(const T* host_stream, unsigned long long seed, long long i);
template <> __device__ __forceinline__ double stream_at<double>(const double* s, unsigned long long seed, long long i) {
    return s ? s[i] : gauss_at(seed, (unsigned long long)i);
}
template <> __device__ __forceinline__ cplx stream_at<cplx>(const cplx* s, unsigned long long seed, long long i) {
    if (s) return s[i];
    const double f = 0.70710678118654752440;
    return make_double2(f * gauss_at(seed, 2ull * i), f * gauss_at(seed, 2ull * i + 1));
}

}  // namespace qil
