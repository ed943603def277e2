/* qil_oracle_c.c -- plain-C restatement of the scalar rules of the hot path, independent of numpy.
 *
 * TEST INFRASTRUCTURE ONLY (like oracle/qil_oracle.py): nothing under qilaplace.jl_b200/ links or calls it.
 * It is compiled by __graft_entry__.build() / oracle/build_c.py with gcc and cross-checked against the numpy
 * oracle and the golden vectors in tests/test_oracle_golden.py.
 *
 *   qil_ref_truncate_rank   ITensors/NDTensors `truncate!!` as used by every svd(...; cutoff, maxdim, mindim) of the
 *                           path (src/linalg/rsvd.jl:103, src/signals/SignalConverters.jl:84,266, src/mps.jl:929,946):
 *                           on P = sigma^2 (descending) drop from the tail while n > maxdim, then keep dropping
 *                           while discarded + P[n-1] <= cutoff * sum(P) and n > mindim.
 *   qil_ref_bits_from_integer  big-endian configuration of an Integer (src/mps.jl:633-645): site 1 = MSB.
 *   qil_ref_coefficient     <bits|psi> * amplitude by the left-to-right vector chain (src/mps.jl:669-678), cores
 *                           [l][s][r] row-major complex128, stored back to back.
 */
#include <stdint.h>
#include <stdlib.h>

int qil_ref_truncate_rank(const double* sigma, int n, double cutoff, long long maxdim, long long mindim) {
    if (n <= 1) return n;
    int r = n;
    double err = 0.0, scale = 0.0;
    while ((long long)r > maxdim) { err += sigma[r - 1] * sigma[r - 1]; --r; }
    for (int i = 0; i < n; ++i) scale += sigma[i] * sigma[i];
    if (scale == 0.0) scale = 1.0;
    while ((long long)r > mindim && err + sigma[r - 1] * sigma[r - 1] <= cutoff * scale) {
        err += sigma[r - 1] * sigma[r - 1];
        --r;
    }
    return r < 1 ? 1 : r;
}

/* returns 0, or 1 when the value needs more than n bits / is negative (ArgumentError in the reference) */
int qil_ref_bits_from_integer(long long value, int n, uint8_t* bits) {
    if (value < 0) return 1;
    if (n < 63 && (value >> n) != 0) return 1;
    for (int i = 0; i < n; ++i) bits[i] = (uint8_t)((value >> (n - 1 - i)) & 1);
    return 0;
}

/* bond has n+1 entries (bond[0] = bond[n] = 1); cores holds interleaved (re, im) doubles */
int qil_ref_coefficient(int n, const int64_t* bond, const double* cores, const uint8_t* bits, double amplitude,
                        double* out_re_im) {
    int64_t maxb = 1;
    for (int i = 0; i <= n; ++i) if (bond[i] > maxb) maxb = bond[i];
    double* v = (double*)calloc((size_t)4 * maxb, sizeof(double));
    if (!v) return 2;
    double* w = v + 2 * maxb;
    v[0] = 1.0; v[1] = 0.0;
    const double* core = cores;
    for (int i = 0; i < n; ++i) {
        const int64_t cl = bond[i], cr = bond[i + 1];
        if (bits[i] > 1) { free(v); return 1; }
        for (int64_t r = 0; r < cr; ++r) {
            double re = 0.0, im = 0.0;
            for (int64_t l = 0; l < cl; ++l) {
                const double* m = core + 2 * ((l * 2 + bits[i]) * cr + r);
                re += v[2 * l] * m[0] - v[2 * l + 1] * m[1];
                im += v[2 * l] * m[1] + v[2 * l + 1] * m[0];
            }
            w[2 * r] = re; w[2 * r + 1] = im;
        }
        double* t = v; v = w; w = t;
        core += 2 * cl * 2 * cr;
    }
    out_re_im[0] = amplitude * v[0];
    out_re_im[1] = amplitude * v[1];
    free(v < w ? v : w);
    return 0;
}
