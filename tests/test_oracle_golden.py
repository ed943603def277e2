"""Pin the CPU oracle against the reference's own golden vectors.

CPU only.  Every expected value below comes from tests/golden/reference_goldens.json (transcribed
from the reference repository, file:line inside the fixture) or from a closed form the reference's
tests use as their analytic oracle (DFT, analytical_dt, analytical_zt).
"""
import math

import os

import numpy as np
import pytest

import qil_oracle as O


def test_truncation_rule():
    # relative, cumulative, on sigma^2 (NDTensors truncate!!)
    S = np.array([1.0, 1e-3, 1e-6, 1e-9])
    assert O.truncate_rank(S, cutoff=0.0) == 4
    assert O.truncate_rank(S, cutoff=1e-13) == 3      # 1e-18 <= 1e-13 * scale, 1e-12 + 1e-18 is not
    assert O.truncate_rank(S, cutoff=2e-12) == 2
    assert O.truncate_rank(S, cutoff=1.0) == 1        # mindim = 1
    assert O.truncate_rank(S, cutoff=0.0, maxdim=2) == 2
    assert O.truncate_rank(S, cutoff=1.0, mindim=3) == 3
    assert O.truncate_rank(np.array([0.0]), cutoff=1.0) == 1
    assert O.truncate_rank(np.array([1.0, 0.0, 0.0]), cutoff=0.0) == 1  # cutoff=0 drops exact zeros


def test_coefficient_kats(goldens):
    g = goldens["coefficient_kats"]
    x = np.array(g["x"], dtype=float)
    for method, kw in (("svd", {}), ("rsvd", {})):
        cores, c = O.signal_mps(x, method=method, **kw)
        assert abs(c - np.linalg.norm(x)) < 1e-12
        for i in range(8):
            assert abs(O.coefficient(cores, c, O.bits_from_integer(i, 3)) - x[i]) < 1e-12
    # hand-built MPS of test/test_mps.jl:404-427
    A1 = np.zeros((1, 2, 1)); A1[0, 1, 0] = 1.0
    A2 = np.zeros((1, 2, 1)); A2[0, 0, 0] = 1.0
    A3 = np.zeros((1, 2, 1)); A3[0, 1, 0] = 0.5
    assert O.coefficient([A1, A2, A3], 1.0, g["handbuilt_bits"]) == pytest.approx(g["handbuilt_value"], rel=1e-12)
    with pytest.raises(ValueError):
        O.coefficient([A1, A2, A3], 1.0, [1, 0])
    with pytest.raises(ValueError):
        O.coefficient([A1, A2, A3], 1.0, [2, 0, 1])
    with pytest.raises(ValueError):
        O.bits_from_integer(0b1000, 3)
    # mps_to_vector orders (test/test_mps.jl:448-466)
    cores, c = O.tt_svd(x)
    assert np.allclose(O.mps_to_vector(cores, c), x, atol=1e-12)
    rev = np.array([x[O.bitrev(i, 3)] for i in range(8)])
    assert np.allclose(O.mps_to_vector(cores, c, reverse=True), rev, atol=1e-12)


def test_random_roundtrip_svd_and_rsvd():
    # test/test_signal_converters.jl:25-111,131-139: random n=5,6 reconstructions, RSVD vs SVD 1e-10
    rng = np.random.default_rng(7)
    for n in (5, 6):
        for cplx in (False, True):
            x = rng.standard_normal(2**n) + (1j * rng.standard_normal(2**n) if cplx else 0)
            cs, c = O.tt_svd(x)
            assert np.allclose(O.mps_to_vector(cs, c), x, atol=1e-12)
            cr, c2 = O.tt_rsvd(x, k=2 ** (n // 2), p=0)
            assert np.allclose(O.mps_to_vector(cr, c2), x, atol=1e-10)


def test_rsvd_properties():
    # test/test_rsvd.jl:27-120
    rng = np.random.default_rng(3)
    A = rng.standard_normal((100, 10)) @ rng.standard_normal((10, 100))
    U, S, Vh = O.rsvd(A, k=10, p=5)
    assert S.size == 10
    assert np.linalg.norm(U @ np.diag(S) @ Vh - A) / np.linalg.norm(A) < 1e-10
    assert np.allclose(U.T @ U, np.eye(10), atol=1e-10)
    assert np.allclose(Vh @ Vh.T, np.eye(10), atol=1e-10)
    assert np.all(np.diff(S) <= 0) and np.all(S >= 0)
    U, S, Vh = O.rsvd(A, k=10, p=5, maxdim=4)
    assert S.size == 4
    U, S, Vh = O.rsvd(A, k=20, p=5, cutoff=1e-12, maxdim=25)
    assert S.size == 10
    U, S, Vh = O.rsvd(A, k=20, p=5, cutoff=1.0, maxdim=25, mindim=3)
    assert S.size == 3
    U1, S1, V1 = O.rsvd(A, k=10, p=5, random_seed=99)
    U2, S2, V2 = O.rsvd(A, k=10, p=5, random_seed=99)
    assert np.array_equal(S1, S2) and np.array_equal(U1, U2)


def test_qft_bond_series_and_dft(goldens):
    g = goldens["mpo_max_bond_series"]
    for n in range(2, 13):
        W = O.build_qft_mpo(n, cutoff=1e-15, maxdim=O.BIG)
        assert max(O.mpo_bonds(W)) == g["qft"][n - g["n_start"]], n
    # test/test_qft_transformer.jl:331-464: every basis state and a random complex signal, n = 2..5
    rng = np.random.default_rng(11)
    for n in range(2, 6):
        N = 2**n
        W = O.build_qft_mpo(n, cutoff=1e-14, maxdim=1000)
        F = np.exp(-2j * math.pi * np.outer(np.arange(N), np.arange(N)) / N) / math.sqrt(N)
        for j in list(range(N)) + ["rand"]:
            if j == "rand":
                x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
            else:
                x = np.zeros(N); x[j] = 1.0
            cores, c = O.tt_svd(x)
            v = O.mps_to_vector(O.apply_mpo_mps(W, cores), c)
            fn = np.empty(N, dtype=complex)
            for i in range(N):
                fn[O.bitrev(i, n)] = v[i]
            assert np.linalg.norm(fn - F @ x) < 1e-10


def test_dft_tutorial_bonds(goldens):
    g = goldens["dft_tutorial_n4"]
    x = O.generate_signal(g["n"], kind="sin", dt=g["dt"], freq=2 * math.pi)
    cores, c = O.tt_svd(x)
    assert O.bonds_of(cores) == g["mps_bonds"]
    W = O.build_qft_mpo(g["n"], cutoff=g["qft_cutoff"], maxdim=g["qft_maxdim"])
    assert O.mpo_bonds(W) == g["mpo_bonds"]
    out = O.apply_mpo_mps(W, cores)
    assert O.bonds_of(out) == g["product_bonds"]
    v = O.mps_to_vector(out, c)
    F = np.fft.fft(x) / math.sqrt(16)
    assert max(abs(v[i] - F[O.bitrev(i, 4)]) for i in range(16)) < 1e-13


def _analytical_dt_coeff(j, k, N, wr):
    # test/test_dt_transformer.jl:60-92: main register holds k (LSB first), copy keeps j
    return math.exp(-wr * j * k / N) / math.sqrt(N)


def test_dt_bond_series_and_basis_states(goldens):
    g = goldens["mpo_max_bond_series"]
    for n in range(2, 9):
        W = O.build_dt_mpo(n, 2 * math.pi, cutoff=1e-15, maxdim=O.BIG)
        assert max(O.mpo_bonds(W)) == g["dt"][n - g["n_start"]], n
    for n in (1, 2, 3, 4):
        N = 2**n
        for wr in (0.0, 0.75, 1.0, 2.0, 5.0):
            W = O.build_dt_mpo(n, wr)
            for j in range(N):
                x = np.zeros(N); x[j] = 1.0
                cores, c = O.signal_ztmps(x)
                out = O.apply_mpo_mps(W, cores)
                err = 0.0
                for k in range(N):
                    for jj in range(N):
                        got = O.coefficient(out, c, O.interleave(O.bits_lsb(k, n), O.bits_msb(jj, n)))
                        want = _analytical_dt_coeff(j, k, N, wr) if jj == j else 0.0
                        err = max(err, abs(got - want))
                assert err <= 1e-7, (n, wr, j, err)


def test_zt_bond_series_and_basis_states(goldens):
    g = goldens["mpo_max_bond_series"]
    for n in range(2, 8):
        W = O.build_zt_mpo(n, 2 * math.pi, cutoff=1e-15, maxdim=O.BIG)
        assert max(O.mpo_bonds(W)) == g["zt"][n - g["n_start"]], n
    # test/test_zt_transformer.jl:68-110 (err <= 2e-7 in Frobenius norm over the (k,l) grid)
    for n in (1, 2, 3, 4):
        N = 2**n
        for wr in (0.0, 0.75, 1.0, 2.0, 5.0):
            W = O.build_zt_mpo(n, wr)
            for j in range(N):
                x = np.zeros(N); x[j] = 1.0
                cores, c = O.signal_ztmps(x)
                out = O.apply_mpo_mps(W, cores)
                Z = np.array([[O.coefficient(out, c, O.interleave(O.bits_lsb(k, n), O.bits_lsb(l, n)))
                               for l in range(N)] for k in range(N)])
                ref = np.array([[np.exp(-(wr * k + 2j * math.pi * l) / N * j) / N for l in range(N)]
                                for k in range(N)])
                assert np.linalg.norm(Z - ref) <= 2e-7


def test_zt_tutorial_n2_table(goldens):
    g = goldens["zt_tutorial_n2"]
    n, N = 2, 4
    x = np.array([g["a"] ** j * math.cos(g["omega0_over_pi"] * math.pi * j) for j in range(N)])
    cores, c = O.signal_ztmps(x, cutoff=g["encode_cutoff"], maxdim=g["encode_maxdim"])
    # element access sanity check of the tutorial (zt.jl:72-77)
    assert abs(O.coefficient(cores, c, O.interleave(O.bits_msb(2, n), O.bits_msb(2, n))) - x[2]) < 1e-14
    W = O.build_zt_mpo(n, 2 * math.pi * g["omega_r_over_2pi"], cutoff=g["mpo_cutoff"], maxdim=g["mpo_maxdim"])
    assert O.mpo_bonds(W) == g["mpo_bonds"]
    out = O.apply_mpo_mps(W, cores)
    re = np.array(g["chi_5digits_re"]); im = np.array(g["chi_5digits_im"])
    for k in range(N):
        for l in range(N):
            chi = O.coefficient(out, c, O.interleave(O.bits_lsb(k, n), O.bits_lsb(l, n)))
            assert abs(chi.real - re[k, l]) < 6e-6 and abs(chi.imag - im[k, l]) < 6e-6
            ref = sum(x[j] * np.exp(-(2 * math.pi * k + 2j * math.pi * l) / N * j) for j in range(N)) / N
            assert abs(chi - ref) / abs(ref) < 1e-13


def test_zt_tutorial_n20_bond_list(goldens):
    g = goldens["zt_tutorial_n20"]
    N = 2 ** g["n"]
    a = g["a_abs"] * np.exp(1j * g["a_arg"])
    j = np.arange(N)
    x = a**j * np.cos(g["omega0"] * j)
    for kw in (dict(method="rsvd", k=g["k"], p=g["p"], q=g["q"]), dict(method="svd")):
        cores, c = O.signal_ztmps(x, cutoff=g["cutoff"], maxdim=g["maxdim"], **kw)
        b = O.bonds_of(cores)
        assert b[1::2] == g["bonds_main"]
        assert b[0::2] == g["bonds_copy"]
    # RSVD-encoded coefficients reproduce the signal
    idx = [0, 1, 2, 12345, N // 2, N - 1]
    for i in idx:
        bits = O.bits_msb(i, g["n"])
        got = O.coefficient(cores, c, O.interleave(bits, bits))
        assert abs(got - x[i]) < 1e-9 * np.abs(x).max()


def test_signal_tutorial_n10(goldens):
    g = goldens["signal_tutorial_n10"]
    n = g["n"]
    x = O.generate_signal(n, kind="sin_decay", dt=1.0 / 2**n,
                          freq=[2 * math.pi * f for f in g["freq_over_2pi"]],
                          decay_rate=g["decay_rate"], phase=g["phase"])
    cs, c = O.tt_svd(x, cutoff=g["cutoff"], maxdim=g["maxdim"])
    assert max(O.bonds_of(cs)) == g["max_bond_svd"]
    err = np.linalg.norm(O.mps_to_vector(cs, c) - x) / np.linalg.norm(x)
    assert err == pytest.approx(g["rel_err_svd"], rel=0.01)
    cr, c2 = O.tt_rsvd(x, cutoff=g["cutoff"], maxdim=g["maxdim"], k=g["rsvd_k"])
    assert max(O.bonds_of(cr)) == g["max_bond_rsvd"]
    err = np.linalg.norm(O.mps_to_vector(cr, c2) - x) / np.linalg.norm(x)
    assert err == pytest.approx(g["rel_err_rsvd"], rel=0.05)


def test_compress_and_canonicalize():
    # test/test_mps.jl:156-180 (norm preserved by canonicalize), :331-369 (compress! maxdim=2)
    rng = np.random.default_rng(5)
    bonds = [1, 3, 4, 3, 1]
    cores = [rng.standard_normal((bonds[i], 2, bonds[i + 1])) for i in range(4)]
    n0 = O.mps_norm(cores)
    dense = O.mps_to_vector(cores)
    assert abs(n0 - np.linalg.norm(dense)) < 1e-10
    for d in ("right", "left"):
        cc = O.canonicalize(cores, d)
        assert abs(O.mps_norm(cc) - n0) < 1e-10
        assert np.allclose(O.mps_to_vector(cc), dense, atol=1e-10)
    cc, amp = O.compress(cores, 1.0, maxdim=2, tol=1e-8, sweeps=2)
    assert max(O.bonds_of(cc)) <= 2
    assert abs(O.mps_norm(cc) - 1.0) < 1e-10
    cc, amp = O.compress(cores, 1.0)  # lossless at default tol
    assert np.allclose(O.mps_to_vector(cc, amp), dense, atol=1e-10)


def test_apply_vs_dense():
    # test/test_apply.jl: random MPO (bond 2) x random MPS (bond 3) vs dense contraction
    rng = np.random.default_rng(9)
    n = 4
    wb = [1, 2, 2, 2, 1]; pb = [1, 3, 3, 3, 1]
    W = [rng.standard_normal((wb[i], 2, 2, wb[i + 1])) + 1j * rng.standard_normal((wb[i], 2, 2, wb[i + 1]))
         for i in range(n)]
    P = [rng.standard_normal((pb[i], 2, pb[i + 1])) for i in range(n)]
    dense = O.mpo_to_dense(W) @ O.mps_to_vector(P)
    out = O.apply_mpo_mps(W, P)
    assert O.bonds_of(out) == [6, 6, 6]
    assert np.allclose(O.mps_to_vector(out), dense, atol=1e-12)
    W2 = [rng.standard_normal((wb[i], 2, 2, wb[i + 1])) for i in range(n)]
    W12 = O.apply_mpo_mpo(W, W2)  # W acts first
    assert np.allclose(O.mpo_to_dense(W12), O.mpo_to_dense(W2) @ O.mpo_to_dense(W), atol=1e-12)


def test_coefficient_grid_oracle_consistency():
    """The grid helper is nothing but the reference's chain looped over the free sites: all-free == mps_to_vector
    (mps.jl:716-729, both orders), and x = 1..8 reads back through it (test/test_signal_converters.jl:146-191)."""
    x = np.arange(1.0, 9.0)
    cores, c = O.tt_svd(x)
    assert np.allclose(O.coefficient_grid(cores, c, [2, 2, 2]), x, atol=1e-12)
    assert np.allclose(O.coefficient_grid(cores, c, [1, 2, 2]), x[4:], atol=1e-12)
    assert np.allclose(O.coefficient_grid(cores, c, [2, 0, 2]), x[[0, 1, 4, 5]], atol=1e-12)
    assert np.allclose(O.coefficient_grid(cores, c, [2, 2, 2], out_bit=[0, 1, 2]),
                       O.mps_to_vector(cores, c, reverse=True), atol=1e-12)
    rng = np.random.default_rng(3)
    cores = [rng.standard_normal(s) for s in [(1, 2, 3), (3, 2, 4), (4, 2, 2), (2, 2, 1)]]
    assert np.allclose(O.coefficient_grid(cores, 2.0, [2] * 4), O.mps_to_vector(cores, 2.0), atol=1e-13)


def test_pole_scan_modes_host_logic():
    """Host-side grid description of a (k, l) pole-scan block (api.pole_scan_modes) against the reference's own loop
    `coefficient(psi, interleave(lsb(k), lsb(l)))` (docs/src/tutorials/zt.jl:152-157), evaluated by the oracle."""
    import qilaplace_b200 as q
    n = 4
    rng = np.random.default_rng(11)
    bonds = [1, 2, 3, 4, 3, 4, 3, 2, 1]
    cores = [rng.standard_normal((bonds[i], 2, bonds[i + 1])) + 1j * rng.standard_normal((bonds[i], 2, bonds[i + 1]))
             for i in range(2 * n)]

    def direct(ks, ls):
        return np.array([[O.coefficient(cores, 1.0, O.interleave(O.bits_lsb(k, n), O.bits_lsb(l, n))) for l in ls]
                         for k in ks])

    # full table
    mode, ob = q.pole_scan_modes(n, 0, 0, n, n)
    got = O.coefficient_grid(cores, 1.0, mode, ob).reshape(2**n, 2**n)
    assert np.allclose(got, direct(range(2**n), range(2**n)), atol=1e-13)
    # aligned contiguous block k in [8, 12), l in [4, 6)
    mode, ob = q.pole_scan_modes(n, 8, 4, 2, 1)
    got = O.coefficient_grid(cores, 1.0, mode, ob).reshape(4, 2)
    assert np.allclose(got, direct(range(8, 12), range(4, 6)), atol=1e-13)
    # strided (coarse) scan: k = 1 + 2a (origin bit outside the free range), l = 4b
    mode, ob = q.pole_scan_modes(n, 1, 0, 3, 2, stride_log2_k=1, stride_log2_l=2)
    got = O.coefficient_grid(cores, 1.0, mode, ob).reshape(8, 4)
    assert np.allclose(got, direct(range(1, 16, 2), range(0, 16, 4)), atol=1e-13)
    # a misaligned origin is rejected
    with pytest.raises(q.ArgumentError):
        q.pole_scan_modes(n, 2, 0, 2, 2)


def _c_oracle():
    import ctypes as C
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("qil_build_c", os.path.join(root, "oracle", "build_c.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = C.CDLL(mod.build())
    lib.qil_ref_truncate_rank.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_double, C.c_longlong, C.c_longlong]
    lib.qil_ref_bits_from_integer.argtypes = [C.c_longlong, C.c_int, C.POINTER(C.c_uint8)]
    lib.qil_ref_coefficient.argtypes = [C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_uint8),
                                        C.c_double, C.POINTER(C.c_double)]
    return lib, C


def test_plain_c_oracle_agrees_with_numpy_oracle_and_kats():
    """oracle/qil_oracle_c.c (gcc, no numpy) restates the truncation rule, the Integer -> bits convention and the
    coefficient chain; it must agree with the numpy oracle and with the reference's KATs."""
    lib, C = _c_oracle()
    rng = np.random.default_rng(5)
    # truncation rule (NDTensors truncate!!) on random graded spectra and the edge cases of the rule
    for trial in range(200):
        n = int(rng.integers(1, 24))
        s = np.sort(np.abs(rng.standard_normal(n)) * 10.0 ** (-rng.uniform(0, 14, n)))[::-1].copy()
        cutoff = float(10.0 ** (-rng.uniform(0, 26))) if trial % 5 else 0.0
        maxdim = int(rng.integers(1, 30)) if trial % 3 == 0 else O.BIG
        mindim = int(rng.integers(1, 4))
        got = lib.qil_ref_truncate_rank(s.ctypes.data_as(C.POINTER(C.c_double)), n, cutoff,
                                        min(maxdim, 2**62), mindim)
        assert got == O.truncate_rank(s, cutoff, maxdim, mindim), (s, cutoff, maxdim, mindim)
    # Integer configurations are big-endian (mps.jl:633-645)
    bits = (C.c_uint8 * 5)()
    assert lib.qil_ref_bits_from_integer(5, 3, bits) == 0 and list(bits[:3]) == [1, 0, 1] == O.bits_from_integer(5, 3)
    assert lib.qil_ref_bits_from_integer(8, 3, bits) == 1 and lib.qil_ref_bits_from_integer(-1, 3, bits) == 1
    # coefficient chain: x = 1..8 reads back (test/test_signal_converters.jl:146-191) and random complex MPS
    x = np.arange(1.0, 9.0)
    cores, c = O.tt_svd(x)

    def c_coeff(cores, amp, b):
        n = len(cores)
        bond = np.array([1] + [int(k.shape[2]) for k in cores], dtype=np.int64)
        flat = np.concatenate([np.ascontiguousarray(k, dtype=np.complex128).ravel() for k in cores]).view(np.float64)
        bb = np.asarray(b, dtype=np.uint8)
        out = np.zeros(2)
        rc = lib.qil_ref_coefficient(n, bond.ctypes.data_as(C.POINTER(C.c_int64)),
                                     flat.ctypes.data_as(C.POINTER(C.c_double)),
                                     bb.ctypes.data_as(C.POINTER(C.c_uint8)), float(amp),
                                     out.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        return complex(out[0], out[1])

    for i in range(8):
        assert abs(c_coeff(cores, c, O.bits_msb(i, 3)) - x[i]) < 1e-12
    bonds = [1, 2, 4, 3, 5, 2, 1]
    cores = [rng.standard_normal((bonds[i], 2, bonds[i + 1])) + 1j * rng.standard_normal((bonds[i], 2, bonds[i + 1]))
             for i in range(6)]
    for _ in range(20):
        b = rng.integers(0, 2, 6)
        assert abs(c_coeff(cores, 0.7, b) - O.coefficient(cores, 0.7, b)) < 1e-12
