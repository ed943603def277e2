"""GPU parity for the randomized-SVD path (K1/K2 streaming DMMA+TMA GEMMs, TSQR, D&C encoder) through the
C ABI.  The oracle and the GPU are fed the SAME normal stream (Omega[c][j] = stream[c + C*j]), so the
comparison does not depend on a random-number generator."""
import math

import numpy as np
import pytest

import qil_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _stream(n, cplx, seed=1234):
    rng = np.random.default_rng(seed)
    if cplx:
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / math.sqrt(2.0)
    return rng.standard_normal(n)


def _omega_fn(stream, L):
    def fn(cols, iscomplex):
        return stream[: cols * L].reshape(L, cols).T
    return fn


def _lowrank(rng, m, n, r, cplx):
    X = rng.standard_normal((m, r))
    Y = rng.standard_normal((r, n))
    if cplx:
        X = X + 1j * rng.standard_normal((m, r))
        Y = Y + 1j * rng.standard_normal((r, n))
    return X @ Y


# ---------------------------------------------------------------------------------------------
# rsvd on a matrix: test/test_rsvd.jl:27-120
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape,rank", [((100, 100), 10), ((50, 80), 5), ((2048, 1024), 10), ((1024, 4096), 7),
                                        ((4096, 512), 12)])
@pytest.mark.parametrize("q_iter", [0, 2])
def test_rsvd_lowrank(q, shape, rank, cplx, q_iter):
    m, n = shape
    rng = np.random.default_rng(m + n + rank + cplx)
    A = _lowrank(rng, m, n, rank, cplx)
    U, S, Vh = q.rsvd(A, k=rank, p=5, q=q_iter)
    assert S.size == rank
    assert np.linalg.norm((U * S) @ Vh - A) / np.linalg.norm(A) < 1e-10
    assert np.abs(U.conj().T @ U - np.eye(rank)).max() < 1e-10
    assert np.abs(Vh @ Vh.conj().T - np.eye(rank)).max() < 1e-10
    assert np.all(np.diff(S) <= 0) and np.all(S >= 0)
    Sref = np.linalg.svd(A, compute_uv=False)[:rank]
    assert np.abs(S - Sref).max() <= 1e-10 * Sref[0]


def test_rsvd_options_and_errors(q):
    rng = np.random.default_rng(3)
    A = _lowrank(rng, 100, 100, 10, False)
    assert q.rsvd(A, k=10, p=5, maxdim=4)[1].size == 4
    assert q.rsvd(A, k=20, p=5, cutoff=1e-12, maxdim=25)[1].size == 10
    assert q.rsvd(A, k=20, p=5, cutoff=1.0, maxdim=25, mindim=3)[1].size == 3
    U1, S1, V1 = q.rsvd(A, k=10, p=5, random_seed=99)
    U2, S2, V2 = q.rsvd(A, k=10, p=5, random_seed=99)
    assert np.array_equal(S1, S2) and np.array_equal(U1, U2) and np.array_equal(V1, V2)   # same seed -> same result
    with pytest.raises(q.ErrorException):
        q.rsvd(np.zeros((0, 5)))
    # host-supplied stream == oracle with the same Omega
    st = _stream(100 * 15, False)
    U, S, Vh = q.rsvd(A, k=10, p=5, normal_stream=st)
    Uo, So, Vho = O.rsvd(A, k=10, p=5, omega=_omega_fn(st, 15)(100, False))
    assert np.abs(S - So).max() <= 1e-12 * So[0]
    assert np.abs((U * S) @ Vh - (Uo * So) @ Vho).max() <= 1e-11 * So[0]


@pytest.mark.parametrize("cplx", [False, True])
def test_rsvd_streaming_kernels_match_oracle_with_same_omega(q, cplx):
    # full-rank-ish matrix with decaying spectrum: the result depends on Omega, so this pins K1/K2 themselves
    rng = np.random.default_rng(17 + cplx)
    m, n = 1024, 2048
    k, p = 20, 10
    L = k + p
    Uf, _ = np.linalg.qr(rng.standard_normal((m, 64)) + (1j * rng.standard_normal((m, 64)) if cplx else 0))
    Vf, _ = np.linalg.qr(rng.standard_normal((n, 64)) + (1j * rng.standard_normal((n, 64)) if cplx else 0))
    A = (Uf * np.logspace(0, -6, 64)) @ Vf.conj().T
    st = _stream(n * L, cplx)
    for q_iter in (0, 1):
        U, S, Vh = q.rsvd(A, k=k, p=p, q=q_iter, maxdim=L, normal_stream=st)
        Uo, So, Vho = O.rsvd(A, k=k, p=p, q=q_iter, maxdim=L, omega=_omega_fn(st, L)(n, cplx))
        assert S.size == So.size
        assert np.abs(S - So).max() <= 1e-10 * So[0]
        assert np.abs((U * S) @ Vh - (Uo * So) @ Vho).max() <= 1e-9 * So[0]


# ---------------------------------------------------------------------------------------------
# signal_mps(:rsvd)
# ---------------------------------------------------------------------------------------------
def test_encode_rsvd_kats(q, goldens):
    x = np.array(goldens["coefficient_kats"]["x"], dtype=float)
    psi = q.signal_mps(x, method="rsvd")
    assert abs(psi.amplitude - np.linalg.norm(x)) < 1e-12
    for i in range(8):
        assert abs(q.coefficient(psi, i) - x[i]) < 1e-10
    ps = q.signal_mps(x, method="svd")
    assert np.abs(q.mps_to_vector(psi) - q.mps_to_vector(ps)).max() < 1e-10
    assert abs(q.coefficient(q.signal_mps(np.array([3.0, 4.0]), method="rsvd"), 1) - 4.0) < 1e-12


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n", [5, 6, 9])
def test_encode_rsvd_random_full_rank(q, n, cplx):
    # test/test_signal_converters.jl:131-139,194-202: with k >= full rank RSVD reproduces the signal
    rng = np.random.default_rng(n + cplx)
    x = rng.standard_normal(2**n) + (1j * rng.standard_normal(2**n) if cplx else 0)
    psi = q.signal_mps(x, method="rsvd", k=2 ** ((n + 1) // 2), p=0)
    assert np.abs(q.mps_to_vector(psi) - x).max() <= 1e-10 * np.abs(x).max()


def test_encode_rsvd_signal_tutorial(q, goldens):
    g = goldens["signal_tutorial_n10"]
    n = g["n"]
    x = q.generate_signal(n, kind="sin_decay", dt=1.0 / 2**n, freq=[2 * math.pi * f for f in g["freq_over_2pi"]],
                          decay_rate=g["decay_rate"], phase=g["phase"])
    k = g["rsvd_k"]
    st = _stream(2 ** n * (k + 10), False)
    psi = q.signal_mps(x, method="rsvd", cutoff=g["cutoff"], maxdim=g["maxdim"], k=k, normal_stream=st)
    co, c = O.tt_rsvd(x, cutoff=g["cutoff"], maxdim=g["maxdim"], k=k, omega_fn=_omega_fn(st, k + 10))
    assert psi.bonds == O.bonds_of(co)
    assert max(psi.bonds) == g["max_bond_rsvd"]
    got = q.mps_to_vector(psi)
    assert np.abs(got - O.mps_to_vector(co, c)).max() <= 1e-9 * np.abs(x).max()
    assert np.linalg.norm(got - x) / np.linalg.norm(x) == pytest.approx(g["rel_err_rsvd"], rel=0.05)


@pytest.mark.parametrize("n,kind", [(20, "real"), (21, "real"), (20, "complex"), (22, "complex")])
def test_encode_rsvd_streaming_path_matches_oracle(q, n, kind):
    """n >= 20 goes through the TMA/DMMA streaming kernels at the top split (C2/C3-type workloads)."""
    N = 2**n
    j = np.arange(N)
    if kind == "real":
        dt = 1.0 / (2.5 * N)
        x = np.sin(1.0 * dt * j) * np.exp(-0.08 * dt * j) + np.sin(2.5 * dt * j) * np.exp(-0.03 * dt * j)
        kw = dict(k=15, p=5, q=2, cutoff=1e-12)
    else:
        x = (1.00015 * np.exp(0.002j)) ** (j * 2.0 ** (20 - n)) * np.cos(0.0061 * j * 2.0 ** (20 - n))
        kw = dict(k=50, p=5, q=2, cutoff=1e-12, maxdim=128)
    L = kw["k"] + kw["p"]
    cols_top = 2 ** (n - n // 2)
    st = _stream(cols_top * L, kind == "complex")
    psi = q.signal_mps(x, method="rsvd", normal_stream=st, **kw)
    co, c = O.tt_rsvd(x, omega_fn=_omega_fn(st, L), **kw)
    # (the D&C bonds need not equal the sequential TT-SVD bonds: each split truncates relative to its own
    #  factor, SignalConverters.jl:166-179 -- parity is with the same algorithm)
    assert psi.bonds == O.bonds_of(co)
    assert abs(psi.amplitude - c) <= 1e-12 * c
    rng = np.random.default_rng(n)
    idx = np.concatenate([[0, 1, N // 2, N - 1], rng.integers(0, N, 4000)])
    bits = ((idx[:, None] >> np.arange(n - 1, -1, -1)[None, :]) & 1).astype(np.uint8)
    got = q.coefficients(psi, bits)
    want = O.coefficient_batch(co, c, bits)
    scale = np.abs(x).max()
    assert np.abs(got - want).max() <= TOL * scale
    assert np.abs(got - x[idx]).max() <= 1e-5 * scale        # cutoff 1e-12 on sigma^2
    # device-generated Omega (no host stream): same bonds, same amplitudes to the truncation level
    psi2 = q.signal_mps(x, method="rsvd", **kw)
    assert psi2.bonds == psi.bonds
    assert np.abs(q.coefficients(psi2, bits) - got).max() <= 1e-8 * scale


def test_signal_ztmps_rsvd_tutorial_bonds(q, goldens):
    g = goldens["zt_tutorial_n20"]
    N = 2 ** g["n"]
    a = g["a_abs"] * np.exp(1j * g["a_arg"])
    j = np.arange(N)
    x = a**j * np.cos(g["omega0"] * j)
    z = q.signal_ztmps(x, method="rsvd", k=g["k"], p=g["p"], q=g["q"], cutoff=g["cutoff"], maxdim=g["maxdim"])
    assert z.bonds_main == g["bonds_main"]
    assert z.bonds_copy == g["bonds_copy"]
    idx = [0, 1, 2, 12345, N // 2, N - 1]
    bits = np.array([O.interleave(O.bits_msb(i, g["n"]), O.bits_msb(i, g["n"])) for i in idx], dtype=np.uint8)
    assert np.abs(q.coefficients(z, bits) - x[idx]).max() <= 1e-9 * np.abs(x).max()
