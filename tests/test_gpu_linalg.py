"""GPU parity for the dense factorizations behind every ITensors svd/qr/factorize call of the path
(K3/K4/K5), through the C ABI."""
import numpy as np
import pytest

import qil_oracle as O

pytestmark = pytest.mark.gpu


def _rand(rng, m, n, cplx):
    A = rng.standard_normal((m, n))
    if cplx:
        A = A + 1j * rng.standard_normal((m, n))
    return A


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(1, 1), (4, 4), (8, 3), (3, 8), (64, 20), (300, 30), (5000, 24), (16384, 32),
                                   (100, 100), (40, 90)])
def test_qr(q, shape, cplx):
    m, n = shape
    rng = np.random.default_rng(m * 131 + n + cplx)
    A = _rand(rng, m, n, cplx)
    for positive in (False, True):
        Q, R = q.qr(A, positive=positive)
        k = min(m, n)
        assert Q.shape == (m, k) and R.shape == (k, n)
        assert np.abs(Q @ R - A).max() <= 1e-12 * max(1.0, np.abs(A).max()) * np.sqrt(max(m, n))
        assert np.abs(Q.conj().T @ Q - np.eye(k)).max() <= 1e-13 * np.sqrt(m)
        assert np.abs(np.tril(R, -1)).max() == 0.0
        if positive:
            d = np.diagonal(R)
            assert np.all(np.abs(d.imag) <= 1e-15 * np.abs(d).max()) and np.all(d.real >= 0)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(257, 20), (300, 7), (1024, 30), (2560, 8), (4096, 20), (16384, 20), (32768, 20),
                                   (37000, 5), (40000, 20), (16384, 6), (9000, 17)])
def test_qr_tall_panels_one_launch_tsqr(q, shape, cplx):
    """Shapes that walk the plans of the cooperative one-launch TSQR (qil_tsqr.cu): two and three levels, 128- and 256-row
    level-0 blocks, every column capacity of the register-resident factor (8 / 16 / 24 / 32), the sizes of the n = 28
    and n = 30 top splits, and panels beyond one CTA per SM (three-launch fallback).  Rank-deficient columns included:
    Q must stay orthonormal to rounding (rsvd.jl:83)."""
    m, n = shape
    rng = np.random.default_rng(m * 7 + n + cplx)
    A = _rand(rng, m, n, cplx)
    if n >= 6:
        A[:, n - 2] = A[:, 0] * 0.5 - A[:, 1]            # exact dependence
        A[:, n - 1] *= 1e-14                              # a column at rounding level
    for positive in (True, False):
        Q, R = q.qr(A, positive=positive)
        assert Q.shape == (m, n) and R.shape == (n, n)
        assert np.abs(Q @ R - A).max() <= 1e-12 * np.abs(A).max() * np.sqrt(m)
        assert np.abs(Q.conj().T @ Q - np.eye(n)).max() <= 1e-13 * np.sqrt(m)
        assert np.abs(np.tril(R, -1)).max() == 0.0
        if positive:
            d = np.diagonal(R)
            assert np.all(np.abs(d.imag) <= 1e-15 * np.abs(d).max()) and np.all(d.real >= 0)


@pytest.mark.parametrize("cplx", [False, True])
def test_qr_rank_deficient_tall(q, cplx):
    # Y = A*Omega of a low-rank A: Q must stay orthonormal to rounding (rsvd.jl:83)
    rng = np.random.default_rng(5 + cplx)
    Y = _rand(rng, 4096, 3, cplx) @ _rand(rng, 3, 30, cplx)
    Q, R = q.qr(Y, positive=True)
    assert np.abs(Q.conj().T @ Q - np.eye(30)).max() <= 1e-12
    assert np.abs(Q @ R - Y).max() <= 1e-11 * np.abs(Y).max()


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (4, 4), (6, 4), (4, 6), (30, 30), (64, 20), (20, 64), (2, 4096),
                                   (4096, 6), (60, 16384), (100, 100), (48, 200), (300, 260), (130, 500)])
def test_svd_full(q, shape, cplx):
    m, n = shape
    rng = np.random.default_rng(m * 17 + n * 3 + cplx)
    A = _rand(rng, m, n, cplx)
    U, S, Vh = q.svd_trunc(A, cutoff=0.0)
    k = min(m, n)
    Sref = np.linalg.svd(A, compute_uv=False)
    assert S.size == k
    assert np.abs(S - Sref).max() <= 1e-12 * Sref[0]
    assert np.all(np.diff(S) <= 0)
    assert np.abs(U.conj().T @ U - np.eye(k)).max() <= 1e-12
    assert np.abs(Vh @ Vh.conj().T - np.eye(k)).max() <= 1e-12
    assert np.abs((U * S) @ Vh - A).max() <= 1e-12 * Sref[0] * np.sqrt(max(m, n))


@pytest.mark.parametrize("cplx", [False, True])
def test_svd_truncation_rule_matches_oracle(q, cplx):
    rng = np.random.default_rng(77 + cplx)
    # graded spectrum spanning 1 .. 1e-14
    for (m, n) in [(40, 40), (64, 24), (24, 64), (8, 2048)]:
        k = min(m, n)
        U0, _ = np.linalg.qr(_rand(rng, m, k, cplx))
        V0, _ = np.linalg.qr(_rand(rng, n, k, cplx))
        s = np.logspace(0, -13.5, k)
        A = (U0 * s) @ V0.conj().T
        for cutoff, maxdim, mindim in [(1e-15, None, 1), (1e-12, None, 1), (1e-9, None, 1), (1e-20, 5, 1),
                                       (0.0, None, 1), (1.0, None, 3), (1e-25, None, 1)]:
            U, S, Vh = q.svd_trunc(A, cutoff=cutoff, maxdim=maxdim, mindim=mindim)
            Uo, So, Vho = O.svd_trunc(A, cutoff, O.BIG if maxdim is None else maxdim, mindim)
            # sigma below ~1e-16*sigma_max is rounding noise in LAPACK; compare ranks where the decision
            # is not inside that noise
            # skip knife-edge decisions (discarded weight within 1e-6 of cutoff*scale)
            md = O.BIG if maxdim is None else maxdim
            Sfull = np.linalg.svd(A, compute_uv=False)
            robust = (O.truncate_rank(Sfull, cutoff * (1 - 1e-6), md, mindim)
                      == O.truncate_rank(Sfull, cutoff * (1 + 1e-6), md, mindim))
            if cutoff >= 1e-25 and robust:
                assert S.size == So.size, (m, n, cutoff, maxdim, S.size, So.size)
            r = min(S.size, So.size)
            assert np.abs(S[:r] - So[:r]).max() <= 1e-13
            # compare on the common rank (equal unless the decision was a knife edge)
            assert np.abs((U[:, :r] * S[:r]) @ Vh[:r] - (Uo[:, :r] * So[:r]) @ Vho[:r]).max() <= 1e-12


def test_svd_small_singular_values_relative_accuracy(q):
    # one-sided Jacobi keeps tiny singular values accurate (needed by compress!'s ~1e-25 cutoff)
    rng = np.random.default_rng(1)
    n = 16
    U0, _ = np.linalg.qr(rng.standard_normal((n, n)))
    s = np.logspace(0, -11, n)
    A = (U0 * s) @ np.eye(n)      # columns scaled: Jacobi-friendly grading
    U, S, Vh = q.svd_trunc(A, cutoff=0.0)
    assert np.abs(S / s - 1).max() < 1e-10


def test_truncation_margin_reports_the_closest_cutoff_decision(q):
    """qil_truncation_margin: |w / (cutoff * sum sigma^2) - 1| of the closest decision (rank-parity diagnostic)."""
    ctx = q.default_context()
    sig2 = np.array([1.0, 1e-6, 1e-13])
    rng = np.random.default_rng(3)
    U, _ = np.linalg.qr(rng.standard_normal((6, 3)))
    V, _ = np.linalg.qr(rng.standard_normal((5, 3)))
    A = (U * np.sqrt(sig2)) @ V.T
    ctx.truncation_margin(reset=True)
    assert ctx.truncation_margin(reset=False) == float("inf")
    Uo, S, Vh = q.svd_trunc(A, cutoff=1e-12)
    assert S.size == 2
    m = ctx.truncation_margin(reset=True)
    # dropped weight 1e-13 against the threshold 1e-12 * (1 + 1e-6 + 1e-13): 0.9 away; the kept value is 1e6 away
    assert abs(m - 0.9) < 1e-3
    # a knife edge: sigma_3^2 = cutoff * total up to 1e-9 -> the margin says so
    tot = 1.0 + 1e-6
    sig2 = np.array([1.0, 1e-6, 1e-12 * tot * (1 + 1e-9)])
    A = (U * np.sqrt(sig2)) @ V.T
    q.svd_trunc(A, cutoff=1e-12)
    assert ctx.truncation_margin(reset=True) < 1e-6
