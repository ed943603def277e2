"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the
same seeded inputs and against the reference's golden vectors.  Tolerance for floating point: 1e-10
relative to the largest amplitude (BASELINE.json north_star), tighter where the reference's own tests
are tighter."""
import math

import numpy as np
import pytest

import qil_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-10


def _rand_mps(rng, bonds, cplx):
    cores = []
    for i in range(len(bonds) - 1):
        c = rng.standard_normal((bonds[i], 2, bonds[i + 1]))
        if cplx:
            c = c + 1j * rng.standard_normal(c.shape)
        cores.append(c / math.sqrt(bonds[i] * 2))
    return cores


def _rand_mpo(rng, bonds, cplx):
    cores = []
    for i in range(len(bonds) - 1):
        c = rng.standard_normal((bonds[i], 2, 2, bonds[i + 1]))
        if cplx:
            c = c + 1j * rng.standard_normal(c.shape)
        cores.append(c / math.sqrt(bonds[i] * 2))
    return cores


def _hash_bits(B, n, seed):
    """Counter-based hash so the bitstrings are reproducible anywhere."""
    idx = np.arange(B * n, dtype=np.uint64).reshape(B, n)
    z = idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)
    z ^= z >> np.uint64(30); z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27); z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return (z & np.uint64(1)).astype(np.uint8)


# ------------------------------------------------------------------------------------------
# coefficient (K8)
# ------------------------------------------------------------------------------------------
def test_coefficient_kats(q, goldens):
    g = goldens["coefficient_kats"]
    x = np.array(g["x"], dtype=float)
    cores, c = O.tt_svd(x)
    psi = q.SignalMPS.from_cores(cores, c)
    for i in range(8):
        assert abs(q.coefficient(psi, i) - x[i]) < 1e-12
    assert abs(q.coefficient(psi, "101") - x[5]) < 1e-12
    assert abs(q.coefficient(psi, "[1,0,1]") - x[5]) < 1e-12
    assert abs(q.coefficient(psi, (1, 0, 1)) - x[5]) < 1e-12
    assert abs(q.coefficient(psi, 1, 0, 1) - x[5]) < 1e-12
    assert abs(psi[1, 0, 1] - x[5]) < 1e-12
    A1 = np.zeros((1, 2, 1)); A1[0, 1, 0] = 1.0
    A2 = np.zeros((1, 2, 1)); A2[0, 0, 0] = 1.0
    A3 = np.zeros((1, 2, 1)); A3[0, 1, 0] = 0.5
    hb = q.SignalMPS.from_cores([A1, A2, A3])
    assert q.coefficient(hb, g["handbuilt_bits"]) == pytest.approx(g["handbuilt_value"], rel=1e-12)
    assert q.coefficient(hb, 0b101) == pytest.approx(0.5, rel=1e-12)
    for bad in ([1, 0], [2, 0, 1], "[1, 2, 1]", 0b1000):
        with pytest.raises(q.ArgumentError):
            q.coefficient(hb, bad)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("bonds", [
    [1, 2, 3, 2, 1],
    [1, 2, 4, 8, 16, 24, 24, 16, 8, 4, 2, 1],
    [1, 2, 8, 37, 61, 124, 192, 122, 90, 33, 16, 4, 1],
    [1, 7, 300, 5, 1],
])
def test_coefficient_batch_random(q, bonds, cplx):
    rng = np.random.default_rng(len(bonds) * 7 + cplx)
    cores = _rand_mps(rng, bonds, cplx)
    n = len(cores)
    for B in (1, 5, 1000):
        bits = _hash_bits(B, n, 1234 + B)
        want = O.coefficient_batch(cores, 1.75, bits)
        psi = q.SignalMPS.from_cores(cores, 1.75)
        got = q.coefficients(psi, bits)
        scale = np.abs(want).max()
        assert np.abs(got - want).max() <= TOL * scale
    assert q.coefficients(psi, np.zeros((0, n), dtype=np.uint8)).shape == (0,)


def test_coefficient_all_bitstrings_reconstruct_signal(q):
    # docs/src/tutorials/signal.jl:149-150: every coefficient of an encoded signal
    n = 10
    x = q.generate_signal(n, kind="sin_decay", freq=[1.0, 2.5], decay_rate=[0.08, 0.03])
    cores, c = O.tt_svd(x, cutoff=1e-15)
    psi = q.SignalMPS.from_cores(cores, c)
    bits = np.array([O.bits_msb(i, n) for i in range(2**n)], dtype=np.uint8)
    got = q.coefficients(psi, bits)
    want = O.coefficient_batch(cores, c, bits)
    assert np.abs(got - want).max() <= 1e-13 * np.abs(x).max()
    assert np.abs(got - x).max() <= 1e-6 * np.abs(x).max()   # cutoff 1e-15 on sigma^2


# ------------------------------------------------------------------------------------------
# apply (K6 / K7)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wc,pc", [(False, False), (True, False), (False, True), (True, True)])
def test_apply_mpo_mps_random(q, wc, pc):
    rng = np.random.default_rng(21 + 2 * wc + pc)
    wb = [1, 2, 5, 8, 3, 1]
    pb = [1, 2, 3, 4, 2, 1]
    W = _rand_mpo(rng, wb, wc)
    P = _rand_mps(rng, pb, pc)
    want = O.apply_mpo_mps(W, P)
    out = q.apply(q.SingleSiteMPO.from_cores(W), q.SignalMPS.from_cores(P, 2.5))
    assert out.amplitude == 2.5
    assert out.bonds == O.bonds_of(want)
    for a, b in zip(out.cores(), want):
        assert np.abs(a - b).max() <= 1e-14 * max(1.0, np.abs(b).max())
    # `*` alias and the dense cross-check of test/test_apply.jl
    out2 = q.SingleSiteMPO.from_cores(W) * q.SignalMPS.from_cores(P)
    dense = O.mpo_to_dense(W) @ O.mps_to_vector(P)
    assert np.allclose(O.mps_to_vector(out2.cores()), dense, atol=1e-12)


def test_apply_identity_and_pauli_x(q):
    n = 4
    rng = np.random.default_rng(2)
    P = _rand_mps(rng, [1, 2, 3, 2, 1], False)
    I = [np.eye(2).reshape(1, 2, 2, 1) for _ in range(n)]
    X = [np.array([[0.0, 1.0], [1.0, 0.0]]).reshape(1, 2, 2, 1) for _ in range(n)]
    psi = q.SignalMPS.from_cores(P)
    v = O.mps_to_vector(P)
    assert np.allclose(O.mps_to_vector((q.SingleSiteMPO.from_cores(I) * psi).cores()), v, atol=1e-14)
    assert np.allclose(O.mps_to_vector((q.SingleSiteMPO.from_cores(X) * psi).cores()), v[::-1], atol=1e-14)


def test_apply_errors(q):
    rng = np.random.default_rng(3)
    W = q.SingleSiteMPO.from_cores(_rand_mpo(rng, [1, 2, 2, 1], False))
    psi = q.SignalMPS.from_cores(_rand_mps(rng, [1, 2, 2, 2, 1], False))
    with pytest.raises(q.ArgumentError):
        q.apply(W, psi)
    Wp = q.PairedSiteMPO.from_cores(_rand_mpo(rng, [1, 2, 2, 2, 1], False))
    with pytest.raises(q.ArgumentError):
        q.apply(Wp, psi)


@pytest.mark.parametrize("c1,c2", [(False, False), (True, True), (True, False)])
def test_apply_mpo_mpo(q, c1, c2):
    rng = np.random.default_rng(31 + c1 + 2 * c2)
    W1 = _rand_mpo(rng, [1, 2, 3, 2, 1], c1)
    W2 = _rand_mpo(rng, [1, 3, 2, 4, 1], c2)
    want = O.apply_mpo_mpo(W1, W2)
    out = q.apply(q.SingleSiteMPO.from_cores(W1), q.SingleSiteMPO.from_cores(W2))
    assert out.bonds == O.mpo_bonds(want)
    for a, b in zip(out.cores(), want):
        assert np.abs(a - b).max() <= 1e-14
    assert np.allclose(O.mpo_to_dense(out.cores()), O.mpo_to_dense(W2) @ O.mpo_to_dense(W1), atol=1e-12)


def test_apply_mpo_mpo_window(q):
    # unequal lengths (test/test_apply.jl): the shorter MPO acts on sites 2..3 of a 4-site MPO
    rng = np.random.default_rng(41)
    W1 = _rand_mpo(rng, [1, 2, 3, 2, 1], False)
    W2 = _rand_mpo(rng, [1, 2, 1], False)
    out = q.apply(q.SingleSiteMPO.from_cores(W1), q.SingleSiteMPO.from_cores(W2), start1=1, start2=0)
    want = O.apply_mpo_mpo(W1, W2, start1=1, start2=0)
    for a, b in zip(out.cores(), want):
        assert a.shape == b.shape and np.abs(a - b).max() <= 1e-14


def test_zt_tutorial_table_through_apply_and_coefficient(q, goldens):
    """n=2 z-transform table of docs/src/tutorials/zt.md:283-304: MPO and MPS come from the oracle,
    apply + coefficient run on the GPU."""
    g = goldens["zt_tutorial_n2"]
    n, N = 2, 4
    x = np.array([g["a"] ** j * math.cos(g["omega0_over_pi"] * math.pi * j) for j in range(N)])
    cores, c = O.signal_ztmps(x, cutoff=1e-14, maxdim=64)
    W = O.build_zt_mpo(n, 2 * math.pi, cutoff=1e-14, maxdim=64)
    out = q.PairedSiteMPO.from_cores(W) * q.ZTMPS.from_cores(cores, c)
    bits = np.array([O.interleave(O.bits_lsb(k, n), O.bits_lsb(l, n)) for k in range(N) for l in range(N)],
                    dtype=np.uint8)
    chi = q.coefficients(out, bits).reshape(N, N)
    re = np.array(g["chi_5digits_re"]); im = np.array(g["chi_5digits_im"])
    assert np.abs(chi.real - re).max() < 6e-6 and np.abs(chi.imag - im).max() < 6e-6
    ref = np.array([[sum(x[j] * np.exp(-(2 * math.pi * k + 2j * math.pi * l) / N * j) for j in range(N)) / N
                     for l in range(N)] for k in range(N)])
    assert (np.abs(chi - ref) / np.abs(ref)).max() < 1e-13


def test_cores_into_one_host_buffer(q):
    rng = np.random.default_rng(9)
    for cplx in (False, True):
        cores = _rand_mps(rng, [1, 2, 4, 3, 2, 1], cplx)
        psi = q.SignalMPS.from_cores(cores, 2.0)
        nbytes = sum(c.nbytes for c in cores)
        buf = bytearray(nbytes + 64)
        got = psi.cores_into(buf)
        for a, b in zip(got, psi.cores()):
            assert a.shape == b.shape and np.array_equal(a, b)
        for a, b in zip(got, cores):
            assert np.array_equal(a, b)
        with pytest.raises(q.ArgumentError):
            psi.cores_into(bytearray(nbytes - 8))
