"""Truncating MPO x MPS (zip-up sweep) on the GPU against its numpy restatement (same algorithm) and
against the reference's exact apply (apply.jl:75-122)."""
import math

import numpy as np
import pytest

import qil_oracle as O

pytestmark = pytest.mark.gpu


def _relerr(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("kind,n,cutoff,maxdim", [("random", 8, 1e-14, None), ("random", 10, 1e-6, 12), ("sin_decay", 12, 1e-12, None),
                                                  ("complex", 9, 1e-10, None)])
def test_zipup_qft_matches_oracle_and_exact_apply(q, kind, n, cutoff, maxdim):
    N = 2**n
    rng = np.random.default_rng(n)
    if kind == "random":            # the :random benchmark input: full-rank MPS, fused bonds explode in the exact apply
        x = rng.standard_normal(N)
    elif kind == "complex":
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    else:
        t = np.arange(N) / (2.5 * N)
        x = np.sin(t) * np.exp(-0.08 * t) + np.sin(2.5 * t) * np.exp(-0.03 * t)
    psi = q.signal_mps(x, cutoff=1e-14)
    W = q.build_qft_mpo(n, cutoff=1e-14)
    got = q.apply_zipup(W, psi, cutoff=cutoff, maxdim=maxdim)
    want = O.apply_zipup(W.cores(), psi.cores(), cutoff, O.BIG if maxdim is None else maxdim)
    assert got.bonds == O.bonds_of(want)
    if maxdim is not None:
        assert max(got.bonds) <= maxdim
    v = q.mps_to_vector(got)
    vo = O.mps_to_vector(want, psi.amplitude)
    assert np.abs(v - vo).max() <= 1e-10 * np.abs(vo).max()
    exact = q.mps_to_vector(W * psi)
    if maxdim is None:
        assert _relerr(v, exact) <= 10 * math.sqrt(n * cutoff) + 1e-12
        assert max(got.bonds) <= max((W * psi).bonds)
    f = np.fft.fft(x) / math.sqrt(N)
    rev = np.array([O.bitrev(i, n) for i in range(N)])
    if maxdim is None:
        assert _relerr(v, f[rev]) <= 10 * math.sqrt(n * cutoff) + 1e-9


def test_zipup_zt_paired(q):
    """zT MPO on a paired-register state: the zip-up bonds stay far below the fused D * chi bonds of the exact apply."""
    n = 8
    N = 2**n
    j = np.arange(N)
    x = (1.0003 * np.exp(0.01j)) ** j * np.cos(0.07 * j)
    z = q.signal_ztmps(x, cutoff=1e-12)
    W = q.build_zt_mpo(z, 2 * math.pi, cutoff=1e-12, maxdim=128)
    exact = W * z
    zz = q.apply_zipup(W, z, cutoff=1e-20)
    assert max(zz.bonds) < max(exact.bonds)
    a, b = q.mps_to_vector(zz), q.mps_to_vector(exact)
    assert _relerr(a, b) <= 1e-8
    want = O.apply_zipup(W.cores(), z.cores(), 1e-20, O.BIG)
    assert zz.bonds == O.bonds_of(want)
    with pytest.raises(q.ArgumentError):
        q.apply_zipup(q.build_qft_mpo(4), z)
