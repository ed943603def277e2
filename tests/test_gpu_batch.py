"""Batched encode of independent signals (BASELINE configs[1] family: sin_decay signals with shifted frequencies,
signal_mps(:rsvd, maxdim=64), QFT apply) -- qil_encode_rsvd_batch_dev against the one-signal entry point and the
CPU oracle."""
import numpy as np
import pytest

import qil_oracle as O

pytestmark = pytest.mark.gpu


def _family(n, count):
    N = 2**n
    t = np.arange(N) / (2.5 * N)
    return np.stack([np.sin((1 + 0.01 * b) * t) * np.exp(-0.08 * t) + np.sin((2.5 + 0.01 * b) * t) * np.exp(-0.03 * t)
                     for b in range(count)])


@pytest.mark.parametrize("n,count,workers,qit", [(14, 9, 4, 0), (20, 6, 16, 0), (16, 5, 4, 2), (20, 3, 4, 1)])
def test_batch_encode_matches_single_and_oracle(q, n, count, workers, qit):
    import torch
    N = 2**n
    xs = _family(n, count)
    ctx = q.default_context()
    d = torch.from_numpy(xs).cuda()
    torch.cuda.synchronize()
    kw = dict(k=20, p=10, q=qit, cutoff=1e-14, maxdim=64)
    batch = q.signal_mps_batch_dev(ctx, d.data_ptr(), N, count, False, workers=workers, **kw)
    assert len(batch) == count
    W = q.build_qft_mpo(n, cutoff=1e-14, maxdim=128, ctx=ctx)
    rng = np.random.default_rng(0)
    idx = rng.integers(0, N, 512)
    bits = ((idx[:, None] >> np.arange(n - 1, -1, -1)[None, :]) & 1).astype(np.uint8)
    outs = q.apply_batch(W, batch)                   # one launch for the whole batch
    assert len(outs) == count
    for b in range(count):
        single = W * batch[b]
        assert outs[b].bonds == single.bonds
        for ca, cb in zip(outs[b].cores(), single.cores()):
            assert np.array_equal(ca, cb)
    for b in range(count):
        one = q.signal_mps_dev(ctx, d[b].data_ptr(), N, False, method="rsvd", **kw)
        assert batch[b].bonds == one.bonds
        assert abs(batch[b].amplitude - one.amplitude) <= 1e-13 * one.amplitude
        got = q.coefficients(batch[b], bits)
        # batched and single encodes sum their split-K partials in different orders; with power iterations that rounding
        # noise is amplified by the conditioning of the algorithm (oracle vs oracle on a 1-ulp perturbed input: 4e-11 at n=20)
        # (the single encode factors its tall panels with the one-launch WY TSQR, a batch with the three-launch reflector
        # form: two Householder QRs that differ at rounding level, hence 1e-11 and not 1e-12 without power iterations)
        assert np.abs(got - q.coefficients(one, bits)).max() <= (1e-11 if qit == 0 else 1e-9) * np.abs(xs[b]).max()
        assert np.abs(got - xs[b][idx]).max() <= 1e-6 * np.abs(xs[b]).max()
        if b in (0, count - 1):
            co, c = O.tt_rsvd(xs[b], **kw)
            assert batch[b].bonds == O.bonds_of(co)
            # (the oracle draws its own Omega: with power iterations the two sketches agree to the truncation level only)
            assert np.abs(got - O.coefficient_batch(co, c, bits)).max() <= (1e-10 if qit == 0 else 5e-8) * np.abs(xs[b]).max()
            # QFT apply on the batched result == FFT with bit-reversed output (test_qft_transformer.jl:427-463)
            out = W * batch[b]
            f = np.fft.fft(xs[b]) / np.sqrt(N)
            rev = np.array([O.bitrev(int(i), n) for i in idx])
            assert np.abs(q.coefficients(out, bits) - f[rev]).max() <= 1e-6 * np.abs(f).max()


def test_batch_encode_empty_and_errors(q):
    import torch
    ctx = q.default_context()
    assert q.signal_mps_batch_dev(ctx, 0, 16, 0, False) == []
    d = torch.zeros(64, dtype=torch.float64, device="cuda")
    with pytest.raises(q.ArgumentError):
        q.signal_mps_batch_dev(ctx, d.data_ptr(), 16, 4, False, k=0)


@pytest.mark.parametrize("b", [95, 135])
def test_c2_family_q0_bonds_follow_the_oracle_deep_in_the_tree(q, b):
    """Members of the configs[1] family whose bonds three levels down the tree were inflated (6-7 instead of 4) when
    the sketch was narrowed without a power iteration (q = 0): the narrowing is now limited to q >= 1, and the leading
    bonds follow the oracle again."""
    n = 20
    N = 2**n
    t = np.arange(N) / (2.5 * N)
    x = np.sin((1 + 0.01 * b) * t) * np.exp(-0.08 * t) + np.sin((2.5 + 0.01 * b) * t) * np.exp(-0.03 * t)
    kw = dict(k=20, p=10, q=0, cutoff=1e-14, maxdim=64)
    psi = q.signal_mps(x, method="rsvd", **kw)
    co, c = O.tt_rsvd(x, **kw)
    assert psi.bonds[:6] == O.bonds_of(co)[:6] == [2, 4, 4, 4, 4, 4]


def test_batch_encode_shared_stream_bonds_equal_oracle_for_every_signal(q):
    """configs[1] family with ONE host-drawn normal stream given to the oracle and to the batched encoder (every signal
    and every split reuse it, as the reference reseeds per rsvd call): bonds identical for every signal, amplitudes
    within 1e-10."""
    import torch
    n, count = 20, 24
    N = 2**n
    xs = _family(n, count)
    ctx = q.default_context()
    d = torch.from_numpy(xs).cuda()
    kw = dict(k=20, p=10, q=0, cutoff=1e-14, maxdim=64)
    L = kw["k"] + kw["p"]
    cols = 2 ** (n - n // 2)
    stream = np.random.default_rng(1234).standard_normal(cols * L)
    sd = torch.from_numpy(stream).cuda()
    torch.cuda.synchronize()
    batch = q.signal_mps_batch_dev(ctx, d.data_ptr(), N, count, False, normal_stream_dev=sd.data_ptr(),
                                   stream_len=stream.size, **kw)
    rng = np.random.default_rng(1)
    idx = rng.integers(0, N, 256)
    bits = ((idx[:, None] >> np.arange(n - 1, -1, -1)[None, :]) & 1).astype(np.uint8)
    fn = lambda c_, iscomplex: stream[: c_ * L].reshape(L, c_).T
    for b in range(count):
        co, c = O.tt_rsvd(xs[b], omega_fn=fn, **kw)
        assert batch[b].bonds == O.bonds_of(co), b
        got = q.coefficients(batch[b], bits)
        assert np.abs(got - O.coefficient_batch(co, c, bits)).max() <= 1e-10 * np.abs(xs[b]).max()
