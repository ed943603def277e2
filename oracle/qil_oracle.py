"""CPU ORACLE for the QILaplace.jl hot path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy/LAPACK restatement of the reference algorithms (Julia + ITensors.jl 0.9,
neither of which can run in this environment).  It is the checker for the CUDA product path.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing under ``qilaplace.jl_b200/`` imports it, and the product path
raises when ``libqilcuda.so`` is missing instead of falling back to this file.

PINNING (tests/test_oracle_golden.py): the restatement reproduces the reference's own goldens --
x=[1..8] coefficient KATs (test/test_signal_converters.jl:114-202), the hand-built MPS KAT
(test/test_mps.jl:404-427), QFT == DFT with bit-reversed output on basis states and random complex
input (test/test_qft_transformer.jl:331-464), DT / zT analytic formulas on all basis states
(test/test_dt_transformer.jl:60-92,211-238, test/test_zt_transformer.jl:20-110), the n=2 4x4 chi
table and MPO bonds of docs/src/tutorials/zt.md:185-304, the n=4 bonds of docs/src/tutorials/dft.md,
the n=20 ZTMPS bond list of zt.md:344-406 and the QFT/DT/zT max-bond series stored in
scripts/benchmark/results/mpo_bond_dim.jld2.
UNPINNED: the Gaussian test matrix of `rsvd` (Julia's Xoshiro stream after Random.seed!(1234)) -- RSVD
results are Omega-independent to rounding only when k+p >= numerical rank.

The arithmetic that the reference delegates to ITensors.jl (third-party, compat "0.9", not vendored
under /root/reference) is restated from its published behaviour:
  * truncation rule of NDTensors `truncate!!` (relative, cumulative, on sigma^2),
  * `factorize` decision rule (QR when nothing truncates, SVD when cutoff <= 1e-12, else eigen),
  * thin `qr` (optionally `positive=true`), `combiner(i, j)` = first index fastest.

Array conventions (row-major numpy): MPS core ``M[l, s, r]``; MPO core ``W[l, p, s, r]`` with
``p`` = primed/input leg and ``s`` = unprimed/output leg; a length-2^n signal is viewed MSB-first
(site 1 = most significant bit), which is the same tensor as the reference's column-major, permuted
ITensor (src/signals/SignalConverters.jl:39-41).
"""
from __future__ import annotations

import math
import numpy as np

BIG = 2**62


# --------------------------------------------------------------------------------------------
# ITensors / NDTensors semantics
# --------------------------------------------------------------------------------------------
def truncate_rank(S, cutoff=0.0, maxdim=BIG, mindim=1):
    """Number of singular values kept by NDTensors `truncate!!` applied to P = S.^2 (descending).

    Call sites in the reference: rsvd.jl:103, SignalConverters.jl:84,266, mps.jl:929,946,
    qft_transformer.jl:82, dt_transformer.jl:213,261 (all via ITensors.svd(...; cutoff, maxdim)).
    """
    P = np.asarray(S, dtype=np.float64) ** 2
    n = P.size
    if n <= 1:
        return n
    maxdim = int(min(maxdim, BIG))
    err = 0.0
    while n > maxdim:
        err += P[n - 1]
        n -= 1
    scale = float(P.sum())
    if scale == 0.0:
        scale = 1.0
    while n > mindim and err + P[n - 1] <= cutoff * scale:
        err += P[n - 1]
        n -= 1
    return max(n, 1)


def svd_trunc(M, cutoff=0.0, maxdim=BIG, mindim=1):
    """ITensors.svd(M, rows; cutoff, maxdim, mindim) -> U[:, :r], S[:r], Vh[:r, :]."""
    try:
        U, S, Vh = np.linalg.svd(M, full_matrices=False)
    except np.linalg.LinAlgError:  # gesdd failure -> gesvd, as LinearAlgebra does
        import scipy.linalg as sla
        U, S, Vh = sla.svd(M, full_matrices=False, lapack_driver="gesvd")
    r = truncate_rank(S, cutoff, maxdim, mindim)
    return U[:, :r], S[:r], Vh[:r, :]


def qr_thin(M, positive=False):
    """ITensors.qr(M, rows; positive) -- thin QR; positive=true makes diag(R) real and >= 0."""
    Q, R = np.linalg.qr(M, mode="reduced")
    if positive:
        d = np.diagonal(R).copy()
        ph = np.where(np.abs(d) > 0, d / np.where(np.abs(d) > 0, np.abs(d), 1.0), 1.0)
        Q = Q * ph[None, :]
        R = np.conj(ph)[:, None] * R
    return Q, R


def eigen_trunc_left(M, cutoff, maxdim, mindim=1):
    """factorize(...; which_decomp="eigen", ortho="left"): M ~ L (L^H M), L from eig(M M^H)."""
    G = M @ M.conj().T
    w, V = np.linalg.eigh(G)
    order = np.argsort(-w)
    w = w[order]
    V = V[:, order]
    w = np.where(w < 0, 0.0, w)
    r = truncate_rank(np.sqrt(w), cutoff, maxdim, mindim)
    L = V[:, :r]
    return L, L.conj().T @ M


def factorize(M, ortho="left", cutoff=None, maxdim=None):
    """ITensors.factorize decision rule; returns (L, R) with M ~ L @ R.

    ortho="left": L is an isometry (U, S*V); ortho="right": R is (U*S, V).
    No cutoff and no effective maxdim -> thin QR; cutoff <= 1e-12 -> SVD; else eigen.
    """
    m, n = M.shape
    md = min(m, n) if maxdim is None else min(int(min(maxdim, BIG)), min(m, n))
    might_truncate = (cutoff is not None) or md < min(m, n)
    if not might_truncate:
        if ortho == "left":
            return np.linalg.qr(M, mode="reduced")
        Q, R = np.linalg.qr(M.T, mode="reduced")
        return R.T, Q.T
    if cutoff is None or cutoff <= 1e-12:
        U, S, Vh = svd_trunc(M, 0.0 if cutoff is None else cutoff, md)
        if ortho == "left":
            return U, S[:, None] * Vh
        return U * S[None, :], Vh
    if ortho == "left":
        return eigen_trunc_left(M, cutoff, md)
    Lt, Rt = eigen_trunc_left(M.T, cutoff, md)
    return Rt.T, Lt.T


# --------------------------------------------------------------------------------------------
# signals (src/signals/Signals.jl) -- deterministic kinds only
# --------------------------------------------------------------------------------------------
def generate_signal(n, kind="sin", dt=None, freq=None, **kw):
    """generate_signal (Signals.jl:188-235) for :sin, :sin_decay, :abs_cos_power_p8.

    The Xoshiro-seeded kinds (:random, :multi_sin, :multi_sin_exp) cannot be regenerated outside Julia.
    """
    N = 2**n
    f = 2 * math.pi if freq is None else freq
    if dt is None:
        fmax = max(abs(v) for v in f) if isinstance(f, (list, tuple, np.ndarray)) else abs(f)
        dt = 1.0 if fmax == 0 else 1.0 / (fmax * N)
    j = np.arange(N, dtype=np.float64)
    vec = isinstance(f, (list, tuple, np.ndarray))
    if kind == "sin":
        if vec:
            ph = kw.get("phase", [0.0] * len(f))
            return sum(np.sin(w * dt * j + p) for w, p in zip(f, ph))
        return np.sin(f * dt * j + kw.get("phase", 0.0))
    if kind == "sin_decay":
        dr = kw["decay_rate"]
        if vec:
            ph = kw.get("phase") or [0.0] * len(f)
            return sum(np.sin(w * dt * j + p) * np.exp(-l * dt * j) for w, l, p in zip(f, dr, ph))
        return np.sin(f * dt * j + kw.get("phase", 0.0)) * np.exp(-dr * dt * j)
    if kind == "abs_cos_power_p8":
        return np.abs(np.cos(2 * math.pi * dt * j)) ** kw.get("power", 0.8)
    raise ValueError(f"Unsupported signal kind: {kind}")


# --------------------------------------------------------------------------------------------
# signal -> MPS  (src/signals/SignalConverters.jl, src/linalg/rsvd.jl)
# --------------------------------------------------------------------------------------------
def array_to_tensor(x):
    """_array_to_tensor (SignalConverters.jl:16-46): n = round(log2 N), zero-pad, normalise."""
    x = np.asarray(x)
    N = x.size
    n = int(round(math.log2(N)))
    if N < 2**n:
        xp = np.zeros(2**n, dtype=np.result_type(x.dtype, np.float64))
        xp[:N] = x
        x = xp
    if x.size != 2**n:
        raise AssertionError("_array_to_tensor: Length of signal vector must be a power of 2")
    x = x.astype(np.complex128 if np.iscomplexobj(x) else np.float64)
    c = float(np.linalg.norm(x))
    return x / c, c, n


def tt_svd(x, cutoff=1e-15, maxdim=BIG):
    """signal_mps(x; method=:svd) (SignalConverters.jl:49-104, 228-233) -> (cores, amplitude)."""
    xn, c, n = array_to_tensor(x)
    if n == 1:
        return [xn.reshape(1, 2, 1)], c
    cores = []
    cur = xn.reshape(1, -1)
    chi = 1
    for _ in range(n - 1):
        M = cur.reshape(chi * 2, -1)
        U, S, Vh = svd_trunc(M, cutoff, maxdim)
        r = S.size
        cores.append(U.reshape(chi, 2, r))
        cur = S[:, None] * Vh
        chi = r
    cores.append(cur.reshape(chi, 2, 1))
    return cores, c


def gaussian_omega(rows, cols, iscomplex, seed=1234):
    """Stand-in for Random.seed!(seed); random_itensor(eltype, cR, alpha) (rsvd.jl:74-76).

    The Julia stream cannot be reproduced here; any N(0,1) matrix gives the same result to rounding
    whenever k+p >= numerical rank.  Complex entries are (N(0,1) + i N(0,1)) / sqrt(2).
    """
    rng = np.random.default_rng(seed)
    if iscomplex:
        return (rng.standard_normal((rows, cols)) + 1j * rng.standard_normal((rows, cols))) / math.sqrt(2.0)
    return rng.standard_normal((rows, cols))


def rsvd(A, k=20, p=10, q=0, random_seed=1234, cutoff=1e-15, maxdim=None, mindim=1, omega=None):
    """rsvd (rsvd.jl:38-121) on a matrix -> U (m x r), S (r), Vh (r x n)."""
    m, n = A.shape
    if m == 0 or n == 0:
        raise RuntimeError("In `rsvd`, left or right index set is empty.")
    if maxdim is None:
        maxdim = k
    l = min(k + p, m, n)
    Om = gaussian_omega(n, l, np.iscomplexobj(A), random_seed) if omega is None else omega[:n, :l]
    Q, _ = qr_thin(A @ Om, positive=True)
    for _ in range(q):
        Qz, _ = qr_thin(A.conj().T @ Q, positive=True)
        Q, _ = qr_thin(A @ Qz, positive=True)
    B = Q.conj().T @ A
    Us, S, Vh = svd_trunc(B, cutoff, maxdim, mindim)
    return Q @ Us, S, Vh


def tt_rsvd(x, cutoff=1e-15, maxdim=BIG, omega_fn=None, **kw):
    """signal_mps(x; method=:rsvd, ...) (SignalConverters.jl:107-196): divide and conquer.

    kwargs k, p, q, random_seed, mindim are forwarded to rsvd; cutoff/maxdim always override (:133).
    """
    xn, c, n = array_to_tensor(x)
    cores = [None] * n
    if n == 1:
        return [xn.reshape(1, 2, 1)], c

    def rec(T, first, last):  # T[lb, 2^(last-first+1), rb]
        lb, _, rb = T.shape
        if first == last:
            cores[first] = T.reshape(lb, 2, rb)
            return
        mid = (first + last + 1) // 2 - 1  # 0-based form of mid = (first + last - 1) ÷ 2 (:161)
        nl = mid - first + 1
        A = T.reshape(lb * 2**nl, -1)
        om = None if omega_fn is None else omega_fn(A.shape[1], np.iscomplexobj(A))
        U, S, Vh = rsvd(A, cutoff=cutoff, maxdim=maxdim, omega=om, **kw)
        r = S.size
        rec(U.reshape(lb, 2**nl, r), first, mid)
        rec((S[:, None] * Vh).reshape(r, -1, rb), mid + 1, last)

    rec(xn.reshape(1, -1, 1), 0, n - 1)
    return cores, c


def signal_mps(x, method="svd", **kw):
    if method == "svd":
        return tt_svd(x, **kw)
    if method == "rsvd":
        return tt_rsvd(x, **kw)
    raise ValueError(f"tensor_to_mps: unknown method {method}. Use :svd or :rsvd.")


def ztmps_split(cores, cutoff=1e-10, maxdim=BIG):
    """Per-site copy-tensor split of signal_ztmps (SignalConverters.jl:258-277).

    Returns the flat 2n-site chain main1, copy1, main2, copy2, ... (the `_as_signal_2n` view,
    mps.jl:421-445): Amain[chi_l, 2, c], Acopy[c, 2, chi_r].
    """
    out = []
    for M in cores:
        l, _, r = M.shape
        T = np.zeros((l, 2, 2, r), dtype=M.dtype)
        T[:, 0, 0, :] = M[:, 0, :]
        T[:, 1, 1, :] = M[:, 1, :]
        U, S, Vh = svd_trunc(T.reshape(2 * l, 2 * r), cutoff, maxdim)
        c = S.size
        out.append(U.reshape(l, 2, c))
        out.append((S[:, None] * Vh).reshape(c, 2, r))
    return out


def signal_ztmps(x, cutoff=1e-10, maxdim=BIG, **kw):
    """signal_ztmps (SignalConverters.jl:247-283) -> (flat 2n cores, amplitude)."""
    cores, c = signal_mps(x, cutoff=cutoff, maxdim=maxdim, **kw)
    return ztmps_split(cores, cutoff, maxdim), c


# --------------------------------------------------------------------------------------------
# MPS algorithms (src/mps.jl)
# --------------------------------------------------------------------------------------------
def bonds_of(cores):
    return [int(c.shape[-1]) for c in cores[:-1]]


def coefficient(cores, amplitude, bits):
    """coefficient (mps.jl:669-678): <bits|psi> * amplitude by a left-to-right vector chain."""
    if len(bits) != len(cores):
        raise ValueError(f"coefficient: expected {len(cores)} entries, got {len(bits)}")
    v = np.ones(1, dtype=cores[0].dtype)
    for M, b in zip(cores, bits):
        if not 0 <= int(b) < M.shape[1]:
            raise ValueError(f"coefficient: bit value {b} outside [0,{M.shape[1] - 1}]")
        v = v @ M[:, int(b), :]
    return amplitude * v[0]


def coefficient_batch(cores, amplitude, bits):
    """Vectorised `coefficient` over a (B, n) bit matrix (same arithmetic per row)."""
    bits = np.asarray(bits)
    B = bits.shape[0]
    dt = np.result_type(*[c.dtype for c in cores])
    v = np.ones((B, 1), dtype=dt)
    for i, M in enumerate(cores):
        b = bits[:, i].astype(bool)
        nv = np.empty((B, M.shape[2]), dtype=dt)
        nv[~b] = v[~b] @ M[:, 0, :]
        nv[b] = v[b] @ M[:, 1, :]
        v = nv
    return amplitude * v[:, 0]


def bits_from_integer(value, n):
    """_bits_from_integer (mps.jl:633-645): big-endian, site 1 = MSB."""
    if value < 0:
        raise ValueError("coefficient: integer configuration must be non-negative")
    bits = [(value >> (n - 1 - i)) & 1 for i in range(n)]
    if value >> n:
        raise ValueError(f"coefficient: integer {value} requires more than {n} bits")
    return bits


def mps_to_vector(cores, amplitude=1.0, reverse=False):
    """mps_to_vector (mps.jl:716-729): MSB-first by default, bit-reversed with reverse=true."""
    T = cores[0].reshape(-1, cores[0].shape[2])
    for M in cores[1:]:
        T = (T @ M.reshape(M.shape[0], -1)).reshape(-1, M.shape[2])
    v = T.reshape(-1)
    if reverse:
        n = len(cores)
        v = v.reshape((2,) * n).transpose(tuple(range(n - 1, -1, -1))).reshape(-1)
    return v * amplitude


def coefficient_grid(cores, amplitude, site_mode, out_bit=None):
    """Every combination of the free sites (site_mode == 2), each evaluated by the reference's chain
    (mps.jl:669-678), i.e. the loops of docs/src/tutorials/zt.jl:152-157.  Entry j of out_bit is the output-index
    bit of the j-th free site; default big-endian (mps.jl:633-645)."""
    mode = [int(m) for m in site_mode]
    free = [i for i, m in enumerate(mode) if m == 2]
    F = len(free)
    ob = list(range(F - 1, -1, -1)) if out_bit is None else [int(b) for b in out_bit]
    idx = np.arange(2**F, dtype=np.int64)
    bits = np.empty((2**F, len(cores)), dtype=np.uint8)
    for i, m in enumerate(mode):
        if m != 2:
            bits[:, i] = m
    for j, site in enumerate(free):
        bits[:, site] = (idx >> ob[j]) & 1
    return coefficient_batch(cores, amplitude, bits)


def mps_norm(cores):
    """norm (mps.jl:754-765): sqrt(|<psi|psi>|) by a transfer-matrix chain; ignores amplitude."""
    E = np.ones((1, 1), dtype=np.complex128)
    for M in cores:
        E = np.einsum("ab,asc,bsd->cd", E, M, M.conj())
    return math.sqrt(abs(E[0, 0]))


def canonicalize(cores, direction, center=None, cutoff=1e-12, maxdim=BIG):
    """canonicalize! (mps.jl:787-840). direction 'right' sweeps 1..c-1, 'left' sweeps N..c+1."""
    if direction not in ("right", "left"):
        raise ValueError("Direction must be :right or :left")
    N = len(cores)
    cores = [c.copy() for c in cores]
    if direction == "right":
        c = N if center is None else center
        if not 1 <= c <= N:
            raise IndexError(f"Center out of range [1,{N}]")
        for i in range(c - 1):
            l, _, r = cores[i].shape
            L, R = factorize(cores[i].reshape(l * 2, r), "left", cutoff, maxdim)
            k = L.shape[1]
            cores[i] = L.reshape(l, 2, k)
            nxt = cores[i + 1]
            cores[i + 1] = (R @ nxt.reshape(nxt.shape[0], -1)).reshape(k, 2, nxt.shape[2])
    else:
        c = 1 if center is None else center
        if not 1 <= c <= N:
            raise IndexError(f"Center out of range [1,{N}]")
        for i in range(N - 1, c - 1, -1):
            l, _, r = cores[i].shape
            L, R = factorize(cores[i].reshape(l, 2 * r), "right", cutoff, maxdim)
            k = L.shape[1]
            cores[i] = R.reshape(k, 2, r)
            prv = cores[i - 1]
            cores[i - 1] = (prv.reshape(-1, l) @ L).reshape(prv.shape[0], 2, k)
    return cores


def compress(cores, amplitude, maxdim=BIG, tol=1e-12, sweeps=1):
    """compress! (mps.jl:913-973) -> (cores, amplitude)."""
    N = len(cores)
    if N < 2:
        raise IndexError("SignalMPS must have at least 2 sites.")
    cutoff = tol**2 / ((N - 1) * sweeps)
    cores = canonicalize(cores, "left")
    for _ in range(sweeps):
        for j in range(N - 1):
            l = cores[j].shape[0]
            r = cores[j + 1].shape[2]
            th = cores[j].reshape(l * 2, -1) @ cores[j + 1].reshape(-1, 2 * r)
            U, S, Vh = svd_trunc(th, cutoff, maxdim)
            k = S.size
            cores[j] = U.reshape(l, 2, k)
            cores[j + 1] = (S[:, None] * Vh).reshape(k, 2, r)
        for j in range(N - 2, -1, -1):
            l = cores[j].shape[0]
            r = cores[j + 1].shape[2]
            th = cores[j].reshape(l * 2, -1) @ cores[j + 1].reshape(-1, 2 * r)
            U, S, Vh = svd_trunc(th, cutoff, maxdim)
            k = S.size
            cores[j] = (U * S[None, :]).reshape(l, 2, k)
            cores[j + 1] = Vh.reshape(k, 2, r)
    cores = canonicalize(cores, "left")
    nrm = mps_norm(cores)
    if nrm != 0:
        amplitude = amplitude * nrm
        cores[0] = cores[0] * (1.0 / nrm)
    return cores, amplitude


# --------------------------------------------------------------------------------------------
# apply (src/linalg/apply.jl)
# --------------------------------------------------------------------------------------------
def apply_mpo_mps(W, psi):
    """apply(W::SingleSiteMPO, psi::SignalMPS) (apply.jl:75-122): exact, W bond fastest."""
    if len(W) != len(psi):
        raise ValueError("apply: MPO and MPS must have the same number of sites.")
    out = []
    for Wi, Mi in zip(W, psi):
        a, _, _, b = Wi.shape
        l, _, r = Mi.shape
        T = np.einsum("apsb,lpr->lasrb", Wi, Mi)
        out.append(T.reshape(l * a, 2, r * b))
    return out


def right_canonical(cores):
    """Orthogonality centre to site 1 by untruncated SVDs from the right (each site i > 1 becomes Vh)."""
    cores = [c.copy() for c in cores]
    for i in range(len(cores) - 1, 0, -1):
        l, _, r = cores[i].shape
        U, S, Vh = svd_trunc(cores[i].reshape(l, 2 * r), 0.0, BIG, 1)
        k = S.size
        cores[i] = Vh.reshape(k, 2, r)
        prv = cores[i - 1]
        cores[i - 1] = (prv.reshape(-1, l) @ (U * S)).reshape(prv.shape[0], 2, k)
    return cores


def apply_zipup(W, psi, cutoff=1e-12, maxdim=BIG):
    """Truncating MPO x MPS (zip-up; NOT a reference function -- SURVEY.md 8f-3): right-canonical psi, then a
    left-to-right sweep T[b,s,d,c] = sum C[b,a,l] psi_i[l,p,c] W_i[a,p,s,d], (b s) x (d c) = U S Vh with the path's
    truncation rule, core_i = U, carry = S Vh.  Restated here so that the CUDA sweep has a checker."""
    if len(W) != len(psi):
        raise ValueError("apply: MPO and MPS must have the same number of sites.")
    psi = right_canonical(psi)
    n = len(psi)
    C = np.ones((1, 1, 1))
    out = []
    for i in range(n):
        X = np.einsum("bal,lpc->bapc", C, psi[i])
        T = np.einsum("bapc,apsd->bsdc", X, W[i])
        b, _, d, c = T.shape
        if i == n - 1:
            out.append(T.reshape(b, 2, 1))
        else:
            U, S, Vh = svd_trunc(T.reshape(b * 2, d * c), cutoff, maxdim, 1)
            r = S.size
            out.append(U.reshape(b, 2, r))
            C = (S[:, None] * Vh).reshape(r, d, c)
    return out


def apply_mpo_mpo(W1, W2, start1=0, start2=0):
    """apply(W1, W2) (apply.jl:124-199): W1 acts first; fused bond has the W1 bond fastest.

    Cores carry explicit dim-1 boundary bonds.  `start1`/`start2` give the first matching site of
    each operand (0-based); the longer operand is the base whose non-overlapping cores are kept.
    """
    n1, n2 = len(W1), len(W2)
    match = min(n1 - start1, n2 - start2)
    if n1 >= n2:
        base, bstart = [w.copy() for w in W1], start1
    else:
        base, bstart = [w.copy() for w in W2], start2
    for i in range(match):
        A = W1[start1 + i]
        B = W2[start2 + i]
        a, _, _, b = A.shape
        c, _, _, d = B.shape
        T = np.einsum("apmb,cmsd->capsdb", A, B)
        base[bstart + i] = T.reshape(c * a, 2, 2, d * b)
    return base


# --------------------------------------------------------------------------------------------
# gates (src/circuits/*.jl).  ctl(g0, g1): bond-diagonal core with g0 on bond value 1, g1 on value 2.
# --------------------------------------------------------------------------------------------
_I2 = np.eye(2)
_H = np.array([[1.0, 1.0], [1.0, -1.0]]) / math.sqrt(2.0)


def _P(theta):
    return np.diag([1.0, np.exp(-1j * theta)])


def _R(f):
    return np.diag([1.0, math.exp(-f)])


def _Hd(wr):
    return np.array([[1.0, 1.0], [1.0, math.exp(-wr / 2.0)]]) / math.sqrt(2.0)


def _ctl(g0, g1, left=True, right=True, dtype=np.float64):
    W = np.zeros((2 if left else 1, 2, 2, 2 if right else 1), dtype=dtype)
    for b, g in enumerate((g0, g1)):
        W[b if left else 0, :, :, b if right else 0] += g
    return W


def control_hphase_mpo(k):
    """control_Hphase_mpo (qft_gates.jl:43-97)."""
    if k == 1:
        return [_H.reshape(1, 2, 2, 1).astype(np.complex128)]
    cores = []
    W = np.zeros((1, 2, 2, 2), dtype=np.complex128)
    for b in range(2):
        W[0, :, b, b] = _H[:, b]  # W[p, s, b] = H[p, s] * [s == b]
    cores.append(W)
    for l in range(2, k):
        cores.append(_ctl(_I2, _P(2 * math.pi / 2.0**l), dtype=np.complex128))
    cores.append(_ctl(_I2, _P(2 * math.pi / 2.0**k), right=False, dtype=np.complex128))
    return cores


def control_damping_mpo(k, wr):
    """control_damping_mpo (dt_gates.jl:30-130) on 2k interleaved sites (main1, copy1, ...)."""
    if k == 1:
        return [_Hd(wr).reshape(1, 2, 2, 1), _I2.reshape(1, 2, 2, 1).copy()]
    cores = []
    for l in range(1, k):
        cores.append(_ctl(_I2, _R(wr * 2.0 ** (l - k - 1)), left=(l > 1)))
        cores.append(_ctl(_I2, _I2))
    W = np.zeros((2, 2, 2, 2))
    Hd = _Hd(wr)
    for b in range(2):
        W[b, b, :, b] = Hd[b, :]  # [p == b] * Hd[p, s]
    cores.append(W)
    cores.append(_ctl(_I2, _I2, right=False))
    return cores


def control_damping_copy_mpo(n, k, wr):
    """control_damping_copy_mpo (dt_gates.jl:133-229) on sites 2k-1..2n; L = n-k+1 pairs."""
    L = n - k + 1
    if L == 1:
        return [_I2.reshape(1, 2, 2, 1).copy(), _I2.reshape(1, 2, 2, 1).copy()]
    cores = []
    W = np.zeros((1, 2, 2, 2))
    W[0, :, :, 0] = _I2
    cores.append(W)
    W = np.zeros((2, 2, 2, 2))
    for b in range(2):
        W[0, b, b, b] = 1.0  # projector |b><b| on copy[1], left bond value 1, right bond value b
    cores.append(W)
    for j in range(2, L + 1):
        cores.append(_ctl(_I2, _R(wr * 2.0 ** (j - 2))))
        cores.append(_ctl(_I2, _I2, right=(j < L)))
    return cores


def control_hphase_ztmps_mpo(k):
    """control_Hphase_ztmps_mpo (zt_gates.jl:12-114) on 2k interleaved sites."""
    cd = np.complex128
    if k == 1:
        return [_I2.reshape(1, 2, 2, 1).astype(cd), _H.reshape(1, 2, 2, 1).astype(cd)]
    cores = []
    W = np.zeros((1, 2, 2, 2), dtype=cd)
    W[0, :, :, 0] = _I2
    W[0, :, :, 1] = _I2
    cores.append(W)
    cores.append(_ctl(_I2, _P(2 * math.pi / 2.0**k), dtype=cd))
    for j in range(2, k):
        cores.append(_ctl(_I2, _I2, dtype=cd))
        cores.append(_ctl(_I2, _P(2 * math.pi / 2.0 ** (k - j + 1)), dtype=cd))
    cores.append(_ctl(_I2, _I2, dtype=cd))
    W = np.zeros((2, 2, 2, 1), dtype=cd)
    for b in range(2):
        W[b, b, :, 0] = _H[b, :]  # [p == b] * H[p, s]
    cores.append(W)
    return cores


# --------------------------------------------------------------------------------------------
# MPO builders (src/transforms/*.jl)
# --------------------------------------------------------------------------------------------
def mpo_bonds(W):
    return [int(w.shape[3]) for w in W[:-1]]


def build_qft_mpo(n, cutoff=1e-14, maxdim=1000):
    """build_qft_mpo (qft_transformer.jl:121-165): zip-up (QR) then zip-down (truncated SVD)."""
    if n == 1:
        return control_hphase_mpo(1)
    qft = control_hphase_mpo(n)
    for it in range(1, n):
        m2 = control_hphase_mpo(n - it)
        L1, L2 = n, n - it
        new = [w for w in qft]
        # zip_up_mpos (:13-66): bottom -> top; T carries (l1, l2, q)
        T = np.ones((1, 1, 1), dtype=np.complex128)
        for irev in range(L2):
            i1, i2 = L1 - 1 - irev, L2 - 1 - irev
            core = np.einsum("apmb,cmsd,bdq->acpsq", qft[i1], m2[i2], T)
            a, c, _, _, qd = core.shape
            Mt = core.reshape(a * c, 4 * qd).T
            Q, R = np.linalg.qr(Mt, mode="reduced")
            kk = Q.shape[1]
            new[i1] = Q.T.reshape(kk, 2, 2, qd)
            T = R.T.reshape(a, c, kk)
        top = L1 - L2 - 1
        new[top] = np.einsum("lpsa,aq->lpsq", new[top], T[:, 0, :])
        # zip_down_mpos (:69-101): oc = it (1-based) ... L-1
        for kx in range(it - 1, n - 1):
            l, _, _, r = new[kx].shape
            U, S, Vh = svd_trunc(new[kx].reshape(l * 4, r), cutoff, maxdim)
            kk = S.size
            new[kx] = U.reshape(l, 2, 2, kk)
            nxt = new[kx + 1]
            new[kx + 1] = ((S[:, None] * Vh) @ nxt.reshape(r, -1)).reshape(kk, 2, 2, nxt.shape[3])
        qft = new
    return qft


def _combine_down(M1, M2):
    """zip_to_combine_mpos, "down" branch (dt_transformer.jl:38-95); M1 acts first, then M2."""
    n1, n2 = len(M1), len(M2)
    new = [w for w in M1]
    T = np.ones((1, 1, 1), dtype=np.result_type(M1[0].dtype, M2[0].dtype))
    for kx in range(n2):
        core = np.einsum("qac,apmb,cmsd->qpsbd", T, M1[kx], M2[kx])
        qd, _, _, b, d = core.shape
        if kx == n2 - 1 and n1 == n2:
            new[kx] = core.reshape(qd, 2, 2, 1)  # empty right index set: Q*R re-multiplied (:73,:90-94)
            T = None
            break
        Q, R = np.linalg.qr(core.reshape(qd * 4, b * d), mode="reduced")
        kk = Q.shape[1]
        new[kx] = Q.reshape(qd, 2, 2, kk)
        T = R.reshape(kk, b, d)
    if T is not None:
        new[n2] = np.einsum("qa,apsb->qpsb", T[:, :, 0], new[n2])
    return new


def _combine_up(M1, M2):
    """zip_to_combine_mpos, "up" branch (dt_transformer.jl:97-153)."""
    n1, n2 = len(M1), len(M2)
    new = [w for w in M1]
    T = np.ones((1, 1, 1), dtype=np.result_type(M1[0].dtype, M2[0].dtype))
    for kx in range(n2):
        i1, i2 = n1 - 1 - kx, n2 - 1 - kx
        core = np.einsum("apmb,cmsd,bdq->acpsq", M1[i1], M2[i2], T)
        a, c, _, _, qd = core.shape
        Q, R = np.linalg.qr(core.reshape(a * c, 4 * qd).T, mode="reduced")
        kk = Q.shape[1]
        new[i1] = Q.T.reshape(kk, 2, 2, qd)
        T = R.T.reshape(a, c, kk)
    tgt = n1 - n2 - 1
    new[tgt] = np.einsum("lpsa,aq->lpsq", new[tgt], T[:, 0, :])
    return new


def _compress_mpo(M, direction, cutoff, maxdim):
    """zip_to_compress_mpo (dt_transformer.jl:167-288) over the full chain."""
    L = len(M)
    if L < 2:
        return M
    new = [w for w in M]
    if direction == "down":
        for i in range(L - 1):
            l, _, _, r = new[i].shape
            Q, R = np.linalg.qr(new[i].reshape(l * 4, r), mode="reduced")
            kk = Q.shape[1]
            new[i] = Q.reshape(l, 2, 2, kk)
            nxt = new[i + 1]
            new[i + 1] = (R @ nxt.reshape(r, -1)).reshape(kk, 2, 2, nxt.shape[3])
        for i in range(L - 1, 0, -1):
            l = new[i - 1].shape[0]
            r = new[i].shape[3]
            th = new[i - 1].reshape(l * 4, -1) @ new[i].reshape(-1, 4 * r)
            U, S, Vh = svd_trunc(th, cutoff, maxdim)
            kk = S.size
            new[i] = Vh.reshape(kk, 2, 2, r)
            new[i - 1] = (U * S[None, :]).reshape(l, 2, 2, kk)
    elif direction == "up":
        for i in range(L - 1, 0, -1):
            l, _, _, r = new[i].shape
            Q, R = np.linalg.qr(new[i].reshape(l, 4 * r).T, mode="reduced")
            kk = Q.shape[1]
            new[i] = Q.T.reshape(kk, 2, 2, r)
            prv = new[i - 1]
            new[i - 1] = (prv.reshape(-1, l) @ R.T).reshape(prv.shape[0], 2, 2, kk)
        for i in range(L - 1):
            l = new[i].shape[0]
            r = new[i + 1].shape[3]
            th = new[i].reshape(l * 4, -1) @ new[i + 1].reshape(-1, 4 * r)
            U, S, Vh = svd_trunc(th, cutoff, maxdim)
            kk = S.size
            new[i] = U.reshape(l, 2, 2, kk)
            new[i + 1] = (S[:, None] * Vh).reshape(kk, 2, 2, r)
    else:
        raise ValueError(f"zip_to_compress_mpo: Unknown direction '{direction}'.")
    return new


def _extend_identity_pair(M):
    dt = M[0].dtype
    return M + [np.eye(2, dtype=dt).reshape(1, 2, 2, 1).copy(), np.eye(2, dtype=dt).reshape(1, 2, 2, 1).copy()]


def build_dt_mpo(n, wr, cutoff=1e-14, maxdim=1000):
    """build_dt_mpo (dt_transformer.jl:312-407) -> flat 2n cores (main1, copy1, ...)."""
    if n == 1:
        return control_damping_mpo(1, wr)
    M = control_damping_mpo(1, wr)
    for k in range(2, n + 1):
        M = _extend_identity_pair(M)
        M = _combine_down(M, control_damping_mpo(k, wr))
        M = _compress_mpo(M, "down", cutoff, maxdim)
    for k in range(1, n):
        blk = control_damping_copy_mpo(n, k, wr)
        M = _combine_down(M, blk) if len(blk) == len(M) else _combine_up(M, blk)
        M = _compress_mpo(M, "up", cutoff, maxdim)
    return M


def build_zt_mpo(n, wr, cutoff=1e-14, maxdim=1000):
    """build_zt_mpo (zt_transformer.jl:41-106)."""
    Wdt = build_dt_mpo(n, wr, cutoff, maxdim)
    if n == 1:
        return apply_mpo_mpo(Wdt, control_hphase_ztmps_mpo(1))
    Wq = control_hphase_ztmps_mpo(1)
    for k in range(2, n + 1):
        Wq = _extend_identity_pair(Wq)
        Wq = _combine_down(Wq, control_hphase_ztmps_mpo(k))
        Wq = _compress_mpo(Wq, "down", cutoff, maxdim)
    Wzt = apply_mpo_mpo(Wdt, Wq)
    return _compress_mpo(Wzt, "down", cutoff, maxdim)


# --------------------------------------------------------------------------------------------
# helpers used by the tests (mirrors of test/preamble_test.jl and the tutorials)
# --------------------------------------------------------------------------------------------
def bits_msb(v, n):
    return [(v >> (n - 1 - i)) & 1 for i in range(n)]


def bits_lsb(v, n):
    return [(v >> i) & 1 for i in range(n)]


def interleave(main_bits, copy_bits):
    out = []
    for a, b in zip(main_bits, copy_bits):
        out += [int(a), int(b)]
    return out


def bitrev(i, n):
    return int(format(i, f"0{n}b")[::-1], 2) if n > 0 else 0


def mpo_to_dense(W):
    """Dense operator O[out, in] with site 1 = MSB (test/preamble_test.jl:65-125)."""
    T = W[0]
    n = len(W)
    cur = T.reshape(T.shape[1], T.shape[2], T.shape[3])  # [p, s, r]
    P = cur.reshape(2, 2, -1)
    acc = P  # [pin, sout, r]
    for Wi in W[1:]:
        acc = np.einsum("xyr,rpsb->xpysb", acc, Wi)
        dx = acc.shape[0] * 2
        acc = acc.reshape(dx, dx, -1)
    return acc[:, :, 0].T  # [out, in]
