"""CPU statement of the ROW-SHARDED top split of the divide-and-conquer encoder, built from the
oracle's pieces and torch.distributed collectives.  The reference has no distributed path, so this file is the
specification the CUDA implementation (qil_encode.cu: rsvd_split_sharded / tsqr_sharded) follows: same exchange
steps, same order.  Test infrastructure only (imports the oracle)."""
import numpy as np
import torch
import torch.distributed as dist

import qil_oracle as O


def _allreduce(a):
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.float64).copy())
    dist.all_reduce(t)
    return t.numpy().view(a.dtype).reshape(a.shape)


def _allgather_rows(a):
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.float64).copy())
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    return np.concatenate([p.numpy().view(a.dtype).reshape(a.shape) for p in parts], axis=0)


def tsqr_sharded(Yg):
    """Local thin QR, all-gather of the l x l R factors, QR of the stack (positive), local combine."""
    l = Yg.shape[1]
    Qg, Rg = O.qr_thin(Yg, positive=True)
    Q2, _ = O.qr_thin(_allgather_rows(Rg), positive=True)
    g = dist.get_rank()
    return Qg @ Q2[g * l:(g + 1) * l, :]


def tt_rsvd_sharded(x_local, n, k=20, p=10, q=0, cutoff=1e-15, maxdim=O.BIG, mindim=1, omega_fn=None):
    """Every rank passes its contiguous chunk of the length-2^n signal and gets the full list of cores + amplitude."""
    G = dist.get_world_size()
    mid = n // 2 - 1
    R, C = 2 ** (mid + 1), 2 ** (n - 1 - mid)
    A = np.asarray(x_local).reshape(R // G, C)
    ic = np.iscomplexobj(A)
    ss = _allreduce(np.array([np.vdot(A, A).real]))
    c = float(np.sqrt(ss[0]))
    l = min(k + p, R, C)
    Om = (O.gaussian_omega(C, l, ic) if omega_fn is None else omega_fn(C, ic))[:C, :l]
    Q = tsqr_sharded(A @ Om)
    for _ in range(q):
        Qz, _ = O.qr_thin(_allreduce(A.conj().T @ Q), positive=True)
        Q = tsqr_sharded(A @ Qz)
    B = _allreduce(A.conj().T @ Q).conj().T / c
    Us, S, Vh = O.svd_trunc(B, cutoff, maxdim, mindim)
    U = _allgather_rows(Q @ Us)
    r = S.size
    # the two subtrees are small: replicated on every rank with the single-process recursion of the oracle
    cores = [None] * n

    def rec(T, first, last):
        lb, _, rb = T.shape
        if first == last:
            cores[first] = T.reshape(lb, 2, rb)
            return
        m = (first + last + 1) // 2 - 1
        nl = m - first + 1
        M = T.reshape(lb * 2**nl, -1)
        om = None if omega_fn is None else omega_fn(M.shape[1], np.iscomplexobj(M))
        U2, S2, Vh2 = O.rsvd(M, k=k, p=p, q=q, cutoff=cutoff, maxdim=maxdim, mindim=mindim, omega=om)
        rec(U2.reshape(lb, 2**nl, S2.size), first, m)
        rec((S2[:, None] * Vh2).reshape(S2.size, -1, rb), m + 1, last)

    rec(U.reshape(1, R, r), 0, mid)
    rec((S[:, None] * Vh).reshape(r, C, 1), mid + 1, n - 1)
    return cores, c
