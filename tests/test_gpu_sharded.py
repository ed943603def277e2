"""Row-sharded encode of ONE signal (qil_encode_rsvd_sharded_dev + the torch.distributed collectives of
qilaplace_b200.parallel.TorchComm) against the single-device CUDA encode and the CPU oracle.

On a one-GPU box the ranks share cuda:0 and the process group is gloo (NCCL refuses two ranks on one device), which
drives the identical library code through the host-staged variant of the two callbacks; with >= 2 GPUs the group
is NCCL, one device per rank."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _signal(n, cplx):
    N = 2**n
    t = np.arange(N) / (2.5 * N)
    x = np.sin(1.0 * t) * np.exp(-0.08 * t) + np.sin(2.5 * t) * np.exp(-0.03 * t)
    return x * np.exp(0.3j * t) if cplx else x


def _free_port():
    """A port nobody listens on right now (a fixed pid-derived port collided once when the whole file ran back to back)."""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _worker(rank, world, port, ret, default_stream=False, peer=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import qil_oracle as O
    import qilaplace_b200 as q
    from qilaplace_b200 import parallel
    ngpu = torch.cuda.device_count()
    nccl = ngpu >= world
    dev = rank if nccl else 0
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if nccl else "gloo", rank=rank, world_size=world)
    try:
        # own stream, or torch's (legacy default) stream like bench.py
        ctx = q.Context(dev, stream=torch.cuda.current_stream().cuda_stream) if default_stream else q.Context(dev)
        if peer:      # the library's own exchange kernels over CUDA-IPC peer memory
            comm = parallel.PeerComm(ctx, parallel.encode_exchange_bytes(2**22, 15, 5, True))
        else:
            comm = parallel.TorchComm(ctx)
        for n, cplx, qit in ((20, False, 2), (21, True, 1), (22, False, 0)):
            N = 2**n
            x = _signal(n, cplx)
            lo, hi = rank * N // world, (rank + 1) * N // world
            xl = torch.from_numpy(np.ascontiguousarray(x[lo:hi])).to(f"cuda:{dev}")
            torch.cuda.synchronize()
            psi = parallel.signal_mps_sharded_dev(comm, xl.data_ptr(), N, cplx, k=15, p=5, q=qit, cutoff=1e-12)
            # single-device CUDA encode of the whole signal on this rank's device
            xf = torch.from_numpy(np.ascontiguousarray(x)).to(f"cuda:{dev}")
            torch.cuda.synchronize()
            one = q.signal_mps_dev(ctx, xf.data_ptr(), N, cplx, method="rsvd", k=15, p=5, q=qit, cutoff=1e-12)
            assert psi.bonds == one.bonds, (psi.bonds, one.bonds)
            assert abs(psi.amplitude - one.amplitude) < 1e-12 * one.amplitude
            v, v1 = q.mps_to_vector(psi), q.mps_to_vector(one)
            assert np.abs(v - v1).max() < 1e-10 * np.abs(v1).max()
            # CPU oracle (single process, reference algorithm)
            ref, cref = O.tt_rsvd(x, k=15, p=5, q=qit, cutoff=1e-12)
            assert psi.bonds == O.bonds_of(ref)
            vref = O.mps_to_vector(ref, cref)
            assert np.abs(v - vref).max() < 1e-10 * np.abs(vref).max()
            assert np.linalg.norm(v - x) < 1e-5 * np.linalg.norm(x)
        if peer:
            comm.close()
        else:
            assert comm.calls["allreduce"] > 0 and comm.calls["allgather"] > 0
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,default_stream,peer", [(1, False, False), (2, False, False), (2, True, False),
                                                       (1, False, True), (2, False, True), (2, True, True),
                                                       (4, False, False), (4, False, True)])
def test_row_sharded_encode_matches_single_device_and_oracle(world, default_stream, peer):
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret, default_stream, peer), nprocs=world, join=True)
    assert dict(ret) == {r: "ok" for r in range(world)}


def _worker_big(rank, world, port, ret, peer):
    """n = 26 bench-family signal over `world` real devices (NCCL / NVLink): bonds identical to the one-GPU encode."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import qilaplace_b200 as q
    from qilaplace_b200 import parallel
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ctx = q.Context(rank)
        n = 26
        N = 2**n
        kw = dict(k=15, p=5, q=2, cutoff=1e-12)
        comm = (parallel.PeerComm(ctx, parallel.encode_exchange_bytes(N, 15, 5, False)) if peer else parallel.TorchComm(ctx))
        j = torch.arange(N, dtype=torch.float64, device=f"cuda:{rank}")
        t = j / (2.5 * N)
        x = torch.sin(1.0 * t) * torch.exp(-0.08 * t) + torch.sin(2.5 * t) * torch.exp(-0.03 * t)
        del j, t
        lo, hi = rank * N // world, (rank + 1) * N // world
        xl = x[lo:hi].contiguous()
        torch.cuda.synchronize()
        one = q.signal_mps_dev(ctx, x.data_ptr(), N, False, method="rsvd", **kw)
        for rep in range(3):
            psi = parallel.signal_mps_sharded_dev(comm, xl.data_ptr(), N, False, **kw)
            assert psi.bonds == one.bonds, (rep, psi.bonds, one.bonds)
            assert abs(psi.amplitude - one.amplitude) < 1e-12 * one.amplitude
        rng = np.random.default_rng(0)
        idx = rng.integers(0, N, 2048)
        bits = ((idx[:, None] >> np.arange(n - 1, -1, -1)[None, :]) & 1).astype(np.uint8)
        a, b = q.coefficients(psi, bits), q.coefficients(one, bits)
        assert np.abs(a - b).max() < 1e-8 * np.abs(b).max()
        if peer:
            comm.close()
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [4, 8])
@pytest.mark.parametrize("peer", [False, True])
def test_row_sharded_encode_many_gpus_bonds_identical(world, peer):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_big, args=(world, port, ret, peer), nprocs=world, join=True)
    assert dict(ret) == {r: "ok" for r in range(world)}


def _worker_scan(rank, world, port, ret):
    """MPS broadcast (device to device over the group) + k-row-sharded pole scan with the 3-number arg-max exchange
    (docs/src/tutorials/zt.jl:296-326) against the one-device scan."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import math
    import torch
    import torch.distributed as dist
    import qilaplace_b200 as q
    from qilaplace_b200 import parallel
    ngpu = torch.cuda.device_count()
    nccl = ngpu >= world
    dev = rank if nccl else 0
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if nccl else "gloo", rank=rank, world_size=world)
    try:
        ctx = q.Context(dev)
        n = 12
        N = 2**n
        psi = None
        if rank == 0:
            j = np.arange(N)
            x = (1.0002 * np.exp(0.002j)) ** j * np.cos(0.11 * j)
            z = q.signal_ztmps(x, ctx=ctx, cutoff=1e-12)
            W = q.build_zt_mpo(z, 2 * math.pi, cutoff=1e-12, maxdim=128, ctx=ctx)
            psi = W * z
        out = parallel.broadcast_mps(psi, ctx, src=0)
        k, l, av = parallel.pole_scan_argmax_sharded(out, 0, 0, 6, 6, n - 6, n - 6)
        k1, l1, av1, _ = q.pole_scan_argmax(out, 0, 0, 6, 6, n - 6, n - 6)
        assert (k, l) == (k1, l1) and abs(av - av1) <= 1e-14 * av1
        gathered = [None] * world
        dist.all_gather_object(gathered, (k, l, av, out.bonds))
        assert all(g == gathered[0] for g in gathered)          # every rank holds the same MPS and the same peak
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_broadcast_mps_and_sharded_pole_scan(world):
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_scan, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {r: "ok" for r in range(world)}
