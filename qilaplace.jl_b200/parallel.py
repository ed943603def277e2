"""Sharding of the path's independent units across ranks (one process per GPU, torch.distributed).

The hot path shards without any data-path collective: independent signals are encoded by
different ranks, and a coefficient grid / pole scan is split by bitstring rows after one broadcast of the
(small) MPS.  Only the results are gathered.  The compute callable defaults to the CUDA path; the CPU test
suite passes its own callable to exercise this host logic over gloo."""
from __future__ import annotations

import numpy as np


def shard_range(total, world, rank):
    """Contiguous balanced split of `total` units: the first `total % world` ranks get one extra."""
    base, rem = divmod(int(total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist


def broadcast_cores(cores, amplitude, src=0, group=None):
    """Broadcast an MPS (list of numpy cores + amplitude) from `src` to every rank."""
    dist = _dist()
    payload = [cores, amplitude] if dist.get_rank(group) == src else [None, None]
    dist.broadcast_object_list(payload, src=src, group=group)
    return payload[0], payload[1]


def broadcast_mps(psi, ctx, src=0, group=None):
    """Device-to-device broadcast of an MPS over the process group (NCCL over NVLink on the GPU box): bonds and amplitude
    travel as a small object, every core goes straight from the source rank's HBM into a freshly allocated core on each
    receiver (qil_mps_alloc / qil_mps_core_ptr) -- no host staging of the tensor data.  `psi` is ignored off `src`."""
    import ctypes as C
    import torch
    from . import _lib, api
    dist = _dist()
    rank = dist.get_rank(group)
    meta = [None]
    if rank == src:
        meta = [(psi.nsites_flat, bool(psi.is_complex), [1] + list(psi.bonds) + [1], float(psi.amplitude),
                 isinstance(psi, api.ZTMPS))]
    dist.broadcast_object_list(meta, src=src, group=group)
    nsites, is_c, bond, amp, is_zt = meta[0]
    if rank == src:
        out = psi
    else:
        b = np.asarray(bond, dtype=np.int64)
        h = _lib.c_mps()
        _lib.call("qil_mps_alloc", ctx.handle, int(nsites), int(is_c), C.c_void_p(b.ctypes.data), float(amp), C.byref(h))
        out = (api.ZTMPS if is_zt else api.SignalMPS)(ctx, h)
    dev = torch.device("cuda", ctx.device)
    nccl = dist.get_backend(group) == "nccl"
    ctx.sync()
    for i in range(nsites):
        ptr, cnt = C.c_void_p(), C.c_int64()
        _lib.call("qil_mps_core_ptr", out.handle, i, C.byref(ptr), C.byref(cnt))
        t = torch.as_tensor(_DevBuf(ptr.value, cnt.value * (2 if is_c else 1)), device=dev)
        if nccl:
            dist.broadcast(t, src=src, group=group)
        else:                       # functional fallback (gloo): through the host
            hbuf = t.cpu()
            dist.broadcast(hbuf, src=src, group=group)
            if rank != src:
                t.copy_(hbuf)
    torch.cuda.synchronize(dev)
    return out


def pole_scan_argmax_sharded(psi, k0=0, l0=0, log2_k=None, log2_l=None, stride_log2_k=0, stride_log2_l=0, group=None):
    """Aligned (k, l) block split by k-rows over the ranks: rank g evaluates the 2^(log2_k - log2 G) rows
    whose top free k-bits equal g (grid + arg-max on its own device), then a 3-number all-gather picks the global peak
    (largest |chi|, lowest flat index on ties).  Every rank must hold the same MPS (see broadcast_mps)."""
    import torch
    from . import api
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = psi.nsites_flat // 2
    log2_k = n - stride_log2_k if log2_k is None else log2_k
    log2_l = n - stride_log2_l if log2_l is None else log2_l
    lg = int(np.log2(world))
    if (1 << lg) != world or lg > log2_k:
        raise api.ArgumentError("pole_scan_argmax_sharded: world size must be a power of two <= the number of k rows")
    sub = log2_k - lg
    k0_loc = k0 + (rank << (stride_log2_k + sub))
    k, l, av, v = api.pole_scan_argmax(psi, k0_loc, l0, sub, log2_l, stride_log2_k, stride_log2_l)
    flat = ((k - k0) >> stride_log2_k) * (1 << log2_l) + ((l - l0) >> stride_log2_l)
    mine = torch.tensor([av, float(flat), float(k), float(l)], dtype=torch.float64)
    if dist.get_backend(group) == "nccl":
        mine = mine.cuda(psi.ctx.device)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    best = max((t.cpu().tolist() for t in allv), key=lambda t: (t[0], -t[1]))
    return int(best[2]), int(best[3]), float(best[0])


def coefficients_sharded(cores, amplitude, bits, compute=None, group=None):
    """Every rank evaluates its contiguous slice of `bits` (B x n) and all ranks receive all B results."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bits = np.ascontiguousarray(bits)
    lo, hi = shard_range(bits.shape[0], world, rank)
    if compute is None:
        from . import api

        def compute(c, a, b):
            return api.coefficients(api.SignalMPS.from_cores(c, a), b)
    local = np.asarray(compute(cores, amplitude, bits[lo:hi]), dtype=np.complex128)
    sizes = [shard_range(bits.shape[0], world, r) for r in range(world)]
    maxlen = max(h - l for l, h in sizes)
    buf = torch.zeros(maxlen, dtype=torch.complex128)
    buf[: hi - lo] = torch.from_numpy(local)
    gathered = [torch.zeros(maxlen, dtype=torch.complex128) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    return np.concatenate([g.numpy()[: h - l] for g, (l, h) in zip(gathered, sizes)])


def argmax_abs_sharded(values_local, offset, group=None):
    """Global (|value|, index) maximum of a sharded result vector (pole-scan peak, docs/src/tutorials/zt.jl:300-310)."""
    import torch
    dist = _dist()
    mag = np.abs(np.asarray(values_local))
    if mag.size:
        i = int(mag.argmax())
        best = torch.tensor([float(mag[i]), float(offset + i)], dtype=torch.float64)
    else:
        best = torch.tensor([-1.0, -1.0], dtype=torch.float64)
    world = dist.get_world_size(group)
    allb = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allb, best, group=group)
    vals = [(float(b[0]), int(b[1])) for b in allb]
    v, idx = max(vals, key=lambda t: (t[0], -t[1]))
    return v, idx


# --------------------------------------------------------------------------------------------------
# One signal row-sharded over the ranks: the exchange steps of the divide-and-conquer top split
# (all-gather of the TSQR R factors, all-reduce of the partial projections) run through torch.distributed;
# the library calls back into the two collectives below with device pointers.
# --------------------------------------------------------------------------------------------------
class _DevBuf:
    """A raw device allocation exposed through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class TorchComm:
    """struct qil_comm backed by a torch.distributed process group (NCCL over NVLink on the GPU box).

    With the NCCL backend the collectives are enqueued relative to the context's stream (made torch's current
    stream for the duration of the callback), so no host synchronisation happens inside the encode.  Any other
    backend (gloo) is served through a host staging copy; it exists for functional tests only."""

    def __init__(self, ctx, group=None):
        import ctypes as C
        import torch
        from . import _lib
        dist = _dist()
        self.ctx, self.group = ctx, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cuda", ctx.device)
        self.nccl = dist.get_backend(group) == "nccl"
        sp = C.c_void_p()
        _lib.call("qil_get_stream", ctx.handle, C.byref(sp))
        # a NULL handle is the legacy default stream, which is torch's default stream; ExternalStream(0) would
        # silently hand out a pool stream with no ordering against the library's work
        self.stream = (torch.cuda.ExternalStream(sp.value, device=self.device) if sp.value
                       else torch.cuda.default_stream(self.device))
        self.calls = {"allreduce": 0, "allgather": 0, "bytes": 0}
        self.error = None

        def wrap(ptr, count):
            return torch.as_tensor(_DevBuf(ptr, count), device=self.device)

        def allreduce(_user, d_buf, count):
            try:
                with torch.cuda.stream(self.stream):
                    t = wrap(d_buf, count)
                    if self.nccl:
                        # all-gather + summation in RANK ORDER on every rank: the same bits everywhere and the same
                        # bits as the library's peer-memory exchange (qil_peer.cu) -- ncclAllReduce sums in an order
                        # that depends on the ring/tree it picks, which moved a knife-edge bond at 4 ranks in round 1
                        parts = torch.empty((self.world, int(count)), dtype=torch.float64, device=self.device)
                        dist.all_gather_into_tensor(parts.view(-1), t, group=group)
                        t.copy_(parts[0])
                        for r in range(1, self.world):
                            t.add_(parts[r])
                    else:
                        h = t.cpu()
                        dist.all_reduce(h, group=group)
                        t.copy_(h)
                self.calls["allreduce"] += 1
                self.calls["bytes"] += 8 * int(count)
                return 0
            except Exception as e:   # never let an exception cross the C boundary
                self.error = e
                return 1

        def allgather(_user, d_send, d_recv, count):
            try:
                with torch.cuda.stream(self.stream):
                    s = wrap(d_send, count)
                    r = wrap(d_recv, count * self.world)
                    if self.nccl:
                        dist.all_gather_into_tensor(r, s, group=group)
                    else:
                        hs = s.cpu()
                        parts = [torch.empty_like(hs) for _ in range(self.world)]
                        dist.all_gather(parts, hs, group=group)
                        r.copy_(torch.cat(parts))
                self.calls["allgather"] += 1
                self.calls["bytes"] += 8 * int(count) * self.world
                return 0
            except Exception as e:
                self.error = e
                return 1

        self._cb = (_lib.QilComm.ALLREDUCE(allreduce), _lib.QilComm.ALLGATHER(allgather))   # keep alive
        self.struct = _lib.QilComm(self.rank, self.world, None, self._cb[0], self._cb[1])


class PeerComm:
    """struct qil_comm implemented by the library itself over NVLink peer memory (qil_peer.cu): exchange buffers are
    shared between the ranks through CUDA IPC, the all-reduce / all-gather are the library's own kernels (the
    reduction runs inside the exchange kernel).  torch.distributed is used ONCE, to swap the 64-byte IPC handles."""

    def __init__(self, ctx, payload_bytes, group=None):
        import ctypes as C
        from . import _lib
        dist = _dist()
        self.ctx, self.group = ctx, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.error = None
        self.handle = C.c_void_p()
        mine = (C.c_ubyte * 64)()
        _lib.call("qil_peer_create", ctx.handle, self.rank, self.world, C.c_int64(int(payload_bytes)),
                  C.byref(self.handle), mine)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(mine), group=group)
        blob = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(handles))
        _lib.call("qil_peer_connect", self.handle, blob)
        self.struct = _lib.QilComm()
        _lib.call("qil_peer_comm", self.handle, C.byref(self.struct))
        dist.barrier(group=group)          # every rank is mapped before the first kernel raises a flag
        self.calls = {"native": True}

    def close(self):
        from . import _lib
        if self.handle:
            self.ctx.sync()                     # this rank's exchange kernels are done before it reports in ...
            _dist().barrier(group=self.group)   # ... so after the barrier nobody can still be reading a peer's buffer
            _lib.load().qil_peer_destroy(self.handle)
            self.handle = None


def encode_exchange_bytes(N_total, k, p, is_complex):
    """Payload a PeerComm needs for signal_mps_sharded_dev: the largest message is the C x l partial projection
    (or the R x r block of U, which is smaller)."""
    n = int(round(np.log2(N_total)))
    C = 2 ** (n - n // 2)
    return 16 * (2 if is_complex else 1) * (k + p) * C


def signal_mps_sharded_dev(comm, d_x_local, N_total, is_complex, cutoff=1e-15, maxdim=None, k=20, p=10, q=0,
                           random_seed=1234, mindim=1, adaptive=False):
    """signal_mps(x; method=:rsvd) for ONE signal whose rank-th contiguous chunk of N_total / world samples lives at
    device pointer d_x_local on this rank.  Every rank makes the call and receives the same SignalMPS."""
    import ctypes as C
    from . import _lib, api
    h = _lib.c_mps()
    try:
        _lib.call("qil_encode_rsvd_sharded_dev", comm.ctx.handle, C.byref(comm.struct), int(is_complex),
                  C.c_void_p(int(d_x_local)), C.c_int64(N_total), int(k), int(p), int(q), C.c_int64(random_seed),
                  float(cutoff), C.c_int64(api._maxdim_arg(maxdim)), C.c_int64(mindim), None, C.c_int64(0),
                  C.c_int64(api.RSVD_ADAPTIVE if adaptive else 0), C.byref(h))
    except Exception:
        if comm.error is not None:
            raise comm.error
        raise
    return api.SignalMPS(comm.ctx, h)


# --------------------------------------------------------------------------------------------------
# Streaming many host signals through one GPU: double-buffered upload on a copy stream, so that the PCIe transfer of
# signal i+1 overlaps the encode of signal i (the encode itself is 5-6 ms at n = 28, the 2 GiB upload ~40 ms).
# --------------------------------------------------------------------------------------------------
class SignalUploader:
    """Double-buffered host -> device staging for a stream of equally sized signals.

        up = SignalUploader(ctx, N, is_complex=False)
        up.submit(x0)                                  # pinned host array / tensor; returns immediately
        for x_next in ...:
            d_x = up.acquire()                         # device pointer of the oldest submitted signal (stream-ordered)
            up.submit(x_next)                          # its upload overlaps the work below
            psi = signal_mps_dev(ctx, d_x, N, ...)
            up.release()

    torch supplies the copy stream, the events and the device buffers; the context must live on torch's current
    stream (Context(device, stream=torch.cuda.current_stream().cuda_stream))."""

    def __init__(self, ctx, N, is_complex=False, depth=2):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.dev = torch.device("cuda", ctx.device)
        dt = torch.complex128 if is_complex else torch.float64
        self.bufs = [torch.empty(int(N), dtype=dt, device=self.dev) for _ in range(depth)]
        self.ready = [torch.cuda.Event() for _ in range(depth)]      # upload of buffer i finished
        self.freed = [torch.cuda.Event() for _ in range(depth)]      # consumer of buffer i finished
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.head = 0      # next buffer to fill
        self.tail = 0      # next buffer to consume
        self.inflight = 0
        self.used = [False] * depth

    def submit(self, x_host):
        torch = self.torch
        if self.inflight == len(self.bufs):
            raise RuntimeError("SignalUploader: every buffer is in flight; acquire/release one first")
        t = x_host if isinstance(x_host, torch.Tensor) else torch.from_numpy(x_host)
        i = self.head
        with torch.cuda.stream(self.copy_stream):
            if self.used[i]:
                self.copy_stream.wait_event(self.freed[i])   # the previous consumer of this buffer is done
            self.bufs[i].copy_(t, non_blocking=True)
            self.ready[i].record(self.copy_stream)
        self.head = (i + 1) % len(self.bufs)
        self.inflight += 1

    def acquire(self):
        if self.inflight == 0:
            raise RuntimeError("SignalUploader: nothing submitted")
        i = self.tail
        self.torch.cuda.current_stream(self.dev).wait_event(self.ready[i])
        return self.bufs[i].data_ptr()

    def release(self):
        i = self.tail
        self.freed[i].record(self.torch.cuda.current_stream(self.dev))
        self.used[i] = True
        self.tail = (i + 1) % len(self.bufs)
        self.inflight -= 1
