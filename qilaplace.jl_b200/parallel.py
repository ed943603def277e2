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
