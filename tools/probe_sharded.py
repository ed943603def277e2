"""Bond lists of the row-sharded and the single-GPU encode of the bench signal (torchrun --nproc-per-node G)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
import qilaplace_b200 as q
from qilaplace_b200 import parallel
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
N = 2**n
dev = torch.device("cuda", rank)
ctx = q.Context(rank, stream=torch.cuda.current_stream().cuda_stream)
j = torch.arange(N, dtype=torch.float64, device=dev)
t = j * (1.0 / (2.5 * N))
x = torch.sin(1.0 * t) * torch.exp(-0.08 * t) + torch.sin(2.5 * t) * torch.exp(-0.03 * t)
del j, t
ctx.truncation_margin(reset=True)
one = q.signal_mps_dev(ctx, x.data_ptr(), N, False, method="rsvd", **bench.ALGO)
m1 = ctx.truncation_margin(reset=True)
comm = parallel.PeerComm(ctx, parallel.encode_exchange_bytes(N, bench.ALGO["k"], bench.ALGO["p"], False))
NL = N // world
xl = x[rank * NL:(rank + 1) * NL].contiguous()
ps = parallel.signal_mps_sharded_dev(comm, xl.data_ptr(), N, False, **bench.ALGO)
m2 = ctx.truncation_margin(reset=True)
if rank == 0:
    print("env", {k: v for k, v in os.environ.items() if k.startswith("QIL_")})
    print("single ", one.bonds, "margin %.4f" % m1)
    print("sharded", ps.bonds, "margin %.4f" % m2, "equal", one.bonds == ps.bonds)
comm.close()
dist.destroy_process_group()
