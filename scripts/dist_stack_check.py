"""torchrun check: the row stack built by all ranks together equals the one a plan builds alone (bit for bit)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import reach_b200
from reach_b200 import _lib, workloads as W
from reach_b200.lgg09 import Lgg09Plan
from reach_b200.sharding import shard_directions, build_stack_distributed
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = _lib.Context(lr)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
S = int(sys.argv[3]) if len(sys.argv) > 3 else 40
wl = W.synthetic_c5(n=n, nsteps=nsteps)
local_dirs, rows_by_rank, mloc = shard_directions(wl.dirs, world, rank)


def stack_of(plan):
    ptr, nbe, ld, S_ = plan.stack_info()

    class H:
        __cuda_array_interface__ = {"shape": ((S_ + 1) * nbe, ld), "typestr": "<f8", "data": (int(ptr), False), "version": 3, "strides": None}
    return torch.as_tensor(H(), device=torch.device("cuda", lr)), nbe


a = Lgg09Plan(ctx, n, mloc, nsteps, time_block=S)
a.upload(wl.Phi, local_dirs, wl.Omega0, wl.V)
ra = a.run().sfmat().copy()
sa, nbe = stack_of(a)
sa = sa.clone()
b = Lgg09Plan(ctx, n, mloc, nsteps, time_block=S)
b.upload(wl.Phi, local_dirs, wl.Omega0, wl.V)
steps = build_stack_distributed(b)
sb, _ = stack_of(b)
sb = sb.clone()
rb_ = b.run().sfmat()
d = (sa - sb).abs()
bad = (d > 0).any(dim=1).nonzero().flatten()
print("rank %d: %d steps, stack max|diff| %.3e, differing rows %d (first %s, block %s), rho equal %s, rho max|diff| %.3e"
      % (rank, steps, float(d.max()), len(bad), int(bad[0]) if len(bad) else None, int(bad[0]) // nbe if len(bad) else None,
         np.array_equal(ra, rb_), float(np.abs(ra - rb_).max())), flush=True)
for it in range(3):                                       # the plan is re-run: every run rebuilds the stack with all ranks
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float64, device="cuda")
    dist.barrier()
    torch.cuda.synchronize()
    build_stack_distributed(b)
    torch.cuda.synchronize()
    sb2, _ = stack_of(b)
    sb2 = sb2.clone()
    r2 = b.run().sfmat()
    d2 = (sa - sb2).abs()
    bad = (d2 > 0).any(dim=1).nonzero().flatten()
    print("rank %d rerun %d: stack max|diff| %.3e rows %d (first %s), rho equal %s max|diff| %.3e" % (
        rank, it, float(d2.max()), len(bad), int(bad[0]) if len(bad) else None, np.array_equal(ra, r2), float(np.abs(ra - r2).max())), flush=True)
    del flush
r3 = b.run().sfmat()                                       # the same plan, building the stack itself again
sb3, _ = stack_of(b)
print("rank %d local run after the shared builds: rho equal %s max|diff| %.3e, stack max|diff| %.3e" % (
    rank, np.array_equal(ra, r3), float(np.abs(ra - r3).max()), float((sa - sb3).abs().max())), flush=True)
r4 = a.run().sfmat()
print("rank %d plan a again: rho equal %s" % (rank, np.array_equal(ra, r4)), flush=True)
dist.destroy_process_group()
