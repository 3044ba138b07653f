"""Time-block sweep for one 8-GPU shard of C5 (250 chains) on the INT8 path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import reach_b200
from reach_b200 import _lib, workloads as W
from reach_b200.lgg09 import Lgg09Plan
from reach_b200.sharding import shard_directions
ctx = _lib.default_context(0)
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
wl = W.synthetic_c5(nsteps=nsteps)
local_dirs, rows_by_rank, mloc = shard_directions(wl.dirs, world, 0)
for S in (0, 14, 20, 28, 40, 56):
    plan = Lgg09Plan(ctx, wl.n, mloc, nsteps, time_block=S)
    plan.upload(wl.Phi, local_dirs, wl.Omega0, wl.V)
    plan.run(); plan.run()
    st = plan.stats()
    print("chains %d S=%d: %.2f ms, gemm %.3f ms x %d (share %.3f), launches %d" % (st["nchains"], st["time_block"], st["ms"], st["gemm_ms"], st["gemm_launches"], st["gemm_ms"] * st["gemm_launches"] / st["ms"], st["launches"]), flush=True)
    del plan
