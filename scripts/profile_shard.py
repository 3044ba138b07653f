"""One shard of C5 (rank 0 of `world` GPUs), full length, for the ncu launch list (gpu__time_duration per kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import reach_b200
from reach_b200 import _lib, workloads as W
from reach_b200.lgg09 import Lgg09Plan
from reach_b200.sharding import shard_directions
ctx = _lib.default_context(0)
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
S = int(sys.argv[3]) if len(sys.argv) > 3 else 0
wl = W.synthetic_c5(nsteps=nsteps)
local_dirs, rows_by_rank, mloc = shard_directions(wl.dirs, world, 0)
plan = Lgg09Plan(ctx, wl.n, mloc, nsteps, time_block=S)
plan.upload(wl.Phi, local_dirs, wl.Omega0, wl.V)
plan.run()
print(plan.stats())
