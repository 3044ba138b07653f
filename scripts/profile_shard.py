import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import reach_b200
from reach_b200 import _lib, workloads as W
from reach_b200.lgg09 import Lgg09Plan
from reach_b200.sharding import shard_directions
ctx = _lib.default_context(0)
wl = W.synthetic_c5(nsteps=140)
local_dirs, rows_by_rank, mloc = shard_directions(wl.dirs, 8, 0)
plan = Lgg09Plan(ctx, wl.n, mloc, 140, time_block=14)
plan.upload(wl.Phi, local_dirs, wl.Omega0, wl.V)
plan.run()
print(plan.stats())
