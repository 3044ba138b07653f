"""Short C5 run for ncu: a few passes of the main-loop kernel (n=2000, 2000 chains)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import reach_b200
from reach_b200 import _lib, workloads as W
from reach_b200.lgg09 import Lgg09Plan
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
path = sys.argv[3] if len(sys.argv) > 3 else "auto"
ctx = _lib.default_context(0)
wl = W.synthetic_c5()
plan = Lgg09Plan(ctx, wl.n, wl.dirs.shape[1], nsteps, time_block=S, path=path)
plan.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
plan.run()
print(plan.stats())
