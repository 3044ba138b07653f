"""Short C1 / C2 runs for ncu (small-n paths)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import reach_b200
from reach_b200 import _lib, workloads as W
from reach_b200.lgg09 import Lgg09Plan
which = sys.argv[1] if len(sys.argv) > 1 else "c1"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 0
path = sys.argv[3] if len(sys.argv) > 3 else "auto"
ctx = _lib.default_context(0)
wl = W.building_c1() if which == "c1" else W.iss_c2(nsteps=4096)
plan = Lgg09Plan(ctx, wl.n, wl.dirs.shape[1], wl.nsteps, time_block=S, path=path)
plan.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
plan.run()
plan.run()
print(plan.stats())
