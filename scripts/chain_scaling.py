import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import reach_b200
from reach_b200 import _lib, workloads as W
from reach_b200.lgg09 import Lgg09Plan
ctx = _lib.default_context(0)
wl = W.iss_c2(nsteps=4096)
n = wl.n
for mu in (1, 8, 64, 148, 270):
    dirs = np.hstack([np.eye(n)[:, :mu], -np.eye(n)[:, :mu]])
    plan = Lgg09Plan(ctx, n, 2 * mu, wl.nsteps, path="chain")
    plan.upload(wl.Phi, dirs, wl.Omega0, wl.V)
    plan.run(); plan.run()
    st = plan.stats()
    print("chains %d: %.3f ms -> %.0f ns per step" % (mu, st["ms"], st["ms"] * 1e6 / wl.nsteps), flush=True)
