"""C2 (ISS) on the compact-chain path vs the persistent chain kernel: time and agreement."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import reach_b200
from reach_b200 import _lib, workloads as W
from reach_b200.lgg09 import Lgg09Plan
ctx = _lib.default_context(0)
wl = W.iss_c2()
res = {}
for path in ("compact", "chain"):
    plan = Lgg09Plan(ctx, wl.n, wl.dirs.shape[1], wl.nsteps, path=path)
    plan.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
    for _ in range(3):
        plan.run()
        st = plan.stats()
    res[path] = plan.sfmat()
    print(path, "%.3f ms, %d launches -> %.3e evals/s" % (st["ms"], st["launches"], wl.dirs.shape[1] * wl.nsteps / (st["ms"] * 1e-3)), flush=True)
    plan.close()
a, b = res["compact"], res["chain"]
scale = np.maximum(np.abs(b), np.abs(b).max(axis=1, keepdims=True))
print("max scaled diff compact vs chain: %.3e" % (np.abs(a - b) / scale).max())
