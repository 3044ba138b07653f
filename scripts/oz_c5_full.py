"""Full-length C5 (10^4 steps): DMMA path vs INT8 digit-plane path, accumulated difference."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import reach_b200
from reach_b200 import _lib, workloads as W
from reach_b200.lgg09 import Lgg09Plan
ctx = _lib.default_context(0)
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
wl = W.synthetic_c5(nsteps=nsteps)
res = {}
for (path, S, T) in (("gemm", 0, 0), ("int8", 8, 8), ("int8", 8, 7)):
    plan = Lgg09Plan(ctx, wl.n, wl.dirs.shape[1], nsteps, time_block=S, path=path, digit_planes=T)
    plan.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
    plan.run()
    st = plan.stats()
    res[(path, T)] = plan.sfmat()
    print(path, T, st, "evals/s %.3e" % (wl.dirs.shape[1] * nsteps / st["ms"] * 1e3), flush=True)
    del plan
ref = res[("gemm", 0)]
rows = np.abs(ref).max(axis=1, keepdims=True)
for key in (("int8", 8), ("int8", 7)):
    d = np.abs(res[key] - ref)
    sc = d / np.maximum(np.abs(ref), rows)
    print(key, "max scaled diff %.3e; by step range:" % sc.max(),
          ["%.1e" % sc[:, a:b].max() for a, b in ((0, 100), (100, 1000), (1000, 5000), (5000, nsteps))], flush=True)
    # pure relative where |rho| is not tiny
    rel = d / np.abs(ref)
    print("   pure relative (all entries): max %.3e, 99.9th pct %.3e" % (rel.max(), np.quantile(rel, 0.999)))
