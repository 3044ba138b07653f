"""Round 2: C5 full length on the DMMA path and the guarded INT8 path: time per pass and accumulated scaled difference
to the DMMA path.  REACH_OZ_NOPAIR=1 in the environment switches the two-level (N = 256) MMAs off (A/B)."""
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
for (name, path, tout) in (("dmma", "gemm", None), ("int8", "auto", None)):
    if tout is None:
        os.environ.pop("REACH_OZ_TOUT", None)
    else:
        os.environ["REACH_OZ_TOUT"] = tout
    plan = Lgg09Plan(ctx, wl.n, wl.dirs.shape[1], nsteps, path=path)
    plan.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
    plan.run(); plan.run()
    st = plan.stats()
    res[name] = plan.sfmat()
    print(name, st, plan.path_info() if hasattr(plan, "path_info") else "", "evals/s %.4e" % (wl.dirs.shape[1] * nsteps / st["ms"] * 1e3), flush=True)
    del plan
ref = res["dmma"]
rows = np.abs(ref).max(axis=1, keepdims=True)
for key in res:
    if key == "dmma":
        continue
    d = np.abs(res[key] - ref)
    sc = d / np.maximum(np.abs(ref), rows)
    print(key, "max scaled diff %.3e; by step range:" % sc.max(),
          ["%.1e" % sc[:, a:b].max() for a, b in ((0, 100), (100, 1000), (1000, 5000), (5000, nsteps))], flush=True)
