"""Auto path selection with the DMMA guard: C5 (should stay on int8) and HEAT02 (should fall back)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import reach_b200
from reach_b200 import _lib, workloads as W
from reach_b200.lgg09 import Lgg09Plan
ctx = _lib.default_context(0)
def scaled_err(got, ref):
    scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=1, keepdims=True))
    return float((np.abs(got - ref) / scale).max())
for name, wl in (("c5", W.synthetic_c5(nsteps=600)), ("c3", W.heat_c3(nsteps=600))):
    base = None
    for path in ("gemm", "auto", "int8"):
        plan = Lgg09Plan(ctx, wl.n, wl.dirs.shape[1], wl.nsteps, path=path)
        plan.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
        plan.run(); plan.run()
        got = plan.sfmat()
        if base is None: base = got
        print(name, path, "scaled diff vs DMMA %.3e" % scaled_err(got, base), plan.stats(), flush=True)
        del plan
