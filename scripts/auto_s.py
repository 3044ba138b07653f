"""Time block / tile the auto-tuner picks for the BASELINE configs, and the C1 run time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import reach_b200
from reach_b200 import _lib, workloads as W
from reach_b200.lgg09 import Lgg09Plan
ctx = _lib.default_context(0)
for name, wl, run in (("C1", W.building_c1(), True), ("C2/gemm", W.iss_c2(), False), ("C3", W.heat_c3(nsteps=10000), False),
                      ("C5", W.synthetic_c5(), False), ("C5/8", W.synthetic_c5(), False)):
    dirs = wl.dirs
    if name == "C5/8":
        dirs = dirs[:, list(range(250)) + list(range(2000, 2250))]
    plan = Lgg09Plan(ctx, wl.n, dirs.shape[1], wl.nsteps, path="gemm" if name == "C2/gemm" else "auto")
    plan.upload(wl.Phi, dirs, wl.Omega0, wl.V)
    if run:
        for _ in range(4):
            plan.run()
    print(name, plan.stats(), flush=True)
    plan.close()
