"""LGG09 through the INT8 digit-plane GEMM path vs the DMMA path vs the oracle (GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import reach_b200
from reach_b200 import _lib, workloads as W
from reach_b200.lgg09 import Lgg09Plan
from oracle import lgg09_oracle as LO
ctx = _lib.default_context(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "small"

def scaled_err(got, ref):
    scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=1, keepdims=True))
    return float((np.abs(got - ref) / scale).max())

if mode == "small":
    for (n, nsteps) in ((200, 64), (333, 100)):
        wl = W.synthetic_c5(n=n, nsteps=nsteps)
        ref = LO.reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, wl.V)
        for (path, S, T) in (("gemm", 0, 0), ("int8", 1, 8), ("int8", 3, 8), ("int8", 4, 7), ("int8", 0, 6)):
            plan = Lgg09Plan(ctx, wl.n, wl.dirs.shape[1], nsteps, time_block=S, path=path, digit_planes=T)
            plan.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
            plan.run()
            got = plan.sfmat()
            print("n=%d steps=%d path=%s S=%d T=%d: scaled err vs oracle %.3e  %s" % (n, nsteps, path, S, T, scaled_err(got, ref), plan.stats()), flush=True)
else:
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    wl = W.synthetic_c5(nsteps=nsteps) if mode == "c5" else W.heat_c3(nsteps=nsteps)
    base = None
    for (path, S, T) in (("gemm", 0, 0), ("int8", 0, 8), ("int8", 0, 7), ("int8", 4, 8), ("int8", 8, 8), ("int8", 12, 8)):
        plan = Lgg09Plan(ctx, wl.n, wl.dirs.shape[1], nsteps, time_block=S, path=path, digit_planes=T)
        plan.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
        plan.run(); plan.run()
        st = plan.stats()
        got = plan.sfmat()
        if base is None:
            base = got
        print("%s n=%d steps=%d path=%s S=%d T=%d: %.2f ms (%.3e evals/s), gemm %.3f ms x %d, scaled diff vs DMMA path %.3e"
              % (mode, wl.n, nsteps, path, st["time_block"], T, st["ms"], wl.dirs.shape[1] * nsteps / st["ms"] * 1e3,
                 st["gemm_ms"], st["gemm_launches"], scaled_err(got, base)), flush=True)
        del plan
