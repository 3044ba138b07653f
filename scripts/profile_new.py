"""Short runs of the compact-chain path (C2 ISS) and the LGG09 split batch (Building) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import reach_b200 as rb
from reach_b200 import _lib, workloads as W, discretization as D, sets as S
from reach_b200.lgg09 import Lgg09Plan, Lgg09BatchPlan
which = sys.argv[1] if len(sys.argv) > 1 else "compact"
ctx = _lib.default_context(0)
if which == "compact":
    wl = W.iss_c2()
    plan = Lgg09Plan(ctx, wl.n, wl.dirs.shape[1], wl.nsteps)
    plan.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
    plan.run()
    plan.run()
    print(plan.stats())
else:
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
    A, B, X0, U0 = W.building_sets()
    U = D.normalize_input(B, U0)
    ivps = [D.forward(A, X, 0.002, U) for X in S.split(X0, [2] * k + [1] * (48 - k))]
    batch = Lgg09BatchPlan(ctx, ivps[0].Phi, rb.BoxDirections(48).matrix(), [iv.Omega0 for iv in ivps], ivps[0].V, nsteps)
    batch.run()
    batch.run()
    print(batch.stats())
