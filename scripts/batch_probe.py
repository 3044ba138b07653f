"""LGG09 split caller: Building BLDF01 with X0 split into 2^k boxes, batch vs one plan per split."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import reach_b200 as rb
from reach_b200 import _lib, workloads as W, discretization as D, sets as S
from reach_b200.lgg09 import Lgg09BatchPlan, Lgg09Plan
ctx = _lib.default_context(0)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 10
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
A, B, X0, U0 = W.building_sets()
U = D.normalize_input(B, U0)
splits = S.split(X0, [2] * k + [1] * (48 - k))
t0 = time.perf_counter()
base = D.forward(A, splits[0], 0.002, U)
ivps = [base] + [D.forward(A, X, 0.002, U) for X in splits[1:]]
print("host discretisation of %d splits: %.1f s" % (len(splits), time.perf_counter() - t0), flush=True)
dirs = rb.BoxDirections(48).matrix()
t0 = time.perf_counter()
batch = Lgg09BatchPlan(ctx, base.Phi, dirs, [iv.Omega0 for iv in ivps], base.V, nsteps)
print("batch create/upload: %.2f s" % (time.perf_counter() - t0), flush=True)
for _ in range(3):
    batch.run()
    st = batch.stats()
print("batch: %d splits x %d dirs x %d steps: %.3f ms, %s" % (len(splits), dirs.shape[1], nsteps, st["ms"], st), flush=True)
print("  -> %.3e evals/s, rho %.2f GB" % (len(splits) * dirs.shape[1] * nsteps / (st["ms"] * 1e-3), len(splits) * dirs.shape[1] * nsteps * 8 / 1e9))
plan = Lgg09Plan(ctx, 48, dirs.shape[1], nsteps)
plan.upload(base.Phi, dirs, ivps[-1].Omega0, base.V)
for _ in range(3):
    plan.run()
one = plan.stats()["ms"]
print("one plan per split: %.3f ms each -> %.1f ms for all (%.1fx the batch)" % (one, one * len(splits), one * len(splits) / st["ms"]))
a, b = batch.sfmat(len(splits) - 1), plan.sfmat()
scale = np.maximum(np.abs(b), np.abs(b).max(axis=1, keepdims=True))
print("max scaled diff batch vs plan (last split): %.3e" % (np.abs(a - b) / scale).max())
