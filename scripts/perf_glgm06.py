"""Exploratory GLGM06 timing on the GPU box: config C4 (Building, 1024 splits, max_order 10)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import reach_b200 as rb
from reach_b200 import _lib, workloads as W
from reach_b200.glgm06 import _run_batch, GLGM06
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
per_dim = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = _lib.default_context(0)
t = time.time()
Phi, Oms, V = W.building_c4(per_dim, nsteps)
if len(sys.argv) > 3:
    Oms = Oms[:int(sys.argv[3])]          # e.g. 128: one rank's share at 8 GPUs
print("host discretization of %d splits: %.2f s" % (len(Oms), time.time() - t))
alg = GLGM06(δ=0.002, max_order=10)
t = time.time()
plan = _run_batch(alg, Phi, Oms, V, nsteps, ctx)
print("first run incl. upload: %.3f s" % (time.time() - t), plan.stats())
for _ in range(2):
    plan.run()
    print("run:", plan.stats())
d = np.zeros(48); d[24] = 1.0
t = time.time(); sup = plan.support(d); print("support over %s zonotopes: %.3f s, max %.6e" % (sup.shape, time.time() - t, sup.max()))
t = time.time(); c, G = plan.fetch(len(Oms) - 1, nsteps - 1); print("fetch one zonotope: %.4f s" % (time.time() - t), G.shape)
