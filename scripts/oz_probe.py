"""Accuracy / throughput probe of the INT8-tcgen05 digit-plane GEMM against NumPy (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import reach_b200
from reach_b200 import _lib
ctx = _lib.default_context(0)
rng = np.random.default_rng(0)
cases = [(128, 64, 32), (128, 64, 64), (256, 128, 160), (300, 200, 1000), (2048, 2048, 2000)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (m, n, k) in cases:
    X = rng.standard_normal((m, k)) * np.exp(rng.uniform(-3, 3, (m, 1)))
    Y = rng.standard_normal((n, k)) * np.exp(rng.uniform(-3, 3, (n, 1)))
    ref = X @ Y.T
    scale = np.abs(X).max(1)[:, None] * np.abs(Y).max(1)[None, :] * k
    for T in (8, 7, 6):
        Cm, tf, sl = ctx.ozgemm_probe(X, Y, slices=T, iters=3 if m * n * k > 1e8 else 1)
        err = np.abs(Cm - ref)
        print("m=%d n=%d k=%d T=%d: max|err|/(k rowmax colmax) = %.3e, max|err|/max|C| = %.3e, %.1f TFLOP/s-equivalent, slicing %.3f ms"
              % (m, n, k, T, (err / scale).max(), err.max() / np.abs(ref).max(), tf, sl), flush=True)
