"""Accuracy / throughput probe of the INT8-tcgen05 residue (CRT) GEMM against NumPy (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import reach_b200
from reach_b200 import _lib
ctx = _lib.default_context(0)
rng = np.random.default_rng(0)
cases = [(128, 128, 64), (128, 256, 64), (256, 256, 32), (100, 64, 37), (300, 200, 1000), (384, 640, 130), (2048, 2048, 2000)]
mods = (14,)
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:] if "x" in a]
    mm = [int(a) for a in sys.argv[1:] if "x" not in a]
    if mm:
        mods = tuple(mm)
for (m, n, k) in cases:
    X = rng.standard_normal((m, k)) * np.exp(rng.uniform(-3, 3, (m, 1)))
    Y = rng.standard_normal((n, k)) * np.exp(rng.uniform(-3, 3, (n, 1)))
    if m * n * k <= 4e7:
        ref = (X.astype(np.longdouble) @ Y.T.astype(np.longdouble)).astype(np.float64)
    else:
        ref = X @ Y.T
    scale = np.sqrt((X * X).sum(1))[:, None] * np.sqrt((Y * Y).sum(1))[None, :]
    for N in mods:
        Cm, tf, sl = ctx.crtgemm_probe(X, Y, moduli=N, iters=3 if m * n * k > 1e8 else 1)
        err = np.abs(Cm - ref)
        i, j = np.unravel_index(np.argmax(err / scale), err.shape)
        print("m=%d n=%d k=%d N=%d: max|err|/(|x||y|) = %.3e at (%d,%d) got %.6e want %.6e, max|err|/max|C| = %.3e, %.1f TFLOP/s-equivalent, residues %.3f ms"
              % (m, n, k, N, (err / scale).max(), i, j, Cm[i, j], ref[i, j], err.max() / np.abs(ref).max(), tf, sl), flush=True)
