"""One large digit-plane GEMM for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import reach_b200
from reach_b200 import _lib
ctx = _lib.default_context(0)
rng = np.random.default_rng(0)
m, n, k = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "2048x8192x2000").split("x"))
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
X = rng.standard_normal((m, k)); Y = rng.standard_normal((n, k))
_, tf, sl = ctx.ozgemm_probe(X, Y, slices=T, iters=2, want_result=False)
print(m, n, k, T, tf, sl)
