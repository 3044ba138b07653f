"""One residue-GEMM launch shape for ncu / timing experiments: python scripts/crt_profile.py m n k [moduli] [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import reach_b200
from reach_b200 import _lib
ctx = _lib.default_context(0)
m, n, k = (int(v) for v in sys.argv[1:4])
N = int(sys.argv[4]) if len(sys.argv) > 4 else 14
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
rng = np.random.default_rng(0)
X = rng.standard_normal((m, k)); Y = rng.standard_normal((n, k))
_, tf, sl = ctx.crtgemm_probe(X, Y, moduli=N, iters=iters, want_result=False)
print("m=%d n=%d k=%d N=%d flags=%s: %.1f TFLOP/s-equivalent = %.0f TOP/s INT8, %.3f ms per launch, residues %.3f ms"
      % (m, n, k, N, os.environ.get("REACH_CRT_FLAGS", "0"), tf, tf * N, 2.0 * m * n * k / tf / 1e9, sl), flush=True)
