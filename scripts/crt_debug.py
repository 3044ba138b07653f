"""Debug of the residue GEMM: raw INT32 accumulators of one modulus of the first 256 x 256 tile against NumPy."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
MODS = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193]
imod = int(sys.argv[1]) if len(sys.argv) > 1 else 0
N = 14
os.environ["REACH_CRT_DEBUG"] = str(imod)
import reach_b200
from reach_b200 import _lib
ctx = _lib.default_context(0)
rng = np.random.default_rng(1)
m, n, k = 256, 256, 64
X = rng.standard_normal((m, k)); Y = rng.standard_normal((n, k))
P = 1
for p in MODS[:N]: P *= p
tot = P.bit_length() - 1 - 2
alpha = (tot + 1) // 2; beta = tot - alpha
def scaled(A, bits):
    out = np.zeros(A.shape, dtype=object)
    for r in range(A.shape[0]):
        mx = np.abs(A[r]).max()
        em = int(np.floor(np.log2(mx)))
        ss = ((A[r] * 2.0 ** -em) ** 2).sum()
        en = int(np.floor(np.log2(np.sqrt(ss) * 1.000001))) + 1
        s = bits - (em + en)
        out[r] = [int(np.rint(np.ldexp(v, s))) for v in A[r]]
    return out
def res(Ai, p):
    R = np.zeros(Ai.shape, dtype=np.int64)
    for idx, v in np.ndenumerate(Ai):
        r = v % p
        if r > (p - 1) // 2: r -= p
        R[idx] = r
    return R
p = MODS[imod]
Xi, Yi = scaled(X, alpha), scaled(Y, beta)
want = res(Xi, p) @ res(Yi, p).T
Cm, tf, sl = ctx.crtgemm_probe(X, Y, moduli=N)
got = Cm.view(np.int32).reshape(-1)[:2 * 128 * 256].reshape(256, 256)
print("modulus", p, "alpha", alpha, "beta", beta)
print("want[:3,:6]\n", want[:3, :6]); print("got[:3,:6]\n", got[:3, :6])
print("match all:", np.array_equal(want, got))
for (r0, c0) in ((0, 0), (0, 128), (128, 0), (128, 128)):
    print("block", r0, c0, "match", np.array_equal(want[r0:r0 + 128, c0:c0 + 128], got[r0:r0 + 128, c0:c0 + 128]))
# does got match some permutation of blocks?
for (r0, c0) in ((0, 0), (0, 128), (128, 0), (128, 128)):
    for (r1, c1) in ((0, 0), (0, 128), (128, 0), (128, 128)):
        if np.array_equal(want[r0:r0 + 128, c0:c0 + 128], got[r1:r1 + 128, c1:c1 + 128]): print("want block", r0, c0, "== got block", r1, c1)
print("row sums equal?", np.array_equal(np.sort(want, axis=None), np.sort(got, axis=None)))
