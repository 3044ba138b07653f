"""Extract the benchmark model matrices from the reference's JLD2 example data into small
.npz fixtures (run once in the build container; /root/reference does not exist on the GPU box).

  examples/Building/building.jld2 -> tests/golden/models/building.npz  (A 48x48, B 48, C 48)
  examples/ISS/ISS.jld2           -> tests/golden/models/iss.npz       (A CSC 270x270, B 270x3, C 3x270)
  examples/Heat3D/HEAT02.jld2     -> tests/golden/models/heat02.npz    (A CSC 1000x1000, xind)

Usage: python tools/make_model_fixtures.py [/root/reference]
"""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from jld2_min import JLD2File  # noqa: E402

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "models")
os.makedirs(out, exist_ok=True)


def dense(x):
    return x.toarray() if hasattr(x, "toarray") else np.asarray(x)


def csc_parts(M):
    M = M.tocsc()
    return dict(data=M.data, indices=M.indices.astype(np.int32), indptr=M.indptr.astype(np.int32),
                shape=np.array(M.shape, dtype=np.int64))


f = JLD2File(os.path.join(ref, "examples/Building/building.jld2"))
np.savez_compressed(os.path.join(out, "building.npz"), A=f.read("A"), B=f.read("B"), C=f.read("C"))

f = JLD2File(os.path.join(ref, "examples/ISS/ISS.jld2"))
A = csc_parts(f.read("A"))
np.savez_compressed(os.path.join(out, "iss.npz"), B=dense(f.read("B")), C=dense(f.read("C")),
                    **{"A_" + k: v for k, v in A.items()})

f = JLD2File(os.path.join(ref, "examples/Heat3D/HEAT02.jld2"))
A = csc_parts(f.read("A"))
np.savez_compressed(os.path.join(out, "heat02.npz"), xind=f.read("xind"),
                    **{"A_" + k: v for k, v in A.items()})
for n in sorted(os.listdir(out)):
    print(n, os.path.getsize(os.path.join(out, n)))
