import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import reach_b200  # noqa: E402,F401  (registers the package under this name)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """Device context; the product path has no CPU fallback, so this fails loudly without a GPU."""
    from reach_b200 import _lib
    return _lib.default_context(0)


def model(name):
    return np.load(os.path.join(ROOT, "tests", "golden", "models", name + ".npz"))


def csc(m, prefix="A_"):
    import scipy.sparse as sp
    return sp.csc_matrix((m[prefix + "data"], m[prefix + "indices"], m[prefix + "indptr"]),
                         shape=tuple(m[prefix + "shape"]))


def building_problem():
    """examples/Building/Building.jl:58-73 (BLDF01)."""
    import reach_b200 as rb
    m = model("building")
    c = np.r_[np.full(10, 0.000225), np.zeros(38)]
    r = np.r_[np.full(10, 0.000025), np.zeros(14), 0.0001, np.zeros(23)]
    return rb.IVP(m["A"], rb.Hyperrectangle(c, r), m["B"], rb.Interval(0.8, 1.0))


def iss_problem():
    """examples/ISS/ISS.jl:81-91 (ISSF01)."""
    import reach_b200 as rb
    m = model("iss")
    A = csc(m).toarray()
    X0 = rb.BallInf(np.zeros(270), 0.0001)
    U = rb.Hyperrectangle(low=[0.0, 0.8, 0.9], high=[0.1, 1.0, 1.0])
    return rb.IVP(A, X0, m["B"], U), m["C"]


def assert_support_parity(got, ref, tol=1e-10):
    """north_star tolerance: 1e-10 relative, where 'relative' is to max(|ρ_ref|, row scale) --
    pure relative error is ill-posed at the zero crossings of ρ (SURVEY §7 hard part 4)."""
    assert got.shape == ref.shape
    rowscale = np.max(np.abs(ref), axis=1, keepdims=True)
    denom = np.maximum(np.abs(ref), rowscale)
    denom[denom == 0] = 1.0
    err = np.abs(got - ref) / denom
    assert np.all(np.isfinite(got)), "non-finite support values"
    assert err.max() <= tol, "max scaled error %.3e at %s" % (err.max(), np.unravel_index(err.argmax(), err.shape))
    return err.max()
