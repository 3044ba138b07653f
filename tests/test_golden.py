"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle must keep
reproducing them (CPU), and the CUDA path must match them (GPU)."""
import os
import sys

import numpy as np
import pytest

import reach_b200 as rb
from reach_b200 import discretization as D
from conftest import assert_support_parity

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as MG  # noqa: E402


def test_oracle_reproduces_lgg09_golden():
    want = np.load(os.path.join(HERE, "golden", "lgg09_golden.npz"))
    got = MG.lgg09_cases()
    for k in want.files:
        np.testing.assert_allclose(got[k], want[k], rtol=1e-12, atol=1e-15, err_msg=k)


def test_oracle_reproduces_glgm06_golden():
    want = np.load(os.path.join(HERE, "golden", "glgm06_golden.npz"))
    got = MG.glgm06_cases()
    assert float(want["min_gap"]) > 1e-9          # selection is well defined on this fixture
    for k in want.files:
        if k.startswith("sel_"):
            np.testing.assert_array_equal(got[k], want[k], err_msg=k)
        else:
            np.testing.assert_allclose(got[k], want[k], rtol=1e-12, atol=1e-15, err_msg=k)


@pytest.mark.gpu
def test_gpu_matches_lgg09_golden(ctx):
    from reach_b200.lgg09 import reach_inhomog_LGG09
    want = np.load(os.path.join(HERE, "golden", "lgg09_golden.npz"))
    A, B, X0, U0 = MG.building()
    U = D.normalize_input(B, U0)
    ivp = D.forward(A, X0, 0.004, U)
    dirs = rb.LGG09(δ=0.004, vars=(25,), n=48).template.matrix()
    rho = reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, 5000, None, ivp.V, ctx=ctx).sfmat()
    assert_support_parity(rho[:, ::50], want["bldf01_e25_every50"])
    np.testing.assert_allclose(rho.max(axis=1), want["bldf01_e25_max"], rtol=1e-10)
    ivp = D.forward(A, X0, 0.002, U)
    for S in (1, 32):
        rho = reach_inhomog_LGG09(rb.BoxDirections(48).matrix(), ivp.Omega0, ivp.Phi, 300, None, ivp.V, ctx=ctx,
                                  time_block=S).sfmat()
        assert_support_parity(rho, want["c1_head300"])
    n = 9
    X0 = rb.Zonotope(want["rand9_c"], want["rand9_G"])
    Uz = rb.LinearMap(want["rand9_M"], rb.Hyperrectangle(np.array([0.1, -0.1]), np.array([0.05, 0.02])))
    ivp = D.no_bloating(want["rand9_inputs_A"], X0, 0.05, Uz)
    dirs = np.vstack([rb.OctDirections(4).matrix(), np.zeros((n - 4, 32))])
    rho = reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, 80, None, ivp.V, ctx=ctx).sfmat()
    assert_support_parity(rho, want["rand9_oct"])


@pytest.mark.gpu
def test_gpu_matches_glgm06_golden(ctx):
    from reach_b200.glgm06 import _run_batch, GLGM06
    want = np.load(os.path.join(HERE, "golden", "glgm06_golden.npz"))
    X0 = rb.Hyperrectangle(want["x0c"], want["x0r"])
    U = rb.LinearMap(want["M"], rb.Hyperrectangle(np.array([0.2, 0.0]), np.array([0.05, 0.1])))
    ivp = D.forward_zono(want["A"], X0, 0.02, U)
    plan = _run_batch(GLGM06(δ=0.02, max_order=3), ivp.Phi, [ivp.Omega0], ivp.V, 60, ctx)
    for k in (0, 1, 2, 3, 10, 30, 59):
        c, G, sel = plan.fetch(0, k, with_selection=True)
        scale = max(np.abs(want["G_%d" % k]).max(), 1e-300)
        assert G.shape == want["G_%d" % k].shape
        assert np.abs(G - want["G_%d" % k]).max() <= 1e-10 * scale
        assert np.abs(c - want["c_%d" % k]).max() <= 1e-10 * max(np.abs(want["c_%d" % k]).max(), scale)
        if "sel_%d" % k in want.files:
            np.testing.assert_array_equal(sel[sel >= 0], want["sel_%d" % k])     # identical selection
