"""Device-side Φ, Φ₁, Φ₂ (reach_discretize_phi) against the host restatement of
src/Discretization/Exponentiation.jl (SciPy `expm` of the block matrices)."""
import numpy as np
import pytest
import scipy.linalg

import reach_b200 as rb
from reach_b200 import discretization as D
from reach_b200.lgg09 import reach_inhomog_LGG09
from oracle import lgg09_oracle as LO
from conftest import building_problem, iss_problem, assert_support_parity

pytestmark = pytest.mark.gpu


def host_phis(A, delta):
    prev = D.set_device_context(None)
    try:
        return scipy.linalg.expm(A * delta), D.Phi1(A, delta), D.Phi2(A, delta)
    finally:
        D.set_device_context(prev)


def check(got, ref, tol=2e-13):
    for g, r, name in zip(got, ref, ("Phi", "Phi1", "Phi2")):
        err = np.abs(g - r).max() / np.abs(r).max()
        assert err <= tol, "%s: max error %.3e of max|.|" % (name, err)


@pytest.mark.parametrize("n,delta,scale", [(1, 0.1, 1.0), (5, 0.05, 1.0), (48, 0.01, 30.0), (130, 0.2, 3.0), (600, 0.01, 1.0)])
def test_random_matrices(ctx, n, delta, scale):
    rng = np.random.default_rng(n)
    A = scale * (-np.eye(n) + (0.5 / np.sqrt(n)) * rng.standard_normal((n, n)))
    check(D.device_phi(ctx, A, delta), host_phis(A, delta))


def test_exponentiation_jl_6_26_diagonal(ctx):
    """test/discretization/exponentiation.jl:6-14: _exp([10 0; 0 20], 0.1) = diag(e, e^2); the reference compares
    its dense `exp` with `==`, the device route (Taylor + 2 squarings) is held to a few ulp and exact zeros."""
    Phi, P1, P2 = D.device_phi(ctx, np.array([[10.0, 0.0], [0.0, 20.0]]), 0.1)
    want = np.array([[np.exp(1.0), 0.0], [0.0, np.exp(2.0)]])
    assert Phi[0, 1] == 0.0 and Phi[1, 0] == 0.0
    np.testing.assert_allclose(np.diag(Phi), np.diag(want), rtol=4e-16 * 8)
    np.testing.assert_allclose(np.diag(P1), [(np.exp(1.0) - 1) / 10.0, (np.exp(2.0) - 1) / 20.0], rtol=1e-14)
    np.testing.assert_allclose(np.diag(P2), [(np.exp(1.0) - 2) / 100.0, (np.exp(2.0) - 3) / 400.0], rtol=1e-14)


def test_building_needs_scaling_and_matches(ctx):
    """‖A‖∞ δ = 23.7 at δ = 0.002: six squarings; also Φ₂(|A|, δ), the matrix the Forward model bloats with."""
    prob = building_problem()
    A = np.asarray(prob.A)
    check(D.device_phi(ctx, A, 0.002), host_phis(A, 0.002))
    check(D.device_phi(ctx, np.abs(A), 0.002), host_phis(np.abs(A), 0.002))


def test_iss_keeps_the_block_structure(ctx):
    """ISS: expm(Aδ) has exactly 540 non-zeros (135 decoupled 2x2 blocks); products of exact zeros stay zero."""
    prob, _ = iss_problem()
    A = np.asarray(prob.A)
    Phi, P1, P2 = D.device_phi(ctx, A, 6e-4)
    ref = host_phis(A, 6e-4)
    check((Phi, P1, P2), ref)
    assert np.count_nonzero(Phi) == np.count_nonzero(ref[0]) == 540


def test_forward_model_on_device_feeds_the_hot_path(ctx):
    """discretize(Forward) with the matrix functions computed on the device, then LGG09: same flowpipe as with the
    host discretisation (SciPy) to the support-value tolerance."""
    prob = building_problem()
    U = D.normalize_input(prob.B, prob.U)
    host = D.forward(prob.A, prob.X0, 0.002, U)
    prev = D.set_device_context(ctx)
    try:
        dev = D.forward(prob.A, prob.X0, 0.002, U)
    finally:
        D.set_device_context(prev)
    assert np.abs(dev.Phi - host.Phi).max() <= 1e-13
    dirs = rb.BoxDirections(48).matrix()
    ref = LO.reach_inhomog_LGG09(dirs, host.Omega0, host.Phi, 200, host.V)
    got = reach_inhomog_LGG09(dirs, dev.Omega0, dev.Phi, 200, None, dev.V, ctx=ctx).sfmat()
    assert_support_parity(got, ref, tol=1e-9)     # two different exponentials upstream of the path, not the path itself


def test_solve_with_device_discretisation(ctx):
    prob = building_problem()
    alg = rb.LGG09(δ=0.004, vars=(25,), n=48)
    a = rb.solve(prob, NSTEPS=300, alg=alg, ctx=ctx)
    b = rb.solve(prob, NSTEPS=300, alg=alg, ctx=ctx, device_discretize=True)
    assert D._device_ctx is None                      # restored
    assert_support_parity(b.F.support_function_matrix(), a.F.support_function_matrix(), tol=1e-9)


def test_argument_errors(ctx):
    with pytest.raises(ValueError):
        D.device_phi(ctx, np.eye(3), 0.0)
    with pytest.raises(ValueError):
        D.device_phi(ctx, np.full((2, 2), np.nan), 0.1)
    with pytest.raises(ValueError):
        D.device_phi(ctx, np.ones((2, 3)), 0.1)
