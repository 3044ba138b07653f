"""BOX algorithm on the device (reach_box, C ABI) against the CPU oracle (Algorithms/BOX/reach_homog.jl,
reach_inhomog.jl restated in oracle/box_oracle.py)."""
import numpy as np
import pytest
import scipy.linalg

import reach_b200 as rb
from reach_b200 import box as B, discretization as D
from reach_b200.solve import solve, IVP
from oracle import box_oracle as BO, lazysets_oracle as LS
from conftest import building_problem

pytestmark = pytest.mark.gpu


def _sys(n, seed, delta=0.05):
    rng = np.random.default_rng(seed)
    A = -np.eye(n) + (0.5 / np.sqrt(n)) * rng.standard_normal((n, n))
    return A, scipy.linalg.expm(A * delta), rng


def _close(got, ref, tol=1e-10):
    scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=1, keepdims=True))
    scale[scale == 0] = 1.0
    err = (np.abs(got - ref) / scale).max()
    assert err <= tol, err


@pytest.mark.parametrize("n,nsteps", [(1, 5), (7, 40), (48, 120), (130, 33), (300, 20)])
@pytest.mark.parametrize("S", [0, 1, 5])
def test_box_homogeneous_and_inhomogeneous(ctx, n, nsteps, S):
    A, Phi, rng = _sys(n, n)
    c0, r0 = rng.uniform(-1, 1, n), 0.1 * rng.uniform(0, 1, n)
    cU, rU = rng.uniform(-0.1, 0.1, n), 0.02 * rng.uniform(0, 1, n)
    Cr, Rr = BO.reach_homog_BOX(c0, r0, Phi, nsteps)
    Cg, Rg = B._reach(c0, r0, Phi, nsteps, ctx=ctx, time_block=S)
    _close(Cg, Cr)
    _close(Rg, Rr)
    assert np.array_equal(Cg[:, 0], c0) and np.array_equal(Rg[:, 0], r0)       # F[1] = Omega0 itself
    assert (Rg >= 0).all()
    Cr, Rr = BO.reach_inhomog_BOX(c0, r0, Phi, nsteps, cU, rU)
    Cg, Rg = B._reach(c0, r0, Phi, nsteps, cU, rU, ctx=ctx, time_block=S)
    _close(Cg, Cr)
    _close(Rg, Rr)


def test_box_centre_is_not_polluted_by_the_radius(ctx):
    """The centre comes from the linear part alone: with c0 = 0 it is exactly zero however large the radius, and a
    centre 1e-12 of the radius keeps its relative accuracy (it would not if it were recovered as (rho+ - rho-)/2)."""
    n, nsteps = 40, 30
    _, Phi, rng = _sys(n, 9)
    r0 = rng.uniform(1, 2, n)
    Cg, Rg = B._reach(np.zeros(n), r0, Phi, nsteps, ctx=ctx)
    assert np.all(Cg == 0.0)
    c0 = 1e-12 * rng.uniform(1, 2, n)
    Cr, _ = BO.reach_homog_BOX(c0, r0, Phi, nsteps)
    Cg, _ = B._reach(c0, r0, Phi, nsteps, ctx=ctx)
    _close(Cg, Cr)


def test_box_recursive(ctx):
    n, nsteps = 25, 30
    _, Phi, rng = _sys(n, 4)
    c0, r0 = rng.uniform(-1, 1, n), 0.1 * rng.uniform(0, 1, n)
    Cr, Rr = BO.reach_homog_BOX(c0, r0, Phi, nsteps, recursive=True)
    Cg, Rg = B._reach(c0, r0, Phi, nsteps, recursive=True, ctx=ctx)
    _close(Cg, Cr)
    _close(Rg, Rr)
    # wrapping effect: the recursive boxes contain the unwrapped ones
    _, Rn = B._reach(c0, r0, Phi, nsteps, ctx=ctx)
    assert (Rg >= Rn - 1e-12).all()
    with pytest.raises(NotImplementedError):
        B._reach(c0, r0, Phi, nsteps, c0, r0, recursive=True, ctx=ctx)


def test_box_post_building(ctx):
    """`solve(prob; alg=BOX(δ))` on BLDF01: the lazy Omega0 and V are first overapproximated by their box hulls
    (post.jl:29,40), then propagated; checked against the oracle loops fed with oracle-side hulls, and against the
    box-template LGG09 flowpipe, which it must contain (equal at the first set)."""
    prob = building_problem()
    sol = solve(prob, T=0.2, alg=rb.BOX(δ=0.002), ctx=ctx)
    assert len(sol.F) == 100
    C, R = sol.F.ext["centers"], sol.F.ext["radii"]
    divp = rb.discretize(prob, 0.002, rb.Forward())
    I = np.eye(48)
    hull = lambda X: (np.array([LS.rho(I[i], X) for i in range(48)]), np.array([LS.rho(-I[i], X) for i in range(48)]))
    hi0, nlo0 = hull(divp.Omega0)
    hiU, nloU = hull(divp.V)
    Cr, Rr = BO.reach_inhomog_BOX((hi0 - nlo0) / 2, (hi0 + nlo0) / 2, divp.Phi, 100, (hiU - nloU) / 2, (hiU + nloU) / 2)
    _close(C, Cr)
    _close(R, Rr)
    ref = solve(prob, T=0.2, alg=rb.LGG09(δ=0.002, template=rb.BoxDirections(48)), ctx=ctx)
    sf = ref.F.device.sfmat()
    hi, lo = sf[:48], -sf[48:]
    scale = np.abs(sf).max()
    assert np.abs((C + R)[:, 0] - hi[:, 0]).max() <= 1e-12 * scale
    assert ((C + R) >= hi - 1e-12 * scale).all() and ((C - R) <= lo + 1e-12 * scale).all()
    e25 = np.zeros(48); e25[24] = 1.0
    assert rb.rho(e25, sol.F[10]) == pytest.approx((C + R)[24, 10], rel=1e-15)
    assert sol.F[3].dt.lo == pytest.approx(3 * 0.002) and sol.F[3].dt.hi == pytest.approx(4 * 0.002)


def test_box_invariant_stops_the_flowpipe(ctx):
    n, nsteps = 6, 80
    A = -0.5 * np.eye(n)
    Phi = scipy.linalg.expm(A * 0.1)
    c0, r0 = np.full(n, 2.0), np.full(n, 0.1)
    X = rb.HalfSpace(-np.eye(n)[0], -1.0)                       # x1 >= 1
    Cr, _ = BO.reach_homog_BOX(c0, r0, Phi, nsteps, X=("halfspace", X.a, X.b))
    F = B.reach_homog_BOX(rb.Hyperrectangle(c0, r0), Phi, nsteps, 0.1, X, ctx=ctx)
    assert len(F) == Cr.shape[1] < nsteps
    assert F[len(F) - 1].set.center[0] + F[len(F) - 1].set.radius[0] >= 1.0


def test_box_rejects_bad_arguments(ctx):
    Phi = np.eye(3)
    with pytest.raises(ValueError):
        B._reach(np.zeros(3), -np.ones(3), Phi, 5, ctx=ctx)       # negative radius
    with pytest.raises(AssertionError):
        B._reach(np.zeros(2), np.ones(2), Phi, 5, ctx=ctx)
    with pytest.raises(ValueError):
        B.post(rb.BOX(δ=0.1), D.forward(-np.eye(3), rb.Hyperrectangle(np.zeros(3), np.ones(3)), 0.1), ctx=ctx)   # NSTEPS missing
