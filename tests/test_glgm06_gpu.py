"""Parity of the CUDA GLGM06 path (through the C ABI) against the CPU oracle: generator values
within 1e-10 (relative to the zonotope's scale) and IDENTICAL generator selection."""
import numpy as np
import pytest

import reach_b200 as rb
from reach_b200 import discretization as D
from reach_b200.glgm06 import _run_batch, GLGM06
from oracle import glgm06_oracle as GO, lazysets_oracle as Z
from conftest import building_problem

pytestmark = pytest.mark.gpu
TOL = 1e-10


def check_zonotope(got, ref, what=""):
    (c, G), (cr, Gr) = got, ref
    assert G.shape == Gr.shape, "%s: %s generators, reference has %s" % (what, G.shape, Gr.shape)
    scale = max(np.abs(Gr).max(initial=0.0), np.abs(cr).max(initial=0.0), 1e-300)
    assert np.abs(c - cr).max(initial=0.0) <= TOL * scale, what
    assert np.abs(G - Gr).max(initial=0.0) <= TOL * scale, what


def stable(n, seed, delta):
    rng = np.random.default_rng(seed)
    A = -np.eye(n) + (0.5 / np.sqrt(n)) * rng.standard_normal((n, n))
    return A


def test_flowpipes_jl_43_44_on_the_device(ctx):
    """test/flowpipes/flowpipes.jl:40-44 through solve() exactly as the reference calls it: `solve(prob; tspan=dt)`
    with no algorithm => the default GLGM06 with δ = diam(tspan) / 100 = 0.2 (solve.jl:326-352)."""
    A = np.array([[0.0, 1.0], [-1.0, 0.0]])
    sol = rb.solve(rb.IVP(A, rb.Singleton([1.0, 0.0])), tspan=(0.0, 20.0), ctx=ctx)
    assert isinstance(sol.alg, rb.GLGM06) and sol.alg.δ == 0.2
    assert len(sol) == 100
    d = np.ones(2)
    assert rb.rho(d, sol) < 2.0
    assert abs(rb.rho(d, sol) - 1.4491791666187368) < 1e-13
    rhos = [R.rho(d) for R in sol]
    k = int(np.argmax(rhos))
    Zk = sol[k].set
    sigma = Zk.center + Zk.generators @ np.sign(Zk.generators.T @ d)
    np.testing.assert_allclose(sigma, [0.7502062459270877, 0.6989729206916551], rtol=1e-12)
    np.testing.assert_allclose(rb.sigma(d, sol), [0.7502062459270877, 0.6989729206916551], rtol=1e-12)   # σ(d, fp) itself


def test_flowpipes_jl_56_70_projectile_inclusion_on_the_device(ctx):
    """test/flowpipes/flowpipes.jl:56-70: GLGM06(δ=1e-2, Forward()), T = 20 ⇒ sol ⊆ {24 x1 + x3 <= 375};
    support values of all 2000 zonotopes against the oracle's."""
    from test_oracle_kat import projectile_problem
    A, X0, U = projectile_problem()
    alg = rb.GLGM06(δ=1e-2, approx_model=rb.Forward(setops="zono"))
    sol = rb.solve(rb.IVP(A, X0, np.eye(4), U), T=20.0, alg=alg, ctx=ctx)
    assert len(sol) == 2000
    d = np.array([24.0, 0.0, 1.0, 0.0])
    assert rb.rho(d, sol) <= 375.0
    ivp = D.forward_zono(A, X0, 1e-2, U)
    ref = GO.post_GLGM06(ivp.Phi, ivp.Omega0, 2000, 5, ivp.V)
    want = max(Z.rho_zonotope(d, c, G) for c, G in ref)
    assert abs(rb.rho(d, sol) - want) <= 1e-10 * abs(want)
    for k in (0, 1, 500, 1999):
        check_zonotope((sol[k].set.center, sol[k].set.generators), ref[k], "step %d" % k)


@pytest.mark.parametrize("n,p_extra,nsteps", [(2, 0, 30), (5, 3, 60), (48, 0, 25), (70, 10, 12)])
def test_homogeneous_matches_oracle(ctx, n, p_extra, nsteps):
    rng = np.random.default_rng(n)
    A = stable(n, n, 0.05)
    X0 = rb.Zonotope(rng.standard_normal(n), 0.1 * rng.standard_normal((n, n + p_extra)))
    ivp = D.first_order_zonotope(A, X0, 0.05)
    plan = _run_batch(GLGM06(δ=0.05, max_order=50), ivp.Phi, [ivp.Omega0], None, nsteps, ctx)
    ref = GO.post_GLGM06(ivp.Phi, ivp.Omega0, nsteps, 50)
    for k in range(nsteps):
        check_zonotope(plan.fetch(0, k), ref[k], "step %d" % k)


def test_homogeneous_singleton_has_one_zero_generator(ctx):
    """reach_homog.jl:52-54: p == 0 ⇒ one zero generator."""
    A = stable(3, 1, 0.1)
    import scipy.linalg
    Phi = scipy.linalg.expm(A * 0.1)
    plan = _run_batch(GLGM06(δ=0.1), Phi, [rb.Singleton([1.0, 2.0, 3.0])], None, 5, ctx)
    ref = GO.post_GLGM06(Phi, rb.Singleton([1.0, 2.0, 3.0]), 5, 5)
    for k in range(5):
        check_zonotope(plan.fetch(0, k), ref[k])
        assert plan.fetch(0, k)[1].shape == (3, 1)


@pytest.mark.parametrize("n,r,nsteps", [(5, 5, 200), (8, 1, 40), (8, 3, 120), (13, 2, 60)])
def test_inhomogeneous_values_and_selection(ctx, n, r, nsteps):
    """test/algorithms/GLGM06.jl:81-107 shapes (linear5D, motor-like 8-D incl. max_order=1) with
    full numeric parity and identical GIR05 selection."""
    rng = np.random.default_rng(100 + n)
    A = stable(n, n, 0.02)
    X0 = rb.Hyperrectangle(rng.standard_normal(n), np.where(rng.uniform(size=n) < 0.7, 0.1, 0.0) + 0.0)
    q = max(1, n // 3)
    U = rb.LinearMap(rng.standard_normal((n, q)), rb.Hyperrectangle(0.1 * rng.standard_normal(q), 0.05 * np.ones(q)))
    ivp = D.forward_zono(A, X0, 0.02, U)
    plan = _run_batch(GLGM06(δ=0.02, max_order=r), ivp.Phi, [ivp.Omega0], ivp.V, nsteps, ctx)
    ref, sels, gaps = GO.post_GLGM06(ivp.Phi, ivp.Omega0, nsteps, r, ivp.V, return_diag=True)
    reduced = 0
    for k in range(nsteps):
        c, G, sel = plan.fetch(0, k, with_selection=True)
        check_zonotope((c, G), ref[k], "step %d" % k)
        if k > 0 and sels[k] is not None:
            reduced += 1
            gap = gaps[k - 1]
            if r > 1 and gap is not None and gap > 1e-9:      # selection well defined
                keep = sel[sel >= 0]
                np.testing.assert_array_equal(keep, sels[k], err_msg="kept generators differ at step %d" % k)
                assert np.all(sel[len(keep):] == -1)
    assert reduced > 0, "the test never exercised the order reduction"


def test_building_split_batch_c4_small(ctx):
    """Config C4 in miniature: Building, Forward(:zono), max_order=10, 8 splits, 60 steps; every
    split agrees with an independent oracle run, and the batch shares Φ and W_k."""
    prob = building_problem()
    splits = rb.split(prob.X0, [2, 2, 2] + [1] * 45)
    alg = rb.GLGM06(δ=0.002, max_order=10, approx_model=rb.Forward(setops="zono"))
    nsteps = 60
    sol = rb.solve(rb.IVP(prob.A, splits, prob.B, prob.U), NSTEPS=nsteps, alg=alg, ctx=ctx)
    assert len(sol.F) == 8
    U = D.normalize_input(prob.B, prob.U)
    for s in (0, 3, 7):
        ivp = D.forward_zono(prob.A, splits[s], 0.002, U)
        ref = GO.post_GLGM06(ivp.Phi, ivp.Omega0, nsteps, 10, ivp.V)
        for k in (0, 1, 2, 9, 10, 11, 30, 59):
            R = sol.F[s][k]
            check_zonotope((R.set.center, R.set.generators), ref[k], "split %d step %d" % (s, k))
    # ρ(e25, ·) over the whole batch on the device == max over oracle zonotopes
    d = np.zeros(48)
    d[24] = 1.0
    plan = sol.F.ext["plan"]
    sup = plan.support(d)
    ivp = D.forward_zono(prob.A, splits[5], 0.002, U)
    ref = GO.post_GLGM06(ivp.Phi, ivp.Omega0, nsteps, 10, ivp.V)
    want = np.array([Z.rho_zonotope(d, c, G) for c, G in ref])
    np.testing.assert_allclose(sup[5], want, rtol=1e-10, atol=1e-10 * np.abs(want).max())
    assert abs(rb.rho(d, sol.F[5]) - want.max()) <= 1e-10 * abs(want.max())


def test_initial_order_reduction_post_jl_26(ctx):
    """post.jl:26: Ω₀ with more than max_order·n generators is reduced before the loop."""
    rng = np.random.default_rng(5)
    n = 6
    A = stable(n, 2, 0.05)
    import scipy.linalg
    Phi = scipy.linalg.expm(A * 0.05)
    Om = rb.Zonotope(rng.standard_normal(n), rng.standard_normal((n, 40)))
    V = rb.BallInf(np.zeros(n), 0.01)
    for r in (2, 4):
        plan = _run_batch(GLGM06(δ=0.05, max_order=r), Phi, [Om, Om], V, 20, ctx)
        ref = GO.post_GLGM06(Phi, Om, 20, r, V)
        for k in range(20):
            check_zonotope(plan.fetch(1, k), ref[k], "r=%d step %d" % (r, k))


def test_reduce_order_gir05_kernel(ctx):
    rng = np.random.default_rng(9)
    n, p = 7, 60
    G = rng.standard_normal((n, p))
    G[:, 4] = 0.0
    G[2, 4] = 3.0                 # axis-aligned generators tie at weight 0: stability decides
    G[:, 17] = 0.0
    G[5, 17] = -1.0
    Zt = rb.Zonotope(rng.standard_normal(n), G)
    for r in (1, 2, 5, 9):
        Zr, sel = rb.reduce_order(Zt, r, ctx=ctx, with_selection=True)
        cr, Gr, keep = Z.reduce_order_gir05(Zt.center, G, r, return_selection=True)
        check_zonotope((Zr.center, Zr.generators), (cr, Gr), "r=%d" % r)
        if keep is not None and r > 1:
            np.testing.assert_array_equal(sel[sel >= 0], keep)
    with pytest.raises(ValueError):
        rb.reduce_order(Zt, 0, ctx=ctx)


def test_cut_inside_a_tie_group(ctx):
    """SURVEY §7 hard part 5: when the cut lands inside a group of exactly tied weights (many
    axis-aligned generators), index order must decide, as Julia's stable sortperm does."""
    n = 4
    G = np.hstack([np.diag([1.0, 2.0, 3.0, 4.0]), np.diag([0.5, 0.25, 0.125, 0.0625]),
                   np.diag([7.0, 6.0, 5.0, 4.5]), np.array([[1.0, 1.0, 0.0, 0.0]]).T])
    Zt = rb.Zonotope(np.zeros(n), G)              # 13 generators, 12 of weight exactly 0
    Zr, sel = rb.reduce_order(Zt, 2, ctx=ctx, with_selection=True)
    cr, Gr, keep = Z.reduce_order_gir05(Zt.center, G, 2, return_selection=True)
    np.testing.assert_array_equal(sel[sel >= 0], keep)
    check_zonotope((Zr.center, Zr.generators), (cr, Gr))


def test_glgm06_errors(ctx):
    with pytest.raises(ValueError):
        rb.GLGM06(δ=0.1, max_order=0)
    A = stable(3, 1, 0.1)
    with pytest.raises(ValueError):
        rb.solve(rb.IVP(A, rb.BallInf(np.zeros(3), 1.0)), alg=rb.GLGM06(δ=0.1), ctx=ctx)   # no T / NSTEPS


def test_invariant_early_termination(ctx):
    """reach_inhomog.jl:46-86 / reach_homog.jl:76-147: stop before the first set disjoint from X."""
    n = 4
    A = np.diag([-1.0, -2.0, -0.5, -1.5])
    X0 = rb.Hyperrectangle(np.full(n, 5.0), np.full(n, 0.5))
    inv = rb.HalfSpace(-np.eye(n)[0], -2.0)                       # x1 >= 2
    for U in (None, rb.BallInf(np.zeros(n), 0.01)):
        prob = rb.IVP(A, X0, None, U, inv)
        sol = rb.solve(prob, NSTEPS=200, alg=rb.GLGM06(δ=0.05, max_order=3, approx_model=rb.Forward(setops="zono")), ctx=ctx)
        ivp = D.forward_zono(A, X0, 0.05, U)
        ref = GO.post_GLGM06(ivp.Phi, ivp.Omega0, 200, 3, ivp.V)
        kept = GO.first_disjoint_step(ref, inv)
        assert 1 < kept < 200 and len(sol) == kept
        check_zonotope((sol[kept - 1].set.center, sol[kept - 1].set.generators), ref[kept - 1])
        d = np.eye(n)[1]
        want = max(Z.rho_zonotope(d, c, G) for c, G in ref[:kept])
        assert abs(rb.rho(d, sol) - want) <= 1e-10 * abs(want)
