"""Pin the CPU oracle against every known-answer test the reference holds for this path
(SURVEY §8c).  CPU only."""
import numpy as np

import reach_b200 as rb
from reach_b200 import discretization as D
from oracle import lgg09_oracle as LO, glgm06_oracle as GO, lazysets_oracle as Z
from conftest import building_problem, model


def test_flowpipes_jl_43_44_harmonic_oscillator_glgm06():
    """test/flowpipes/flowpipes.jl:40-44: x' = [0 1; -1 0] x, X0 = {(1, 0)}, T = 20, default
    algorithm (GLGM06, δ = T/100, FirstOrderZonotope): ρ(ones(2), fp) < 2 and
    σ(ones(2), fp) ≈ [0.7502062459270877, 0.6989729206916551]."""
    A = np.array([[0.0, 1.0], [-1.0, 0.0]])
    ivp = D.first_order_zonotope(A, rb.Singleton([1.0, 0.0]), 0.2)
    F = GO.post_GLGM06(ivp.Phi, ivp.Omega0, 100, max_order=5)
    d = np.ones(2)
    rhos = [Z.rho_zonotope(d, c, G) for c, G in F]
    k = int(np.argmax(rhos))
    assert max(rhos) < 2.0
    assert abs(max(rhos) - 1.4491791666187368) < 1e-14
    np.testing.assert_allclose(Z.sigma_zonotope(d, *F[k]), [0.7502062459270877, 0.6989729206916551], rtol=1e-13)


def test_building_jl_41_55_bldf01_lgg09():
    """test/models/generated/Building.jl:41-47 (Forward, δ=0.004) and :49-55 (NoBloating, δ=0.01)."""
    prob = building_problem()
    U = D.normalize_input(prob.B, prob.U)
    dirs = LO.vars_directions(25, 48)
    ivp = D.forward(prob.A, prob.X0, 0.004, U)
    out = LO.reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, 5000, ivp.V)
    assert out[0].max() <= 5.1e-3 and not (out[0].max() <= 4e-3)
    assert not (out[0, -1] <= -0.78e-3)
    assert abs(out[0].max() - 4.8602388967852826e-3) < 1e-15      # value recorded in SURVEY Appendix A.2
    ivp = D.no_bloating(prob.A, prob.X0, 0.01, U)
    out = LO.reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, 2000, ivp.V)
    assert out[0].max() <= 5.1e-3 and not (out[0].max() <= 4e-3)
    assert not (out[0, -1] <= -0.78e-3)


def test_building_jl_59_72_bldc01_homogeneous():
    """test/models/generated/Building.jl:59-72: constant input as an extra state (n = 49)."""
    m = model("building")
    A, B = m["A"], m["B"]
    Ae = np.zeros((49, 49))
    Ae[:48, :48] = A
    Ae[:48, 48] = B
    c = np.r_[np.full(10, 0.000225), np.zeros(38), 0.9]
    r = np.r_[np.full(10, 0.000025), np.zeros(14), 0.0001, np.zeros(23), 0.1]
    X0 = rb.Hyperrectangle(c, r)
    dirs = LO.vars_directions(25, 49)
    ivp = D.forward(Ae, X0, 0.005)
    out = LO.reach_homog_LGG09(dirs, ivp.Omega0, ivp.Phi, 4000)
    assert out[0].max() <= 5.1e-3 and not (out[0].max() <= 4e-3)
    assert not (out[0, -1] <= -0.78e-3)


def test_generator_count_pins():
    """test/algorithms/GLGM06.jl:61 (linear5D: 16 generators), test/continuous/solve.jl:15-16
    (motor: 13 generators), test/algorithms/GLGM06.jl:102-104 (max_order=1 ⇒ n generators)."""
    rng = np.random.default_rng(0)
    A5 = -np.eye(5) + 0.1 * rng.standard_normal((5, 5))
    X0 = rb.BallInf(np.ones(5), 0.1)
    ivp = D.first_order_zonotope(A5, X0, 0.01)
    assert ivp.Omega0.generators.shape == (5, 16)
    A8 = -np.eye(8) + 0.1 * rng.standard_normal((8, 8))
    X8 = rb.Hyperrectangle(np.r_[0.00225, 0.0, 0.0, 0.0, 0.0, 0.0, 0.00225, 0.0],
                           np.r_[0.00025, 0.0, 0.0, 0.0, 0.0, 0.0, 0.00025, 0.0])
    U8 = rb.Hyperrectangle(np.r_[0.23, 0.3], np.r_[0.07, 0.1])
    B8 = np.zeros((8, 2))
    B8[3, 0], B8[7, 1] = -1.0, -1.0
    ivp = D.first_order_zonotope(A8, X8, 1e-3, D.normalize_input(B8, U8))
    assert ivp.Omega0.generators.shape == (8, 13)
    F = GO.post_GLGM06(ivp.Phi, ivp.Omega0, 30, max_order=1, U=ivp.V)
    assert all(G.shape == (8, 8) for _, G in F[1:])
    # building split for config C4: 2*11 + 1 + 48 generators
    prob = building_problem()
    ivp = D.first_order_zonotope(prob.A, prob.X0, 0.002, D.normalize_input(prob.B, prob.U))
    assert ivp.Omega0.generators.shape == (48, 71)


def test_direction_order_lgg09_jl_52_55():
    """test/algorithms/LGG09.jl:30-60: vars=(1,3), n=5 ⇒ [+e1, +e3, −e1, −e3]."""
    D_ = LO.vars_directions((1, 3), 5)
    e = np.eye(5)
    np.testing.assert_array_equal(D_, np.stack([e[0], e[2], -e[0], -e[2]], axis=1))
    np.testing.assert_array_equal(rb.LGG09(δ=0.1, vars=(1, 3), n=5).template.matrix(), D_)
    np.testing.assert_array_equal(rb.LGG09(δ=0.1, vars=1, n=5).template.matrix(), LO.vars_directions(1, 5))
    np.testing.assert_array_equal(rb.LGG09(δ=0.1, template="box", n=3).template.matrix(), LO.box_directions(3))


def test_template_reach_set_rho_reachsets_jl_20_29():
    """test/reachsets/reachsets.jl:20-29: ρ(d, R) == 2.0 and ρ(2d, R) == 4.0."""
    dirs = rb.BoxDirections(2)
    sf = np.array([2.0, 1.0, 0.0, 3.0])
    R = rb.TemplateReachSet(dirs, sf, rb.TimeInterval(0, 1))
    d = np.array([1.0, 0.0])
    assert R.rho(d) == 2.0 and R.rho(2 * d) == 4.0


def test_gir05_rules():
    """reduce_order(Z, r, GIR05()): unchanged if r·n ≥ p; r == 1 boxes everything; otherwise the
    m = p − ⌊n(r−1)⌋ smallest-weight generators are boxed, ties kept in index order (stable)."""
    rng = np.random.default_rng(1)
    n, p = 4, 14
    G = rng.standard_normal((n, p))
    G[:, 3] = [0, 2.0, 0, 0]          # axis-aligned generators: weight exactly 0
    G[:, 9] = [0, 0, -1.5, 0]
    c = rng.standard_normal(n)
    _, G2 = Z.reduce_order_gir05(c, G, 4)
    assert G2 is G or np.array_equal(G2, G)
    _, G1 = Z.reduce_order_gir05(c, G, 1)
    np.testing.assert_allclose(G1, np.diag(np.abs(G).sum(axis=1)))
    _, G3, keep = Z.reduce_order_gir05(c, G, 3, return_selection=True)
    assert G3.shape == (n, 3 * n)
    w = Z.gir05_weights(G)
    order = np.argsort(w, kind="stable")
    m = p - n * 2
    assert list(order[:2]) == [3, 9]                      # zero-weight ties in index order
    np.testing.assert_array_equal(keep, order[m:])
    np.testing.assert_array_equal(G3[:, :p - m], G[:, order[m:]])
    np.testing.assert_allclose(np.diag(G3[:, p - m:]), np.abs(G[:, order[:m]]).sum(axis=1))
    # the reduced zonotope encloses the original: support function can only grow
    for _ in range(20):
        d = rng.standard_normal(n)
        assert Z.rho_zonotope(d, c, G3) >= Z.rho_zonotope(d, c, G) - 1e-12


def projectile_problem():
    """test/flowpipes/flowpipes.jl:56-70: x' = Ax + u, x(0) = (0, 5, 100, 0), u = (0, 0, 0, -9.81)."""
    import reach_b200 as rb
    A = np.array([[0.0, 0.5, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.7], [0.0, 0.0, 0.0, 0.0]])
    return A, rb.Singleton([0.0, 5.0, 100.0, 0.0]), rb.Singleton([0.0, 0.0, 0.0, -9.81])


def test_flowpipes_jl_56_70_projectile_inclusion_glgm06_forward():
    """`sol ⊆ {24 x1 + x3 <= 375}` for GLGM06(δ=1e-2, Forward()), T = 20: every reach-set's support value in
    the constraint's direction stays below the offset (and the flowpipe really gets close to it)."""
    from reach_b200 import discretization as D
    from oracle import glgm06_oracle as GO, lazysets_oracle as Z
    A, X0, U = projectile_problem()
    ivp = D.forward_zono(A, X0, 1e-2, U)
    F = GO.post_GLGM06(ivp.Phi, ivp.Omega0, 2000, 5, ivp.V)
    d = np.array([24.0, 0.0, 1.0, 0.0])
    rhos = np.array([Z.rho_zonotope(d, c, G) for c, G in F])
    assert len(F) == 2000 and rhos.max() <= 375.0
    assert rhos.max() > 300.0
