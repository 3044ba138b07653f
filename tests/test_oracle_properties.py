"""Independent derivations that pin the CPU oracle beyond the reference's known-answer tests (CPU only).

The reference's own tests hold three numeric anchors for this path (tests/test_oracle_kat.py); everything else
rests on the oracle restating the loops correctly.  Here the loops are checked against closed forms that do not share
their code: the exact discrete reach set by explicit matrix powers and vertex / sample enumeration, the soundness
of the GIR05 reduction (the reduced zonotope contains the original), exactness of GLGM06 while no reduction happens,
and the algebraic properties of the support function (positive homogeneity, sub-additivity, monotonicity in time of
the running input sum)."""
import itertools

import numpy as np
import pytest

pytest.importorskip("hypothesis")           # collection must never fail on a box without it
from hypothesis import given, settings, strategies as st  # noqa: E402

import reach_b200 as rb
from oracle import lgg09_oracle as LO, glgm06_oracle as GO, lazysets_oracle as Z


def _system(seed, n):
    rng = np.random.default_rng(seed)
    A = -0.3 * np.eye(n) + 0.4 * rng.standard_normal((n, n)) / np.sqrt(n)
    Phi = np.eye(n) + 0.1 * A + 0.005 * A @ A                  # any matrix will do: the recurrences are algebraic
    c, r = rng.uniform(-1, 1, n), rng.uniform(0.0, 0.5, n)
    cu, ru = 0.1 * rng.uniform(-1, 1, n), 0.05 * rng.uniform(0.0, 1.0, n)
    return rng, Phi, rb.Hyperrectangle(c, r), rb.Hyperrectangle(cu, ru)


def _vertices(H):
    c, r = np.asarray(H.center), np.asarray(H.radius)
    return [c + np.array(s) * r for s in itertools.product((-1.0, 1.0), repeat=len(c))]


@settings(max_examples=25, deadline=None, derandomize=True)
@given(st.integers(0, 10 ** 6), st.integers(1, 4), st.integers(1, 9))
def test_lgg09_homogeneous_equals_the_support_of_the_explicit_reach_set(seed, n, nsteps):
    """ρℓ[j, k] = max over the vertices v of X0 of ℓ_j · Φ^k v (the exact reach set Φ^k X0 is a polytope)."""
    rng, Phi, X0, _ = _system(seed, n)
    dirs = rng.standard_normal((n, 5))
    got = LO.reach_homog_LGG09(dirs, X0, Phi, nsteps)
    V = np.array(_vertices(X0)).T                              # n × 2^n
    for k in range(nsteps):
        want = (dirs.T @ np.linalg.matrix_power(Phi, k) @ V).max(axis=1)
        np.testing.assert_allclose(got[:, k], want, rtol=1e-12, atol=1e-13)


@settings(max_examples=25, deadline=None, derandomize=True)
@given(st.integers(0, 10 ** 6), st.integers(1, 4), st.integers(1, 9))
def test_lgg09_inhomogeneous_equals_the_minkowski_sum_in_closed_form(seed, n, nsteps):
    """Ω_k = Φ^k Ω₀ ⊕ ⊕_{i<k} Φ^i V  ⇒  ρ(ℓ, Ω_k) = ρ((Φ^k)ᵀℓ, Ω₀) + Σ_{i<k} ρ((Φ^i)ᵀℓ, V), each term the box formula
    d·c + |d|·r written out here, not taken from the oracle."""
    rng, Phi, X0, V = _system(seed, n)
    dirs = rng.standard_normal((n, 4))
    got = LO.reach_inhomog_LGG09(dirs, X0, Phi, nsteps, V)

    def box(d, H):
        return float(d @ np.asarray(H.center) + np.abs(d) @ np.asarray(H.radius))

    for j in range(dirs.shape[1]):
        for k in range(nsteps):
            want = box(np.linalg.matrix_power(Phi, k).T @ dirs[:, j], X0)
            want += sum(box(np.linalg.matrix_power(Phi, i).T @ dirs[:, j], V) for i in range(k))
            assert got[j, k] == pytest.approx(want, rel=1e-12, abs=1e-13)
    zero = LO.reach_inhomog_LGG09(dirs, rb.ZeroSet(n), Phi, nsteps, V)      # reach_inhomog.jl:66-81
    np.testing.assert_allclose(zero, got - LO.reach_homog_LGG09(dirs, X0, Phi, nsteps), rtol=0, atol=1e-12)


@settings(max_examples=25, deadline=None, derandomize=True)
@given(st.integers(0, 10 ** 6), st.integers(1, 4))
def test_lgg09_contains_every_simulated_trajectory(seed, n):
    """x_{k+1} = Φ x_k + u_k with x_0 ∈ X0, u_k ∈ V: every sampled trajectory satisfies ℓ·x_k ≤ ρℓ[ℓ, k]."""
    rng, Phi, X0, V = _system(seed, n)
    dirs = LO.box_directions(n)
    rho = LO.reach_inhomog_LGG09(dirs, X0, Phi, 12, V)
    for _ in range(20):
        x = np.asarray(X0.center) + rng.uniform(-1, 1, n) * np.asarray(X0.radius)
        for k in range(12):
            assert np.all(dirs.T @ x <= rho[:, k] + 1e-12)
            x = Phi @ x + np.asarray(V.center) + rng.uniform(-1, 1, n) * np.asarray(V.radius)


@settings(max_examples=25, deadline=None, derandomize=True)
@given(st.integers(0, 10 ** 6), st.integers(1, 5))
def test_support_function_rules_are_positively_homogeneous_and_subadditive(seed, n):
    """ρ(λd, X) = λ ρ(d, X) for λ > 0 and ρ(d₁ + d₂, X) ≤ ρ(d₁, X) + ρ(d₂, X) for every set type the path accepts,
    and the combinators obey their definitions (sum for ⊕ and ×, max for the convex hull, Mᵀd for linear maps)."""
    rng, Phi, H, B = _system(seed, n)
    G = rng.standard_normal((n, n + 2))
    Zt = rb.Zonotope(rng.standard_normal(n), G)
    sets = [H, B, Zt, rb.BallInf(rng.standard_normal(n), 0.3), rb.Singleton(rng.standard_normal(n)),
            rb.MinkowskiSum(H, Zt), rb.ConvexHull(H, Zt), rb.LinearMap(Phi, H), rb.Scale(0.25, Zt)]
    d1, d2 = rng.standard_normal(n), rng.standard_normal(n)
    for X in sets:
        r1, r2 = Z.rho(d1, X), Z.rho(d2, X)
        assert Z.rho(2.5 * d1, X) == pytest.approx(2.5 * r1, rel=1e-12, abs=1e-12)
        assert Z.rho(d1 + d2, X) <= r1 + r2 + 1e-12
        assert Z.rho(d1, X) + Z.rho(-d1, X) >= -1e-12                       # width is non-negative
    assert Z.rho(d1, rb.MinkowskiSum(H, Zt)) == pytest.approx(Z.rho(d1, H) + Z.rho(d1, Zt), rel=1e-13)
    assert Z.rho(d1, rb.ConvexHull(H, Zt)) == max(Z.rho(d1, H), Z.rho(d1, Zt))
    assert Z.rho(d1, rb.LinearMap(Phi, H)) == pytest.approx(Z.rho(Phi.T @ d1, H), rel=1e-13)
    assert Z.rho(d1, rb.Scale(0.25, Zt)) == pytest.approx(0.25 * Z.rho(d1, Zt), rel=1e-13)
    assert Z.rho(d1, rb.ZeroSet(n)) == 0.0
    CP = rb.CartesianProduct(H, Zt)
    dd = np.concatenate([d1, d2])
    assert Z.rho(dd, CP) == pytest.approx(Z.rho(d1, H) + Z.rho(d2, Zt), rel=1e-13)


@settings(max_examples=30, deadline=None, derandomize=True)
@given(st.integers(0, 10 ** 6), st.integers(1, 5), st.integers(1, 4), st.integers(0, 14))
def test_gir05_reduction_contains_the_original_zonotope_and_has_the_documented_shape(seed, n, r, extra):
    """Soundness: ρ(d, reduce_order(Z, r)) ≥ ρ(d, Z) in every direction, equality of centres, exactly
    ⌊n(r−1)⌋ + n generators when a reduction happens, Z itself when r·n ≥ p, and the box part is diagonal and holds
    the column-wise absolute sum of the dropped generators."""
    rng = np.random.default_rng(seed)
    p = n * r + extra
    c, G = rng.standard_normal(n), rng.standard_normal((n, max(p, 1)))
    cr, Gr, keep = Z.reduce_order_gir05(c, G, r, return_selection=True)
    np.testing.assert_array_equal(cr, c)
    if r * n >= G.shape[1]:
        np.testing.assert_array_equal(Gr, G)
        return
    assert Gr.shape == (n, n * (r - 1) + n)
    kept = Gr[:, :n * (r - 1)]
    np.testing.assert_array_equal(kept, G[:, keep])                         # kept columns are copied, in sorted order
    w = Z.gir05_weights(G)
    assert np.all(np.diff(w[keep]) >= 0)                                    # ascending weights
    dropped = np.setdiff1d(np.arange(G.shape[1]), keep)
    assert len(dropped) == G.shape[1] - n * (r - 1)
    assert w[dropped].max() <= w[keep].min() if len(keep) else True         # the smallest weights are boxed
    D = Gr[:, n * (r - 1):]
    np.testing.assert_array_equal(D, np.diag(np.diag(D)))
    np.testing.assert_allclose(np.diag(D), np.abs(G[:, dropped]).sum(axis=1), rtol=1e-13)
    for _ in range(10):
        d = rng.standard_normal(n)
        assert Z.rho_zonotope(d, cr, Gr) >= Z.rho_zonotope(d, c, G) - 1e-12


@settings(max_examples=20, deadline=None, derandomize=True)
@given(st.integers(0, 10 ** 6), st.integers(2, 4), st.integers(2, 7))
def test_glgm06_is_exact_while_no_reduction_happens_and_sound_afterwards(seed, n, nsteps):
    """With max_order large enough that GIR05 never triggers, Z_k is exactly Φ^k Z₀ ⊕ ⊕_{i<k} Φ^i U, so its support
    function equals the closed form; with a small max_order it can only grow (over-approximation)."""
    rng, Phi, X0, U = _system(seed, n)
    dirs = rng.standard_normal((n, 6))
    exact = GO.post_GLGM06(Phi, X0, nsteps, max_order=1000, U=U)
    small = GO.post_GLGM06(Phi, X0, nsteps, max_order=2, U=U)
    closed = LO.reach_inhomog_LGG09(dirs, X0, Phi, nsteps, U)               # pinned above against explicit powers
    for k in range(nsteps):
        for j in range(dirs.shape[1]):
            e = Z.rho_zonotope(dirs[:, j], *exact[k])
            assert e == pytest.approx(closed[j, k], rel=1e-11, abs=1e-12)
            assert Z.rho_zonotope(dirs[:, j], *small[k]) >= e - 1e-11
        assert small[k][1].shape[1] <= max(2 * n, small[0][1].shape[1])     # the order cap holds at every step
    # new generators enter on the left, the accumulated ones stay on the right (setops.jl:7-25)
    c1, G1 = Z.linear_map_minkowski_sum(Phi, Z.to_zonotope(X0), Z.to_zonotope(U))
    c0, G0 = Z.to_zonotope(X0)
    cu, Gu = Z.to_zonotope(U)
    np.testing.assert_allclose(c1, Phi @ c0 + cu, rtol=1e-14)
    np.testing.assert_allclose(G1, np.hstack([Phi @ G0, Gu]), rtol=1e-14)


@settings(max_examples=20, deadline=None, derandomize=True)
@given(st.integers(0, 10 ** 6), st.integers(1, 4), st.integers(2, 8))
def test_glgm06_homogeneous_is_the_linear_image(seed, n, nsteps):
    """reach_homog.jl:33-73: Z_k = Φ Z_{k-1} with no reduction inside the loop: c_k = Φ^k c₀, G_k = Φ^k G₀ up to
    the association of the products."""
    rng, Phi, X0, _ = _system(seed, n)
    c0, G0 = Z.to_zonotope(X0)
    F = GO.reach_homog_GLGM06(c0, G0, Phi, nsteps)
    for k, (c, G) in enumerate(F):
        P = np.linalg.matrix_power(Phi, k)
        np.testing.assert_allclose(c, P @ c0, rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(G, P @ G0, rtol=1e-11, atol=1e-13)
    none = GO.reach_homog_GLGM06(c0, np.zeros((n, 0)), Phi, 3)              # p = 0 ⇒ one zero generator (:52-54)
    assert none[2][1].shape == (n, 1) and not none[2][1].any()
