"""Parity of the CUDA LGG09 path (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest

import reach_b200 as rb
from reach_b200 import discretization as D, workloads as W
from reach_b200.lgg09 import reach_inhomog_LGG09, reach_homog_LGG09, post, Lgg09Plan
from oracle import lgg09_oracle as LO
from conftest import building_problem, iss_problem, assert_support_parity

pytestmark = pytest.mark.gpu


def stable_system(n, seed, delta=0.05):
    rng = np.random.default_rng(seed)
    A = -np.eye(n) + (0.5 / np.sqrt(n)) * rng.standard_normal((n, n))
    import scipy.linalg
    return A, scipy.linalg.expm(A * delta)


def random_sets(n, seed):
    rng = np.random.default_rng(seed + 1000)
    X0 = rb.Hyperrectangle(rng.uniform(-1, 1, n), 0.1 * rng.uniform(0, 1, n))
    q = max(1, n // 3)
    Bm = rng.standard_normal((n, q))
    U0 = rb.Hyperrectangle(rng.uniform(-0.2, 0.2, q), 0.05 * rng.uniform(0, 1, q))
    Z = rb.Zonotope(rng.uniform(-1, 1, n), 0.1 * rng.standard_normal((n, n + 3)))
    return X0, rb.LinearMap(Bm, U0), Z


@pytest.mark.parametrize("n,ndirs,nsteps", [(1, 2, 7), (5, 10, 40), (13, 7, 33), (48, 96, 64), (100, 37, 50)])
@pytest.mark.parametrize("S", [1, 4, 7])
def test_random_forward_inhomog(ctx, n, ndirs, nsteps, S):
    A, _ = stable_system(n, n)
    X0, U, _ = random_sets(n, n)
    ivp = D.forward(A, X0, 0.05, U)
    rng = np.random.default_rng(7)
    dirs = np.hstack([np.eye(n), -np.eye(n)])[:, :ndirs] if ndirs <= 2 * n else rng.standard_normal((n, ndirs))
    if ndirs % 2:
        dirs = np.hstack([dirs[:, :-1], rng.standard_normal((n, 1))])
    ref = LO.reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, ivp.V)
    plan = reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, None, ivp.V, ctx=ctx, time_block=S)
    assert plan.stats()["time_block"] == S
    assert_support_parity(plan.sfmat(), ref)


@pytest.mark.parametrize("tile", [(128, 128), (64, 64), (16, 128), (64, 128)])
@pytest.mark.parametrize("share", [True, False])
def test_tile_configs_and_chain_sharing(ctx, tile, share):
    n, nsteps = 70, 45
    A, _ = stable_system(n, 3)
    X0, U, _ = random_sets(n, 3)
    ivp = D.forward(A, X0, 0.05, U)
    rng = np.random.default_rng(11)
    dirs = np.hstack([np.eye(n), -np.eye(n), rng.standard_normal((n, 5))])
    ref = LO.reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, ivp.V)
    plan = reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, None, ivp.V, ctx=ctx, time_block=2,
                               tile_m=tile[0], tile_n=tile[1], share_negated=share)
    st = plan.stats()
    assert (st["tile_m"], st["tile_n"]) == tile
    assert st["nchains"] == (n + 5 if share else 2 * n + 5)
    assert_support_parity(plan.sfmat(), ref)


def test_chain_sharing_is_bit_identical(ctx):
    n, nsteps = 33, 60
    A, _ = stable_system(n, 5)
    X0, U, _ = random_sets(n, 5)
    ivp = D.forward(A, X0, 0.05, U)
    dirs = np.hstack([np.eye(n), -np.eye(n)])
    a = reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, None, ivp.V, ctx=ctx, time_block=1, share_negated=True).sfmat()
    b = reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, None, ivp.V, ctx=ctx, time_block=1, share_negated=False).sfmat()
    assert np.array_equal(a, b)


@pytest.mark.parametrize("S", [1, 8])
def test_homogeneous_set_types(ctx, S):
    n, nsteps = 20, 50
    A, Phi = stable_system(n, 9)
    X0, U, Z = random_sets(n, 9)
    dirs = rb.OctDirections(4).matrix()
    dirs = np.vstack([dirs, np.zeros((n - 4, dirs.shape[1]))])
    cases = {
        "zonotope": Z,
        "box": X0,
        "ballinf": rb.BallInf(np.arange(n) / n, 0.3),
        "singleton": rb.Singleton(np.linspace(-1, 1, n)),
        "forward_homog": D.forward(A, X0, 0.05).Omega0,
        "ch_of_zono_and_mapped_zono": rb.ConvexHull(Z, rb.MinkowskiSum(rb.LinearMap(Phi, Z), X0)),
        "cartesian": rb.CartesianProduct(rb.Interval(-1, 2), rb.Hyperrectangle(np.zeros(n - 1), np.ones(n - 1))),
        "scaled_mapped": rb.Scale(0.5, rb.LinearMap(np.random.default_rng(1).standard_normal((n, n)), Z)),
    }
    for name, Om in cases.items():
        ref = LO.reach_homog_LGG09(dirs, Om, Phi, nsteps)
        got = reach_homog_LGG09(dirs, Om, Phi, nsteps, ctx=ctx, time_block=S).sfmat()
        assert_support_parity(got, ref), name


def test_zero_initial_set_and_nsteps_one(ctx):
    n = 6
    A, Phi = stable_system(n, 2)
    X0, U, Z = random_sets(n, 2)
    dirs = np.hstack([np.eye(n), -np.eye(n)])
    ref = LO.reach_inhomog_LGG09(dirs, rb.ZeroSet(n), Phi, 30, U)
    got = reach_inhomog_LGG09(dirs, rb.ZeroSet(n), Phi, 30, None, U, ctx=ctx).sfmat()
    assert_support_parity(got, ref)
    for Om, V in ((X0, None), (X0, U), (D.forward(A, X0, 0.05, U).Omega0, D.forward(A, X0, 0.05, U).V)):
        ref = LO.reach_inhomog_LGG09(dirs, Om, Phi, 1, V) if V is not None else LO.reach_homog_LGG09(dirs, Om, Phi, 1)
        got = reach_inhomog_LGG09(dirs, Om, Phi, 1, None, V, ctx=ctx).sfmat()
        assert_support_parity(got, ref)


def test_building_bldf01_reference_assertions(ctx):
    """test/models/generated/Building.jl:41-47 (Forward, δ=0.004) and :49-55 (NoBloating, δ=0.01)."""
    prob = building_problem()
    x25 = np.zeros(48)
    x25[24] = 1.0
    sol = rb.solve(prob, T=20.0, alg=rb.LGG09(δ=0.004, vars=(25,), n=48), ctx=ctx)
    assert len(sol) == 5000 and sol.F.support_function_matrix().shape == (2, 5000)
    assert rb.rho(x25, sol) <= 5.1e-3
    assert not (rb.rho(x25, sol) <= 4e-3)
    assert not (rb.rho(x25, sol(20.0)) <= -0.78e-3)
    ivp = D.forward(prob.A, prob.X0, 0.004, D.normalize_input(prob.B, prob.U))
    ref = LO.reach_inhomog_LGG09(LO.vars_directions(25, 48), ivp.Omega0, ivp.Phi, 5000, ivp.V)
    assert_support_parity(sol.F.support_function_matrix(), ref)
    assert abs(rb.rho(x25, sol) - 4.8602388967852826e-3) < 1e-12

    sol = rb.solve(prob, T=20.0, alg=rb.LGG09(δ=0.01, vars=(25,), n=48, approx_model=rb.NoBloating()), ctx=ctx)
    assert rb.rho(x25, sol) <= 5.1e-3
    assert not (rb.rho(x25, sol) <= 4e-3)
    assert not (rb.rho(x25, sol(20.0)) <= -0.78e-3)


def test_building_box_template_c1(ctx):
    """BASELINE config C1: BLDF01, box template (96 dirs), δ=0.002, 10^4 steps; oracle on the
    first 2000 steps, all time-block settings agree on the rest."""
    prob = building_problem()
    ivp = D.forward(prob.A, prob.X0, 0.002, D.normalize_input(prob.B, prob.U))
    dirs = rb.BoxDirections(48).matrix()
    ref = LO.reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, 2000, ivp.V)
    full = {}
    for S in (1, 0):
        plan = reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, 10000, None, ivp.V, ctx=ctx, time_block=S)
        full[S] = plan.sfmat()
        assert_support_parity(full[S][:, :2000], ref)
    assert_support_parity(full[0], full[1])


def test_iss_issf01_two_directions(ctx):
    """examples/ISS/ISS.jl:117-120: template ±C₃, δ=6e-4 (first 3000 steps against the oracle)."""
    prob, Cm = iss_problem()
    ivp = D.forward(prob.A, prob.X0, 6e-4, D.normalize_input(prob.B, prob.U))
    dirs = np.stack([Cm[2], -Cm[2]], axis=1)
    ref = LO.reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, 3000, ivp.V)
    got = reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, 3000, None, ivp.V, ctx=ctx).sfmat()
    assert_support_parity(got, ref)


def test_invariant_early_termination(ctx):
    n = 4
    A = np.diag([-1.0, -2.0, -0.5, -1.5])
    import scipy.linalg
    Phi = scipy.linalg.expm(A * 0.1)
    X0 = rb.Hyperrectangle(np.full(n, 5.0), np.full(n, 0.5))
    dirs = rb.BoxDirections(n).matrix()
    inv = rb.HalfSpace(-np.eye(n)[0], -2.0)          # x1 >= 2
    ref, k = LO.reach_LGG09_invariant(dirs, X0, Phi, 100, inv)
    plan = reach_homog_LGG09(dirs, X0, Phi, 100, X=inv, ctx=ctx)
    assert plan.nsteps_done == k and 0 < k < 100
    assert_support_parity(plan.sfmat(), ref)
    box = rb.Hyperrectangle(np.full(n, 4.0), np.full(n, 3.0))   # x_i in [1, 7]
    ref, k = LO.reach_LGG09_invariant(dirs, X0, Phi, 100, box)
    plan = reach_homog_LGG09(dirs, X0, Phi, 100, X=box, ctx=ctx)
    assert plan.nsteps_done == k and 0 < k < 100


def test_post_api_and_template_reach_set(ctx):
    """test/algorithms/LGG09.jl:62-117,151-170 (shapes) and test/reachsets/reachsets.jl:20-29."""
    A, _ = stable_system(5, 1)
    X0 = rb.BallInf(np.ones(5), 0.1)
    for tmpl in ("box", "oct"):
        sol = rb.solve(rb.IVP(A, X0), T=1.0, alg=rb.LGG09(δ=0.01, template=tmpl, n=5), ctx=ctx)
        assert len(sol) == 100 and sol[0].dim == 5
        assert sol.F.support_function_matrix().shape == (10 if tmpl == "box" else 50, 100)
    R = sol[3]
    d = rb.BoxDirections(5).matrix()[:, 1]
    assert R.rho(2 * d) == 2 * R.rho(d)
    assert abs(sol.tend - 1.0) < 1e-9 and sol.tstart == 0.0
    x = rb.sigma(d, sol)                                         # σ(d, fp): a point of the reach-set attaining ρ(d, fp)
    assert abs(d @ x - rb.rho(d, sol)) <= 1e-9 * max(1.0, abs(rb.rho(d, sol)))
    assert tuple(rb.shift(sol, 1.0).tspan) == tuple(sol.tspan + 1.0)                # test/algorithms/GLGM06.jl:85
    with pytest.raises(AssertionError):
        post(rb.LGG09(δ=0.01, template="box", n=4), D.no_bloating(A, X0, 0.01), 10, ctx=ctx)
    with pytest.raises(ValueError):
        post(rb.LGG09(δ=0.01, template="box", n=5), D.no_bloating(A, X0, 0.01), ctx=ctx)


def test_error_behaviour(ctx):
    A, Phi = stable_system(3, 1)
    X0 = rb.BallInf(np.zeros(3), 1.0)
    with pytest.raises(ValueError):
        reach_homog_LGG09(np.eye(3), X0, Phi, 0, ctx=ctx)
    with pytest.raises(ValueError):
        reach_homog_LGG09(np.eye(3), rb.Universe(3), Phi, 5, ctx=ctx)
    with pytest.raises(ValueError):
        reach_homog_LGG09(np.eye(3), X0, Phi, 5, ctx=ctx, time_block=4096)
    bad = Phi.copy()
    bad[0, 0] = np.nan
    with pytest.raises(ValueError):
        reach_homog_LGG09(np.eye(3), X0, bad, 5, ctx=ctx)


@pytest.mark.parametrize("n,ndirs,nsteps", [(1, 2, 7), (5, 10, 40), (33, 40, 64), (48, 96, 200), (100, 37, 50), (120, 12, 30)])
def test_persistent_chain_kernel_matches_oracle(ctx, n, ndirs, nsteps):
    """path=2: warp-per-chain kernel with Φᵀ in shared memory (ELL), strictly sequential association."""
    A, _ = stable_system(n, n + 1)
    X0, U, Z = random_sets(n, n + 1)
    ivp = D.forward(A, X0, 0.05, U)
    rng = np.random.default_rng(17)
    dirs = np.hstack([np.eye(n), -np.eye(n)])[:, :ndirs] if ndirs <= 2 * n else rng.standard_normal((n, ndirs))
    dirs[:, -1] = rng.standard_normal(n)
    if n // 3 <= 8:            # the mapped input has n//3 rows; the chain kernel holds at most 8
        ref = LO.reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, ivp.V)
        plan = reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, None, ivp.V, ctx=ctx, path="chain")
    else:                      # more mapped rows than the kernel holds: use the homogeneous problem
        hom = D.forward(A, X0, 0.05)
        ref = LO.reach_homog_LGG09(dirs, hom.Omega0, hom.Phi, nsteps)
        plan = reach_homog_LGG09(dirs, hom.Omega0, hom.Phi, nsteps, ctx=ctx, path="chain")
    assert plan.stats()["tile_m"] == 0           # the chain kernel ran, not the GEMM path
    assert_support_parity(plan.sfmat(), ref)


def test_chain_kernel_zonotope_sets_and_nsteps_one(ctx):
    n = 12
    A, Phi = stable_system(n, 4)
    X0, U, Z = random_sets(n, 4)
    Zs = rb.Zonotope(Z.center, Z.generators[:, :6])
    dirs = np.hstack([np.eye(n), -np.eye(n)])
    for Om, V, k in ((Zs, None, 25), (rb.ConvexHull(X0, rb.LinearMap(Phi, Zs)), rb.BallInf(np.zeros(n), 0.1), 25),
                     (X0, None, 1), (rb.ZeroSet(n), rb.BallInf(np.zeros(n), 0.1), 9)):
        ref = LO.reach_inhomog_LGG09(dirs, Om, Phi, k, V) if V is not None else LO.reach_homog_LGG09(dirs, Om, Phi, k)
        got = reach_inhomog_LGG09(dirs, Om, Phi, k, None, V, ctx=ctx, path="chain").sfmat()
        assert_support_parity(got, ref)


def test_iss_sparse_phi_takes_the_chain_path(ctx):
    """ISS: Φ has 540 non-zeros of 72 900; the auto-tuner must pick the persistent chain kernel
    (the device analogue of the reference's `sparse=true`), and both paths agree with the oracle."""
    prob, Cm = iss_problem()
    ivp = D.forward(prob.A, prob.X0, 6e-4, D.normalize_input(prob.B, prob.U))
    assert np.count_nonzero(ivp.Phi) == 540
    dirs = rb.BoxDirections(270).matrix()
    ref = LO.reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, 400, ivp.V)
    auto = reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, 400, None, ivp.V, ctx=ctx)
    assert auto.stats()["path"] == "compact"      # 135 decoupled 2x2 oscillators: two rows per box direction
    assert_support_parity(auto.sfmat(), ref)
    chain = reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, 400, None, ivp.V, ctx=ctx, path="chain")
    assert chain.stats()["path"] == "chain"
    assert_support_parity(chain.sfmat(), ref)
    # the reference's own ISS run: the two directions +-C3 couple all 270 rows -> not compactable
    d2 = np.column_stack([Cm[2], -Cm[2]])
    ref2 = LO.reach_inhomog_LGG09(d2, ivp.Omega0, ivp.Phi, 400, ivp.V)
    two = reach_inhomog_LGG09(d2, ivp.Omega0, ivp.Phi, 400, None, ivp.V, ctx=ctx)
    assert two.stats()["path"] == "chain"
    assert_support_parity(two.sfmat(), ref2)
    gemm = reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, 400, None, ivp.V, ctx=ctx, path="gemm")
    assert gemm.stats()["tile_m"] > 0
    assert_support_parity(gemm.sfmat(), ref)


def block_diagonal_system(nblocks, g, seed, delta=0.05):
    """Φ = expm(Aδ) of `nblocks` decoupled g×g systems, rows shuffled so the blocks are not contiguous."""
    import scipy.linalg
    rng = np.random.default_rng(seed)
    n = nblocks * g
    A = np.zeros((n, n))
    for b in range(nblocks):
        A[b * g:(b + 1) * g, b * g:(b + 1) * g] = -np.eye(g) + 0.7 * rng.standard_normal((g, g))
    perm = rng.permutation(n)
    A = A[np.ix_(perm, perm)]
    Phi = np.zeros((n, n))
    inv = np.argsort(perm)
    for b in range(nblocks):
        idx = inv[b * g:(b + 1) * g]
        Phi[np.ix_(idx, idx)] = scipy.linalg.expm(A[np.ix_(idx, idx)] * delta)      # exact zeros off the blocks
    return A, Phi, [inv[b * g:(b + 1) * g] for b in range(nblocks)]


@pytest.mark.parametrize("nblocks,g,nsteps", [(40, 2, 700), (9, 3, 300), (7, 4, 150), (5, 8, 333), (3, 16, 200), (12, 1, 130)])
def test_compact_chains_decoupled_blocks(ctx, nblocks, g, nsteps):
    """path 4: every chain lives in the closure of its support under Φ's sparsity graph; lane groups × time blocks."""
    A, Phi, blocks = block_diagonal_system(nblocks, g, 5 * nblocks + g)
    n = nblocks * g
    X0, U, Z = random_sets(n, n + 2)
    rng = np.random.default_rng(3)
    q = min(8, max(1, n // 3))
    Bm = rng.standard_normal((n, q))
    U0 = rb.Hyperrectangle(rng.uniform(-0.2, 0.2, q), 0.05 * rng.uniform(0, 1, q))
    V = rb.MinkowskiSum(rb.LinearMap(0.05 * Bm, U0), rb.BallInf(np.zeros(n), 1e-3))
    Om = rb.ConvexHull(X0, rb.MinkowskiSum(rb.LinearMap(Phi, X0), V))
    dirs = np.hstack([np.eye(n), -np.eye(n)])
    extra = np.zeros((n, 2))
    extra[blocks[0], 0] = rng.standard_normal(g)                    # a direction inside one block
    if 2 * g <= 16 and nblocks > 1:
        extra[np.r_[blocks[0], blocks[1]], 1] = rng.standard_normal(2 * g)   # spanning two blocks
    else:
        extra[blocks[-1], 1] = rng.standard_normal(g)
    dirs = np.hstack([dirs, extra])
    ref = LO.reach_inhomog_LGG09(dirs, Om, Phi, nsteps, V)
    plan = reach_inhomog_LGG09(dirs, Om, Phi, nsteps, None, V, ctx=ctx)
    assert plan.stats()["path"] == "compact"
    assert_support_parity(plan.sfmat(), ref)
    forced = reach_inhomog_LGG09(dirs, Om, Phi, nsteps, None, V, ctx=ctx, path="compact")
    assert np.array_equal(forced.sfmat(), plan.sfmat())
    gemm = reach_inhomog_LGG09(dirs, Om, Phi, nsteps, None, V, ctx=ctx, path="gemm")
    assert gemm.stats()["path"] == "gemm"
    assert_support_parity(gemm.sfmat(), ref)
    Zs = rb.Zonotope(Z.center, Z.generators[:, :6])              # six dense generators = six mapped rows
    hom_ref = LO.reach_homog_LGG09(dirs, Zs, Phi, nsteps)
    hom = reach_homog_LGG09(dirs, Zs, Phi, nsteps, ctx=ctx)
    assert hom.stats()["path"] == "compact"
    assert_support_parity(hom.sfmat(), hom_ref)


@pytest.mark.parametrize("n,nsteps", [(1, 1), (2, 65), (5, 1000), (13, 64), (16, 129)])
def test_compact_chains_tiny_dense_systems(ctx, n, nsteps):
    A, _ = stable_system(n, n + 40)
    X0, U, _ = random_sets(n, n + 40)
    ivp = D.forward(A, X0, 0.05, U)
    dirs = np.hstack([np.eye(n), -np.eye(n), np.random.default_rng(n).standard_normal((n, 3))])
    ref = LO.reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, ivp.V)
    plan = reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, None, ivp.V, ctx=ctx)
    assert plan.stats()["path"] == "compact"
    assert_support_parity(plan.sfmat(), ref)
    seq = reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, None, ivp.V, ctx=ctx, path="chain")
    assert seq.stats()["path"] == "chain"
    assert_support_parity(plan.sfmat(), seq.sfmat())


def test_compact_path_rejects_coupled_systems(ctx):
    n = 48
    A, Phi = stable_system(n, 2)
    with pytest.raises(NotImplementedError):
        reach_homog_LGG09(np.eye(n)[:, :4], rb.BallInf(np.zeros(n), 1.0), Phi, 3, ctx=ctx, path="compact")
    auto = reach_homog_LGG09(np.eye(n)[:, :4], rb.BallInf(np.zeros(n), 1.0), Phi, 3, ctx=ctx)
    assert auto.stats()["path"] != "compact"


def test_chain_path_rejects_what_it_cannot_hold(ctx):
    n = 600
    A, Phi = stable_system(n, 2)
    with pytest.raises(NotImplementedError):
        reach_homog_LGG09(np.eye(n)[:, :4], rb.BallInf(np.zeros(n), 1.0), Phi, 3, ctx=ctx, path="chain")


def test_to_intervals_flowpipe_jl_311_331(ctx):
    prob = building_problem()
    sol = rb.solve(prob, NSTEPS=300, alg=rb.LGG09(δ=0.004, vars=(25,), n=48), ctx=ctx)
    iv = sol.F.to_intervals()
    mat = sol.F.support_function_matrix()
    assert len(iv) == 300
    for k in (0, 17, 299):
        assert iv[k].set.hi == mat[0, k] and iv[k].set.lo == -mat[1, k]
        assert iv[k].set.lo <= iv[k].set.hi
    with pytest.raises(ValueError):
        sol.F.to_intervals(rows=(1,))


def _decay_problem(n=6, seed=2):
    import scipy.linalg
    rng = np.random.default_rng(seed)
    A = -np.diag(rng.uniform(0.5, 2.0, n)) + 0.2 * rng.standard_normal((n, n))
    Phi = scipy.linalg.expm(A * 0.05)
    X0 = rb.Hyperrectangle(np.full(n, 5.0) + rng.uniform(-1, 1, n), np.full(n, 0.3))
    return Phi, X0, rng


def test_invariant_general_halfspace_with_box_template_matches_the_exact_lp_test(ctx):
    """reach_homog.jl:418 `_isdisjoint(X, set(F[k]))` is an exact polyhedral test (an LP in LazySets).  With a box
    template the device decides any half-space exactly from the +-e_i rows (min a.x over a box is separable)."""
    Phi, X0, rng = _decay_problem()
    n = Phi.shape[0]
    dirs = rb.BoxDirections(n).matrix()
    a = -rng.uniform(0.2, 1.0, n)                        # a.x <= b with a dense normal that is no template direction
    ref_all = LO.reach_homog_LGG09(dirs, X0, Phi, 150)
    lo = -ref_all[n:, :]                                 # lower bounds of the box hull per step
    mins = (a[:, None] * np.where(a[:, None] > 0, lo, ref_all[:n, :])).sum(axis=0)        # min a.x per step
    b = float(0.5 * (mins[5] + mins[60]))                # crossed somewhere between steps 5 and 60
    inv = rb.HalfSpace(a, b)
    ref, k = LO.reach_LGG09_invariant(dirs, X0, Phi, 150, inv, exact=True)
    assert 5 <= k <= 60
    for opts in ({}, {"path": "gemm", "time_block": 4}, {"path": "chain"}):
        plan = reach_homog_LGG09(dirs, X0, Phi, 150, X=inv, ctx=ctx, **opts)
        info = plan.invariant_info()
        assert info["faces"] == 1 and info["covered"] == 1 and info["exact"]
        assert plan.nsteps_done == k, (opts, plan.nsteps_done, k)
        assert_support_parity(plan.sfmat(), ref)
    # several faces (HPolyhedron): each face is tested; exactness is only claimed for one face or box-vs-box
    inv2 = rb.HPolyhedron([rb.HalfSpace(a, b), rb.HalfSpace(-np.eye(n)[1], -1e-3)])
    plan = reach_homog_LGG09(dirs, X0, Phi, 150, X=inv2, ctx=ctx)
    info = plan.invariant_info()
    assert info["faces"] == 2 and info["covered"] == 2 and not info["exact"]
    ref2, k2 = LO.reach_LGG09_invariant(dirs, X0, Phi, 150, inv2, exact=True)
    assert plan.nsteps_done >= k2 and plan.nsteps_done <= k        # never a false stop; no later than the single face
    # box invariant with a box template: exact (axis by axis)
    box = rb.Hyperrectangle(np.full(n, 4.0), np.full(n, 3.0))
    plan = reach_homog_LGG09(dirs, X0, Phi, 150, X=box, ctx=ctx)
    assert plan.invariant_info()["exact"]
    assert plan.nsteps_done == LO.reach_LGG09_invariant(dirs, X0, Phi, 150, box, exact=True)[1]


def test_invariant_sufficient_only_and_unsupported_templates(ctx):
    Phi, X0, rng = _decay_problem(seed=3)
    n = Phi.shape[0]
    a = -rng.uniform(0.2, 1.0, n)
    oct_dirs = rb.OctDirections(n).matrix()
    ref_all = LO.reach_homog_LGG09(oct_dirs, X0, Phi, 150)
    inv = rb.HalfSpace(a, float(a @ (X0.center * 0.35)))
    ref, k = LO.reach_LGG09_invariant(oct_dirs, X0, Phi, 150, inv, exact=True)
    assert 0 < k < 150
    plan = reach_homog_LGG09(oct_dirs, X0, Phi, 150, X=inv, ctx=ctx)
    info = plan.invariant_info()
    assert info["covered"] == 1 and not info["exact"]    # the box rows of the oct template: sufficient, not exact
    assert k <= plan.nsteps_done <= 150                  # may stop later than the reference, never earlier
    assert_support_parity(plan.sfmat()[:, :k], ref)
    # a template without the rows the normal needs: the library refuses instead of silently ignoring the invariant
    some = rng.standard_normal((n, 5))
    with pytest.raises(NotImplementedError):
        reach_homog_LGG09(some, X0, Phi, 50, X=inv, ctx=ctx)
    # ... unless the normal itself is in the template
    with_normal = np.hstack([some, -a[:, None] * 2.5])
    plan = reach_homog_LGG09(with_normal, X0, Phi, 150, X=inv, ctx=ctx)
    assert plan.invariant_info()["exact"]
    assert plan.nsteps_done == LO.reach_LGG09_invariant(with_normal, X0, Phi, 150, inv, exact=True)[1]


def test_invariant_stops_launching_device_work(ctx):
    """Early termination is real: once a step is certified disjoint no further passes are launched."""
    wl = W.building_c1()
    n = wl.n
    full = Lgg09Plan(ctx, n, wl.dirs.shape[1], wl.nsteps, time_block=64, path="gemm")
    full.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
    full.run()
    ref = full.sfmat()
    # x_i <= b for the coordinate whose lower bound -rho(-e_i) peaks latest (Building's transient is over after ~400 of
    # the 10^4 steps), with b between the earlier maximum and that peak: first violated shortly before the peak
    lo = -ref[n:, :]
    i = int(lo.argmax(axis=1).argmax())
    kpk = int(lo[i].argmax())
    assert kpk > 200
    b = float(0.5 * (lo[i, :kpk - 30].max() + lo[i, kpk]))
    lower = lo[i]
    k_first = int(np.nonzero(lower > b)[0][0])
    assert kpk - 30 <= k_first <= kpk
    e = np.zeros(n)
    e[i] = 1.0
    plan = Lgg09Plan(ctx, n, wl.dirs.shape[1], wl.nsteps, time_block=64, path="gemm")
    plan.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V, rb.HalfSpace(e, b))
    plan.run()
    info = plan.invariant_info()
    assert plan.nsteps_done == k_first and info["exact"]
    assert info["steps_computed"] < wl.nsteps // 2                     # stopped within a few passes of the violation
    assert info["steps_computed"] >= k_first
    assert plan.stats()["launches"] < full.stats()["launches"] // 2
    assert np.array_equal(plan.sfmat(), ref[:, :k_first])
    assert np.all(plan.max_over_steps()[0] == ref[:, :k_first].max(axis=1))
    # a violation at step 0 leaves an empty flowpipe: rho(d, fp) over nothing is -inf
    plan0 = Lgg09Plan(ctx, n, wl.dirs.shape[1], 500, path="gemm")
    plan0.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V, rb.HalfSpace(e, float(lower[0] - 1.0)))
    plan0.run()
    assert plan0.nsteps_done == 0 and plan0.sfmat().shape == (2 * n, 0)
    assert np.all(np.isneginf(plan0.max_over_steps()[0]))


@pytest.mark.parametrize("n", [1, 6, 48])
def test_vector_output_twins_single_direction(ctx, n):
    """L5: `reach_homog_dir_LGG09!` / `reach_inhomog_dir_LGG09!` with a Vector output (reach_homog.jl:80-106,
    reach_inhomog.jl:98-128; called with n = 1, NSTEPS = 1 at test/algorithms/LGG09.jl:120-128): ONE direction
    (ndirs = 1 on the device), Φ passed transposed as in the reference."""
    from reach_b200.lgg09 import reach_homog_dir_LGG09, reach_inhomog_dir_LGG09
    if n == 1:
        # the reference's own call: X0, hcat([1.0]), [1.0], NSTEPS = 1, pre-filled vector [1.0]
        X0 = rb.Interval(0.0, 1.0)
        v = reach_homog_dir_LGG09(np.array([1.0]), X0, np.array([[1.0]]), [1.0], 1, cache=False, ctx=ctx)
        assert len(v) == 1 and v[0] == 1.0                                   # ρ([1], [0,1]) = 1
        v = reach_inhomog_dir_LGG09(np.array([1.0]), X0, np.array([[1.0]]), rb.Interval(-0.5, 0.25), [1.0], 1,
                                    cache=True, ctx=ctx)
        assert len(v) == 1 and v[0] == 1.0                                   # s_0 = 0
    A, Phi = stable_system(n, 40 + n)
    X0, U, Z = random_sets(n, 40 + n)
    ivp = D.forward(A, X0, 0.05, U)
    rng = np.random.default_rng(n)
    ell = rng.standard_normal(n)
    for nsteps in (1, 2, 37):
        ref = LO.reach_inhomog_LGG09(ell.reshape(-1, 1), ivp.Omega0, ivp.Phi, nsteps, ivp.V)
        out = np.full(nsteps + 2, np.nan)
        got = reach_inhomog_dir_LGG09(out, ivp.Omega0, ivp.Phi.T, ivp.V, ell, nsteps, ctx=ctx)
        assert got is out and np.all(np.isnan(out[nsteps:]))
        assert_support_parity(out[:nsteps].reshape(1, -1), ref)
        refh = LO.reach_homog_LGG09(ell.reshape(-1, 1), Z, ivp.Phi, nsteps)
        gh = reach_homog_dir_LGG09(np.empty(nsteps), Z, ivp.Phi.T, ell, nsteps, cache=False, ctx=ctx)
        assert_support_parity(gh.reshape(1, -1), refh)
    with pytest.raises(IndexError):
        reach_homog_dir_LGG09(np.empty(3), Z, ivp.Phi.T, ell, 4, ctx=ctx)


def test_rho_of_a_non_template_direction_falls_back_to_the_polyhedron(ctx):
    """ρ(d, fp) for d outside the template: max_k ρ(d, set(F[k])) by LP on the fetched support values
    (TemplateReachSet.jl:112-114, AbstractFlowpipe.jl:35-37).  With a box template the LP optimum is the box hull's
    support function, which the oracle evaluates in closed form."""
    n, nsteps = 6, 25
    A, _ = stable_system(n, 5)
    X0, U, _ = random_sets(n, 5)
    ivp = D.forward(A, X0, 0.05, U)
    alg = rb.LGG09(δ=0.05, template="box", n=n)
    fp = post(alg, ivp, nsteps, ctx=ctx)
    M = fp.support_function_matrix()
    Dm = alg.template.matrix()
    d = np.array([1.0, -0.5, 0.0, 2.0, 0.0, -1.0])
    best = -np.inf
    # rows of +e_i / -e_i
    pos = [next(j for j in range(2 * n) if np.array_equal(Dm[:, j], np.eye(n)[i])) for i in range(n)]
    neg = [next(j for j in range(2 * n) if np.array_equal(Dm[:, j], -np.eye(n)[i])) for i in range(n)]
    for k in range(nsteps):
        val = sum(d[i] * (M[pos[i], k] if d[i] > 0 else -M[neg[i], k]) for i in range(n))
        best = max(best, val)
    assert abs(fp.rho(d) - best) <= 1e-9 * max(1.0, abs(best))
    assert abs(fp[3].rho(d) - sum(d[i] * (M[pos[i], 3] if d[i] > 0 else -M[neg[i], 3]) for i in range(n))) <= 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("n,nsteps,S,pieces,path", [(130, 60, 7, 3, "gemm"), (200, 90, 20, 8, "gemm"), (520, 48, 6, 2, "int8")])
def test_row_stack_built_in_steps_by_row_ranges_is_bit_identical(ctx, n, nsteps, S, pieces, path):
    """reach_lgg09_plan_stack_step: what the ranks of a sharded run do (sharding.build_stack_distributed), here with the row
    ranges of all `pieces` ranks taken one after the other on one GPU -- same products in the same order as plan_run's
    own build, so the support values must be bit-identical; steps out of order and runs without every step are refused."""
    from reach_b200 import workloads as W
    from reach_b200.lgg09 import Lgg09Plan
    from reach_b200.sharding import stack_row_split
    wl = W.synthetic_c5(n=n, nsteps=nsteps, seed=n)
    ref = Lgg09Plan(ctx, n, wl.dirs.shape[1], nsteps, time_block=S, path=path)
    ref.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
    want = ref.run().sfmat().copy()
    plan = Lgg09Plan(ctx, n, wl.dirs.shape[1], nsteps, time_block=S, path=path)
    plan.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
    ptr, nbe, ld, S_ = plan.stack_info()
    assert S_ == S and ptr and ld >= n and nbe >= n
    with pytest.raises(rb._lib.ReachError):
        plan.stack_step(1, 0, 8)                         # step 1 before step 0
    step = 0
    while True:
        rows, dest = plan.stack_step(step)
        if rows == 0:
            break
        assert dest == (nbe if step == 0 else ((1 << (step - 1)) + 1) * nbe)
        for r in range(pieces):
            chunk, lo, hi = stack_row_split(rows, pieces, r)
            plan.stack_step(step, lo, hi)
        step += 1
    plan.stack_ready()
    got = plan.run().sfmat()
    assert np.array_equal(got, want)
    assert np.array_equal(plan.run().sfmat(), want)      # the flag is consumed: this run builds the stack itself
    plan.stack_step(0, 0, nbe)
    with pytest.raises(rb._lib.ReachError):
        plan.stack_ready()                               # not every step taken
