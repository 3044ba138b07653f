"""LGG09 over a vector of initial sets (the split caller, src/Continuous/solve.jl:84-117) on the device:
every split must equal the single-problem oracle run of that split."""
import numpy as np
import pytest

import reach_b200 as rb
from reach_b200 import discretization as D
from reach_b200 import sets as S
from reach_b200.lgg09 import Lgg09BatchPlan, reach_inhomog_LGG09
from oracle import lgg09_oracle as LO
from conftest import building_problem, assert_support_parity
from test_lgg09_gpu import stable_system, random_sets

pytestmark = pytest.mark.gpu


def random_splits(n, nsplits, seed):
    rng = np.random.default_rng(seed)
    return [rb.Hyperrectangle(rng.uniform(-1, 1, n), 0.1 * rng.uniform(0, 1, n)) for _ in range(nsplits)]


@pytest.mark.parametrize("n,nsplits,nsteps", [(1, 3, 5), (5, 1, 40), (13, 7, 90), (48, 19, 130), (100, 4, 33)])
def test_batch_forward_inhomog_matches_oracle_per_split(ctx, n, nsplits, nsteps):
    A, _ = stable_system(n, n + 3)
    _, U, _ = random_sets(n, n + 3)
    ivps = [D.forward(A, X0, 0.05, U) for X0 in random_splits(n, nsplits, n)]
    rng = np.random.default_rng(5)
    dirs = np.hstack([np.eye(n), -np.eye(n), rng.standard_normal((n, 3))])
    batch = Lgg09BatchPlan(ctx, ivps[0].Phi, dirs, [iv.Omega0 for iv in ivps], ivps[0].V, nsteps).run()
    st = batch.stats()
    assert st["alternatives"] == 2 and st["launches"] > 0
    for i, iv in enumerate(ivps):
        ref = LO.reach_inhomog_LGG09(dirs, iv.Omega0, iv.Phi, nsteps, iv.V)
        assert_support_parity(batch.sfmat(i), ref)
    # rho(d, MixedFlowpipe) = max over splits and steps, on the device
    every = np.stack([batch.sfmat(i) for i in range(nsplits)])
    assert np.array_equal(batch.max_over_steps(), every.max(axis=(0, 2)))
    assert np.array_equal(batch.max_over_steps(nsplits - 1), every[-1].max(axis=1))


def test_batch_homogeneous_zonotope_initial_sets(ctx):
    """Per-split mapped rows (dense generators) and a single alternative; no input: s_k = 0."""
    n, nsteps = 12, 70
    A, Phi = stable_system(n, 9)
    rng = np.random.default_rng(9)
    Zs = [rb.Zonotope(rng.uniform(-1, 1, n), 0.1 * rng.standard_normal((n, 4))) for _ in range(5)]
    dirs = np.hstack([np.eye(n), -np.eye(n)])
    batch = Lgg09BatchPlan(ctx, Phi, dirs, Zs, None, nsteps).run()
    assert batch.stats()["alternatives"] == 1
    for i, Z in enumerate(Zs):
        assert_support_parity(batch.sfmat(i), LO.reach_homog_LGG09(dirs, Z, Phi, nsteps))


def test_batch_mixed_alternative_counts_and_single_step(ctx):
    n = 6
    A, Phi = stable_system(n, 21)
    X = random_splits(n, 3, 21)
    V = rb.BallInf(np.zeros(n), 0.02)
    oms = [X[0], rb.ConvexHull(X[1], rb.LinearMap(Phi, X[1])), rb.ConvexHull(X[2], rb.MinkowskiSum(rb.LinearMap(Phi, X[2]), V))]
    dirs = np.hstack([np.eye(n), -np.eye(n)])
    for nsteps in (1, 2, 65):
        batch = Lgg09BatchPlan(ctx, Phi, dirs, oms, V, nsteps).run()
        for i, Om in enumerate(oms):
            assert_support_parity(batch.sfmat(i), LO.reach_inhomog_LGG09(dirs, Om, Phi, nsteps, V))


@pytest.mark.parametrize("nalt", [3, 4, 8])
def test_batch_nested_convex_hulls_use_the_wide_max(ctx, nalt):
    """3..8 alternatives per split (nested convex hulls): the epilogue's max runs over 4 / 8 rows (lanes differing in
    bits 2-4); splits with fewer alternatives repeat their first one."""
    n, nsteps = 9, 50
    A, Phi = stable_system(n, 31)
    rng = np.random.default_rng(31)
    V = rb.BallInf(np.zeros(n), 0.02)

    def hull(k, seed):
        r = np.random.default_rng(seed)
        sets = [rb.Hyperrectangle(r.uniform(-1, 1, n), 0.1 * r.uniform(0, 1, n)) for _ in range(k)]
        sets[1] = rb.MinkowskiSum(rb.LinearMap(Phi, sets[0]), V)          # one alternative evaluated at r_{k+1}
        out = sets[0]
        for X in sets[1:]:
            out = rb.ConvexHull(out, X)
        return out

    oms = [hull(nalt, 1), hull(2, 2), hull(nalt, 3)]
    dirs = np.hstack([np.eye(n), -np.eye(n), rng.standard_normal((n, 2))])
    batch = Lgg09BatchPlan(ctx, Phi, dirs, oms, V, nsteps).run()
    assert batch.stats()["alternatives"] == (4 if nalt <= 4 else 8)
    for i, Om in enumerate(oms):
        assert_support_parity(batch.sfmat(i), LO.reach_inhomog_LGG09(dirs, Om, Phi, nsteps, V))


def test_batch_building_splits_equal_the_single_plan(ctx):
    """Building BLDF01 with X0 split into 16 boxes: the batch equals 16 separate device runs (and the oracle)."""
    prob = building_problem()
    splits = S.split(prob.X0, [2] * 4 + [1] * 44)
    U = D.normalize_input(prob.B, prob.U)
    ivps = [D.forward(prob.A, X0, 0.002, U) for X0 in splits]
    dirs = rb.BoxDirections(48).matrix()
    nsteps = 400
    batch = Lgg09BatchPlan(ctx, ivps[0].Phi, dirs, [iv.Omega0 for iv in ivps], ivps[0].V, nsteps).run()
    for i in (0, 5, 15):
        single = reach_inhomog_LGG09(dirs, ivps[i].Omega0, ivps[i].Phi, nsteps, None, ivps[i].V, ctx=ctx).sfmat()
        assert_support_parity(batch.sfmat(i), single)
    assert_support_parity(batch.sfmat(3), LO.reach_inhomog_LGG09(dirs, ivps[3].Omega0, ivps[3].Phi, nsteps, ivps[3].V))


def test_batch_large_n_uses_the_128_tiles(ctx):
    n, nsteps = 520, 12
    A, Phi = stable_system(n, 2, delta=0.01)
    X = random_splits(n, 3, 2)
    V = rb.BallInf(np.zeros(n), 0.05)
    oms = [rb.ConvexHull(x, rb.MinkowskiSum(rb.LinearMap(Phi, x), V)) for x in X]
    dirs = np.hstack([np.eye(n), -np.eye(n)])[:, :300]
    batch = Lgg09BatchPlan(ctx, Phi, dirs, oms, V, nsteps).run()
    for i in (0, 2):
        assert_support_parity(batch.sfmat(i), LO.reach_inhomog_LGG09(dirs, oms[i], Phi, nsteps, V))


def test_solve_with_a_vector_of_initial_sets_runs_one_batch(ctx):
    """solve(IVP(S, [X0_1..X0_m]); alg=LGG09) -> MixedFlowpipe (solve.jl:84-117, MixedFlowpipe.jl:20-23,66-68)."""
    prob = building_problem()
    splits = S.split(prob.X0, [2, 2] + [1] * 46)
    alg = rb.LGG09(δ=0.004, vars=(25,), n=48)
    sol = rb.solve(rb.IVP(prob.A, splits, prob.B, prob.U), NSTEPS=150, alg=alg, ctx=ctx)
    assert len(sol.F) == 4 and "batch" in sol.F.ext
    singles = [rb.solve(rb.IVP(prob.A, X0, prob.B, prob.U), NSTEPS=150, alg=alg, ctx=ctx) for X0 in splits]
    for fp, one in zip(sol.F, singles):
        assert len(fp) == 150
        assert_support_parity(fp.support_function_matrix(), one.F.support_function_matrix())
        assert fp[7].dt == one.F[7].dt
    e25 = np.zeros(48)
    e25[24] = 1.0
    assert sol.F.rho(e25) == max(fp.rho(e25) for fp in sol.F)
    assert abs(sol.F.rho(e25) - max(one.F.rho(e25) for one in singles)) <= 1e-10 * abs(sol.F.rho(e25))


def test_batch_time_chunks_of_the_column_operand(ctx, monkeypatch):
    """The [y;|y|] column operand is produced in time chunks (2 GB by default); force chunks of one time block so that a
    run crosses many chunk boundaries, including a ragged last one, and compare with the unchunked run."""
    n, nsplits, nsteps = 13, 5, 333
    A, _ = stable_system(n, 77)
    _, U, _ = random_sets(n, 77)
    ivps = [D.forward(A, X0, 0.05, U) for X0 in random_splits(n, nsplits, 77)]
    dirs = np.hstack([np.eye(n), -np.eye(n)])
    whole = Lgg09BatchPlan(ctx, ivps[0].Phi, dirs, [iv.Omega0 for iv in ivps], ivps[0].V, nsteps).run()
    assert whole.stats()["chunk_steps"] >= nsteps
    monkeypatch.setenv("REACH_BATCH_CHUNK_BYTES", "1")
    parts = Lgg09BatchPlan(ctx, ivps[0].Phi, dirs, [iv.Omega0 for iv in ivps], ivps[0].V, nsteps).run()
    st = parts.stats()
    assert st["chunk_steps"] == st["time_block"] < nsteps
    for i in range(nsplits):
        assert np.array_equal(parts.sfmat(i), whole.sfmat(i))
    assert_support_parity(parts.sfmat(2), LO.reach_inhomog_LGG09(dirs, ivps[2].Omega0, ivps[2].Phi, nsteps, ivps[2].V))


def test_batch_argument_errors(ctx):
    n = 4
    A, Phi = stable_system(n, 1)
    X = random_splits(n, 2, 1)
    with pytest.raises(ValueError):
        Lgg09BatchPlan(ctx, Phi, np.eye(n + 1), X, None, 5)
    with pytest.raises(ValueError):
        Lgg09BatchPlan(ctx, Phi, np.eye(n), [], None, 5)
    with pytest.raises(ValueError):
        Lgg09BatchPlan(ctx, Phi, np.eye(n), [rb.BallInf(np.zeros(n + 1), 1.0)], None, 5)
    with pytest.raises(ValueError):
        Lgg09BatchPlan(ctx, Phi, np.eye(n), X, None, 0)
