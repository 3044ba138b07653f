"""INT8 tcgen05 digit-plane GEMM (csrc/oz_gemm.cuh): accuracy of the GEMM itself against NumPy, parity of the LGG09
path built on it against the CPU oracle, and the DMMA guard of the automatically selected path."""
import numpy as np
import pytest
import scipy.linalg

import reach_b200 as rb
from reach_b200 import discretization as D, workloads as W
from reach_b200.lgg09 import reach_inhomog_LGG09, reach_homog_LGG09
from oracle import lgg09_oracle as LO
from conftest import assert_support_parity

pytestmark = pytest.mark.gpu


def _rel(Cm, ref):
    return float(np.abs(Cm - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (7, 5, 3), (128, 128, 32), (130, 129, 33), (300, 200, 1000), (513, 257, 2000)])
def test_digit_plane_gemm_matches_numpy(ctx, m, n, k):
    rng = np.random.default_rng(m * 1000 + n)
    X = rng.standard_normal((m, k)) * np.exp(rng.uniform(-3, 3, (m, 1)))
    Y = rng.standard_normal((n, k)) * np.exp(rng.uniform(-3, 3, (n, 1)))
    ref = X @ Y.T
    bound = np.abs(X).max(1)[:, None] * np.abs(Y).max(1)[None, :] * k
    for T, tol_norm, tol_bound in ((8, 2e-14, 2.0 ** -50), (7, 2e-14, 2.0 ** -50), (6, 4e-12, 2.0 ** -44), (5, 1e-9, 2.0 ** -36),
                                   (4, 3e-7, 2.0 ** -28)):
        Cm, _, _ = ctx.ozgemm_probe(X, Y, slices=T)
        assert np.all(np.isfinite(Cm))
        assert _rel(Cm, ref) <= tol_norm, (T, _rel(Cm, ref))
        # the documented bound: truncation below T K 2^(-8T) of rowmax(X) rowmax(Y), per entry (8-bit digits)
        assert (np.abs(Cm - ref) / bound).max() <= tol_bound, (T, (np.abs(Cm - ref) / bound).max())


def test_digit_plane_gemm_is_exact_on_small_integers(ctx):
    rng = np.random.default_rng(5)
    X = rng.integers(-1000, 1000, (200, 333)).astype(np.float64)
    Y = rng.integers(-1000, 1000, (150, 333)).astype(np.float64)
    Cm, _, _ = ctx.ozgemm_probe(X, Y, slices=8)
    assert np.array_equal(Cm, X @ Y.T)           # 10-bit inputs: nothing is truncated, the sums are exact


def test_digit_plane_gemm_zero_rows_extreme_scales_and_determinism(ctx):
    rng = np.random.default_rng(6)
    X = rng.standard_normal((140, 96))
    Y = rng.standard_normal((260, 96))
    X[3] = 0.0
    Y[17] = 0.0
    X[5] *= 2.0 ** 300
    X[6] *= 2.0 ** -400
    Y[9] *= 2.0 ** -500
    Y[10] *= 2.0 ** 200
    ref = X @ Y.T
    C1, _, _ = ctx.ozgemm_probe(X, Y, slices=8)
    C2, _, _ = ctx.ozgemm_probe(X, Y, slices=8)
    assert np.array_equal(C1, C2)                # integer accumulation: independent of scheduling
    assert np.all(C1[3] == 0.0) and np.all(C1[:, 17] == 0.0)
    scale = np.abs(X).max(1)[:, None] * np.abs(Y).max(1)[None, :] * 96
    ok = scale > 0
    assert (np.abs(C1 - ref)[ok] / scale[ok]).max() <= 2.0 ** -50


def test_digit_plane_gemm_rejects_bad_arguments(ctx):
    X = np.ones((4, 4))
    with pytest.raises(ValueError):
        ctx.ozgemm_probe(X, X, slices=3)
    with pytest.raises(ValueError):
        ctx.ozgemm_probe(X, X, slices=9)


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (100, 64, 37), (300, 200, 1000), (384, 640, 130), (129, 257, 513)])
@pytest.mark.parametrize("moduli", [12, 14, 16])
def test_residue_gemm_matches_exact_products(ctx, m, n, k, moduli):
    """crt_gemm.cuh: one INT8 GEMM per modulus + Chinese remainder reconstruction.  Error model: every operand element is
    rounded at 2^-bits of its ROW NORM (bits = 46 / 54 / 61 for 12 / 14 / 16 moduli), nothing else is inexact."""
    rng = np.random.default_rng(m + n + k)
    X = rng.standard_normal((m, k)) * np.exp(rng.uniform(-3, 3, (m, 1)))
    Y = rng.standard_normal((n, k)) * np.exp(rng.uniform(-3, 3, (n, 1)))
    ref = (X.astype(np.longdouble) @ Y.T.astype(np.longdouble)).astype(np.float64)
    Cm, _, _ = ctx.crtgemm_probe(X, Y, moduli=moduli)
    scale = np.sqrt((X * X).sum(1))[:, None] * np.sqrt((Y * Y).sum(1))[None, :]
    bits = {12: 46, 14: 54, 16: 61}[moduli]
    tol = max(2.0 ** -(bits - 2), 2.0 ** -51)          # 2^-51: the rounding of the FP64 result itself
    assert (np.abs(Cm - ref) / scale).max() <= tol


@pytest.mark.parametrize("m,n,k", [(100, 30000, 96), (384, 40000, 64), (640, 20000, 200)])
def test_residue_gemm_many_tiles_per_cta_pair_and_odd_row_tile_counts(ctx, m, n, k):
    """More 256 x 256 tiles than CTA pairs (each pair walks several tiles: both scratch halves, the slice ring and the
    accumulator hand-over wrap around), with 1 / 3 / 5 row tiles: the second CTA of the last pair multiplies a dummy tile."""
    rng = np.random.default_rng(m + k)
    X = rng.standard_normal((m, k)) * np.exp(rng.uniform(-2, 2, (m, 1)))
    Y = rng.standard_normal((n, k)) * np.exp(rng.uniform(-2, 2, (n, 1)))
    Cm, _, _ = ctx.crtgemm_probe(X, Y)
    ref = X @ Y.T
    scale = np.sqrt((X * X).sum(1))[:, None] * np.sqrt((Y * Y).sum(1))[None, :]
    assert (np.abs(Cm - ref) / scale).max() <= 2.0 ** -49


def test_residue_gemm_is_deterministic_and_handles_zero_and_tiny_rows(ctx):
    rng = np.random.default_rng(7)
    X = rng.standard_normal((260, 300))
    Y = rng.standard_normal((140, 300))
    X[5] = 0.0
    X[9] *= 1e-290
    Y[3] = 0.0
    Y[11] *= 1e250
    X[17, :] = 0.0
    X[17, 4] = 3.0                                   # a single entry: the row norm equals the row maximum
    C1, _, _ = ctx.crtgemm_probe(X, Y)
    C2, _, _ = ctx.crtgemm_probe(X, Y)
    assert np.array_equal(C1, C2)
    ref = X @ Y.T
    assert np.all(C1[5] == 0.0) and np.all(C1[:, 3] == 0.0)
    scale = np.sqrt((X * X).sum(1))[:, None] * np.sqrt((Y * Y).sum(1))[None, :]
    ok = scale > 0
    assert (np.abs(C1 - ref)[ok] / scale[ok]).max() <= 2.0 ** -50


def test_residue_gemm_rejects_bad_arguments(ctx):
    X = np.ones((4, 4))
    with pytest.raises(ValueError):
        ctx.crtgemm_probe(X, X, moduli=9)
    with pytest.raises(ValueError):
        ctx.crtgemm_probe(X, X, moduli=17)


def _synthetic(n, nsteps, seed=0):
    return W.synthetic_c5(n=n, nsteps=nsteps, seed=seed)


@pytest.mark.parametrize("planes,expect", [(0, 14), (7, 7), (13, 13)])   # default = residues mod 14 moduli; 7 digit planes; 13 moduli
@pytest.mark.parametrize("n,nsteps,S", [(130, 30, 1), (200, 48, 3), (333, 40, 0), (600, 24, 5)])
def test_lgg09_int8_path_matches_oracle(ctx, n, nsteps, S, planes, expect):
    wl = _synthetic(n, nsteps, seed=n)
    ref = LO.reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, wl.V)
    plan = reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, time_block=S, path="int8", digit_planes=planes)
    st = plan.stats()
    assert st["path"] == "int8" and st["digit_planes"] == expect and (st["tile_m"], st["tile_n"]) == (128, 128)
    err = assert_support_parity(plan.sfmat(), ref)
    assert err <= 2e-11                          # measured ~1e-12: two decades inside the tolerance
    # same numbers on a second run (deterministic), and the DMMA path agrees
    again = reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, time_block=S, path="int8", digit_planes=planes)
    assert np.array_equal(again.sfmat(), plan.sfmat())
    dmma = reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, time_block=S, path="gemm")
    assert dmma.stats()["path"] == "gemm"
    assert_support_parity(plan.sfmat(), dmma.sfmat(), tol=2e-11)


def test_lgg09_int8_path_homogeneous_custom_directions_and_mapped_rows(ctx):
    n, nsteps = 150, 25
    rng = np.random.default_rng(3)
    A = -np.eye(n) + (0.5 / np.sqrt(n)) * rng.standard_normal((n, n))
    X0 = rb.Hyperrectangle(rng.uniform(-1, 1, n), 0.1 * rng.uniform(0, 1, n))
    q = 3
    U = rb.LinearMap(rng.standard_normal((n, q)), rb.Hyperrectangle(rng.uniform(-0.2, 0.2, q), 0.05 * rng.uniform(0, 1, q)))
    dirs = np.hstack([np.eye(n), -np.eye(n)[:, :40], rng.standard_normal((n, 21))])
    ivp = D.forward(A, X0, 0.05, U)
    ref = LO.reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, ivp.V)
    plan = reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, None, ivp.V, ctx=ctx, path="int8")
    assert plan.stats()["path"] == "int8"
    assert_support_parity(plan.sfmat(), ref)
    ivp_h = D.forward(A, X0, 0.05)
    ref_h = LO.reach_homog_LGG09(dirs, ivp_h.Omega0, ivp_h.Phi, nsteps)
    plan_h = reach_homog_LGG09(dirs, ivp_h.Omega0, ivp_h.Phi, nsteps, None, ctx=ctx, path="int8", time_block=4)
    assert_support_parity(plan_h.sfmat(), ref_h)


def test_lgg09_int8_path_unsupported_shapes_raise(ctx):
    wl = _synthetic(48, 10)
    with pytest.raises(NotImplementedError):     # n + mapped rows < 128
        reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, 10, None, wl.V, ctx=ctx, path="int8")
    wl = _synthetic(160, 10)
    for bad in (3, 9, 17):                       # neither 4..8 digit planes nor 10..16 moduli
        with pytest.raises(ValueError):
            reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, 10, None, wl.V, ctx=ctx, path="int8", digit_planes=bad)


def test_auto_selects_guarded_int8_and_keeps_it_on_a_well_scaled_problem(ctx):
    n, nsteps = 520, 70
    wl = _synthetic(n, nsteps, seed=1)
    ref = LO.reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, wl.V)
    plan = reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, time_block=2)
    st = plan.stats()
    assert st["path"] == "int8" and st["guard_checks"] >= 3 and st["fallback_pass"] == -1
    assert 0.0 < st["guard_max"] <= 1e-12
    assert_support_parity(plan.sfmat(), ref)


def test_auto_falls_back_to_dmma_on_wide_dynamic_range(ctx):
    """1-D heat conduction: the rows of Phi^k decay over many decades away from the diagonal and the initial set only
    touches a few cells, so the support values of far cells are made of entries far below the row maximum -- a
    fixed-point digit-plane representation per row cannot hold them.  The guard must notice at pass 0 and the
    result must still match the oracle to 1e-10 (it then comes from the FP64 tensor pipe)."""
    n, nsteps = 512, 40
    A = (np.diag(-2.0 * np.ones(n)) + np.diag(np.ones(n - 1), 1) + np.diag(np.ones(n - 1), -1)) * 100.0
    c = np.zeros(n)
    r = np.zeros(n)
    c[:6] = 1.0
    r[:6] = 0.1
    X0 = rb.Hyperrectangle(c, r)
    ivp = D.forward(A, X0, 0.01)
    dirs = np.hstack([np.eye(n)[:, :128], -np.eye(n)[:, :128]])      # cells 0..127: values down to ~1e-130, no underflow
    ref = LO.reach_homog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps)
    assert np.abs(ref).max(axis=1).min() > 1e-200
    plan = reach_homog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, None, ctx=ctx)
    st = plan.stats()
    assert st["path"] == "int8" and st["fallback_pass"] == 0 and st["guard_max"] > 1e-12
    assert_support_parity(plan.sfmat(), ref)
    # forced (unguarded) INT8 really is outside the tolerance here: the guard is not decorative
    forced = reach_homog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, None, ctx=ctx, path="int8")
    rowscale = np.abs(ref).max(axis=1, keepdims=True)
    denom = np.maximum(np.abs(ref), rowscale)
    denom[denom == 0] = 1.0
    assert (np.abs(forced.sfmat() - ref) / denom).max() > 1e-10


@pytest.mark.parametrize("nsteps", [1, 2, 3])
def test_lgg09_int8_path_tiny_step_counts(ctx, nsteps):
    wl = _synthetic(140, nsteps, seed=2)
    ref = LO.reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, wl.V)
    for S in (0, 1, 4):
        plan = reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, path="int8", time_block=S)
        assert_support_parity(plan.sfmat(), ref)


def test_lgg09_int8_path_with_invariant_stops_early(ctx):
    """Early termination (reach_inhomog.jl:174-218) is independent of the main-loop kernel."""
    n, nsteps = 136, 60
    wl = _synthetic(n, nsteps, seed=4)
    ref = LO.reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, wl.V)
    # x_1 <= b with b between the first and the last value of rho(-e_1): -rho(-e_1) is the lower bound of x_1
    lower = -ref[n, :]
    b = 0.5 * (lower[0] + lower.max()) if lower.max() > lower[0] else lower[0] - 1.0
    X = rb.HalfSpace(np.eye(n)[0], float(b))
    a = reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, X, wl.V, ctx=ctx, path="int8")
    g = reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, X, wl.V, ctx=ctx, path="gemm")
    assert a.nsteps_done == g.nsteps_done


def test_guard_failure_at_a_late_pass_replays_from_the_last_verified_pass(ctx, monkeypatch):
    """A check that fails at pass t >= 64 must not keep the unverified INT8 passes since the last good check: the state
    kept after that check is restored and every pass since then is recomputed on the FP64 tensor pipe
    (REACH_OZ_TRIP_AT is the library's test hook: the first check at or after that pass is treated as failed)."""
    n, nsteps = 520, 140
    wl = _synthetic(n, nsteps, seed=7)
    ref = LO.reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, wl.V)
    monkeypatch.setenv("REACH_OZ_TRIP_AT", "3")          # checks run at passes 0, 1, 64, 128, last: trips at 64
    plan = reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, time_block=1)
    monkeypatch.delenv("REACH_OZ_TRIP_AT")
    st = plan.stats()
    assert st["path"] == "int8" and st["fallback_pass"] == 64
    assert st["replay_from"] == 2 and st["replayed_passes"] == 63        # passes 0 and 1 were verified, 2..64 are redone
    assert_support_parity(plan.sfmat(), ref)
    # from pass 2 on the numbers come from the DMMA kernel: they agree with the pure DMMA run far inside the tolerance
    dmma = reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, time_block=1, path="gemm")
    assert_support_parity(plan.sfmat(), dmma.sfmat(), tol=1e-12)
    # a failure at pass 0 restarts from the initial state: the whole run is FP64 tensor-pipe output
    monkeypatch.setenv("REACH_OZ_TRIP_AT", "0")
    again = reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, time_block=1)
    monkeypatch.delenv("REACH_OZ_TRIP_AT")
    st = again.stats()
    assert st["fallback_pass"] == 0 and st["replay_from"] == 0 and st["replayed_passes"] == 1
    assert_support_parity(again.sfmat(), dmma.sfmat(), tol=1e-13)
    # time-blocked run with the side stream active, failure at the last check
    monkeypatch.setenv("REACH_OZ_TRIP_AT", "65")
    tb = reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, 700, None, wl.V, ctx=ctx, time_block=4)
    monkeypatch.delenv("REACH_OZ_TRIP_AT")
    st = tb.stats()
    assert st["fallback_pass"] == 128 and st["replay_from"] == 65 and st["replayed_passes"] == 64
    ref700 = LO.reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, 700, wl.V)
    assert_support_parity(tb.sfmat(), ref700)


def _decaying_mode_problem(n=512, slow=0.999, fast=0.7):
    """Phi = Q diag(lambda) Q' (symmetric) with Q the normalised Hadamard matrix (all entries +-1/sqrt(n)): half the
    modes decay like slow^k, the others like fast^k.  Omega0 is the segment [-q_f, q_f] along a fast mode, so
    rho(r_k, Omega0) = |q_f . r_k| = fast^k / sqrt(n) for every box direction, while the rows r_k keep entries of size
    slow^k: the support value is made of ever smaller parts of its operands."""
    Q = scipy.linalg.hadamard(n) / np.sqrt(n)
    lam = np.full(n, slow)
    lam[n // 2:] = fast
    Phi = (Q * lam) @ Q.T
    Phi = 0.5 * (Phi + Phi.T)
    g = Q[:, n - 1].copy()
    return Phi, rb.Zonotope(np.zeros(n), g.reshape(n, 1))


def test_guard_trips_by_itself_at_a_late_pass_when_the_dynamic_range_grows_with_time(ctx):
    """Passes 0 and 1 are harmless (the support value is a large part of its operands), by pass 64 the support value is
    fast^64 ~ 1e-10 of them: the INT8 and the FP64 kernels then differ far above 1e-12 of the feature, the check at
    pass 64 fails, passes 2..64 are recomputed and the run finishes on the FP64 tensor pipe."""
    n, nsteps = 512, 100
    Phi, Om = _decaying_mode_problem(n)
    dirs = np.hstack([np.eye(n), -np.eye(n)])
    ref = LO.reach_homog_LGG09(dirs, Om, Phi, nsteps)
    plan = reach_homog_LGG09(dirs, Om, Phi, nsteps, None, ctx=ctx, time_block=1)
    st = plan.stats()
    assert st["path"] == "int8" and st["fallback_pass"] == 64, st
    assert st["replay_from"] == 2 and st["replayed_passes"] == 63
    assert_support_parity(plan.sfmat(), ref)
    dmma = reach_homog_LGG09(dirs, Om, Phi, nsteps, None, ctx=ctx, time_block=1, path="gemm")
    assert_support_parity(plan.sfmat(), dmma.sfmat(), tol=1e-12)


def test_sentinel_chains_report_the_accumulated_drift(ctx):
    n, nsteps = 520, 300
    wl = _synthetic(n, nsteps, seed=9)
    plan = reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, time_block=2)
    st = plan.stats()
    assert st["path"] == "int8" and st["fallback_pass"] == -1 and st["guard_checks"] >= 5
    assert 0.0 < st["drift_max"] <= 1e-11 and st["replay_from"] == -1 and st["replayed_passes"] == 0
    assert_support_parity(plan.sfmat(), LO.reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, wl.V))


# ---- pass groups: few chains (one shard of a multi-GPU run) -> several passes per pair of launches ----------------

def _few_chains(n, nchains, nsteps, seed):
    wl = _synthetic(n, nsteps, seed=seed)
    dirs = np.hstack([np.eye(n)[:, :nchains], -np.eye(n)[:, :nchains]])
    return wl, dirs


@pytest.mark.parametrize("n,nchains,nsteps,S", [(200, 37, 29, 3), (300, 130, 64, 1), (600, 250, 41, 4), (520, 64, 90, 0)])
def test_pass_groups_forced_int8_match_oracle_and_ungrouped_run(ctx, monkeypatch, n, nchains, nsteps, S):
    """Unguarded INT8 path with few chains: every pass is grouped (J = 16 / 8 / 8 / 16 stacked passes, ragged last group,
    group 0 starts at block 0).  Same numbers as the oracle, and as the run without groups up to the reassociation through
    (Phi')^(jS)."""
    wl, dirs = _few_chains(n, nchains, nsteps, seed=n + nchains)
    rng = np.random.default_rng(1)
    if nchains == 37:                                     # custom directions too: an odd number of chains without negation
        dirs = np.hstack([dirs, rng.standard_normal((n, 3))])
    ref = LO.reach_inhomog_LGG09(dirs, wl.Omega0, wl.Phi, nsteps, wl.V)
    plan = reach_inhomog_LGG09(dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, time_block=S, path="int8")
    st = plan.stats()
    assert st["path"] == "int8"
    passes = -(-(nsteps - 1 + 1) // st["time_block"])     # Forward model: features of r_{k+1} are needed
    assert st["gemm_launches"] < passes, (st["gemm_launches"], passes)       # grouped
    assert_support_parity(plan.sfmat(), ref)
    monkeypatch.setenv("REACH_OZ_NO_GROUPS", "1")
    single = reach_inhomog_LGG09(dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, time_block=S, path="int8")
    monkeypatch.delenv("REACH_OZ_NO_GROUPS")
    assert single.stats()["gemm_launches"] >= passes
    assert_support_parity(plan.sfmat(), single.sfmat(), tol=1e-11)


def test_pass_groups_with_the_guard_and_a_trip_inside_a_grouped_run(ctx, monkeypatch):
    """Auto path (guarded) with 120 chains: checked passes 0, 1, 64, 128, last stay single, the passes between them run in
    groups of up to 16.  A failed check at pass 64 restores the state kept after pass 1 and redoes passes 2..64 on the
    FP64 tensor pipe -- the grouped INT8 passes in between leave nothing behind."""
    n, nchains, nsteps = 520, 120, 333
    wl, dirs = _few_chains(n, nchains, nsteps, seed=11)
    ref = LO.reach_inhomog_LGG09(dirs, wl.Omega0, wl.Phi, nsteps, wl.V)
    plan = reach_inhomog_LGG09(dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, time_block=2)
    st = plan.stats()
    assert st["path"] == "int8" and st["fallback_pass"] == -1 and st["guard_checks"] == 5, st
    assert st["gemm_launches"] <= 5 + 2 + 11, st          # 5 single checked passes + ceil(62/16) + ceil(63/16) + ceil(36/16) groups
    assert 0.0 < st["guard_max"] <= 1e-12 and 0.0 < st["drift_max"] <= 1e-11
    assert_support_parity(plan.sfmat(), ref)
    monkeypatch.setenv("REACH_OZ_TRIP_AT", "3")
    trip = reach_inhomog_LGG09(dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, time_block=2)
    monkeypatch.delenv("REACH_OZ_TRIP_AT")
    st = trip.stats()
    assert st["fallback_pass"] == 64 and st["replay_from"] == 2 and st["replayed_passes"] == 63, st
    assert_support_parity(trip.sfmat(), ref)
    dmma = reach_inhomog_LGG09(dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, time_block=2, path="gemm")
    assert_support_parity(trip.sfmat()[:, 4:], dmma.sfmat()[:, 4:], tol=1e-12)     # steps of passes >= 2: FP64 pipe output


def test_guarded_residue_path_starts_with_13_moduli_and_escalates_on_a_marginal_first_check(ctx, monkeypatch):
    """Auto path: 13 moduli (the digit planes' accuracy class) unless the first guard check is above the escalation
    threshold; then the run restarts at pass 0 with 14 moduli -- no DMMA fallback, same results within the tolerance."""
    n, nsteps = 520, 40
    wl = _synthetic(n, nsteps, seed=5)
    ref = LO.reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, wl.V)
    plan = reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, time_block=4)
    st = plan.stats()
    assert st["path"] == "int8" and st["digit_planes"] == 13 and st["fallback_pass"] == -1 and st["guard_max"] < 2.5e-13
    assert_support_parity(plan.sfmat(), ref)
    monkeypatch.setenv("REACH_CRT_ESCALATE_TOL", "1e-30")            # every first check is "marginal"
    esc = reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, time_block=4)
    se = esc.stats()
    assert se["path"] == "int8" and se["digit_planes"] == 14 and se["fallback_pass"] == -1
    assert se["guard_max"] < st["guard_max"]                         # the guard statistics are those of the 14-modulus run
    assert_support_parity(esc.sfmat(), ref)
    again = esc.run().sfmat()                                        # the plan stays at 14 moduli: same numbers
    assert np.array_equal(again, esc.sfmat()) and esc.stats()["digit_planes"] == 14
    monkeypatch.delenv("REACH_CRT_ESCALATE_TOL")
    monkeypatch.setenv("REACH_CRT_NO_ADAPT", "1")
    fixed = reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, None, wl.V, ctx=ctx, time_block=4)
    assert fixed.stats()["digit_planes"] == 14
    assert np.array_equal(fixed.sfmat(), esc.sfmat())                # an escalated run equals a run that started with 14
