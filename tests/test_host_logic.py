"""Host-side logic: set flattening, templates, discretization, sharding maps (CPU only)."""
import numpy as np
import pytest

import reach_b200 as rb
from reach_b200 import discretization as D, support_program as sp, _lib, sharding
from reach_b200.zonotope_ops import to_zonotope
from oracle import lazysets_oracle as Z, lgg09_oracle as LO


def eval_flat(alts, r0, r1):
    """Evaluate a flattened program in plain NumPy (test helper)."""
    best = None
    for alt in alts:
        v = 0.0
        for t in alt:
            d = r1 if t.which else r0
            if t.M is not None:
                d = t.M.T @ d
            if t.c is not None:
                v += float(t.c @ d)
            if t.kind == _lib.REACH_TERM_BOX and t.r is not None:
                v += float(t.r @ np.abs(d))
            if t.kind == _lib.REACH_TERM_ZONO:
                v += float(np.abs(t.G.T @ d).sum())
        best = v if best is None else max(best, v)
    return best


def test_flatten_matches_lazy_support_function():
    rng = np.random.default_rng(0)
    n = 6
    Phi = rng.standard_normal((n, n))
    X0 = rb.Hyperrectangle(rng.standard_normal(n), rng.uniform(0, 1, n))
    Zt = rb.Zonotope(rng.standard_normal(n), rng.standard_normal((n, 9)))
    U = rb.LinearMap(rng.standard_normal((n, 2)), rb.Hyperrectangle(rng.standard_normal(2), rng.uniform(0, 1, 2)))
    sets = [
        X0, Zt, rb.BallInf(np.ones(n), 0.5), rb.Singleton(np.arange(n)), rb.ZeroSet(n),
        rb.ConvexHull(X0, rb.MinkowskiSum(rb.MinkowskiSum(rb.LinearMap(Phi, X0), rb.MinkowskiSum(rb.Scale(0.1, U), X0)), Zt)),
        rb.MinkowskiSum(rb.ConvexHull(X0, Zt), rb.ConvexHull(rb.LinearMap(Phi, Zt), U)),
        rb.CartesianProduct(rb.Interval(-1, 3), rb.Zonotope(np.zeros(n - 1), rng.standard_normal((n - 1, 4)))),
        rb.Scale(-2.0, rb.LinearMap(rng.standard_normal((n, n)), X0)),
    ]
    for X in sets:
        alts = sp.flatten(X, Phi)
        sp.to_c(alts, n)           # builds the ctypes structs without error
        for _ in range(5):
            r0 = rng.standard_normal(n)
            want = Z.rho(r0, X)
            got = eval_flat(alts, r0, Phi.T @ r0)
            assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), type(X).__name__


def test_flatten_marks_phi_image_as_next_direction():
    n = 3
    Phi = np.arange(9.0).reshape(3, 3)
    X0 = rb.BallInf(np.zeros(n), 1.0)
    alts = sp.flatten(rb.ConvexHull(X0, rb.LinearMap(Phi, X0)), Phi)
    assert [[t.which for t in a] for a in alts] == [[0], [1]]
    assert all(t.M is None for a in alts for t in a)
    alts = sp.flatten(rb.LinearMap(Phi + 1, X0), Phi)
    assert alts[0][0].which == 0 and alts[0][0].M is not None


def test_flatten_rejects_unbounded_and_unknown():
    with pytest.raises(ValueError):
        sp.flatten(rb.Universe(2))
    with pytest.raises(NotImplementedError):
        sp.flatten(rb.HalfSpace([1.0, 0.0], 1.0))


def test_templates():
    assert rb.BoxDirections(3).matrix().shape == (3, 6)
    O = rb.OctDirections(3).matrix()
    assert O.shape == (3, 18) and len({tuple(c) for c in O.T}) == 18
    with pytest.raises(ValueError):
        rb.LGG09(δ=0.1)
    with pytest.raises(ValueError):
        rb.LGG09(δ=0.1, vars=(1,))
    assert len(rb.LGG09(δ=0.1, dirs=[np.array([1.0, 0.0]), np.array([0.0, -1.0])]).template) == 2
    assert len(rb.LGG09(δ=0.1, template=np.array([1.0, 2.0])).template) == 1


def test_forward_discretization_contains_exact_reach_set_samples():
    """Ω₀ of the Forward model covers x(t), t ∈ [0, δ], for sampled x0, u (sanity of the host-side
    restatement of ForwardModule.jl:137-164)."""
    import scipy.linalg
    rng = np.random.default_rng(2)
    n, delta = 4, 0.05
    A = -np.eye(n) + 0.3 * rng.standard_normal((n, n))
    X0 = rb.Hyperrectangle(rng.standard_normal(n), 0.1 * np.ones(n))
    Bm = rng.standard_normal((n, 2))
    U0 = rb.Hyperrectangle(np.array([0.5, -0.2]), np.array([0.1, 0.2]))
    ivp = D.forward(A, X0, delta, D.normalize_input(Bm, U0))
    dirs = rng.standard_normal((n, 12))
    sup = Z.rho(dirs, ivp.Omega0)
    for _ in range(40):
        x0 = X0.center + X0.radius * rng.uniform(-1, 1, n)
        u = U0.center + U0.radius * rng.uniform(-1, 1, 2)
        for t in np.linspace(0, delta, 5):
            M = np.zeros((n + 1, n + 1))
            M[:n, :n] = A * t
            M[:n, n] = Bm @ u * t
            x = (scipy.linalg.expm(M) @ np.r_[x0, 1.0])[:n]
            assert np.all(dirs.T @ x <= sup + 1e-12)


def test_split_order_first_dimension_fastest():
    H = rb.Hyperrectangle(np.zeros(3), np.ones(3))
    parts = rb.split(H, [2, 2, 1])
    assert len(parts) == 4
    np.testing.assert_allclose(parts[0].center, [-0.5, -0.5, 0])
    np.testing.assert_allclose(parts[1].center, [0.5, -0.5, 0])
    np.testing.assert_allclose(parts[2].center, [-0.5, 0.5, 0])
    np.testing.assert_allclose(parts[0].radius, [0.5, 0.5, 1.0])


def test_to_zonotope_skips_flat_dimensions():
    c, G = to_zonotope(rb.Hyperrectangle(np.zeros(4), np.array([1.0, 0.0, 2.0, 0.0])))
    assert G.shape == (4, 2) and G[0, 0] == 1.0 and G[2, 1] == 2.0


def test_shard_directions_keeps_negations_together():
    n = 5
    dirs = np.hstack([np.eye(n), -np.eye(n), np.ones((n, 1))])
    groups = sharding.chain_groups(dirs)
    assert groups[:n] == [[i, n + i] for i in range(n)] and groups[n] == [2 * n]
    seen = []
    for r in range(3):
        local, rows_by_rank, mloc = sharding.shard_directions(dirs, 3, r)
        rows = rows_by_rank[r]
        assert local.shape == (n, mloc)
        np.testing.assert_array_equal(local[:, :len(rows)], dirs[:, rows])
        assert not local[:, len(rows):].any()
        seen += list(rows)
    assert sorted(seen) == list(range(2 * n + 1))
    local, rows_by_rank, mloc = sharding.shard_directions(dirs, 1, 0)
    np.testing.assert_array_equal(local, dirs)


def test_oracle_invariant_variant_stops_and_drops_the_triggering_step():
    n = 2
    Phi = np.diag([0.5, 0.9])
    X0 = rb.Hyperrectangle(np.array([4.0, 4.0]), np.array([0.1, 0.1]))
    dirs = LO.box_directions(n)
    out, k = LO.reach_LGG09_invariant(dirs, X0, Phi, 50, rb.HalfSpace(np.array([-1.0, 0.0]), -1.0))
    # x1 upper bound: 4.1 * 0.5^k ; disjoint from {x1 >= 1} once 4.1*0.5^k < 1  ->  k = 3
    assert k == 3 and out.shape == (4, 3)


def test_nsteps_is_ceil_of_the_floating_point_quotient():
    """compute_nsteps (src/Continuous/solve.jl:284-289): `ceil(Int, T / δ)` -- 9.3 / 0.3 = 31.000000000000004 ⇒ 32."""
    from reach_b200.solve import _nsteps
    assert _nsteps(9.3, 0.3) == 32
    assert _nsteps(20.0, 0.002) == 10000
    assert _nsteps(20.0, 6e-4) == 33334
    with pytest.raises(AssertionError):
        _nsteps(0.0, 0.1)


def test_solve_rejects_nonzero_initial_time():
    """_get_T (solve.jl:272-282): 'this algorithm can only handle zero initial time'."""
    ivp = rb.IVP(-np.eye(2), rb.Hyperrectangle(np.zeros(2), np.ones(2)))
    with pytest.raises(AssertionError, match="zero initial time"):
        rb.solve(ivp, tspan=(1.0, 2.0), alg=rb.LGG09(δ=0.1, template="box", n=2))


def test_template_reach_set_lp_fallback():
    """ρ(d, TemplateReachSet) for a direction outside the template = ρ(d, set(R)) (TemplateReachSet.jl:112-114): an
    LP over the template polyhedron.  Box template ⇒ the LP optimum is the box support function; an oct template is
    tighter; a polyhedron unbounded in d gives +inf."""
    from reach_b200.flowpipe import TemplateReachSet, zeroT, lp_support
    from reach_b200.directions import get_template
    box = get_template("box", 3)
    lo, hi = np.array([-1.0, 0.5, -2.0]), np.array([2.0, 1.5, 3.0])
    sf = np.r_[hi, -lo] if np.array_equal(box.matrix()[:, :3], np.eye(3)) else None
    assert sf is not None
    R = TemplateReachSet(box, sf, zeroT)
    d = np.array([1.0, -2.0, 0.5])
    expect = sum(di * (h if di > 0 else l) for di, l, h in zip(d, lo, hi))
    assert abs(R.rho(d) - expect) < 1e-9
    assert R.rho(2.0 * box.matrix()[:, 1]) == 2.0 * sf[1]              # positive multiple: no LP
    # x <= 1 only: unbounded in y
    assert lp_support(np.array([0.0, 1.0]), np.array([[1.0], [0.0]]), np.array([1.0])) == float("inf")
    # octagon {|x|,|y| <= 1, |x+y| <= 1.5, |x-y| <= 2}: ρ((2,1)) is attained at the vertex (1, 0.5)
    Doct = np.array([[1, 0], [0, 1], [-1, 0], [0, -1], [1, 1], [-1, -1], [1, -1], [-1, 1]], dtype=float).T
    sfo = np.array([1, 1, 1, 1, 1.5, 1.5, 2, 2], dtype=float)
    assert abs(lp_support(np.array([2.0, 1.0]), Doct, sfo) - 2.5) < 1e-9


def test_post_batch_rejects_splits_that_do_not_share_the_input_set():
    """The split callers share Φ and V across the splits (`_solve_distributed`, solve.jl:84-117: same system, different
    X0); a batch whose discretised problems differ in V must be refused, not silently run with the first one."""
    from reach_b200 import lgg09
    A = -np.eye(3)
    X0 = rb.Hyperrectangle(np.zeros(3), 0.1 * np.ones(3))
    a = D.forward(A, X0, 0.1, rb.BallInf(np.zeros(3), 0.05))
    b = D.forward(A, X0, 0.1, rb.BallInf(np.zeros(3), 0.06))
    alg = rb.LGG09(δ=0.1, template="box", n=3)
    with pytest.raises(ValueError, match="share the input set"):
        lgg09.post_batch(alg, [a, b], 5, ctx=object())
    assert lgg09._same_set(a.V, a.V) and not lgg09._same_set(a.V, b.V)


def test_stack_row_split_partitions_every_step_and_stays_inside_the_slack_rows():
    """sharding.stack_row_split: the row ranges the ranks of a sharded run take from one doubling step of the row stack
    (reach_lgg09_plan_stack_step) -- disjoint, complete, chunks of a multiple of 8 rows, and the in-place all-gather of
    world equal chunks overruns the step's rows by fewer than the 256 slack rows the stack buffer ends with."""
    from reach_b200.sharding import stack_row_split
    for world in (1, 2, 3, 4, 8):
        for rows in (8, 48, 270, 2000, 2001, 16000, 52 * 2000):
            covered = []
            chunks = set()
            for r in range(world):
                chunk, lo, hi = stack_row_split(rows, world, r)
                chunks.add(chunk)
                assert chunk % 8 == 0 and 0 <= lo <= hi <= rows and hi - lo <= chunk
                assert lo == min(rows, r * chunk)
                covered.extend(range(lo, hi))
            assert covered == list(range(rows))
            assert len(chunks) == 1
            assert world * chunks.pop() - rows < 256


def _host_flowpipe(n=5, delta=0.1, dt0=None):
    from reach_b200.flowpipe import Flowpipe, ReachSet, zeroT
    from reach_b200.sets import Interval
    return Flowpipe(n, lambda k, dt: ReachSet(Interval(float(k), k + 1.0), dt), delta, dt0 if dt0 is not None else zeroT)


def test_flowpipe_shift_and_time_evaluation_flowpipe_jl_97_169():
    """`tspan(shift(sol, 1.0)) == tspan(sol) + 1.0` (test/algorithms/GLGM06.jl:85); fp(t) returns the set(s) whose time
    span contains t and throws outside (Flowpipe.jl:97-112)."""
    from reach_b200.flowpipe import ReachSolution, shift
    fp = _host_flowpipe()
    sh = shift(fp, 1.0)
    assert tuple(sh.tspan) == tuple(fp.tspan + 1.0)                 # interval addition (outward rounding)
    assert len(sh) == len(fp) and sh.ext is fp.ext
    for k in range(len(fp)):
        assert tuple(sh[k].dt) == tuple(fp[k].dt + 1.0)
        assert sh[k].set.lo == fp[k].set.lo                     # the sets themselves are untouched
    assert tuple(fp.tspan) == (0.0, fp[len(fp) - 1].dt[1])       # the original is not modified
    sol = ReachSolution(fp, "alg")
    assert tuple(shift(sol, 2.0).tspan) == tuple(fp.tspan + 2.0) and shift(sol, 2.0).alg == "alg"
    inner = fp(0.25)
    assert not isinstance(inner, list) and tuple(inner.dt)[0] <= 0.25 <= tuple(inner.dt)[1]
    edge = fp(fp[1].dt[1])                                       # a shared end point belongs to two reach-sets
    assert isinstance(edge, list) and len(edge) == 2
    with pytest.raises(ValueError):
        fp(fp.tspan[1] + 1.0)
    assert len(fp((0.05, 0.25))) == 3
    with pytest.raises(ValueError):
        fp((9.0, 10.0))


def test_mixed_flowpipe_interface_mixedflowpipe_jl_31_68():
    from reach_b200.flowpipe import MixedFlowpipe
    a, b = _host_flowpipe(3), _host_flowpipe(4, dt0=rb.TimeInterval(1.0, 1.0))
    mfp = MixedFlowpipe([a, b])
    assert len(mfp) == 2 and mfp.numrsets() == 7 and mfp[1] is b
    with pytest.raises(ValueError):
        mfp.tspan()
    assert tuple(mfp.tspan(1)) == tuple(b.tspan)
    d = np.array([1.0])
    assert mfp.rho(d) == max(a.rho(d), b.rho(d)) == 4.0          # MixedFlowpipe.jl:66-68


def test_support_vector_of_reach_sets_and_flowpipes_abstractflowpipe_jl_56_58():
    from reach_b200.flowpipe import Flowpipe, ReachSet, TemplateReachSet, zeroT
    Z = [rb.Zonotope(np.array([0.1 * k, 0.0]), np.array([[1.0, 0.2], [0.0, 0.5]])) for k in range(4)]
    fp = Flowpipe(4, lambda k, dt: ReachSet(Z[k], dt), 0.1)
    d = np.ones(2)
    assert fp.argmax_rho(d) == 3 and rb.rho(d, fp) == pytest.approx(2.0)
    np.testing.assert_allclose(rb.sigma(d, fp), [1.5, 0.5])
    assert d @ rb.sigma(d, fp) == pytest.approx(rb.rho(d, fp))
    H = ReachSet(rb.Hyperrectangle([1.0, 2.0], [0.5, 0.25]), zeroT)
    np.testing.assert_array_equal(H.sigma(np.array([1.0, -1.0])), [1.5, 1.75])
    T = TemplateReachSet(rb.BoxDirections(2), np.array([1.0, 2.0, 0.5, 0.0]), zeroT)      # [-0.5, 1] x [0, 2]
    np.testing.assert_allclose(T.sigma(np.array([1.0, 1.0])), [1.0, 2.0])
    np.testing.assert_allclose(T.sigma(np.array([-1.0, 0.5])), [-0.5, 2.0])
    unbounded = TemplateReachSet(rb.CustomDirections([[1.0, 0.0]]), np.array([1.0]), zeroT)
    with pytest.raises(ValueError):
        unbounded.sigma(np.array([0.0, 1.0]))


def test_solve_argument_handling_solve_jl_195_352():
    """test/continuous/solve.jl:80-125 (ways to give the time span, T and tspan together => ArgumentError) and
    :5-21 (default algorithm) against the pure argument helpers of the mirror -- no device call."""
    from reach_b200.solve import _get_tspan, _get_cpost, _default_cpost, _promote_tspan
    for given in ((0.0, 2.0), [0.0, 2.0], np.array([0.0, 2.0]), rb.Interval(0.0, 2.0), rb.TimeInterval(0.0, 2.0), 2.0, 2):
        assert tuple(_promote_tspan(given)) == (0.0, 2.0)
        assert tuple(_get_tspan((), tspan=given)) == (0.0, 2.0)
        assert tuple(_get_tspan((given,))) == (0.0, 2.0)              # solve(p, (0.0, 2.0)), solve(p, 2.0)
    assert tuple(_get_tspan((), T=2.0)) == (0.0, 2.0)
    assert _get_tspan((), NSTEPS=10) is None                          # "defined a posteriori"
    for kw in ({"T": 1.0, "tspan": (0.0, 2.0)}, {"tspan": (0.0, 2.0), "NSTEPS": 3}, {"T": 1.0, "NSTEPS": 3}):
        with pytest.raises(ValueError):                               # solve.jl:220-229
            _get_tspan((), **kw)
    with pytest.raises(ValueError):
        _get_tspan(())                                                # solve.jl:257-261
    with pytest.raises(ValueError):
        _promote_tspan([0.0, 1.0, 2.0])                               # solve.jl:208-210
    a = rb.GLGM06(δ=0.1)
    assert _get_cpost((), alg=a) is a and _get_cpost((), algorithm=a) is a and _get_cpost((), opC=a) is a
    assert _get_cpost((a,)) is a and _get_cpost(((0.0, 1.0), a)) is a and _get_cpost(((0.0, 1.0),)) is None
    with pytest.raises(ValueError):
        _get_cpost((1.0, 2.0, 3.0))
    A = -np.eye(8)
    d = _default_cpost(rb.IVP(A, rb.Hyperrectangle(np.zeros(8), np.ones(8))), rb.TimeInterval(0.0, 0.001))
    assert isinstance(d, rb.GLGM06) and d.δ == 0.001 / 100 and isinstance(d.approx_model, rb.FirstOrderZonotope)
    assert d.static is False and d.max_order == 5                    # test/continuous/solve.jl:5-9
    assert _default_cpost(rb.IVP(A, rb.Singleton(np.ones(8))), rb.TimeInterval(0.0, 1.0), static=True).static is True
    lazy = rb.ConvexHull(rb.BallInf(np.zeros(8), 1.0), rb.BallInf(np.ones(8), 1.0))
    assert isinstance(_default_cpost(rb.IVP(A, lazy), rb.TimeInterval(0.0, 1.0)).approx_model, rb.Forward)
    assert _default_cpost(rb.IVP(A, lazy), rb.TimeInterval(0.0, 1.0), δ=0.25).δ == 0.25
    assert _default_cpost(rb.IVP(A, lazy), rb.TimeInterval(0.0, 1.0), N=50).δ == 1.0 / 50
    with pytest.raises(NotImplementedError):                          # statedim 1 => INT in the reference
        _default_cpost(rb.IVP(-np.eye(1), rb.Interval(0.0, 1.0)), rb.TimeInterval(0.0, 1.0))
    with pytest.raises(ValueError):                                   # both given: refused before any device work
        rb.solve(rb.IVP(A, rb.Singleton(np.ones(8))), T=1.0, tspan=(0.0, 2.0), alg=a)


def test_flowpipe_interface_solve_jl_24_80():
    """The assertions of test/continuous/solve.jl:33-80 on a host flowpipe with δ = 0.1 over [0, 1]: time spans of
    reach-sets and slices, the callable forms, the border case, the dimension.  They need the outward rounding of
    the time stamps: ten rounded-to-nearest additions of 0.1 end below 1.0."""
    F = _host_flowpipe(10, 0.1)
    δ = 0.1
    assert F.tstart == 0.0 and F.tend == pytest.approx(1.0) and F.tend >= 1.0
    assert F[0].tstart == 0.0 and F[0].tend == pytest.approx(0.1)
    assert F[-1].tstart == pytest.approx(0.9) and F[-1].tend == pytest.approx(1.0)
    assert F[0:3].tstart == pytest.approx(0.0) and F[0:3].tend == pytest.approx(3 * δ)
    assert tuple(F[0:3].tspan) == pytest.approx((0.0, 3 * δ))
    assert F((0.0, 1.0)) == F[0:]                                 # :64
    assert F(0.05) == F[0]                                        # :65
    assert F(1.0) == F[-1]                                        # :66
    assert F((0.05, 0.15)) == F[0:2]                              # :67
    assert F(rb.Interval(0.05, 0.15)) == F[0:2]
    assert tuple(F((0.05, 1.05)).tspan) == tuple(F.tspan)         # :70
    with pytest.raises(ValueError):                               # :73
        F((2, 3))
    assert F(0.1) == F[0:2]                                       # :77 two reach-sets on the border
    assert F.dim == 1                                             # :80
    assert F[2] == F[2] and F[2] != F[3]                          # same set, different time span


def test_time_interval_addition_rounds_outwards():
    from reach_b200.flowpipe import TimeInterval
    from fractions import Fraction
    dt = TimeInterval(0.0, 0.1)
    lo, hi = Fraction(0), Fraction(0.1)
    for _ in range(1000):
        dt = dt + 0.1
        lo, hi = lo + Fraction(0.1), hi + Fraction(0.1)
        assert Fraction(dt.lo) <= lo and hi <= Fraction(dt.hi)    # encloses the exact sums of the float 0.1
    assert dt.hi - dt.lo < 0.1 + 1e-10                            # and stays tight: at most one ulp per addition
    assert tuple(TimeInterval(1.0, 2.0) + TimeInterval(0.5, 0.25)) == (1.5, 2.25)    # exact sums are not widened


def test_template_reach_set_interface_templatereachset_jl_57_115():
    from reach_b200.flowpipe import TemplateReachSet, TimeInterval
    R = TemplateReachSet(rb.BoxDirections(2), np.array([1.0, 2.0, 0.5, 0.0]), TimeInterval(0.0, 1.0))
    assert R.dim == 2 and R.vars == (1, 2) and R.directions is R.dirs
    assert tuple(R.tspan) == (0.0, 1.0) and R.tstart == 0.0 and R.tend == 1.0
    assert R.support_functions(1) == 2.0 and list(R.support_functions()) == [1.0, 2.0, 0.5, 0.0]
    P = R.set                                                     # HPolyhedron of one half-space per direction
    assert len(P.constraints) == 4 and rb.dim(P) == 2
    for c, d, b in zip(P.constraints, R.dirs, R.sf):
        np.testing.assert_array_equal(c.a, d)
        assert c.b == b
    assert R.rho(np.array([0.0, 3.0])) == 6.0                     # positive multiple of a template direction
    assert R.rho(np.array([1.0, 1.0])) == pytest.approx(3.0)      # anything else: the LP over set(R)
    assert R == TemplateReachSet(rb.BoxDirections(2), np.array([1.0, 2.0, 0.5, 0.0]), TimeInterval(0.0, 1.0))
    assert R != TemplateReachSet(rb.BoxDirections(2), np.array([1.0, 2.0, 0.5, 0.1]), TimeInterval(0.0, 1.0))


def test_abstract_flowpipe_interface_flowpipes_jl_5_37():
    """Indexing, counts, sets, time domain and variables as test/flowpipes/flowpipes.jl:20-36 asserts them."""
    from reach_b200.flowpipe import Flowpipe, ReachSet
    Zs = [rb.Zonotope(np.array([0.1 * k, 1.0]), np.eye(2)) for k in range(6)]
    fp = Flowpipe(6, lambda k, dt: ReachSet(Zs[k], dt), 0.5)
    assert fp.dim == 2 and fp.vars == (1, 2)                       # :17, :36
    assert fp[0] == next(iter(fp)) and fp[-1] == fp[len(fp) - 1]   # first / last
    assert len(fp) == fp.numrsets() == 6                           # :24
    assert fp.set_of(2) is Zs[2] and fp[2].set is Zs[2]            # :25
    assert fp.set_of(slice(2, 5)) == Zs[2:5] and len(fp.set_of()) == 6 and fp.set_of([0, 5]) == [Zs[0], Zs[5]]
    assert fp.tstart == 0.0 and fp.tend == pytest.approx(3.0) and tuple(fp.tspan) == pytest.approx((0.0, 3.0))
