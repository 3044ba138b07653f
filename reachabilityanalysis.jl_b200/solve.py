"""`solve` for linear systems: the caller of the hot path.

Mirrors src/Continuous/solve.jl:49-117 (argument handling, NSTEPS = ceil(T/δ) :284-289,
vector-of-initial-sets form :84-117) and src/Algorithms/linear_post.jl:2-22
(normalize → discretize → discrete post).  Everything here is host code that stays Julia in
the reference; only `post` touches the device."""
import math
from dataclasses import dataclass

import numpy as np

from . import sets as S
from . import discretization as D
from .flowpipe import ReachSolution, MixedFlowpipe, TimeInterval, zeroT
from . import lgg09 as _lgg09


@dataclass
class IVP:
    """x' = A x + B u, u ∈ U, x ∈ X, x(0) ∈ X0  (B, U, X optional)."""
    A: object
    X0: object
    B: object = None
    U: object = None
    X: object = None


def _nsteps(T, delta):
    """compute_nsteps (solve.jl:284-289): `ceil(Int, T / δ)` on the floating-point quotient, exactly as the
    reference -- T = 9.3, δ = 0.3 gives 32 there (9.3 / 0.3 = 31.000000000000004) and here."""
    assert T > 0, "the time horizon should be positive"
    return int(math.ceil(T / delta))


def discretize(ivp, delta, model):
    """`discretize(ivp, δ, approx_model)` for the models the hot path consumes."""
    U = D.normalize_input(ivp.B, ivp.U) if ivp.U is not None else None
    name = getattr(model, "name", type(model).__name__)
    if name == "Forward":
        f = D.forward_zono if getattr(model, "setops", "lazy") == "zono" else D.forward
        return f(ivp.A, ivp.X0, delta, U, ivp.X)
    if name == "NoBloating":
        return D.no_bloating(ivp.A, ivp.X0, delta, U, ivp.X)
    if name == "FirstOrderZonotope":
        return D.first_order_zonotope(ivp.A, ivp.X0, delta, U, ivp.X)
    raise NotImplementedError("approximation model %s" % name)


def post_continuous(alg, ivp, T=None, NSTEPS=None, Δt0=zeroT, ctx=None):
    """linear_post.jl:2-22."""
    delta = alg.δ
    if NSTEPS is None:
        if T is None:
            raise ValueError("the time horizon `T` (or `NSTEPS`) should be specified")
        NSTEPS = _nsteps(T, delta)
    divp = discretize(ivp, delta, alg.approx_model)
    if isinstance(alg, _lgg09.LGG09):
        return _lgg09.post(alg, divp, NSTEPS, Δt0=Δt0, ctx=ctx)
    from . import box as _box
    if isinstance(alg, _box.BOX):
        return _box.post(alg, divp, NSTEPS, Δt0=Δt0, ctx=ctx)
    from . import glgm06 as _glgm06
    return _glgm06.post(alg, divp, NSTEPS, Δt0=Δt0, ctx=ctx)


DEFAULT_NSTEPS = 100          # solve.jl:324: discrete steps when only the time horizon is given


def _promote_tspan(t):
    """`_promote_tspan` (solve.jl:195-212): a number T means [0, T]; tuples, length-2 arrays, LazySets intervals and
    time intervals are taken as they are.  Returns None for anything that is not a time span."""
    if isinstance(t, TimeInterval):
        return t
    if isinstance(t, S.Interval):
        return TimeInterval(t.lo, t.hi)
    if isinstance(t, (bool, np.bool_)):
        return None
    if isinstance(t, (int, float, np.integer, np.floating)):
        return TimeInterval(0.0, t)
    if isinstance(t, (tuple, list, np.ndarray)):
        if len(t) != 2:
            raise ValueError("the length of a time span must be two, but it is of length %d" % len(t))
        return TimeInterval(t[0], t[1])
    return None


def _get_tspan(args, T=None, tspan=None, NSTEPS=None):
    """`_get_tspan` (solve.jl:214-263).  Returns the time interval, or None when only NSTEPS is given (the
    reference's empty interval, "defined a posteriori")."""
    if tspan is not None and T is not None:
        raise ValueError("cannot parse the time horizon `T` and the time span `tspan` simultaneously; "
                         "use one or the other")
    if tspan is not None and NSTEPS is not None:
        raise ValueError("cannot parse the time span `tspan` and the number of steps `NSTEPS` simultaneously; "
                         "use one or the other")
    if T is not None and NSTEPS is not None:
        raise ValueError("cannot parse the time horizon `T` and the number of steps `NSTEPS` simultaneously; "
                         "use one or the other")
    for given in (tspan, T):
        if given is not None:
            dt = _promote_tspan(given)
            if dt is None:
                raise ValueError("cannot parse %r as a time span" % (given,))
            return dt
    if NSTEPS is not None:
        return None
    if args:
        dt = _promote_tspan(args[0])
        if dt is not None:
            return dt
    raise ValueError("the time span has not been specified, but is required for `solve`; you should specify "
                     "either the time horizon `T=...`, the time span `tspan=...`, or the number of steps, "
                     "`NSTEPS=...`")


def _is_post(x):
    from .box import BOX
    from .glgm06 import GLGM06
    return isinstance(x, (_lgg09.LGG09, GLGM06, BOX))


def _get_cpost(args, alg=None, algorithm=None, opC=None):
    """`_get_cpost` (solve.jl:291-317): keyword first, then the first or second positional argument."""
    for given in (alg, algorithm, opC):
        if given is not None:
            return given
    args = [a for a in args if a is not None] if args and args[0] is not None else []
    if args and _is_post(args[0]):
        return args[0]
    if len(args) == 2 and _is_post(args[1]):
        return args[1]
    if len(args) > 2:
        raise ValueError("the number of arguments, %d, is not valid" % len(args))
    return None


def _default_cpost(ivp, dt, static=False, δ=None, N=None, NSTEPS=None, num_steps=None):
    """`_default_cpost` for affine systems (solve.jl:326-352): step δ = diam(Δt) / 100 unless `δ`, `N`, `NSTEPS` or
    `num_steps` says otherwise; GLGM06 with the default FirstOrderZonotope model when X0 is zonotope-like, with
    Forward() otherwise.  One-dimensional systems default to INT in the reference, which is outside this path."""
    from .glgm06 import GLGM06
    if δ is None:
        steps = next((v for v in (N, NSTEPS, num_steps) if v is not None), DEFAULT_NSTEPS)
        if dt is None:
            raise ValueError("the default algorithm needs a time span to choose its step size")
        δ = (dt.hi - dt.lo) / steps
    if np.shape(ivp.A)[0] == 1:
        raise NotImplementedError("one-dimensional systems default to the INT algorithm, which is outside the "
                                  "LGG09 / GLGM06 device path; pass `alg=` explicitly")
    X0 = ivp.X0[0] if isinstance(ivp.X0, (list, tuple)) else ivp.X0
    zonotope_like = isinstance(X0, (S.Hyperrectangle, S.BallInf, S.Interval, S.Singleton, S.Zonotope, S.ZeroSet))
    if zonotope_like:
        return GLGM06(δ=δ, static=static)
    return GLGM06(δ=δ, static=static, approx_model=_lgg09.Forward())


def solve(ivp, *args, T=None, alg=None, NSTEPS=None, tspan=None, algorithm=None, opC=None, threading=False,
          ctx=None, device_discretize=False, static=False, δ=None, N=None, num_steps=None):
    """`solve(ivp, args...; kwargs...)` (solve.jl:49-76): the time span comes from `tspan`, `T`, `NSTEPS` or the first
    positional argument (`_get_tspan`), the algorithm from `alg` / `algorithm` / `opC` or a positional argument
    (`_get_cpost`), else the reference's default for affine systems (`_default_cpost`).  A list of initial sets runs
    every split through the same Φ / input chain in one batch (`_solve_distributed`, solve.jl:103-117) and returns
    a MixedFlowpipe.  `static` given outside the algorithm is ignored when an algorithm is passed
    (test/continuous/solve.jl:19-21)."""
    if device_discretize:
        # Φ, Φ₁, Φ₂ on the device too (reach_discretize_phi) instead of the host `exp`
        from . import _lib as _L
        ctx = ctx or _L.default_context()
        prev = D.set_device_context(ctx)
        try:
            return solve(ivp, *args, T=T, alg=alg, NSTEPS=NSTEPS, tspan=tspan, algorithm=algorithm, opC=opC,
                         threading=threading, ctx=ctx, static=static, δ=δ, N=N, num_steps=num_steps)
        finally:
            D.set_device_context(prev)
    dt = _get_tspan(args, T, tspan, NSTEPS)
    alg = _get_cpost(args, alg, algorithm, opC)
    if alg is None:
        alg = _default_cpost(ivp, dt, static=static, δ=δ, N=N, NSTEPS=NSTEPS, num_steps=num_steps)
    if dt is not None:
        # _get_T (solve.jl:272-282): the linear algorithms only take time spans of the form (0, T)
        assert dt.lo == 0, "this algorithm can only handle zero initial time"
        T = dt.hi
    else:
        T = None
    dt0 = zeroT
    if isinstance(ivp.X0, (list, tuple)):
        from . import box as _box
        if not isinstance(alg, (_lgg09.LGG09, _box.BOX)):
            from . import glgm06 as _glgm06
            F = _glgm06.post_batch(alg, ivp, T=T, NSTEPS=NSTEPS, Δt0=dt0, ctx=ctx)
        elif isinstance(alg, _lgg09.LGG09) and ivp.X is None:
            # Φ, the template and V are shared: one batch on the device (lgg09.post_batch)
            if NSTEPS is None:
                if T is None:
                    raise ValueError("the time horizon `T` (or `NSTEPS`) should be specified")
                NSTEPS = _nsteps(T, alg.δ)
            divps = [discretize(IVP(ivp.A, X0, ivp.B, ivp.U, None), alg.δ, alg.approx_model) for X0 in ivp.X0]
            F = _lgg09.post_batch(alg, divps, NSTEPS, Δt0=dt0, ctx=ctx)
        else:
            F = MixedFlowpipe([post_continuous(alg, IVP(ivp.A, X0, ivp.B, ivp.U, ivp.X), T, NSTEPS, dt0, ctx)
                               for X0 in ivp.X0])
        return ReachSolution(F, alg)
    F = post_continuous(alg, ivp, T, NSTEPS, dt0, ctx)
    return ReachSolution(F, alg)
