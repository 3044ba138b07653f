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
from .flowpipe import ReachSolution, MixedFlowpipe, zeroT
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


def solve(ivp, T=None, alg=None, NSTEPS=None, tspan=None, algorithm=None, opC=None, threading=False,
          ctx=None, device_discretize=False):
    """`solve(ivp; T, alg)` (solve.jl:49-76); a list of initial sets runs every split through the
    same Φ / input chain in one batch (`_solve_distributed`, solve.jl:103-117) and returns a
    MixedFlowpipe."""
    if device_discretize:
        # Φ, Φ₁, Φ₂ on the device too (reach_discretize_phi) instead of the host `exp`
        from . import _lib as _L
        ctx = ctx or _L.default_context()
        prev = D.set_device_context(ctx)
        try:
            return solve(ivp, T, alg, NSTEPS, tspan, algorithm, opC, threading, ctx, False)
        finally:
            D.set_device_context(prev)
    alg = alg or algorithm or opC
    if alg is None:
        raise ValueError("an algorithm (`alg=LGG09(...)` or `alg=GLGM06(...)`) is required")
    if tspan is not None:
        # _get_T (solve.jl:272-282): the linear algorithms only take time spans of the form (0, T)
        assert tspan[0] == 0, "this algorithm can only handle zero initial time"
        T = tspan[1]
        dt0 = zeroT
    else:
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
