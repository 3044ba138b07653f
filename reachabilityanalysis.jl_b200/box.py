"""BOX: hyperrectangular flowpipes for linear systems (src/Algorithms/BOX/BOX.jl:51-84, post.jl:4-46) on the device path.

`c_k = Φ^(k-1) c_1 + W_k.c`, `r_k = |Φ^(k-1)| r_1 + W_k.r` (reach_homog.jl:8-47, reach_inhomog.jl:8-60): row i of Φ^k is the
direction chain (Φᵀ)^k e_i of the LGG09 recurrence, so `reach_box` runs that recurrence over the box template and writes the
(lin, abs) pair of every chain -- the centre and the radius -- instead of ρ(±e_i) = abs ± lin.  There is no CPU fallback."""
import ctypes as C

import numpy as np

from . import _lib as L
from . import sets as S
from .flowpipe import Flowpipe, ReachSet, zeroT


class BOX:
    """`BOX(; δ, approx_model=Forward(sih=:concrete, exp=:base, setops=:lazy), static=false, dim=missing,
    recursive=false)` (BOX.jl:51-84)."""

    def __init__(self, δ=None, delta=None, approx_model=None, static=False, dim=None, recursive=False):
        d = δ if δ is not None else delta
        if d is None:
            raise TypeError("BOX: the step size δ is required")
        self.δ = float(d)
        if approx_model is None:
            from .lgg09 import Forward
            approx_model = Forward(sih="concrete", exp="base", setops="lazy")
        self.approx_model = approx_model
        self.static = bool(static)
        self.dim = dim
        self.recursive = bool(recursive)

    step_size = property(lambda self: self.δ)


def box_hull(X, ctx=None):
    """`_overapproximate(X, Hyperrectangle)` (post.jl:29,40): the tight box of a lazy set from its support function in
    ±e_i -- a one-step LGG09 run over the box template on the device."""
    if isinstance(X, S.Hyperrectangle):
        return X
    if isinstance(X, S.BallInf):
        return S.Hyperrectangle(X.center, np.full(len(X.center), X.radius))
    from .lgg09 import reach_homog_LGG09
    n = S.dim(X)
    dirs = np.hstack([np.eye(n), -np.eye(n)])
    rho = reach_homog_LGG09(dirs, X, np.eye(n), 1, None, ctx=ctx).sfmat()[:, 0]
    return S.Hyperrectangle(low=-rho[n:], high=rho[:n])


def _disjoint(X, c, r):
    """`_isdisjoint(X, Hyperrectangle(c, r))` for HalfSpace / Hyperrectangle / BallInf / Universe invariants."""
    if X is None or isinstance(X, S.Universe):
        return False
    if isinstance(X, S.HalfSpace):
        return float(X.a @ c - np.abs(X.a) @ r) > X.b
    if isinstance(X, (S.Hyperrectangle, S.BallInf)):
        H = box_hull(X)
        return bool(np.any(c - r > H.center + H.radius) or np.any(c + r < H.center - H.radius))
    raise NotImplementedError("BOX: invariant of type %s is not supported on the device path" % type(X).__name__)


def _reach(c0, r0, Phi, NSTEPS, cU=None, rU=None, recursive=False, ctx=None, **opts):
    ctx = ctx or L.default_context()
    Phi = L.fmat(Phi)
    n = Phi.shape[0]
    c0 = np.ascontiguousarray(c0, dtype=np.float64)
    r0 = np.ascontiguousarray(r0, dtype=np.float64)
    if c0.shape != (n,) or r0.shape != (n,):
        raise AssertionError("the initial box has dimension %d, Φ is %s" % (c0.shape[0], Phi.shape))
    o = L.reach_lgg09_opts()
    ctx.lib.reach_lgg09_default_opts(C.byref(o))
    o.time_block = int(opts.get("time_block", 0))
    o.path = {"auto": 0, "gemm": 1, "chain": 2, "int8": 3, "compact": 4}.get(opts.get("path", 0), opts.get("path", 0))
    cu = None if cU is None else np.ascontiguousarray(cU, dtype=np.float64)
    ru = None if rU is None else np.ascontiguousarray(rU, dtype=np.float64)
    Cm = np.empty((NSTEPS, n))        # == n x NSTEPS column-major
    Rm = np.empty((NSTEPS, n))
    ctx.check(ctx.lib.reach_box(ctx.handle, n, int(NSTEPS), L.dptr(Phi), L.dptr(c0), L.dptr(r0), L.dptr(cu), L.dptr(ru),
                                int(bool(recursive)), C.byref(o), L.dptr(Cm), L.dptr(Rm)))
    return Cm.T, Rm.T


def _fill(Cm, Rm, δ, Δt0, X):
    """Flowpipe of ReachSet{Hyperrectangle}; an invariant stops it at the first disjoint box
    (`_isdisjoint(X, Hₖ) && break`, then `resize!(F, k-1)`)."""
    nvalid = Cm.shape[1]
    if not (X is None or isinstance(X, S.Universe)):
        for k in range(1, Cm.shape[1]):
            if _disjoint(X, Cm[:, k], Rm[:, k]):
                nvalid = k
                break

    def make(k, dt):
        return ReachSet(S.Hyperrectangle(Cm[:, k].copy(), Rm[:, k].copy()), dt)

    return Flowpipe(nvalid, make, δ, Δt0, ext={"centers": Cm[:, :nvalid], "radii": Rm[:, :nvalid]})


def reach_homog_BOX(Ω0, Φ, NSTEPS, δ, X=None, recursive=False, Δt0=zeroT, ctx=None, **opts):
    """`reach_homog_BOX!` (reach_homog.jl:8-47 non-recursive, :50-69 recursive, :72-118 with an invariant)."""
    if recursive and not (X is None or isinstance(X, S.Universe)):
        raise NotImplementedError("BOX: no recursive method with an invariant (as in the reference)")
    Cm, Rm = _reach(Ω0.center, Ω0.radius, Φ, NSTEPS, recursive=recursive, ctx=ctx, **opts)
    return _fill(Cm, Rm, δ, Δt0, X)


def reach_inhomog_BOX(Ω0, Φ, NSTEPS, δ, X, U, recursive=False, Δt0=zeroT, ctx=None, **opts):
    """`reach_inhomog_BOX!` (reach_inhomog.jl:8-60, :63-118 with an invariant); only `recursive=false` exists."""
    if recursive:
        raise NotImplementedError("reach_inhomog_BOX!: no method for recursive::Val{true} (MethodError in the reference)")
    Cm, Rm = _reach(Ω0.center, Ω0.radius, Φ, NSTEPS, U.center, U.radius, ctx=ctx, **opts)
    return _fill(Cm, Rm, δ, Δt0, X)


def post(alg, ivp, NSTEPS=None, Δt0=zeroT, ctx=None, **kwargs):
    """Discrete `post(alg::BOX, ivp::IVP{<:AbstractDiscreteSystem}, NSTEPS; Δt0)` (post.jl:4-46)."""
    if NSTEPS is None:
        NSTEPS = kwargs.pop("NSTEPS", None)
        if NSTEPS is None:
            raise ValueError("`NSTEPS` not specified")
    Ω0 = box_hull(ivp.Omega0, ctx)                    # post.jl:29
    X = getattr(ivp, "X", None)
    if ivp.V is None:
        F = reach_homog_BOX(Ω0, ivp.Phi, NSTEPS, alg.δ, X, alg.recursive, Δt0, ctx=ctx, **kwargs)
    else:
        U = box_hull(ivp.V, ctx)                      # post.jl:40
        F = reach_inhomog_BOX(Ω0, ivp.Phi, NSTEPS, alg.δ, X, U, alg.recursive, Δt0, ctx=ctx, **kwargs)
    return F
