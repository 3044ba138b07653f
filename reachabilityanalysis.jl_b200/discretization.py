"""Host-side restatement of the three approximation models that *produce* the hot path's input
(Φ, Ω₀, V).  In the reference these stay in Julia and are out of scope for the device
(north_star: "the discretization output … stays drop-in"); they live here only because Julia is
not available to generate inputs for the ARCH-COMP configurations.  Not on the timed path.
The matrix functions Φ, Φ₁, Φ₂ can be computed on the device instead (`set_device_context`,
`device_phi` → `reach_discretize_phi`); SciPy's `expm` is then the checker.

  Forward            : src/Discretization/ForwardModule.jl:86-102 (homog), :137-164 (inhomog)
  NoBloating         : src/Discretization/NoBloatingModule.jl:72-105
  FirstOrderZonotope : src/Discretization/FirstOrderZonotopeModule.jl:73-144
  Φ₁, Φ₂             : src/Discretization/Exponentiation.jl:263-291, 397-427 (block exponentials)
  normalisation      : src/Continuous/normalization.jl:343-354 (x' = Ax + Bu ⇒ U := B·U₀)

`scipy.linalg.expm` stands in for Julia's `exp(::Matrix)`; the two differ at the 1e-16 level,
which is irrelevant because parity of the hot path is defined *given* (Φ, Ω₀, V).
"""
from dataclasses import dataclass
import numpy as np
import scipy.linalg
import scipy.sparse as sp

from . import sets as S
from .zonotope_ops import (to_zonotope, convex_hull_zonotope, symmetric_interval_hull_of_map,
                           norm_inf)


@dataclass
class DiscreteIVP:
    """What `discretize` returns: x_{k+1} = Φ x_k ⊕ V, x_0 ∈ Ω₀, invariant X."""
    Phi: np.ndarray
    Omega0: object
    V: object = None            # None ⇔ homogeneous (`!hasinput(ivp)`, LGG09/post.jl:31)
    X: object = None            # None ⇔ Universe
    delta: float = 0.0

    @property
    def n(self):
        return self.Phi.shape[0]


def _dense(A):
    return A.toarray() if sp.issparse(A) else np.asarray(A, dtype=np.float64)


_memo = {}


def _memoized(kind, A, delta, fn):
    """Φ and Φ₂ depend only on (A, δ): the split caller (solve.jl:107-117 re-discretizes every
    split) reuses them across the initial sets of a batch."""
    key = (kind, A.shape, float(delta), hash(A.tobytes()))
    if key not in _memo:
        if len(_memo) > 8:
            _memo.clear()
        _memo[key] = fn()
    return _memo[key]


_device_ctx = None


def set_device_context(ctx):
    """Route Φ, Φ₁, Φ₂ through the device (`reach_discretize_phi`, csrc/expm.cu) instead of SciPy's `expm`;
    `None` switches back.  Returns the previous setting."""
    global _device_ctx
    prev, _device_ctx = _device_ctx, ctx
    _memo.clear()
    return prev


def device_phi(ctx, A, delta, want=("Phi", "Phi1", "Phi2")):
    """(Φ, Φ₁, Φ₂)(A, δ) on the device (Exponentiation.jl:168-171, 263-291, 397-427); entries not in `want` are None."""
    import ctypes as C
    from . import _lib as L
    A = L.fmat(_dense(A))
    n = A.shape[0]
    if A.shape != (n, n):
        raise ValueError("A must be square, got %s" % (A.shape,))
    outs = {k: (np.empty((n, n), order="F") if k in want else None) for k in ("Phi", "Phi1", "Phi2")}
    s = C.c_int32()
    ctx.check(ctx.lib.reach_discretize_phi(ctx.handle, n, L.dptr(A), float(delta), L.dptr(outs["Phi"]),
                                           L.dptr(outs["Phi1"]), L.dptr(outs["Phi2"]), C.byref(s)))
    return outs["Phi"], outs["Phi1"], outs["Phi2"]


def _exp(A, delta):
    A = _dense(A)
    if _device_ctx is not None:
        return _memoized("exp", A, delta, lambda: device_phi(_device_ctx, A, delta, ("Phi",))[0])
    return _memoized("exp", A, delta, lambda: scipy.linalg.expm(A * delta))


def Phi1(A, delta):
    """Φ₁(A,δ) = Σ δ^{i+1}/(i+1)! A^i via the 2n×2n block exponential (Exponentiation.jl:275-291)."""
    A = _dense(A)
    n = A.shape[0]
    if _device_ctx is not None:
        return device_phi(_device_ctx, A, delta, ("Phi1",))[1]
    P = np.zeros((2 * n, 2 * n))
    P[:n, :n] = A * delta
    P[:n, n:] = delta * np.eye(n)
    return scipy.linalg.expm(P)[:n, n:2 * n]


def Phi2(A, delta):
    """Φ₂(A,δ) = Σ δ^{i+2}/(i+2)! A^i via the 3n×3n block exponential (Exponentiation.jl:407-426)."""
    A = _dense(A)
    n = A.shape[0]

    def compute():
        if _device_ctx is not None:
            return device_phi(_device_ctx, A, delta, ("Phi2",))[2]
        P = np.zeros((3 * n, 3 * n))
        P[:n, :n] = A * delta
        P[:n, n:2 * n] = delta * np.eye(n)
        P[n:2 * n, 2 * n:] = delta * np.eye(n)
        return scipy.linalg.expm(P)[:n, 2 * n:]

    return _memoized("phi2", A, delta, compute)


def normalize_input(B, U0):
    """normalization.jl:343-354: x' = Ax + Bu, u ∈ U₀  ⇒  x' = Ax + u, u ∈ LinearMap(B, U₀)."""
    if B is None:
        return U0
    B = _dense(B)
    if B.ndim == 1:
        B = B.reshape(-1, 1)
    return S.LinearMap(B, U0)


def forward(A, X0, delta, U=None, X=None):
    """`discretize(ivp, δ, Forward())` with the default lazy set operations.
    U is the *normalised* input set (dimension n) or None."""
    Ad = _dense(A)
    Phi = _exp(Ad, delta)
    P2 = Phi2(np.abs(Ad), delta)
    Eplus = symmetric_interval_hull_of_map(
        P2, symmetric_interval_hull_of_map(Ad @ Ad, X0))           # ForwardModule.jl:94 / :147
    if U is None:
        Omega0 = S.ConvexHull(X0, S.MinkowskiSum(S.LinearMap(Phi, X0), Eplus))   # :96
        return DiscreteIVP(Phi, Omega0, None, X, delta)
    Epsi = symmetric_interval_hull_of_map(P2, symmetric_interval_hull_of_map(Ad, U))   # :150
    V = S.MinkowskiSum(S.Scale(delta, U), Epsi)                                    # :153
    Omega0 = S.ConvexHull(
        X0, S.MinkowskiSum(S.MinkowskiSum(S.LinearMap(Phi, X0), V), Eplus))        # :156
    return DiscreteIVP(Phi, Omega0, V, X, delta)


def forward_zono(A, X0, delta, U=None, X=None):
    """`Forward(setops=:zono)`: the same sets, concretised to zonotopes
    (Discretization/ApplySetops.jl:74-93) -- the route GLGM06 is designed for with inputs."""
    ivp = forward(A, X0, delta, U, X)
    c, G = to_zonotope(ivp.Omega0)
    V = None
    if ivp.V is not None:
        cv, Gv = to_zonotope(ivp.V)
        V = S.Zonotope(cv, Gv)
    return DiscreteIVP(ivp.Phi, S.Zonotope(c, G), V, X, delta)


def no_bloating(A, X0, delta, U=None, X=None):
    """`discretize(ivp, δ, NoBloating())`: Ω₀ = X₀; V = Φ₁(A,δ)·U (NoBloatingModule.jl:72-105)."""
    Ad = _dense(A)
    Phi = _exp(Ad, delta)
    if U is None:
        return DiscreteIVP(Phi, X0, None, X, delta)
    M = Phi1(Ad, delta)
    if isinstance(U, S.Singleton):
        V = S.Singleton(M @ U.element)
    else:
        V = S.LinearMap(M, U)
    return DiscreteIVP(Phi, X0, V, X, delta)


def first_order_zonotope(A, X0, delta, U=None, X=None):
    """`discretize(ivp, δ, FirstOrderZonotope())` (FirstOrderZonotopeModule.jl:73-144)."""
    Ad = _dense(A)
    n = Ad.shape[0]
    norm_A = float(np.abs(Ad).sum(axis=1).max())          # opnorm(A, Inf)
    Phi = _exp(Ad, delta)
    c0, G0 = to_zonotope(X0)
    Zd = (Phi @ c0, Phi @ G0)                             # linear_map(Φ, X0)  :135
    cz, Gz = convex_hull_zonotope((c0, G0), Zd)           # :138
    alpha = (np.exp(norm_A * delta) - 1 - norm_A * delta) * norm_inf(X0)   # :141
    if U is None:
        bloat = alpha
        V = None
    else:
        mu = norm_inf(U)
        beta = mu * delta if abs(norm_A) < 1e-14 else ((np.exp(norm_A * delta) - 1) * mu) / norm_A  # :114
        bloat = alpha + beta
        V = S.BallInf(np.zeros(n), beta)                  # :120
    # minkowski_sum(Z, BallInf(0, bloat)): generators of the ball are bloat·e_i (none if bloat == 0)
    Gb = bloat * np.eye(n) if bloat != 0 else np.zeros((n, 0))
    Omega0 = S.Zonotope(cz, np.hstack([Gz, Gb]))
    return DiscreteIVP(Phi, Omega0, V, X, delta)
