"""Host-side set types for the LGG09 / GLGM06 hot path.

These mirror the (LazySets.jl) set types that the reference's discretization hands to the
discrete `post` as Ω₀ and V -- see `src/Discretization/ForwardModule.jl:137-164`
(`ConvexHull(X0, Φ*X0 ⊕ V ⊕ E⁺)`, `δ*U ⊕ Eψ0`), `FirstOrderZonotopeModule.jl:98-127`
(`Zonotope ⊕ BallInf`), `NoBloatingModule.jl:72-105` -- and nothing else.  They are plain
data carriers: no support function is evaluated here.  The device evaluates a *flattened*
form produced by `support_program.flatten_*`; the CPU oracle (`oracle/`) has its own
evaluator that walks these objects by duck typing.
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np

__all__ = [
    "Hyperrectangle", "BallInf", "Interval", "Singleton", "ZeroSet", "Universe", "Zonotope",
    "LinearMap", "Scale", "MinkowskiSum", "ConvexHull", "CartesianProduct", "HalfSpace", "HPolyhedron",
    "dim", "split",
]


def _vec(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))


@dataclass(frozen=True, eq=False)
class Hyperrectangle:
    center: np.ndarray
    radius: np.ndarray

    def __init__(self, center=None, radius=None, *, low=None, high=None):
        if low is not None:
            low, high = _vec(low), _vec(high)
            center, radius = (low + high) / 2, (high - low) / 2
        c, r = _vec(center), _vec(radius)
        if c.shape != r.shape:
            raise ValueError("center and radius must have the same length")
        if np.any(r < 0):
            raise ValueError("radius must be non-negative")
        object.__setattr__(self, "center", c)
        object.__setattr__(self, "radius", r)


@dataclass(frozen=True, eq=False)
class BallInf:
    center: np.ndarray
    radius: float

    def __init__(self, center, radius):
        if radius < 0:
            raise ValueError("radius must be non-negative")
        object.__setattr__(self, "center", _vec(center))
        object.__setattr__(self, "radius", float(radius))


@dataclass(frozen=True, eq=False)
class Interval:
    lo: float
    hi: float

    def __init__(self, lo, hi):
        if hi < lo:
            raise ValueError("empty interval")
        object.__setattr__(self, "lo", float(lo))
        object.__setattr__(self, "hi", float(hi))


@dataclass(frozen=True, eq=False)
class Singleton:
    element: np.ndarray

    def __init__(self, element):
        object.__setattr__(self, "element", _vec(element))


@dataclass(frozen=True, eq=False)
class ZeroSet:
    n: int


@dataclass(frozen=True, eq=False)
class Universe:
    n: int


@dataclass(frozen=True, eq=False)
class Zonotope:
    center: np.ndarray
    generators: np.ndarray  # n x p

    def __init__(self, center, generators):
        c = _vec(center)
        G = np.asarray(generators, dtype=np.float64)
        if G.ndim == 1:
            G = G.reshape(len(c), -1)
        if G.shape[0] != len(c):
            raise ValueError("generator matrix must have dim(Z) rows")
        object.__setattr__(self, "center", c)
        object.__setattr__(self, "generators", np.ascontiguousarray(G))

    @property
    def ngens(self):
        return self.generators.shape[1]

    @property
    def order(self):
        return self.ngens / len(self.center)


@dataclass(frozen=True, eq=False)
class LinearMap:
    M: np.ndarray
    X: object

    def __init__(self, M, X):
        M = np.asarray(M, dtype=np.float64)
        if M.ndim == 1:
            M = M.reshape(-1, 1)
        if M.shape[1] != dim(X):
            raise ValueError("LinearMap: %s does not act on a %d-dimensional set" % (M.shape, dim(X)))
        object.__setattr__(self, "M", M)
        object.__setattr__(self, "X", X)


@dataclass(frozen=True, eq=False)
class Scale:
    """`α * X` (the reference writes `δ * U`, ForwardModule.jl:153)."""
    alpha: float
    X: object


@dataclass(frozen=True, eq=False)
class MinkowskiSum:
    X: object
    Y: object

    def __init__(self, X, Y):
        if dim(X) != dim(Y):
            raise ValueError("MinkowskiSum of sets with different dimensions")
        object.__setattr__(self, "X", X)
        object.__setattr__(self, "Y", Y)


@dataclass(frozen=True, eq=False)
class ConvexHull:
    X: object
    Y: object

    def __init__(self, X, Y):
        if dim(X) != dim(Y):
            raise ValueError("ConvexHull of sets with different dimensions")
        object.__setattr__(self, "X", X)
        object.__setattr__(self, "Y", Y)


@dataclass(frozen=True, eq=False)
class CartesianProduct:
    X: object
    Y: object


@dataclass(frozen=True, eq=False)
class HalfSpace:
    """{x : a·x ≤ b} -- only used as an invariant for the early-terminating variants."""
    a: np.ndarray
    b: float

    def __init__(self, a, b):
        object.__setattr__(self, "a", _vec(a))
        object.__setattr__(self, "b", float(b))


@dataclass(frozen=True, eq=False)
class HPolyhedron:
    """{x : a_i·x ≤ b_i for all i} -- an invariant made of several half-spaces."""
    constraints: tuple

    def __init__(self, constraints):
        cs = tuple(constraints)
        if not cs or not all(isinstance(c, HalfSpace) for c in cs):
            raise ValueError("HPolyhedron needs a non-empty list of HalfSpace constraints")
        object.__setattr__(self, "constraints", cs)


def dim(X) -> int:
    if isinstance(X, (Hyperrectangle, BallInf, Zonotope)):
        return len(X.center)
    if isinstance(X, Interval):
        return 1
    if isinstance(X, Singleton):
        return len(X.element)
    if isinstance(X, (ZeroSet, Universe)):
        return X.n
    if isinstance(X, LinearMap):
        return X.M.shape[0]
    if isinstance(X, Scale):
        return dim(X.X)
    if isinstance(X, (MinkowskiSum, ConvexHull)):
        return dim(X.X)
    if isinstance(X, CartesianProduct):
        return dim(X.X) + dim(X.Y)
    if isinstance(X, HalfSpace):
        return len(X.a)
    if isinstance(X, HPolyhedron):
        return len(X.constraints[0].a)
    raise TypeError("not a set: %r" % (type(X),))


def split(H: Hyperrectangle, parts):
    """`split(H, parts)`: uniform grid of sub-boxes, first dimension varying fastest (the order
    LazySets' `split` produces and `solve(IVP(S, [X0_1..X0_m]))` consumes,
    `src/Continuous/solve.jl:84-101`)."""
    parts = [int(p) for p in parts]
    if len(parts) != len(H.center):
        raise ValueError("one partition count per dimension")
    lo = H.center - H.radius
    rad = H.radius / np.array(parts)
    out = []
    idx = [0] * len(parts)
    total = int(np.prod(parts))
    for _ in range(total):
        c = lo + rad * (2 * np.array(idx) + 1)
        out.append(Hyperrectangle(c, rad.copy()))
        for d in range(len(parts)):
            idx[d] += 1
            if idx[d] < parts[d]:
                break
            idx[d] = 0
    return out
