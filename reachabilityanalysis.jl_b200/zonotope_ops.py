"""Host-side zonotope conversions used *before* the hot loop (GLGM06/post.jl:25,54:
`_convert_or_overapproximate(Zonotope, ·)`, Discretization/Overapproximate.jl:21-33, and the
Girard-2005 convex-hull rule FirstOrderZonotopeModule.jl:137 relies on).  These run once per
problem on the host, exactly as they do in the reference; order reduction and the propagation
loop are on the device."""
import numpy as np

from . import sets as S


def to_zonotope(X):
    """Return (c, G) of the zonotope equal to (or, for ConvexHull, covering) X.  A
    hyperrectangle yields one generator r_i·e_i per non-zero radius, in index order."""
    if isinstance(X, S.Zonotope):
        return X.center.copy(), X.generators.copy()
    if isinstance(X, (S.Hyperrectangle, S.BallInf)):
        c = X.center
        n = len(c)
        r = X.radius if isinstance(X, S.Hyperrectangle) else np.full(n, X.radius)
        nz = np.flatnonzero(r)
        G = np.zeros((n, len(nz)))
        G[nz, np.arange(len(nz))] = r[nz]
        return c.copy(), G
    if isinstance(X, S.Interval):
        r = (X.hi - X.lo) / 2
        return np.array([(X.lo + X.hi) / 2]), (np.array([[r]]) if r != 0 else np.zeros((1, 0)))
    if isinstance(X, S.Singleton):
        return X.element.copy(), np.zeros((len(X.element), 0))
    if isinstance(X, S.ZeroSet):
        return np.zeros(X.n), np.zeros((X.n, 0))
    if isinstance(X, S.LinearMap):
        c, G = to_zonotope(X.X)
        return X.M @ c, X.M @ G
    if isinstance(X, S.Scale):
        c, G = to_zonotope(X.X)
        return X.alpha * c, X.alpha * G
    if isinstance(X, S.MinkowskiSum):
        c1, G1 = to_zonotope(X.X)
        c2, G2 = to_zonotope(X.Y)
        return c1 + c2, np.hstack([G1, G2])
    if isinstance(X, S.ConvexHull):
        return convex_hull_zonotope(to_zonotope(X.X), to_zonotope(X.Y))
    raise TypeError("cannot represent %s as a zonotope" % type(X).__name__)


def convex_hull_zonotope(Z1, Z2):
    """Zonotope cover of CH(Z1 ∪ Z2): c = (c₁+c₂)/2, G = [(G₁+G₂)/2 | (c₁−c₂)/2 | (G₁−G₂)/2]
    (plus the surplus generators of the higher-order operand)."""
    (c1, G1), (c2, G2) = Z1, Z2
    if G2.shape[1] > G1.shape[1]:
        (c1, G1), (c2, G2) = (c2, G2), (c1, G1)
    p2 = G2.shape[1]
    c = (c1 + c2) / 2
    if G1.shape[1] == p2:
        G = np.hstack([G1 + G2, (c1 - c2)[:, None], G1 - G2]) / 2
    else:
        G = np.hstack([(G1[:, :p2] + G2) / 2, ((c1 - c2) / 2)[:, None], (G1[:, :p2] - G2) / 2,
                       G1[:, p2:]])
    return c, G


def interval_hull_of_map(M, X):
    """(centre, radius) of the interval hull of M·X for the box-like / zonotopic X used here."""
    if isinstance(X, S.LinearMap):
        return interval_hull_of_map(M @ X.M, X.X)
    if isinstance(X, S.Scale):
        return interval_hull_of_map(M * X.alpha, X.X)
    if isinstance(X, S.MinkowskiSum):
        c1, r1 = interval_hull_of_map(M, X.X)
        c2, r2 = interval_hull_of_map(M, X.Y)
        return c1 + c2, r1 + r2
    c, G = to_zonotope(X)
    return M @ c, np.abs(M @ G).sum(axis=1)


def symmetric_interval_hull_of_map(M, X):
    """`symmetric_interval_hull(M * X)` = Hyperrectangle(0, |M c| + |M| r)."""
    c, r = interval_hull_of_map(M, X)
    return S.Hyperrectangle(np.zeros(len(c)), np.abs(c) + r)


def norm_inf(X):
    """`norm(X, Inf)`: largest absolute vertex coordinate."""
    c, r = interval_hull_of_map(np.eye(S.dim(X)), X)
    return float(np.max(np.abs(c) + r)) if len(c) else 0.0
