"""Restatement of the LazySets.jl rules the hot path calls (TEST INFRASTRUCTURE ONLY).

LazySets.jl is a third-party dependency of the reference (Project.toml:34, compat "7") and is
not under /root/reference, so each rule below is the library's published algorithm; the call
sites that reach it are cited.  Sets are walked by duck typing on the class name, so this
module does not import the product package.

  ρ(d, X)                : LGG09/reach_homog.jl:55, reach_inhomog.jl:56-57
  linear_map, minkowski  : Continuous/setops.jl:7-25, GLGM06/reach_homog.jl:24,67
  reduce_order(·, GIR05) : GLGM06/post.jl:26, reach_inhomog.jl:30,36
  convert/overapproximate(Zonotope, ·) : GLGM06/post.jl:25,54 via Discretization/Overapproximate.jl:21-33
"""
import numpy as np


def _name(X):
    return type(X).__name__


def rho(d, X):
    """Support function ρ(d, X).  `d` is a vector (n,) or a block of directions (n, m); returns a
    scalar or (m,)."""
    d = np.asarray(d, dtype=np.float64)
    single = d.ndim == 1
    D = d.reshape(len(d), -1)
    out = _rho(D, X)
    return float(out[0]) if single else out


def _rho(D, X):
    k = _name(X)
    if k == "Hyperrectangle":
        # res += d_i * (c_i ± r_i) by the sign of d_i (d_i == 0 contributes nothing)
        return np.sum(D * (X.center[:, None] + np.sign(D) * X.radius[:, None]), axis=0)
    if k == "BallInf":
        return np.sum(D * (X.center[:, None] + np.sign(D) * X.radius), axis=0)
    if k == "Interval":
        d0 = D[0]
        return d0 * np.where(d0 > 0, X.hi, X.lo)
    if k == "Singleton":
        return X.element @ D
    if k == "ZeroSet":
        return np.zeros(D.shape[1])
    if k == "Universe":
        return np.where(np.any(D != 0, axis=0), np.inf, 0.0)
    if k == "Zonotope":
        # ρ = d·c + Σ_j |g_j·d|
        return X.center @ D + np.sum(np.abs(X.generators.T @ D), axis=0)
    if k == "LinearMap":
        return _rho(X.M.T @ D, X.X)
    if k == "Scale":
        return _rho(X.alpha * D, X.X)
    if k == "MinkowskiSum":
        return _rho(D, X.X) + _rho(D, X.Y)
    if k == "ConvexHull":
        return np.maximum(_rho(D, X.X), _rho(D, X.Y))
    if k == "CartesianProduct":
        nx = _dim(X.X)
        return _rho(D[:nx], X.X) + _rho(D[nx:], X.Y)
    raise TypeError("ρ not defined for %s" % k)


def _dim(X):
    k = _name(X)
    if k in ("Hyperrectangle", "BallInf", "Zonotope"):
        return len(X.center)
    if k == "Interval":
        return 1
    if k == "Singleton":
        return len(X.element)
    if k in ("ZeroSet", "Universe"):
        return X.n
    if k == "LinearMap":
        return X.M.shape[0]
    if k == "Scale":
        return _dim(X.X)
    if k in ("MinkowskiSum", "ConvexHull"):
        return _dim(X.X)
    if k == "CartesianProduct":
        return _dim(X.X) + _dim(X.Y)
    raise TypeError(k)


# ---------------------------------------------------------------------------------------------
# zonotopes: (c, G) pairs of plain arrays
# ---------------------------------------------------------------------------------------------

def to_zonotope(X):
    """`_convert_or_overapproximate(Zonotope, X)` for the zonotopic sets the discretizations
    produce (GLGM06/post.jl:25,54).  A hyperrectangle contributes one generator r_i·e_i per
    NON-ZERO radius, in index order (pinned by the generator counts in
    test/algorithms/GLGM06.jl:61 and test/continuous/solve.jl:15-16)."""
    k = _name(X)
    if k == "Zonotope":
        return X.center.copy(), X.generators.copy()
    if k in ("Hyperrectangle", "BallInf"):
        c = X.center
        n = len(c)
        r = X.radius if k == "Hyperrectangle" else np.full(n, X.radius)
        nz = np.nonzero(r)[0]
        G = np.zeros((n, len(nz)))
        G[nz, np.arange(len(nz))] = r[nz]
        return c.copy(), G
    if k == "Interval":
        r = (X.hi - X.lo) / 2
        G = np.array([[r]]) if r != 0 else np.zeros((1, 0))
        return np.array([(X.lo + X.hi) / 2]), G
    if k == "Singleton":
        return X.element.copy(), np.zeros((len(X.element), 0))
    if k == "ZeroSet":
        return np.zeros(X.n), np.zeros((X.n, 0))
    if k == "LinearMap":
        c, G = to_zonotope(X.X)
        return X.M @ c, X.M @ G
    if k == "Scale":
        c, G = to_zonotope(X.X)
        return X.alpha * c, X.alpha * G
    if k == "MinkowskiSum":
        c1, G1 = to_zonotope(X.X)
        c2, G2 = to_zonotope(X.Y)
        return c1 + c2, np.hstack([G1, G2])
    if k == "ConvexHull":
        return ch_zonotope(to_zonotope(X.X), to_zonotope(X.Y))
    raise TypeError("cannot convert %s to a zonotope" % k)


def ch_zonotope(Z1, Z2):
    """`overapproximate(ConvexHull(Z1, Z2), Zonotope)` (Girard 2005 "mean" rule) --
    FirstOrderZonotopeModule.jl:137."""
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


def gir05_weights(G):
    """w_j = ‖g_j‖₁ − ‖g_j‖∞ (LazySets `_weighted_gens!(…, ::GIR05)`)."""
    a = np.abs(G)
    return a.sum(axis=0) - (a.max(axis=0) if G.shape[0] else 0.0)


def reduce_order_gir05(c, G, r, return_selection=False):
    """`reduce_order(Zonotope(c, G), r, GIR05())`.

    r·n ≥ p: unchanged.  r == 1: box everything, unsorted.  Otherwise the m = p − ⌊n(r−1)⌋
    generators of smallest weight (stable ascending sort -- Julia's `sortperm` guarantees that
    equal weights keep ascending index order) are replaced by their interval hull; the result is
    [kept generators in sorted order | n×n diagonal]."""
    n, p = G.shape
    if r < 1:
        raise ValueError("the target order should be at least 1, but it is %s" % (r,))
    if r * n >= p:
        return (c, G, None) if return_selection else (c, G)
    if r == 1:
        D = np.diag(np.abs(G).sum(axis=1))
        return (c, D, np.zeros(0, dtype=np.int64)) if return_selection else (c, D)
    w = gir05_weights(G)
    idx = np.argsort(w, kind="stable")
    m = p - int(np.floor(n * (r - 1)))
    box = idx[:m]
    keep = idx[m:]
    diag = np.zeros(n)
    for j in box:  # summed in sorted order, as `_interval_hull(G, view(indices, 1:m))` does
        diag += np.abs(G[:, j])
    Gred = np.hstack([G[:, keep], np.diag(diag)])
    return (c, Gred, keep) if return_selection else (c, Gred)


def gir05_boundary_gap(G, r):
    """Relative weight gap across the keep/box cut (None if no reduction happens).  A gap near
    rounding noise means 'identical generator selection' is not well defined for this input."""
    n, p = G.shape
    if r * n >= p or r == 1:
        return None
    w = np.sort(gir05_weights(G), kind="stable")
    m = p - int(np.floor(n * (r - 1)))
    scale = max(abs(w[m]), abs(w[m - 1]), 1e-300)
    return (w[m] - w[m - 1]) / scale


def linear_map_minkowski_sum(M, Z1, Z2):
    """Continuous/setops.jl:7-25: c = M c₁ + c₂ ; G = [M G₁ | G₂]."""
    (c1, G1), (c2, G2) = Z1, Z2
    return M @ c1 + c2, np.hstack([M @ G1, G2])


def rho_zonotope(d, c, G):
    return float(d @ c + np.abs(G.T @ d).sum())


def sigma_zonotope(d, c, G):
    """Support vector of a zonotope: c + Σ_j sign(g_j·d) g_j."""
    return c + G @ np.sign(G.T @ d)
