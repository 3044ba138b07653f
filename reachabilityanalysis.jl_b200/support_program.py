"""Flatten the lazy set trees Ω₀ / V into the ABI's `reach_setexpr`:

    ρ(r_k; Ω₀) = max over alternatives of Σ terms,      ρ(r_k; V) = Σ terms

following the LazySets support-function rules that `ρ(rᵢ, Ω₀)` / `ρ(rᵢ, U)` dispatch to in the
reference loop (LGG09/reach_inhomog.jl:56-57): MinkowskiSum ↦ sum, ConvexHull ↦ max,
LinearMap(M,X) ↦ ρ(Mᵀd, X), Scale(α,X) ↦ ρ(αd, X), CartesianProduct ↦ sum over blocks of d.
A `LinearMap(Φ, ·)` at the top level becomes `which = 1` (evaluate at r_{k+1} = Φᵀ r_k), which is
what lets the device do the Φᵀ r mat-vec once per step instead of twice.

No support function is *evaluated* here -- this only rearranges the description of the sets.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import sets as S
from . import _lib as L


@dataclass
class Term:
    kind: int                 # L.REACH_TERM_BOX / L.REACH_TERM_ZONO
    which: int
    M: np.ndarray = None      # n x q (None: identity)
    c: np.ndarray = None
    r: np.ndarray = None
    G: np.ndarray = None      # dim x p


def _compose(T, alpha, M):
    """New pre-map after descending through LinearMap(M, ·): d_local = (T·M)ᵀ d · alpha."""
    if T is None:
        return np.array(M, dtype=np.float64, copy=True)
    return T @ M


def _leaf(T, alpha, which, kind, c=None, r=None, G=None):
    # fold the scalar into the description: ρ(αd, Box(c,r)) = ρ(d, Box(αc, |α|r));  zonotope alike
    if alpha != 1.0:
        c = None if c is None else alpha * c
        r = None if r is None else abs(alpha) * r
        G = None if G is None else alpha * G
    return Term(kind, which, None if T is None else np.asfortranarray(T), c, r,
                None if G is None else np.asfortranarray(G))


def _flatten(X, T, alpha, which, Phi):
    """Return a list of alternatives, each a list of Terms."""
    if isinstance(X, S.Hyperrectangle):
        return [[_leaf(T, alpha, which, L.REACH_TERM_BOX, X.center.copy(), X.radius.copy())]]
    if isinstance(X, S.BallInf):
        return [[_leaf(T, alpha, which, L.REACH_TERM_BOX, X.center.copy(),
                       np.full(len(X.center), X.radius))]]
    if isinstance(X, S.Interval):
        return [[_leaf(T, alpha, which, L.REACH_TERM_BOX, np.array([(X.lo + X.hi) / 2]),
                       np.array([(X.hi - X.lo) / 2]))]]
    if isinstance(X, S.Singleton):
        return [[_leaf(T, alpha, which, L.REACH_TERM_BOX, X.element.copy(), None)]]
    if isinstance(X, S.ZeroSet):
        return [[]]
    if isinstance(X, S.Zonotope):
        return [[_leaf(T, alpha, which, L.REACH_TERM_ZONO, X.center.copy(), None, X.generators.copy())]]
    if isinstance(X, S.LinearMap):
        if T is None and alpha == 1.0 and which == 0 and Phi is not None and X.M.shape == Phi.shape \
                and np.array_equal(X.M, Phi):
            return _flatten(X.X, None, 1.0, 1, Phi)
        return _flatten(X.X, _compose(T, alpha, X.M), alpha, which, Phi)
    if isinstance(X, S.Scale):
        return _flatten(X.X, T, alpha * X.alpha, which, Phi)
    if isinstance(X, S.MinkowskiSum):
        a, b = _flatten(X.X, T, alpha, which, Phi), _flatten(X.Y, T, alpha, which, Phi)
        return [x + y for x in a for y in b]
    if isinstance(X, S.ConvexHull):
        return _flatten(X.X, T, alpha, which, Phi) + _flatten(X.Y, T, alpha, which, Phi)
    if isinstance(X, S.CartesianProduct):
        nx, ny = S.dim(X.X), S.dim(X.Y)
        n_here = nx + ny
        base = np.eye(n_here) if T is None else T
        a = _flatten(X.X, base[:, :nx], alpha, which, Phi)
        b = _flatten(X.Y, base[:, nx:], alpha, which, Phi)
        return [x + y for x in a for y in b]
    if isinstance(X, S.Universe):
        raise ValueError("the support function of a Universe is unbounded; it cannot be an initial or input set")
    raise NotImplementedError(
        "sets of type %s cannot be flattened for the device (LP-based support functions are out of scope)"
        % type(X).__name__)


def flatten(X, Phi=None):
    """Alternatives (list of lists of Term) for ρ(·, X)."""
    return _flatten(X, None, 1.0, 0, None if Phi is None else np.asarray(Phi))


@dataclass
class CSetExpr:
    """Owns the ctypes structs and the arrays they point into."""
    expr: L.reach_setexpr
    keep: list = field(default_factory=list)

    def ptr(self):
        return C.byref(self.expr)


def to_c(alts, n):
    keep = []
    terms = [t for alt in alts for t in alt]
    arr = (L.reach_term * max(1, len(terms)))()
    for i, t in enumerate(terms):
        q = 0 if t.M is None else t.M.shape[1]
        d = q if q else n
        if t.M is not None and t.M.shape[0] != n:
            raise ValueError("linear map with %d rows in a %d-dimensional problem" % (t.M.shape[0], n))
        for name in ("c", "r"):
            v = getattr(t, name)
            if v is not None and len(v) != d:
                raise ValueError("term %s has length %d, expected %d" % (name, len(v), d))
        if t.G is not None and t.G.shape[0] != d:
            raise ValueError("generator matrix has %d rows, expected %d" % (t.G.shape[0], d))
        arr[i].kind, arr[i].which, arr[i].q = t.kind, t.which, q
        arr[i].p = 0 if t.G is None else t.G.shape[1]
        for name in ("M", "c", "r", "G"):
            v = getattr(t, name)
            if v is not None:
                v = np.asfortranarray(v, dtype=np.float64)
                keep.append(v)
                setattr(arr[i], name, L.dptr(v))
    lens = (C.c_int32 * len(alts))(*[len(a) for a in alts])
    expr = L.reach_setexpr(len(alts), C.cast(lens, L.c_int32_p), C.cast(arr, C.POINTER(L.reach_term)))
    keep += [arr, lens]
    return CSetExpr(expr, keep)


def invariant_to_c(X, n):
    """`X` argument of reach_*_LGG09!: None / Universe -> NULL."""
    if X is None or isinstance(X, S.Universe):
        return None, []
    if isinstance(X, S.HalfSpace):
        a = np.ascontiguousarray(X.a)
        b = np.array([X.b])
        return L.reach_invariant(L.REACH_INV_HALFSPACE, 1, L.dptr(a), L.dptr(b)), [a, b]
    if isinstance(X, S.HPolyhedron):
        A = np.ascontiguousarray(np.stack([h.a for h in X.constraints]))     # row-major m x n
        if A.shape[1] != n:
            raise ValueError("invariant of dimension %d in a %d-dimensional problem" % (A.shape[1], n))
        b = np.array([h.b for h in X.constraints])
        return L.reach_invariant(L.REACH_INV_POLYHEDRON, len(b), L.dptr(A), L.dptr(b)), [A, b]
    if isinstance(X, (S.Hyperrectangle, S.BallInf)):
        r = X.radius if isinstance(X, S.Hyperrectangle) else np.full(n, X.radius)
        lo = np.ascontiguousarray(X.center - r)
        hi = np.ascontiguousarray(X.center + r)
        return L.reach_invariant(L.REACH_INV_BOX, 0, L.dptr(lo), L.dptr(hi)), [lo, hi]
    raise NotImplementedError("invariant of type %s is not supported on the device" % type(X).__name__)
