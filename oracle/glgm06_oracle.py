"""GLGM06 zonotope flowpipe, restated from the reference (TEST INFRASTRUCTURE ONLY).

Follows src/Algorithms/GLGM06/post.jl:25-30,54-57, reach_homog.jl:33-73 (homogeneous: Φ applied
repeatedly to the *set*, no order reduction in the loop) and reach_inhomog.jl:6-43
(inhomogeneous: the *power* P_k = P_{k-1}·Φ applied to Ω₀, two GIR05 reductions per step),
with Continuous/setops.jl:7-25 for `linear_map_minkowski_sum`.
"""
import numpy as np

from .lazysets_oracle import (to_zonotope, reduce_order_gir05, linear_map_minkowski_sum,
                              gir05_boundary_gap)


def post_GLGM06(Phi, Omega0, nsteps, max_order=5, U=None, return_diag=False):
    """GLGM06/post.jl:4-61 for a discrete IVP (Φ, Ω₀[, U]).  Returns a list of (c, G) per step
    (F[1] is Ω₀ itself)."""
    c0, G0 = to_zonotope(Omega0)                       # :25
    c0, G0 = reduce_order_gir05(c0, G0, max_order)     # :26
    if U is None:
        return reach_homog_GLGM06(c0, G0, Phi, nsteps)
    cU, GU = to_zonotope(U)                            # :54
    return reach_inhomog_GLGM06(c0, G0, Phi, nsteps, max_order, cU, GU, return_diag=return_diag)


def reach_homog_GLGM06(c0, G0, Phi, nsteps):
    """reach_homog.jl:33-73 (preallocate=true, the default): Z_k = Φ·Z_{k-1}.  p == 0 becomes
    one zero generator (:52-54)."""
    n = len(c0)
    if G0.shape[1] == 0:
        G0 = np.zeros((n, 1))
    out = [(c0.copy(), G0.copy())]
    c, G = c0, G0
    for _ in range(1, nsteps):
        c = Phi @ c                                    # linear_map!(Zout[k], Φ, Zout[k-1]) :67
        G = Phi @ G
        out.append((c, G))
    return out


def reach_inhomog_GLGM06(c0, G0, Phi, nsteps, max_order, cU, GU, return_diag=False):
    """reach_inhomog.jl:6-43."""
    out = [(c0.copy(), G0.copy())]                     # :20
    W = (cU, GU)                                       # :22
    P = Phi.copy()                                     # :23-24
    gaps = []
    sels = [None]
    for _ in range(1, nsteps):
        cR, GR = linear_map_minkowski_sum(P, (c0, G0), W)          # :29
        if return_diag:
            gaps.append(gir05_boundary_gap(GR, max_order))
            cR, GR, keep = reduce_order_gir05(cR, GR, max_order, return_selection=True)
            sels.append(keep)
        else:
            cR, GR = reduce_order_gir05(cR, GR, max_order)         # :30
        out.append((cR, GR))                                       # :33
        cW, GW = linear_map_minkowski_sum(P, (cU, GU), W)          # :35
        W = reduce_order_gir05(cW, GW, max_order)                  # :36
        P = P @ Phi                                                # :38
    if return_diag:
        return out, sels, gaps
    return out


def first_disjoint_step(F, X):
    """Early termination of the invariant variants (reach_homog.jl:76-147, reach_inhomog.jl:46-86):
    `_isdisjoint(X, R_k) && break` is evaluated for k >= 2 (F[1] = Omega0 is stored unchecked) and
    the disjoint set is not stored.  Returns the number of sets kept.  For a half-space
    {a.x <= b} and a zonotope R the test is exact: disjoint iff  -rho(-a, R) > b."""
    a, b = X.a, X.b
    for k in range(1, len(F)):
        c, G = F[k]
        if -(float(-a @ c) + np.abs(G.T @ (-a)).sum()) > b:
            return k
    return len(F)
