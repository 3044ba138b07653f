"""TEST INFRASTRUCTURE -- CPU restatement (NumPy float64, sequential loops) of the BOX algorithm of
ReachabilityAnalysis.jl.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

Follows src/Algorithms/BOX/reach_homog.jl:8-47 (non-recursive), :50-69 (recursive), :72-118 (invariant),
reach_inhomog.jl:8-60 (non-recursive), :63-118 (invariant).  Parity unpinned by the reference's own tests beyond
types and set inclusion (test/algorithms/BOX.jl); the arithmetic restated here is the pin.
"""
import numpy as np


def _disjoint_box(X, c, r):
    """`_isdisjoint(X, Hyperrectangle(c, r))` for the invariants the device path knows: (a, b) half-space a.x <= b,
    or (lo, hi) box."""
    if X is None:
        return False
    kind, p, q = X
    if kind == "halfspace":           # min over the box of a.x = a.c - |a|.r  >  b
        return float(p @ c - np.abs(p) @ r) > float(q)
    if kind == "box":
        return bool(np.any(c - r > q) or np.any(c + r < p))
    raise ValueError(kind)


def reach_homog_BOX(c0, r0, Phi, NSTEPS, recursive=False, X=None):
    c0, r0, Phi = np.asarray(c0, float), np.asarray(r0, float), np.asarray(Phi, float)
    C, R = [c0.copy()], [r0.copy()]
    if recursive:                                     # reach_homog.jl:50-69
        absPhi = np.abs(Phi)
        for k in range(1, NSTEPS):
            C.append(Phi @ C[-1])
            R.append(absPhi @ R[-1])
        return np.array(C).T, np.array(R).T
    P = Phi.copy()                                    # Phi_power_k
    for k in range(1, NSTEPS):
        c = P @ c0                                    # mul!(c[k], Phi_power_k, c[1])
        r = np.abs(P) @ r0
        if _disjoint_box(X, c, r):
            break
        C.append(c)
        R.append(r)
        P = P @ Phi                                   # mul!(cache, Phi_power_k, Phi)
    return np.array(C).T, np.array(R).T


def reach_inhomog_BOX(c0, r0, Phi, NSTEPS, cU, rU, X=None):
    c0, r0, Phi = np.asarray(c0, float), np.asarray(r0, float), np.asarray(Phi, float)
    cU, rU = np.asarray(cU, float), np.asarray(rU, float)
    C, R = [c0.copy()], [r0.copy()]
    Wc, Wr = cU.copy(), rU.copy()                     # Wk+ = copy(U)
    P = Phi.copy()
    for k in range(1, NSTEPS):
        aP = np.abs(P)
        c = P @ c0 + Wc
        r = aP @ r0 + Wr
        if _disjoint_box(X, c, r):
            break
        C.append(c)
        R.append(r)
        Wc = Wc + P @ cU
        Wr = Wr + aP @ rU
        P = P @ Phi
    return np.array(C).T, np.array(R).T
