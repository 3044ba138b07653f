"""LGG09 support-function recurrence, restated from the reference (TEST INFRASTRUCTURE ONLY).

Follows src/Algorithms/LGG09/reach_homog.jl:49-62 and reach_inhomog.jl:49-64 (ZeroSet
specialisation :66-81): per direction ℓ, r₀ = ℓ, and for i = 1..NSTEPS

    ρℓ[j, i] = ρ(r_i, Ω₀) + s_i ;  s_{i+1} = s_i + ρ(r_i, U) ;  r_{i+1} = Φᵀ r_i

in exactly that (sequential) order.  The loop over directions is vectorised as a block of
columns -- which is the reference's own invariant-variant shape (`mul!(rᵢ₊₁, Φᵀ, rᵢ)` with one
column per direction, reach_inhomog.jl:174-218) -- so one GEMM per step replaces ndirs GEMVs.
"""
import numpy as np

from .lazysets_oracle import rho, _name


def box_directions(n):
    """`BoxDirections(n)`: +e₁..+e_n then −e₁..−e_n (LGG09.jl:112; order pinned for the `vars`
    form by test/algorithms/LGG09.jl:52-55)."""
    I = np.eye(n)
    return np.hstack([I, -I])


def vars_directions(vars_, n):
    """`_get_template(vars, n)`: [+e_v for v in vars] then [−e_v …] (LGG09.jl:117-131); 1-based."""
    vars_ = [vars_] if np.isscalar(vars_) else list(vars_)
    D = np.zeros((n, 2 * len(vars_)))
    for j, v in enumerate(vars_):
        D[v - 1, j] = 1.0
        D[v - 1, len(vars_) + j] = -1.0
    return D


def reach_homog_LGG09(dirs, Omega0, Phi, nsteps):
    """reach_homog.jl:6-33,49-62.  dirs: (n, ndirs), one column per direction.
    Returns ρℓ (ndirs, nsteps)."""
    PhiT = np.ascontiguousarray(Phi.T)          # :18
    R = np.array(dirs, dtype=np.float64, copy=True)
    out = np.empty((R.shape[1], nsteps))
    for i in range(nsteps):
        out[:, i] = rho(R, Omega0)              # :55
        R = PhiT @ R                            # :58-59
    return out


def reach_inhomog_LGG09(dirs, Omega0, Phi, nsteps, U):
    """reach_inhomog.jl:5-33,49-64.  Returns ρℓ (ndirs, nsteps)."""
    PhiT = np.ascontiguousarray(Phi.T)
    R = np.array(dirs, dtype=np.float64, copy=True)
    m = R.shape[1]
    out = np.empty((m, nsteps))
    s = np.zeros(m)
    zero_init = _name(Omega0) == "ZeroSet"      # :66-81
    for i in range(nsteps):
        out[:, i] = s if zero_init else rho(R, Omega0) + s   # :56
        s = s + rho(R, U)                       # :57
        R = PhiT @ R                            # :60-61
    return out


def reach_LGG09_invariant(dirs, Omega0, Phi, nsteps, X, U=None, exact=False):
    """Invariant variants (reach_homog.jl:382-424, reach_inhomog.jl:174-218) for the invariants
    this build supports on the device: a HalfSpace a·x ≤ b or a Hyperrectangle.  The template
    reach-set F[k] = {x : dirs_j·x ≤ ρℓ[j,k]} is declared disjoint from X when a template
    direction certifies it -- the test `_isdisjoint(X, set(F[k]))` resolves to for these X when
    ±a (resp. ±e_i) are template directions:
       HalfSpace:      some direction j == −a/‖…‖ scaled:  −ρ(−a, F) > b
       Hyperrectangle: some i with  −ρ(−e_i, F) > hi_i  or  ρ(e_i, F) < lo_i.
    Returns (ρℓ[:, :k_stop], k_stop): the step that triggered the break is computed but dropped
    (`resize!(F, k - 1)`, reach_homog.jl:421-423).  exact=True uses the reference's own exact polyhedral test
    (template_disjoint_lp, one LP per step) instead of the template-row certificates."""
    PhiT = np.ascontiguousarray(Phi.T)
    R = np.array(dirs, dtype=np.float64, copy=True)
    m = R.shape[1]
    out = np.empty((m, nsteps))
    s = np.zeros(m)
    k = 0
    while k < nsteps:
        out[:, k] = rho(R, Omega0) + s
        if U is not None:
            s = s + rho(R, U)
        R = PhiT @ R
        if (template_disjoint_lp if exact else template_disjoint)(dirs, out[:, k], X):
            break
        k += 1
    return out[:, :k], k


def _faces(X):
    k = _name(X)
    if k == "HalfSpace":
        return [(np.asarray(X.a, dtype=float), float(X.b))]
    if k == "HPolyhedron":
        return [(np.asarray(h.a, dtype=float), float(h.b)) for h in X.constraints]
    if k in ("Hyperrectangle", "BallInf"):
        n = len(X.center)
        r = X.radius if k == "Hyperrectangle" else np.full(n, X.radius)
        out = []
        for i in range(n):
            e = np.zeros(n)
            e[i] = 1.0
            out.append((e, X.center[i] + r[i]))
            out.append((-e, -(X.center[i] - r[i])))
        return out
    raise TypeError("invariant %s not supported" % k)


def template_disjoint_lp(dirs, sf, X):
    """The reference's test itself: `_isdisjoint(X, set(F[k]))` (reach_inhomog.jl:210) with the default method is the
    exact emptiness of {x : dirs_j.x <= sf_j} ∩ X, which LazySets decides with a linear program (for a single half-space:
    min a.x over the polytope > b).  Here: one LP feasibility problem (SciPy / HiGHS)."""
    if _name(X) == "Universe":
        return False
    from scipy.optimize import linprog
    faces = _faces(X)
    A = np.vstack([np.asarray(dirs, dtype=float).T] + [a[None, :] for a, _ in faces])
    b = np.concatenate([np.asarray(sf, dtype=float), [bb for _, bb in faces]])
    keep = np.isfinite(b)
    res = linprog(np.zeros(A.shape[1]), A_ub=A[keep], b_ub=b[keep], bounds=[(None, None)] * A.shape[1], method="highs")
    return res.status == 2            # infeasible: the sets are disjoint


def template_disjoint(dirs, sf, X):
    k = _name(X)
    if k == "Universe":
        return False
    if k == "HalfSpace":
        # need a template direction d = −λ a (λ>0): max over F of (−a·x)·λ = sf  ⇒  min a·x = −sf/λ
        for j in range(dirs.shape[1]):
            d = dirs[:, j]
            lam = _neg_multiple(d, X.a)
            if lam is not None and -sf[j] / lam > X.b:
                return True
        return False
    if k == "Hyperrectangle":
        lo, hi = X.center - X.radius, X.center + X.radius
        for j in range(dirs.shape[1]):
            d = dirs[:, j]
            nz = np.nonzero(d)[0]
            if len(nz) != 1:
                continue
            i = nz[0]
            if d[i] > 0 and sf[j] / d[i] < lo[i]:
                return True
            if d[i] < 0 and -sf[j] / (-d[i]) > hi[i]:
                return True
        return False
    raise TypeError("invariant %s not supported" % k)


def _neg_multiple(d, a):
    """λ > 0 with d == −λ a exactly, else None."""
    nz = np.nonzero(a)[0]
    if len(nz) == 0:
        return None
    lam = -d[nz[0]] / a[nz[0]]
    if lam > 0 and np.array_equal(d, -lam * a):
        return lam
    return None
