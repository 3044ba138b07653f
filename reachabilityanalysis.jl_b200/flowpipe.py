"""Result containers: TemplateReachSet, ReachSet, Flowpipe, MixedFlowpipe.

Mirrors src/ReachSets/TemplateReachSet.jl:32-37,97-115, src/ReachSets/ReachSet.jl:24-27,
src/Flowpipes/Flowpipe.jl:22-25,97-128,304-331, AbstractFlowpipe.jl:35-37 and
MixedFlowpipe.jl:20-23,66-68 -- with the difference this build is about: the support-value
matrix / the zonotope generators live in device memory (the plan object) and are only brought
to the host when asked for; `rho(d, fp)` over a template direction is a device reduction.
Time stamps stay on the host exactly as in the reference (`Δt += δ`, reach_homog.jl:26-30).
"""
import math

import numpy as np


def _add_rounded(a, b, up):
    """a + b rounded towards +inf (`up`) or -inf: the rounded-to-nearest sum moved one ulp outwards when the exact
    error of the addition (Knuth's TwoSum) has that sign.  Works on floats and on arrays."""
    s = a + b
    bb = s - a
    err = (a - (s - bb)) + (b - bb)
    if isinstance(s, np.ndarray):
        return np.where(err > 0, np.nextafter(s, np.inf), s) if up else np.where(err < 0, np.nextafter(s, -np.inf), s)
    if up:
        return math.nextafter(s, math.inf) if err > 0 else s
    return math.nextafter(s, -math.inf) if err < 0 else s


class TimeInterval(tuple):
    """`TimeInterval` (an IntervalArithmetic interval in the reference, TimeInterval.jl:54-56): sums round outwards,
    lower bound down and upper bound up, so that repeated `Δt += δ` never loses a time point -- `F(1.0) == F[end]`
    holds for δ = 0.1 although ten rounded-to-nearest additions of 0.1 stop below 1 (test/continuous/solve.jl:66)."""

    def __new__(cls, lo, hi):
        return super().__new__(cls, (float(lo), float(hi)))

    lo = property(lambda s: s[0])
    hi = property(lambda s: s[1])

    def __add__(self, other):
        if isinstance(other, TimeInterval):
            return TimeInterval(_add_rounded(self[0], other[0], False), _add_rounded(self[1], other[1], True))
        other = float(other)
        return TimeInterval(_add_rounded(self[0], other, False), _add_rounded(self[1], other, True))

    def __contains__(self, t):
        return self[0] <= t <= self[1]


zeroT = TimeInterval(0.0, 0.0)


class TemplateReachSet:
    """(dirs, sf, Δt): sf is a view into column k of the ndirs × NSTEPS matrix."""

    def __init__(self, dirs, sf, dt):
        self.dirs, self.sf, self.dt = dirs, sf, dt

    tspan = property(lambda s: s.dt)
    tstart = property(lambda s: s.dt.lo)
    tend = property(lambda s: s.dt.hi)
    dim = property(lambda s: s.dirs.n)
    vars = property(lambda s: tuple(range(1, s.dirs.n + 1)))          # TemplateReachSet.jl:71, 1-based like the reference
    directions = property(lambda s: s.dirs)                          # :73

    @property
    def set(self):
        """`set(R)` (TemplateReachSet.jl:57-60): the polyhedron {x : d_i·x ≤ sf_i ∀i} as a list of half-spaces."""
        from .sets import HPolyhedron, HalfSpace
        return HPolyhedron([HalfSpace(d, float(self.sf[i])) for i, d in enumerate(self.dirs)])

    def support_functions(self, i=None):
        return self.sf if i is None else self.sf[i]

    def __eq__(self, other):
        return (isinstance(other, TemplateReachSet) and tuple(self.dt) == tuple(other.dt) and self.dirs == other.dirs
                and np.array_equal(np.asarray(self.sf), np.asarray(other.sf)))

    __hash__ = None

    def rho(self, d):
        """TemplateReachSet.jl:97-115: exact match, then positive multiple; no LP fallback."""
        d = np.asarray(d, dtype=np.float64)
        D = self.dirs.matrix()
        for i in range(D.shape[1]):
            if np.array_equal(d, D[:, i]):
                return float(self.sf[i])
        for i in range(D.shape[1]):
            ok, k = samedir(d, D[:, i])
            if ok:
                return float(self.sf[i]) * k
        # TemplateReachSet.jl:112-114: `ρ(d, set(R))`, the support function of the polyhedron the template describes --
        # a linear program (LazySets' ρ(d, ::HPolyhedron)); host side, as in the reference
        return lp_support(d, D, np.asarray(self.sf, dtype=np.float64))

    def sigma(self, d):
        """σ(d, R) = σ(d, set(R)): a maximiser of d·x over the template polyhedron (LazySets' σ(d, ::HPolyhedron) is
        the same linear program as its ρ); host side, like the LP fallback of `rho`."""
        return lp_support(d, self.dirs.matrix(), np.asarray(self.sf, dtype=np.float64), argmax=True)


def lp_support(d, D, sf, argmax=False):
    """ρ(d, {x : D[:, j]·x ≤ sf[j] ∀j}) = max d·x by linear programming (what LazySets' `ρ(d, P::HPolyhedron)` solves).
    +inf when the polyhedron is unbounded in direction d; raises on an empty polyhedron, like LazySets.
    `argmax=True` returns the maximiser instead (raises when unbounded)."""
    from scipy.optimize import linprog
    res = linprog(-np.asarray(d, dtype=np.float64), A_ub=np.ascontiguousarray(D.T), b_ub=sf,
                  bounds=[(None, None)] * D.shape[0], method="highs")
    if res.status == 3:
        if argmax:
            raise ValueError("the support vector in direction %s is undefined because the polyhedron is unbounded" % (d,))
        return float("inf")
    if res.status == 2:
        raise ValueError("the support function in direction(s) %s is undefined because the polyhedron is empty" % (d,))
    if res.status != 0:
        raise RuntimeError("LP solver: %s" % res.message)
    return np.array(res.x) if argmax else float(-res.fun)


def _same_value(a, b):
    """Value equality of sets / arrays / plain fields (the host set types are frozen records holding arrays)."""
    if type(a) is not type(b):
        return False
    if isinstance(a, np.ndarray):
        return a.shape == b.shape and bool(np.array_equal(a, b))
    if hasattr(a, "__dataclass_fields__"):
        return all(_same_value(getattr(a, f), getattr(b, f)) for f in a.__dataclass_fields__)
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_same_value(x, y) for x, y in zip(a, b))
    return a == b


class FlowpipeSlice(list):
    """`F[i:j]`, `F(Δt)`: a run of reach-sets with the time-span interface of the reference's views
    (`tstart(F[1:3])`, `tspan(F(0.05 .. 1.05))`, test/continuous/solve.jl:43-75)."""

    tstart = property(lambda s: min(R.tstart for R in s))
    tend = property(lambda s: max(R.tend for R in s))
    tspan = property(lambda s: TimeInterval(s.tstart, s.tend))


def samedir(u, v):
    """(True, k) iff u = k·v with k > 0."""
    nz = np.flatnonzero(v)
    if len(nz) == 0:
        return False, 0.0
    k = u[nz[0]] / v[nz[0]]
    if k > 0 and np.allclose(u, k * v, rtol=1e-12, atol=0.0):
        return True, float(k)
    return False, 0.0


class ReachSet:
    """(set, Δt) with set a Zonotope (GLGM06), a Hyperrectangle (BOX) or an Interval (`to_intervals`)."""

    def __init__(self, Z, dt):
        self.set, self.dt = Z, dt

    tspan = property(lambda s: s.dt)
    tstart = property(lambda s: s.dt.lo)
    tend = property(lambda s: s.dt.hi)

    @property
    def dim(self):
        from .sets import dim
        return dim(self.set)

    @property
    def vars(self):
        return tuple(range(1, self.dim + 1))

    def __eq__(self, other):
        """Same set (by value) over the same time span, like `==` of the reference's immutable structs."""
        return isinstance(other, ReachSet) and tuple(self.dt) == tuple(other.dt) and _same_value(self.set, other.set)

    __hash__ = None

    def rho(self, d):
        d = np.asarray(d, dtype=np.float64)
        if hasattr(self.set, "generators"):
            c, G = self.set.center, self.set.generators
            return float(d @ c + np.abs(G.T @ d).sum())
        if hasattr(self.set, "lo") and hasattr(self.set, "hi") and not hasattr(self.set, "radius"):     # Interval
            return float(d[0] * (self.set.hi if d[0] > 0 else self.set.lo))
        return float(d @ self.set.center + np.abs(d) @ self.set.radius)      # Hyperrectangle (BOX)

    def sigma(self, d):
        """Support vector of the set: c + Σ_j sign(g_j·d) g_j for a zonotope, the vertex c + sign(d)·r for a box."""
        d = np.asarray(d, dtype=np.float64)
        if hasattr(self.set, "generators"):
            c, G = self.set.center, self.set.generators
            return c + G @ np.sign(G.T @ d)
        if hasattr(self.set, "lo") and hasattr(self.set, "hi") and not hasattr(self.set, "radius"):     # Interval
            return np.array([self.set.hi if d[0] > 0 else self.set.lo])
        return self.set.center + np.sign(d) * self.set.radius


class Flowpipe:
    """Indexable sequence of reach-sets + `ext` dictionary (Flowpipe.jl:22-25).

    `make(k)` materialises reach-set k on demand; `nsets` is their number."""

    def __init__(self, nsets, make, delta, dt0=zeroT, ext=None, device=None):
        self._n = int(nsets)
        self._make = make
        self.delta = float(delta)
        self.dt0 = dt0
        self.ext = ext if ext is not None else {}
        self.device = device            # the plan wrapper holding the device-resident data
        # Δt_k by repeated addition, as the reference does
        self._t = None

    def _times(self):
        if self._t is None:
            lo = np.empty(self._n)
            hi = np.empty(self._n)
            dt = TimeInterval(0.0, self.delta) + self.dt0
            for k in range(self._n):
                lo[k], hi[k] = dt
                dt = dt + self.delta
            self._t = (lo, hi)
        return self._t

    def shift(self, t0):
        """`shift(fp, t0)` (Flowpipe.jl:153-169, ReachSet.jl:52-54): every reach-set keeps its set and gets the time
        span `tspan(R) + t0`; the device-resident data and `ext` are shared, nothing is copied or recomputed."""
        lo, hi = self._times()
        out = Flowpipe(self._n, self._make, self.delta, self.dt0, self.ext, self.device)
        out._t = (_add_rounded(lo, float(t0), False), _add_rounded(hi, float(t0), True))
        return out

    def __len__(self):
        return self._n

    def __getitem__(self, k):
        if isinstance(k, slice):
            return FlowpipeSlice(self[i] for i in range(*k.indices(self._n)))
        if k < 0:
            k += self._n
        if not 0 <= k < self._n:
            raise IndexError(k)
        lo, hi = self._times()
        return self._make(k, TimeInterval(lo[k], hi[k]))

    def __iter__(self):
        return (self[k] for k in range(self._n))

    def numrsets(self):
        """Flowpipe.jl: the number of reach-sets (== len(fp), test/flowpipes/flowpipes.jl:24)."""
        return self._n

    def set_of(self, k=None):
        """`set(fp, k)` (AbstractFlowpipe.jl:108-140): the set of reach-set k; a list of sets for a slice / index list;
        every set for `set(fp)` (the reference wraps those in a UnionSetArray)."""
        if k is None:
            return [R.set for R in self]
        if isinstance(k, slice):
            return [R.set for R in self[k]]
        if isinstance(k, (list, tuple, range, np.ndarray)):
            return [self[int(i)].set for i in k]
        return self[k].set

    @property
    def vars(self):
        """AbstractFlowpipe.jl:211: the variables of the first reach-set (1-based, like the reference)."""
        return self[0].vars

    @property
    def dim(self):
        """AbstractFlowpipe.jl:71-75: the dimension of the first reach-set; undefined for an empty flowpipe."""
        if self._n == 0:
            raise ValueError("the dimension is not defined because this flowpipe is empty")
        return self[0].dim

    @property
    def tstart(self):
        return self._times()[0][0]

    @property
    def tend(self):
        return self._times()[1][-1]

    @property
    def tspan(self):
        return TimeInterval(self.tstart, self.tend)

    def __call__(self, t):
        """Flowpipe.jl:97-128: the reach-set(s) whose time span contains t."""
        lo, hi = self._times()
        if np.isscalar(t):
            idx = np.flatnonzero((lo <= t) & (t <= hi))
            if len(idx) == 0:
                raise ValueError("time %s does not belong to the time span %s" % (t, tuple(self.tspan)))
            sets = FlowpipeSlice(self[i] for i in idx[:2])      # the first occurrence and, on a border, its successor
            return sets[0] if len(sets) == 1 else sets
        if hasattr(t, "lo") and hasattr(t, "hi") and not isinstance(t, tuple):
            t = (t.lo, t.hi)                                    # a LazySets-style Interval
        t0, t1 = t
        idx = np.flatnonzero((hi >= t0) & (lo <= t1))
        if len(idx) == 0:
            raise ValueError("the time interval %s does not intersect the time span %s of the given flowpipe"
                             % (tuple(t), tuple(self.tspan)))
        return FlowpipeSlice(self[i] for i in idx)

    def support_function_matrix(self):
        """Flowpipe.jl:304-306."""
        return self.ext["sfmat"]() if callable(self.ext.get("sfmat")) else self.ext["sfmat"]

    def to_intervals(self, rows=(1, 2)):
        """Flowpipe.jl:311-331: rows = (idx_pos_dir, idx_neg_dir), 1-based rows of the support
        function matrix of a positive and the matching negative direction; returns the list of
        (lo, hi, Δt) = (−ρ(−d), ρ(d), tspan) per reach-set.  Only the two rows are fetched from
        the device."""
        if len(rows) != 2:
            raise ValueError("expected the number of rows of the support function matrix to be 2, got %d" % len(rows))
        plan = self.device
        if plan is not None and hasattr(plan, "fetch_rows"):
            pos, neg = plan.fetch_rows([rows[0] - 1, rows[1] - 1])
        else:
            mat = self.support_function_matrix()
            if mat.shape[0] < 2:
                raise ValueError("the number of rows of the support function matrix should be at least 2, got %d"
                                 % mat.shape[0])
            pos, neg = mat[rows[0] - 1], mat[rows[1] - 1]
        lo, hi = self._times()
        from .sets import Interval
        return [ReachSet(Interval(-neg[k], pos[k]), TimeInterval(lo[k], hi[k])) for k in range(self._n)]

    def rho(self, d):
        """ρ(d, fp) = max_k ρ(d, F[k]) (AbstractFlowpipe.jl:35-37)."""
        if self.device is not None and hasattr(self.device, "rho_flowpipe"):
            return self.device.rho_flowpipe(d)
        return max(R.rho(d) for R in self)

    def argmax_rho(self, d):
        """Index of the first reach-set attaining ρ(d, fp).  With a device-resident flowpipe the per-step support
        values come from one device reduction and only the index travels."""
        if self.device is not None and hasattr(self.device, "rho_per_step"):
            return int(np.argmax(self.device.rho_per_step(d)))
        if self.device is not None and hasattr(self.device, "argmax_over_steps"):
            k = self.device.argmax_over_steps(d)
            if k is not None and 0 <= k < self._n:
                return k
        return int(np.argmax([R.rho(d) for R in self]))

    def sigma(self, d):
        """σ(d, fp) (AbstractFlowpipe.jl:56-58): the flowpipe behaves like the union of its reach-sets, so this is the
        support vector of a reach-set attaining the maximum; only that one set is materialised."""
        return self[self.argmax_rho(d)].sigma(d)


class MixedFlowpipe:
    """Vector of flowpipes, one per initial-set split (MixedFlowpipe.jl:20-23); ρ = max (:66-68)."""

    def __init__(self, flowpipes, ext=None):
        self.flowpipes = list(flowpipes)
        self.ext = ext if ext is not None else {}

    def __len__(self):
        return len(self.flowpipes)

    def __getitem__(self, i):
        return self.flowpipes[i]

    def __iter__(self):
        return iter(self.flowpipes)

    def rho(self, d):
        dev = getattr(self, "device", None)
        if dev is not None and hasattr(dev, "rho_flowpipe"):
            try:
                return dev.rho_flowpipe(d)       # one device reduction over every split and step
            except NotImplementedError:
                pass
        return max(fp.rho(d) for fp in self.flowpipes)

    def numrsets(self):
        """MixedFlowpipe.jl:36."""
        return sum(len(fp) for fp in self.flowpipes)

    def tspan(self, i=None):
        """MixedFlowpipe.jl:40-43: the time span of a mixed flowpipe is only defined per flowpipe (0-based i)."""
        if i is None:
            raise ValueError("for mixed flowpipes you should specify the index, as in `tspan(fp, i)`")
        return self.flowpipes[i].tspan


class ReachSolution:
    """(F, alg, ext) (Flowpipes/solutions.jl:60-64); forwards the flowpipe interface."""

    def __init__(self, F, alg, ext=None):
        self.F, self.alg, self.ext = F, alg, ext if ext is not None else {}

    def __getattr__(self, name):
        return getattr(self.F, name)

    def __len__(self):
        return len(self.F)

    def __getitem__(self, k):
        return self.F[k]

    def __iter__(self):
        return iter(self.F)

    def __call__(self, t):
        return self.F(t)

    def shift(self, t0):
        """solutions.jl:131-133."""
        return ReachSolution(self.F.shift(t0), self.alg, self.ext)


def shift(X, t0):
    """`shift(fp | sol, t0)` as a free function, like the reference's."""
    return X.shift(t0)


def rho(d, X):
    """ρ(d, ·) for reach-sets, flowpipes and solutions."""
    return X.rho(d)


def sigma(d, X):
    """σ(d, ·) for reach-sets, flowpipes and solutions."""
    return X.sigma(d)
