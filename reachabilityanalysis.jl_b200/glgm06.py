"""GLGM06 (Girard–Le Guernic–Maler zonotope flowpipe): host-side mirror of the reference's
operator interface over the device path.

  struct / constructor   : src/Algorithms/GLGM06/GLGM06.jl:51-81
  discrete `post`        : src/Algorithms/GLGM06/post.jl:4-61
  reach_homog_GLGM06!    : src/Algorithms/GLGM06/reach_homog.jl:33-73   (-> reach_glgm06_plan_*)
  reach_inhomog_GLGM06!  : src/Algorithms/GLGM06/reach_inhomog.jl:6-43  (-> reach_glgm06_plan_*)
  split caller           : src/Continuous/solve.jl:84-117 (`_solve_distributed`) -> post_batch

`static`, `dim`, `ngens` and `preallocate` are accepted for drop-in compatibility and ignored
(they select CPU memory-layout tricks).  The reduction method is GIR05, the reference's default.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from . import sets as S
from .flowpipe import Flowpipe, MixedFlowpipe, ReachSet, zeroT
from .zonotope_ops import to_zonotope
from .lgg09 import Forward, FirstOrderZonotope  # noqa: F401


@dataclass
class GLGM06:
    δ: float
    approx_model: object
    max_order: int
    static: bool
    dim: object
    ngens: object
    preallocate: bool
    reduction_method: str
    disjointness_method: object

    def __init__(self, *, δ=None, delta=None, approx_model=None, max_order=5, static=False, dim=None,
                 ngens=None, preallocate=True, reduction_method="GIR05", disjointness_method=None):
        if δ is None:
            δ = delta
        if δ is None:
            raise TypeError("GLGM06: the step size `δ` is required")
        if reduction_method != "GIR05":
            raise NotImplementedError("only the GIR05 reduction method is implemented on the device")
        if int(max_order) < 1:
            raise ValueError("the target order should be at least 1, but it is %s" % (max_order,))
        self.δ = float(δ)
        self.approx_model = approx_model if approx_model is not None else FirstOrderZonotope()
        self.max_order = int(max_order)
        self.static, self.dim, self.ngens, self.preallocate = static, dim, ngens, preallocate
        self.reduction_method, self.disjointness_method = reduction_method, disjointness_method

    @property
    def delta(self):
        return self.δ


class Glgm06Plan:
    """Device-resident GLGM06 run for a batch of initial zonotopes sharing Φ and U."""

    def __init__(self, ctx, n, nsteps, nsplits, max_order, p0, pU):
        self.ctx, self.lib = ctx, ctx.lib
        self.n, self.nsteps, self.nsplits = int(n), int(nsteps), int(nsplits)
        self.max_order, self.p0, self.pU = int(max_order), int(p0), int(pU)
        self.handle = C.c_void_p()
        ctx.check(self.lib.reach_glgm06_plan_create(ctx.handle, self.n, self.nsteps, self.nsplits,
                                                    self.max_order, self.p0, self.pU, C.byref(self.handle)))

    def upload(self, Phi, c0, G0, cU=None, GU=None):
        n = self.n
        Phi = L.fmat(Phi)
        if Phi.shape != (n, n):
            raise ValueError("Φ is %s, expected (%d, %d)" % (Phi.shape, n, n))
        c0 = np.ascontiguousarray(np.asarray(c0, dtype=np.float64).reshape(self.nsplits, n))
        G0 = np.ascontiguousarray(np.asarray(G0, dtype=np.float64).reshape(self.nsplits, self.p0, n))
        cUa = None if cU is None else np.ascontiguousarray(cU, dtype=np.float64)
        GUa = None if GU is None else np.ascontiguousarray(np.asarray(GU, dtype=np.float64).reshape(max(self.pU, 0), n))
        self._keep = (Phi, c0, G0, cUa, GUa)
        self.ctx.check(self.lib.reach_glgm06_plan_upload(self.handle, L.dptr(Phi), L.dptr(c0), L.dptr(G0),
                                                         L.dptr(cUa), L.dptr(GUa)))
        return self

    def run(self, sync=True):
        self.ctx.check(self.lib.reach_glgm06_plan_run(self.handle))
        if sync:
            self.sync()
        return self

    def sync(self):
        self.ctx.check(self.lib.reach_glgm06_plan_sync(self.handle))

    def ngens(self, split, step):
        out = C.c_int()
        self.ctx.check(self.lib.reach_glgm06_plan_ngens(self.handle, split, step, C.byref(out)))
        return out.value

    def fetch(self, split, step, with_selection=False):
        """(c, G[, sel]) of one zonotope; G is n × ngens."""
        ng = self.ngens(split, step)
        c = np.empty(self.n)
        Gt = np.empty((max(ng, 1), self.n))            # rows = generators == n × ng column-major
        sel = np.empty(max(ng, 1), dtype=np.int32)
        self.ctx.check(self.lib.reach_glgm06_plan_fetch(self.handle, split, step, L.dptr(c), L.dptr(Gt),
                                                        sel.ctypes.data_as(L.c_int32_p)))
        G = Gt[:ng].T.copy()
        return (c, G, sel[:ng]) if with_selection else (c, G)

    def support(self, d):
        """ρ(d, Z) of every stored zonotope, on the device: array (nsplits, nsteps)."""
        d = np.ascontiguousarray(d, dtype=np.float64)
        if d.shape != (self.n,):
            raise ValueError("direction of length %d in a %d-dimensional flowpipe" % (d.size, self.n))
        out = np.empty((self.nsplits, self.nsteps), order="F")
        self.ctx.check(self.lib.reach_glgm06_plan_support(self.handle, L.dptr(d), L.dptr(out)))
        return out

    def stats(self):
        ms, launches, stored = C.c_double(), C.c_int64(), C.c_int64()
        self.ctx.check(self.lib.reach_glgm06_plan_stats(self.handle, C.byref(ms), C.byref(launches), C.byref(stored)))
        return dict(ms=ms.value, launches=launches.value, bytes_stored=stored.value)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.reach_glgm06_plan_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def reduce_order(Z, r, ctx=None, with_selection=False):
    """`reduce_order(Z, r, GIR05())` on the device (GLGM06/post.jl:26)."""
    ctx = ctx or L.default_context()
    c, G = to_zonotope(Z)
    n, p = G.shape
    Gt = np.ascontiguousarray(G.T)
    out = np.empty((max(p, r * n, 1), n))
    sel = np.empty(max(p, r * n, 1), dtype=np.int32)
    pred = C.c_int()
    ctx.check(ctx.lib.reach_reduce_order_gir05(ctx.handle, n, p, int(r), L.dptr(Gt), L.dptr(out), C.byref(pred),
                                               sel.ctypes.data_as(L.c_int32_p)))
    Zr = S.Zonotope(c, out[:pred.value].T.copy())
    return (Zr, sel[:pred.value]) if with_selection else Zr


class _SplitView:
    """What a Flowpipe of split `s` needs from the shared plan."""

    def __init__(self, plan, split):
        self.plan, self.split = plan, split
        self._sup = {}

    def rho_flowpipe(self, d):
        d = np.asarray(d, dtype=np.float64)
        key = d.tobytes()
        if key not in self.plan.__dict__.setdefault("_support_cache", {}):
            self.plan._support_cache[key] = self.plan.support(d)
        return float(self.plan._support_cache[key][self.split][:self.nsets].max())

    def rho_per_step(self, d):
        """ρ(d, Z_k) of every step of this split (one device pass over all zonotopes, cached per direction)."""
        self.rho_flowpipe(d)
        return self.plan._support_cache[np.asarray(d, dtype=np.float64).tobytes()][self.split][:self.nsets]


def _flowpipe_of_split(plan, split, alg, dt0, nsets=None):
    def make(k, dt):
        c, G = plan.fetch(split, k)
        return ReachSet(S.Zonotope(c, G), dt)

    view = _SplitView(plan, split)
    view.nsets = plan.nsteps if nsets is None else nsets
    return Flowpipe(view.nsets, make, alg.δ, dt0, {}, device=view)


def _sets_kept(plan, X):
    """Invariant variants (reach_homog.jl:76-147, reach_inhomog.jl:46-86): per split, the number of
    reach-sets before the first one that is disjoint from X (`_isdisjoint(X, R_k) && break`; the
    first set is stored unchecked, the disjoint one is dropped).  Decided on the device from the
    support function of every stored zonotope: for a half-space {a·x ≤ b} the test
    −ρ(−a, R_k) > b is exact; for a box invariant each face is tested this way (the
    `BoxEnclosure` disjointness method -- sufficient, not necessary)."""
    n, nsteps = plan.n, plan.nsteps
    if X is None or isinstance(X, S.Universe):
        return np.full(plan.nsplits, nsteps, dtype=np.int64)
    if isinstance(X, S.HalfSpace):
        faces = [(X.a, X.b)]
    elif isinstance(X, (S.Hyperrectangle, S.BallInf)):
        r = X.radius if isinstance(X, S.Hyperrectangle) else np.full(n, X.radius)
        faces = []
        for i in range(n):
            e = np.zeros(n)
            e[i] = 1.0
            faces.append((e, X.center[i] + r[i]))
            faces.append((-e, -(X.center[i] - r[i])))
    else:
        raise NotImplementedError("invariant of type %s is not supported on the device" % type(X).__name__)
    bad = np.zeros((plan.nsplits, nsteps), dtype=bool)
    for a, b in faces:
        bad |= -plan.support(-np.asarray(a, dtype=np.float64)) > b
    bad[:, 0] = False
    first = np.where(bad.any(axis=1), bad.argmax(axis=1), nsteps)
    return first.astype(np.int64)


def _run_batch(alg, Phi, Omega0_list, V, NSTEPS, ctx):
    ctx = ctx or L.default_context()
    zs = [to_zonotope(Om) for Om in Omega0_list]                 # post.jl:25
    n = len(zs[0][0])
    p0 = zs[0][1].shape[1]
    for c, G in zs:
        if G.shape != (n, p0):
            raise ValueError("all initial sets of a batch must have the same number of generators")
    if V is not None:
        cU, GU = to_zonotope(V)                                  # post.jl:54
        pU = GU.shape[1]
        GUt = np.ascontiguousarray(GU.T)
    else:
        cU, GUt, pU = None, None, -1
    plan = Glgm06Plan(ctx, n, NSTEPS, len(zs), alg.max_order, p0, pU)
    c0 = np.stack([c for c, _ in zs])
    G0 = np.stack([np.ascontiguousarray(G.T) for _, G in zs]) if p0 > 0 else np.zeros((len(zs), 0, n))
    plan.upload(Phi, c0, G0, cU, GUt)
    plan.run()
    return plan


def post(alg, ivp, NSTEPS=None, Δt0=zeroT, ctx=None, **kwargs):
    """Discrete `post(alg::GLGM06, ivp, NSTEPS; Δt0)` (post.jl:4-61) for one initial set."""
    if NSTEPS is None:
        NSTEPS = kwargs.get("NSTEPS")
        if NSTEPS is None:
            raise ValueError("`NSTEPS` not specified")
    plan = _run_batch(alg, ivp.Phi, [ivp.Omega0], ivp.V, NSTEPS, ctx)
    return _flowpipe_of_split(plan, 0, alg, Δt0, int(_sets_kept(plan, ivp.X)[0]))


def post_batch(alg, ivp, T=None, NSTEPS=None, Δt0=zeroT, ctx=None):
    """`solve(IVP(S, [X0_1 … X0_m]); alg=GLGM06(…))`: every split goes through the same Φ and the
    same input chain W_k, computed once (the reference re-runs normalize + discretize + post per
    split, solve.jl:107-117).  Returns a MixedFlowpipe whose flowpipes share one device plan."""
    from .solve import discretize, IVP, _nsteps
    if NSTEPS is None:
        if T is None:
            raise ValueError("the time horizon `T` (or `NSTEPS`) should be specified")
        NSTEPS = _nsteps(T, alg.δ)
    divps = [discretize(IVP(ivp.A, X0, ivp.B, ivp.U, ivp.X), alg.δ, alg.approx_model) for X0 in ivp.X0]
    # one plan only makes sense if the splits really share Φ and the input zonotope (they do for the reference's models:
    # neither depends on X0); checked rather than assumed
    v0 = to_zonotope(divps[0].V) if divps[0].V is not None else None
    for d in divps[1:]:
        if not np.array_equal(d.Phi, divps[0].Phi):
            raise ValueError("the splits must share Φ")
        vd = to_zonotope(d.V) if d.V is not None else None
        if (vd is None) != (v0 is None) or (v0 is not None and not (np.array_equal(vd[0], v0[0]) and np.array_equal(vd[1], v0[1]))):
            raise ValueError("the splits must share the input set V")
    plan = _run_batch(alg, divps[0].Phi, [d.Omega0 for d in divps], divps[0].V, NSTEPS, ctx)
    kept = _sets_kept(plan, ivp.X)
    return MixedFlowpipe([_flowpipe_of_split(plan, s, alg, Δt0, int(kept[s])) for s in range(len(divps))],
                         {"plan": plan})
