"""LGG09 (Le Guernic–Girard support-function algorithm): host-side mirror of the reference's
operator interface over the device path.

  struct / constructor : src/Algorithms/LGG09/LGG09.jl:46-103
  discrete `post`      : src/Algorithms/LGG09/post.jl:4-50
  reach_homog_LGG09!   : src/Algorithms/LGG09/reach_homog.jl:6-33   (-> reach_lgg09_plan_*)
  reach_inhomog_LGG09! : src/Algorithms/LGG09/reach_inhomog.jl:5-33 (-> reach_lgg09_plan_*)

`threaded`, `sparse`, `cache` and `static` are accepted for drop-in compatibility and ignored
by the device path (the reference itself ignores `static`, post.jl:34): every direction runs in
one batched pass, so there is nothing to thread or cache.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from . import sets as S
from .directions import AbstractDirections, get_template, template_from_vars
from .flowpipe import Flowpipe, TemplateReachSet, zeroT, samedir
from .support_program import flatten, to_c, invariant_to_c


class _ApproxModel:
    """Approximation models are plain option records and compare by value, like the reference's structs."""

    def __eq__(self, other):
        return type(self) is type(other) and vars(self) == vars(other)

    __hash__ = None

    def __repr__(self):
        return "%s(%s)" % (self.name, ", ".join("%s=%r" % kv for kv in vars(self).items()))


class Forward(_ApproxModel):
    name = "Forward"

    def __init__(self, sih="concrete", exp="base", setops="lazy"):
        self.sih, self.exp, self.setops = sih, exp, setops


class NoBloating(_ApproxModel):
    name = "NoBloating"

    def __init__(self, exp="base", setops="lazy"):
        self.exp, self.setops = exp, setops


class FirstOrderZonotope(_ApproxModel):
    name = "FirstOrderZonotope"

    def __init__(self, exp="base"):
        self.exp = exp


@dataclass
class LGG09:
    δ: float
    approx_model: object
    template: AbstractDirections
    static: bool
    threaded: bool
    sparse: bool
    cache: bool
    vars: object
    # device knobs (no counterpart in the reference)
    time_block: int = 0
    share_negated: bool = True

    def __init__(self, *, δ=None, delta=None, approx_model=None, template=None, dirs=None, vars=None,
                 n=None, dim=None, static=False, threaded=False, sparse=False, cache=True,
                 time_block=0, share_negated=True):
        if δ is None:
            δ = delta
        if δ is None:
            raise TypeError("LGG09: the step size `δ` is required")
        _n = n if n is not None else dim
        _template = dirs if dirs is not None else template
        if vars is not None:
            if _n is None:
                raise ValueError("the ambient dimension, `n` (or `dim`), should be specified")
            directions = template_from_vars(vars, _n)
        elif _template is not None:
            directions = get_template(_template, _n)
        else:
            raise ValueError("either `vars`, `dirs` or `template` should be specified")
        self.δ = float(δ)
        self.approx_model = approx_model if approx_model is not None else Forward()
        self.template = directions
        self.static, self.threaded, self.sparse, self.cache = static, threaded, sparse, cache
        self.vars = vars
        self.time_block = int(time_block)
        self.share_negated = bool(share_negated)

    @property
    def delta(self):
        return self.δ


def step_size(alg):
    return alg.δ


def _as_torch(ptr, ndirs, nsteps, device):
    """View the device-resident ρℓ (ndirs × nsteps, column-major) as a torch tensor [nsteps, ndirs]
    without copying (plumbing for strided fetches and NCCL)."""
    import torch

    class _Holder:
        __cuda_array_interface__ = {"shape": (nsteps, ndirs), "typestr": "<f8", "data": (int(ptr), False),
                                    "version": 3, "strides": None}
    return torch.as_tensor(_Holder(), device=torch.device("cuda", device))


class Lgg09Plan:
    """Device-resident LGG09 run: owns a `reach_lgg09_plan` (inputs, ρ matrix, statistics)."""

    def __init__(self, ctx, n, ndirs, nsteps, time_block=0, share_negated=True, tile_m=0, tile_n=0, path=0,
                 digit_planes=0):
        self.ctx, self.lib = ctx, ctx.lib
        self.n, self.ndirs, self.nsteps = int(n), int(ndirs), int(nsteps)
        opts = L.reach_lgg09_opts()
        self.lib.reach_lgg09_default_opts(C.byref(opts))
        opts.time_block, opts.share_negated = int(time_block), int(bool(share_negated))
        opts.tile_m, opts.tile_n = int(tile_m), int(tile_n)
        opts.path = {"auto": 0, "gemm": 1, "chain": 2, "int8": 3, "compact": 4}.get(path, path)
        opts.reserved[0] = int(digit_planes)      # 0 = residues modulo 14 moduli (default); 4..8 digit planes; 10..16 moduli
        self.handle = C.c_void_p()
        ctx.check(self.lib.reach_lgg09_plan_create(ctx.handle, self.n, self.ndirs, self.nsteps,
                                                   C.byref(opts), C.byref(self.handle)))
        self._dirs = None
        self._sfmat = None

    def upload(self, Phi, dirs, Omega0, V=None, X=None):
        """Flatten the sets and copy the inputs to the device (host→device traffic only)."""
        n = self.n
        Phi = L.fmat(Phi)
        dirs = L.fmat(dirs)
        if Phi.shape != (n, n):
            raise ValueError("Φ is %s, expected (%d, %d)" % (Phi.shape, n, n))
        if dirs.shape != (n, self.ndirs):
            raise ValueError("the problem's dimension %d doesn't match the dimension of the template "
                             "directions, %d" % (n, dirs.shape[0]))
        if S.dim(Omega0) != n:
            raise ValueError("Ω₀ has dimension %d, expected %d" % (S.dim(Omega0), n))
        om = to_c(flatten(Omega0, Phi), n)
        v = to_c(flatten(V, None), n) if V is not None else None   # never the r_{k+1} shortcut: rho(r_k, V) is evaluated at r_k
        if v is not None and v.expr.nalt != 1:
            raise NotImplementedError("input sets containing a convex hull are not supported")
        inv, keep = invariant_to_c(X, n)
        self._keep = (Phi, dirs, om, v, inv, keep)
        self.ctx.check(self.lib.reach_lgg09_plan_upload(
            self.handle, L.dptr(Phi), L.dptr(dirs), om.ptr(), v.ptr() if v is not None else None,
            C.byref(inv) if inv is not None else None))
        self._dirs = dirs
        self._sfmat = None
        self.h2d_bytes = Phi.nbytes + dirs.nbytes
        return self

    def set_output(self, device_ptr):
        self.ctx.check(self.lib.reach_lgg09_plan_set_output(self.handle, C.c_void_p(int(device_ptr))))

    # ---- the row stack in steps (reach_lgg09_plan_stack_*): the ranks of a sharded run split its rows ----
    def stack_info(self):
        """(device pointer, rows per block, leading dimension in doubles, time block S) of the plan's row stack."""
        ptr, nbe, ld, S = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int32()
        self.ctx.check(self.lib.reach_lgg09_plan_stack_info(self.handle, C.byref(ptr), C.byref(nbe), C.byref(ld), C.byref(S)))
        return ptr.value, nbe.value, ld.value, S.value

    def stack_step(self, step, row_lo=-1, row_hi=-1):
        """Rows [row_lo, row_hi) of build step `step` (asynchronous); returns (rows of the step, first stack row they go to).
        With row_lo < 0 nothing is computed."""
        rows, dest = C.c_int64(), C.c_int64()
        self.ctx.check(self.lib.reach_lgg09_plan_stack_step(self.handle, int(step), int(row_lo), int(row_hi), C.byref(rows),
                                                            C.byref(dest)))
        return rows.value, dest.value

    def stack_ready(self):
        self.ctx.check(self.lib.reach_lgg09_plan_stack_ready(self.handle))

    def run(self, sync=True):
        self.ctx.check(self.lib.reach_lgg09_plan_run(self.handle))
        self._sfmat = None
        if sync:
            self.sync()
        return self

    def sync(self):
        self.ctx.check(self.lib.reach_lgg09_plan_sync(self.handle))

    @property
    def nsteps_done(self):
        out = C.c_int64()
        self.ctx.check(self.lib.reach_lgg09_plan_nsteps_done(self.handle, C.byref(out)))
        return out.value

    def fetch(self, step0=0, nsteps=None):
        """ρℓ[:, step0:step0+nsteps] as an (ndirs, nsteps) array (device→host copy)."""
        nsteps = self.nsteps - step0 if nsteps is None else nsteps
        out = np.empty((self.ndirs, nsteps), order="F")
        self.ctx.check(self.lib.reach_lgg09_plan_fetch(self.handle, step0, nsteps, L.dptr(out)))
        return out

    def sfmat(self):
        if self._sfmat is None:
            self._sfmat = self.fetch(0, self.nsteps_done)
        return self._sfmat

    def invariant_info(self):
        """Coverage of the invariant by the device test: faces, faces with a certificate, whether the test equals
        the reference's exact `_isdisjoint(X, set(F[k]))`, and how many steps the last run computed before it
        stopped launching (include/reach.h, reach_lgg09_plan_invariant_info)."""
        f, c, e, k = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        self.ctx.check(self.lib.reach_lgg09_plan_invariant_info(self.handle, C.byref(f), C.byref(c), C.byref(e), C.byref(k)))
        return dict(faces=f.value, covered=c.value, exact=bool(e.value), steps_computed=k.value)

    def fetch_rows(self, rows):
        """Selected rows of ρℓ (all steps): strided device→host copies of 8·NSTEPS bytes each."""
        m = self.sfmat() if self._sfmat is not None else None
        if m is not None:
            return [m[r] for r in rows]
        import torch
        t = _as_torch(self.device_rho, self.ndirs, self.nsteps, self.ctx.device)
        self.sync()
        return [t[:self.nsteps_done, r].cpu().numpy() for r in rows]

    def max_over_steps(self):
        vmax = np.empty(self.ndirs)
        amax = np.empty(self.ndirs, dtype=np.int64)
        self.ctx.check(self.lib.reach_lgg09_plan_max_over_steps(
            self.handle, L.dptr(vmax), amax.ctypes.data_as(L.c_int64_p)))
        return vmax, amax

    def rho_flowpipe(self, d):
        """ρ(d, fp) = max over steps, reduced on the device when d is (a positive multiple of) a
        template direction (TemplateReachSet.jl:97-115 + AbstractFlowpipe.jl:35-37)."""
        d = np.asarray(d, dtype=np.float64)
        D = self._dirs
        for i in range(D.shape[1]):
            if np.array_equal(d, D[:, i]):
                return float(self.max_over_steps()[0][i])
        for i in range(D.shape[1]):
            ok, k = samedir(d, D[:, i])
            if ok:
                return float(self.max_over_steps()[0][i]) * k
        # not a template direction: max over the reach-sets of ρ(d, set(R_k)), one LP per step on the fetched support
        # values (TemplateReachSet.jl:112-114 through AbstractFlowpipe.jl:35-37) -- host side, as in the reference
        from .flowpipe import lp_support
        M = self.sfmat()
        return max(lp_support(d, D, M[:, k]) for k in range(M.shape[1]))

    def argmax_over_steps(self, d):
        """Step index attaining ρ(d, fp) for (a positive multiple of) a template direction, from the same device
        reduction as `rho_flowpipe`; None for any other direction."""
        d = np.asarray(d, dtype=np.float64)
        D = self._dirs
        for i in range(D.shape[1]):
            if np.array_equal(d, D[:, i]) or samedir(d, D[:, i])[0]:
                return int(self.max_over_steps()[1][i])
        return None

    @property
    def device_rho(self):
        return self.lib.reach_lgg09_plan_device_rho(self.handle)

    def stats(self):
        ms, gms = C.c_double(), C.c_double()
        launches, gl = C.c_int64(), C.c_int64()
        S_, tm, tn, mu = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self.ctx.check(self.lib.reach_lgg09_plan_stats(
            self.handle, C.byref(ms), C.byref(launches), C.byref(S_), C.byref(tm), C.byref(tn),
            C.byref(mu), C.byref(gms), C.byref(gl)))
        path, planes = C.c_int32(), C.c_int32()
        gmax, gchecks, fb = C.c_double(), C.c_int64(), C.c_int64()
        self.ctx.check(self.lib.reach_lgg09_plan_path(self.handle, C.byref(path), C.byref(planes), C.byref(gmax),
                                                      C.byref(gchecks), C.byref(fb)))
        drift, rfrom, rpl = C.c_double(), C.c_int64(), C.c_int64()
        self.ctx.check(self.lib.reach_lgg09_plan_guard(self.handle, C.byref(drift), C.byref(rfrom), C.byref(rpl)))
        return dict(ms=ms.value, launches=launches.value, time_block=S_.value, tile_m=tm.value,
                    tile_n=tn.value, nchains=mu.value, gemm_ms=gms.value, gemm_launches=gl.value,
                    path={1: "gemm", 2: "chain", 3: "int8", 4: "compact"}[path.value], digit_planes=planes.value,
                    guard_max=gmax.value, guard_checks=gchecks.value, fallback_pass=fb.value,
                    drift_max=drift.value, replay_from=rfrom.value, replayed_passes=rpl.value)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.reach_lgg09_plan_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Lgg09BatchPlan:
    """LGG09 over a vector of initial sets sharing Φ, the template and V (`reach_lgg09_batch_*`): the split caller
    `_solve_distributed` (src/Continuous/solve.jl:84-117) without re-running the chains per split.  The result
    stays on the device as [split][step][dir]."""

    def __init__(self, ctx, Phi, dirs, Omega0s, V, nsteps):
        self.ctx, self.lib = ctx, ctx.lib
        Phi, dirs = L.fmat(Phi), L.fmat(dirs)
        n = Phi.shape[0]
        if Phi.shape != (n, n):
            raise ValueError("Φ must be square, got %s" % (Phi.shape,))
        if dirs.shape[0] != n:
            raise ValueError("the problem's dimension %d doesn't match the dimension of the template "
                             "directions, %d" % (n, dirs.shape[0]))
        if len(Omega0s) < 1:
            raise ValueError("at least one initial set is required")
        for Om in Omega0s:
            if S.dim(Om) != n:
                raise ValueError("Ω₀ has dimension %d, expected %d" % (S.dim(Om), n))
        self.n, self.ndirs, self.nsteps, self.nsplits = n, dirs.shape[1], int(nsteps), len(Omega0s)
        oms = [to_c(flatten(Om, Phi), n) for Om in Omega0s]
        arr = (L.reach_setexpr * len(oms))(*[o.expr for o in oms])
        v = to_c(flatten(V, None), n) if V is not None else None   # never the r_{k+1} shortcut: rho(r_k, V) is evaluated at r_k
        if v is not None and v.expr.nalt != 1:
            raise NotImplementedError("input sets containing a convex hull are not supported")
        self._keep = (Phi, dirs, oms, arr, v)
        self._dirs = dirs
        self.handle = C.c_void_p()
        ctx.check(self.lib.reach_lgg09_batch_create(
            ctx.handle, n, self.ndirs, self.nsteps, self.nsplits, L.dptr(Phi), L.dptr(dirs),
            C.cast(arr, C.POINTER(L.reach_setexpr)), v.ptr() if v is not None else None, C.byref(self.handle)))
        self._cache = {}

    def run(self, sync=True):
        self.ctx.check(self.lib.reach_lgg09_batch_run(self.handle))
        self._cache = {}
        if sync:
            self.ctx.check(self.lib.reach_lgg09_batch_sync(self.handle))
        return self

    def fetch(self, split, step0=0, nsteps=None):
        """ρℓ of one split, (ndirs, nsteps), device→host."""
        nsteps = self.nsteps - step0 if nsteps is None else nsteps
        out = np.empty((self.ndirs, nsteps), order="F")
        self.ctx.check(self.lib.reach_lgg09_batch_fetch(self.handle, int(split), step0, nsteps, L.dptr(out)))
        return out

    def sfmat(self, split):
        if split not in self._cache:
            self._cache[split] = self.fetch(split)
        return self._cache[split]

    def max_over_steps(self, split=-1):
        """max over steps of every template row; split = -1: over all splits too (ρ(d, MixedFlowpipe))."""
        out = np.empty(self.ndirs)
        self.ctx.check(self.lib.reach_lgg09_batch_max_over_steps(self.handle, int(split), L.dptr(out)))
        return out

    def rho_flowpipe(self, d, split=-1):
        d = np.asarray(d, dtype=np.float64)
        D = self._dirs
        for i in range(D.shape[1]):
            ok, k = samedir(d, D[:, i])
            if ok:
                return float(self.max_over_steps(split)[i]) * k
        # not a template direction: max over the reach-sets of ρ(d, set(R_k)), one LP per step on the fetched support
        # values (TemplateReachSet.jl:112-114 through AbstractFlowpipe.jl:35-37) -- host side, as in the reference
        from .flowpipe import lp_support
        splits = range(self.nsplits) if split < 0 else [split]
        return max(lp_support(d, D, self.sfmat(i)[:, k]) for i in splits for k in range(self.nsteps))

    def stats(self):
        ms, launches, chunk = C.c_double(), C.c_int64(), C.c_int64()
        S_, na, kc = C.c_int32(), C.c_int32(), C.c_int32()
        self.ctx.check(self.lib.reach_lgg09_batch_stats(self.handle, C.byref(ms), C.byref(launches), C.byref(S_),
                                                        C.byref(na), C.byref(kc), C.byref(chunk)))
        return dict(ms=ms.value, launches=launches.value, time_block=S_.value, alternatives=na.value,
                    feature_cols=kc.value, chunk_steps=chunk.value)

    @property
    def device_rho(self):
        return self.lib.reach_lgg09_batch_device_rho(self.handle)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.reach_lgg09_batch_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _SplitView:
    """One split of a batch, with the interface Flowpipe expects from a device plan."""

    def __init__(self, batch, split):
        self.batch, self.split = batch, split

    def sfmat(self):
        return self.batch.sfmat(self.split)

    def fetch_rows(self, rows):
        m = self.sfmat()
        return [m[r] for r in rows]

    def rho_flowpipe(self, d):
        return self.batch.rho_flowpipe(d, self.split)


def _same_set(X, Y):
    """Structural equality of two (flattened) sets: same alternatives, same terms, same numbers."""
    a, b = flatten(X, None), flatten(Y, None)
    if len(a) != len(b) or any(len(x) != len(y) for x, y in zip(a, b)):
        return False
    for x, y in zip(a, b):
        for t, u in zip(x, y):
            if t.kind != u.kind or t.which != u.which:
                return False
            for name in ("M", "c", "r", "G"):
                v, w = getattr(t, name), getattr(u, name)
                if (v is None) != (w is None) or (v is not None and not np.array_equal(v, w)):
                    return False
    return True


def post_batch(alg, ivps, NSTEPS, Δt0=zeroT, ctx=None):
    """The vector-of-initial-sets form of `solve` (src/Continuous/solve.jl:84-117) for LGG09: `ivps` are the
    discretised problems of the splits (same Φ and V, different Ω₀).  Returns a MixedFlowpipe whose flowpipes
    share one device-resident batch; ρ(d, ·) over a template direction is one device reduction."""
    from .flowpipe import MixedFlowpipe
    ctx = ctx or L.default_context()
    template = alg.template
    first = ivps[0]
    if first.n != template.n:
        raise AssertionError("the problems' dimension %d doesn't match the dimension of the template "
                             "directions, %d" % (first.n, template.n))
    if any(iv.X is not None for iv in ivps):
        raise NotImplementedError("the batched split caller does not take an invariant")
    for iv in ivps[1:]:
        if not np.array_equal(iv.Phi, first.Phi):
            raise ValueError("the splits must share Φ")
        if (iv.V is None) != (first.V is None) or (first.V is not None and not _same_set(iv.V, first.V)):
            raise ValueError("the splits must share the input set V")
    batch = Lgg09BatchPlan(ctx, first.Phi, template.matrix(), [iv.Omega0 for iv in ivps], first.V, NSTEPS)
    batch.run()
    fps = []
    for i in range(len(ivps)):
        view = _SplitView(batch, i)

        def make(k, dt, view=view):
            return TemplateReachSet(template, view.sfmat()[:, k], dt)

        fps.append(Flowpipe(NSTEPS, make, alg.δ, Δt0, {"sfmat": view.sfmat, "alg_vars": alg.vars}, device=view))
    mixed = MixedFlowpipe(fps, {"batch": batch})
    mixed.device = batch
    return mixed


def reach_homog_LGG09(dirs, Omega0, Phi, NSTEPS, X=None, ctx=None, **opts):
    """`reach_homog_LGG09!` (reach_homog.jl:6-33): returns the device-resident plan; `.sfmat()`
    is the ndirs × NSTEPS matrix ρℓ."""
    return reach_inhomog_LGG09(dirs, Omega0, Phi, NSTEPS, X, None, ctx=ctx, **opts)


def reach_inhomog_LGG09(dirs, Omega0, Phi, NSTEPS, X=None, U=None, ctx=None, **opts):
    """`reach_inhomog_LGG09!` (reach_inhomog.jl:5-33)."""
    ctx = ctx or L.default_context()
    D = dirs.matrix() if isinstance(dirs, AbstractDirections) else np.asarray(dirs, dtype=np.float64)
    Phi = np.asarray(Phi, dtype=np.float64)
    plan = Lgg09Plan(ctx, Phi.shape[0], D.shape[1], NSTEPS, **opts)
    plan.upload(Phi, D, Omega0, U, X)
    plan.run()
    return plan


def post(alg, ivp, NSTEPS=None, Δt0=zeroT, ctx=None, **kwargs):
    """Discrete `post(alg::LGG09, ivp::IVP{<:AbstractDiscreteSystem}, NSTEPS; Δt0)` (post.jl:4-50).

    `ivp` is a `discretization.DiscreteIVP` (Φ, Ω₀, V, X).  Returns a Flowpipe of
    TemplateReachSets whose `ext` holds `:sfmat` (fetched lazily) and `:alg_vars`."""
    template = alg.template
    if ivp.n != template.n:
        raise AssertionError("the problems' dimension %d doesn't match the dimension of the template "
                             "directions, %d" % (ivp.n, template.n))
    if NSTEPS is None:
        NSTEPS = kwargs.get("NSTEPS")
        if NSTEPS is None:
            raise ValueError("`NSTEPS` not specified")
    plan = reach_inhomog_LGG09(template, ivp.Omega0, ivp.Phi, NSTEPS, ivp.X, ivp.V, ctx=ctx,
                               time_block=alg.time_block, share_negated=alg.share_negated)
    nsets = plan.nsteps_done
    info = plan.invariant_info()
    if info["faces"] and not info["exact"]:
        # the reference decides `_isdisjoint(X, set(F[k]))` with an LP per step; the device test is sufficient only here
        import warnings
        warnings.warn("LGG09 invariant: %d of %d faces are tested on the device and the test is sufficient, not exact, for "
                      "this template -- the flowpipe may be longer than the reference's (never shorter)"
                      % (info["covered"], info["faces"]), RuntimeWarning, stacklevel=2)

    def make(k, dt):
        return TemplateReachSet(template, plan.sfmat()[:, k], dt)

    ext = {"sfmat": plan.sfmat, "alg_vars": alg.vars}
    return Flowpipe(nsets, make, alg.δ, Δt0, ext, device=plan)


def reach_homog_dir_LGG09(rho_vec, Omega0, PhiT, ell, NSTEPS, cache=True, ctx=None, **opts):
    """`reach_homog_dir_LGG09!(ρvec_ℓ, Ω₀, Φᵀ, ℓ, NSTEPS, cache)` — the vector-output twin (reach_homog.jl:80-106):
    ρvec_ℓ[i] = ρ((Φᵀ)^i ℓ, Ω₀), i = 0..NSTEPS-1, for ONE direction.  Note the reference passes Φ **transposed** here.
    Same device kernels as the matrix form with ndirs = 1 (one chain; `cache` has nothing to select on the device).
    `rho_vec` (length ≥ NSTEPS) is filled in place and returned."""
    return reach_inhomog_dir_LGG09(rho_vec, Omega0, PhiT, None, ell, NSTEPS, cache, ctx=ctx, **opts)


def reach_inhomog_dir_LGG09(rho_vec, Omega0, PhiT, U, ell, NSTEPS, cache=True, ctx=None, **opts):
    """`reach_inhomog_dir_LGG09!(ρvec_ℓ, Ω₀, Φᵀ, U, ℓ, NSTEPS, cache)` (reach_inhomog.jl:98-128):
    ρvec_ℓ[i] = ρ(r_i, Ω₀) + s_i, s_{i+1} = s_i + ρ(r_i, U), r_{i+1} = Φᵀ r_i."""
    ctx = ctx or L.default_context()
    PhiT = np.asarray(PhiT, dtype=np.float64)
    ell = np.asarray(ell, dtype=np.float64).reshape(-1)
    if PhiT.shape != (len(ell), len(ell)):
        raise ValueError("Φᵀ is %s, the direction has length %d" % (PhiT.shape, len(ell)))
    rho_vec = np.asarray(rho_vec) if not isinstance(rho_vec, np.ndarray) else rho_vec
    if rho_vec.shape[0] < NSTEPS:
        raise IndexError("ρvec_ℓ has length %d < NSTEPS = %d (BoundsError in the reference)" % (rho_vec.shape[0], NSTEPS))
    plan = Lgg09Plan(ctx, len(ell), 1, NSTEPS, **opts)
    plan.upload(np.ascontiguousarray(PhiT.T), ell.reshape(-1, 1), Omega0, U)
    plan.run()
    rho_vec[:NSTEPS] = plan.fetch(0, NSTEPS)[0]
    plan.close()
    return rho_vec
