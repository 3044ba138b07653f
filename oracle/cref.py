"""ctypes loader of oracle/c/libreach_ref.so, the C/OpenMP restatement of the reference's CPU loops
(TEST INFRASTRUCTURE / CPU BASELINE ONLY -- see oracle/c/reach_ref.c).

Only tests/, __graft_entry__.smoke() and the CPU legs of bench.py may import this module.  It does not import the
product package: the flattened support program is rebuilt here from duck-typed set objects with the same struct
layout as include/reach.h.
"""
import ctypes as C
import glob
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CDIR = os.path.join(_HERE, "c")
c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class _Term(C.Structure):
    _fields_ = [("kind", C.c_int32), ("which", C.c_int32), ("q", C.c_int32), ("p", C.c_int32),
                ("M", c_double_p), ("c", c_double_p), ("r", c_double_p), ("G", c_double_p)]


class _SetExpr(C.Structure):
    _fields_ = [("nalt", C.c_int32), ("alt_len", c_int32_p), ("terms", C.POINTER(_Term))]


_GEMM_T = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_double, c_double_p,
                      C.c_int64, c_double_p, C.c_int64, C.c_double, c_double_p, C.c_int64)
_lib = None
_gemm = None
_blas = None


def _cpu_tag():
    try:
        flags = [l for l in open("/proc/cpuinfo") if l.startswith("flags")][0]
    except Exception:
        flags = "generic"
    return hashlib.sha1(flags.encode()).hexdigest()[:12]


def _build_native():
    """-march=native build for the machine this runs on (the timed CPU arm should use the host's vector units)."""
    out = os.path.join(_CDIR, "build", _cpu_tag(), "libreach_ref.so")
    src = os.path.join(_CDIR, "reach_ref.c")
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    tmp = out + ".%d.tmp" % os.getpid()
    cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-fno-math-errno", "-shared", "-o", tmp, src, "-lm"]
    subprocess.run(cmd, check=True, capture_output=True)
    os.replace(tmp, out)
    return out


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = None
    try:
        path = _build_native()
    except Exception:
        generic = os.path.join(_CDIR, "libreach_ref.so")
        if os.path.exists(generic):
            path = generic
    if path is None:
        raise RuntimeError("oracle/c/libreach_ref.so is not built and gcc is unavailable (make -C oracle/c)")
    # OpenBLAS (pthreads) and libgomp both spin by default; between a GEMM and the threaded epilogue they would fight
    # over the same cores
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")
    lib = C.CDLL(path)
    lib.ref_num_threads.restype = C.c_int
    lib.ref_lgg09.restype = C.c_int
    lib.ref_lgg09.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, c_double_p, C.c_int, c_double_p,
                              C.POINTER(_SetExpr), C.POINTER(_SetExpr), C.c_int, C.c_int, c_double_p, C.c_int, C.c_void_p]
    lib.ref_reduce_order_gir05.restype = C.c_int
    lib.ref_reduce_order_gir05.argtypes = [C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_int32_p, c_int32_p]
    lib.ref_glgm06.restype = C.c_int
    lib.ref_glgm06.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, C.c_int,
                               c_double_p, c_double_p, C.c_int, c_double_p, c_double_p, C.c_int, c_int32_p,
                               C.POINTER(C.c_int64), c_double_p, c_double_p, C.c_int, c_int32_p, c_int32_p, C.c_int, C.c_int]
    _lib = lib
    return lib


def openblas_dgemm():
    """cblas_dgemm of the OpenBLAS NumPy ships (ILP64 build, symbol scipy_cblas_dgemm64_) -- the same library family
    Julia's LinearAlgebra calls for `mul!` -- as a raw function pointer for ref_lgg09's GEMM-per-step shape."""
    global _gemm, _blas
    if _gemm is not None:
        return _gemm
    libs = glob.glob(os.path.join(os.path.dirname(np.__file__), "..", "numpy.libs", "libscipy_openblas64_*.so"))
    if not libs:
        return None
    _blas = C.CDLL(libs[0])
    fn = getattr(_blas, "scipy_cblas_dgemm64_", None)
    if fn is None:
        return None
    _gemm = C.cast(fn, C.c_void_p)
    return _gemm


def set_blas_threads(nthreads):
    openblas_dgemm()
    if _blas is not None and hasattr(_blas, "scipy_openblas_set_num_threads64_"):
        _blas.scipy_openblas_set_num_threads64_(C.c_int(int(nthreads)))


def num_threads():
    return load().ref_num_threads()


# ---- flattened support program from duck-typed sets (same rules as oracle/lazysets_oracle.rho) ----
def _name(X):
    return type(X).__name__


def _flatten(X, T, alpha, which, Phi):
    k = _name(X)

    def leaf(kind, c=None, r=None, G=None):
        if alpha != 1.0:
            c = None if c is None else alpha * c
            r = None if r is None else abs(alpha) * r
            G = None if G is None else alpha * G
        return [[dict(kind=kind, which=which, M=T, c=c, r=r, G=G)]]

    if k == "Hyperrectangle":
        return leaf(0, np.array(X.center, dtype=float), np.array(X.radius, dtype=float))
    if k == "BallInf":
        return leaf(0, np.array(X.center, dtype=float), np.full(len(X.center), float(X.radius)))
    if k == "Interval":
        return leaf(0, np.array([(X.lo + X.hi) / 2]), np.array([(X.hi - X.lo) / 2]))
    if k == "Singleton":
        return leaf(0, np.array(X.element, dtype=float), None)
    if k == "ZeroSet":
        return [[]]
    if k == "Zonotope":
        return leaf(1, np.array(X.center, dtype=float), None, np.array(X.generators, dtype=float))
    if k == "LinearMap":
        M = np.asarray(X.M, dtype=float)
        if T is None and alpha == 1.0 and which == 0 and Phi is not None and M.shape == Phi.shape and np.array_equal(M, Phi):
            return _flatten(X.X, None, 1.0, 1, Phi)
        return _flatten(X.X, M.copy() if T is None else T @ M, alpha, which, Phi)
    if k == "Scale":
        return _flatten(X.X, T, alpha * X.alpha, which, Phi)
    if k == "MinkowskiSum":
        a, b = _flatten(X.X, T, alpha, which, Phi), _flatten(X.Y, T, alpha, which, Phi)
        return [x + y for x in a for y in b]
    if k == "ConvexHull":
        return _flatten(X.X, T, alpha, which, Phi) + _flatten(X.Y, T, alpha, which, Phi)
    raise TypeError("cannot flatten %s" % k)


class _CExpr:
    def __init__(self, alts, n):
        self.keep = []
        terms = [t for alt in alts for t in alt]
        arr = (_Term * max(1, len(terms)))()
        for i, t in enumerate(terms):
            q = 0 if t["M"] is None else t["M"].shape[1]
            arr[i].kind, arr[i].which, arr[i].q = t["kind"], t["which"], q
            arr[i].p = 0 if t["G"] is None else t["G"].shape[1]
            for name in ("M", "c", "r", "G"):
                v = t[name]
                if v is not None:
                    v = np.asfortranarray(v, dtype=np.float64)
                    self.keep.append(v)
                    setattr(arr[i], name, v.ctypes.data_as(c_double_p))
        lens = (C.c_int32 * len(alts))(*[len(a) for a in alts])
        self.expr = _SetExpr(len(alts), C.cast(lens, c_int32_p), C.cast(arr, C.POINTER(_Term)))
        self.keep += [arr, lens]


def reach_lgg09(dirs, Omega0, Phi, nsteps, U=None, shape=2, sparse=False, literal=False, share=False, nthreads=0):
    """rho_l (ndirs x nsteps) by the C restatement.  shape 1: per-direction chains over the cores (B1); shape 2: one
    GEMM per step (B2, OpenBLAS threads = nthreads or all)."""
    lib = load()
    Phi = np.asfortranarray(Phi, dtype=np.float64)
    dirs = np.asfortranarray(dirs, dtype=np.float64)
    n, ndirs = dirs.shape
    om = _CExpr(_flatten(Omega0, None, 1.0, 0, Phi), n)
    v = _CExpr(_flatten(U, None, 1.0, 0, None), n) if U is not None else None
    out = np.empty((ndirs, int(nsteps)), order="F")
    gemm = openblas_dgemm() if shape == 2 else None
    if shape == 2 and nthreads:
        set_blas_threads(nthreads)
    rc = lib.ref_lgg09(int(shape), n, ndirs, int(nsteps), Phi.ctypes.data_as(c_double_p), int(bool(sparse)),
                       dirs.ctypes.data_as(c_double_p), C.byref(om.expr), C.byref(v.expr) if v is not None else None,
                       int(bool(literal)), int(bool(share)), out.ctypes.data_as(c_double_p), int(nthreads), gemm)
    if rc != 0:
        raise RuntimeError("ref_lgg09 failed (rc=%d)" % rc)
    return out


def reduce_order_gir05(G, r):
    lib = load()
    G = np.asfortranarray(G, dtype=np.float64)
    n, p = G.shape
    out = np.zeros((n, max(p, r * n)), order="F")
    sel = np.zeros(max(p, 1), dtype=np.int32)
    nsel = C.c_int32(0)
    pout = lib.ref_reduce_order_gir05(n, p, int(r), G.ctypes.data_as(c_double_p), out.ctypes.data_as(c_double_p),
                                      sel.ctypes.data_as(c_int32_p), C.byref(nsel))
    if pout < 0:
        raise RuntimeError("ref_reduce_order_gir05 failed")
    return out[:, :pout].copy(), (None if nsel.value < 0 else sel[:nsel.value].copy())


def reach_inhomog_glgm06(Phi, c0, G0, nsteps, max_order, cU, GU, d=None, samples=(), shared_chain=True, nthreads=0):
    """GLGM06 over a batch of initial zonotopes: c0 (n, nsplits), G0 (nsplits, n, p0) already reduced (post.jl:26).
    Returns (rho or None [nsplits, nsteps], list of (c, G, sel) for `samples` = [(split, step), ...])."""
    lib = load()
    Phi = np.asfortranarray(Phi, dtype=np.float64)
    c0 = np.asfortranarray(c0, dtype=np.float64)
    n, nsplits = c0.shape
    G0 = np.ascontiguousarray(np.stack([np.asfortranarray(g).T for g in G0]))      # [split][p0][n] == n x p0 col-major each
    p0 = G0.shape[1]
    cU = np.ascontiguousarray(cU, dtype=np.float64)
    GU = np.asfortranarray(GU, dtype=np.float64)
    pU = GU.shape[1]
    gcap = max(p0, max_order * n) + pU + n
    ns = len(samples)
    ss = np.array([s for s, _ in samples] or [0], dtype=np.int32)
    sk = np.array([k for _, k in samples] or [0], dtype=np.int64)
    sc = np.zeros((max(ns, 1), n))
    sG = np.zeros((max(ns, 1), gcap, n))
    sn = np.zeros(max(ns, 1), dtype=np.int32)
    ssel = np.zeros((max(ns, 1), gcap), dtype=np.int32)
    rho = np.empty((nsplits, int(nsteps)), order="F") if d is not None else None
    dd = np.ascontiguousarray(d, dtype=np.float64) if d is not None else None
    rc = lib.ref_glgm06(n, int(nsteps), nsplits, int(max_order), Phi.ctypes.data_as(c_double_p),
                        c0.ctypes.data_as(c_double_p), G0.ctypes.data_as(c_double_p), p0,
                        cU.ctypes.data_as(c_double_p), GU.ctypes.data_as(c_double_p), pU,
                        dd.ctypes.data_as(c_double_p) if dd is not None else None,
                        rho.ctypes.data_as(c_double_p) if rho is not None else None, ns,
                        ss.ctypes.data_as(c_int32_p), sk.ctypes.data_as(C.POINTER(C.c_int64)),
                        sc.ctypes.data_as(c_double_p), sG.ctypes.data_as(c_double_p), gcap,
                        sn.ctypes.data_as(c_int32_p), ssel.ctypes.data_as(c_int32_p), int(bool(shared_chain)), int(nthreads))
    if rc != 0:
        raise RuntimeError("ref_glgm06 failed (rc=%d)" % rc)
    out = []
    for q in range(ns):
        p = int(sn[q])
        sel = ssel[q, :p].copy()
        out.append((sc[q].copy(), sG[q, :p].T.copy(), None if (p and sel[0] == -2) else sel))
    return rho, out
