"""ctypes binding of libreach_cuda.so (the C ABI in include/reach.h).

There is no fallback: if the shared library is missing or no B200 is visible, calls raise."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libreach_cuda.so")

REACH_OK, REACH_EINVAL, REACH_ECUDA, REACH_ENOMEM, REACH_EUNSUPPORTED, REACH_ESTATE = 0, -1, -2, -3, -4, -5
REACH_TERM_BOX, REACH_TERM_ZONO = 0, 1
REACH_INV_NONE, REACH_INV_HALFSPACE, REACH_INV_BOX, REACH_INV_POLYHEDRON = 0, 1, 2, 3

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class ReachError(RuntimeError):
    """Device / library failure (the Julia wrapper rethrows these as ErrorException)."""


class reach_term(C.Structure):
    _fields_ = [("kind", C.c_int32), ("which", C.c_int32), ("q", C.c_int32), ("p", C.c_int32),
                ("M", c_double_p), ("c", c_double_p), ("r", c_double_p), ("G", c_double_p)]


class reach_setexpr(C.Structure):
    _fields_ = [("nalt", C.c_int32), ("alt_len", c_int32_p), ("terms", C.POINTER(reach_term))]


class reach_invariant(C.Structure):
    _fields_ = [("kind", C.c_int32), ("m", C.c_int32), ("a", c_double_p), ("b", c_double_p)]


class reach_lgg09_opts(C.Structure):
    _fields_ = [("time_block", C.c_int32), ("share_negated", C.c_int32), ("tile_m", C.c_int32),
                ("tile_n", C.c_int32), ("path", C.c_int32), ("reserved", C.c_int32 * 3)]


# name -> (restype, argtypes); every symbol include/reach.h declares
SIGNATURES = {
    "reach_version": (C.c_int, []),
    "reach_device_count": (C.c_int, []),
    "reach_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "reach_destroy": (None, [C.c_void_p]),
    "reach_last_error": (C.c_char_p, [C.c_void_p]),
    "reach_device_name": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "reach_sm_count": (C.c_int, [C.c_void_p]),
    "reach_stream": (C.c_void_p, [C.c_void_p]),
    "reach_synchronize": (C.c_int, [C.c_void_p]),
    "reach_lgg09_default_opts": (None, [C.POINTER(reach_lgg09_opts)]),
    "reach_lgg09": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, c_double_p, c_double_p,
                              C.POINTER(reach_setexpr), C.POINTER(reach_setexpr),
                              C.POINTER(reach_invariant), C.POINTER(reach_lgg09_opts), c_double_p,
                              c_int64_p]),
    "reach_lgg09_plan_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64,
                                          C.POINTER(reach_lgg09_opts), C.POINTER(C.c_void_p)]),
    "reach_lgg09_plan_upload": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.POINTER(reach_setexpr),
                                          C.POINTER(reach_setexpr), C.POINTER(reach_invariant)]),
    "reach_lgg09_plan_set_output": (C.c_int, [C.c_void_p, C.c_void_p]),
    "reach_lgg09_plan_run": (C.c_int, [C.c_void_p]),
    "reach_lgg09_plan_stack_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                              C.POINTER(C.c_int32)]),
    "reach_lgg09_plan_stack_step": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_int64),
                                              C.POINTER(C.c_int64)]),
    "reach_lgg09_plan_stack_ready": (C.c_int, [C.c_void_p]),
    "reach_lgg09_plan_sync": (C.c_int, [C.c_void_p]),
    "reach_lgg09_plan_fetch": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_double_p]),
    "reach_lgg09_plan_nsteps_done": (C.c_int, [C.c_void_p, c_int64_p]),
    "reach_lgg09_plan_max_over_steps": (C.c_int, [C.c_void_p, c_double_p, c_int64_p]),
    "reach_lgg09_plan_device_rho": (C.c_void_p, [C.c_void_p]),
    "reach_lgg09_plan_stats": (C.c_int, [C.c_void_p, c_double_p, c_int64_p, c_int32_p, c_int32_p,
                                         c_int32_p, c_int32_p, c_double_p, c_int64_p]),
    "reach_lgg09_plan_path": (C.c_int, [C.c_void_p, c_int32_p, c_int32_p, c_double_p, c_int64_p, c_int64_p]),
    "reach_lgg09_plan_invariant_info": (C.c_int, [C.c_void_p, c_int32_p, c_int32_p, c_int32_p, c_int64_p]),
    "reach_lgg09_plan_guard": (C.c_int, [C.c_void_p, c_double_p, c_int64_p, c_int64_p]),
    "reach_lgg09_plan_destroy": (None, [C.c_void_p]),
    "reach_lgg09_batch_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, c_double_p, c_double_p,
                                           C.POINTER(reach_setexpr), C.POINTER(reach_setexpr), C.POINTER(C.c_void_p)]),
    "reach_lgg09_batch_run": (C.c_int, [C.c_void_p]),
    "reach_lgg09_batch_sync": (C.c_int, [C.c_void_p]),
    "reach_lgg09_batch_fetch": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, c_double_p]),
    "reach_lgg09_batch_max_over_steps": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "reach_lgg09_batch_stats": (C.c_int, [C.c_void_p, c_double_p, c_int64_p, c_int32_p, c_int32_p, c_int32_p, c_int64_p]),
    "reach_lgg09_batch_device_rho": (C.c_void_p, [C.c_void_p]),
    "reach_lgg09_batch_destroy": (None, [C.c_void_p]),
    "reach_box": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                            C.c_int, C.POINTER(reach_lgg09_opts), c_double_p, c_double_p]),
    "reach_glgm06_plan_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.POINTER(C.c_void_p)]),
    "reach_glgm06_plan_upload": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                           c_double_p]),
    "reach_glgm06_plan_run": (C.c_int, [C.c_void_p]),
    "reach_glgm06_plan_sync": (C.c_int, [C.c_void_p]),
    "reach_glgm06_plan_ngens": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_int)]),
    "reach_glgm06_plan_fetch": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_double_p, c_double_p, c_int32_p]),
    "reach_glgm06_plan_support": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "reach_glgm06_plan_stats": (C.c_int, [C.c_void_p, c_double_p, c_int64_p, c_int64_p]),
    "reach_glgm06_plan_destroy": (None, [C.c_void_p]),
    "reach_glgm06": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, c_double_p, c_double_p,
                               c_double_p, C.c_int, c_double_p, c_double_p, C.c_int, c_double_p,
                               c_double_p, C.c_int, c_int32_p]),
    "reach_multi_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int]),
    "reach_multi_destroy": (None, [C.c_void_p]),
    "reach_multi_ndev": (C.c_int, [C.c_void_p]),
    "reach_multi_ctx": (C.c_void_p, [C.c_void_p, C.c_int]),
    "reach_multi_last_error": (C.c_char_p, [C.c_void_p]),
    "reach_multi_lgg09": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, c_double_p, c_double_p,
                                    C.POINTER(reach_setexpr), C.POINTER(reach_setexpr), C.POINTER(reach_lgg09_opts),
                                    c_double_p, C.POINTER(C.c_void_p), c_double_p]),
    "reach_multi_glgm06": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, c_double_p, c_double_p,
                                     c_double_p, C.c_int, c_double_p, c_double_p, C.c_int, c_double_p,
                                     c_double_p, C.c_int, c_int32_p]),
    "reach_reduce_order_gir05": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p,
                                           C.POINTER(C.c_int), c_int32_p]),
    "reach_discretize_phi": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_double, c_double_p, c_double_p, c_double_p,
                                       c_int32_p]),
    "reach_dgemm_probe": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p]),
    "reach_ozgemm_probe": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p,
                                     c_double_p, C.c_int, c_double_p, c_double_p]),
    "reach_crtgemm_probe": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p,
                                      c_double_p, C.c_int, c_double_p, c_double_p]),
    "reach_crt_constants": (C.c_int, [C.c_int, c_int32_p, c_double_p, c_double_p, c_double_p, c_int32_p, c_int32_p]),
}

_lib = None


def load():
    """Load libreach_cuda.so; raises if it has not been built (`python __graft_entry__.py`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ReachError(
            "libreach_cuda.so is not built (%s missing). Build it with `python -c 'import "
            "__graft_entry__ as g; g.build()'` or `make -C reachabilityanalysis.jl_b200/csrc`. "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dptr(a):
    """double* view of a float64 C- or F-contiguous array (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and (a.flags.c_contiguous or a.flags.f_contiguous)
    return a.ctypes.data_as(c_double_p)


def fmat(a):
    """Column-major float64 copy (what the ABI expects for matrices)."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


class Context:
    """One device + one stream (`reach_ctx`)."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.reach_create(C.byref(h), int(device))
        if rc != REACH_OK:
            raise ReachError("reach_create(device=%d) failed (%d): %s"
                             % (device, rc, self.lib.reach_last_error(None).decode()))
        self.handle = h
        self.device = int(device)

    def check(self, rc):
        if rc == REACH_OK:
            return
        msg = self.lib.reach_last_error(self.handle).decode()
        if rc == REACH_EINVAL:
            raise ValueError(msg)          # ArgumentError on the Julia side
        if rc == REACH_EUNSUPPORTED:
            raise NotImplementedError(msg)
        raise ReachError("libreach_cuda error %d: %s" % (rc, msg))

    @property
    def name(self):
        buf = C.create_string_buffer(256)
        self.check(self.lib.reach_device_name(self.handle, buf, 256))
        return buf.value.decode()

    @property
    def sm_count(self):
        return self.lib.reach_sm_count(self.handle)

    def synchronize(self):
        self.check(self.lib.reach_synchronize(self.handle))

    def dgemm_probe(self, m, n, k, iters=5):
        out = C.c_double()
        self.check(self.lib.reach_dgemm_probe(self.handle, m, n, k, iters, C.byref(out)))
        return out.value

    def ozgemm_probe(self, X, Y, slices=8, iters=1, want_result=True):
        """C = X @ Y.T through the INT8-tcgen05 digit-plane GEMM; returns (C or None, TFLOP/s, slicing ms)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        m, k = X.shape
        n = Y.shape[0]
        assert Y.shape[1] == k
        Cm = np.empty((m, n)) if want_result else None
        tf, sl = C.c_double(), C.c_double()
        self.check(self.lib.reach_ozgemm_probe(self.handle, m, n, k, slices, dptr(X), dptr(Y), dptr(Cm), iters,
                                               C.byref(tf), C.byref(sl)))
        return Cm, tf.value, sl.value

    def crtgemm_probe(self, X, Y, moduli=14, iters=1, want_result=True):
        """C = X @ Y.T through the INT8-tcgen05 residue (CRT) GEMM; returns (C or None, TFLOP/s, residue conversion ms)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        m, k = X.shape
        n = Y.shape[0]
        assert Y.shape[1] == k
        Cm = np.empty((m, n)) if want_result else None
        tf, sl = C.c_double(), C.c_double()
        self.check(self.lib.reach_crtgemm_probe(self.handle, m, n, k, moduli, dptr(X), dptr(Y), dptr(Cm), iters,
                                                C.byref(tf), C.byref(sl)))
        return Cm, tf.value, sl.value

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.reach_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiContext:
    """Several devices behind one handle (`reach_multi`): single process, directions / splits sharded inside the
    library, the LGG09 all-gather done by peer / host 2-D copies (include/reach.h)."""

    def __init__(self, device_ids):
        self.lib = load()
        ids = (C.c_int * len(device_ids))(*[int(d) for d in device_ids])
        h = C.c_void_p()
        rc = self.lib.reach_multi_create(C.byref(h), ids, len(device_ids))
        if rc != REACH_OK:
            raise ReachError("reach_multi_create(%s) failed (%d): %s"
                             % (list(device_ids), rc, self.lib.reach_multi_last_error(None).decode()))
        self.handle = h
        self.device_ids = [int(d) for d in device_ids]

    @property
    def ndev(self):
        return self.lib.reach_multi_ndev(self.handle)

    def check(self, rc):
        if rc == REACH_OK:
            return
        msg = self.lib.reach_multi_last_error(self.handle).decode()
        if rc == REACH_EINVAL:
            raise ValueError(msg)
        if rc == REACH_EUNSUPPORTED:
            raise NotImplementedError(msg)
        raise ReachError("libreach_cuda error %d: %s" % (rc, msg))

    def lgg09(self, Phi, dirs, Omega0_c, V_c, nsteps, opts=None, rho_dev=None, want_host=True):
        """`reach_multi_lgg09`: Omega0_c / V_c are flattened programs (support_program.to_c).  Returns
        (rho [ndirs x nsteps] or None, ms per device)."""
        Phi, dirs = fmat(Phi), fmat(dirs)
        n, ndirs = dirs.shape
        out = np.empty((ndirs, int(nsteps)), order="F") if want_host else None
        ms = np.zeros(self.ndev)
        devp = None
        if rho_dev is not None:
            devp = (C.c_void_p * self.ndev)(*[C.c_void_p(int(p)) if p else None for p in rho_dev])
        self.check(self.lib.reach_multi_lgg09(self.handle, n, ndirs, int(nsteps), dptr(Phi), dptr(dirs), Omega0_c.ptr(),
                                              V_c.ptr() if V_c is not None else None,
                                              C.byref(opts) if opts is not None else None, dptr(out), devp, dptr(ms)))
        return out, ms

    def glgm06(self, Phi, c0, G0, cU, GU, nsteps, max_order, gcap):
        """`reach_multi_glgm06`: c0 [n x nsplits], G0 [n x p0 x nsplits] (Fortran order); returns (c, G, ngens) in the
        one-shot layout of reach_glgm06."""
        Phi = fmat(Phi)
        n = Phi.shape[0]
        c0 = np.asfortranarray(c0, dtype=np.float64)
        G0 = np.asfortranarray(G0, dtype=np.float64)
        nsplits, p0 = c0.shape[1], G0.shape[1]
        pU = -1 if GU is None else np.asarray(GU).shape[1]
        cUa = None if cU is None else np.ascontiguousarray(cU, dtype=np.float64)
        GUa = None if GU is None else fmat(GU)
        c = np.empty((n, nsplits, int(nsteps)), order="F")
        G = np.zeros((n, gcap, nsplits, int(nsteps)), order="F")
        ng = np.zeros((nsplits, int(nsteps)), dtype=np.int32, order="F")
        self.check(self.lib.reach_multi_glgm06(self.handle, n, int(nsteps), nsplits, int(max_order), dptr(Phi), dptr(c0),
                                               dptr(G0), p0, dptr(cUa), dptr(GUa), pU, dptr(c), dptr(G), int(gcap),
                                               ng.ctypes.data_as(c_int32_p)))
        return c, G, ng

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.reach_multi_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=0):
    ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx
