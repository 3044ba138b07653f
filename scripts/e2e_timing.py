"""Where the end-to-end time of the one-shot `reach_lgg09` call goes on C5 (REACH_TIMING=1 prints the library's own
phases to stderr): pinned host buffers in, pinned host buffer out."""
import os, sys, time
os.environ["REACH_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
import reach_b200
from reach_b200 import _lib, workloads as W
from reach_b200.support_program import flatten, to_c
ctx = _lib.default_context(0)
wl = W.synthetic_c5()
n, ndirs, nsteps = wl.n, wl.dirs.shape[1], wl.nsteps
Phi_h = torch.empty((n, n), dtype=torch.float64, pin_memory=True).numpy(); Phi_h[...] = wl.Phi.T
dirs_h = torch.empty((ndirs, n), dtype=torch.float64, pin_memory=True).numpy(); dirs_h[...] = wl.dirs.T
rho_h = torch.empty((nsteps, ndirs), dtype=torch.float64, pin_memory=True).numpy()
om, v = to_c(flatten(wl.Omega0, wl.Phi), n), to_c(flatten(wl.V, None), n)
for i in range(4):
    t0 = time.perf_counter()
    ctx.check(ctx.lib.reach_lgg09(ctx.handle, n, ndirs, nsteps, _lib.dptr(Phi_h), _lib.dptr(dirs_h), om.ptr(), v.ptr(), None, None,
                                  _lib.dptr(rho_h), None))
    print("call %d: %.1f ms" % (i, 1e3 * (time.perf_counter() - t0)), flush=True)
