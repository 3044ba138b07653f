"""Exploratory timings on the GPU box (not the bench): DMMA kernel vs cuBLAS DGEMM, and the
LGG09 configs at reduced step counts."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import reach_b200 as rb
from reach_b200 import _lib, workloads as W
from reach_b200.lgg09 import Lgg09Plan

ctx = _lib.default_context(0)
print("device", ctx.name, "SMs", ctx.sm_count)

def cublas(nn, iters=5):
    a = torch.randn(nn, nn, dtype=torch.float64, device="cuda")
    b = torch.randn(nn, nn, dtype=torch.float64, device="cuda")
    torch.matmul(a, b); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        torch.matmul(a, b)
    e1.record(); torch.cuda.synchronize()
    return 2.0 * nn**3 * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12

for nn in (2048, 4096, 8192):
    print("n=%d  cuBLAS DGEMM %.2f TF   reach DMMA %.2f TF" % (nn, cublas(nn), ctx.dgemm_probe(nn, nn, nn, 5)))

def run(name, wl, nsteps, **opts):
    plan = Lgg09Plan(ctx, wl.n, wl.dirs.shape[1], nsteps, **opts)
    plan.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
    plan.run(); plan.run()
    st = plan.stats()
    evals = wl.dirs.shape[1] * nsteps
    flops = 2.0 * wl.n**2 * st["nchains"] * nsteps
    print("%-10s n=%d dirs=%d steps=%d S=%d tile=%dx%d: %.2f ms  %.3e evals/s  %.2f TF (executed)  gemm %.3f ms x %d, launches %d"
          % (name, wl.n, wl.dirs.shape[1], nsteps, st["time_block"], st["tile_m"], st["tile_n"], st["ms"],
             evals / (st["ms"] * 1e-3), flops / (st["ms"] * 1e-3) / 1e12, st["gemm_ms"], st["gemm_launches"], st["launches"]))
    plan.close()

which = sys.argv[1:] or ["c1", "c2", "c3", "c5"]
if "c1" in which:
    wl = W.building_c1()
    for S in (0, 1, 16, 64, 256):
        run("C1", wl, 10000, time_block=S, path="gemm")
    run("C1 chain", wl, 10000, path="chain")
if "c2" in which:
    wl = W.iss_c2()
    for S in (0, 1, 8, 32):
        run("C2", wl, 4096, time_block=S, path="gemm")
    run("C2 chain", wl, 4096, path="chain")
    run("C2 auto full", wl, 33334)
if "c3" in which:
    wl = W.heat_c3()
    for S in (0, 1, 2, 4):
        run("C3", wl, 256, time_block=S)
if "c5" in which:
    wl = W.synthetic_c5()
    for S, tm, tn in ((0, 0, 0), (1, 128, 128), (2, 128, 128), (4, 128, 128), (1, 64, 128), (4, 64, 128)):
        run("C5", wl, 128, time_block=S, tile_m=tm, tile_n=tn)
