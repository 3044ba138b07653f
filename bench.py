#!/usr/bin/env python
"""bench.py -- LGG09 support evaluations/s on BASELINE.json config C5
(synthetic stable LTI n=2000, box template = 4000 directions, 10^4 steps), sharded by direction
over N GPUs of one node.  One "step" = one full pass of the hot path (all 10^4 time steps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  See DESIGN.md §measurement for what each key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# torchrun exports OMP_NUM_THREADS=1; rank 0 also runs the CPU baseline, which must use the host's cores
if os.environ.get("RANK", "0") == "0":
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    os.environ.pop("OPENBLAS_NUM_THREADS", None)

import numpy as np  # noqa: E402  (after the thread-count fix above)

METRIC = "LGG09 support evals/s (dirs x steps)"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=2000, help="state dimension of the synthetic workload")
    ap.add_argument("--nsteps", type=int, default=10000, help="time steps per pass")
    ap.add_argument("--time-block", type=int, default=0)
    ap.add_argument("--no-dist-stack", action="store_true",
                    help="N > 1: every rank builds the whole row stack itself (default: the ranks share the build)")
    ap.add_argument("--path", default="auto", choices=["auto", "gemm", "int8"],
                    help="main-loop kernel: auto (INT8 digit-plane GEMM guarded by DMMA passes), gemm (FP64 DMMA), int8 (unguarded)")
    ap.add_argument("--cpu-sample-steps", type=int, default=0, help="time steps of the CPU sample (0: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-glgm06", action="store_true", help="skip the GLGM06 C4 flowpipe wall-time leg")
    ap.add_argument("--glgm06-steps", type=int, default=500)
    ap.add_argument("--no-fp64-leg", action="store_true",
                    help="skip the FP64-tensor-pipe pass of C5 and the full-run INT8-vs-DMMA comparison (N = 1 only)")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short C1 / C2 / split-caller legs (rank 0, N = 1 only)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def cpu_lgg09_arm(wl, budget_s=12.0, min_steps=100, sparse=False, repeats=1, want_output=True):
    """The reference's CPU algorithm for the LGG09 loop, as the C/OpenMP restatement oracle/c/reach_ref.c (Julia is
    not installed), on all host cores, in BOTH shapes the reference has:
      B1  per-direction chains spread over the cores   (reach_inhomog.jl:35-64; sparse Phi as LGG09/post.jl:25-27)
      B2  one GEMM per time step over the whole direction block + fused support epilogue (reach_inhomog.jl:174-218)
    each with the l / -l sharing the reference only notes (reach_homog.jl:108-110) -- a stronger baseline than the
    literal loops.  Both are calibrated on a few steps, the faster one is timed on a bounded sample (first `sample`
    time steps of the same workload, all directions; the cost per step is constant)."""
    from oracle import cref
    ndirs = wl.dirs.shape[1]
    share = True
    try:                                       # sharing needs a template of the form [D, -D]
        cref.reach_lgg09(wl.dirs, wl.Omega0, wl.Phi, 1, wl.V, shape=1, share=True)
    except RuntimeError:
        share = False
    cands = [("B1 per-direction chains over the cores" + (", sparse Phi" if sparse else ""), dict(shape=1, sparse=sparse, share=share)),
             ("B2 GEMM per step (OpenBLAS) + fused support epilogue", dict(shape=2, share=share))]
    calib = {}
    for name, kw in cands:
        k = 4
        while True:
            t0 = time.perf_counter()
            cref.reach_lgg09(wl.dirs, wl.Omega0, wl.Phi, k, wl.V, **kw)
            dt = time.perf_counter() - t0
            if dt > 0.25 or k >= wl.nsteps:
                break
            k = min(wl.nsteps, k * 4)
        calib[name] = dt / k
    best = min(calib, key=calib.get)
    kw = dict(cands)[best]
    sample = int(max(min(min_steps, wl.nsteps), min(wl.nsteps, budget_s / calib[best])))
    total, out = 0.0, None
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = cref.reach_lgg09(wl.dirs, wl.Omega0, wl.Phi, sample, wl.V, **kw)
        total += time.perf_counter() - t0
    # the same shape without the l / -l sharing: what the reference's loops literally do
    lit_rate = None
    if share:
        lit = dict(kw, share=False)
        k = max(2, min(sample, int(2.0 / calib[best])))
        t0 = time.perf_counter()
        cref.reach_lgg09(wl.dirs, wl.Omega0, wl.Phi, k, wl.V, **lit)
        lit_rate = ndirs * k / (time.perf_counter() - t0)
    desc = "first %d of %d time steps, all %d directions" % (sample, wl.nsteps, ndirs)
    info = {"value": ndirs * sample * repeats / total, "unit": UNIT, "cores": cref.num_threads(), "kind": "port",
            "sample": desc, "shape": best,
            "shapes_calibrated_evals_per_s": {k_: ndirs / v for k_, v in calib.items()},
            "literal_unshared_value": lit_rate,
            "note": "C/OpenMP restatement of the reference loops (oracle/c/reach_ref.c, gcc -O3 -march=native -fopenmp, "
                    "OpenBLAS dgemm for B2); the faster of the two shapes; l / -l chains %s" % ("shared" if share else "not shared (template is not of the form [D, -D])")}
    return info, (out if want_output else None), sample, total


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from reach_b200 import workloads as W
    wl = W.synthetic_c5(n=args.n, nsteps=args.nsteps)
    ndirs = wl.dirs.shape[1]
    if args.cpu_sample_steps:
        budget, min_steps = 1e-9, args.cpu_sample_steps
    else:
        budget, min_steps = 12.0, 100
    for _ in range(min(args.warmup, 1)):
        cpu_lgg09_arm(wl, budget_s=1.0, min_steps=4, want_output=False)
    info, _, sample, total = cpu_lgg09_arm(wl, budget_s=budget, min_steps=min_steps, repeats=args.steps, want_output=False)
    value = info["value"]
    desc = info["sample"] + ", per bench step"
    info["sample"] = desc
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": wl.name, "n": wl.n, "ndirs": ndirs, "nsteps": args.nsteps,
                   "sample": desc},
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            for i, nm in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith("active"):
                    reasons.add(nm)
        # take the upper half (samples under load) for the median
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cublas_fp64_peak(torch, nn=8192, iters=6):
    """FP64 roofline denominator: cuBLAS DGEMM nn^3 measured in this run (MEASURED_PEAKS.json
    holds HBM and bf16 only)."""
    a = torch.randn(nn, nn, dtype=torch.float64, device="cuda")
    b = torch.randn(nn, nn, dtype=torch.float64, device="cuda")
    for _ in range(2):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * nn ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    torch.cuda.empty_cache()
    return best


def cublaslt_int8_peak(torch, nn=8192, burst_iters=6, sustained_s=2.0):
    """INT8 tensor roofline denominator: cuBLASLt INT8 GEMM nn^3 (torch._int_mm) measured in this run, as a burst
    (best single launch) and sustained (back to back for ~sustained_s seconds, the regime of a kernel timed inside a
    seconds-long pass).  MEASURED_PEAKS.json holds HBM and bf16 only."""
    a = torch.randint(-64, 64, (nn, nn), dtype=torch.int8, device="cuda")
    b = torch.randint(-64, 64, (nn, nn), dtype=torch.int8, device="cuda")
    for _ in range(2):
        torch._int_mm(a, b)
    torch.cuda.synchronize()
    ops = 2.0 * nn ** 3
    best = 0.0
    for _ in range(burst_iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch._int_mm(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, ops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    iters = max(10, int(sustained_s / (ops / (best * 1e12))))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        torch._int_mm(a, b)
    e1.record()
    torch.cuda.synchronize()
    sustained = ops * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b
    torch.cuda.empty_cache()
    return best, sustained


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def scaled_diff(got, ref):
    scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=1, keepdims=True))
    scale[scale == 0] = 1.0
    return float((np.abs(got - ref) / scale).max())


def glgm06_leg(args, ctx, torch, dist, world, rank, peak_tf):
    """Second half of the metric: GLGM06 flowpipe wall-time on config C4 (Building, 1024 initial-set
    splits, max_order 10, Forward(:zono) model), splits sharded over the ranks, no collective."""
    from reach_b200 import workloads as W
    from reach_b200.glgm06 import _run_batch, GLGM06, Glgm06Plan
    from reach_b200.zonotope_ops import to_zonotope
    nsteps = args.glgm06_steps
    Phi, Oms, V = W.building_c4(2, nsteps)
    from reach_b200.sharding import shard_splits
    own = shard_splits(len(Oms), world, rank)
    mine = Oms[own.start:own.stop]
    plan = _run_batch(GLGM06(δ=0.002, max_order=10), Phi, mine, V, nsteps, ctx)       # warm-up run
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    for _ in range(2):
        plan.run()
    ms = []
    for _ in range(3):
        flush.zero_()
        if world > 1:
            dist.barrier()
        plan.run()
        ms.append(plan.stats()["ms"])
    t = sum(ms) / len(ms)
    if world > 1:
        tt = torch.tensor([t], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt.item())
    st = plan.stats()
    n = Phi.shape[0]
    p0r = min(plan.p0, 10 * n)
    # algorithmic work of the dominant kernel (the batched product P_k [G0 | c0] of every split and step)
    flops = 2.0 * n * n * (p0r + 1) * len(mine) * nsteps
    out = {"workload": "C4 Building GLGM06 max_order=10, Forward(setops=:zono), %d splits x %d steps" % (len(Oms), nsteps),
           "flowpipe_wall_ms": t, "zonotopes_per_s": len(Oms) * nsteps / (t * 1e-3), "splits_per_gpu": len(mine),
           "generators_per_zonotope": 480, "launches": st["launches"], "bytes_stored_per_gpu": st["bytes_stored"],
           "timer": "CUDA events around the run on the library stream, mean of 3 after 3 warm-up runs, 256 MB L2 flush "
                    "between runs, max over ranks",
           "roofline": {"bound": "tensor", "kernel": "nt_dgemm_kernel<...,GenEpi> (FP64 DMMA: P_k [G0|c0] of all splits and steps, "
                                                    "fused per-generator 1-norm / inf-norm)",
                        "achieved": flops / (t * 1e-3) / 1e12, "peak": peak_tf or None, "unit": "TFLOP/s",
                        "frac": (flops / (t * 1e-3) / 1e12 / peak_tf) if peak_tf else None,
                        "flops_per_run": flops,
                        "note": "whole-run time (product + shared chain + selection) against the product's algorithmic flops; "
                                "peak = cuBLAS DGEMM 8192^3 measured in this run"}}
    # e2e: host arrays in (upload), run, and the user-visible result out: rho(d, Z) of every stored zonotope for one
    # direction (device reduction, 8 B per zonotope D2H) + one materialised zonotope -- the flowpipe itself stays in HBM
    zs = [to_zonotope(Om) for Om in mine]
    cU, GU = to_zonotope(V)
    c0 = np.stack([c for c, _ in zs])
    G0 = np.stack([np.ascontiguousarray(G.T) for _, G in zs])
    GUt = np.ascontiguousarray(GU.T)
    d = np.zeros(n)
    d[24] = 1.0
    e2e = []
    for it in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        p2 = Glgm06Plan(ctx, n, nsteps, len(mine), 10, G0.shape[1], GUt.shape[0])
        p2.upload(Phi, c0, G0, cU, GUt)
        p2.run()
        sup = p2.support(d)
        p2.fetch(0, nsteps - 1)
        dt = time.perf_counter() - t0
        p2.close()
        if it:
            e2e.append(dt)
    te = sum(e2e) / len(e2e)
    if world > 1:
        tt = torch.tensor([te], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te = float(tt.item())
    out["e2e"] = {"flowpipe_wall_ms": te * 1e3, "zonotopes_per_s": len(Oms) * nsteps / te,
                  "h2d_bytes_per_step": int(Phi.nbytes + c0.nbytes + G0.nbytes + GUt.nbytes + cU.nbytes),
                  "d2h_bytes_per_step": int(sup.nbytes + 8 * n * 481),
                  "timer": "host wall clock: plan create + upload from host arrays + run + rho(e_25, Z) of all zonotopes to the "
                           "host + one zonotope fetched; mean of 2 after 1 warm-up, max over ranks"}
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import cref, lazysets_oracle as LZ
        ns = min(len(Oms), 128)
        zr = [LZ.reduce_order_gir05(*LZ.to_zonotope(Om), 10) for Om in Oms[:ns]]
        cc = np.stack([z[0] for z in zr], axis=1)
        best = None
        for shared in (True, False):
            t0 = time.perf_counter()
            rho, _ = cref.reach_inhomog_glgm06(Phi, cc, [z[1] for z in zr], nsteps, 10, cU, GU, d=d, shared_chain=shared)
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, shared, rho)
        out["cpu_baseline"] = {"zonotopes_per_s": ns * nsteps / best[0], "flowpipe_wall_ms_extrapolated": best[0] * 1e3 * len(Oms) / ns,
                               "cores": cref.num_threads(), "kind": "port",
                               "shape": "splits over the cores, input chain W_k %s" % ("computed once and shared" if best[1] else "recomputed per split (the reference's literal loop)"),
                               "sample": "%d of %d splits x %d steps (C/OpenMP restatement of reach_inhomog_GLGM06!, "
                                         "oracle/c/reach_ref.c)" % (ns, len(Oms), nsteps)}
        if own.start == 0:
            got = plan.support(d)[:ns]
            out["cpu_baseline"]["max_rel_diff_rho_vs_gpu"] = float(np.abs(got[:min(ns, len(mine))] - best[2][:min(ns, len(mine))]).max() / np.abs(best[2]).max())
    plan.close()
    return out


def lgg09_host_call(ctx, wl, torch, **optkw):
    """The one-shot C-ABI call reach_lgg09 with pinned HOST buffers for workload wl: returns (call, rho_h, h2d, d2h)."""
    import ctypes as C
    from reach_b200 import _lib
    from reach_b200.support_program import flatten, to_c
    n, ndirs, nsteps = wl.n, wl.dirs.shape[1], wl.nsteps
    Phi_h = torch.empty((n, n), dtype=torch.float64, pin_memory=True).numpy()
    Phi_h[...] = wl.Phi.T                                                              # buffer holds Phi col-major
    dirs_h = torch.empty((ndirs, n), dtype=torch.float64, pin_memory=True).numpy()
    dirs_h[...] = wl.dirs.T
    rho_h = torch.empty((nsteps, ndirs), dtype=torch.float64, pin_memory=True).numpy()
    om = to_c(flatten(wl.Omega0, wl.Phi), n)
    v = to_c(flatten(wl.V, None), n) if wl.V is not None else None
    opts = _lib.reach_lgg09_opts()
    ctx.lib.reach_lgg09_default_opts(C.byref(opts))
    opts.time_block = optkw.get("time_block", 0)
    opts.path = {"auto": 0, "gemm": 1, "int8": 3}[optkw.get("path", "auto")]
    keep = (Phi_h, dirs_h, om, v, opts)

    def call():
        ctx.check(ctx.lib.reach_lgg09(ctx.handle, n, ndirs, nsteps, _lib.dptr(Phi_h), _lib.dptr(dirs_h), om.ptr(),
                                      v.ptr() if v is not None else None, None, C.byref(opts), _lib.dptr(rho_h), None))
        return keep
    return call, rho_h, Phi_h.nbytes + dirs_h.nbytes + 4 * n * 8, rho_h.nbytes


def other_configs_leg(ctx, torch, peak_tf, no_cpu):
    """The other BASELINE configs, each with its own device-resident time, e2e (one-shot reach_lgg09 with pinned host
    buffers), roofline and CPU baseline: C1 Building (DMMA GEMM over the power stack), C2 ISS (compact chains: 135
    decoupled 2x2 blocks), C3 HEAT02 (n = 1000, 2001 directions, all 10^4 steps; the INT8 guard sends it to the DMMA
    kernel at pass 0); plus LGG09 over 256 initial-set splits of Building (one batch)."""
    import reach_b200 as rb
    from reach_b200 import workloads as W, discretization as D, sets as S
    from reach_b200.lgg09 import Lgg09Plan, Lgg09BatchPlan
    hbm = measured_peaks().get("hbm_gbs")
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    out = {}
    for key, wl, sparse in (("C1", W.building_c1(), False), ("C2", W.iss_c2(), True), ("C3", W.heat_c3(), False)):
        ndirs = wl.dirs.shape[1]
        plan = Lgg09Plan(ctx, wl.n, ndirs, wl.nsteps)
        plan.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
        for _ in range(3):
            plan.run()
        ms, gms = [], []
        for _ in range(3):
            flush.zero_()
            plan.run()
            ms.append(plan.stats()["ms"])
            gms.append(plan.stats()["gemm_ms"])
        st = plan.stats()
        t = sum(ms) / len(ms)
        ent = {"workload": wl.name, "ndirs": ndirs, "nsteps": wl.nsteps, "ms": t, "ms_best": min(ms),
               "value": ndirs * wl.nsteps / (t * 1e-3), "unit": UNIT, "path": st["path"],
               "launches": st["launches"], "time_block": st["time_block"],
               "timer": "CUDA events on the library stream, mean of 3 after 3 warm-up runs, 256 MB L2 flush between runs"}
        if st["path"] == "int8":      # guarded: fallback_pass >= 0 means the run was redone on the FP64 (DMMA) kernel
            ent.update(guard_checks=st["guard_checks"], guard_max=st["guard_max"], fallback_pass=st["fallback_pass"],
                       replay_from=st["replay_from"])
        flops = 2.0 * wl.n * wl.n * st["nchains"] * wl.nsteps
        if st["path"] == "compact":
            bytes_ = 8.0 * ndirs * wl.nsteps
            ent["roofline"] = {"bound": "hbm", "kernel": "lgg09_compact_kernel (lane group per (chain, time block))",
                               "achieved": bytes_ / (t * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                               "frac": (bytes_ / (t * 1e-3) / 1e9 / hbm) if hbm else None,
                               "algorithmic_bytes": bytes_, "traffic": None,
                               "note": "algorithmic bytes = rho written once (8 B per evaluation); the kernel is "
                                       "instruction-issue bound (ncu: profiles/r1s3_compact_and_split_batch_ncu.md), "
                                       "peak = MEASURED_PEAKS.json hbm_gbs"}
        else:
            g = sum(gms) / len(gms)
            fl = 2.0 * wl.n * wl.n * st["nchains"] * st["time_block"]
            ent["roofline"] = {"bound": "tensor", "kernel": "nt_dgemm(_ws)_kernel<...,Lgg09Epi> (FP64 DMMA + fused support epilogue)",
                               "achieved": flops / (t * 1e-3) / 1e12, "peak": peak_tf or None, "unit": "TFLOP/s",
                               "frac": (flops / (t * 1e-3) / 1e12 / peak_tf) if peak_tf else None,
                               "main_launch": {"achieved": fl / (g * 1e-3) / 1e12 if g > 0 else None, "launch_ms": g,
                                               "launches": st["gemm_launches"],
                                               "share_of_run": st["gemm_launches"] * g / t if t > 0 else None},
                               "traffic": None,
                               "note": "achieved = algorithmic 2 n^2 flop per chain-step over the WHOLE run time (stack build, "
                                       "running sums and evaluation included); main_launch = the main-loop GEMM alone; "
                                       "peak = cuBLAS DGEMM 8192^3 measured in this run"}
        got = plan.fetch(0, wl.nsteps)
        plan.close()
        call, rho_h, h2d, d2h = lgg09_host_call(ctx, wl, torch)
        call()
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            call()
            ts.append(time.perf_counter() - t0)
        te = sum(ts) / len(ts)
        ent["e2e"] = {"value": ndirs * wl.nsteps / te, "unit": UNIT, "ms": te * 1e3, "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h, "timer": "host wall clock around reach_lgg09(host buffers), mean of 3",
                      "bit_identical_to_device_resident_run": bool(np.array_equal(rho_h.T, got))}
        if not no_cpu:
            info, ref, sample, _ = cpu_lgg09_arm(wl, budget_s=6.0, min_steps=100, sparse=sparse)
            info["max_scaled_diff_vs_gpu"] = scaled_diff(got[:, :sample], ref)
            ent["cpu_baseline"] = info
        out[key] = ent
    # early termination (reach_inhomog.jl:198-215): Building with an invariant x_i <= b that is first violated near the
    # end of the transient (the latest step at which any lower bound of the box hull peaks); time saved = full run - this
    try:
        wl = W.building_c1()
        n = wl.n
        full = Lgg09Plan(ctx, n, 2 * n, wl.nsteps)
        full.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
        full.run()
        ref = full.sfmat()
        lo = -ref[n:, :]
        i = int(lo.argmax(axis=1).argmax())
        kpk = int(lo[i].argmax())
        b = float(0.5 * (lo[i, :max(1, kpk - 30)].max() + lo[i, kpk]))
        e = np.zeros(n)
        e[i] = 1.0
        res = {}
        for name, S_ in (("auto_time_block", 0), ("time_block_64", 64)):
            f2 = Lgg09Plan(ctx, n, 2 * n, wl.nsteps, time_block=S_)
            f2.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
            inv = Lgg09Plan(ctx, n, 2 * n, wl.nsteps, time_block=S_)
            inv.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V, S.HalfSpace(e, b))
            tf, ti = [], []
            for _ in range(4):
                f2.run()
                tf.append(f2.stats()["ms"])
                inv.run()
                ti.append(inv.stats()["ms"])
            info = inv.invariant_info()
            res[name] = {"time_block": inv.stats()["time_block"], "full_run_ms": min(tf[1:]), "with_invariant_ms": min(ti[1:]),
                         "nsteps_done": inv.nsteps_done, "steps_computed": info["steps_computed"], "exact_test": info["exact"],
                         "launches": inv.stats()["launches"], "full_run_launches": f2.stats()["launches"]}
            f2.close()
            inv.close()
        out["C1_invariant_early_exit"] = {"workload": "Building BLDF01 LGG09 box template, invariant x_%d <= %.6g (first violated at step %d of %d)"
                                          % (i + 1, b, res["time_block_64"]["nsteps_done"], wl.nsteps), **res,
                                          "note": "the pass loop stops launching once a step is certified disjoint (host polls the flag two passes behind the device)"}
        full.close()
    except Exception as ex:                      # a diagnostic leg must not take the bench line down
        out["C1_invariant_early_exit"] = {"error": repr(ex)}
    A, B, X0, U0 = W.building_sets()
    U = D.normalize_input(B, U0)
    ivps = [D.forward(A, X, 0.002, U) for X in S.split(X0, [2] * 8 + [1] * 40)]
    nsteps = 10000
    batch = Lgg09BatchPlan(ctx, ivps[0].Phi, rb.BoxDirections(48).matrix(), [iv.Omega0 for iv in ivps], ivps[0].V, nsteps)
    ms = []
    for _ in range(3):
        batch.run()
        ms.append(batch.stats()["ms"])
    t = min(ms[1:])
    out["C1_splits"] = {"workload": "Building BLDF01 LGG09 box template, X0 split into 256 boxes (solve.jl:84-117), one batch",
                        "nsplits": len(ivps), "ndirs": 96, "nsteps": nsteps, "ms": t,
                        "evals_per_s": len(ivps) * 96 * nsteps / (t * 1e-3), "launches": batch.stats()["launches"],
                        "per_split_plan_ms": out["C1"]["ms"]}
    batch.close()
    return out


def fp64_path_leg(args, ctx, torch, wl, out_T, peak_tf):
    """The same C5 pass on the FP64 tensor pipe only (path = gemm, DMMA kernel): its evals/s and roofline, and the
    difference between the two paths over EVERY support value of the run (the INT8 result is in out_T)."""
    from reach_b200.lgg09 import Lgg09Plan
    n, ndirs, nsteps = wl.n, wl.dirs.shape[1], wl.nsteps
    ref_T = torch.empty((nsteps, ndirs), dtype=torch.float64, device="cuda")
    plan = Lgg09Plan(ctx, n, ndirs, nsteps, time_block=args.time_block, path="gemm")
    plan.set_output(ref_T.data_ptr())
    plan.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
    plan.run()                                   # warm-up pass
    plan.run()
    st = plan.stats()
    flops = 2.0 * n * n * st["nchains"] * st["time_block"]
    ach = flops / (st["gemm_ms"] * 1e-3) / 1e12 if st["gemm_ms"] > 0 else None
    # scaled difference |a - b| / max(|b|, max_k |b[j, k]|) over all ndirs x nsteps values, on the device
    rowscale = ref_T.abs().amax(dim=0, keepdim=True)
    denom = torch.maximum(ref_T.abs(), rowscale)
    denom[denom == 0] = 1.0
    diff = float(((out_T - ref_T).abs() / denom).max().item())
    leg = {"value": ndirs * nsteps / (st["ms"] * 1e-3), "unit": UNIT, "ms_per_step": st["ms"], "passes_timed": 1,
           "main_kernel": st["path"], "time_block": st["time_block"],
           "roofline": {"bound": "tensor", "kernel": "nt_dgemm_ws_kernel<...,Lgg09PackedEpi> (FP64 DMMA + fused support epilogue)",
                        "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": (ach / peak_tf) if ach and peak_tf else None,
                        "launch_ms": st["gemm_ms"], "launches_per_step": st["gemm_launches"],
                        "gemm_share_of_step": st["gemm_launches"] * st["gemm_ms"] / st["ms"],
                        "peak_source": "cuBLAS DGEMM 8192^3 measured in this run"}}
    plan.close()
    del ref_T, denom, rowscale
    torch.cuda.empty_cache()
    return leg, {"max_scaled_diff": diff, "values_compared": ndirs * nsteps, "tolerance_north_star": 1e-10,
                 "scale": "|a-b| / max(|b|, max_k |b[j,k]|), b = FP64 DMMA path"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import reach_b200  # noqa: F401
    from reach_b200 import _lib, workloads as W
    from reach_b200.lgg09 import Lgg09Plan
    from reach_b200.sharding import shard_directions, gather_support_values, build_stack_distributed

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with --nproc-per-node %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = _lib.Context(local_rank)
    wl = W.synthetic_c5(n=args.n, nsteps=args.nsteps)
    if world > 1:
        # one problem for all ranks: expm on the hosts of two ranks can differ in the last bit (BLAS threading), and the ranks
        # build ONE row stack together -- every rank must hold rank 0's Phi, Omega0 and V
        box = [wl if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        wl = box[0]
    n, ndirs, nsteps = wl.n, wl.dirs.shape[1], args.nsteps
    local_dirs, rows_by_rank, mloc = shard_directions(wl.dirs, world, rank)

    # device-resident inputs and output
    out_T = torch.empty((nsteps, mloc), dtype=torch.float64, device="cuda")     # == mloc x nsteps col-major
    plan = Lgg09Plan(ctx, n, mloc, nsteps, time_block=args.time_block, path=args.path)
    plan.set_output(out_T.data_ptr())
    plan.upload(wl.Phi, local_dirs, wl.Omega0, wl.V)
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    dist_stack = world > 1 and not args.no_dist_stack
    if dist_stack:
        try:
            plan.stack_info()                     # GEMM paths only (a chain-kernel plan has no row stack)
        except NotImplementedError:
            dist_stack = False

    def one_pass():
        """Whole hot path, inputs resident in HBM: recurrence on this rank's chains + all-gather."""
        flush.zero_()
        barrier()
        build_ms = 0.0
        if dist_stack:
            # the row stack is built once per run by all ranks together (row ranges of every doubling step + in-place
            # all-gathers over NVLink) instead of once per rank; its time is part of the step
            t0 = time.perf_counter()
            build_stack_distributed(plan)
            torch.cuda.synchronize()
            build_ms = (time.perf_counter() - t0) * 1e3
        plan.run(sync=True)
        st = plan.stats()
        ms = st["ms"] + build_ms
        st["stack_build_ms"] = build_ms
        full = None
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            full = gather_support_values(out_T, rows_by_rank, ndirs)
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        return ms, st, full

    peak_tf = cublas_fp64_peak(torch) if rank == 0 else 0.0
    int8_burst, int8_sustained = cublaslt_int8_peak(torch) if rank == 0 else (0.0, 0.0)
    for _ in range(args.warmup):
        one_pass()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    total_ms, launches, gemm_ms_acc, st = 0.0, 0, 0.0, None
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        ms, st, _ = one_pass()
        total_ms += ms
        launches += st["launches"]
        gemm_ms_acc += st["gemm_ms"]
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = ndirs * nsteps * args.steps / (total_ms * 1e-3)

    # ---- e2e: the one-shot C-ABI call with HOST buffers (pinned), H2D + compute + D2H (+ gather) ----
    import ctypes as C
    from reach_b200.support_program import flatten, to_c
    Phi_h = torch.empty((n, n), dtype=torch.float64, pin_memory=True).numpy()          # column-major view below
    Phi_h[...] = wl.Phi.T                                                              # buffer holds Phi col-major
    dirs_h = torch.empty((mloc, n), dtype=torch.float64, pin_memory=True).numpy()
    dirs_h[...] = local_dirs.T
    rho_h = torch.empty((nsteps, mloc), dtype=torch.float64, pin_memory=True).numpy()
    om, v = to_c(flatten(wl.Omega0, wl.Phi), n), to_c(flatten(wl.V, None), n)
    opts = _lib.reach_lgg09_opts()
    ctx.lib.reach_lgg09_default_opts(C.byref(opts))
    opts.time_block = args.time_block
    opts.path = {"auto": 0, "gemm": 1, "int8": 3}[args.path]
    h2d = Phi_h.nbytes + dirs_h.nbytes + 4 * n * 8
    d2h = rho_h.nbytes
    full_h = torch.empty((nsteps, ndirs), dtype=torch.float64, pin_memory=True) if (world > 1 and rank == 0) else None

    def e2e_pass():
        barrier()
        t0 = time.perf_counter()
        if dist_stack:
            # the same C-ABI calls the one-shot reach_lgg09 makes (plan create, upload from host buffers, run, fetch to the
            # host, destroy), with the row stack built by the ranks together in between (reach_lgg09_plan_stack_*)
            p2 = Lgg09Plan(ctx, n, mloc, nsteps, time_block=args.time_block, path=args.path)
            ctx.check(ctx.lib.reach_lgg09_plan_upload(p2.handle, _lib.dptr(Phi_h), _lib.dptr(dirs_h), om.ptr(), v.ptr(), None))
            build_stack_distributed(p2)
            p2.run(sync=True)
            ctx.check(ctx.lib.reach_lgg09_plan_fetch(p2.handle, 0, nsteps, _lib.dptr(rho_h)))
            p2.close()
        else:
            ctx.check(ctx.lib.reach_lgg09(ctx.handle, n, mloc, nsteps, _lib.dptr(Phi_h), _lib.dptr(dirs_h), om.ptr(),
                                          v.ptr(), None, C.byref(opts), _lib.dptr(rho_h), None))
        if world > 1:
            # the only exchange: all-gather of the per-rank blocks; rank 0 lands the full matrix on its host
            loc = torch.from_numpy(rho_h).cuda(non_blocking=True)
            full = gather_support_values(loc, rows_by_rank, ndirs)
            if full_h is not None:
                full_h.copy_(full, non_blocking=True)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    e2e_pass()
    e2e_t = [e2e_pass() for _ in range(max(1, args.steps))]
    e2e_s = sum(e2e_t)
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = ndirs * nsteps * len(e2e_t) / e2e_s
    # e2e result must be the same numbers as the device-resident run
    same = bool(np.array_equal(rho_h, out_T.cpu().numpy()))
    e2e_maxdiff = float(np.abs(rho_h - out_T.cpu().numpy()).max())

    glgm06 = None if args.no_glgm06 else glgm06_leg(args, ctx, torch, dist, world, rank, peak_tf)

    if rank == 0:
        S = st["time_block"]
        gemm_ms = gemm_ms_acc / args.steps                                  # mean duration of one launch
        # algorithmic: 2 n^2 per chain-step; a launch covers S steps of every chain, or (pass groups, few chains per GPU)
        # S steps of several consecutive passes at once -- so: all chain-steps of the run / main-loop launches of the run
        steps_covered = -(-nsteps // S) * S
        flops_per_launch = 2.0 * n * n * st["nchains"] * steps_covered / max(1, st["gemm_launches"])
        achieved = flops_per_launch / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None   # FP64-equivalent TFLOP/s
        traffic, ncu, l2sm = None, {}, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "lgg09_main_kernel_traffic.json")))
            ent = tj.get(st["path"])
            nscale = 1.0
            if st["path"] == "int8" and ent and ent.get("moduli", ent.get("digit_planes")) != st.get("digit_planes"):
                if ent.get("moduli") and st.get("digit_planes", 0) >= 10:
                    nscale = st["digit_planes"] / float(ent["moduli"])     # same kernel, other modulus count: bytes scale with it
                else:
                    ent = tj.get("int8_digit_planes")        # the capture must be of the kernel that ran
                    if ent and ent.get("digit_planes") != st.get("digit_planes"):
                        ent = None
            if ent and ent.get("n") == n and ent.get("chains") == st["nchains"]:
                # DRAM bytes of one launch from the ncu --set full capture; only rescaled if the capture's time block differs
                traffic = ent["dram_bytes_per_launch"] * (S / float(ent["time_block"])) * nscale
                if ent.get("l2_to_sm_bytes_per_launch") and tj.get("l2_to_sm_peak"):
                    l2sm = {"bytes_per_launch": ent["l2_to_sm_bytes_per_launch"] * (S / float(ent["time_block"])) * nscale,
                            "peak_tbs": tj["l2_to_sm_peak"]["tbs"], "peak_source": tj["l2_to_sm_peak"]["source"]}
                ncu = {"time_block_of_capture": ent["time_block"], "moduli_of_capture": ent.get("moduli"), "source": ent.get("source"),
                       "pipe_tensor_active_pct": ent.get("pipe_tensor_active_pct"),
                       "tc_shared_operand_feed_pct": ent.get("tc_shared_operand_feed_pct")}
        except Exception:
            pass
        peaks = measured_peaks()
        share = st["gemm_launches"] * gemm_ms / (total_ms / args.steps)
        if st["path"] == "int8":
            T = st["digit_planes"]
            # digit planes (4..8): T(T+1)/2 plane pairs per FP64 product; residue path (10..16): one INT8 product per modulus
            crt = T >= 10
            pairs = T if crt else T * (T + 1) // 2
            int8_ops = flops_per_launch * pairs                              # INT8 MMA work actually issued
            ach_int8 = int8_ops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
            bf16 = peaks.get("bf16_tflops_sustained")
            peak_int8 = 2.0 * bf16 if bf16 else int8_sustained              # dense INT8 tensor rate = 2 x dense bf16
            roofline = {
                "bound": "tensor",
                "kernel": ("crt_gemm_kernel<OzLgg09Epi> (INT8 tcgen05 2-CTA GEMM per modulus, %d coprime moduli, FP64 by Chinese "
                           "remainder reconstruction, fused support epilogue)" % T) if crt else
                          ("oz_gemm_kernel<%d, OzLgg09Epi> (INT8 tcgen05 digit-plane GEMM, FP64 by error-free slicing, "
                           "fused support epilogue)" % T),
                "achieved": ach_int8, "peak": peak_int8, "unit": "TOP/s (INT8 tensor)",
                "frac": (ach_int8 / peak_int8) if ach_int8 and peak_int8 else None,
                "peak_source": ("2 x MEASURED_PEAKS.json bf16_tflops_sustained (%.1f): dense INT8 tcgen05 rate is twice the bf16 "
                                "rate; sustained figure because the kernel is timed inside a seconds-long pass" % bf16) if bf16 else
                               "cuBLASLt INT8 GEMM measured in this run (MEASURED_PEAKS.json absent)",
                "frac_vs_cublaslt_int8": {"sustained": (ach_int8 / int8_sustained) if ach_int8 and int8_sustained else None,
                                          "burst": (ach_int8 / int8_burst) if ach_int8 and int8_burst else None,
                                          "cublaslt_sustained_tops": int8_sustained, "cublaslt_burst_tops": int8_burst,
                                          "how": "torch._int_mm 8192^3 measured in this run"},
                "ncu": ncu,
                "int8_ops_per_launch": int8_ops, "int8_products_per_fp64_product": pairs,
                "scheme": "residues (Ozaki II)" if crt else "digit planes (Ozaki I)", "moduli_or_digit_planes": T,
                "fp64_equivalent": {"achieved_tflops": achieved, "cublas_dgemm_tflops": peak_tf,
                                    "ratio": (achieved / peak_tf) if achieved and peak_tf else None,
                                    "note": "algorithmic 2 n^2 flop per chain-step; cuBLAS DGEMM 8192^3 measured in this run"},
                "guard": {"checks": st["guard_checks"], "max_rel_diff_vs_dmma": st["guard_max"],
                          "max_drift_vs_fp64_sentinel_chains": st["drift_max"],
                          "fallback_pass": st["fallback_pass"], "replay_from": st["replay_from"],
                          "tolerance": 1e-12, "drift_tolerance": 1e-11},
                "flops_per_launch": flops_per_launch, "launch_ms": gemm_ms,
                "launches_per_step": st["gemm_launches"],
                "gemm_share_of_step": share,
                "traffic": traffic}
            if l2sm and gemm_ms > 0 and crt and world == 1:
                # what actually bounds the residue kernel: the residue images delivered from L2 to the SMs' shared memory
                tbs = l2sm["bytes_per_launch"] / (gemm_ms * 1e-3) / 1e12
                roofline["operand_delivery_l2_to_sm"] = {
                    "achieved_tbs": tbs, "peak_tbs": l2sm["peak_tbs"], "frac": tbs / l2sm["peak_tbs"],
                    "bytes_per_launch": l2sm["bytes_per_launch"], "peak_source": l2sm["peak_source"],
                    "note": "bytes from the ncu capture (l1tex__m_xbar2l1tex_read_bytes), duration live; the tensor pipe waits for "
                            "operand stages 37 % of a launch (per-role cycle counters, profiles/r2b_lgg09_crt_gemm_ncu_full_s20.md)"}
        else:
            roofline = {
                "bound": "tensor", "kernel": "nt_dgemm_ws_kernel<...,Lgg09PackedEpi> (FP64 DMMA + fused support epilogue)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": (achieved / peak_tf) if achieved and peak_tf else None,
                "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul f64) measured in this run, burst; "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "ncu": ncu,
                "flops_per_launch": flops_per_launch, "launch_ms": gemm_ms,
                "launches_per_step": st["gemm_launches"],
                "gemm_share_of_step": share,
                "traffic": traffic}
        step_ms = total_ms / args.steps
        breakdown = {"main_gemm_ms": st["gemm_launches"] * gemm_ms, "other_ms": step_ms - st["gemm_launches"] * gemm_ms,
                     "note": "other = row-stack build + digit-plane slicing + combine kernels not hidden behind the GEMM + guard checks"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.name, "n": n, "ndirs": ndirs, "nsteps": nsteps,
                       "ndirs_per_gpu": mloc, "chains_per_gpu": st["nchains"], "time_block": S,
                       "tile": "%dx%d" % (st["tile_m"], st["tile_n"]), "main_kernel": st["path"],
                       "sharding": "template directions",
                       "collective": ("all-gather of per-rank rho blocks" + ("; in-place all-gathers of the row stack's doubling steps "
                                      "(each rank builds 1/N of the stack: %.1f ms per run incl. the gathers)" % st.get("stack_build_ms", 0.0)
                                      if dist_stack else "")) if world > 1 else "none",
                       "l2": "working set (row stack %d MB + rho %d MB) exceeds the 126 MB L2; 512 MB flush "
                             "between passes" % ((S + 1) * 2048 * n * 8 * (2 if st["path"] == "int8" else 1) >> 20,
                                                 mloc * nsteps * 8 >> 20)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "timer": ("host wall clock around reach_lgg09_plan_create / _upload(host buffers) / shared stack build / _run / "
                              "_fetch(host buffer) / _destroy + the all-gather, max over ranks") if dist_stack else
                             "host wall clock around reach_lgg09(host buffers), max over ranks",
                    "bit_identical_to_device_resident_run": same, "max_abs_diff_to_device_resident_run": e2e_maxdiff},
            "gpu_launches": launches,
            "wall_s_timed_region": wall,
            "roofline": roofline,
        }
        line["step_breakdown"] = breakdown
        if glgm06 is not None:
            line["glgm06_c4"] = glgm06
        if world == 1 and not args.no_fp64_leg and st["path"] == "int8":
            leg, diff = fp64_path_leg(args, ctx, torch, wl, out_T, peak_tf)
            line["fp64_path"] = leg
            line["int8_vs_dmma_full_run"] = diff
        if not args.no_cpu_baseline:
            if args.cpu_sample_steps:
                info, ref, sample, _ = cpu_lgg09_arm(wl, budget_s=1e-9, min_steps=args.cpu_sample_steps)
            else:
                info, ref, sample, _ = cpu_lgg09_arm(wl, budget_s=12.0, min_steps=100)
            if world == 1:
                info["max_scaled_diff_vs_gpu"] = scaled_diff(out_T[:sample].cpu().numpy().T, ref)
            line["cpu_baseline"] = info
        if world == 1 and not args.no_other_configs:
            line["configs"] = other_configs_leg(ctx, torch, peak_tf, args.no_cpu_baseline)
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
