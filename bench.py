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
    ap.add_argument("--path", default="auto", choices=["auto", "gemm", "int8"],
                    help="main-loop kernel: auto (INT8 digit-plane GEMM guarded by DMMA passes), gemm (FP64 DMMA), int8 (unguarded)")
    ap.add_argument("--cpu-sample-steps", type=int, default=0, help="time steps of the CPU sample (0: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-glgm06", action="store_true", help="skip the GLGM06 C4 flowpipe wall-time leg")
    ap.add_argument("--glgm06-steps", type=int, default=500)
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short C1 / C2 / split-caller legs (rank 0, N = 1 only)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(wl, sample_steps):
    """The reference's CPU algorithm (oracle port, GEMM-per-step shape of reach_inhomog.jl:174-218,
    multi-threaded BLAS) on the first `sample_steps` time steps of the same workload."""
    from oracle import lgg09_oracle as LO
    t0 = time.perf_counter()
    out = LO.reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, sample_steps, wl.V)
    dt = time.perf_counter() - t0
    return out, dt


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        n = 1
    return min(n, len(os.sched_getaffinity(0))) if n else len(os.sched_getaffinity(0))


def auto_sample_steps(wl):
    # aim at ~10-20 s of CPU work: time two steps, scale
    from oracle import lgg09_oracle as LO
    t0 = time.perf_counter()
    LO.reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, 3, wl.V)
    per = (time.perf_counter() - t0) / 3
    return int(max(5, min(400, 12.0 / max(per, 1e-4))))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from reach_b200 import workloads as W
    wl = W.synthetic_c5(n=args.n, nsteps=args.nsteps)
    sample = args.cpu_sample_steps or auto_sample_steps(wl)
    ndirs = wl.dirs.shape[1]
    for _ in range(min(args.warmup, 1)):
        cpu_reference_sample(wl, max(2, sample // 10))
    total = 0.0
    for _ in range(args.steps):
        _, dt = cpu_reference_sample(wl, sample)
        total += dt
    value = ndirs * sample * args.steps / total
    cores = cpu_threads()
    desc = "first %d of %d time steps, all %d directions, per bench step" % (sample, args.nsteps, ndirs)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": wl.name, "n": wl.n, "ndirs": ndirs, "nsteps": args.nsteps,
                   "sample": desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                         "note": "NumPy/OpenBLAS restatement of the reference loops (Julia is not installed); "
                                 "one multi-threaded GEMM per time step + vectorised support functions"},
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


def glgm06_leg(args, ctx, torch, dist, world, rank):
    """Second half of the metric: GLGM06 flowpipe wall-time on config C4 (Building, 1024 initial-set
    splits, max_order 10, Forward(:zono) model), splits sharded over the ranks, no collective."""
    from reach_b200 import workloads as W
    from reach_b200.glgm06 import _run_batch, GLGM06
    nsteps = args.glgm06_steps
    Phi, Oms, V = W.building_c4(2, nsteps)
    from reach_b200.sharding import shard_splits
    own = shard_splits(len(Oms), world, rank)
    mine = Oms[own.start:own.stop]
    plan = _run_batch(GLGM06(δ=0.002, max_order=10), Phi, mine, V, nsteps, ctx)       # warm-up run
    ms = []
    for _ in range(3):
        if world > 1:
            dist.barrier()
        plan.run()
        ms.append(plan.stats()["ms"])
    t = min(ms)
    if world > 1:
        tt = torch.tensor([t], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt.item())
    st = plan.stats()
    out = {"workload": "C4 Building GLGM06 max_order=10, Forward(setops=:zono), %d splits x %d steps" % (len(Oms), nsteps),
           "flowpipe_wall_ms": t, "zonotopes_per_s": len(Oms) * nsteps / (t * 1e-3), "splits_per_gpu": len(mine),
           "generators_per_zonotope": 480, "launches": st["launches"], "bytes_stored_per_gpu": st["bytes_stored"],
           "timer": "CUDA events around the run on the library stream, best of 3, max over ranks"}
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import glgm06_oracle as GO
        k = min(100, nsteps)
        t0 = time.perf_counter()
        GO.post_GLGM06(Phi, Oms[0], k, 10, V)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"zonotopes_per_s": k / dt, "cores": 1, "kind": "port",
                               "sample": "1 split x %d steps (NumPy restatement of reach_inhomog_GLGM06!)" % k}
    plan.close()
    return out


def other_configs_leg(ctx):
    """The other BASELINE configs and the split caller, device-resident, best of 3 (CUDA events on the library
    stream): C1 Building (DMMA GEMM over the power stack), C2 ISS (compact chains: 135 decoupled 2x2 blocks),
    C3 HEAT02 (n = 1000, 2001 directions; first 2000 of the 10^4 steps, stack build included; the INT8 guard sends
    it to the DMMA kernel at pass 0), LGG09 over 256 initial-set splits of Building (one batch)."""
    import reach_b200 as rb
    from reach_b200 import workloads as W, discretization as D, sets as S
    from reach_b200.lgg09 import Lgg09Plan, Lgg09BatchPlan
    out = {}
    for key, wl in (("C1", W.building_c1()), ("C2", W.iss_c2()), ("C3", W.heat_c3(nsteps=2000))):
        plan = Lgg09Plan(ctx, wl.n, wl.dirs.shape[1], wl.nsteps)
        plan.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
        ms = []
        for _ in range(4):
            plan.run()
            ms.append(plan.stats()["ms"])
        st = plan.stats()
        t = min(ms[1:])
        out[key] = {"workload": wl.name, "ndirs": wl.dirs.shape[1], "nsteps": wl.nsteps, "ms": t,
                    "evals_per_s": wl.dirs.shape[1] * wl.nsteps / (t * 1e-3), "path": st["path"],
                    "launches": st["launches"], "time_block": st["time_block"]}
        if st["path"] == "int8":      # guarded: fallback_pass >= 0 means the run continued on the FP64 (DMMA) kernel
            out[key].update(guard_checks=st["guard_checks"], guard_max=st["guard_max"], fallback_pass=st["fallback_pass"])
        plan.close()
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


def run_b200(args):
    import torch
    import torch.distributed as dist
    import reach_b200  # noqa: F401
    from reach_b200 import _lib, workloads as W
    from reach_b200.lgg09 import Lgg09Plan
    from reach_b200.sharding import shard_directions, gather_support_values

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
    n, ndirs, nsteps = wl.n, wl.dirs.shape[1], args.nsteps
    local_dirs, rows_by_rank, mloc = shard_directions(wl.dirs, world, rank)

    # device-resident inputs and output
    out_T = torch.empty((nsteps, mloc), dtype=torch.float64, device="cuda")     # == mloc x nsteps col-major
    plan = Lgg09Plan(ctx, n, mloc, nsteps, time_block=args.time_block, path=args.path)
    plan.set_output(out_T.data_ptr())
    plan.upload(wl.Phi, local_dirs, wl.Omega0, wl.V)
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")

    def one_pass():
        """Whole hot path, inputs resident in HBM: recurrence on this rank's chains + all-gather."""
        flush.zero_()
        barrier()
        plan.run(sync=True)
        st = plan.stats()
        ms = st["ms"]
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
    om, v = to_c(flatten(wl.Omega0, wl.Phi), n), to_c(flatten(wl.V, wl.Phi), n)
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

    glgm06 = None if args.no_glgm06 else glgm06_leg(args, ctx, torch, dist, world, rank)

    if rank == 0:
        S = st["time_block"]
        gemm_ms = gemm_ms_acc / args.steps                                  # mean duration of one launch
        flops_per_launch = 2.0 * n * n * st["nchains"] * S                  # algorithmic: 2 n^2 per chain-step
        achieved = flops_per_launch / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None   # FP64-equivalent TFLOP/s
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "lgg09_main_kernel_traffic.json")))
            ent = tj.get(st["path"])
            if ent and ent.get("n") == n and ent.get("chains") == st["nchains"]:
                # DRAM bytes of one launch from the ncu --set full capture, scaled to this run's time block
                traffic = ent["dram_bytes_per_launch"] * (S / float(ent["time_block"]))
        except Exception:
            pass
        if st["path"] == "int8":
            T = st["digit_planes"]
            pairs = T * (T + 1) // 2
            int8_ops = flops_per_launch * pairs                              # INT8 MMA work actually issued
            ach_int8 = int8_ops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
            roofline = {
                "bound": "tensor",
                "kernel": "oz_gemm_kernel<%d, OzLgg09Epi> (INT8 tcgen05 digit-plane GEMM, FP64 by error-free slicing, "
                          "fused support epilogue)" % T,
                "achieved": ach_int8, "peak": int8_sustained, "unit": "TOP/s (INT8 tensor)",
                "frac": (ach_int8 / int8_sustained) if ach_int8 and int8_sustained else None,
                "peak_source": "cuBLASLt INT8 GEMM 8192^3 (torch._int_mm) measured in this run, sustained over ~2 s "
                               "(the kernel is timed inside a seconds-long pass); burst %.0f TOP/s; "
                               "MEASURED_PEAKS.json has no INT8 entry" % int8_burst,
                "int8_ops_per_launch": int8_ops, "plane_pairs": pairs, "digit_planes": T,
                "fp64_equivalent": {"achieved_tflops": achieved, "cublas_dgemm_tflops": peak_tf,
                                    "ratio": (achieved / peak_tf) if achieved and peak_tf else None,
                                    "note": "algorithmic 2 n^2 flop per chain-step; cuBLAS DGEMM 8192^3 measured in this run"},
                "guard": {"checks": st["guard_checks"], "max_rel_diff_vs_dmma": st["guard_max"],
                          "fallback_pass": st["fallback_pass"], "tolerance": 1e-12},
                "flops_per_launch": flops_per_launch, "launch_ms": gemm_ms,
                "launches_per_step": st["gemm_launches"],
                "gemm_share_of_step": st["gemm_launches"] * gemm_ms / (total_ms / args.steps),
                "traffic": traffic}
        else:
            roofline = {
                "bound": "tensor", "kernel": "nt_dgemm_ws_kernel<...,Lgg09PackedEpi> (FP64 DMMA + fused support epilogue)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": (achieved / peak_tf) if achieved and peak_tf else None,
                "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul f64) measured in this run, burst; "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "flops_per_launch": flops_per_launch, "launch_ms": gemm_ms,
                "launches_per_step": st["gemm_launches"],
                "gemm_share_of_step": st["gemm_launches"] * gemm_ms / (total_ms / args.steps),
                "traffic": traffic}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.name, "n": n, "ndirs": ndirs, "nsteps": nsteps,
                       "ndirs_per_gpu": mloc, "chains_per_gpu": st["nchains"], "time_block": S,
                       "tile": "%dx%d" % (st["tile_m"], st["tile_n"]), "main_kernel": st["path"],
                       "sharding": "template directions",
                       "collective": "all-gather of per-rank rho blocks" if world > 1 else "none",
                       "l2": "working set (row stack %d MB + rho %d MB) exceeds the 126 MB L2; 512 MB flush "
                             "between passes" % ((S + 1) * 2048 * n * 8 * (2 if st["path"] == "int8" else 1) >> 20,
                                                 mloc * nsteps * 8 >> 20)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "timer": "host wall clock around reach_lgg09(host buffers), max over ranks",
                    "bit_identical_to_device_resident_run": same},
            "gpu_launches": launches,
            "wall_s_timed_region": wall,
            "roofline": roofline,
        }
        if glgm06 is not None:
            line["glgm06_c4"] = glgm06
        if world == 1 and not args.no_other_configs:
            line["other_configs"] = other_configs_leg(ctx)
        if not args.no_cpu_baseline:
            sample = args.cpu_sample_steps or auto_sample_steps(wl)
            ref, dt = cpu_reference_sample(wl, sample)
            got = out_T[:sample].cpu().numpy().T if world == 1 else None
            line["cpu_baseline"] = {
                "value": ndirs * sample / dt, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                "sample": "first %d of %d time steps, all %d directions" % (sample, nsteps, ndirs)}
            if got is not None:
                scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=1, keepdims=True))
                line["cpu_baseline"]["max_scaled_diff_vs_gpu"] = float((np.abs(got - ref) / scale).max())
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
