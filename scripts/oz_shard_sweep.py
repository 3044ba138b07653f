"""One shard of C5 (the directions of rank 0 of a `world`-GPU run) on the INT8 path: time per pass for the auto
configuration and a few forced time blocks, with and without pass groups (REACH_OZ_NO_GROUPS)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import reach_b200
from reach_b200 import _lib, workloads as W
from reach_b200.lgg09 import Lgg09Plan
from reach_b200.sharding import shard_directions
ctx = _lib.default_context(0)
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
worlds = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8]
Ss = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
wl = W.synthetic_c5(nsteps=nsteps)
ref = None
for world in worlds:
    local_dirs, rows_by_rank, mloc = shard_directions(wl.dirs, world, 0)
    for groups in ([int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 1]):   # 0: auto J, 1: no groups, k: J = k
        os.environ.pop("REACH_OZ_NO_GROUPS", None)
        os.environ.pop("REACH_OZ_GROUP", None)
        if groups == 1:
            os.environ["REACH_OZ_NO_GROUPS"] = "1"
        elif groups > 1:
            os.environ["REACH_OZ_GROUP"] = str(groups)
        for S in Ss:
            plan = Lgg09Plan(ctx, wl.n, mloc, nsteps, time_block=S)
            plan.upload(wl.Phi, local_dirs, wl.Omega0, wl.V)
            plan.run(); plan.run()
            st = plan.stats()
            out = plan.sfmat()
            key = world
            if ref is None or ref[0] != key:
                ref = (key, out)
            d = np.abs(out - ref[1]) / np.maximum(np.abs(ref[1]), np.abs(ref[1]).max(axis=1, keepdims=True))
            print("world %d chains %d groups=%s S=%d: %.2f ms, gemm %.3f ms x %d (share %.3f), launches %d, guard %.1e drift %.1e fb %d; vs first %.1e"
                  % (world, st["nchains"], groups, st["time_block"], st["ms"], st["gemm_ms"], st["gemm_launches"],
                     st["gemm_ms"] * st["gemm_launches"] / st["ms"], st["launches"], st["guard_max"], st["drift_max"], st["fallback_pass"], d.max()), flush=True)
            del plan
