"""Multi-GPU sharding of the LGG09 path: template directions are independent
(`Threads.@threads for j in eachindex(ℓ)`, reach_inhomog.jl:35-40), so each rank owns a
contiguous block of direction *chains* (a direction and its negation stay on the same rank so
they keep sharing (Φᵀ)^k ℓ), and the only collective is one all-gather of the per-rank
support-value blocks.  One process per GPU; `torch.distributed` (NCCL on the GPUs, gloo in the
CPU tests) is plumbing only."""
import numpy as np


def chain_groups(dirs):
    """Group directions into chains: ℓ and −ℓ (exact negation) share a chain.  Returns a list of
    index lists in first-appearance order."""
    n, m = dirs.shape
    index = {}
    groups = []
    for j in range(m):
        d = dirs[:, j]
        nz = np.flatnonzero(d)
        sign = -1.0 if len(nz) and d[nz[0]] < 0 else 1.0
        key = (sign * d + 0.0).tobytes()
        g = index.get(key)
        if g is None or len(groups[g]) >= 2:
            index[key] = len(groups)
            groups.append([j])
        else:
            groups[g].append(j)
    return groups


def shard_directions(dirs, world, rank):
    """Directions of `rank`: (local_dirs [n × mloc], global_rows [mloc_valid], mloc) with every
    rank padded (zero directions, ρ ≡ 0) to the same mloc so the all-gather is uniform."""
    if world == 1:
        # single GPU: keep the caller's order, nothing to gather
        return np.asfortranarray(dirs), [np.arange(dirs.shape[1], dtype=np.int64)], dirs.shape[1]
    groups = chain_groups(dirs)
    per = (len(groups) + world - 1) // world
    rows_by_rank = []
    for r in range(world):
        rows = [j for g in groups[r * per:(r + 1) * per] for j in g]
        rows_by_rank.append(np.array(rows, dtype=np.int64))
    mloc = max(1, max(len(r) for r in rows_by_rank))
    rows = rows_by_rank[rank]
    local = np.zeros((dirs.shape[0], mloc), order="F")
    local[:, :len(rows)] = dirs[:, rows]
    return local, rows_by_rank, mloc


_gather_cache = {}


def gather_support_values(local_T, rows_by_rank, ndirs, group=None):
    """All-gather the per-rank blocks and place the rows.

    local_T: torch tensor [nsteps, mloc] (i.e. the column-major mloc × nsteps block the device
    path produced).  Returns a tensor [nsteps, ndirs] == column-major ndirs × nsteps, the
    reference's ρℓ layout.  The per-rank row indices are built once per (sharding, device) and kept on the
    device, and the gather / output buffers are reused between calls (the returned tensor is overwritten by
    the next call)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    nsteps, mloc = local_T.shape
    key = (id(rows_by_rank), str(local_T.device), nsteps, mloc, ndirs, local_T.dtype)
    ent = _gather_cache.get(key)
    if ent is None:
        ent = {"rows": [torch.as_tensor(np.asarray(rows_by_rank[r], dtype=np.int64), device=local_T.device) for r in range(world)],
               "flat": torch.empty((world, nsteps, mloc), dtype=local_T.dtype, device=local_T.device),
               "out": torch.empty((nsteps, ndirs), dtype=local_T.dtype, device=local_T.device),
               "rows_by_rank": rows_by_rank}                       # keeps id() stable while cached
        _gather_cache.clear()
        _gather_cache[key] = ent
    flat, out = ent["flat"], ent["out"]
    dist.all_gather_into_tensor(flat.view(world * nsteps, mloc), local_T.contiguous(), group=group)
    for r in range(world):                                         # out[:, rows_r] = block of rank r (its valid columns)
        rows = ent["rows"][r]
        if len(rows):
            out.index_copy_(1, rows, flat[r, :, :len(rows)])
    return out


def shard_splits(nsplits, world, rank):
    """Initial-set splits of `rank` (the GLGM06 / LGG09 split callers, solve.jl:84-117): contiguous, balanced
    to within one split.  The splits are independent problems, so there is no data-path collective; each rank
    keeps its flowpipes on its own device."""
    base, extra = divmod(int(nsplits), int(world))
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def reduce_mixed_flowpipe_support(local_max, group=None):
    """ρ(d_j, MixedFlowpipe) over ALL splits from the per-rank maxima over the rank's own splits and steps
    (MixedFlowpipe.jl:66-68): one all-reduce(MAX) of an ndirs vector -- the only exchange a property check on a
    sharded split run needs."""
    import torch.distributed as dist
    out = local_max.clone()
    dist.all_reduce(out, op=dist.ReduceOp.MAX, group=group)
    return out


def stack_row_split(rows, world, rank):
    """Rows of one build step taken by `rank`: equal chunks of a multiple of 8 rows (the last ranks may get fewer or none).
    Returns (chunk, lo, hi)."""
    chunk = -(-rows // world)
    chunk = -(-chunk // 8) * 8
    lo = min(rows, rank * chunk)
    hi = min(rows, (rank + 1) * chunk)
    return chunk, lo, hi


def build_stack_distributed(plan, group=None, stack=None):
    """Build the row stack [I;E](Φᵀ)^b of `plan` (GEMM paths) with the ranks of `group` sharing the work: every doubling step's
    output rows are split over the ranks (`reach_lgg09_plan_stack_step`) and all-gathered in place over NVLink, instead of
    every rank repeating the whole build (the largest per-rank fixed cost of a sharded LGG09 run: `profiles/`).  All ranks
    must hold plans with the same Φ, mapped rows and time block.  The next `plan.run()` uses the stack as it is; the
    products and their order are those of a single-GPU run, so the results are bit-identical.  Returns the number of
    steps taken.

    `stack` (tests): a [rows, ld] tensor over the plan's stack memory; by default a zero-copy CUDA view of
    `plan.stack_info()`'s device pointer."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ptr, nbe, ld, S = plan.stack_info()
    if stack is None:
        dev = torch.cuda.current_device()

        class _Holder:                                        # the whole stack buffer (+ its 256 slack rows) as [rows, ld] float64
            __cuda_array_interface__ = {"shape": ((S + 1) * nbe + 256, ld), "typestr": "<f8", "data": (int(ptr), False),
                                        "version": 3, "strides": None}
        stack = torch.as_tensor(_Holder(), device=torch.device("cuda", dev))
    step = 0
    while True:
        rows, dest = plan.stack_step(step)
        if rows == 0:
            break
        chunk, lo, hi = stack_row_split(rows, world, rank)
        plan.stack_step(step, lo, hi)
        plan.ctx.synchronize()                                # the library's stream -> torch's (NCCL) stream
        region = stack[dest:dest + world * chunk]
        dist.all_gather_into_tensor(region, region[rank * chunk:(rank + 1) * chunk].clone() if not stack.is_cuda
                                    else region[rank * chunk:(rank + 1) * chunk], group=group)
        if stack.is_cuda:
            torch.cuda.current_stream().synchronize()         # the next step reads the gathered rows on the library's stream
        step += 1
    plan.stack_ready()
    return step
