"""world_size-2 test of the N>1 path on CPU (gloo): shard the directions, run each shard
(the CPU oracle stands in for the device pass), all-gather, compare with the unsharded run."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import reach_b200 as rb
from reach_b200 import discretization as D, sharding
from oracle import lgg09_oracle as LO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    rng = np.random.default_rng(3)
    n = 7
    A = -np.eye(n) + 0.2 * rng.standard_normal((n, n))
    X0 = rb.Hyperrectangle(rng.standard_normal(n), 0.1 * rng.uniform(0, 1, n))
    U = rb.BallInf(np.zeros(n), 0.05)
    ivp = D.forward(A, X0, 0.05, U)
    dirs = np.hstack([np.eye(n), -np.eye(n), rng.standard_normal((n, 3))])
    return ivp, dirs, 25


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ivp, dirs, nsteps = _problem()
    local, rows_by_rank, mloc = sharding.shard_directions(dirs, world, rank)
    block = LO.reach_inhomog_LGG09(local, ivp.Omega0, ivp.Phi, nsteps, ivp.V)      # mloc x nsteps
    full = sharding.gather_support_values(torch.from_numpy(np.ascontiguousarray(block.T)), rows_by_rank,
                                          dirs.shape[1])
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                                         # the bench's max-over-ranks
    q.put((rank, full.numpy().T.copy(), float(t.item())))
    dist.destroy_process_group()


def test_two_rank_sharded_run_matches_unsharded():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ivp, dirs, nsteps = _problem()
    ref = LO.reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, ivp.V)
    for rank, full, tmax in results:
        assert tmax == 2.0
        np.testing.assert_allclose(full, ref, rtol=0, atol=1e-13)


def _split_problem():
    rng = np.random.default_rng(11)
    n = 5
    A = -np.eye(n) + 0.2 * rng.standard_normal((n, n))
    U = rb.BallInf(np.zeros(n), 0.05)
    X0s = [rb.Hyperrectangle(rng.standard_normal(n), 0.1 * rng.uniform(0, 1, n)) for _ in range(5)]
    ivps = [D.forward(A, X0, 0.05, U) for X0 in X0s]
    dirs = np.hstack([np.eye(n), -np.eye(n)])
    return ivps, dirs, 20


def _split_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ivps, dirs, nsteps = _split_problem()
    mine = sharding.shard_splits(len(ivps), world, rank)
    local = np.full(dirs.shape[1], -np.inf)
    for i in mine:           # the CPU oracle stands in for the device batch of this rank's splits
        rho = LO.reach_inhomog_LGG09(dirs, ivps[i].Omega0, ivps[i].Phi, nsteps, ivps[i].V)
        local = np.maximum(local, rho.max(axis=1))
    full = sharding.reduce_mixed_flowpipe_support(torch.from_numpy(local))
    q.put((rank, list(mine), full.numpy().copy()))
    dist.destroy_process_group()


def test_two_rank_split_sharding_and_mixed_flowpipe_support():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_split_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ivps, dirs, nsteps = _split_problem()
    ref = np.max([LO.reach_inhomog_LGG09(dirs, iv.Omega0, iv.Phi, nsteps, iv.V).max(axis=1) for iv in ivps], axis=0)
    owned = sorted(i for _, mine, _ in results for i in mine)
    assert owned == list(range(len(ivps)))                     # every split on exactly one rank
    for rank, mine, full in results:
        assert len(mine) in (2, 3)
        np.testing.assert_array_equal(full, ref)


def test_shard_splits_is_a_balanced_partition():
    for nsplits in (1, 5, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            parts = [list(sharding.shard_splits(nsplits, world, r)) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(nsplits))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


class _HostPlan:
    """Stands in for a GEMM-path `Lgg09Plan` in the CPU test of `sharding.build_stack_distributed`: the same stepwise row-stack
    build as `reach_lgg09_plan_stack_step` (step 0: block 1 = block 0 x P; step j: blocks s+1.. = blocks 1.. x P^s with
    s = 2^(j-1), after Q_{2s} = Q_s Q_s), in NumPy on a host tensor."""

    class _Ctx:
        def synchronize(self):
            pass

    def __init__(self, Phi, E, S):
        n = Phi.shape[0]
        self.n, self.S = n, S
        self.nbe = -(-(n + E.shape[0]) // 8) * 8
        self.stack = torch.zeros(((S + 1) * self.nbe + 256, n), dtype=torch.float64)
        self.stack[:n, :n] = torch.eye(n, dtype=torch.float64)
        self.stack[n:n + E.shape[0]] = torch.from_numpy(E)
        self.P = {1: Phi.T.copy()}                            # row r of block b = row r of block 0 times (Phi')^b
        self.ctx = self._Ctx()
        self.steps_done, self.ready, self.calls = 0, False, []

    def stack_info(self):
        return 0, self.nbe, self.n, self.S

    def stack_step(self, step, lo=-1, hi=-1):
        s = 0 if step == 0 else 1 << (step - 1)
        if step > 0 and s >= self.S:
            return 0, 0
        rows = self.nbe if step == 0 else min(s, self.S - s) * self.nbe
        dest = self.nbe if step == 0 else (s + 1) * self.nbe
        if lo < 0:
            return rows, dest
        assert step in (self.steps_done, self.steps_done - 1) and 0 <= lo <= hi <= rows
        if step > 0 and 2 * s not in self.P:
            self.P[2 * s] = self.P[s] @ self.P[s]
        src0 = 0 if step == 0 else self.nbe
        X = self.stack[src0 + lo:src0 + hi].numpy()
        self.stack[dest + lo:dest + hi] = torch.from_numpy(X @ self.P[max(s, 1)])
        self.calls.append((step, lo, hi))
        self.steps_done = step + 1
        return rows, dest

    def stack_ready(self):
        need = 1
        s = 1
        while s < self.S:
            need, s = need + 1, s * 2
        assert self.steps_done == need
        self.ready = True


def _stack_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    n, S = 13, 11
    Phi = 0.9 * np.eye(n) + 0.05 * rng.standard_normal((n, n))
    E = rng.standard_normal((3, n))
    plan = _HostPlan(Phi, E, S)
    steps = sharding.build_stack_distributed(plan, stack=plan.stack)
    q.put((rank, steps, plan.ready, plan.stack[:(S + 1) * plan.nbe].numpy().copy(), plan.calls))
    dist.destroy_process_group()


def test_two_ranks_build_the_row_stack_together():
    """`sharding.build_stack_distributed` on gloo: each rank computes its row range of every doubling step, the in-place
    all-gathers complete the stack on both, and the result equals the powers computed directly."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stack_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    rng = np.random.default_rng(11)
    n, S = 13, 11
    Phi = 0.9 * np.eye(n) + 0.05 * rng.standard_normal((n, n))
    E = rng.standard_normal((3, n))
    nbe = 16
    B0 = np.zeros((nbe, n))
    B0[:n] = np.eye(n)
    B0[n:n + 3] = E
    for rank, steps, ready, stack, calls in results:
        assert ready and steps == 5                          # block 1, then s = 1, 2, 4, 8
        for b in range(S + 1):
            want = B0 @ np.linalg.matrix_power(Phi.T, b)
            assert np.allclose(stack[b * nbe:(b + 1) * nbe], want, rtol=1e-12, atol=1e-14)
    assert np.array_equal(results[0][3], results[1][3])      # both ranks hold the same stack, bit for bit
    # the ranks took disjoint row ranges of every step
    for st in range(5):
        r0 = [(lo, hi) for s_, lo, hi in results[0][4] if s_ == st]
        r1 = [(lo, hi) for s_, lo, hi in results[1][4] if s_ == st]
        assert len(r0) == 1 and len(r1) == 1 and r0[0][1] <= r1[0][0]
