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
