"""Several GPUs behind one handle of the C ABI (`reach_multi_*`, include/reach.h): directions / splits are sharded inside
the library, the LGG09 all-gather is done by 2-D copies.  On a one-GPU box the handle is opened on (0, 0) -- two
contexts, two host threads, the same sharding and gather code; on a multi-GPU box on (0, 1, ...)."""
import ctypes as C

import numpy as np
import pytest

import reach_b200 as rb
from reach_b200 import _lib, discretization as D, workloads as W
from reach_b200.support_program import flatten, to_c
from reach_b200.lgg09 import reach_inhomog_LGG09
from oracle import lgg09_oracle as LO, glgm06_oracle as GO
from conftest import assert_support_parity

pytestmark = pytest.mark.gpu


def device_ids(k):
    n = _lib.load().reach_device_count()
    return [i % max(n, 1) for i in range(k)]


@pytest.fixture(scope="module")
def multi2():
    m = _lib.MultiContext(device_ids(2))
    yield m
    m.close()


def test_multi_handle_basics(multi2):
    assert multi2.ndev == 2
    lib = multi2.lib
    assert lib.reach_multi_ctx(multi2.handle, 0) and lib.reach_multi_ctx(multi2.handle, 1)
    assert not lib.reach_multi_ctx(multi2.handle, 2)
    with pytest.raises(_lib.ReachError):
        _lib.MultiContext([9999])
    with pytest.raises(ValueError):
        multi2.lgg09(np.eye(2), np.eye(2), to_c(flatten(rb.Hyperrectangle([0, 0], [1, 1]), None), 2), None, 0)


@pytest.mark.parametrize("n,ndirs_extra,nsteps,ndev", [(6, 0, 40, 2), (48, 5, 120, 2), (48, 3, 64, 3), (130, 0, 33, 2), (520, 0, 30, 2)])
def test_multi_lgg09_is_bit_identical_to_one_device(ctx, n, ndirs_extra, nsteps, ndev):
    rng = np.random.default_rng(n + ndirs_extra)
    A = -np.eye(n) + (0.5 / np.sqrt(n)) * rng.standard_normal((n, n))
    X0 = rb.Hyperrectangle(rng.uniform(-1, 1, n), 0.1 * rng.uniform(0, 1, n))
    U = rb.LinearMap(rng.standard_normal((n, 2)), rb.Hyperrectangle(rng.uniform(-0.2, 0.2, 2), [0.05, 0.02]))
    ivp = D.forward(A, X0, 0.05, U)
    # box template with a few custom directions mixed in (odd count: one chain has no negation)
    dirs = np.hstack([np.eye(n), rng.standard_normal((n, ndirs_extra)), -np.eye(n)])
    om, v = to_c(flatten(ivp.Omega0, ivp.Phi), n), to_c(flatten(ivp.V, None), n)
    m = _lib.MultiContext(device_ids(ndev))
    try:
        rho, ms = m.lgg09(ivp.Phi, dirs, om, v, nsteps)
    finally:
        m.close()
    one = reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, None, ivp.V, ctx=ctx).sfmat()
    if n < 512:
        assert np.array_equal(rho, one)                  # same kernels on the same chains: bit-identical
    else:
        # the INT8 path groups the passes of a shard (J depends on the chains per device): same numbers up to reassociation
        assert_support_parity(rho, one, tol=1e-12)
    assert np.all(ms > 0)
    assert_support_parity(rho, LO.reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, nsteps, ivp.V))


def test_multi_lgg09_device_resident_all_gather(multi2):
    """rho_dev: every device receives the complete ndirs x NSTEPS matrix through 2-D (peer) copies."""
    import torch
    wl = W.building_c1()
    n, ndirs, nsteps = wl.n, wl.dirs.shape[1], 300
    om, v = to_c(flatten(wl.Omega0, wl.Phi), n), to_c(flatten(wl.V, None), n)
    bufs = [torch.full((nsteps, ndirs), float("nan"), dtype=torch.float64, device="cuda:%d" % d) for d in multi2.device_ids]
    torch.cuda.synchronize()
    rho, _ = multi2.lgg09(wl.Phi, wl.dirs, om, v, nsteps, rho_dev=[b.data_ptr() for b in bufs])
    ref = LO.reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, nsteps, wl.V)
    assert_support_parity(rho, ref)
    for b in bufs:                                       # [nsteps, ndirs] row-major == ndirs x nsteps column-major
        assert np.array_equal(b.cpu().numpy().T, rho)
    # device-resident only (no host matrix)
    for b in bufs:
        b.fill_(float("nan"))
    torch.cuda.synchronize()
    out, _ = multi2.lgg09(wl.Phi, wl.dirs, om, v, nsteps, rho_dev=[bufs[0].data_ptr(), None], want_host=False)
    assert out is None and np.array_equal(bufs[0].cpu().numpy().T, rho) and bool(torch.isnan(bufs[1]).all())


@pytest.mark.parametrize("nsplits,ndev", [(5, 2), (3, 3), (1, 2)])
def test_multi_glgm06_matches_one_device_and_oracle(ctx, nsplits, ndev):
    n, nsteps, max_order = 6, 30, 4
    rng = np.random.default_rng(nsplits)
    A = -np.eye(n) + (0.5 / np.sqrt(n)) * rng.standard_normal((n, n))
    import scipy.linalg
    Phi = scipy.linalg.expm(A * 0.05)
    p0, pU = n + 2, 3
    c0 = np.asfortranarray(rng.standard_normal((n, nsplits)))
    G0 = np.asfortranarray(0.1 * rng.standard_normal((n, p0, nsplits)))
    cU, GU = 0.01 * rng.standard_normal(n), 0.02 * rng.standard_normal((n, pU))
    gcap = max(p0, max_order * n)
    m = _lib.MultiContext(device_ids(ndev))
    try:
        c, G, ng = m.glgm06(Phi, c0, G0, cU, GU, nsteps, max_order, gcap)
    finally:
        m.close()
    single = _lib.MultiContext([0])
    try:
        c1, G1, ng1 = single.glgm06(Phi, c0, G0, cU, GU, nsteps, max_order, gcap)
    finally:
        single.close()
    assert np.array_equal(ng, ng1) and np.array_equal(c, c1) and np.array_equal(G, G1)
    for s in range(nsplits):
        ref = GO.post_GLGM06(Phi, rb.Zonotope(c0[:, s], G0[:, :, s]), nsteps, max_order, rb.Zonotope(cU, GU))
        for k in (0, 1, nsteps // 2, nsteps - 1):
            cr, Gr = ref[k]
            assert ng[s, k] == Gr.shape[1]
            scale = max(np.abs(Gr).max(), np.abs(cr).max())
            assert np.abs(c[:, s, k] - cr).max() <= 1e-10 * scale
            assert np.abs(G[:, :ng[s, k], s, k] - Gr).max() <= 1e-10 * scale
