"""Full-length parity of the BASELINE.json configurations: the CUDA path (through the C ABI) against the C/OpenMP
restatement of the reference loops (oracle/c/reach_ref.c, pinned to the NumPy oracle and the reference's KATs by
tests/test_cref.py and tests/test_oracle_kat.py) over EVERY step of the run, not a prefix:

  C1  Building  96 dirs x 10 000 steps      all three device paths (stack GEMM, persistent chain, time block 1)
  C2  ISS      540 dirs x 33 334 steps      compact chains, persistent chain kernel
  C3  HEAT02  2001 dirs x 1 200 steps       of the 10^4-step run (the CPU port needs ~25 ms per step at n = 1000)
  C4  Building GLGM06 1024 splits x 500 steps: rho(e25, Z) of all 512 000 zonotopes, and 8 splits x 20 steps in full
      (centre, generators, GIR05 selection indices)
  C5  n = 2000, 4000 dirs: 100 steps against the CPU port; all 10^4 steps INT8 digit-plane path vs FP64 DMMA path

Tolerance (north_star): 1e-10, relative to max(|rho_ref|, row scale) for support values and to the zonotope's scale for
generators; generator selection identical."""
import numpy as np
import pytest

import reach_b200 as rb
from reach_b200 import workloads as W
from reach_b200.lgg09 import Lgg09Plan
from reach_b200.glgm06 import _run_batch, GLGM06
from oracle import cref, lazysets_oracle as LZ
from conftest import assert_support_parity

pytestmark = pytest.mark.gpu


def run(ctx, wl, nsteps, **opts):
    plan = Lgg09Plan(ctx, wl.n, wl.dirs.shape[1], nsteps, **opts)
    plan.upload(wl.Phi, wl.dirs, wl.Omega0, wl.V)
    plan.run()
    return plan


def test_c1_building_every_step_every_path(ctx):
    wl = W.building_c1()
    ref = cref.reach_lgg09(wl.dirs, wl.Omega0, wl.Phi, wl.nsteps, wl.V, shape=1)
    assert ref.shape == (96, 10000)
    for opts in ({}, {"path": "chain"}, {"time_block": 1, "path": "gemm"}):
        plan = run(ctx, wl, wl.nsteps, **opts)
        assert_support_parity(plan.sfmat(), ref)
        plan.close()


def test_c2_iss_every_step(ctx):
    wl = W.iss_c2()
    ref = cref.reach_lgg09(wl.dirs, wl.Omega0, wl.Phi, wl.nsteps, wl.V, shape=1, sparse=True, share=True)
    assert ref.shape == (540, 33334)
    plan = run(ctx, wl, wl.nsteps)
    assert plan.stats()["path"] == "compact"
    assert_support_parity(plan.sfmat(), ref)
    plan.close()
    plan = run(ctx, wl, wl.nsteps, path="chain")
    assert_support_parity(plan.sfmat(), ref)
    plan.close()


def test_c3_heat02_first_1200_steps_of_the_full_run(ctx):
    wl = W.heat_c3()
    k = 1200
    box = wl.dirs[:, 1:]
    ref = np.vstack([cref.reach_lgg09(wl.dirs[:, :1], wl.Omega0, wl.Phi, k, None, shape=1),
                     cref.reach_lgg09(box, wl.Omega0, wl.Phi, k, None, shape=2, share=True)])
    plan = run(ctx, wl, wl.nsteps)
    got = plan.fetch(0, k)
    assert_support_parity(got, ref)
    plan.close()


def test_c4_glgm06_1024_splits_500_steps(ctx):
    nsteps, r = 500, 10
    Phi, Oms, V = W.building_c4(2, nsteps)
    assert len(Oms) == 1024
    plan = _run_batch(GLGM06(δ=0.002, max_order=r), Phi, Oms, V, nsteps, ctx)
    zs = [LZ.reduce_order_gir05(*LZ.to_zonotope(Om), r) for Om in Oms]
    cU, GU = LZ.to_zonotope(V)
    c0 = np.stack([z[0] for z in zs], axis=1)
    d = np.zeros(48)
    d[24] = 1.0
    rng = np.random.default_rng(4)
    splits = sorted(rng.choice(1024, 8, replace=False).tolist())
    steps = sorted(set([0, 1, 2, 3, 9, 10, 11, 12, 499, 498] + rng.choice(500, 12, replace=False).tolist()))[:20]
    samples = [(s, k) for s in splits for k in steps]
    rho, outs = cref.reach_inhomog_glgm06(Phi, c0, [z[1] for z in zs], nsteps, r, cU, GU, d=d, samples=samples)
    got = plan.support(d)
    assert got.shape == rho.shape == (1024, 500)
    assert np.abs(got - rho).max() <= 1e-10 * np.abs(rho).max()
    reduced = 0
    for (s, k), (c, G, sel) in zip(samples, outs):
        gc, gG, gsel = plan.fetch(s, k, with_selection=True)
        assert gG.shape == G.shape, (s, k)
        scale = max(np.abs(G).max(), np.abs(c).max())
        assert np.abs(gc - c).max() <= 1e-10 * scale and np.abs(gG - G).max() <= 1e-10 * scale, (s, k)
        if sel is not None:
            reduced += 1
            np.testing.assert_array_equal(gsel[gsel >= 0], sel[sel >= 0], err_msg="selection differs at %s" % ((s, k),))
            assert np.all(gsel[len(sel[sel >= 0]):] == -1)
    assert reduced >= 100
    plan.close()


def test_c5_100_steps_against_cpu_and_full_run_int8_vs_dmma(ctx):
    wl = W.synthetic_c5()
    ref = cref.reach_lgg09(wl.dirs, wl.Omega0, wl.Phi, 100, wl.V, shape=2, share=True)
    auto = run(ctx, wl, wl.nsteps)
    st = auto.stats()
    assert st["path"] == "int8" and st["fallback_pass"] == -1 and st["guard_checks"] >= 3
    a = auto.sfmat()
    assert a.shape == (4000, 10000)
    assert_support_parity(a[:, :100], ref)
    auto.close()
    dmma = run(ctx, wl, wl.nsteps, path="gemm")
    b = dmma.sfmat()
    assert_support_parity(b[:, :100], ref)
    err = assert_support_parity(a, b, tol=1e-11)          # every one of the 4·10^7 support values
    print("C5 full run: INT8 digit-plane path vs FP64 DMMA path, max scaled difference %.2e" % err)
    dmma.close()
