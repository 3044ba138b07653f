"""BASELINE.json configurations at (or near) full size.  The oracle covers a prefix of each run;
beyond that the checks are size-independent properties: two independent device paths agree
(power-stack GEMM vs persistent chain kernel; time block 1 vs auto), template widths are
non-negative, positive homogeneity in the direction, and ρ(d, fp) equals the row maximum."""
import numpy as np
import pytest

import reach_b200 as rb
from reach_b200 import workloads as W
from reach_b200.lgg09 import Lgg09Plan
from oracle import lgg09_oracle as LO
from conftest import assert_support_parity

pytestmark = pytest.mark.gpu


def run(ctx, wl, nsteps, dirs=None, **opts):
    dirs = wl.dirs if dirs is None else dirs
    plan = Lgg09Plan(ctx, wl.n, dirs.shape[1], nsteps, **opts)
    plan.upload(wl.Phi, dirs, wl.Omega0, wl.V)
    plan.run()
    return plan


def widths_nonnegative(rho, n):
    """box template: ρ(+e_i) + ρ(−e_i) is the width of the reach-set along x_i."""
    w = rho[:n] + rho[n:2 * n]
    scale = np.abs(rho).max()
    return w.min() >= -1e-12 * scale


def test_c2_iss_full_run_two_device_paths_agree(ctx):
    """C2: ISSF01, box template (540 directions), δ = 6e-4, 33 334 steps."""
    wl = W.iss_c2()
    ref = LO.reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, 300, wl.V)
    chain = run(ctx, wl, wl.nsteps)                       # auto: persistent chain kernel (sparse Φ)
    assert chain.stats()["tile_m"] == 0
    a = chain.sfmat()
    assert a.shape == (540, 33334)
    assert_support_parity(a[:, :300], ref)
    gemm = run(ctx, wl, wl.nsteps, path="gemm")           # dense power-stack GEMM, reassociated
    b = gemm.sfmat()
    assert_support_parity(b[:, :300], ref)
    assert_support_parity(b, a, tol=1e-10)
    assert widths_nonnegative(a, 270)
    vmax, amax = chain.max_over_steps()
    np.testing.assert_array_equal(vmax, a.max(axis=1))
    np.testing.assert_array_equal(amax, a.argmax(axis=1))


def test_c3_heat02_homogeneous(ctx):
    """C3: HEAT02 (n = 1000), centre cell + box template (2001 directions), 10^4 steps."""
    wl = W.heat_c3()
    ref = LO.reach_homog_LGG09(wl.dirs, wl.Omega0, wl.Phi, 40)
    auto = run(ctx, wl, wl.nsteps)
    a = auto.sfmat()
    assert a.shape == (2001, 10000)
    assert_support_parity(a[:, :40], ref)
    seq = run(ctx, wl, 2000, time_block=1)                # the reference's association, shorter run
    assert_support_parity(a[:, :2000], seq.sfmat())
    assert widths_nonnegative(a[1:], 1000)
    # heat diffuses: the hottest bound stays at the initial maximum 1.1 up to the Forward model's O(δ²) bloating
    assert 1.1 <= a[1:1001].max() <= 1.1 * (1 + 1e-3)


def test_c5_synthetic_reduced_steps(ctx):
    """C5: n = 2000, 4000 directions; oracle on 12 steps, homogeneity + widths on 200 steps."""
    wl = W.synthetic_c5()
    ref = LO.reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, 12, wl.V)
    plan = run(ctx, wl, 200)
    a = plan.sfmat()
    assert_support_parity(a[:, :12], ref)
    assert widths_nonnegative(a, 2000)
    sub = wl.dirs[:, [0, 7, 2000, 2007]]
    b = run(ctx, wl, 200, dirs=2.0 * sub).sfmat()
    c = run(ctx, wl, 200, dirs=sub).sfmat()
    assert_support_parity(b, 2.0 * c, tol=1e-12)          # ρ(2d) = 2ρ(d)
    assert_support_parity(c, a[[0, 7, 2000, 2007]], tol=1e-12)


def test_c1_building_full_run_paths_agree(ctx):
    wl = W.building_c1()
    a = run(ctx, wl, wl.nsteps).sfmat()
    b = run(ctx, wl, wl.nsteps, path="chain").sfmat()
    assert_support_parity(a, b)
    assert widths_nonnegative(a, 48)
