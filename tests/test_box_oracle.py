"""BOX oracle (oracle/box_oracle.py) against the LGG09 oracle: on a hyperrectangular Omega0 / U the BOX flowpipe is the
box-template support-function flowpipe (rho(+-e_i) = +-c_i + r_i), reach_homog.jl:8-47 vs LGG09/reach_homog.jl:49-62."""
import numpy as np
import scipy.linalg

import reach_b200 as rb
from oracle import box_oracle as BO, lgg09_oracle as LO


def test_box_oracle_equals_lgg09_oracle_on_boxes():
    rng = np.random.default_rng(0)
    n, nsteps = 9, 40
    A = -np.eye(n) + 0.2 * rng.standard_normal((n, n))
    Phi = scipy.linalg.expm(0.05 * A)
    c0, r0 = rng.uniform(-1, 1, n), 0.1 * rng.uniform(0, 1, n)
    cU, rU = rng.uniform(-0.1, 0.1, n), 0.02 * rng.uniform(0, 1, n)
    dirs = np.hstack([np.eye(n), -np.eye(n)])
    C, R = BO.reach_inhomog_BOX(c0, r0, Phi, nsteps, cU, rU)
    rho = LO.reach_inhomog_LGG09(dirs, rb.Hyperrectangle(c0, r0), Phi, nsteps, rb.Hyperrectangle(cU, rU))
    assert np.allclose(C + R, rho[:n], rtol=0, atol=1e-12)
    assert np.allclose(R - C, rho[n:], rtol=0, atol=1e-12)
    Ch, Rh = BO.reach_homog_BOX(c0, r0, Phi, nsteps)
    rho = LO.reach_homog_LGG09(dirs, rb.Hyperrectangle(c0, r0), Phi, nsteps)
    assert np.allclose(Ch + Rh, rho[:n], rtol=0, atol=1e-12)


def test_box_oracle_recursive_wraps_and_invariant_truncates():
    rng = np.random.default_rng(1)
    n = 5
    th = 0.3
    Phi = np.eye(n)
    Phi[:2, :2] = [[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]
    c0, r0 = np.ones(n), 0.1 * np.ones(n)
    _, Rn = BO.reach_homog_BOX(c0, r0, Phi, 30)
    _, Rr = BO.reach_homog_BOX(c0, r0, Phi, 30, recursive=True)
    assert (Rr >= Rn - 1e-15).all() and Rr[0, -1] > 2 * Rn[0, -1]          # wrapping effect of the rotation
    C, _ = BO.reach_homog_BOX(c0, r0, 0.9 * np.eye(n), 50, X=("halfspace", -np.eye(n)[0], -0.5))
    assert C.shape[1] < 50 and C[0, -1] + 0.1 >= 0.5
