"""Generate the committed golden vectors for the hot path.

The reference (Julia) cannot run in the build container, so these vectors come from the CPU
oracle (oracle/), which is itself pinned by the reference's known-answer tests
(tests/test_oracle_kat.py).  They freeze the oracle's output on small seeded inputs so that
(a) an accidental change of the oracle is caught on CPU and (b) the GPU path is compared against
numbers that do not depend on the oracle code at test time.

Usage: python tests/golden/make_golden.py     (writes tests/golden/*.npz)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import reach_b200 as rb  # noqa: E402
from reach_b200 import discretization as D  # noqa: E402
from oracle import lgg09_oracle as LO, glgm06_oracle as GO  # noqa: E402


def building():
    m = np.load(os.path.join(HERE, "models", "building.npz"))
    c = np.r_[np.full(10, 0.000225), np.zeros(38)]
    r = np.r_[np.full(10, 0.000025), np.zeros(14), 0.0001, np.zeros(23)]
    return m["A"], m["B"], rb.Hyperrectangle(c, r), rb.Interval(0.8, 1.0)


def lgg09_cases():
    out = {}
    # 1. Building BLDF01, Forward, δ = 0.004, directions ±e25, every 50th of 5000 steps
    A, B, X0, U0 = building()
    ivp = D.forward(A, X0, 0.004, D.normalize_input(B, U0))
    dirs = LO.vars_directions(25, 48)
    rho = LO.reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, 5000, ivp.V)
    out["bldf01_e25_every50"] = rho[:, ::50]
    out["bldf01_e25_max"] = rho.max(axis=1)
    # 2. Building, box template, δ = 0.002, first 300 steps (config C1 head)
    ivp = D.forward(A, X0, 0.002, D.normalize_input(B, U0))
    out["c1_head300"] = LO.reach_inhomog_LGG09(LO.box_directions(48), ivp.Omega0, ivp.Phi, 300, ivp.V)
    # 3. seeded random system, zonotopic Ω₀, mapped box input, oct template on 4 coordinates
    rng = np.random.default_rng(42)
    n = 9
    A = -np.eye(n) + 0.2 * rng.standard_normal((n, n))
    X0 = rb.Zonotope(rng.standard_normal(n), 0.1 * rng.standard_normal((n, 12)))
    U = rb.LinearMap(rng.standard_normal((n, 2)), rb.Hyperrectangle(np.array([0.1, -0.1]), np.array([0.05, 0.02])))
    ivp = D.no_bloating(A, X0, 0.05, U)
    dirs = np.vstack([rb.OctDirections(4).matrix(), np.zeros((n - 4, 32))])
    out["rand9_inputs_A"] = A
    out["rand9_c"], out["rand9_G"] = X0.center, X0.generators
    out["rand9_M"] = U.M
    out["rand9_oct"] = LO.reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, 80, ivp.V)
    return out


def glgm06_cases():
    out = {}
    rng = np.random.default_rng(7)
    n = 6
    A = -np.eye(n) + 0.3 * rng.standard_normal((n, n))
    X0 = rb.Hyperrectangle(rng.standard_normal(n), np.array([0.1, 0.0, 0.2, 0.1, 0.0, 0.3]))
    U = rb.LinearMap(rng.standard_normal((n, 2)), rb.Hyperrectangle(np.array([0.2, 0.0]), np.array([0.05, 0.1])))
    ivp = D.forward_zono(A, X0, 0.02, U)
    out["A"], out["x0c"], out["x0r"], out["M"] = A, X0.center, X0.radius, U.M
    F, sels, gaps = GO.post_GLGM06(ivp.Phi, ivp.Omega0, 60, 3, ivp.V, return_diag=True)
    for k in (0, 1, 2, 3, 10, 30, 59):
        out["c_%d" % k], out["G_%d" % k] = F[k]
        if sels[k] is not None:
            out["sel_%d" % k] = sels[k]
    out["min_gap"] = np.array(min(g for g in gaps if g is not None))
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "lgg09_golden.npz"), **lgg09_cases())
    np.savez_compressed(os.path.join(HERE, "glgm06_golden.npz"), **glgm06_cases())
    for f in ("lgg09_golden.npz", "glgm06_golden.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))
