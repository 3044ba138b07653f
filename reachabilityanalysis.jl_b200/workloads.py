"""The benchmark configurations of BASELINE.json as (Φ, Ω₀, V, directions) tuples, generated on
the host.  Input generation is not on the timed path."""
import os
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

from . import sets as S
from . import discretization as D
from .directions import BoxDirections

_MODELS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "models")


@dataclass
class Workload:
    name: str
    Phi: np.ndarray
    Omega0: object
    V: object
    dirs: np.ndarray
    nsteps: int
    delta: float

    @property
    def n(self):
        return self.Phi.shape[0]


def _model(name):
    return np.load(os.path.join(_MODELS, name + ".npz"))


def _csc(m):
    return sp.csc_matrix((m["A_data"], m["A_indices"], m["A_indptr"]), shape=tuple(m["A_shape"]))


def building_sets():
    """examples/Building/Building.jl:58-62."""
    m = _model("building")
    c = np.r_[np.full(10, 0.000225), np.zeros(38)]
    r = np.r_[np.full(10, 0.000025), np.zeros(14), 0.0001, np.zeros(23)]
    return m["A"], m["B"], S.Hyperrectangle(c, r), S.Interval(0.8, 1.0)


def building_c1(delta=0.002, nsteps=10000):
    """C1: BLDF01, Forward model, box template (96 directions), δ=0.002, T=20."""
    A, B, X0, U0 = building_sets()
    ivp = D.forward(A, X0, delta, D.normalize_input(B, U0))
    return Workload("C1 Building BLDF01 n=48 box", ivp.Phi, ivp.Omega0, ivp.V, BoxDirections(48).matrix(), nsteps, delta)


def iss_c2(delta=6e-4, nsteps=33334):
    """C2: ISSF01 (examples/ISS/ISS.jl:81-91), Forward model, box template (540 directions)."""
    m = _model("iss")
    A = _csc(m).toarray()
    X0 = S.BallInf(np.zeros(270), 0.0001)
    U0 = S.Hyperrectangle(low=[0.0, 0.8, 0.9], high=[0.1, 1.0, 1.0])
    ivp = D.forward(A, X0, delta, D.normalize_input(m["B"], U0))
    return Workload("C2 ISS ISSF01 n=270 box", ivp.Phi, ivp.Omega0, ivp.V, BoxDirections(270).matrix(), nsteps, delta)


def heat_c3(delta=0.004, nsteps=10000):
    """C3: HEAT02 (n=1000), homogeneous.  The reference ships only A and `xind` (no IVP), so the
    initial set follows the ARCH-COMP convention: heated cells (xind) in [0.9, 1.1], the rest 0;
    directions = the centre cell followed by the box template."""
    m = _model("heat02")
    A = _csc(m).toarray()
    x = m["xind"]
    X0 = S.Hyperrectangle(np.where(x != 0, 1.0, 0.0), np.where(x != 0, 0.1, 0.0))
    ivp = D.forward(A, X0, delta)
    centre = np.zeros((1000, 1))
    centre[555 - 1, 0] = 1.0          # cell (5,5,5) of the 10x10x10 mesh, 1-based
    dirs = np.hstack([centre, BoxDirections(1000).matrix()])
    return Workload("C3 Heat3D HEAT02 n=1000", ivp.Phi, ivp.Omega0, None, dirs, nsteps, delta)


def synthetic_c5(n=2000, delta=0.01, nsteps=10000, seed=0):
    """C5: A = −I + (0.5/√n)·N(0,1), Φ = expm(Aδ), X₀ = box(c~U(−1,1), r~0.1·U(0,1)), B = I,
    U₀ = BallInf(0, 0.05), Forward model, box template (2n directions).  The Forward model's
    3n×3n block exponential is replaced by its (rapidly converging) Taylor series Φ₂(|A|,δ) =
    Σ δ^{i+2}/(i+2)! |A|^i, which is what that block computes."""
    rng = np.random.default_rng(seed)
    A = -np.eye(n) + (0.5 / np.sqrt(n)) * rng.standard_normal((n, n))
    c, r = rng.uniform(-1, 1, n), 0.1 * rng.uniform(0, 1, n)
    X0 = S.Hyperrectangle(c, r)
    U = S.BallInf(np.zeros(n), 0.05)
    Ad = A * delta
    # Φ = expm(Aδ) by scaling-free Taylor (‖Aδ‖ ≈ 0.015)
    Phi = np.eye(n)
    term = np.eye(n)
    for i in range(1, 14):
        term = term @ Ad / i
        Phi = Phi + term
    Aabs = np.abs(A)
    P2 = np.eye(n) * (delta ** 2 / 2)
    term = np.eye(n) * (delta ** 2 / 2)
    for i in range(1, 12):
        term = term @ (Aabs * delta) / (i + 2)
        P2 = P2 + term
    A2 = A @ A
    rE = P2 @ (np.abs(A2 @ c) + np.abs(A2) @ r)                       # E⁺ (ForwardModule.jl:147)
    rpsi = P2 @ (np.abs(A) @ np.full(n, 0.05))                        # Eψ (:150), U centred at 0
    Eplus = S.Hyperrectangle(np.zeros(n), rE)
    V = S.MinkowskiSum(S.Scale(delta, U), S.Hyperrectangle(np.zeros(n), rpsi))
    Omega0 = S.ConvexHull(X0, S.MinkowskiSum(S.MinkowskiSum(S.LinearMap(Phi, X0), V), Eplus))
    return Workload("C5 synthetic stable LTI n=%d box" % n, Phi, Omega0, V, BoxDirections(n).matrix(), nsteps, delta)


def building_c4(nsplits_per_dim=2, nsteps=1000, delta=0.002):
    """C4: Building BLDF01, X0 split into 2^10 = 1024 boxes along its 10 non-flat leading dimensions
    (`split(X0, [2×10, 1×38])`, first dimension fastest), GLGM06 with max_order = 10 and the
    Forward(setops=:zono) model (the literal default FirstOrderZonotope is
    arithmetically valid but useless at this ‖A‖δ).  Returns (Φ, [Ω₀ per split], V)."""
    A, B, X0, U0 = building_sets()
    parts = [nsplits_per_dim] * 10 + [1] * 38
    splits = S.split(X0, parts)
    U = D.normalize_input(B, U0)
    ivps = [D.forward_zono(A, X, delta, U) for X in splits]
    return ivps[0].Phi, [i.Omega0 for i in ivps], ivps[0].V
