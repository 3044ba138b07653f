"""The C/OpenMP restatement (oracle/c/reach_ref.c -- checker and timed CPU arm) against the NumPy oracle, which in
turn reproduces the reference's known-answer tests (test_oracle_kat.py).  CPU only."""
import os

import numpy as np
import pytest

import reach_b200 as rb
from reach_b200 import discretization as D, workloads as W
from oracle import cref, lgg09_oracle as LO, glgm06_oracle as GO, lazysets_oracle as LZ
from conftest import assert_support_parity, model

HERE = os.path.dirname(os.path.abspath(__file__))


def _rand_problem(rng, n, with_input=True, zono=False):
    A = -np.eye(n) + (0.6 / np.sqrt(n)) * rng.standard_normal((n, n))
    if zono:
        X0 = rb.Zonotope(rng.standard_normal(n), 0.1 * rng.standard_normal((n, n + 3)))
    else:
        X0 = rb.Hyperrectangle(rng.uniform(-1, 1, n), 0.1 * rng.uniform(0, 1, n))
    U = None
    if with_input:
        M = rng.standard_normal((n, 2))
        U = rb.LinearMap(M, rb.Hyperrectangle(np.array([0.1, -0.2]), np.array([0.05, 0.02])))
    return D.forward(A, X0, 0.02, U)


@pytest.mark.parametrize("shape,kw", [(1, {}), (1, {"literal": True}), (1, {"sparse": True}), (1, {"share": True}),
                                      (2, {}), (2, {"share": True})])
@pytest.mark.parametrize("zono", [False, True])
def test_lgg09_shapes_match_numpy_oracle(shape, kw, zono):
    rng = np.random.default_rng(3)
    ivp = _rand_problem(rng, 13, zono=zono)
    dirs = rb.BoxDirections(13).matrix()
    ref = LO.reach_inhomog_LGG09(dirs, ivp.Omega0, ivp.Phi, 120, ivp.V)
    got = cref.reach_lgg09(dirs, ivp.Omega0, ivp.Phi, 120, ivp.V, shape=shape, **kw)
    assert_support_parity(got, ref, tol=1e-12)


def test_lgg09_homogeneous_zero_set_and_custom_directions():
    rng = np.random.default_rng(5)
    ivp = _rand_problem(rng, 7, with_input=False)
    dirs = rng.standard_normal((7, 5))
    ref = LO.reach_homog_LGG09(dirs, ivp.Omega0, ivp.Phi, 60)
    for shape in (1, 2):
        assert_support_parity(cref.reach_lgg09(dirs, ivp.Omega0, ivp.Phi, 60, None, shape=shape), ref, tol=1e-12)
    ivp = _rand_problem(rng, 7)
    ref = LO.reach_inhomog_LGG09(dirs, rb.ZeroSet(7), ivp.Phi, 60, ivp.V)          # reach_inhomog.jl:66-81
    assert_support_parity(cref.reach_lgg09(dirs, rb.ZeroSet(7), ivp.Phi, 60, ivp.V, shape=1), ref, tol=1e-12)
    with pytest.raises(RuntimeError):                                                  # share needs [D, -D]
        cref.reach_lgg09(dirs[:, :4], ivp.Omega0, ivp.Phi, 10, ivp.V, shape=1, share=True)


def test_lgg09_building_kat_through_c():
    """test/models/generated/Building.jl:44-47 through the C restatement (same bounds the NumPy oracle is pinned on)."""
    m = model("building")
    c = np.r_[np.full(10, 0.000225), np.zeros(38)]
    r = np.r_[np.full(10, 0.000025), np.zeros(14), 0.0001, np.zeros(23)]
    ivp = D.forward(m["A"], rb.Hyperrectangle(c, r), 0.004, D.normalize_input(m["B"], rb.Interval(0.8, 1.0)))
    dirs = LO.vars_directions(25, 48)
    rho = cref.reach_lgg09(dirs, ivp.Omega0, ivp.Phi, 5000, ivp.V, shape=1)
    assert 4e-3 < rho[0].max() <= 5.1e-3
    assert rho[0, -1] > -0.78e-3
    np.testing.assert_allclose(rho[0].max(), 4.8602388967852826e-3, rtol=1e-10)
    np.testing.assert_allclose(rho[0, -1], 1.0835336749422586e-3, rtol=1e-10)


def test_lgg09_sparse_iss_matches_dense():
    wl = W.iss_c2(nsteps=150)
    a = cref.reach_lgg09(wl.dirs, wl.Omega0, wl.Phi, 150, wl.V, shape=1, sparse=True, share=True)
    b = LO.reach_inhomog_LGG09(wl.dirs, wl.Omega0, wl.Phi, 150, wl.V)
    assert_support_parity(a, b, tol=1e-12)


def test_gir05_matches_numpy_including_ties():
    rng = np.random.default_rng(11)
    for n, p, r in [(5, 40, 3), (4, 9, 2), (6, 30, 1), (3, 5, 4)]:
        G = rng.standard_normal((n, p))
        G[:, 3] = 0.0
        G[:, 7 % p] = np.eye(n)[:, 0] * 0.3            # axis-aligned: weight exactly 0, tie with the zero column
        G[:, 5 % p] = np.eye(n)[:, 1] * 0.7
        _, want, keep = LZ.reduce_order_gir05(np.zeros(n), G, r, return_selection=True)
        got, sel = cref.reduce_order_gir05(G, r)
        np.testing.assert_allclose(got, want, rtol=1e-14, atol=1e-15)   # r = 1: NumPy sums pairwise
        if keep is None:
            assert sel is None
        else:
            np.testing.assert_array_equal(sel, keep)


def test_glgm06_batch_matches_numpy_oracle():
    Phi, Oms, V = W.building_c4(2, 40)
    zs = [LZ.reduce_order_gir05(*LZ.to_zonotope(Om), 10) for Om in Oms[:3]]
    cU, GU = LZ.to_zonotope(V)
    c0 = np.stack([z[0] for z in zs], axis=1)
    d = np.zeros(48)
    d[24] = 1.0
    samples = [(s, k) for s in range(3) for k in (0, 1, 7, 39)]
    for shared in (True, False):
        rho, outs = cref.reach_inhomog_glgm06(Phi, c0, [z[1] for z in zs], 40, 10, cU, GU, d=d, samples=samples,
                                              shared_chain=shared)
        for s in range(3):
            ref, sels, gaps = GO.post_GLGM06(Phi, Oms[s], 40, 10, V, return_diag=True)
            assert min(g for g in gaps if g is not None) > 1e-9
            for (ss, k), (c, G, sel) in zip(samples, outs):
                if ss != s:
                    continue
                np.testing.assert_allclose(c, ref[k][0], rtol=0, atol=1e-13)
                np.testing.assert_allclose(G, ref[k][1], rtol=0, atol=1e-13)
                if sels[k] is not None:
                    np.testing.assert_array_equal(sel[sel >= 0], sels[k])
            want = np.array([LZ.rho_zonotope(d, *ref[k]) for k in range(40)])
            np.testing.assert_allclose(rho[s], want, rtol=1e-12, atol=1e-15)
