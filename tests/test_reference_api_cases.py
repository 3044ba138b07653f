"""The struct / constructor cases of the reference's own algorithm tests, restated for the host mirror
(CPU only: no device call).  Each test names the reference lines it follows; `ArgumentError` maps to
`ValueError`."""
import numpy as np
import pytest

import reach_b200 as rb
from reach_b200.directions import get_template
from reach_b200.lgg09 import step_size


def _sev(i, n, v):
    d = np.zeros(n)
    d[i - 1] = v
    return d


def test_lgg09_struct_lgg09_jl_5_28():
    alg = rb.LGG09(δ=0.01, vars=(1,), dim=1)                       # :14
    with pytest.raises(ValueError):                                # :17
        rb.LGG09(δ=0.01)
    rb.LGG09(δ=0.01, dirs=[1.0])                                   # :18
    rb.LGG09(δ=0.01, template=rb.BoxDirections(1))                 # :19
    with pytest.raises(ValueError):                                # :20
        rb.LGG09(δ=0.01, vars=(1,))
    with pytest.raises(ValueError):                                # :21
        rb.LGG09(δ=0.01, dim=1)
    assert step_size(alg) == 0.01                                  # :24
    assert alg.template.matrix().dtype == np.float64               # :25 numtype


def test_lgg09_template_construction_lgg09_jl_30_66():
    dirs = np.zeros(5)
    dirs[0] = 1.0
    alg = rb.LGG09(δ=0.01, template=dirs)                          # :32-36 a single vector
    assert alg.template == rb.CustomDirections([dirs])
    two = [_sev(1, 5, 1.0), _sev(1, 5, -1.0)]
    assert rb.LGG09(δ=0.01, template=two).template == rb.CustomDirections(two)    # :39-41
    assert rb.LGG09(δ=0.01, dirs=two).template == rb.CustomDirections(two)        # :44-46 alias
    with pytest.raises(ValueError):                                # :49
        rb.LGG09(δ=0.01, vars=(1))
    alg = rb.LGG09(δ=0.01, vars=(1), n=5)                          # :50-51
    assert [list(d) for d in alg.template] == [list(d) for d in two]
    d13 = [_sev(1, 5, 1.0), _sev(3, 5, 1.0), _sev(1, 5, -1.0), _sev(3, 5, -1.0)]
    alg = rb.LGG09(δ=0.01, vars=(1, 3), n=5)                       # :53-56
    assert [list(d) for d in alg.template] == [list(d) for d in d13]
    alg2 = rb.LGG09(δ=0.01, vars=(1, 3), dim=5)                    # :59-60 alias
    assert [list(d) for d in alg2.template] == [list(d) for d in d13]
    with pytest.raises((ValueError, AssertionError)):              # LGG09.jl:119 m <= n
        rb.LGG09(δ=0.01, vars=(1, 2, 3), n=2)


def test_lgg09_equivalent_definitions_lgg09_jl_157_165():
    box5d = rb.BoxDirections(5)
    alg0 = rb.LGG09(δ=0.01, template=box5d)
    alg1 = rb.LGG09(δ=0.01, template="box", n=5)
    alg2 = rb.LGG09(δ=0.01, template="box", dim=5)
    assert alg0 == alg1 == alg2                                    # :160
    assert [list(d) for d in alg0.template] == [list(d) for d in alg1.template]   # :162
    alg3 = rb.LGG09(δ=0.01, template="oct", dim=5)                 # :165-166
    assert isinstance(alg3.template, rb.OctDirections) and alg3.template.dim == 5
    assert alg0 != alg3
    assert alg0 != rb.LGG09(δ=0.02, template=box5d)
    assert alg0 != rb.LGG09(δ=0.01, template=box5d, approx_model=rb.NoBloating())
    assert alg0 == rb.LGG09(δ=0.01, template=box5d, approx_model=rb.Forward(sih="concrete", exp="base", setops="lazy"))


def test_lgg09_accepts_the_reference_options_lgg09_jl_92_105():
    box1d = rb.CustomDirections([[1.0], [-1.0]])
    for kw in ({"sparse": True}, {"threaded": True}, {"cache": False}, {"static": False}):
        alg = rb.LGG09(δ=0.01, template=box1d, **kw)
        assert getattr(alg, next(iter(kw))) == next(iter(kw.values()))
        assert len(alg.template) == 2 and alg.template.dim == 1


def test_polar_directions_lgg09_jl_184_188():
    t = rb.PolarDirections(40)
    M = t.matrix()
    assert M.shape == (2, 40) and t.dim == 2 and len(t) == 40
    np.testing.assert_allclose(np.hypot(M[0], M[1]), 1.0, rtol=0, atol=1e-15)
    np.testing.assert_array_equal(M[:, 0], [1.0, 0.0])
    np.testing.assert_allclose(M[:, 10], [0.0, 1.0], atol=1e-15)           # φ = π/2
    np.testing.assert_allclose(M[:, 20], -M[:, 0], atol=1e-15)            # antipodal pairs for even Nφ
    assert rb.LGG09(δ=0.2, template=t).template is t
    assert rb.PolarDirections(40) == t and rb.PolarDirections(41) != t
    with pytest.raises(ValueError):
        rb.PolarDirections(0)


def test_template_equality_is_by_value():
    assert rb.BoxDirections(3) == rb.BoxDirections(3) and rb.BoxDirections(3) != rb.BoxDirections(4)
    assert rb.OctDirections(3) == rb.OctDirections(3) and rb.OctDirections(3) != rb.BoxDirections(3)
    assert rb.CustomDirections(list(rb.BoxDirections(3))) == rb.BoxDirections(3)     # same directions, same order
    assert get_template("box", 3) == rb.BoxDirections(3)
    with pytest.raises(ValueError):
        get_template("box")
    with pytest.raises(ValueError):
        get_template("hex", 3)
    with pytest.raises(ValueError):                                # CustomDirections: directions of one dimension
        rb.CustomDirections([[1.0, 0.0], [1.0]])


def test_glgm06_struct_glgm06_jl_5_34():
    alg = rb.GLGM06(δ=0.01)                                        # :10
    assert alg.δ == 0.01 and alg.delta == 0.01                     # :13
    assert alg.max_order == 5 and alg.preallocate is True          # GLGM06.jl:66-81 defaults
    assert alg.reduction_method == "GIR05" and isinstance(alg.approx_model, rb.FirstOrderZonotope)
    rb.GLGM06(δ=0.01, static=True, dim=1, ngens=10)                # :31 accepted (dense storage on the device)
    rb.GLGM06(δ=0.01, preallocate=False)                           # :76
    rb.GLGM06(δ=0.01, approx_model=rb.NoBloating())                # :65
    assert rb.GLGM06(δ=1e-3, static=True, dim=8, max_order=1, ngens=8).max_order == 1   # :101
    assert rb.GLGM06(δ=0.01) == rb.GLGM06(δ=0.01) and rb.GLGM06(δ=0.01) != rb.GLGM06(δ=0.01, max_order=6)
    with pytest.raises(TypeError):
        rb.GLGM06()
    with pytest.raises(ValueError):
        rb.GLGM06(δ=0.01, max_order=0)
    with pytest.raises(NotImplementedError):
        rb.GLGM06(δ=0.01, reduction_method="COMB03")


def test_box_struct_box_jl_5_27():
    alg = rb.BOX(δ=0.01)                                           # :10
    assert alg.step_size == 0.01 and alg.δ == 0.01                 # :13
    assert alg.recursive is False and alg.static is False
    assert alg.approx_model == rb.Forward(sih="concrete", exp="base", setops="lazy")   # BOX.jl:60
    assert rb.BOX(δ=0.01, static=True, dim=1).dim == 1             # :19
    assert rb.BOX(δ=0.01, recursive=True).recursive is True        # :83
    with pytest.raises(TypeError):
        rb.BOX()


def test_box_and_oct_templates_in_lazysets_iteration_order():
    """Opt-in row order for callers that compare with a Julia run ([MEM]: BoxDirections(2) collects to
    [1,0], [0,1], [0,-1], [-1,0]); the default order of the mirror and of the oracle is unchanged."""
    np.testing.assert_array_equal(rb.BoxDirections(2, lazysets_order=True).matrix().T, [[1, 0], [0, 1], [0, -1], [-1, 0]])
    np.testing.assert_array_equal(rb.BoxDirections(2).matrix().T, [[1, 0], [0, 1], [-1, 0], [0, -1]])
    M = rb.OctDirections(3, lazysets_order=True).matrix()
    assert M.shape == (3, 18)
    np.testing.assert_array_equal(M[:, :4].T, [[1, 1, 0], [1, -1, 0], [-1, 1, 0], [-1, -1, 0]])
    np.testing.assert_array_equal(M[:, 12:], rb.BoxDirections(3, lazysets_order=True).matrix())
    a, b = rb.BoxDirections(4, lazysets_order=True).matrix(), rb.BoxDirections(4).matrix()
    assert sorted(map(tuple, a.T)) == sorted(map(tuple, b.T))          # the same set of directions
    assert rb.BoxDirections(3) != rb.BoxDirections(3, lazysets_order=True)
