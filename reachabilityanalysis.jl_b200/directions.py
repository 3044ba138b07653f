"""Template direction sets (`AbstractDirections` of LazySets as used by LGG09.jl:105-131)."""
import numpy as np


class AbstractDirections:
    n: int

    def matrix(self):
        """n x ndirs array, one column per direction, in iteration order."""
        raise NotImplementedError

    def __len__(self):
        return self.matrix().shape[1]

    def __iter__(self):
        M = self.matrix()
        return (M[:, j] for j in range(M.shape[1]))

    @property
    def dim(self):
        return self.n

    def __eq__(self, other):
        """Templates compare by value (`alg0 == alg1 == alg2`, test/algorithms/LGG09.jl:158-159)."""
        if not isinstance(other, AbstractDirections):
            return NotImplemented
        if type(self) is type(other) and type(self) in (BoxDirections, OctDirections, PolarDirections):
            return self._key() == other._key()
        return self.n == other.n and np.array_equal(self.matrix(), other.matrix())

    __hash__ = None

    def _key(self):
        return self.n


class BoxDirections(AbstractDirections):
    """+e₁..+e_n then −e₁..−e_n (LGG09.jl:112).

    The rows of ρℓ follow the iteration order of the template.  The device path itself is order-agnostic (it takes
    the direction matrix), and the Julia wrapper passes `collect(template)`, i.e. LazySets' own order, which for the
    negative half is believed to be −e_n..−e₁ ([MEM]: LazySets is not vendored in the reference tree and no reference
    test pins it).  This mirror and the oracle default to the order stated in the first line; `lazysets_order=True`
    gives +e₁..+e_n, −e_n..−e₁ for callers that need the rows to line up with a Julia run."""

    def __init__(self, n, lazysets_order=False):
        self.n = int(n)
        self.lazysets_order = bool(lazysets_order)

    def matrix(self):
        I = np.eye(self.n)
        return np.hstack([I, -I[:, ::-1]]) if self.lazysets_order else np.hstack([I, -I])

    def _key(self):
        return (self.n, self.lazysets_order)


class OctDirections(AbstractDirections):
    """±e_i ± e_j for i < j, pair by pair in the order (+,+), (+,−), (−,+), (−,−), followed by the
    box directions (LazySets' iteration order; LGG09.jl:113).  `lazysets_order` is passed to the box part."""

    def __init__(self, n, lazysets_order=False):
        self.n = int(n)
        self.lazysets_order = bool(lazysets_order)

    def matrix(self):
        n = self.n
        cols = []
        for i in range(n):
            for j in range(i + 1, n):
                for si, sj in ((1, 1), (1, -1), (-1, 1), (-1, -1)):
                    d = np.zeros(n)
                    d[i], d[j] = si, sj
                    cols.append(d)
        D = np.array(cols).T if cols else np.zeros((n, 0))
        return np.hstack([D, BoxDirections(n, self.lazysets_order).matrix()])

    def _key(self):
        return (self.n, self.lazysets_order)


class PolarDirections(AbstractDirections):
    """`PolarDirections(Nφ)` of LazySets ([MEM]): the Nφ planar unit vectors (cos φ_i, sin φ_i) with
    φ = range(0, 2π, length = Nφ + 1)[1:Nφ] (test/algorithms/LGG09.jl:184-188 runs LGG09 over `PolarDirections(40)`)."""

    def __init__(self, Nphi):
        if int(Nphi) <= 0:
            raise ValueError("Nφ = %s is invalid; it should be at least 1" % (Nphi,))
        self.Nphi = int(Nphi)
        self.n = 2

    def matrix(self):
        phi = np.linspace(0.0, 2.0 * np.pi, self.Nphi + 1)[: self.Nphi]
        return np.vstack([np.cos(phi), np.sin(phi)])

    def _key(self):
        return self.Nphi


class CustomDirections(AbstractDirections):
    def __init__(self, directions, n=None):
        dirs = [np.asarray(d, dtype=np.float64).reshape(-1) for d in directions]
        if not dirs:
            raise ValueError("at least one direction is required")
        self.n = int(n) if n is not None else len(dirs[0])
        for d in dirs:
            if len(d) != self.n:
                raise ValueError("direction of length %d in a %d-dimensional template" % (len(d), self.n))
        self._M = np.array(dirs).T.copy()

    def matrix(self):
        return self._M


def template_from_vars(vars_, n):
    """`_get_template(vars, n)`: [+e_v …, −e_v …] (LGG09.jl:117-131, 1-based variables; the order
    is pinned by test/algorithms/LGG09.jl:52-55)."""
    vs = [vars_] if np.isscalar(vars_) else list(vars_)
    if len(vs) > n:
        raise ValueError("the number of variables should not exceed `n`")
    m = len(vs)
    D = np.zeros((n, 2 * m))
    for i, v in enumerate(vs):
        if not 1 <= v <= n:
            raise ValueError("variable %d outside 1..%d" % (v, n))
        D[v - 1, i] = 1.0
        D[v - 1, i + m] = -1.0
    return CustomDirections(list(D.T), n)


def get_template(template, n=None):
    """`_get_template` dispatch (LGG09.jl:105-115)."""
    if isinstance(template, AbstractDirections):
        return template
    if isinstance(template, str):
        if n is None:
            raise ValueError("the ambient dimension, `n` (or `dim`), should be specified")
        if template == "box":
            return BoxDirections(n)
        if template == "oct":
            return OctDirections(n)
        raise ValueError("unknown template %r" % template)
    arr = np.asarray(template, dtype=np.float64)
    if arr.ndim == 1:
        return CustomDirections([arr])
    return CustomDirections(list(template))
