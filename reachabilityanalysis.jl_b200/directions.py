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


class BoxDirections(AbstractDirections):
    """+e₁..+e_n then −e₁..−e_n (LGG09.jl:112)."""

    def __init__(self, n):
        self.n = int(n)

    def matrix(self):
        I = np.eye(self.n)
        return np.hstack([I, -I])


class OctDirections(AbstractDirections):
    """±e_i ± e_j for i < j, pair by pair in the order (+,+), (+,−), (−,+), (−,−), followed by the
    box directions (LazySets' iteration order; LGG09.jl:113)."""

    def __init__(self, n):
        self.n = int(n)

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
        return np.hstack([D, BoxDirections(n).matrix()])


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
