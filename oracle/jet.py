"""ORACLE (test infrastructure) -- second-order forward-mode AD ("jets").

The reference never hand-derives anything: residual, tangent and the adjoint
partials are UFL symbolic derivatives of ONE energy expression
(``derivative(...)`` at /root/reference/GOLDFISH/nonmatching_opt.py:440,449 and
/root/reference/GOLDFISH/operations/int_energy_exop.py:34,46,53).  The oracle
mirrors that: the energy is written once (kl_shell.py, penalty.py) and its
gradient/Hessian come from this exact (to round-off) Taylor arithmetic,
vectorised over quadrature points.  The CUDA path differentiates by other
means (hand-derived first variation + lane-parallel dual numbers), so the two
are independent derivations of the same quantity.
"""
import numpy as np


class Jet:
    __slots__ = ("v", "g", "h")
    __array_ufunc__ = None  # make ndarray (op) Jet defer to Jet.__r*__

    def __init__(self, v, g, h):
        self.v, self.g, self.h = v, g, h

    # -- construction -------------------------------------------------------
    @staticmethod
    def variables(vals):
        """vals: (n, m) -> list of m independent Jets."""
        n, m = vals.shape
        out = []
        for k in range(m):
            g = np.zeros((n, m))
            g[:, k] = 1.0
            out.append(Jet(vals[:, k].copy(), g, np.zeros((n, m, m))))
        return out

    def _const_like(self, c):
        c = np.broadcast_to(np.asarray(c, dtype=np.float64), self.v.shape)
        return Jet(c.copy(), np.zeros_like(self.g), np.zeros_like(self.h))

    # -- arithmetic -----------------------------------------------------------
    def __add__(self, o):
        if isinstance(o, Jet):
            return Jet(self.v + o.v, self.g + o.g, self.h + o.h)
        return Jet(self.v + o, self.g, self.h)

    __radd__ = __add__

    def __neg__(self):
        return Jet(-self.v, -self.g, -self.h)

    def __sub__(self, o):
        if isinstance(o, Jet):
            return Jet(self.v - o.v, self.g - o.g, self.h - o.h)
        return Jet(self.v - o, self.g, self.h)

    def __rsub__(self, o):
        return (-self) + o

    def __mul__(self, o):
        if isinstance(o, Jet):
            gg = self.g[:, :, None] * o.g[:, None, :]
            return Jet(self.v * o.v,
                       self.v[:, None] * o.g + o.v[:, None] * self.g,
                       self.v[:, None, None] * o.h + o.v[:, None, None] * self.h
                       + gg + gg.transpose(0, 2, 1))
        o = np.asarray(o, dtype=np.float64)
        if o.ndim == 0:
            return Jet(self.v * o, self.g * o, self.h * o)
        return Jet(self.v * o, self.g * o[:, None], self.h * o[:, None, None])

    __rmul__ = __mul__

    def _unary(self, f, f1, f2):
        gg = self.g[:, :, None] * self.g[:, None, :]
        return Jet(f, f1[:, None] * self.g,
                   f1[:, None, None] * self.h + f2[:, None, None] * gg)

    def recip(self):
        r = 1.0 / self.v
        return self._unary(r, -r * r, 2.0 * r * r * r)

    def sqrt(self):
        s = np.sqrt(self.v)
        return self._unary(s, 0.5 / s, -0.25 / (s * self.v))

    def __truediv__(self, o):
        if isinstance(o, Jet):
            return self * o.recip()
        return self * (1.0 / np.asarray(o, dtype=np.float64))

    def __rtruediv__(self, o):
        return self.recip() * o

    def __pow__(self, k):
        if k == 2:
            return self * self
        if k == 3:
            return self * self * self
        raise NotImplementedError


# -- small vector helpers on lists of Jets ------------------------------------
def dot(a, b):
    s = a[0] * b[0]
    for x, y in zip(a[1:], b[1:]):
        s = s + x * y
    return s


def cross(a, b):
    return [a[1] * b[2] - a[2] * b[1],
            a[2] * b[0] - a[0] * b[2],
            a[0] * b[1] - a[1] * b[0]]


def norm(a):
    d = dot(a, a)
    return d.sqrt() if isinstance(d, Jet) else np.sqrt(d)


def unit(a):
    r = 1.0 / norm(a)
    return [x * r for x in a]


def axpy(al, x, y):
    return [al * xi + yi for xi, yi in zip(x, y)]


def scale(al, x):
    return [al * xi for xi in x]


def vsub(x, y):
    return [a - b for a, b in zip(x, y)]


def vadd(x, y):
    return [a + b for a, b in zip(x, y)]
