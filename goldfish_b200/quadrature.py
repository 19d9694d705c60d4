"""Quadrature tables handed to the CUDA kernels (host side, setup time).

The reference integrates on tIGAr's extracted FE mesh: two triangles per knot
span, FFC default scheme of degree ``quad_deg`` (SURVEY.md Appendix A.1;
``quad_deg = quad_deg_const * p`` at /root/reference/GOLDFISH/tests/test_tbeam.py:31).
To reproduce its sums the kernels receive exactly those points, mapped to the
unit square of a span: degree <= 6 -> 12-point Strang-Fix/Dunavant rule,
degree > 6 -> collapsed Gauss-Jacobi with m = (deg+2)//2 points per axis.
"""
import numpy as np
from numpy.polynomial.legendre import leggauss
from scipy.special import roots_jacobi

_DUNAVANT6 = (
    (0.249286745170910, 0.116786275726379),
    (0.063089014491502, 0.050844906370207),
    (0.053145049844817, 0.310352451033784, 0.082851075618374),
)


def _triangle(deg):
    if deg <= 6:
        (a1, w1), (a2, w2), (a3, b3, w3) = _DUNAVANT6
        pts, wts = [], []
        for a, w in ((a1, w1), (a2, w2)):
            pts += [(a, a), (1 - 2 * a, a), (a, 1 - 2 * a)]
            wts += [w] * 3
        c3 = 1 - a3 - b3
        pts += [(a3, b3), (b3, a3), (a3, c3), (c3, a3), (b3, c3), (c3, b3)]
        wts += [w3] * 6
        return np.array(pts), 0.5 * np.array(wts)
    m = (deg + 2) // 2
    gs, ws = leggauss(m)
    s, ws = 0.5 * (gs + 1.0), 0.5 * ws
    gr, wr = roots_jacobi(m, 1.0, 0.0)
    r, wr = 0.5 * (gr + 1.0), 0.25 * wr
    pts = np.array([(si * (1.0 - rj), rj) for si in s for rj in r])
    wts = np.array([wi * wj for wi in ws for wj in wr])
    return pts, wts


def span_rule(deg):
    """(pts[nq,2], wts[nq] (sum 1), tw_lin[nq,4]) on the unit square of a knot
    span.  Triangle A = (v00, v10, v11), triangle B = (v00, v01, v11) (dolfin
    'right' diagonal, vertices sorted by index).  tw_lin are the CG1 (V_linear)
    barycentric weights of the span vertices (v00, v10, v01, v11)."""
    p, w = _triangle(deg)
    x, y = p[:, 0], p[:, 1]
    A = np.stack([x + y, y], axis=1)
    B = np.stack([y, x + y], axis=1)
    pts = np.concatenate([A, B])
    wts = np.concatenate([w, w])
    n = len(w)
    tw = np.zeros((2 * n, 4))
    X, Y = pts[:n, 0], pts[:n, 1]
    tw[:n, 0] = 1 - X; tw[:n, 1] = X - Y; tw[:n, 3] = Y
    X, Y = pts[n:, 0], pts[n:, 1]
    tw[n:, 0] = 1 - Y; tw[n:, 2] = Y - X; tw[n:, 3] = X
    return pts, wts, tw


def gauss01(m):
    g, w = leggauss(m)
    return 0.5 * (g + 1.0), 0.5 * w
