"""ORACLE (test infrastructure) -- quadrature rules of the reference FE detour.

tIGAr's ``ExtractedSpline(generator, quad_deg)`` with ``useRect=False``
(/root/reference/GOLDFISH/tests/test_tbeam.py:19,31-32) integrates on a mesh
with TWO TRIANGLES per knot span ("right" diagonal of dolfin.UnitSquareMesh)
using FFC's default scheme of degree ``quad_deg``:
  * degree <= 6 : tabulated symmetric rules (degree 6 -> 12 points, Strang-Fix/
    Dunavant),
  * degree  > 6 : collapsed Gauss-Jacobi with m = (deg+2)//2 points per axis
    (FIAT ``make_quadrature``), i.e. deg 9 -> 25, deg 12 -> 49 points/triangle.
SURVEY.md Appendix A.1.  Parity unpinned (FFC/FIAT not installed here): the
rules below are checked for polynomial exactness in tests/test_oracle_basics.py.

The element rule lives on the unit square [0,1]^2 of a knot span; the same
(table of points, weights) is what the CUDA path receives as input, so the two
sides integrate with identical points.
"""
import numpy as np
from scipy.special import roots_jacobi


def gauss_jacobi_01(m, alpha):
    """m-point Gauss-Jacobi rule on [0,1] for weight (1-r)^alpha."""
    x, w = roots_jacobi(m, alpha, 0.0)
    return 0.5 * (x + 1.0), w / 2.0 ** (alpha + 1.0)


def triangle_collapsed(m):
    """FIAT CollapsedQuadratureTriangleRule on the unit triangle
    {(x,y): x,y>=0, x+y<=1}: x = s(1-r), y = r, s~GJ(0,0), r~GJ(1,0)."""
    s, ws = gauss_jacobi_01(m, 0.0)
    r, wr = gauss_jacobi_01(m, 1.0)
    S, R = np.meshgrid(s, r, indexing="ij")
    WS, WR = np.meshgrid(ws, wr, indexing="ij")
    pts = np.stack([(S * (1.0 - R)).ravel(), R.ravel()], axis=1)
    wts = (WS * WR).ravel()
    return pts, wts


def _sym3(a):
    b = 1.0 - 2.0 * a
    return [(a, a), (b, a), (a, b)]


def _sym6(a, b):
    c = 1.0 - a - b
    return [(a, b), (b, a), (a, c), (c, a), (b, c), (c, b)]


def triangle_tabulated(deg):
    """Tabulated low-degree rules on the unit triangle (weights sum to 1/2)."""
    if deg <= 1:
        return np.array([[1 / 3.0, 1 / 3.0]]), np.array([0.5])
    if deg == 2:
        pts = np.array(_sym3(1.0 / 6.0))
        return pts, np.full(3, 1.0 / 6.0)
    if deg <= 6:
        # degree 3..6 all served here by the 12-point degree-6 rule
        # (Strang & Fix / Dunavant); only deg 6 (= 2p, p = 3, Scordelis-Lo
        # fixture /root/reference/GOLDFISH/tests/test_slr.py:36) is used.
        a1, w1 = 0.249286745170910, 0.116786275726379
        a2, w2 = 0.063089014491502, 0.050844906370207
        a3, b3, w3 = 0.053145049844817, 0.310352451033784, 0.082851075618374
        pts = np.array(_sym3(a1) + _sym3(a2) + _sym6(a3, b3))
        wts = 0.5 * np.array([w1] * 3 + [w2] * 3 + [w3] * 6)
        return pts, wts
    raise ValueError(deg)


def triangle_rule(deg):
    if deg <= 6:
        return triangle_tabulated(deg)
    return triangle_collapsed((deg + 2) // 2)


def element_rule(deg):
    """Rule on the unit square of one knot span = the two dolfin triangles.

    Span vertices v00=(0,0), v10=(1,0), v01=(0,1), v11=(1,1).  dolfin sorts a
    cell's vertices by global index, so triangle A = (v00, v10, v11) and
    triangle B = (v00, v01, v11); reference vertex (0,0)->first, (1,0)->second,
    (0,1)->third.  Returns pts (nq,2), wts (nq,) (sum = 1), tri (nq,) in {0,1}.
    """
    p, w = triangle_rule(deg)
    x, y = p[:, 0], p[:, 1]
    # A: v00 + x (v10 - v00) + y (v11 - v00) = (x + y, y)
    A = np.stack([x + y, y], axis=1)
    # B: v00 + x (v01 - v00) + y (v11 - v00) = (y, x + y)
    B = np.stack([y, x + y], axis=1)
    pts = np.concatenate([A, B])
    wts = np.concatenate([w, w])  # |det| = 1 for both affine maps
    tri = np.concatenate([np.zeros(len(w), int), np.ones(len(w), int)])
    return pts, wts, tri


def gauss_legendre_01(m):
    return gauss_jacobi_01(m, 0.0)
