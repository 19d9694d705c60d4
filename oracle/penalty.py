"""ORACLE (test infrastructure) -- penalty coupling energy on intersection curves.

Restates PENGoLINS ``penalty_energy`` (hanzhao2020/PENGoLINS, un-vendored and
unpinned; reference call site /root/reference/GOLDFISH/nonmatching_opt.py:1109-1116,
setup :422-431, residual/tangent transfer :745-752, :789-801).  Published
formulation: Herrema et al. 2019, Zhao et al. 2022:

  PE = int_Gamma  alpha_d/2 |u^A - u^B|^2
                + alpha_r/2 [ (a3^A.a3^B - A3^A.A3^B)^2 + (an^A.a3^B - An^A.A3^B)^2 ] dGamma
  an = at x a3,  at = unit tangent of the curve pushed forward from patch A,
  alpha_d = alpha E t /(h (1-nu^2)),  alpha_r = alpha E t^3 /(12 h (1-nu^2)).

Discretisation restated here (hypotheses H-pen-*, documented in DESIGN.md,
parity unpinned because PENGoLINS is not installed):
  H-pen-1  mortar fields are CG1 on a 1-D mesh and the measure is VERTEX
           quadrature of degree 0 (docstring :26-29) => every mortar cell c
           contributes at its two end vertices with weight ell_c/2, where
           ell_c = |X^A(v_{c+1}) - X^A(v_c)| is the chord of the CG1
           interpolant of the physical curve (the "line Jacobian").
  H-pen-2  the parametric tangent is cell-wise constant, taken from the side-A
           parametric coordinates of the cell's two vertices.
  H-pen-3  alpha_d, alpha_r are frozen at setup (used as stored, :1115) with
           "minimum" over the two sides (:424) and h = mean of the two sides'
           element sizes.
"""
import numpy as np
from .jet import Jet, dot, cross, unit, scale, vsub, vadd


def _sqrt(x):
    return x.sqrt() if isinstance(x, Jet) else np.sqrt(x)


def penalty_point_energy(uA, duA, uB, duB, XA0, XA1, dXA, dXB, tpar,
                         alpha_d, alpha_r):
    """Energy of one (cell, end-vertex) evaluation.

    uA,uB: 3-vectors; duA,duB,dXA,dXB: [.,1 (3), .,2 (3)];
    XA0, XA1: physical position (side A) of the cell's two vertices;
    tpar: (n,2) parametric tangent on side A (plain numbers).
    """
    chord = vsub(XA1, XA0)
    ell = _sqrt(dot(chord, chord))
    w = 0.5 * ell

    def frame(d, with_tangent):
        g1, g2 = d
        a3 = unit(cross(g1, g2))
        if not with_tangent:
            return a3, None
        at = unit(vadd(scale(tpar[:, 0], g1), scale(tpar[:, 1], g2)))
        an = cross(at, a3)
        return a3, an

    dxA = [vadd(dXA[0], duA[0]), vadd(dXA[1], duA[1])]
    dxB = [vadd(dXB[0], duB[0]), vadd(dXB[1], duB[1])]
    A3A, AnA = frame(dXA, True)
    A3B, _ = frame(dXB, False)
    a3A, anA = frame(dxA, True)
    a3B, _ = frame(dxB, False)
    du = vsub(uA, uB)
    r1 = dot(a3A, a3B) - dot(A3A, A3B)
    r2 = dot(anA, a3B) - dot(AnA, A3B)
    e = w * (0.5 * alpha_d * dot(du, du) + 0.5 * alpha_r * (r1 * r1 + r2 * r2))
    return e
