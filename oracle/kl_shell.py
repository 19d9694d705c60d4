"""ORACLE (test infrastructure) -- Kirchhoff-Love St.Venant-Kirchhoff shell energy.

Restates ShNAPr ``surfaceEnergyDensitySVK(spline, X, x, E, nu, h_th)``
(david-kamensky/ShNAPr, un-vendored and unpinned; reference call site
/root/reference/GOLDFISH/operations/int_energy_exop.py:27-30 and, through
PENGoLINS ``SVK_residual``, /root/reference/GOLDFISH/tests/test_tbeam.py:107-110).
Published algorithm (Kiendl et al. 2009; Kamensky 2021), SURVEY.md Appendix A.3:

    A_a = X_{,a},  A_3 = A_1 x A_2 / |.|,  A_ab = A_a.A_b,  B_ab = -A_a.A_{3,b}
    eps_ab = (a_ab - A_ab)/2,   kap_ab = B_ab - b_ab       (lower case: x = X + u)
    eps_bar, kap_bar = components in the local orthonormal basis
        E_1 = A_1/|A_1|, E_2 = (A_2 - (A_2.E_1)E_1)/|.|   via contravariant A^a
    n_bar = t D voigt(eps_bar),  m_bar = t^3/12 D voigt(kap_bar)
    psi = (voigt(eps_bar).n_bar + voigt(kap_bar).m_bar)/2
    dx  = sqrt(det A_ab) dxi           (tIGAr ``spline.dx``)

Everything is written on scalars that may be floats, numpy arrays or Jets, so
first and second derivatives come from AD exactly as UFL ``derivative`` does
for the reference.  Parity unpinned (ShNAPr not installed); validated by
physics known-answers in tests/ (Scordelis-Lo 0.3006, rigid-body invariance).
"""
import numpy as np
from .jet import Jet, dot, cross, norm, unit, scale, vsub, vadd


def _sqrt(x):
    return x.sqrt() if isinstance(x, Jet) else np.sqrt(x)


def surface_geometry(g1, g2, g11, g22, g12):
    """ShNAPr ``surfaceGeometry``: (a0,a1,a2, metric a, curvature b)."""
    n = cross(g1, g2)
    nn = dot(n, n)
    inv_j = 1.0 / _sqrt(nn)
    a2 = scale(inv_j, n)

    def d_a2(da0, da1):
        # derivative of the unit normal: (I - a2 a2^T) n_{,b} / |n|
        dn = vadd(cross(da0, g2), cross(g1, da1))
        proj = dot(a2, dn)
        return scale(inv_j, vsub(dn, scale(proj, a2)))

    a2_1 = d_a2(g11, g12)
    a2_2 = d_a2(g12, g22)
    a = [[dot(g1, g1), dot(g1, g2)], [dot(g2, g1), dot(g2, g2)]]
    b = [[-dot(g1, a2_1), -dot(g1, a2_2)], [-dot(g2, a2_1), -dot(g2, a2_2)]]
    return a2, a, b


def cartesian_transform(a, a0, a1):
    """Q[i][k] = E_i . A^k of ShNAPr ``covariantRank2TensorToCartesian2D``
    (local orthonormal basis E_i, contravariant basis A^k)."""
    det = a[0][0] * a[1][1] - a[0][1] * a[1][0]
    idet = 1.0 / det
    ac = [[a[1][1] * idet, -a[0][1] * idet], [-a[1][0] * idet, a[0][0] * idet]]
    a0c = vadd(scale(ac[0][0], a0), scale(ac[0][1], a1))
    a1c = vadd(scale(ac[1][0], a0), scale(ac[1][1], a1))
    e0 = unit(a0)
    e1 = unit(vsub(a1, scale(dot(a1, e0), e0)))
    return [[dot(e0, a0c), dot(e0, a1c)], [dot(e1, a0c), dot(e1, a1c)]]


def to_cartesian(T, Q):
    """M_ij = sum_kl T_kl (E_i.A^k)(A^l.E_j) = (Q T Q^T)_ij."""
    QT = [[Q[i][0] * T[0][l] + Q[i][1] * T[1][l] for l in range(2)] for i in range(2)]
    return [[QT[i][0] * Q[j][0] + QT[i][1] * Q[j][1] for j in range(2)] for i in range(2)]


def voigt(T):
    return [T[0][0], T[1][1], 2.0 * T[0][1]]


def shell_energy_density(GX, Gu, t, E, nu):
    """Returns (e, J, psi_m*J, psi_b*J) with e = psi * sqrt(det A) per unit
    parametric area.  GX, Gu: lists of five 3-vectors
    [.,1  .,2  .,11  .,22  .,12]."""
    X1, X2, X11, X22, X12 = GX
    x = [vadd(a, b) for a, b in zip(GX, Gu)]
    x1, x2, x11, x22, x12 = x
    A2, A, B = surface_geometry(X1, X2, X11, X22, X12)
    a2, a, b = surface_geometry(x1, x2, x11, x22, x12)
    eps = [[0.5 * (a[i][j] - A[i][j]) for j in range(2)] for i in range(2)]
    kap = [[B[i][j] - b[i][j] for j in range(2)] for i in range(2)]
    Q = cartesian_transform(A, X1, X2)
    eb = voigt(to_cartesian(eps, Q))
    kb = voigt(to_cartesian(kap, Q))
    c = E / (1.0 - nu * nu)
    D = [[c, c * nu, 0.0], [c * nu, c, 0.0], [0.0, 0.0, c * 0.5 * (1.0 - nu)]]

    def quad(v):
        s = 0.0
        for i in range(3):
            for j in range(3):
                if D[i][j] != 0.0:
                    s = s + v[i] * v[j] * D[i][j]
        return s

    J = _sqrt(A[0][0] * A[1][1] - A[0][1] * A[1][0])
    wm = 0.5 * t * quad(eb) * J
    wb = 0.5 * (t * t * t) * quad(kb) * (1.0 / 12.0) * J
    return wm + wb, J, wm, wb
