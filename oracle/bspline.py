"""ORACLE (test infrastructure, CPU only) -- B-spline / NURBS basis utilities.

Restates, in numpy, the spline evaluation the reference obtains from tIGAr
(`tIGAr.BSplines.BSpline`, un-vendored; call sites
/root/reference/GOLDFISH/utils/opt_utils.py:2-4 and
/root/reference/GOLDFISH/tests/test_tbeam.py:19-32).  Parity unpinned: tIGAr
is not installed in this container, see DESIGN.md section "Oracle".

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may
import this package.  The product (goldfish_b200) never does.

Conventions (/root/reference/GOLDFISH/utils/bsp_utils.py:14
``ij2dof(l,i,j)=i+j*l``): scalar control-point index a = i + j*n_u (u fastest).
"""
import numpy as np


def find_span(knots, p, x):
    """Knot-span index k with knots[k] <= x < knots[k+1] (right-continuous,
    clamped so the last non-empty span owns x == knots[-1])."""
    knots = np.asarray(knots, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n = len(knots) - p - 1  # number of basis functions
    k = np.searchsorted(knots, x, side="right") - 1
    k = np.clip(k, p, n - 1)
    # move left while span is empty (x == end knot)
    return k.astype(np.int64)


def basis_ders(knots, p, x, nder=2):
    """All p+1 non-zero B-spline basis functions and derivatives up to `nder`
    at points x.  Algorithm A2.3 of Piegl & Tiller, vectorised over points.

    Returns (span, ders) with ders[q, d, j] = d-th derivative of N_{span-p+j}.
    """
    knots = np.asarray(knots, dtype=np.float64)
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    nq = x.shape[0]
    span = find_span(knots, p, x)
    ndu = np.zeros((nq, p + 1, p + 1))
    left = np.zeros((nq, p + 1))
    right = np.zeros((nq, p + 1))
    ndu[:, 0, 0] = 1.0
    for j in range(1, p + 1):
        left[:, j] = x - knots[span + 1 - j]
        right[:, j] = knots[span + j] - x
        saved = np.zeros(nq)
        for r in range(j):
            ndu[:, j, r] = right[:, r + 1] + left[:, j - r]
            temp = ndu[:, r, j - 1] / ndu[:, j, r]
            ndu[:, r, j] = saved + right[:, r + 1] * temp
            saved = left[:, j - r] * temp
        ndu[:, j, j] = saved
    ders = np.zeros((nq, nder + 1, p + 1))
    ders[:, 0, :] = ndu[:, :, p]
    for r in range(p + 1):
        s1, s2 = 0, 1
        a = np.zeros((nq, 2, p + 1))
        a[:, 0, 0] = 1.0
        for k in range(1, min(nder, p) + 1):
            d = np.zeros(nq)
            rk = r - k
            pk = p - k
            if r >= k:
                a[:, s2, 0] = a[:, s1, 0] / ndu[:, pk + 1, rk]
                d = a[:, s2, 0] * ndu[:, rk, pk]
            j1 = 1 if rk >= -1 else -rk
            j2 = k - 1 if (r - 1) <= pk else p - r
            for j in range(j1, j2 + 1):
                a[:, s2, j] = (a[:, s1, j] - a[:, s1, j - 1]) / ndu[:, pk + 1, rk + j]
                d = d + a[:, s2, j] * ndu[:, rk + j, pk]
            if r <= pk:
                a[:, s2, k] = -a[:, s1, k - 1] / ndu[:, pk + 1, r]
                d = d + a[:, s2, k] * ndu[:, r, pk]
            ders[:, k, r] = d
            s1, s2 = s2, s1
    fac = p
    for k in range(1, min(nder, p) + 1):
        ders[:, k, :] *= fac
        fac *= (p - k)
    return span, ders


def unique_spans(knots, p):
    """Indices k of the non-empty knot spans [knots[k], knots[k+1])."""
    knots = np.asarray(knots)
    n = len(knots) - p - 1
    ks = np.arange(p, n)
    return ks[knots[ks + 1] > knots[ks]]


def surface_basis(knots_u, knots_v, p_u, p_v, w, xi):
    """Rational bivariate basis phi_a = N_a / W and its parametric derivatives.

    `w` : (n_u*n_v,) weights (u fastest).  `xi`: (nq,2).
    Returns (conn, D) with conn[q, (p_u+1)*(p_v+1)] global scalar CP indices
    (local node l = lu + lv*(p_u+1)) and D[q, 6, nloc] =
    [phi, phi_u, phi_v, phi_uu, phi_vv, phi_uv].

    The reference represents geometry and displacement homogeneously
    (cpFuncs[0..2] = w*P, cpFuncs[3] = w; u = u_hom / w via
    ``spline.rationalize``: /root/reference/GOLDFISH/operations/int_energy_exop.py:24-26),
    so both use the basis N_a / W(xi), W = sum_b N_b w_b.
    """
    xi = np.atleast_2d(xi)
    n_u = len(knots_u) - p_u - 1
    su, du = basis_ders(knots_u, p_u, xi[:, 0], 2)
    sv, dv = basis_ders(knots_v, p_v, xi[:, 1], 2)
    iu = (su - p_u)[:, None] + np.arange(p_u + 1)[None, :]
    iv = (sv - p_v)[:, None] + np.arange(p_v + 1)[None, :]
    conn = (iu[:, None, :] + n_u * iv[:, :, None]).reshape(xi.shape[0], -1)

    def tp(a, b):  # [q, lv, lu] -> flat with lu fastest
        return (dv[:, b, :, None] * du[:, a, None, :]).reshape(xi.shape[0], -1)

    N = tp(0, 0); Nu = tp(1, 0); Nv = tp(0, 1)
    Nuu = tp(2, 0); Nvv = tp(0, 2); Nuv = tp(1, 1)
    wl = np.asarray(w)[conn]
    W = (N * wl).sum(1)[:, None]
    Wu = (Nu * wl).sum(1)[:, None]; Wv = (Nv * wl).sum(1)[:, None]
    Wuu = (Nuu * wl).sum(1)[:, None]; Wvv = (Nvv * wl).sum(1)[:, None]
    Wuv = (Nuv * wl).sum(1)[:, None]
    phi = N / W
    phi_u = (Nu - phi * Wu) / W
    phi_v = (Nv - phi * Wv) / W
    phi_uu = (Nuu - 2.0 * phi_u * Wu - phi * Wuu) / W
    phi_vv = (Nvv - 2.0 * phi_v * Wv - phi * Wvv) / W
    phi_uv = (Nuv - phi_u * Wv - phi_v * Wu - phi * Wuv) / W
    D = np.stack([phi, phi_u, phi_v, phi_uu, phi_vv, phi_uv], axis=1)
    return conn, D
