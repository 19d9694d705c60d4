"""Host-side spline utilities of the product (numpy, setup time only).

Plays the role tIGAr's ``BSpline`` / ``NURBSControlMesh`` play for the reference
(/root/reference/GOLDFISH/tests/test_tbeam.py:18-32): knot spans, basis functions
with first and second derivatives (Cox-de Boor recursion + the derivative
recurrence), and the igakit-style geometry operators the fixtures need
(``ruled``/``elevate``/``refine``, /root/reference/GOLDFISH/tests/test_tbeam.py:5-16).

Scalar control-point ordering is tIGAr's: a = i + j*n_u (u fastest,
/root/reference/GOLDFISH/utils/bsp_utils.py:14).
"""
import numpy as np


def num_basis(knots, p):
    return len(knots) - p - 1


def span_index(knots, p, x):
    """k such that knots[k] <= x < knots[k+1]; x == knots[-1] belongs to the
    last non-empty span."""
    knots = np.asarray(knots, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    k = np.searchsorted(knots, x, side="right") - 1
    return np.clip(k, p, num_basis(knots, p) - 1).astype(np.int64)


def nonempty_spans(knots, p):
    knots = np.asarray(knots, dtype=np.float64)
    ks = np.arange(p, num_basis(knots, p))
    return ks[knots[ks + 1] > knots[ks]]


def _safe_div(a, b):
    out = np.zeros_like(a)
    np.divide(a, b, out=out, where=(b != 0.0))
    return out


def basis_window(knots, p, x, nder=2):
    """Values and derivatives (order 0..nder) of the p+1 B-splines that are
    non-zero on the span of each x.  Returns (k, B) with B[q, r, j] the r-th
    derivative of N_{k-p+j,p} at x[q].

    Cox-de Boor on the local window of p+2 candidate functions per degree;
    derivatives by N^(r)_{i,d} = d (N^(r-1)_{i,d-1}/(t_{i+d}-t_i)
                                   - N^(r-1)_{i+1,d-1}/(t_{i+d+1}-t_{i+1})).
    """
    knots = np.asarray(knots, dtype=np.float64)
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    k = span_index(knots, p, x)
    nq = len(x)
    # pad the knot vector so every window index is valid
    kn = np.concatenate([np.full(p + 1, knots[0]), knots, np.full(p + 1, knots[-1])])
    off = p + 1
    i0 = k - p  # first window function index
    # tab[d][r] : (nq, p+2-? ) values of r-th derivative of N_{i0+j, d}, j = 0..p (+1 slack)
    width = p + 2
    idx = i0[:, None] + np.arange(width)[None, :]          # function indices i
    tab = {}
    N0 = (idx == k[:, None]).astype(np.float64)             # degree 0
    tab[(0, 0)] = N0
    for d in range(1, p + 1):
        ti = kn[idx + off]; tid = kn[idx + d + off]
        ti1 = kn[idx + 1 + off]; tid1 = kn[idx + d + 1 + off]
        for r in range(0, min(d, nder) + 1):
            if r == 0:
                lo = tab[(d - 1, 0)]
                hi = np.concatenate([lo[:, 1:], np.zeros((nq, 1))], axis=1)
                val = _safe_div((x[:, None] - ti) * lo, tid - ti) + \
                    _safe_div((tid1 - x[:, None]) * hi, tid1 - ti1)
            else:
                lo = tab.get((d - 1, r - 1))
                hi = np.concatenate([lo[:, 1:], np.zeros((nq, 1))], axis=1)
                val = d * (_safe_div(lo, tid - ti) - _safe_div(hi, tid1 - ti1))
            tab[(d, r)] = val
    B = np.zeros((nq, nder + 1, p + 1))
    for r in range(0, min(p, nder) + 1):
        B[:, r, :] = tab[(p, r)][:, :p + 1]
    return k, B


def surface_point_tables(ku, kv, pu, pv, weights, xi):
    """Rational basis phi_a = N_a/W at points xi (nq,2): returns
    (conn[nq, nloc], D[nq, 6, nloc]) with kinds
    [phi, phi_u, phi_v, phi_uu, phi_vv, phi_uv] and local node lu + lv*(pu+1).
    Geometry and displacement of the reference are homogeneous fields divided
    by the weight function (``spline.rationalize``), hence N_a/W for both."""
    xi = np.atleast_2d(np.asarray(xi, dtype=np.float64))
    nu = num_basis(ku, pu)
    su, Bu = basis_window(ku, pu, xi[:, 0], 2)
    sv, Bv = basis_window(kv, pv, xi[:, 1], 2)
    iu = (su - pu)[:, None] + np.arange(pu + 1)[None, :]
    iv = (sv - pv)[:, None] + np.arange(pv + 1)[None, :]
    conn = (iu[:, None, :] + nu * iv[:, :, None]).reshape(len(xi), -1)

    def tp(a, b):
        return (Bv[:, b, :, None] * Bu[:, a, None, :]).reshape(len(xi), -1)

    N, Nu, Nv, Nuu, Nvv, Nuv = tp(0, 0), tp(1, 0), tp(0, 1), tp(2, 0), tp(0, 2), tp(1, 1)
    weights = np.asarray(weights, dtype=np.float64)
    if np.all(weights == 1.0):
        return conn, np.stack([N, Nu, Nv, Nuu, Nvv, Nuv], axis=1)
    wl = weights[conn]
    s = lambda A: (A * wl).sum(1)[:, None]
    W, Wu, Wv, Wuu, Wvv, Wuv = s(N), s(Nu), s(Nv), s(Nuu), s(Nvv), s(Nuv)
    f = N / W
    fu = (Nu - f * Wu) / W
    fv = (Nv - f * Wv) / W
    fuu = (Nuu - 2 * fu * Wu - f * Wuu) / W
    fvv = (Nvv - 2 * fv * Wv - f * Wvv) / W
    fuv = (Nuv - fu * Wv - fv * Wu - f * Wuv) / W
    return conn, np.stack([f, fu, fv, fuu, fvv, fuv], axis=1)


# ---------------------------------------------------------------------------
# geometry operators (igakit ``refine`` / ``elevate`` equivalents)
# ---------------------------------------------------------------------------
def knot_insertion_operator(knots, p, new_knots):
    """T with Q = T @ P (Boehm, one knot at a time). Returns (T, new_knots)."""
    knots = np.asarray(knots, dtype=np.float64)
    T = np.eye(num_basis(knots, p))
    for u in np.sort(np.asarray(new_knots, dtype=np.float64)):
        n = num_basis(knots, p)
        k = int(span_index(knots, p, u))
        A = np.zeros((n + 1, n))
        for i in range(n + 1):
            if i <= k - p:
                A[i, i] = 1.0
            elif i > k:
                A[i, i - 1] = 1.0
            else:
                al = (u - knots[i]) / (knots[i + p] - knots[i])
                A[i, i] = al
                A[i, i - 1] = 1.0 - al
        T = A @ T
        knots = np.insert(knots, k + 1, u)
    return T, knots


def degree_elevation_operator(knots, p, t=1):
    """E with Q = E @ P for p -> p+t: collocate the old basis at the Greville
    abscissae of the elevated knot vector (exact: nested spline spaces)."""
    knots = np.asarray(knots, dtype=np.float64)
    if t == 0:
        return np.eye(num_basis(knots, p)), knots
    uniq, mult = np.unique(knots, return_counts=True)
    new = np.repeat(uniq, mult + t)
    q = p + t
    n_new, n_old = num_basis(new, q), num_basis(knots, p)
    grev = np.array([new[i + 1:i + q + 1].mean() for i in range(n_new)])
    kn, Bn = basis_window(new, q, grev, 0)
    ko, Bo = basis_window(knots, p, grev, 0)
    Cn = np.zeros((n_new, n_new)); Co = np.zeros((n_new, n_old))
    for r in range(n_new):
        Cn[r, kn[r] - q:kn[r] + 1] = Bn[r, 0]
        Co[r, ko[r] - p:ko[r] + 1] = Bo[r, 0]
    return np.linalg.solve(Cn, Co), new


class NURBSSurface:
    """Minimal igakit-like NURBS surface: homogeneous control net
    ``control[i, j, 0:4] = (w x, w y, w z, w)``."""

    def __init__(self, knots, degree, control):
        self.knots = [np.asarray(k, dtype=np.float64) for k in knots]
        self.degree = list(degree)
        self.control = np.asarray(control, dtype=np.float64)

    def elevate(self, axis, t):
        if t <= 0:
            return self
        E, new = degree_elevation_operator(self.knots[axis], self.degree[axis], t)
        self.control = np.moveaxis(np.tensordot(E, np.moveaxis(self.control, axis, 0), axes=1), 0, axis)
        self.knots[axis] = new
        self.degree[axis] += t
        return self

    def refine(self, axis, new_knots):
        if len(new_knots) == 0:
            return self
        T, new = knot_insertion_operator(self.knots[axis], self.degree[axis], new_knots)
        self.control = np.moveaxis(np.tensordot(T, np.moveaxis(self.control, axis, 0), axes=1), 0, axis)
        self.knots[axis] = new
        return self

    def flat_control(self):
        """(n_u*n_v, 4) in tIGAr order (u fastest)."""
        return self.control.transpose(1, 0, 2).reshape(-1, 4).copy()


def line(p0, p1):
    c = np.zeros((2, 4)); c[0, :3] = p0; c[1, :3] = p1; c[:, 3] = 1.0
    return [0.0, 0.0, 1.0, 1.0], 1, c


def circle_arc(center, radius, angle):
    """Quadratic rational arc (single segment, sweep < 180 deg), z = center[2]."""
    a0, a1 = angle
    am = 0.5 * (a0 + a1)
    wm = np.cos(0.5 * (a1 - a0))
    c = np.zeros((3, 4))
    pts = [(np.cos(a0), np.sin(a0), 1.0), (np.cos(am) / wm, np.sin(am) / wm, wm),
           (np.cos(a1), np.sin(a1), 1.0)]
    for r, (cx, cy, w) in enumerate(pts):
        c[r, 0] = (center[0] + radius * cx) * w
        c[r, 1] = (center[1] + radius * cy) * w
        c[r, 2] = center[2] * w
        c[r, 3] = w
    return [0.0, 0.0, 0.0, 1.0, 1.0, 1.0], 2, c


def ruled(c0, c1):
    """Linear interpolation between two compatible curves (same knots/degree)."""
    k0, p0, P0 = c0
    k1, p1, P1 = c1
    assert p0 == p1 and np.allclose(k0, k1)
    ctrl = np.stack([P0, P1], axis=1)  # [i, j, 4]
    return NURBSSurface([k0, [0.0, 0.0, 1.0, 1.0]], [p0, 1], ctrl)
