"""CPU: building blocks of the oracle and of the host side."""
from math import factorial
import numpy as np
from oracle import bspline as obs, quadrature as oq
from oracle.jet import Jet
from oracle.kl_shell import shell_energy_density
from goldfish_b200 import bsplines as bsp, quadrature as pq, problems


def test_basis_two_algorithms_agree():
    ku = np.array([0, 0, 0, 0, .2, .5, .5, .7, 1, 1, 1, 1.])
    x = np.concatenate([np.random.default_rng(0).random(64), [0.0, 1.0, 0.5, 0.2]])
    s1, d1 = obs.basis_ders(ku, 3, x, 2)
    s2, d2 = bsp.basis_window(ku, 3, x, 2)
    assert np.array_equal(s1, s2)
    assert np.abs(d1 - d2).max() < 1e-11
    assert np.abs(d1[:, 0].sum(1) - 1).max() < 1e-14      # partition of unity


def test_triangle_rules_exact():
    for deg, npts in ((6, 12), (9, 25), (12, 49)):
        pts, w = oq.triangle_rule(deg)
        assert len(w) == npts
        for a in range(deg + 1):
            for b in range(deg + 1 - a):
                exact = factorial(a) * factorial(b) / factorial(a + b + 2)
                assert abs((w * pts[:, 0] ** a * pts[:, 1] ** b).sum() - exact) < 2e-15


def test_span_rule_tables_identical_oracle_vs_product():
    for deg in (6, 9, 12):
        p1, w1, tri = oq.element_rule(deg)
        p2, w2, tw = pq.span_rule(deg)
        assert np.abs(p1 - p2).max() < 1e-15 and np.abs(w1 - w2).max() < 1e-16
        assert np.abs(tw.sum(1) - 1).max() < 1e-15
        # barycentric interpolation reproduces linear functions
        f = lambda x, y: 2 + 3 * x - 5 * y
        vals = np.array([f(0, 0), f(1, 0), f(0, 1), f(1, 1)])
        assert np.abs(tw @ vals - f(p2[:, 0], p2[:, 1])).max() < 1e-14


def test_jet_against_finite_differences():
    rng = np.random.default_rng(1)
    n = 4
    GX = rng.standard_normal((n, 15)); Gu = 0.1 * rng.standard_normal((n, 15)); t = 0.1 + 0.05 * rng.random(n)

    def ev(GX, Gu, t):
        gX = [[GX[:, 3 * k + c] for c in range(3)] for k in range(5)]
        gu = [[Gu[:, 3 * k + c] for c in range(3)] for k in range(5)]
        return shell_energy_density(gX, gu, t, 1e3, 0.3)[0]
    V = Jet.variables(np.concatenate([GX, Gu, t[:, None]], 1))
    e = shell_energy_density([V[3 * k:3 * k + 3] for k in range(5)], [V[15 + 3 * k:18 + 3 * k] for k in range(5)], V[30], 1e3, 0.3)[0]
    assert np.allclose(e.v, ev(GX, Gu, t), rtol=1e-14)
    h = 1e-6
    for k in range(31):
        d = np.zeros((n, 31)); d[:, k] = h
        fd = (ev(GX + d[:, :15], Gu + d[:, 15:30], t + d[:, 30]) - ev(GX - d[:, :15], Gu - d[:, 15:30], t - d[:, 30])) / (2 * h)
        assert np.allclose(fd, e.g[:, k], rtol=2e-6, atol=1e-7 * np.abs(e.g).max())
    assert np.abs(e.h - e.h.transpose(0, 2, 1)).max() < 1e-10 * np.abs(e.h).max()


def test_nurbs_geometry_exact():
    pr = problems.scordelis_lo(num_el=4)
    for P in pr["patches"][:3]:
        cp = P["cp"]
        xi = np.random.default_rng(2).random((30, 2))
        conn, D = bsp.surface_point_tables(P["knots"][0], P["knots"][1], 3, 3, cp[:, 3], xi)
        X = np.einsum("qa,qac->qc", D[:, 0], cp[:, :3][conn])
        assert np.abs(np.hypot(X[:, 0], X[:, 1]) - 25.0).max() < 1e-12
    assert problems.num_dofs(problems.tbeam()) == 648
    assert problems.num_dofs(problems.plate(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "plate_c1_input.npz"))) == 1449
