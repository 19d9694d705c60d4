"""CPU: the oracle is anchored on physics known-answers and invariants, because
the reference's own tests hold no assertions (SURVEY.md sections 4, 8c)."""
import numpy as np
import pytest
from oracle.model import OracleModel
from oracle import bspline as obs
import cases


@pytest.fixture(scope="module")
def tb():
    pr, kw = cases.tbeam_small()
    m = OracleModel(pr)
    m.set_u(cases.random_state(m.N, m.bc_global))
    return m


def test_scordelis_lo_known_answer():
    """QoI_ref = 0.3006 (/root/reference/GOLDFISH/tests/test_slr.py:50)."""
    pr, _ = cases.slr_small()
    pr = __import__("goldfish_b200.problems", fromlist=["x"]).scordelis_lo(num_el=6)
    m = OracleModel(pr)
    u = m.solve_linear()
    q = pr["qoi"]; P = m.patches[q["patch"]]
    conn, D = obs.surface_basis(P.ku, P.kv, 3, 3, P.cp[:, 3], np.array([q["xi"]]))
    uy = (D[0, 0] * u[P.off + q["field"] * P.ncp + conn[0]]).sum()
    assert abs(-uy - q["ref"]) < 2e-3 * q["ref"]


def _flat_patch_problem(a, b, n0, n1, bc, E, nu, t, body_force=(0.0, 0.0, 0.0), edge_loads=()):
    from goldfish_b200.problems import _ruled_quad, _patch_from_surface
    srf = _ruled_quad([[0, 0, 0], [a, 0, 0], [0, b, 0], [a, b, 0]], n0, n1, 3)
    P = _patch_from_surface(srf, 9, dict(kind="const", values=t), bc, body_force)
    return dict(name="flat", patches=[P], E=E, nu=nu, interfaces=[], penalty_coefficient=1e3, point_loads=[],
                edge_loads=list(edge_loads))


def _probe(m, u, xi, field):
    P = m.patches[0]
    conn, D = obs.surface_basis(P.ku, P.kv, 3, 3, P.cp[:, 3], np.array([xi]))
    return float((D[0, 0] * u[P.off + field * P.ncp + conn[0]]).sum())


def test_navier_plate_known_answer():
    """Simply supported square plate under uniform load, nu = 0.3: w_max = 0.00406235 q a^4 / D with
    D = E t^3 / (12 (1 - nu^2)) (Timoshenko & Woinowsky-Krieger, table 8).  Anchors the bending stiffness
    and its nu dependence, which Scordelis-Lo and the T-beam (both nu = 0) do not see."""
    E, nu, t, a, q = 1.0e7, 0.3, 0.01, 1.0, 1.0e-3
    bc = [(f, d, s, 1) for f in range(3) for d in (0, 1) for s in (0, 1)]
    m = OracleModel(_flat_patch_problem(a, a, 8, 8, bc, E, nu, t, body_force=(0.0, 0.0, q)))
    w = _probe(m, m.solve_linear(), (0.5, 0.5), 2)
    ref = 0.00406235 * q * a ** 4 / (E * t ** 3 / (12 * (1 - nu ** 2)))
    assert abs(w - ref) < 2e-4 * ref


def test_uniaxial_tension_poisson_contraction():
    """Strip under an edge traction N (force per unit length): u_x = N x / (E t), u_y = -nu N y / (E t)
    exactly (the state is in the spline space): membrane stiffness and Poisson coupling."""
    E, nu, t, L, W, N = 1.0e7, 0.3, 0.01, 2.0, 1.0, 5.0
    bc = [(0, 0, 0, 1), (1, 1, 0, 1), (2, 0, 0, 2)]
    pr = _flat_patch_problem(L, W, 4, 3, bc, E, nu, t,
                             edge_loads=[dict(patch=0, direction=0, side=1, traction=(N, 0.0, 0.0))])
    m = OracleModel(pr)
    u = m.solve_linear()
    assert abs(_probe(m, u, (1.0, 1.0), 0) - N * L / (E * t)) < 1e-10 * N * L / (E * t)
    assert abs(_probe(m, u, (1.0, 1.0), 1) + nu * N * W / (E * t)) < 1e-10 * N * W / (E * t)
    assert abs(_probe(m, u, (0.5, 0.5), 0) - 0.5 * N * L / (E * t)) < 1e-10 * N * L / (E * t)


def test_tangent_symmetric_and_fd(tb):
    m = tb
    K = m.stiffness(apply_bcs=False)
    assert abs(K - K.T).max() < 1e-12 * abs(K).max()
    u = m.u.copy(); v = np.random.default_rng(3).standard_normal(m.N); h = 1e-6
    m.set_u(u + h * v); Rp = m.residual(False); m.set_u(u - h * v); Rm = m.residual(False); m.set_u(u)
    assert np.linalg.norm((Rp - Rm) / (2 * h) - K @ v) < 1e-8 * np.linalg.norm(K @ v)


def test_rigid_body_modes_in_kernel():
    """Shell + penalty tangent at u = 0 without BCs annihilates rigid motions."""
    pr, _ = cases.tbeam_small()
    m = OracleModel(pr)
    K0 = m.stiffness(apply_bcs=False)
    for t, w in (((1, 0, 0), (0, 0, 0)), ((0, 0, 1), (0, 0, 0)), ((0, 0, 0), (1, 0, 0)), ((0, 0, 0), (0, 1, 1))):
        r = np.zeros(m.N)
        for P in m.patches:
            X = P.cp[:, :3] / P.cp[:, 3:4]
            d = np.array(t, float)[None] + np.cross(np.array(w, float)[None], X)
            r[P.off:P.off + 3 * P.ncp] = (d * P.cp[:, 3:4]).T.ravel()
        assert np.linalg.norm(K0 @ r) < 1e-12 * abs(K0).max() * np.linalg.norm(r)


def test_adjoint_partials_fd(tb):
    """Mirror of dRIGAdCPIGA_FD (/root/reference/GOLDFISH/nonmatching_opt.py:975-990)."""
    m = tb
    rng = np.random.default_rng(4)
    for field in (0, 2):
        A = m.dRdCP(field, apply_bcs=False)
        cp0 = m.get_cp(field); dv = rng.standard_normal(cp0.size); h = 1e-6
        m.set_cp(field, cp0 + h * dv); Rp = m.residual(False); m.set_cp(field, cp0 - h * dv); Rm = m.residual(False)
        m.set_cp(field, cp0)
        assert np.linalg.norm((Rp - Rm) / (2 * h) - A @ dv) < 1e-7 * np.linalg.norm(A @ dv)
    A = m.dRdt(); th0 = m.theta.copy(); dv = rng.standard_normal(th0.size); h = 1e-7
    m.set_thickness(th0 + h * dv); Rp = m.residual(False); m.set_thickness(th0 - h * dv); Rm = m.residual(False)
    m.set_thickness(th0)
    assert np.linalg.norm((Rp - Rm) / (2 * h) - A @ dv) < 1e-6 * np.linalg.norm(A @ dv)


def test_thickness_homogeneity(tb):
    """Membrane energy ~ t, bending ~ t^3: dW/dt . t = W_m + 3 W_b."""
    m = tb
    from oracle.kl_shell import shell_energy_density
    Wm = Wb = 0.0
    for P in m.patches:
        for sel in m._chunks(P):
            conn = P.conn[sel]; D = P.D[sel]; ne, nq = D.shape[:2]
            Xc = P.cp[:, :3][conn]; uc = m.u[P.off:P.off + 3 * P.ncp].reshape(3, P.ncp).T[conn]
            GX = np.einsum("eqka,eac->eqkc", D[:, :, 1:6], Xc).reshape(ne * nq, 15)
            Gu = np.einsum("eqka,eac->eqkc", D[:, :, 1:6], uc).reshape(ne * nq, 15)
            th = m.theta[P.toff:P.toff + P.nth]
            tq = (P.tw[sel] * th[P.tconn[sel]][:, None, :]).sum(-1).reshape(-1)
            gX = [[GX[:, 3 * k + c] for c in range(3)] for k in range(5)]
            gu = [[Gu[:, 3 * k + c] for c in range(3)] for k in range(5)]
            _, _, wm, wb = shell_energy_density(gX, gu, tq, P.E, P.nu)
            Wm += (wm.reshape(ne, nq) * P.wq[sel]).sum(); Wb += (wb.reshape(ne, nq) * P.wq[sel]).sum()
    assert abs(m.dWdt() @ m.theta - (Wm + 3 * Wb)) < 1e-10 * abs(Wm + 3 * Wb)
    assert abs(m.energy() - (Wm + Wb)) < 1e-12 * abs(Wm + Wb)


def test_oracle_regression_goldens():
    """The oracle reproduces its committed C2 outputs (drift guard)."""
    import os
    g = np.load(os.path.join(cases.GOLDEN, "tbeam_c2_golden.npz"))
    pr, kw = cases.tbeam_c2()
    m = OracleModel(pr)
    m.set_u(g["u"])
    assert np.abs(m.residual() - g["R"]).max() <= 1e-12 * np.abs(g["R"]).max()
    K = m.stiffness()
    assert np.array_equal(K.indptr, g["K_indptr"]) and np.array_equal(K.indices, g["K_indices"])
    assert np.abs(K.data - g["K_data"]).max() <= 1e-12 * np.abs(g["K_data"]).max()


def test_oracle_nonlinear_cantilever_known_answer():
    """The INDEPENDENT numpy oracle (second-order jets of one energy expression) on the geometrically nonlinear cantilever
    of Sze, Liu & Lo (2004), P = 1 of P_max = 4: tip deflections (0.563, 3.015).  Pins the oracle's own tangent and
    large-rotation kinematics against published values, not only against the compiled port."""
    from goldfish_b200 import problems
    from oracle import bspline as obs
    om = OracleModel(problems.cantilever_shear(1.0, ne=4))
    u = om.solve_nonlinear(max_it=30, rtol=1e-7)
    assert om.newton_history[-1] < 1e-7 and len(om.newton_history) < 20
    P = om.patches[1]
    conn, D = obs.surface_basis(P.ku, P.kv, 3, 3, P.cp[:, 3], np.array([[0.5, 1.0]]))
    d = np.array([(D[0, 0] * u[P.off + f * P.ncp + conn[0]]).sum() for f in range(3)])
    assert abs(-d[0] / 0.563 - 1.0) < 6e-3 and abs(d[2] / 3.015 - 1.0) < 3e-3
