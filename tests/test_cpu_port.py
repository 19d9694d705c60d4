"""CPU: the compiled C++/OpenMP port (oracle/c/kl_cpu.cpp: shells + penalty coupling; oracle/c/mf_lu.cpp:
multifrontal LU) agrees with the numpy oracle (second-order jets of one energy expression).  The port shares the
point-level header with the CUDA kernels (dual numbers over a hand-derived first variation), so the numpy oracle is
the independent derivation and this file is what ties the port -- bench.py's CPU arm -- to it."""
import numpy as np
import scipy.sparse as sp
import pytest
import cases
from oracle.model import OracleModel
from oracle.cpu_port import CpuModel
from goldfish_b200 import _capi as capi


@pytest.mark.parametrize("case", ["tbeam_small", "slr_small", "wingbox_small"])
def test_port_matches_numpy_oracle(case):
    pr, kw = getattr(cases, case)()
    cm = CpuModel(pr, **kw)
    om = OracleModel(pr)
    u = cases.random_state(om.N, om.bc_global)
    cm.set_u(u); om.set_u(u)
    cm.assemble(capi.GF_OUT_R | capi.GF_OUT_K | capi.GF_OUT_W | capi.GF_OUT_P | capi.GF_OUT_T)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(cm.residual(), om.residual()) < 1e-11                # shells + penalty + BCs, all from the port
    K = cm.K_matrix(); Ko = om.stiffness()
    assert np.array_equal(K.indptr, Ko.indptr) and np.array_equal(K.indices, Ko.indices)
    assert abs(K - Ko).max() < 1e-11 * abs(Ko).max()
    T = cm.T_matrix()
    assert abs(T - om.dRdt()).max() < 1e-11 * abs(om.dRdt()).max()
    assert abs(cm.WV[0::2].sum() - om.energy()) < 1e-11 * om.energy()
    assert rel(cm.dWdt, om.dWdt()) < 1e-11 and rel(cm.dVdt, om.dVdt()) < 1e-11
    surf = kw["shopt_surf_inds"]
    same = all(list(x) == list(surf[0]) for x in surf)
    Aos = om.dRdCP_fields(kw["opt_field"], surf[0]) if same else [om.dRdCP(f, surf[i]) for i, f in enumerate(kw["opt_field"])]
    for i, f in enumerate(kw["opt_field"]):
        Ao = Aos[i]                                                 # shell + penalty parts (one jet pass for all fields)
        assert abs(cm.P_matrix(i) - Ao).max() < 1e-11 * abs(Ao).max()
        assert rel(cm.dWdP[i], om.dWdCP(f, surf[i])) < 1e-11


def test_port_iteration_matches_numpy_oracle():
    """One full analysis + adjoint iteration (Newton with multifrontal LU, re-factorised transpose for the
    adjoint, total gradients) against the numpy oracle with SuperLU."""
    pr, kw = cases.tbeam_small()
    cm = CpuModel(pr, **kw)
    om = OracleModel(pr)
    _, grads = cm.iteration()
    uo = om.solve_nonlinear(max_it=30, rtol=1e-3)
    assert cm.newton_its == len(om.newton_history) - 1
    assert np.linalg.norm(cm.u - uo) < 1e-9 * np.linalg.norm(uo)
    lam = om.solve(om.stiffness(), om.dWdu(apply_bcs=True), transpose=True)
    assert np.linalg.norm(cm.lam - lam) < 1e-8 * np.linalg.norm(lam)
    go = om.dWdt() - om.dRdt().T @ lam
    assert np.linalg.norm(grads[-1] - go) < 1e-8 * np.linalg.norm(go)
    for i, f in enumerate(kw["opt_field"]):
        gp = om.dWdCP(f, kw["shopt_surf_inds"][i]) - om.dRdCP(f, kw["shopt_surf_inds"][i]).T @ lam
        assert np.linalg.norm(grads[i] - gp) < 1e-8 * np.linalg.norm(gp)


def test_wingbox_total_gradient_against_finite_differences():
    """The reference's own check (check_totals) on the wing-box topology, through the CPU port: adjoint total
    derivative of W_int w.r.t. one patch thickness against a central difference of converged Newton solves."""
    pr, _ = cases.wingbox_small()
    cm = CpuModel(pr, opt_field=[0, 1, 2], shopt_surf_inds=[list(range(len(pr["patches"])))] * 3)
    th0 = cm.theta.copy()
    _, g = cm.iteration(newton_rtol=1e-10)
    ip, h = 3, 1e-4 * th0[3]

    def W_at(d):
        cm.theta[:] = th0; cm.theta[ip] += d
        cm.iteration(newton_rtol=1e-10)
        return cm.W
    fd = (W_at(h) - W_at(-h)) / (2 * h)
    assert abs(fd - g[-1][ip]) < 1e-6 * abs(fd)


def test_multifrontal_lu_matches_superlu():
    """oracle/c/mf_lu.cpp on the tangent pattern of a non-matching 8-patch cylinder with random UNSYMMETRIC values
    (the LU, not a Cholesky, is what the reference's MUMPS run does) against scipy's SuperLU."""
    import scipy.sparse.linalg as spla
    from goldfish_b200 import problems
    cm = CpuModel(problems.cylinder(n_el=10))
    S = cm.S
    lu = cm.direct_solver()
    rng = np.random.default_rng(0)
    A = sp.csr_matrix((rng.standard_normal(len(S.K_indices)), S.K_indices, S.K_indptr), shape=(S.N, S.N)) + sp.diags(np.full(S.N, 300.0))
    A = A.tocsr(); A.sort_indices()
    AT = A.T.tocsr(); AT.sort_indices()
    b = rng.standard_normal(S.N)
    x = lu.factor(A, AT).solve(b)
    assert np.linalg.norm(A @ x - b) < 1e-12 * np.linalg.norm(b)
    xs = spla.splu(A.tocsc()).solve(b)
    assert np.linalg.norm(x - xs) < 1e-11 * np.linalg.norm(xs)
    assert lu.nfronts > 10 and lu.flops > 0


# ---- known answers on closed, multi-patch, non-matching NURBS cylinders (through the compiled port) ----
# Point sources follow the reference (PointSource on the homogeneous FE space,
# /root/reference/GOLDFISH/nonmatching_opt.py:735-738): value * N_a(xi), not value * N_a / W.  Inside a
# rational patch the physical force is therefore value * W(xi); W(0.5) = (1 + cos 45deg) / 2 on a 90-degree arc.
W_MID = 0.5 * (1.0 + np.sqrt(0.5))


def _free_cylinder(n_el, R, L, t, E, nu, jitter, n_axial):
    """Closed cylinder without supports other than six statically determinate dofs (self-equilibrated loads)."""
    from goldfish_b200 import problems
    pr = problems.cylinder(n_el=n_el, R=R, L=L, E=E, nu=nu, h_th=t, pressure_like_load=(0., 0., 0.),
                           jitter=jitter, n_axial=n_axial)
    for P in pr["patches"]:
        P["bc_dofs"] = np.zeros(0, dtype=np.int64)
    P0 = pr["patches"][0]
    n_u, n_v = len(P0["knots"][0]) - 4, len(P0["knots"][1]) - 4
    ncp, a0, a1, a2 = n_u * n_v, 0, n_u - 1, (n_v - 1) * n_u
    P0["bc_dofs"] = np.array([a0, ncp + a0, 2 * ncp + a0, ncp + a1, 2 * ncp + a1, a2], dtype=np.int64)
    return pr


def _solve_linear_port(pr):
    import scipy.sparse.linalg as spla
    cm = CpuModel(pr)
    cm.set_u(np.zeros(cm.S.N))
    cm.assemble(capi.GF_OUT_R | capi.GF_OUT_K)
    return cm, cm.solve(-cm.residual())


def _radial_approach(cm, u, probes):
    from oracle import bspline as obs
    tot = 0.0
    om = OracleModel(cm.problem)
    for s, xi, er in probes:
        P = om.patches[s]
        conn, D = obs.surface_basis(P.ku, P.kv, 3, 3, P.cp[:, 3], np.array([xi]))
        tot -= sum(er[f] * (D[0, 0] * u[P.off + f * P.ncp + conn[0]]).sum() for f in range(2))
    return tot


def test_pinched_ring_known_answer():
    """Ring of four 90-degree NURBS patches (closed by a penalty interface) under two opposite radial loads at
    patch junctions: decrease of the loaded diameter = (pi/4 - 2/pi) P R^3 / (E I) (Timoshenko, curved bars)."""
    R, b, t, E, F = 1.0, 0.1, 0.01, 1.0e7, 1.0e-3
    pr = _free_cylinder(12, R, b, t, E, 0.0, False, 1)
    pr["point_loads"] = [dict(patch=0, field=1, xi=(1.0, 0.5), value=+F), dict(patch=2, field=1, xi=(1.0, 0.5), value=-F)]
    cm, u = _solve_linear_port(pr)
    d = _radial_approach(cm, u, [(0, (1.0, 0.5), (0.0, 1.0)), (2, (1.0, 0.5), (0.0, -1.0))])
    ref = (np.pi / 4 - 2 / np.pi) * F * R ** 3 / (E * b * t ** 3 / 12)
    assert abs(d / ref - 1.0) < 2e-3


def test_pinched_free_cylinder_known_answer():
    """Pinched cylinder with free ends (R = 4.953, L = 10.35, t = 0.094, E = 10.5e6, nu = 0.3125, P = 100):
    radial displacement under the loads 0.1139 (shell obstacle course; inextensional theory 0.1084).  Eight
    NON-MATCHING patches, loads in the middle of two arcs on the interface between the axial halves: curved
    geometry, nu != 0, penalty coupling and rational point sources in one known answer."""
    R, L, t, E, nu, F = 4.953, 10.35, 0.094, 10.5e6, 0.3125, 100.0
    pr = _free_cylinder(12, R, L, t, E, nu, True, 2)
    c = np.sqrt(0.5) * F / W_MID
    pr["point_loads"] = [dict(patch=0, field=0, xi=(0.5, 1.0), value=+c), dict(patch=0, field=1, xi=(0.5, 1.0), value=+c),
                         dict(patch=2, field=0, xi=(0.5, 1.0), value=-c), dict(patch=2, field=1, xi=(0.5, 1.0), value=-c)]
    cm, u = _solve_linear_port(pr)
    s = np.sqrt(0.5)
    w = 0.5 * _radial_approach(cm, u, [(0, (0.5, 1.0), (s, s)), (2, (0.5, 1.0), (-s, -s))])
    assert abs(w / 0.1139 - 1.0) < 1.5e-2


def test_twisted_beam_known_answer():
    """MacNeal-Harder twisted beam (L = 12, w = 1.1, t = 0.32, E = 29e6, nu = 0.22, 90-degree twist, unit tip loads) on two
    NON-MATCHING patches: tip displacement in the load direction 5.424e-3 (in-plane) / 1.754e-3 (out-of-plane); a
    Kirchhoff-Love shell has no transverse shear and converges to 0.995 of both.  Doubly curved geometry, nu != 0, penalty
    coupling; Maxwell-Betti reciprocity of the two load cases is checked too."""
    from goldfish_b200 import problems, bsplines as bsp
    tips = {}
    for name, F in (("in", (0.0, 1.0, 0.0)), ("out", (1.0, 0.0, 0.0))):
        pr = problems.twisted_beam(16, F)
        cm = CpuModel(pr)
        cm.set_u(np.zeros(cm.S.N)); cm.assemble(capi.GF_OUT_R | capi.GF_OUT_K)
        u = cm.solve(-cm.residual())
        P = cm.S.patches[1]
        conn, D = bsp.surface_point_tables(P.ku, P.kv, 3, 3, np.ones(P.ncp), np.array([[0.5, 1.0]]))
        tips[name] = np.array([(D[0, 0] * u[P.dof_off + f * P.ncp + conn[0]]).sum() for f in range(3)])
    assert 0.990 < tips["in"][1] / 5.424e-3 < 1.0
    assert 0.990 < tips["out"][0] / 1.754e-3 < 1.0
    assert abs(tips["in"][0] - tips["out"][1]) < 1e-6 * abs(tips["in"][0])       # u_x(F_y) = u_y(F_x)


def test_pinched_hemisphere_known_answer():
    """Pinched hemisphere with an 18-degree hole (R = 10, t = 0.04, E = 6.825e7, nu = 0.3, F = +-2): radial displacement
    under the loads 0.0940 (shell obstacle course).  Four NON-MATCHING exact NURBS patches closed into a ring: nearly
    inextensional bending of a doubly curved RATIONAL surface -- the case that exposes membrane locking and any error in
    the rational basis derivatives or the curvature terms."""
    from goldfish_b200 import problems, bsplines as bsp
    pr = problems.hemisphere(16)
    cm = CpuModel(pr)
    cm.set_u(np.zeros(cm.S.N)); cm.assemble(capi.GF_OUT_R | capi.GF_OUT_K)
    u = cm.solve(-cm.residual())
    rad = []
    for k in range(4):
        P = cm.S.patches[k]
        conn, D = bsp.surface_point_tables(P.ku, P.kv, 3, 3, P.cp[:, 3], np.array([[0.0, 1.0]]))
        d = np.array([(D[0, 0] * u[P.dof_off + f * P.ncp + conn[0]]).sum() for f in range(3)])
        rad.append(d[0] * np.cos(k * np.pi / 2) + d[1] * np.sin(k * np.pi / 2))
    rad = np.array(rad)
    assert np.all(rad[[0, 2]] > 0) and np.all(rad[[1, 3]] < 0)                 # outward under +F, inward under -F
    assert 0.985 < np.mean(np.abs(rad)) / 0.0940 < 1.0
    assert np.ptp(np.abs(rad)) < 2e-3 * np.mean(np.abs(rad))                   # the four non-matching patches agree


@pytest.mark.parametrize("P,U_ref,W_ref", [(1.0, 0.563, 3.015), (4.0, 3.286, 6.698)])
def test_nonlinear_cantilever_known_answer(P, U_ref, W_ref):
    """The geometrically NONLINEAR benchmark of Sze, Liu & Lo (2004): cantilever under end shear, tip deflection up to
    67 % of the length.  Newton from u = 0 with the full load (as the reference does, no load stepping) on two
    non-matching patches: checks the full St.Venant-Kirchhoff tangent -- geometric stiffness and the second derivatives
    of the curvature -- and the large-rotation kinematics against published values (all other known answers are linear)."""
    from goldfish_b200 import problems, bsplines as bsp
    cm = CpuModel(problems.cantilever_shear(P))
    cm.set_u(np.zeros(cm.S.N))
    ref, hist = None, []
    for it in range(40):
        cm.assemble(capi.GF_OUT_R | capi.GF_OUT_K)
        nrm = np.linalg.norm(cm.R); ref = nrm if it == 0 else ref
        hist.append(nrm / ref)
        if it > 0 and hist[-1] < 1e-7:
            break
        cm.set_u(cm.u + cm.solve(-cm.R))
    assert hist[-1] < 1e-7 and len(hist) < 25 and max(hist) > 1e3        # far from equilibrium on the way, quadratic at the end
    Pp = cm.S.patches[1]
    conn, D = bsp.surface_point_tables(Pp.ku, Pp.kv, 3, 3, np.ones(Pp.ncp), np.array([[0.5, 1.0]]))
    d = np.array([(D[0, 0] * cm.u[Pp.dof_off + f * Pp.ncp + conn[0]]).sum() for f in range(3)])
    assert abs(-d[0] / U_ref - 1.0) < 5e-3 and abs(d[2] / W_ref - 1.0) < 2e-3
