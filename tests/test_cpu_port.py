"""CPU: the compiled C++/OpenMP port (oracle/c/kl_cpu.cpp) agrees with the numpy
oracle -- two independent CPU restatements (jets vs. dual numbers over the
hand-derived first variation) of the same shell quadrature."""
import numpy as np
import scipy.sparse as sp
import pytest
import cases
from oracle.model import OracleModel
from oracle.cpu_port import CpuModel
from goldfish_b200 import _capi as capi


@pytest.mark.parametrize("case", ["tbeam_small", "slr_small"])
def test_port_matches_numpy_oracle(case):
    pr, kw = getattr(cases, case)()
    cm = CpuModel(pr, **kw)
    om = OracleModel(pr)
    u = cases.random_state(om.N, om.bc_global)
    cm.set_u(u); om.set_u(u)
    cm.shell(capi.GF_OUT_R | capi.GF_OUT_K | capi.GF_OUT_W | capi.GF_OUT_P | capi.GF_OUT_T)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(cm.residual(), om.residual()) < 1e-11
    K = cm.K_matrix(); Ko = om.stiffness()
    assert abs(K - Ko).max() < 1e-11 * abs(Ko).max()
    T = sp.csr_matrix((cm.Tv, cm._idx["T"][1], cm._idx["T"][0]), shape=(om.N, om.n_th))
    assert abs(T - om.dRdt()).max() < 1e-11 * abs(om.dRdt()).max()
    assert abs(cm.WV[0::2].sum() - om.energy()) < 1e-11 * om.energy()
    assert rel(cm.dWdt, om.dWdt()) < 1e-11 and rel(cm.dVdt, om.dVdt()) < 1e-11
    for i, f in enumerate(kw["opt_field"]):
        Psh = sp.csr_matrix((cm.Pv[i], cm._idx["P"][i][1], cm._idx["P"][i][0]), shape=(om.N, cm.S.P_ncols[i]))
        Ao = om.dRdCP_fields([f], kw["shopt_surf_inds"][i], penalty=False)[0]
        assert abs(Psh - Ao).max() < 1e-11 * abs(Ao).max()
        assert rel(cm.dWdP[i], om.dWdCP(f, kw["shopt_surf_inds"][i])) < 1e-11


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
    cm.shell(capi.GF_OUT_R | capi.GF_OUT_K)
    return cm, spla.splu(cm.K_matrix().tocsc()).solve(-cm.residual())


def _radial_approach(cm, u, probes):
    from oracle import bspline as obs
    tot = 0.0
    for s, xi, er in probes:
        P = cm.om.patches[s]
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
