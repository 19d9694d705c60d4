"""GPU: the reference-facing facade (NonMatchingOpt + operations) -- same call
sequences as /root/reference/GOLDFISH/operations/disp_imop.py:145-153 and the
OpenMDAO components' check_partials, but asserted."""
import numpy as np
import pytest

import cases
from oracle.model import OracleModel

pytestmark = pytest.mark.gpu


def build_nm(thickness_kind="iga", shape=True):
    from goldfish_b200 import bsplines as bsp, problems
    from goldfish_b200.nonmatching_opt import (NonMatchingOptFFD, SplinePatch, Thickness, ShellLoad, SVK_residual, PointSource)
    pr = problems.tbeam(num_el=4, body_force=(0.0, 0.0, 1.0), thickness_kind=thickness_kind)
    splines = []
    for P in pr["patches"]:
        sp = SplinePatch(P["knots"], P["p"], P["cp"], P["quad_deg"])
        for field in range(3):
            sp.addZeroDofs(field, sp.getSideDofs(1, 0, nLayers=1))
        splines.append(sp)
    h_th = [Thickness(thickness_kind, 0.1) for _ in splines]
    nm = NonMatchingOptFFD(splines, 1.0e7, h_th, 0.0)
    if shape:
        nm.set_shopt_surf_inds([0, 2], [[0, 1], [0, 1]])
    nm.set_thickness_opt(var_thickness=(thickness_kind == "iga"))
    nm.create_mortar_meshes([10])
    nm.mortar_meshes_setup([[0, 1]], [[np.array([[0.5, 0.], [0.5, 1.]]), np.array([[0., 0.], [0., 1.]])]], 1.0e3, 2)
    nm.set_residuals([SVK_residual(dWext=ShellLoad(body_force=(0., 0., 1.))) for _ in splines])
    nm.set_point_sources([PointSource(2, (1., 1.), 10.0)], [0])
    return nm, pr


def test_facade_matches_oracle_and_semantics(built_lib):
    from goldfish_b200.operations import DispImOpeartion, IntEnergyExOperation, VolumeExOperation
    nm, pr = build_nm()
    om = OracleModel(pr)
    disp = DispImOpeartion(nm)
    u = disp.solve_nonlinear(max_it=30, rtol=1e-3)
    uo = om.solve_nonlinear(max_it=30, rtol=1e-3)
    assert np.linalg.norm(u - uo) < 1e-8 * np.linalg.norm(uo)
    nm.update_uIGA(u)
    om.set_u(u)                      # compare every operator at the SAME state
    res = disp.apply_nonlinear()
    # at the converged state R is a small difference of large shell and penalty forces
    # (entries of K ~ 1e9 times u): the meaningful scale is |K| |u|
    scale = (abs(om.stiffness(apply_bcs=False)) @ np.abs(u)).max()
    assert np.abs(res - om.residual()).max() < 1e-11 * scale
    disp.linearize()
    wint, vol = IntEnergyExOperation(nm), VolumeExOperation(nm)
    assert abs(wint.Wint() - om.energy()) < 1e-10 * om.energy()
    assert abs(vol.volume() - om.volume()) < 1e-12 * om.volume()
    assert np.abs(wint.dWintduIGA() - om.dWdu(apply_bcs=True)).max() < 1e-10 * np.abs(om.dWdu()).max()
    assert np.abs(wint.dWintdCPIGA(2) - om.dWdCP(2)).max() < 1e-10 * np.abs(om.dWdCP(2)).max()
    assert np.abs(vol.dvoldCPIGA(0) - om.dVdCP(0)).max() < 1e-10 * max(np.abs(om.dVdCP(0)).max(), 1e-30)
    assert np.abs(wint.dWintdh_th() - om.dWdt()).max() < 1e-10 * np.abs(om.dWdt()).max()
    assert np.abs(vol.dvoldh_th() - om.dVdt()).max() < 1e-12 * np.abs(om.dVdt()).max()
    # apply_linear: accumulate (+=) semantics and fwd/rev transposition
    rng = np.random.default_rng(5)
    N, ncp, nth = nm.vec_iga_dof, nm.vec_scalar_iga_dof, nm.h_th_dof
    du = rng.standard_normal(N); dcp = [rng.standard_normal(ncp) for _ in nm.opt_field]; dt = rng.standard_normal(nth)
    dres0 = rng.standard_normal(N)
    dres = dres0.copy()
    disp.apply_linear_fwd(dcp + [dt], du, dres)
    K = om.stiffness(); Ps = [om.dRdCP(f) for f in nm.opt_field]; T = om.dRdt()
    ref = dres0 + K @ du + sum(P @ d for P, d in zip(Ps, dcp)) + T @ dt
    assert np.abs(dres - ref).max() < 1e-10 * np.abs(ref).max()
    lam = rng.standard_normal(N)
    d_in = [np.ones(ncp) for _ in nm.opt_field] + [np.ones(nth)]; d_out = np.ones(N)
    disp.apply_linear_rev(d_in, d_out, lam)
    assert np.abs(d_out - (1 + K.T @ lam)).max() < 1e-10 * np.abs(K.T @ lam).max()
    for k, P in enumerate(Ps):
        assert np.abs(d_in[k] - (1 + P.T @ lam)).max() < 1e-10 * np.abs(P.T @ lam).max()
    assert np.abs(d_in[-1] - (1 + T.T @ lam)).max() < 1e-10 * np.abs(T.T @ lam).max()
    # solve_linear: overwrite ([:]=) semantics, forward and reverse
    b = rng.standard_normal(N); b[om.bc_global] = 0
    x = np.full(N, 7.0)
    disp.solve_linear_fwd(x, b)
    xo = om.solve(K, b)
    assert np.linalg.norm(x - xo) < 1e-8 * np.linalg.norm(xo)
    y = np.full(N, -3.0)
    disp.solve_linear_rev(b, y)
    assert np.linalg.norm(y - om.solve(K, b, transpose=True)) < 1e-8 * np.linalg.norm(xo)


def test_total_gradient_against_finite_differences(built_lib):
    """check_totals mirror (demos_csdl_alpha/thickness_opt/plate_const_th_opt_wint.py:220-223):
    d W_int(u(t), t) / d t_patch by the adjoint vs central differences of the GPU analysis."""
    from goldfish_b200.operations import DispImOpeartion, IntEnergyExOperation
    nm, pr = build_nm(thickness_kind="const", shape=False)
    disp, wint = DispImOpeartion(nm), IntEnergyExOperation(nm)

    def analysis(th):
        nm.update_h_th(th)
        u = disp.solve_nonlinear(max_it=30, rtol=1e-10)
        nm.update_uIGA(u)
        return wint.Wint()

    th0 = nm.init_h_th.copy()
    analysis(th0)
    disp.linearize()
    dWdu = wint.dWintduIGA(); dWdt = wint.dWintdh_th()
    lam = np.zeros(nm.vec_iga_dof)
    disp.solve_linear_rev(dWdu, lam)
    d_in = [np.zeros(nm.h_th_dof)]
    disp.apply_linear_rev(d_in, None, lam)
    total = dWdt - d_in[0]
    for k in range(len(th0)):
        h = 1e-6 * th0[k]
        e = np.zeros_like(th0); e[k] = h
        fd = (analysis(th0 + e) - analysis(th0 - e)) / (2 * h)
        assert abs(fd - total[k]) < 2e-6 * abs(total[k])


def test_error_behaviour(built_lib):
    from goldfish_b200.opt_utils import update_nest_vec
    nm, pr = build_nm()
    with pytest.raises(ValueError):
        nm.mortar_meshes_setup([[0, 1]], [[np.zeros((2, 2)), np.zeros((2, 2))]], 1e3, 1, penalty_method="maximum")
    with pytest.raises(ValueError):
        nm.dRIGAdCPIGA(1)
    with pytest.raises(TypeError):
        update_nest_vec(np.zeros(3), object())
    with pytest.raises(ValueError):
        nm.update_uIGA(np.zeros(5))
