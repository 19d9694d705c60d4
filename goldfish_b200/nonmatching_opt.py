"""``NonMatchingOpt`` -- the reference's problem object, re-implemented on the
B200 path (drop-in boundary).

Mirrors the public surface of /root/reference/GOLDFISH/nonmatching_opt.py
(class ``NonMatchingOpt(NonMatchingCoupling)`` :7) that the operations, the
OpenMDAO components and the CSDL models consume: same method names, argument
meaning, return conventions (petsc4py-shaped handles) and error behaviour.
What used to be dolfin/tIGAr/PENGoLINS objects becomes plain data:

  ExtractedSpline   -> ``SplinePatch`` (knots, homogeneous control net, zeroDofs, quad_deg)
  UFL residual form -> ``ShellLoad`` (dead body force, dead edge tractions); the
                       SVK Kirchhoff-Love energy itself is built into the kernels
  dolfin PointSource-> ``PointSource(field, xi, value)``
  h_th Function     -> float | ``Thickness(kind, values)``

All arithmetic runs in the CUDA library (goldfish_b200/csrc); nothing here
falls back to the CPU.
"""
import numpy as np
import torch

from . import _capi as capi
from .device_model import DeviceModel
from .symbolic import Symbolic
from .vecmat import DeviceVec, DeviceMat
from .problems import _side_dofs, mortar_coords


class SplinePatch:
    """Stand-in for tIGAr's ``ExtractedSpline`` of one NURBS patch."""
    nsd = 3

    def __init__(self, knots, degree, control, quad_deg, zero_dofs=()):
        self.knots = [np.asarray(k, dtype=np.float64) for k in knots]
        self.p = int(degree[0])
        self.degree = tuple(int(d) for d in degree)
        self.n_u = len(self.knots[0]) - self.degree[0] - 1
        self.n_v = len(self.knots[1]) - self.degree[1] - 1
        self.control = np.array(control, dtype=np.float64).reshape(self.n_u * self.n_v, 4)
        self.quad_deg = int(quad_deg)
        self.zeroDofs = np.unique(np.asarray(zero_dofs, dtype=np.int64))

    @classmethod
    def from_surface(cls, srf, quad_deg, zero_dofs=()):
        return cls(srf.knots, srf.degree, srf.flat_control(), quad_deg, zero_dofs)

    @property
    def num_cp(self):
        return self.n_u * self.n_v

    # tIGAr generator vocabulary (tests/test_tbeam.py:22-29)
    def getSideDofs(self, direction, side, nLayers=1):
        return _side_dofs(self.n_u, self.n_v, direction, side, nLayers)

    def addZeroDofs(self, field, scalar_dofs):
        new = field * self.num_cp + np.asarray(scalar_dofs, dtype=np.int64)
        self.zeroDofs = np.unique(np.concatenate([self.zeroDofs, new]))


class Thickness:
    def __init__(self, kind="const", values=0.0):
        if kind not in ("const", "linear", "iga"):
            raise ValueError("Undefined thickness kind: {}".format(kind))
        self.kind, self.values = kind, values


class ShellLoad:
    """What the UFL ``source_terms`` of the reference scripts describe
    (tests/test_tbeam.py:98-110): dead loads of the SVK residual."""

    def __init__(self, body_force=(0.0, 0.0, 0.0), edge_tractions=()):
        self.body_force = tuple(float(x) for x in body_force)
        self.edge_tractions = list(edge_tractions)   # (direction, side, (tx,ty,tz))


def SVK_residual(spline=None, u=None, z=None, E=None, nu=None, h_th=None, dWext=None):
    """Name kept from PENGoLINS; only the external work term carries information
    here (the internal SVK energy is fixed in the kernels)."""
    return dWext if isinstance(dWext, ShellLoad) else ShellLoad()


class PointSource:
    def __init__(self, field, xi, value):
        self.field, self.xi, self.value = int(field), (float(xi[0]), float(xi[1])), float(value)


class NonMatchingOpt:
    def __init__(self, splines, E, h_th, nu, int_V_family='CG', int_V_degree=1,
                 int_dx_metadata=None, contact=None, comm=None, device=None):
        if contact is not None:
            raise NotImplementedError("contact is outside the accelerated hot path")
        self.splines = list(splines)
        self.num_splines = len(self.splines)
        self.nsd = 3
        self.npd = 2
        self.comm = comm
        self.device = device
        as_list = lambda v: list(v) if isinstance(v, (list, tuple)) else [v] * self.num_splines
        self.E, self.nu = [float(x) for x in as_list(E)], [float(x) for x in as_list(nu)]
        self.h_th = [t if isinstance(t, Thickness) else Thickness("const", float(t)) for t in as_list(h_th)]
        self.opt_shape = False
        self.opt_field = []
        self.shopt_surf_inds = []
        self.opt_thickness = False
        self.var_thickness = False
        self.use_aero_pressure = False
        self.contact = None
        self.vec_scalar_iga_dof_list = [s.num_cp for s in self.splines]
        self.vec_iga_dof_list = [3 * s.num_cp for s in self.splines]
        self.vec_iga_dof = sum(self.vec_iga_dof_list)
        self.vec_scalar_iga_dof = sum(self.vec_scalar_iga_dof_list)
        self.mapping_list, self.mortar_nels, self.mortar_parametric_coords = [], [], []
        self.num_intersections = 0
        self.penalty_coefficient = 1000
        self.loads = [ShellLoad() for _ in self.splines]
        self.point_sources, self.point_source_inds = None, None
        self.init_cp_iga = None
        self._dm = None
        self.nonlinear_its = 0

    # ------------------------------------------------------------ configuration
    def set_shopt_surf_inds(self, opt_field, shopt_surf_inds):
        assert len(opt_field) == len(shopt_surf_inds)
        self.opt_shape = True
        self.opt_field = list(opt_field)
        self.shopt_surf_inds = [list(x) for x in shopt_surf_inds]
        self.shopt_num_desvars = [sum(self.vec_scalar_iga_dof_list[s] for s in inds) for inds in self.shopt_surf_inds]
        self.cpdes_iga_dofs_full = [np.arange(n) for n in self.shopt_num_desvars]
        self._dm = None

    def set_thickness_opt(self, var_thickness=False):
        self.opt_thickness = True
        self.var_thickness = var_thickness
        if var_thickness:
            for t in self.h_th:
                if t.kind != "iga":
                    raise ValueError("var_thickness=True needs Thickness('iga', ...) on every patch")
        self._dm = None

    def create_mortar_meshes(self, mortar_nels, mortar_coords=None):
        self.mortar_nels = [int(n) for n in mortar_nels]
        self.num_intersections = len(self.mortar_nels)

    def mortar_meshes_setup(self, mapping_list, mortar_parametric_coords, penalty_coefficient=1000,
                            transfer_mat_deriv=1, penalty_method="minimum"):
        if penalty_method != "minimum":
            raise ValueError("Undefined penalty method: {}".format(penalty_method))
        if len(mapping_list) != self.num_intersections:
            raise ValueError("mapping_list does not match create_mortar_meshes")
        self.mapping_list = [tuple(int(x) for x in m) for m in mapping_list]
        self.mortar_parametric_coords = [
            [mortar_coords(np.asarray(side, dtype=np.float64), self.mortar_nels[i]) for side in sides]
            for i, sides in enumerate(mortar_parametric_coords)]
        self.penalty_coefficient = penalty_coefficient
        self.transfer_mat_deriv = transfer_mat_deriv
        self._dm = None

    def set_residuals(self, residuals, residuals_deriv=None):
        if len(residuals) != self.num_splines:
            raise ValueError("one residual (ShellLoad) per spline is required")
        self.loads = [r if isinstance(r, ShellLoad) else ShellLoad() for r in residuals]
        self._dm = None

    def set_point_sources(self, point_sources=[], point_source_inds=[]):
        self.point_sources, self.point_source_inds = list(point_sources), list(point_source_inds)
        self._dm = None

    # ------------------------------------------------------------------ backend
    def _problem(self):
        patches = []
        for s, sp in enumerate(self.splines):
            th = self.h_th[s]
            patches.append(dict(p=sp.degree, knots=tuple(sp.knots), cp=sp.control, bc_dofs=sp.zeroDofs,
                                quad_deg=sp.quad_deg, thickness=dict(kind=th.kind, values=th.values),
                                body_force=self.loads[s].body_force, E=self.E[s], nu=self.nu[s]))
        interfaces = [dict(patches=self.mapping_list[i], xi=tuple(self.mortar_parametric_coords[i]))
                      for i in range(self.num_intersections)]
        pls = []
        if self.point_sources is not None:
            for ps, ind in zip(self.point_sources, self.point_source_inds):
                pls.append(dict(patch=ind, field=ps.field, xi=ps.xi, value=ps.value))
        els = []
        for s, ld in enumerate(self.loads):
            for (d, side, trac) in ld.edge_tractions:
                els.append(dict(patch=s, direction=d, side=side, traction=tuple(trac)))
        return dict(name="NonMatchingOpt", patches=patches, E=self.E[0], nu=self.nu[0], interfaces=interfaces,
                    penalty_coefficient=self.penalty_coefficient, point_loads=pls, edge_loads=els)

    @property
    def dm(self):
        """The device model is built on first use, after all set_* calls."""
        if self._dm is None:
            self.problem = self._problem()
            # a caller that already ran the symbolic phase for this topology may hand it over (`_symbolic`)
            # ... or the whole device model of the same problem (`_device_model`: bench.py times the device-resident
            # step and the facade step on ONE model instead of building the set-up twice)
            self._dm = getattr(self, "_device_model", None) or DeviceModel(
                self.problem, self.opt_field, self.shopt_surf_inds, device=self.device, symbolic=getattr(self, "_symbolic", None))
            S = self._dm.sym
            dv = self._dm.device
            self.vec_iga_nest = DeviceVec.zeros(self.vec_iga_dof_list, dv, self._dm)
            self.vec_scalar_iga_nest = DeviceVec.zeros(self.vec_scalar_iga_dof_list, dv, self._dm)
            self.u_iga_nest = DeviceVec(self._dm.u, self.vec_iga_dof_list, self._dm)
            self.h_th_sizes = [P.nth for P in S.patches]
            self.h_th_dof = S.n_th
            self.h_th_nest = DeviceVec(self._dm.theta, self.h_th_sizes, self._dm)
            self.init_h_th = S.theta0.copy()
            self.init_h_th_list = [P.theta0.copy() for P in S.patches]
            self.init_h_th_fe = self.init_h_th
            if self.opt_shape:
                self.cpdes_iga_nest = [DeviceVec.zeros([self.vec_scalar_iga_dof_list[s] for s in inds], dv, self._dm)
                                       for inds in self.shopt_surf_inds]
        return self._dm

    # ------------------------------------------------------------- design/state
    def get_init_CPIGA(self):
        """Initial control points per opt field (homogeneous coordinates of the
        IGA dofs; the reference solves a least-squares fit from FE dofs,
        nonmatching_opt.py:217-228 -- here the IGA dofs are the primary data)."""
        if self.init_cp_iga is None:
            self.init_cp_iga = [np.concatenate([self.splines[s].control[:, f] for s in inds])
                                for f, inds in zip(self.opt_field, self.shopt_surf_inds)]
        return self.init_cp_iga

    def update_uIGA(self, u_array_iga):
        u = np.asarray(u_array_iga, dtype=np.float64)
        if u.size != self.vec_iga_dof:
            raise ValueError("displacement array has wrong size")
        self.dm.set_u(u)

    def update_CPIGA(self, cp_array_iga, field):
        fi = self.opt_field.index(field)
        self.dm.set_cp(field, np.asarray(cp_array_iga, dtype=np.float64), self.shopt_surf_inds[fi])

    def update_h_th(self, h_th_array):
        self.dm.set_theta(np.asarray(h_th_array, dtype=np.float64))

    update_h_th_IGA = update_h_th

    # ------------------------------------------------------------------ results
    def RIGA(self):
        dm = self.dm
        dm.ensure(residual=True)
        return DeviceVec(dm.R.clone(), self.vec_iga_dof_list, dm)

    def dRIGAduIGA(self):
        dm = self.dm
        dm.ensure(tangent=True)
        return DeviceMat(dm, [dm.K], is_K=True)

    def dRIGAdCPIGA(self, field):
        dm = self.dm
        if not self.opt_shape or field not in self.opt_field:
            raise ValueError("field {} is not a shape-optimisation field".format(field))
        dm.ensure(shape=True)
        fi = self.opt_field.index(field)
        parts = [dm.P[fi]] + ([dm.penP[fi][0]] if dm.penP[fi] is not None else [])
        return DeviceMat(dm, parts)

    def dRIGAdh_th(self):
        dm = self.dm
        dm.ensure(thickness=True)
        return DeviceMat(dm, [dm.T])

    def extract_nonmatching_vec(self, vec_list, ind_list=None, scalar=False, apply_bcs=False):
        raise NotImplementedError("the FE detour does not exist on this path: vectors are born in IGA dofs")

    def solve_nonlinear_nonmatching_problem(self, solver="direct", ref_error=None, rtol=1e-3, max_it=20,
                                            zero_mortar_funcs=True, iga_dofs=False, **kw):
        dm = self.dm
        try:
            dm.newton(max_it=max_it, rtol=rtol)
        except capi.GoldfishNotConverged as e:
            raise StopIteration(str(e))
        self.nonlinear_its = len(dm.newton_history) - 1
        u = DeviceVec(dm.u, self.vec_iga_dof_list, dm)
        return (None, u) if iga_dofs else None

    def solve_linear_nonmatching_problem(self, solver="direct", iga_dofs=False, **kw):
        dm = self.dm
        dm.u.zero_(); dm.touch()
        dm.assemble(residual=True, tangent=True)
        rhs = torch.empty_like(dm.R)
        dm.axpby(-1.0, dm.R, 0.0, rhs)
        dm.solve(rhs, dm.u, refactor=True); dm.touch()
        u = DeviceVec(dm.u, self.vec_iga_dof_list, dm)
        return (None, u) if iga_dofs else None


class NonMatchingOptFFD(NonMatchingOpt):
    """Name kept so fixture scripts read the same; the FFD bookkeeping of
    /root/reference/GOLDFISH/nonmatching_opt_ffd.py is setup-time host code
    outside the accelerated path."""
