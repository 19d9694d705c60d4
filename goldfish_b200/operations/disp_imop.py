"""Implicit displacement operation -- numpy-in / numpy-out facade.

Same public surface and in-place semantics as
/root/reference/GOLDFISH/operations/disp_imop.py (class ``DispImOpeartion`` :3,
the spelling is the reference's): ``apply_linear_*`` ACCUMULATE (+=) into the
caller's arrays (:71,:80,:111,:119,:123) and ``solve_linear_*`` OVERWRITE
([:]=, :133,:140).  The thickness derivative is the entry after the shape
fields in ``d_inputs_array_list`` (:82-85).
"""
from ..opt_utils import (get_petsc_vec_array, update_nest_vec, A_x_b, AT_x_b, solve_Ax_b, solve_ATx_b)


class DispImOpeartion(object):
    def __init__(self, nonmatching_opt):
        self.nonmatching_opt = nm = nonmatching_opt
        self.comm = nm.comm
        self.opt_shape, self.opt_field = nm.opt_shape, nm.opt_field
        self.opt_thickness, self.var_thickness = nm.opt_thickness, nm.var_thickness
        self.use_aero_pressure = nm.use_aero_pressure
        nm.dm  # build the device model
        self.dres_iga = nm.vec_iga_nest.copy()
        self.du_iga = nm.vec_iga_nest.copy()
        if self.opt_shape:
            self.dcp_iga = [v.copy() for v in nm.cpdes_iga_nest]
        if self.opt_thickness:
            self.dh_th = nm.h_th_nest.copy()

    def apply_nonlinear(self):
        return get_petsc_vec_array(self.nonmatching_opt.RIGA(), self.comm)

    def solve_nonlinear(self, max_it=30, rtol=1e-3):
        _, u_iga = self.nonmatching_opt.solve_nonlinear_nonmatching_problem(
            max_it=max_it, zero_mortar_funcs=True, rtol=rtol, iga_dofs=True)
        return get_petsc_vec_array(u_iga, self.comm)

    def linearize(self):
        nm = self.nonmatching_opt
        self.dRdu_iga = nm.dRIGAduIGA()
        if self.opt_shape:
            self.dRigadcpiga_list = [nm.dRIGAdCPIGA(field) for field in self.opt_field]
        if self.opt_thickness:
            self.dRigadh_th = nm.dRIGAdh_th()

    def _input_ops(self):
        """(operator, work vector) of every design input, in the order of
        ``d_inputs_array_list``: shape fields first, thickness last."""
        ops = []
        if self.opt_shape:
            ops += list(zip(self.dRigadcpiga_list, self.dcp_iga))
        if self.opt_thickness:
            ops.append((self.dRigadh_th, self.dh_th))
        return ops

    def apply_linear_fwd(self, d_inputs_array_list=None, d_outputs_array=None, d_residuals_array=None):
        if d_residuals_array is None:
            return d_residuals_array
        if d_outputs_array is not None:
            update_nest_vec(d_outputs_array, self.du_iga)
            A_x_b(self.dRdu_iga, self.du_iga, self.dres_iga)
            d_residuals_array[:] += get_petsc_vec_array(self.dres_iga, self.comm)
        if d_inputs_array_list is not None:
            for k, (op, work) in enumerate(self._input_ops()):
                update_nest_vec(d_inputs_array_list[k], work)
                A_x_b(op, work, self.dres_iga)
                d_residuals_array[:] += get_petsc_vec_array(self.dres_iga, self.comm)
        return d_residuals_array

    def apply_linear_rev(self, d_inputs_array_list=None, d_outputs_array=None, d_residuals_array=None):
        if d_residuals_array is not None:
            update_nest_vec(d_residuals_array, self.dres_iga)
            if d_outputs_array is not None:
                AT_x_b(self.dRdu_iga, self.dres_iga, self.du_iga)
                d_outputs_array[:] += get_petsc_vec_array(self.du_iga, self.comm)
            if d_inputs_array_list is not None:
                for k, (op, work) in enumerate(self._input_ops()):
                    AT_x_b(op, self.dres_iga, work)
                    d_inputs_array_list[k][:] += get_petsc_vec_array(work, self.comm)
        return d_inputs_array_list, d_outputs_array

    def solve_linear_fwd(self, d_outputs_array, d_residuals_array):
        K = self.dRdu_iga.copy()
        update_nest_vec(d_residuals_array, self.dres_iga)
        d_outputs_array[:] = solve_Ax_b(K, self.dres_iga, array=True, comm=self.comm)
        return d_outputs_array

    def solve_linear_rev(self, d_outputs_array, d_residuals_array):
        K = self.dRdu_iga.copy()
        update_nest_vec(d_outputs_array, self.du_iga)
        d_residuals_array[:] = solve_ATx_b(K, self.du_iga, array=True, comm=self.comm)
        return d_residuals_array
