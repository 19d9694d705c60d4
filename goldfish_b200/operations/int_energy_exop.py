"""Internal-energy objective, same surface as
/root/reference/GOLDFISH/operations/int_energy_exop.py (``IntEnergyExOperation`` :3):
``Wint()`` :55, ``dWintduIGA(array, apply_bcs)`` :61, ``dWintdCPIGA(field, array)`` :75,
``dWintdh_th(extract, array)`` :92.  W = sum_s int psi_SVK dA comes out of the
same quadrature pass as the residual/tangent."""
import torch
from ..vecmat import DeviceVec
from .. import _capi as capi
import ctypes as C


class IntEnergyExOperation(object):
    def __init__(self, nonmatching_opt, wint_regu=None):
        if wint_regu is not None and any(w is not None for w in wint_regu):
            raise NotImplementedError("wint_regu (user UFL regularisation) is outside the accelerated path")
        self.nonmatching_opt = nonmatching_opt
        self.num_splines = nonmatching_opt.num_splines
        self.splines = nonmatching_opt.splines
        self.opt_shape = nonmatching_opt.opt_shape
        self.opt_thickness = nonmatching_opt.opt_thickness
        if self.opt_shape:
            self.opt_field = nonmatching_opt.opt_field
            self.shopt_surf_inds = nonmatching_opt.shopt_surf_inds

    def _fresh(self, **what):
        dm = self.nonmatching_opt.dm
        dm.ensure(**what)
        return dm

    def Wint(self):
        dm = self._fresh(functionals=True)
        return float(dm.wv_sum[0].item())

    def dWintduIGA(self, array=True, apply_bcs=True):
        nm = self.nonmatching_opt
        dm = self._fresh(thickness=True)
        g = dm.dWdu.clone()
        if apply_bcs:
            capi.check(dm.lib.gf_mask_vec(C.byref(dm.model), C.c_void_p(g.data_ptr()), dm._stream()), "gf_mask_vec")
        v = DeviceVec(g, nm.vec_iga_dof_list, dm)
        return v.array if array else v

    def dWintdCPIGA(self, field, array=True):
        nm = self.nonmatching_opt
        dm = self._fresh(shape=True)
        fi = self.opt_field.index(field)
        v = DeviceVec(dm.dWdP[fi].clone(), [nm.vec_scalar_iga_dof_list[s] for s in self.shopt_surf_inds[fi]], dm)
        return v.array if array else v

    def dWintdh_th(self, extract=False, array=True):
        nm = self.nonmatching_opt
        dm = self._fresh(thickness=True)
        v = DeviceVec(dm.dWdt[:dm.sym.n_th].clone(), nm.h_th_sizes, dm)
        return v.array if array else v
