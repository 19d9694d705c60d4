"""Volume constraint V = sum_s int t dA, same surface as
/root/reference/GOLDFISH/operations/volume_exop.py (``VolumeExOperation`` :3):
``volume()`` :46, ``dvoldh_th(array)`` :52, ``dvoldCPIGA(field, array)`` :69."""
from ..vecmat import DeviceVec


class VolumeExOperation(object):
    def __init__(self, nonmatching_opt, vol_surf_inds=None):
        if vol_surf_inds is not None and sorted(vol_surf_inds) != list(range(nonmatching_opt.num_splines)):
            raise NotImplementedError("vol_surf_inds subsets are not accelerated yet")
        self.nonmatching_opt = nonmatching_opt
        self.num_splines = nonmatching_opt.num_splines
        self.opt_shape = nonmatching_opt.opt_shape
        self.opt_thickness = nonmatching_opt.opt_thickness
        if self.opt_shape:
            self.opt_field = nonmatching_opt.opt_field
            self.shopt_surf_inds = nonmatching_opt.shopt_surf_inds

    def _fresh(self, **what):
        dm = self.nonmatching_opt.dm
        dm.ensure(**what)
        return dm

    def volume(self):
        dm = self._fresh(functionals=True)
        return float(dm.wv_sum[1].item())

    def dvoldh_th(self, array=True):
        nm = self.nonmatching_opt
        dm = self._fresh(thickness=True)
        v = DeviceVec(dm.dVdt[:dm.sym.n_th].clone(), nm.h_th_sizes, dm)
        return v.array if array else v

    def dvoldCPIGA(self, field, array=True):
        nm = self.nonmatching_opt
        dm = self._fresh(shape=True)
        fi = self.opt_field.index(field)
        v = DeviceVec(dm.dVdP[fi].clone(), [nm.vec_scalar_iga_dof_list[s] for s in self.shopt_surf_inds[fi]], dm)
        return v.array if array else v
