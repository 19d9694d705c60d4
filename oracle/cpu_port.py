"""ORACLE / CPU BASELINE (test + measurement infrastructure) -- compiled CPU path.

Drives oracle/c/kl_cpu.cpp (C++/OpenMP shell quadrature + CSR scatter, HOST
memory) together with the numpy oracle's penalty terms and SuperLU solves: the
"restated reference CPU path" timed by bench.py (`cpu_baseline`, `--impl
reference`) with all host threads, and a second, independent CPU check of the
numpy oracle at sizes the numpy AD cannot reach.

It reuses the product's plain-data model description (goldfish_b200.symbolic /
_capi struct layouts are DATA, the arithmetic is oracle/c + oracle/model.py);
nothing in goldfish_b200 imports this file.
"""
import ctypes as C
import os
import subprocess
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from goldfish_b200 import _capi as capi
from goldfish_b200.symbolic import Symbolic
from .model import OracleModel

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "c", "libgfo_cpu.so")


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "c")])
    return LIB


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None and a.size else C.c_void_p(0)


class CpuModel:
    def __init__(self, problem, opt_field=(), shopt_surf_inds=(), with_penalty_oracle=True):
        if not os.path.exists(LIB):
            build()
        self.lib = C.CDLL(LIB)
        self.lib.gfo_shell_assemble.argtypes = [C.POINTER(capi.GfModel), C.c_int, C.POINTER(capi.GfShellOut)]
        self.S = S = Symbolic(problem, opt_field, shopt_surf_inds)
        self.opt_field, self.surf = list(opt_field), [list(x) for x in shopt_surf_inds]
        self.om = OracleModel(problem) if with_penalty_oracle and S.pen["n_eval"] > 0 else None
        a = self.a = {}
        descs = (capi.GfPatchDesc * len(S.patches))()
        for d, P in zip(descs, S.patches):
            d.n_u, d.n_v, d.neu, d.nev = P.n_u, P.n_v, P.neu, P.nev
            d.cp_off, d.dof_off, d.th_off, d.th_kind, d.nth = P.cp_off, P.dof_off, P.th_off, P.th_kind, P.nth
            d.span_u_off, d.span_v_off, d.cpd_u_off, d.cpd_v_off = P.span_u_off, P.span_v_off, P.cpd_u_off, P.cpd_v_off
            d.rational, d.el_off, d.E, d.nu = P.rational, P.el_off, P.E, P.nu
            for f in range(3):
                d.pcol_off[f] = P.pcol_off[f]; d.f[f] = float(P.f[f])
        self._descs = descs
        for k, v, dt in (("elem_patch", S.elem_patch, np.int32), ("elem_eu", S.elem_eu, np.int32), ("elem_ev", S.elem_ev, np.int32),
                         ("color_elem", S.color_elem, np.int32), ("color_ptr", S.color_ptr, np.int32),
                         ("tab_u", S.tab_u, np.float64), ("tab_v", S.tab_v, np.float64),
                         ("first_cp_u", S.first_cp_u, np.int32), ("first_cp_v", S.first_cp_v, np.int32),
                         ("span_h_u", S.span_h_u, np.float64), ("span_h_v", S.span_h_v, np.float64),
                         ("qw", S.qw, np.float64), ("tw_lin", S.tw_lin, np.float64), ("bc", S.bc_mask, np.uint8),
                         ("bc_list", S.bc_list, np.int32), ("row_nlow", S.row_nlow, np.int32)):
            a[k] = np.ascontiguousarray(v, dtype=dt)
        for k, v in S.dirs.items():
            a[k] = np.ascontiguousarray(v, dtype=np.int32)
        self.cp = S.cp0.copy(); self.u = np.zeros(S.N); self.theta = S.theta0.copy()
        self.Kv = np.zeros(S.K_indptr[-1]); self.Tv = np.zeros(S.T_indptr[-1])
        self.Pv = [np.zeros(ip[-1]) for ip in S.P_indptr]
        self.R = np.zeros(S.N); self.WV = np.zeros(2 * S.num_elements); self.dWdu = np.zeros(S.N)
        self.dWdP = [np.zeros(n) for n in S.P_ncols]; self.dVdP = [np.zeros(n) for n in S.P_ncols]
        self.dWdt = np.zeros(S.n_th); self.dVdt = np.zeros(S.n_th); self.dt_el = np.zeros(2 * S.num_elements)
        self._idx = dict(K=(np.ascontiguousarray(S.K_indptr), np.ascontiguousarray(S.K_indices)),
                         T=(np.ascontiguousarray(S.T_indptr), np.ascontiguousarray(S.T_indices)),
                         P=[(np.ascontiguousarray(i), np.ascontiguousarray(j)) for i, j in zip(S.P_indptr, S.P_indices)])
        m = capi.GfModel()
        m.num_patches, m.num_elements, m.nq, m.num_colors = len(S.patches), S.num_elements, S.nq, S.num_colors
        m.N, m.n_scalar, m.n_th = S.N, S.n_scalar, S.n_th
        m.patches = C.cast(descs, C.c_void_p)
        for k in ("elem_patch", "elem_eu", "elem_ev", "color_elem", "tab_u", "tab_v", "first_cp_u", "first_cp_v", "span_h_u",
                  "span_h_v", "qw", "tw_lin", "bc", "bc_list", "row_nlow", "cp_lo_u", "cp_hi_u", "el_lo_u", "el_hi_u",
                  "cp_lo_v", "cp_hi_v", "el_lo_v", "el_hi_v"):
            setattr(m, k, _p(a[k]))
        m.color_ptr_h = _p(a["color_ptr"])
        m.cp, m.u, m.theta, m.n_bc = _p(self.cp), _p(self.u), _p(self.theta), len(S.bc_list)

        def csr(ip, ix, v, ncols):
            c = capi.GfCsr(); c.nrows, c.ncols, c.nnz = S.N, ncols, len(v)
            c.indptr, c.indices, c.vals = _p(ip), _p(ix), _p(v)
            return c
        m.K = csr(*self._idx["K"], self.Kv, S.N)
        for f in range(3):
            if f in S.opt_field:
                i = S.opt_field.index(f)
                m.P[f] = csr(*self._idx["P"][i], self.Pv[i], S.P_ncols[i])
        m.T = csr(*self._idx["T"], self.Tv, S.n_th)
        self.m = m
        o = capi.GfShellOut()
        o.R, o.WV, o.dWdu, o.dWdt, o.dVdt, o.dt_el = _p(self.R), _p(self.WV), _p(self.dWdu), _p(self.dWdt), _p(self.dVdt), _p(self.dt_el)
        for f in range(3):
            if f in S.opt_field:
                i = S.opt_field.index(f)
                o.dWdP[f] = self.dWdP[i].ctypes.data; o.dVdP[f] = self.dVdP[i].ctypes.data
        self.o = o

    def set_u(self, u):
        self.u[:] = u
        if self.om is not None:
            self.om.set_u(u)

    def shell(self, what):
        """Shell part through the compiled port: fills R / K / P / T / functionals in place."""
        if what & capi.GF_OUT_R: self.R[:] = self.S.f_const
        if what & capi.GF_OUT_K: self.Kv[:] = 0
        if what & capi.GF_OUT_P:
            for v in self.Pv + self.dWdP + self.dVdP: v[:] = 0
        if what & capi.GF_OUT_T:
            self.Tv[:] = 0; self.dWdt[:] = 0; self.dVdt[:] = 0; self.dWdu[:] = 0; self.dt_el[:] = 0
        self.lib.gfo_shell_assemble(C.byref(self.m), what, C.byref(self.o))
        if what & capi.GF_OUT_T:
            for P in self.S.patches:
                if P.th_kind == 0:
                    sl = self.dt_el[2 * P.el_off:2 * (P.el_off + P.nel)]
                    self.dWdt[P.th_off] += sl[0::2].sum(); self.dVdt[P.th_off] += sl[1::2].sum()

    def K_matrix(self):
        S = self.S
        K = sp.csr_matrix((self.Kv.copy(), self._idx["K"][1], self._idx["K"][0]), shape=(S.N, S.N))
        if self.om is not None:
            K = K + self.om.stiffness(apply_bcs=True, shell=False, penalty=True)   # BC rows/cols of the part are zeroed there
        K = K.tolil(); bc = S.bc_list
        K[bc, bc] = 1.0
        return K.tocsr()

    def residual(self):
        R = self.R.copy()
        if self.om is not None:
            R += self.om.residual(apply_bcs=False, shell=False, penalty=True, const_loads=False)
        R[self.S.bc_list] = 0.0
        return R

    def iteration(self):
        """One analysis + adjoint iteration of the restated reference CPU path (LU solves)."""
        S = self.S
        t0 = time.perf_counter()
        self.set_u(np.zeros(S.N))
        ref = None
        for it in range(31):
            self.shell(capi.GF_OUT_R | capi.GF_OUT_K)
            R = self.residual()
            nrm = np.linalg.norm(R); ref = nrm if it == 0 else ref
            if it > 0 and nrm / ref < 1e-3:
                break
            K = self.K_matrix()
            self.set_u(self.u + spla.splu(K.tocsc()).solve(-R))
        self.shell(capi.GF_OUT_K | capi.GF_OUT_W | capi.GF_OUT_P | capi.GF_OUT_T)
        K = self.K_matrix()
        rhs = self.dWdu.copy(); rhs[S.bc_list] = 0.0
        lam = spla.splu(K.T.tocsc()).solve(rhs)                      # re-factorised for the adjoint, as the reference does
        grads = []
        Ppen = self.om.dRdCP_fields(self.opt_field, self.surf[0], shell=False, penalty=True) if (self.om is not None and self.opt_field) else None
        for i, f in enumerate(self.opt_field):
            Pm = sp.csr_matrix((self.Pv[i], self._idx["P"][i][1], self._idx["P"][i][0]), shape=(S.N, S.P_ncols[i]))
            g = self.dWdP[i] - Pm.T @ lam
            if Ppen is not None:
                g = g - Ppen[i].T @ lam
            grads.append(g)
        Tm = sp.csr_matrix((self.Tv, self._idx["T"][1], self._idx["T"][0]), shape=(S.N, S.n_th))
        grads.append(self.dWdt - Tm.T @ lam)
        self.newton_its = it
        return time.perf_counter() - t0, grads
