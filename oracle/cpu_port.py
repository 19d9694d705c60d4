"""ORACLE / CPU BASELINE (test + measurement infrastructure) -- compiled CPU path.

Drives oracle/c/kl_cpu.cpp (C++/OpenMP shell quadrature, penalty coupling and CSR
scatter, HOST memory) and oracle/c/mf_lu.cpp (multifrontal sparse LU on a nested
dissection ordering, standing in for MUMPS): the "restated reference CPU path"
timed by bench.py (`cpu_baseline`, `--impl reference`) with all host threads,
and a second CPU check of the numpy oracle at sizes the numpy AD cannot reach.
It shares the point-level header (goldfish_b200/csrc/kl_point.cuh, compiled for the
host) with the CUDA kernels, so it is NOT an independent derivation; the numpy jet
oracle (oracle/model.py) is, and tests/test_cpu_port.py holds this port to it.

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
    def __init__(self, problem, opt_field=(), shopt_surf_inds=(), with_penalty_oracle=False):
        if not os.path.exists(LIB):
            build()
        self.lib = C.CDLL(LIB)
        nthr = int(os.environ.get("GFO_THREADS", "0"))
        if nthr > 0:
            self.lib.gfo_set_threads(nthr)
        self.lib.gfo_shell_assemble.argtypes = [C.POINTER(capi.GfModel), C.c_int, C.POINTER(capi.GfShellOut)]
        self.problem = problem
        self.S = S = Symbolic(problem, opt_field, shopt_surf_inds)
        self.opt_field, self.surf = list(opt_field), [list(x) for x in shopt_surf_inds]
        self.om = None        # (the numpy oracle is only used by the tests that compare this port with it)
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
        self._build_penalty()
        self._lu = None

    # ------------------------------------------------------------------ penalty (compiled, oracle/c/kl_cpu.cpp)
    def _build_penalty(self):
        S = self.S
        lib = self.lib
        lib.gfo_penalty_points.argtypes = [C.POINTER(capi.GfModel), C.POINTER(capi.GfPenalty), C.c_int]
        lib.gfo_penalty_gather_R.argtypes = [C.POINTER(capi.GfPenalty), C.c_void_p]
        lib.gfo_penalty_gather_K.argtypes = [C.POINTER(capi.GfPenalty), C.c_void_p]
        lib.gfo_penalty_gather_P.argtypes = [C.POINTER(capi.GfPenalty), C.POINTER(capi.GfPenaltyP)]
        lib.gfo_bc_set_diag.argtypes = [C.POINTER(capi.GfModel), C.c_double]
        pen = S.pen
        q = capi.GfPenalty()
        q.n_eval = pen["n_eval"]
        self._pen_keep = {}
        self.penP = [None for _ in S.opt_field]
        if pen["n_eval"] > 0:
            for k in ("connA", "connB", "connC0", "connC1", "basA", "basB", "basC0", "basC1", "tpar", "alpha",
                      "dofA", "dofB", "R_ptr", "R_item", "R_row", "K_ptr", "K_item", "K_pos"):
                a = np.ascontiguousarray(pen[k]); self._pen_keep[k] = a
                setattr(q, k, _p(a) if a.size else C.c_void_p(0))
            ne = pen["n_eval"]
            self.pen_g, self.pen_Huu, self.pen_HuX = np.zeros(ne * 18), np.zeros(ne * 324), np.zeros(ne * 324)
            q.g, q.Huu, q.HuX = _p(self.pen_g), _p(self.pen_Huu), _p(self.pen_HuX)
            q.nR, q.nK = pen["nR"], pen["nK"]
            for i, pp in enumerate(S.penP):
                if pp.get("n_dest", 0) == 0:
                    continue
                vals = np.zeros(pp["nnz"])
                structs, keep = [], []
                for rd in pp["rounds"]:
                    s = capi.GfPenaltyP()
                    s.n_dest = rd["n_dest"]
                    arrs = [np.ascontiguousarray(rd[k]) for k in ("ptr", "item_eval", "item_code", "pos")]
                    s.ptr, s.item_eval, s.item_code, s.pos = [_p(a) for a in arrs]
                    s.vals, s.field = _p(vals), pp["field"]
                    structs.append(s); keep.append(arrs)
                M = sp.csr_matrix((vals, pp["indices"], pp["indptr"]), shape=(S.N, S.P_ncols[i]))
                self.penP[i] = (M, structs, keep, vals)
        self.pen = q

    def set_u(self, u):
        self.u[:] = u

    def assemble(self, what):
        """Shells + coupling through the compiled port: R (BC rows zeroed), K (BCs applied, in the fixed
        CSR pattern), P / penalty part of dR/dCP, T and the functionals -- all in place, all threads."""
        S = self.S
        if what & capi.GF_OUT_R: self.R[:] = S.f_const
        if what & capi.GF_OUT_K: self.Kv[:] = 0
        if what & capi.GF_OUT_P:
            for v in self.Pv + self.dWdP + self.dVdP: v[:] = 0
        if what & capi.GF_OUT_T:
            self.Tv[:] = 0; self.dWdt[:] = 0; self.dVdt[:] = 0; self.dWdu[:] = 0; self.dt_el[:] = 0
        self.lib.gfo_shell_assemble(C.byref(self.m), what, C.byref(self.o))
        if what & capi.GF_OUT_T:
            for P in S.patches:
                if P.th_kind == 0:
                    sl = self.dt_el[2 * P.el_off:2 * (P.el_off + P.nel)]
                    self.dWdt[P.th_off] += sl[0::2].sum(); self.dVdt[P.th_off] += sl[1::2].sum()
        with_X = 1 if (what & capi.GF_OUT_P and S.opt_field) else 0
        if S.pen["n_eval"] > 0 and (what & (capi.GF_OUT_R | capi.GF_OUT_K) or with_X):
            self.lib.gfo_penalty_points(C.byref(self.m), C.byref(self.pen), with_X)
            if what & capi.GF_OUT_R:
                self.lib.gfo_penalty_gather_R(C.byref(self.pen), _p(self.R))
            if what & capi.GF_OUT_K:
                self.lib.gfo_penalty_gather_K(C.byref(self.pen), _p(self.Kv))
            if with_X:
                for pp in self.penP:
                    if pp is not None:
                        pp[3][:] = 0
                        for rd in pp[1]:
                            self.lib.gfo_penalty_gather_P(C.byref(self.pen), C.byref(rd))
        if what & capi.GF_OUT_R:
            self.R[S.bc_list] = 0.0
        if what & capi.GF_OUT_K:
            self.lib.gfo_bc_set_diag(C.byref(self.m), 1.0)

    shell = assemble        # older name

    def K_matrix(self):
        """The assembled tangent (BCs applied) as scipy CSR sharing the value array."""
        S = self.S
        return sp.csr_matrix((self.Kv, self._idx["K"][1], self._idx["K"][0]), shape=(S.N, S.N))

    def residual(self):
        return self.R.copy()

    def P_matrix(self, i):
        S = self.S
        Pm = sp.csr_matrix((self.Pv[i], self._idx["P"][i][1], self._idx["P"][i][0]), shape=(S.N, S.P_ncols[i]))
        return Pm if self.penP[i] is None else Pm + self.penP[i][0]

    def T_matrix(self):
        S = self.S
        return sp.csr_matrix((self.Tv, self._idx["T"][1], self._idx["T"][0]), shape=(S.N, S.n_th))

    # ------------------------------------------------------------------ linear solves (multifrontal LU)
    def direct_solver(self):
        """Analysis phase (nested dissection + symbolic), once per pattern."""
        if self._lu is None:
            from .mf_solver import MultifrontalLU
            S = self.S
            n_s = S.n_scalar
            rows, cols = [], []
            for P in S.patches:
                cand, mask, _ = S._own_stencil(P)
                a, mm = np.nonzero(mask)
                rows.append(P.cp_off + a); cols.append(P.cp_off + cand[a, mm])
            if len(S.cpl_keys):
                rows.append(S.cpl_keys // n_s); cols.append(S.cpl_keys % n_s)
            r = np.concatenate(rows); c = np.concatenate(cols)
            G = sp.csr_matrix((np.ones(len(r), dtype=np.int8), (r, c)), shape=(n_s, n_s)); G.sum_duplicates()
            X = S.cp0[:, :3] / S.cp0[:, 3:4]
            node_dofs = np.concatenate([np.stack([P.dof_off + f * P.ncp + np.arange(P.ncp) for f in range(3)], 1) for P in S.patches])
            self._lu = MultifrontalLU(G, X, node_dofs, S.N)
        return self._lu

    def solve(self, b, transpose=False, refine=0):
        """solve_Ax_b / solve_ATx_b (utils/opt_utils.py:156-209): a FRESH LU per call, the transposed system
        re-factorised like the reference does (:199-204)."""
        lu = self.direct_solver()
        K = self.K_matrix()
        if transpose:
            KT = K.T.tocsr(); KT.sort_indices()            # explicit transpose, as the reference forms it
            lu.factor(KT, K)
            A = KT
        else:
            KT = K.T.tocsr(); KT.sort_indices()
            lu.factor(K, KT)
            A = K
        x = lu.solve(b)
        for _ in range(refine):
            x = x + lu.solve(b - A @ x)
        return x

    def iteration(self, newton_rtol=1e-3, timings=None, refine=0):
        """One analysis + adjoint iteration of the restated reference CPU path: Newton from u = 0
        (disp_imop.py:38-44) with an LU solve per step, W/V, linearisation (K, dR/dCP_f, dR/dt), adjoint
        solve with the re-factorised transpose, total gradients."""
        S = self.S
        tm = timings if timings is not None else {}
        t0 = time.perf_counter()

        def lap(name, t):
            now = time.perf_counter(); tm[name] = tm.get(name, 0.0) + now - t; return now
        self.set_u(np.zeros(S.N))
        ref = None
        t = t0
        for it in range(31):
            self.assemble(capi.GF_OUT_R | capi.GF_OUT_K)
            t = lap("assemble_RK", t)
            nrm = np.linalg.norm(self.R); ref = nrm if it == 0 else ref
            if it > 0 and nrm / ref < newton_rtol:
                break
            du = self.solve(-self.R, refine=refine)
            self.set_u(self.u + du)
            t = lap("lu_state", t)
        self.assemble(capi.GF_OUT_K | capi.GF_OUT_W | capi.GF_OUT_P | capi.GF_OUT_T)
        t = lap("linearize", t)
        rhs = self.dWdu.copy(); rhs[S.bc_list] = 0.0
        lam = self.solve(rhs, transpose=True, refine=refine)
        t = lap("lu_adjoint", t)
        grads = [self.dWdP[i] - self.P_matrix(i).T @ lam for i in range(len(self.opt_field))]
        grads.append(self.dWdt - self.T_matrix().T @ lam)
        t = lap("gradients", t)
        self.newton_its = it
        self.lam = lam
        self.W, self.V = self.WV[0::2].sum(), self.WV[1::2].sum()
        return time.perf_counter() - t0, grads
