"""ORACLE (test infrastructure, CPU, numpy/scipy) -- the analysis + adjoint path.

CPU restatement of the reference hot path, following the orchestration of
/root/reference/GOLDFISH/nonmatching_opt.py:
  RIGA            :941-948  (assemble_RFE :726-770, extract_nonmatching_vec :639-658)
  dRIGAduIGA      :950-959  (assemble_dRFEduFE :772-841, extract_nonmatching_mat :660-724)
  dRIGAdCPIGA     :992-1004 (assemble_dRFEdCPFE :843-926, opt_utils.transfer_dRmdcpm_sub
                             /root/reference/GOLDFISH/utils/opt_utils.py:231-260)
  dRIGAdh_th      :1006-1015 (assemble_dRFEdh_th :928-938; no BC rows zeroed, no penalty part)
  solve_Ax_b / solve_ATx_b  /root/reference/GOLDFISH/utils/opt_utils.py:156-209 (sparse LU)
  Newton loop     /root/reference/GOLDFISH/operations/disp_imop.py:38-44 (upstream PENGoLINS,
                  from u = 0, stop when |R|/|R0| < rtol)
  W_int, V        /root/reference/GOLDFISH/operations/int_energy_exop.py:55-107,
                  /root/reference/GOLDFISH/operations/volume_exop.py:46-84

The FE detour of the reference (extraction M, M^T K M) is exact interpolation
(SURVEY.md Appendix A.1), so assembling directly in the spline basis with the
same quadrature points gives the same IGA-space operators.

PARITY UNPINNED: none of dolfin / tIGAr / ShNAPr / PENGoLINS / petsc4py is
installed here and the reference's tests assert nothing.
The oracle is anchored on physics known-answers (tests/test_oracle_physics.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import it.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import bspline as bs
from . import quadrature as qd
from .jet import Jet
from .kl_shell import shell_energy_density
from .penalty import penalty_point_energy


class _Patch:
    pass


class OracleModel:
    def __init__(self, problem, chunk_points=1500):
        self.problem = problem
        self.chunk_points = chunk_points
        self.E = problem["E"]
        self.nu = problem["nu"]
        self.alpha = problem.get("penalty_coefficient", 1.0e3)
        self.patches = []
        off = 0
        soff = 0
        toff = 0
        for s, pd in enumerate(problem["patches"]):
            P = _Patch()
            P.p = tuple(pd["p"])
            P.ku, P.kv = [np.asarray(k, dtype=np.float64) for k in pd["knots"]]
            P.nu_ = len(P.ku) - P.p[0] - 1
            P.nv_ = len(P.kv) - P.p[1] - 1
            P.ncp = P.nu_ * P.nv_
            P.cp = np.array(pd["cp"], dtype=np.float64).reshape(P.ncp, 4)
            P.bc = np.asarray(pd.get("bc_dofs", []), dtype=np.int64)
            P.quad_deg = int(pd["quad_deg"])
            P.body_force = np.asarray(pd.get("body_force", (0.0, 0.0, 0.0)), dtype=np.float64)
            P.E = pd.get("E", self.E)
            P.nu = pd.get("nu", self.nu)
            P.off = off
            P.soff = soff
            off += 3 * P.ncp
            soff += P.ncp
            self._setup_quadrature(P)
            th = pd["thickness"]
            P.th_kind = th["kind"]
            self._setup_thickness(P, th)
            P.toff = toff
            toff += P.nth
            self.patches.append(P)
        self.N = off
        self.n_scalar = soff
        self.n_th = toff
        self.u = np.zeros(self.N)
        self.theta = np.concatenate([P.theta0 for P in self.patches])
        self.bc_global = np.concatenate(
            [P.off + P.bc for P in self.patches]) if self.patches else np.zeros(0, int)
        self.interfaces = []
        for it in problem.get("interfaces", []):
            self._setup_interface(it)
        self.point_loads = problem.get("point_loads", [])
        self.f_const = self._const_force()

    # ------------------------------------------------------------------ setup
    def _setup_quadrature(self, P):
        pts, wts, tri = qd.element_rule(P.quad_deg)
        su = bs.unique_spans(P.ku, P.p[0])
        sv = bs.unique_spans(P.kv, P.p[1])
        P.spans_u, P.spans_v = su, sv
        P.neu, P.nev = len(su), len(sv)
        P.nel = P.neu * P.nev
        P.nq = len(wts)
        # element e = eu + ev*neu
        eu = np.tile(np.arange(P.neu), P.nev)
        ev = np.repeat(np.arange(P.nev), P.neu)
        u0 = P.ku[su][eu]; hu = (P.ku[su + 1] - P.ku[su])[eu]
        v0 = P.kv[sv][ev]; hv = (P.kv[sv + 1] - P.kv[sv])[ev]
        xi = np.empty((P.nel, P.nq, 2))
        xi[:, :, 0] = u0[:, None] + hu[:, None] * pts[None, :, 0]
        xi[:, :, 1] = v0[:, None] + hv[:, None] * pts[None, :, 1]
        P.xi = xi
        P.wq = (hu * hv)[:, None] * wts[None, :]
        P.tri = tri
        P.ref_pts = pts
        P.eu, P.ev = eu, ev
        P.hu, P.hv = hu, hv
        w = P.cp[:, 3]
        conn, D = bs.surface_basis(P.ku, P.kv, P.p[0], P.p[1], w, xi.reshape(-1, 2))
        nloc = conn.shape[1]
        P.nloc = nloc
        P.conn = conn.reshape(P.nel, P.nq, nloc)[:, 0, :].copy()
        P.D = D.reshape(P.nel, P.nq, 6, nloc)
        # non-rational basis values (thickness in IGA dofs is not rationalised)
        _, Dn = bs.surface_basis(P.ku, P.kv, P.p[0], P.p[1], np.ones_like(w), xi.reshape(-1, 2))
        P.Nraw = Dn[:, 0, :].reshape(P.nel, P.nq, nloc)

    def _setup_thickness(self, P, th):
        kind = th["kind"]
        vals = np.atleast_1d(np.asarray(th["values"], dtype=np.float64))
        if kind == "const":
            P.nth = 1
            P.tconn = np.zeros((P.nel, 1), dtype=np.int64)
            P.tw = np.ones((P.nel, P.nq, 1))
        elif kind == "linear":
            # CG1 on the two-triangle mesh: vertex dof = iu + iv*(neu+1)
            P.nth = (P.neu + 1) * (P.nev + 1)
            nvu = P.neu + 1
            v00 = P.eu + P.ev * nvu
            P.tconn = np.stack([v00, v00 + 1, v00 + nvu, v00 + nvu + 1], axis=1)
            x, y = P.ref_pts[:, 0], P.ref_pts[:, 1]
            tw = np.zeros((P.nq, 4))
            A = P.tri == 0  # triangle (v00, v10, v11): x>=y
            tw[A, 0] = 1.0 - x[A]; tw[A, 1] = x[A] - y[A]; tw[A, 3] = y[A]
            B = ~A          # triangle (v00, v01, v11): y>=x
            tw[B, 0] = 1.0 - y[B]; tw[B, 2] = y[B] - x[B]; tw[B, 3] = x[B]
            P.tw = np.broadcast_to(tw[None], (P.nel, P.nq, 4)).copy()
        elif kind == "iga":
            P.nth = P.ncp
            P.tconn = P.conn
            P.tw = P.Nraw
        else:
            raise ValueError(kind)
        if vals.size == 1:
            vals = np.full(P.nth, vals[0])
        assert vals.size == P.nth
        P.theta0 = vals.copy()

    def _element_size(self, P, xi):
        """Physical element size at parametric points (H-pen-3): half the
        parametric cell diameter times the Frobenius norm of dX/dxi."""
        conn, D = bs.surface_basis(P.ku, P.kv, P.p[0], P.p[1], P.cp[:, 3], xi)
        Xc = P.cp[:, :3]
        g1 = np.einsum("qa,qac->qc", D[:, 1], Xc[conn])
        g2 = np.einsum("qa,qac->qc", D[:, 2], Xc[conn])
        su = bs.find_span(P.ku, P.p[0], xi[:, 0]); sv = bs.find_span(P.kv, P.p[1], xi[:, 1])
        hu = P.ku[su + 1] - P.ku[su]; hv = P.kv[sv + 1] - P.kv[sv]
        diam = np.sqrt(hu * hu + hv * hv)
        return 0.5 * diam * np.sqrt((g1 * g1).sum(1) + (g2 * g2).sum(1))

    def _thickness_at(self, P, xi, theta):
        """Thickness field evaluated at arbitrary parametric points."""
        if P.th_kind == "const":
            return np.full(len(xi), theta[0])
        if P.th_kind == "iga":
            conn, D = bs.surface_basis(P.ku, P.kv, P.p[0], P.p[1], np.ones(P.ncp), xi)
            return (D[:, 0] * theta[conn]).sum(1)
        # linear
        su = bs.find_span(P.ku, P.p[0], xi[:, 0]); sv = bs.find_span(P.kv, P.p[1], xi[:, 1])
        eu = np.searchsorted(P.spans_u, su); ev = np.searchsorted(P.spans_v, sv)
        x = (xi[:, 0] - P.ku[su]) / (P.ku[su + 1] - P.ku[su])
        y = (xi[:, 1] - P.kv[sv]) / (P.kv[sv + 1] - P.kv[sv])
        nvu = P.neu + 1
        v00 = eu + ev * nvu
        t00, t10, t01, t11 = theta[v00], theta[v00 + 1], theta[v00 + nvu], theta[v00 + nvu + 1]
        lower = x >= y
        return np.where(lower, (1 - x) * t00 + (x - y) * t10 + y * t11,
                        (1 - y) * t00 + (y - x) * t01 + x * t11)

    def _setup_interface(self, it):
        I = _Patch()
        sA, sB = it["patches"]
        I.sA, I.sB = sA, sB
        xiA = np.asarray(it["xi"][0], dtype=np.float64)
        xiB = np.asarray(it["xi"][1], dtype=np.float64)
        I.xi = (xiA, xiB)
        nv = len(xiA)
        I.nv = nv
        PA, PB = self.patches[sA], self.patches[sB]
        I.connA, I.DA = bs.surface_basis(PA.ku, PA.kv, PA.p[0], PA.p[1], PA.cp[:, 3], xiA)
        I.connB, I.DB = bs.surface_basis(PB.ku, PB.kv, PB.p[0], PB.p[1], PB.cp[:, 3], xiB)
        # evaluations: cell c in 0..nv-2, end vertex c and c+1
        c = np.repeat(np.arange(nv - 1), 2)
        v = c + np.tile([0, 1], nv - 1)
        I.ev_c, I.ev_v = c, v
        d = xiA[c + 1] - xiA[c]
        nrm = np.sqrt((d * d).sum(1))
        nrm = np.where(nrm > 0, nrm, 1.0)
        I.tpar = d / nrm[:, None]
        # frozen penalty parameters at mortar vertices
        hA = self._element_size(PA, xiA); hB = self._element_size(PB, xiB)
        h = 0.5 * (hA + hB)
        tA = self._thickness_at(PA, xiA, PA.theta0); tB = self._thickness_at(PB, xiB, PB.theta0)
        adA = self.alpha * PA.E * tA / (h * (1 - PA.nu ** 2))
        adB = self.alpha * PB.E * tB / (h * (1 - PB.nu ** 2))
        arA = self.alpha * PA.E * tA ** 3 / (12 * h * (1 - PA.nu ** 2))
        arB = self.alpha * PB.E * tB ** 3 / (12 * h * (1 - PB.nu ** 2))
        I.alpha_d = np.minimum(adA, adB)
        I.alpha_r = np.minimum(arA, arB)
        self.interfaces.append(I)

    def _const_force(self):
        """Loads that do not depend on u or the geometry: PointSource
        (/root/reference/GOLDFISH/nonmatching_opt.py:735-738): value * N_a(xi)
        added to the residual (homogeneous basis, not rationalised)."""
        f = np.zeros(self.N)
        for pl in self.point_loads:
            P = self.patches[pl["patch"]]
            xi = np.asarray(pl["xi"], dtype=np.float64).reshape(1, 2)
            conn, D = bs.surface_basis(P.ku, P.kv, P.p[0], P.p[1], np.ones(P.ncp), xi)
            f[P.off + pl["field"] * P.ncp + conn[0]] += pl["value"] * D[0, 0]
        # dead edge tractions  R -= int_edge f . phi_a |X_,tau| dtau  (spline.ds,
        # /root/reference/demos_csdl_alpha/thickness_opt/plate_const_th_opt_wint.py:139-150):
        # Gauss-Legendre with m = (quad_deg+2)//2 points per element edge (FFC
        # default facet scheme).  Geometry taken at setup (thickness-opt only).
        for el in self.problem.get("edge_loads", []):
            P = self.patches[el["patch"]]
            d, side = el["direction"], el["side"]   # edge xi_d = side (0 or 1)
            kt = P.kv if d == 0 else P.ku             # tangential knot vector
            pt = P.p[1] if d == 0 else P.p[0]
            spans = bs.unique_spans(kt, pt)
            m = (P.quad_deg + 2) // 2
            g, w = qd.gauss_legendre_01(m)
            t0 = kt[spans]; h = kt[spans + 1] - kt[spans]
            tau = (t0[:, None] + h[:, None] * g[None, :]).ravel()
            wt = (h[:, None] * w[None, :]).ravel()
            fixed = np.full_like(tau, (P.ku if d == 0 else P.kv)[-1] if side == 1
                                 else (P.ku if d == 0 else P.kv)[0])
            xi = np.stack([fixed, tau], axis=1) if d == 0 else np.stack([tau, fixed], axis=1)
            conn, D = bs.surface_basis(P.ku, P.kv, P.p[0], P.p[1], P.cp[:, 3], xi)
            Xc = P.cp[:, :3][conn]
            gt = np.einsum("qa,qac->qc", D[:, 2 if d == 0 else 1], Xc)
            jac = np.sqrt((gt * gt).sum(1))
            trac = np.asarray(el["traction"], dtype=np.float64)
            for c in range(3):
                np.add.at(f, P.off + c * P.ncp + conn, -(wt * jac)[:, None] * D[:, 0] * trac[c])
        return f

    # --------------------------------------------------------------- state
    def set_u(self, u):
        self.u = np.array(u, dtype=np.float64).copy()

    def set_cp(self, field, cp_array, surf_inds=None):
        """update_CPIGA (/root/reference/GOLDFISH/nonmatching_opt.py:495-506):
        homogeneous coordinate `field` of the listed patches."""
        if surf_inds is None:
            surf_inds = range(len(self.patches))
        o = 0
        for s in surf_inds:
            P = self.patches[s]
            P.cp[:, field] = cp_array[o:o + P.ncp]
            o += P.ncp

    def get_cp(self, field, surf_inds=None):
        if surf_inds is None:
            surf_inds = range(len(self.patches))
        return np.concatenate([self.patches[s].cp[:, field] for s in surf_inds])

    def set_thickness(self, theta):
        self.theta = np.array(theta, dtype=np.float64).copy()

    # ------------------------------------------------------- shell quadrature
    def _shell_jets(self, P, sel, with_X):
        """Energy jets at the quadrature points of elements `sel` of patch P.
        Variable layout: [g_u (15), t (1)] (+ [g_X (15)] first if with_X)."""
        conn = P.conn[sel]                       # (ne, nloc)
        D = P.D[sel]                             # (ne, nq, 6, nloc)
        ne, nq = D.shape[0], D.shape[1]
        Xc = P.cp[:, :3][conn]  # (ne, nloc, 3)
        uc = self.u[P.off:P.off + 3 * P.ncp].reshape(3, P.ncp).T[conn]
        GX = np.einsum("eqka,eac->eqkc", D[:, :, 1:6], Xc).reshape(ne * nq, 15)
        Gu = np.einsum("eqka,eac->eqkc", D[:, :, 1:6], uc).reshape(ne * nq, 15)
        th = self.theta[P.toff:P.toff + P.nth]
        tq = (P.tw[sel] * th[P.tconn[sel]][:, None, :]).sum(-1).reshape(-1)
        if with_X:
            V = Jet.variables(np.concatenate([GX, Gu, tq[:, None]], axis=1))
            gX = [V[3 * k:3 * k + 3] for k in range(5)]
            gu = [V[15 + 3 * k:15 + 3 * k + 3] for k in range(5)]
            tj = V[30]
        else:
            V = Jet.variables(np.concatenate([Gu, tq[:, None]], axis=1))
            gX = [[GX[:, 3 * k + c] for c in range(3)] for k in range(5)]
            gu = [V[3 * k:3 * k + 3] for k in range(5)]
            tj = V[15]
        e, J, _, _ = shell_energy_density(gX, gu, tj, P.E, P.nu)
        uq = np.einsum("eqa,eac->eqc", D[:, :, 0], uc).reshape(ne * nq, 3)
        return e, J, uq

    def _chunks(self, P):
        per = max(1, self.chunk_points // P.nq)
        for a in range(0, P.nel, per):
            yield np.arange(a, min(P.nel, a + per))

    # --------------------------------------------------------------- penalty
    def _penalty_jets(self, I, with_X):
        PA, PB = self.patches[I.sA], self.patches[I.sB]
        v, c = I.ev_v, I.ev_c
        n = len(v)

        def fields(P, conn, D, verts):
            Xc = P.cp[:, :3][conn[verts]]
            uc = self.u[P.off:P.off + 3 * P.ncp].reshape(3, P.ncp).T[conn[verts]]
            Dv = D[verts]
            uval = np.einsum("qa,qac->qc", Dv[:, 0], uc)
            du = np.einsum("qka,qac->qkc", Dv[:, 1:3], uc).reshape(len(verts), 6)
            Xval = np.einsum("qa,qac->qc", Dv[:, 0], Xc)
            dX = np.einsum("qka,qac->qkc", Dv[:, 1:3], Xc).reshape(len(verts), 6)
            return uval, du, Xval, dX

        uA, duA, _, dXA = fields(PA, I.connA, I.DA, v)
        uB, duB, _, dXB = fields(PB, I.connB, I.DB, v)
        _, _, XA0, _ = fields(PA, I.connA, I.DA, c)
        _, _, XA1, _ = fields(PA, I.connA, I.DA, c + 1)
        uvals = np.concatenate([uA, duA, uB, duB], axis=1)          # 18
        Xvals = np.concatenate([XA0, XA1, dXA, dXB], axis=1)        # 18
        if with_X:
            V = Jet.variables(np.concatenate([uvals, Xvals], axis=1))
            Xv = V[18:]
        else:
            V = Jet.variables(uvals)
            Xv = [Xvals[:, k] for k in range(18)]
        uAj, duAj = V[0:3], [V[3:6], V[6:9]]
        uBj, duBj = V[9:12], [V[12:15], V[15:18]]
        XA0j, XA1j = Xv[0:3], Xv[3:6]
        dXAj = [Xv[6:9], Xv[9:12]]
        dXBj = [Xv[12:15], Xv[15:18]]
        e = penalty_point_energy(uAj, duAj, uBj, duBj, XA0j, XA1j, dXAj, dXBj,
                                 I.tpar, I.alpha_d[v], I.alpha_r[v])
        return e

    def _penalty_B(self, I):
        """Basis 'B-matrices' for the 18 u-variables of every evaluation:
        list of (patch, conn[n,16], coef[n,3(kinds),16]) per side."""
        v = I.ev_v
        return [(self.patches[I.sA], I.connA[v], I.DA[v][:, 0:3, :]),
                (self.patches[I.sB], I.connB[v], I.DB[v][:, 0:3, :])]

    # -------------------------------------------------------------- residual
    def residual(self, apply_bcs=True, shell=True, penalty=True, const_loads=True):
        R = np.zeros(self.N)
        for P in (self.patches if shell else []):
            for sel in self._chunks(P):
                e, J, uq = self._shell_jets(P, sel, False)
                ne = len(sel)
                g = e.g[:, :15].reshape(ne, P.nq, 5, 3) * P.wq[sel][:, :, None, None]
                Re = np.einsum("eqka,eqkc->eac", P.D[sel][:, :, 1:6], g)
                # body force: -J f . phi_a
                Jw = (J if not isinstance(J, Jet) else J.v).reshape(ne, P.nq) * P.wq[sel]
                Re -= np.einsum("eqa,eq,c->eac", P.D[sel][:, :, 0], Jw, P.body_force)
                for c in range(3):
                    np.add.at(R, P.off + c * P.ncp + P.conn[sel], Re[:, :, c])
        for I in (self.interfaces if penalty else []):
            e = self._penalty_jets(I, False)
            for side, (P, conn, coef) in enumerate(self._penalty_B(I)):
                g = e.g[:, 9 * side:9 * side + 9].reshape(-1, 3, 3)  # (n,kind,comp)
                Re = np.einsum("qka,qkc->qac", coef, g)
                for c in range(3):
                    np.add.at(R, P.off + c * P.ncp + conn, Re[:, :, c])
        if const_loads:
            R += self.f_const
        if apply_bcs:
            R[self.bc_global] = 0.0
        return R

    # ------------------------------------------------------------- stiffness
    def stiffness(self, apply_bcs=True, shell=True, penalty=True):
        rows, cols, vals = [], [], []
        for P in (self.patches if shell else []):
            for sel in self._chunks(P):
                e, J, uq = self._shell_jets(P, sel, False)
                ne = len(sel)
                H = e.h[:, :15, :15].reshape(ne, P.nq, 5, 3, 5, 3) * \
                    P.wq[sel][:, :, None, None, None, None]
                D5 = P.D[sel][:, :, 1:6]
                Ke = np.einsum("eqka,eqkilj,eqlb->eaibj", D5, H, D5, optimize=True)
                r = (P.off + np.arange(3)[None, None, :] * P.ncp + P.conn[sel][:, :, None])
                rows.append(np.broadcast_to(r[:, :, :, None, None], Ke.shape).ravel())
                cols.append(np.broadcast_to(r[:, None, None, :, :], Ke.shape).ravel())
                vals.append(Ke.ravel())
        for I in (self.interfaces if penalty else []):
            e = self._penalty_jets(I, False)
            Bs = self._penalty_B(I)
            for s0, (P0, conn0, coef0) in enumerate(Bs):
                for s1, (P1, conn1, coef1) in enumerate(Bs):
                    H = e.h[:, 9 * s0:9 * s0 + 9, 9 * s1:9 * s1 + 9].reshape(-1, 3, 3, 3, 3)
                    Ke = np.einsum("qka,qkilj,qlb->qaibj", coef0, H, coef1, optimize=True)
                    r = (P0.off + np.arange(3)[None, None, :] * P0.ncp + conn0[:, :, None])
                    c = (P1.off + np.arange(3)[None, None, :] * P1.ncp + conn1[:, :, None])
                    rows.append(np.broadcast_to(r[:, :, :, None, None], Ke.shape).ravel())
                    cols.append(np.broadcast_to(c[:, None, None, :, :], Ke.shape).ravel())
                    vals.append(Ke.ravel())
        K = self._to_csr(rows, cols, vals, (self.N, self.N))
        if apply_bcs:
            K = self._bc_rows_cols(K, diag=1.0)
        return K

    @staticmethod
    def _to_csr(rows, cols, vals, shape):
        if not rows:
            return sp.csr_matrix(shape)
        A = sp.coo_matrix((np.concatenate(vals),
                           (np.concatenate(rows), np.concatenate(cols))), shape=shape).tocsr()
        A.sum_duplicates()
        A.sort_indices()
        return A

    def _bc_rows_cols(self, K, diag):
        """apply_bcs_mat / zeroRowsColumns keeping the sparsity pattern
        (/root/reference/GOLDFISH/nonmatching_opt.py:693-700)."""
        K = K.tocsr().copy()
        mask = np.zeros(K.shape[0], dtype=bool)
        mask[self.bc_global] = True
        rowidx = np.repeat(np.arange(K.shape[0]), np.diff(K.indptr))
        kill = mask[rowidx] | mask[K.indices]
        K.data[kill] = 0.0
        d = (rowidx == K.indices) & mask[rowidx]
        K.data[d] = diag
        return K

    def _bc_rows(self, A):
        A = A.tocsr().copy()
        mask = np.zeros(A.shape[0], dtype=bool)
        mask[self.bc_global] = True
        rowidx = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
        A.data[mask[rowidx]] = 0.0
        return A

    # ---------------------------------------------------------------- dR/dCP
    def dRdCP(self, field, surf_inds=None, apply_bcs=True):
        """d R_IGA / d (homogeneous CP coordinate `field`) for the patches in
        surf_inds; columns = concatenated scalar CP dofs of those patches."""
        return self.dRdCP_fields([field], surf_inds, apply_bcs)[0]

    def dRdCP_fields(self, fields, surf_inds=None, apply_bcs=True, shell=True, penalty=True):
        """Same for several fields sharing one AD pass (all fields on the same
        patch list)."""
        if surf_inds is None:
            surf_inds = list(range(len(self.patches)))
        coloff = {}
        o = 0
        for s in surf_inds:
            coloff[s] = o
            o += self.patches[s].ncp
        ncol = o
        nf = len(fields)
        rows, cols = [], []
        vals = [[] for _ in fields]
        for s in (surf_inds if shell else []):
            P = self.patches[s]
            for sel in self._chunks(P):
                e, J, uq = self._shell_jets(P, sel, True)
                ne = len(sel)
                wq = P.wq[sel]
                # the X-variables of the jet carry x = X + u with them, so this
                # mixed block is the total derivative wrt X at fixed u_hom
                Hux_all = e.h[:, 15:30, 0:15].reshape(ne, P.nq, 5, 3, 5, 3)
                D5 = P.D[sel][:, :, 1:6]
                r = (P.off + np.arange(3)[None, None, :] * P.ncp + P.conn[sel][:, :, None])
                c = coloff[s] + P.conn[sel]
                shape = (ne, P.nloc, 3, P.nloc)
                rows.append(np.broadcast_to(r[:, :, :, None], shape).ravel())
                cols.append(np.broadcast_to(c[:, None, None, :], shape).ravel())
                for k, field in enumerate(fields):
                    Hux = Hux_all[:, :, :, :, :, field] * wq[:, :, None, None, None]
                    Ae = np.einsum("eqka,eqkil,eqlb->eaib", D5, Hux, D5, optimize=True)
                    # body-force part: -f_i phi_a dJ/dg_X[l,field] D_l phi_b
                    dJ = J.g[:, 0:15].reshape(ne, P.nq, 5, 3)[:, :, :, field] * wq[:, :, None]
                    Ae -= np.einsum("eqa,i,eql,eqlb->eaib", P.D[sel][:, :, 0], P.body_force, dJ, D5,
                                    optimize=True)
                    vals[k].append(Ae.ravel())
        for I in (self.interfaces if penalty else []):
            if I.sA not in coloff and I.sB not in coloff:
                continue
            e = self._penalty_jets(I, True)
            Bs = self._penalty_B(I)
            PA, PB = self.patches[I.sA], self.patches[I.sB]
            v, c_ = I.ev_v, I.ev_c
            HXt = e.h[:, 0:18, 18:36]   # dx = dX + du is formed inside the jets
            Xblocks = []
            if I.sA in coloff:
                Xblocks.append((PA, I.connA[c_], I.DA[c_][:, 0, :], 0))
                Xblocks.append((PA, I.connA[c_ + 1], I.DA[c_ + 1][:, 0, :], 3))
                Xblocks.append((PA, I.connA[v], I.DA[v][:, 1, :], 6))
                Xblocks.append((PA, I.connA[v], I.DA[v][:, 2, :], 9))
            if I.sB in coloff:
                Xblocks.append((PB, I.connB[v], I.DB[v][:, 1, :], 12))
                Xblocks.append((PB, I.connB[v], I.DB[v][:, 2, :], 15))
            for side, (P0, conn0, coef0) in enumerate(Bs):
                for (P1, conn1, coef1, xvar) in Xblocks:
                    s1 = self.patches.index(P1)
                    r = (P0.off + np.arange(3)[None, None, :] * P0.ncp + conn0[:, :, None])
                    c = coloff[s1] + conn1
                    shape = (len(v), conn0.shape[1], 3, conn1.shape[1])
                    rows.append(np.broadcast_to(r[:, :, :, None], shape).ravel())
                    cols.append(np.broadcast_to(c[:, None, None, :], shape).ravel())
                    for k, field in enumerate(fields):
                        Hs = HXt[:, 9 * side:9 * side + 9, xvar + field].reshape(-1, 3, 3)  # (n,kind,comp)
                        Ae = np.einsum("qka,qki,qb->qaib", coef0, Hs, coef1, optimize=True)
                        vals[k].append(Ae.ravel())
        out = []
        for k in range(nf):
            A = self._to_csr(rows, cols, vals[k], (self.N, ncol))
            if apply_bcs:
                A = self._bc_rows(A)
            out.append(A)
        return out

    # ----------------------------------------------------------------- dR/dt
    def dRdt(self):
        """dRIGAdh_th: shell part only, no BC rows zeroed
        (/root/reference/GOLDFISH/nonmatching_opt.py:928-938,1006-1015)."""
        rows, cols, vals = [], [], []
        for P in self.patches:
            for sel in self._chunks(P):
                e, J, uq = self._shell_jets(P, sel, False)
                ne = len(sel)
                Hut = e.h[:, 0:15, 15].reshape(ne, P.nq, 5, 3) * P.wq[sel][:, :, None, None]
                D5 = P.D[sel][:, :, 1:6]
                Ae = np.einsum("eqka,eqki,eqm->eaim", D5, Hut, P.tw[sel], optimize=True)
                r = (P.off + np.arange(3)[None, None, :] * P.ncp + P.conn[sel][:, :, None])
                c = P.toff + P.tconn[sel]
                rows.append(np.broadcast_to(r[:, :, :, None], Ae.shape).ravel())
                cols.append(np.broadcast_to(c[:, None, None, :], Ae.shape).ravel())
                vals.append(Ae.ravel())
        return self._to_csr(rows, cols, vals, (self.N, self.n_th))

    # ------------------------------------------------------------ functionals
    def energy(self):
        W = 0.0
        for P in self.patches:
            for sel in self._chunks(P):
                e, J, uq = self._shell_jets(P, sel, False)
                W += float((e.v.reshape(len(sel), P.nq) * P.wq[sel]).sum())
        return W

    def dWdu(self, apply_bcs=True):
        g = np.zeros(self.N)
        for P in self.patches:
            for sel in self._chunks(P):
                e, J, uq = self._shell_jets(P, sel, False)
                ne = len(sel)
                gg = e.g[:, :15].reshape(ne, P.nq, 5, 3) * P.wq[sel][:, :, None, None]
                Re = np.einsum("eqka,eqkc->eac", P.D[sel][:, :, 1:6], gg)
                for c in range(3):
                    np.add.at(g, P.off + c * P.ncp + P.conn[sel], Re[:, :, c])
        if apply_bcs:
            g[self.bc_global] = 0.0
        return g

    def dWdCP(self, field, surf_inds=None):
        if surf_inds is None:
            surf_inds = list(range(len(self.patches)))
        out = []
        for s in surf_inds:
            P = self.patches[s]
            g = np.zeros(P.ncp)
            for sel in self._chunks(P):
                e, J, uq = self._shell_jets(P, sel, True)
                ne = len(sel)
                gx = e.g[:, 0:15].reshape(ne, P.nq, 5, 3)[:, :, :, field]
                gx = gx * P.wq[sel][:, :, None]
                Ge = np.einsum("eqlb,eql->eb", P.D[sel][:, :, 1:6], gx)
                np.add.at(g, P.conn[sel], Ge)
            out.append(g)
        return np.concatenate(out)

    def dWdt(self):
        g = np.zeros(self.n_th)
        for P in self.patches:
            for sel in self._chunks(P):
                e, J, uq = self._shell_jets(P, sel, False)
                ne = len(sel)
                gt = e.g[:, 15].reshape(ne, P.nq) * P.wq[sel]
                Ge = np.einsum("eq,eqm->em", gt, P.tw[sel])
                np.add.at(g, P.toff + P.tconn[sel], Ge)
        return g

    def _area_jets(self, P, sel):
        conn = P.conn[sel]
        D = P.D[sel]
        ne, nq = D.shape[0], D.shape[1]
        Xc = P.cp[:, :3][conn]
        G = np.einsum("eqka,eac->eqkc", D[:, :, 1:3], Xc).reshape(ne * nq, 6)
        V = Jet.variables(G)
        g1, g2 = V[0:3], V[3:6]
        from .jet import dot
        a11, a22, a12 = dot(g1, g1), dot(g2, g2), dot(g1, g2)
        return (a11 * a22 - a12 * a12).sqrt()

    def volume(self, surf_inds=None):
        V = 0.0
        for s, P in enumerate(self.patches):
            if surf_inds is not None and s not in surf_inds:
                continue
            th = self.theta[P.toff:P.toff + P.nth]
            for sel in self._chunks(P):
                J = self._area_jets(P, sel).v.reshape(len(sel), P.nq)
                tq = (P.tw[sel] * th[P.tconn[sel]][:, None, :]).sum(-1)
                V += float((J * tq * P.wq[sel]).sum())
        return V

    def dVdt(self, surf_inds=None):
        g = np.zeros(self.n_th)
        for s, P in enumerate(self.patches):
            if surf_inds is not None and s not in surf_inds:
                continue
            for sel in self._chunks(P):
                J = self._area_jets(P, sel).v.reshape(len(sel), P.nq)
                Ge = np.einsum("eq,eqm->em", J * P.wq[sel], P.tw[sel])
                np.add.at(g, P.toff + P.tconn[sel], Ge)
        return g

    def dVdCP(self, field, surf_inds=None, vol_surf_inds=None):
        if surf_inds is None:
            surf_inds = list(range(len(self.patches)))
        out = []
        for s in surf_inds:
            P = self.patches[s]
            g = np.zeros(P.ncp)
            if vol_surf_inds is None or s in vol_surf_inds:
                th = self.theta[P.toff:P.toff + P.nth]
                for sel in self._chunks(P):
                    ne = len(sel)
                    Jj = self._area_jets(P, sel)
                    tq = (P.tw[sel] * th[P.tconn[sel]][:, None, :]).sum(-1)
                    dJ = Jj.g.reshape(ne, P.nq, 2, 3)[:, :, :, field] * (tq * P.wq[sel])[:, :, None]
                    Ge = np.einsum("eqlb,eql->eb", P.D[sel][:, :, 1:3], dJ)
                    np.add.at(g, P.conn[sel], Ge)
            out.append(g)
        return np.concatenate(out)

    # ---------------------------------------------------------------- solves
    def solve(self, K, b, transpose=False, refine=0):
        """solve_Ax_b / solve_ATx_b: sparse LU (utils/opt_utils.py:176,204).

        refine > 0: that many steps of iterative refinement with the residual
        accumulated in extended precision (numpy longdouble).  The reference's MUMPS
        solve has none; the goldens use it so that they hold the solution of the
        linear system itself (kappa ~ 1e12 on the C1 plate makes a bare LU solve
        good to ~2e-8 only), which is what an iterative solver converges to."""
        A = (K.T if transpose else K).tocsr()
        lu = spla.splu(A.tocsc())
        b = np.asarray(b, dtype=np.float64)
        x = lu.solve(b)
        if refine:
            rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
            data = A.data.astype(np.longdouble)
            for _ in range(refine):
                prod = data * x.astype(np.longdouble)[A.indices]
                r = b.astype(np.longdouble)
                Ax = np.zeros(A.shape[0], dtype=np.longdouble)
                np.add.at(Ax, rows, prod)
                x = x + lu.solve(np.asarray(r - Ax, dtype=np.float64))
        return x

    def solve_nonlinear(self, max_it=30, rtol=1e-3, verbose=False, refine=0):
        """PENGoLINS solve_nonlinear_nonmatching_problem(iga_dofs=True,
        zero_mortar_funcs=True): Newton from u = 0."""
        self.u = np.zeros(self.N)
        ref = None
        hist = []
        for it in range(max_it + 1):
            R = self.residual()
            nrm = np.linalg.norm(R)
            if it == 0:
                ref = nrm
            rel = nrm / ref if ref > 0 else 0.0
            hist.append(rel)
            if verbose:
                print("newton", it, nrm, rel)
            if it > 0 and rel < rtol:
                break
            if it == max_it:
                break
            K = self.stiffness()
            du = self.solve(K, -R, refine=refine)
            self.u = self.u + du
        self.newton_history = hist
        return self.u.copy()

    def solve_linear(self):
        """solve_linear_nonmatching_problem: one Newton step from u = 0."""
        self.u = np.zeros(self.N)
        R = self.residual()
        K = self.stiffness()
        self.u = self.solve(K, -R)
        return self.u.copy()
