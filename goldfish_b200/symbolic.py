"""Symbolic phase (host, once per topology): DoF maps, CSR patterns, element
colouring, basis tables and the deterministic gather lists of the coupling.

This is what DOLFIN's sparsity builder + PETSc's symbolic MatPtAP +
``create_nest_PETScMat -> convert('seqaij')`` do for the reference
(/root/reference/GOLDFISH/nonmatching_opt.py:660-724): here the pattern of the
IGA-space operator is written down directly.

Pattern definitions (mirrored by the oracle, bit-exact, sorted columns):
  K      rows/cols = displacement dofs; entry iff the two CPs share a Bezier
         element (all 3x3 field pairs) or both support one mortar evaluation.
         Structural zeros of BC rows/columns are kept (zeroRowsColumns keeps
         the pattern).
  dR/dCP_f  shell part: same element rule, columns = scalar CPs of the patches
         in shopt_surf_inds[f]; penalty part kept as a separate small CSR.
  dR/dt  columns = thickness dofs touched by the elements of the row CP.
"""
import numpy as np
from . import bsplines as bsp
from . import quadrature as quad

TH_KIND = {"const": 0, "linear": 1, "iga": 2}


def _ragged_arange(counts):
    """concatenate [arange(c) for c in counts] without a Python loop."""
    counts = np.asarray(counts, dtype=np.int64)
    total = int(counts.sum())
    starts = np.cumsum(counts) - counts
    return np.arange(total, dtype=np.int64) - np.repeat(starts, counts)


class PatchSym:
    pass


class Symbolic:
    def __init__(self, problem, opt_field=(), shopt_surf_inds=(), build_transpose=True, own_patches=None):
        """own_patches: optional boolean array over the patches (patch-sharded runs): the coupling gather lists
        are then built only for destinations whose ROW belongs to an own patch -- the only ones that rank uses --
        which divides the host time and memory of the coupling set-up by about the number of ranks."""
        self.problem = problem
        self.own_patches = None if own_patches is None else np.asarray(own_patches, dtype=bool)
        self.opt_field = list(opt_field)
        self.shopt_surf_inds = [list(x) for x in shopt_surf_inds]
        self.alpha = float(problem.get("penalty_coefficient", 1.0e3))
        self._patches(problem)
        self._elements()
        self._K_pattern()
        self._P_patterns()
        self._T_pattern()
        self._penalty()
        self._const_force()

    # ------------------------------------------------------------------ patches
    def _patches(self, problem):
        self.patches = []
        cp_off = dof_off = th_off = su_off = sv_off = cu_off = cv_off = el_off = 0
        tabs_u, tabs_v, fcu, fcv, shu, shv = [], [], [], [], [], []
        dirs = {k: [] for k in ("cp_lo_u", "cp_hi_u", "el_lo_u", "el_hi_u",
                                "cp_lo_v", "cp_hi_v", "el_lo_v", "el_hi_v")}
        nq = None
        cps, thetas, bcs = [], [], []
        for s, pd in enumerate(problem["patches"]):
            P = PatchSym()
            P.index = s
            P.p = tuple(pd["p"])
            if P.p != (3, 3):
                raise ValueError("the CUDA path is specialised for bicubic patches (p = 3); got %r" % (P.p,))
            P.ku, P.kv = [np.asarray(k, dtype=np.float64) for k in pd["knots"]]
            P.n_u, P.n_v = bsp.num_basis(P.ku, 3), bsp.num_basis(P.kv, 3)
            P.ncp = P.n_u * P.n_v
            P.cp = np.array(pd["cp"], dtype=np.float64).reshape(P.ncp, 4)
            P.rational = int(not np.all(P.cp[:, 3] == 1.0))
            P.quad_deg = int(pd["quad_deg"])
            pts, wts, tw = quad.span_rule(P.quad_deg)
            if nq is None:
                nq = len(wts); self.ref_pts, self.qw, self.tw_lin = pts, wts, tw; self.quad_deg = P.quad_deg
            elif len(wts) != nq or P.quad_deg != self.quad_deg:
                raise ValueError("all patches must share one quadrature degree")
            for d, (kn, n) in enumerate(((P.ku, P.n_u), (P.kv, P.n_v))):
                spans = bsp.nonempty_spans(kn, 3)
                h = kn[spans + 1] - kn[spans]
                x = kn[spans][:, None] + h[:, None] * pts[None, :, d]
                _, B = bsp.basis_window(kn, 3, x.ravel(), 2)          # (ne*nq, 3, 4)
                tab = B.reshape(len(spans), nq, 3, 4)
                first = (spans - 3).astype(np.int32)
                # stencil / element ranges of each CP index
                el_lo = np.full(n, 1 << 30, dtype=np.int64); el_hi = np.full(n, -1, dtype=np.int64)
                for l in range(4):
                    np.minimum.at(el_lo, first + l, np.arange(len(spans)))
                    np.maximum.at(el_hi, first + l, np.arange(len(spans)))
                cp_lo = first[el_lo]; cp_hi = first[el_hi] + 3
                if d == 0:
                    P.spans_u, P.neu, P.first_u, P.h_u = spans, len(spans), first, h
                    tabs_u.append(tab); fcu.append(first); shu.append(h)
                    P.cp_lo_u, P.cp_hi_u, P.el_lo_u, P.el_hi_u = cp_lo, cp_hi, el_lo, el_hi
                    for k, v in (("cp_lo_u", cp_lo), ("cp_hi_u", cp_hi), ("el_lo_u", el_lo), ("el_hi_u", el_hi)):
                        dirs[k].append(v.astype(np.int32))
                else:
                    P.spans_v, P.nev, P.first_v, P.h_v = spans, len(spans), first, h
                    tabs_v.append(tab); fcv.append(first); shv.append(h)
                    P.cp_lo_v, P.cp_hi_v, P.el_lo_v, P.el_hi_v = cp_lo, cp_hi, el_lo, el_hi
                    for k, v in (("cp_lo_v", cp_lo), ("cp_hi_v", cp_hi), ("el_lo_v", el_lo), ("el_hi_v", el_hi)):
                        dirs[k].append(v.astype(np.int32))
            P.nel = P.neu * P.nev
            th = pd["thickness"]
            P.th_kind = TH_KIND[th["kind"]]
            P.nth = {0: 1, 1: (P.neu + 1) * (P.nev + 1), 2: P.ncp}[P.th_kind]
            vals = np.atleast_1d(np.asarray(th["values"], dtype=np.float64))
            P.theta0 = np.full(P.nth, vals[0]) if vals.size == 1 else vals.copy()
            assert P.theta0.size == P.nth
            P.E = float(pd.get("E", problem["E"])); P.nu = float(pd.get("nu", problem["nu"]))
            P.f = np.asarray(pd.get("body_force", (0.0, 0.0, 0.0)), dtype=np.float64)
            P.bc = np.asarray(pd.get("bc_dofs", []), dtype=np.int64)
            P.cp_off, P.dof_off, P.th_off = cp_off, dof_off, th_off
            P.span_u_off, P.span_v_off, P.cpd_u_off, P.cpd_v_off, P.el_off = su_off, sv_off, cu_off, cv_off, el_off
            cp_off += P.ncp; dof_off += 3 * P.ncp; th_off += P.nth
            su_off += P.neu; sv_off += P.nev; cu_off += P.n_u; cv_off += P.n_v; el_off += P.nel
            cps.append(P.cp); thetas.append(P.theta0); bcs.append(P.dof_off + P.bc)
            P.pcol_off = [-1, -1, -1]
            self.patches.append(P)
        self.nq = nq
        self.N, self.n_scalar, self.n_th, self.num_elements = dof_off, cp_off, th_off, el_off
        self.tab_u = np.ascontiguousarray(np.concatenate(tabs_u)); self.tab_v = np.ascontiguousarray(np.concatenate(tabs_v))
        self.first_cp_u = np.concatenate(fcu).astype(np.int32); self.first_cp_v = np.concatenate(fcv).astype(np.int32)
        self.span_h_u = np.concatenate(shu); self.span_h_v = np.concatenate(shv)
        self.dirs = {k: np.concatenate(v) for k, v in dirs.items()}
        self.cp0 = np.concatenate(cps)
        self.theta0 = np.concatenate(thetas)
        self.bc_list = np.sort(np.concatenate(bcs)).astype(np.int32)
        self.bc_mask = np.zeros(self.N, dtype=np.uint8); self.bc_mask[self.bc_list] = 1
        # column offsets of dR/dCP_f
        self.P_ncols = []
        for fi, field in enumerate(self.opt_field):
            o = 0
            for s in self.shopt_surf_inds[fi]:
                self.patches[s].pcol_off[field] = o
                o += self.patches[s].ncp
            self.P_ncols.append(o)
        self.scalar_patch = np.concatenate([np.full(P.ncp, P.index, dtype=np.int32) for P in self.patches])

    # ----------------------------------------------------------------- elements
    def _elements(self):
        ep, eu, ev = [], [], []
        for P in self.patches:
            e = np.arange(P.nel)
            ep.append(np.full(P.nel, P.index)); eu.append(e % P.neu); ev.append(e // P.neu)
        self.elem_patch = np.concatenate(ep).astype(np.int32)
        self.elem_eu = np.concatenate(eu).astype(np.int32)
        self.elem_ev = np.concatenate(ev).astype(np.int32)
        color = (self.elem_eu % 4) + 4 * (self.elem_ev % 4)
        order = np.argsort(color, kind="stable")
        self.color_elem = order.astype(np.int32)
        self.num_colors = 16
        self.color_ptr = np.searchsorted(color[order], np.arange(17)).astype(np.int32)

    # -------------------------------------------------------------- own stencil
    def _own_stencil(self, P):
        """(cand[ncp,49], mask[ncp,49], S[ncp]) neighbours of each CP (sorted)."""
        I = np.tile(np.arange(P.n_u), P.n_v); J = np.repeat(np.arange(P.n_v), P.n_u)
        dI = np.tile(np.arange(-3, 4), 7); dJ = np.repeat(np.arange(-3, 4), 7)
        Ic = I[:, None] + dI[None, :]; Jc = J[:, None] + dJ[None, :]
        mask = (Ic >= P.cp_lo_u[I][:, None]) & (Ic <= P.cp_hi_u[I][:, None]) & \
               (Jc >= P.cp_lo_v[J][:, None]) & (Jc <= P.cp_hi_v[J][:, None])
        cand = Ic + Jc * P.n_u
        return cand, mask, mask.sum(1)

    # ---------------------------------------------------------------- K pattern
    def _interface_tables(self):
        """Per interface: evaluation tables (host).  Evaluations are (cell c,
        end vertex v) in the order c = 0.., v = c, c+1."""
        self.itf = []
        for it in self.problem.get("interfaces", []):
            sA, sB = it["patches"]
            PA, PB = self.patches[sA], self.patches[sB]
            xiA = np.asarray(it["xi"][0], dtype=np.float64); xiB = np.asarray(it["xi"][1], dtype=np.float64)
            cA, DA = bsp.surface_point_tables(PA.ku, PA.kv, 3, 3, PA.cp[:, 3], xiA)
            cB, DB = bsp.surface_point_tables(PB.ku, PB.kv, 3, 3, PB.cp[:, 3], xiB)
            nv = len(xiA)
            c = np.repeat(np.arange(nv - 1), 2); v = c + np.tile([0, 1], nv - 1)
            d = xiA[c + 1] - xiA[c]
            nrm = np.sqrt((d * d).sum(1)); nrm = np.where(nrm > 0, nrm, 1.0)
            self.itf.append(dict(sA=sA, sB=sB, xiA=xiA, xiB=xiB, connA=cA, DA=DA, connB=cB, DB=DB,
                                 c=c, v=v, tpar=d / nrm[:, None]))

    def _K_pattern(self):
        self._interface_tables()
        n_s = self.n_scalar
        # cross-patch coupled scalar pairs
        keys = []
        for T in self.itf:
            PA, PB = self.patches[T["sA"]], self.patches[T["sB"]]
            gA = PA.cp_off + T["connA"]; gB = PB.cp_off + T["connB"]
            if T["sA"] == T["sB"]:
                raise ValueError("self-intersections are not supported")
            a = np.repeat(gA, 16, axis=1).ravel(); b = np.tile(gB, (1, 16)).ravel()
            keys.append(a.astype(np.int64) * n_s + b); keys.append(b.astype(np.int64) * n_s + a)
        keys = np.unique(np.concatenate(keys)) if keys else np.zeros(0, dtype=np.int64)
        self.cpl_keys = keys
        cr = keys // n_s; cc = keys % n_s
        pr = self.scalar_patch[cr]; pc = self.scalar_patch[cc]
        ncoup = np.bincount(cr, minlength=n_s)
        nlow = np.bincount(cr[pc < pr], minlength=n_s)
        self.row_nlow = (3 * nlow).astype(np.int32)
        # per scalar row: own stencil size
        S_all = np.zeros(n_s, dtype=np.int64)
        sten = []
        for P in self.patches:
            cand, mask, S = self._own_stencil(P)
            sten.append((cand, mask, S))
            S_all[P.cp_off:P.cp_off + P.ncp] = S
        self.S_all = S_all
        # vector rows: patch-major, field-major, cp
        rowlen = np.concatenate([np.tile(3 * (S_all[P.cp_off:P.cp_off + P.ncp] + ncoup[P.cp_off:P.cp_off + P.ncp]), 3)
                                 for P in self.patches])
        indptr = np.zeros(self.N + 1, dtype=np.int64); np.cumsum(rowlen, out=indptr[1:])
        indices = np.empty(int(indptr[-1]), dtype=np.int32)
        for P, (cand, mask, S) in zip(self.patches, sten):
            loc = np.cumsum(mask, axis=1) - 1
            a_idx, m_idx = np.nonzero(mask)
            cols = cand[a_idx, m_idx]; lpos = loc[a_idx, m_idx]
            nl3 = self.row_nlow[P.cp_off + a_idx].astype(np.int64)
            Sa = S[a_idx]
            for i in range(3):
                base = indptr[P.dof_off + i * P.ncp + a_idx] + nl3
                for j in range(3):
                    indices[base + j * Sa + lpos] = P.dof_off + j * P.ncp + cols
        # coupling columns
        if len(keys):
            rstart = np.searchsorted(cr, cr, side="left")         # first pair of the row
            gkey = cr * len(self.patches) + pc
            gstart = np.searchsorted(gkey, gkey, side="left")
            gsize = np.searchsorted(gkey, gkey, side="right") - gstart
            rank = np.arange(len(keys)) - gstart
            lower = pc < pr
            Pr_dof = np.array([P.dof_off for P in self.patches])[pr]; Pr_ncp = np.array([P.ncp for P in self.patches])[pr]
            Pr_cp = np.array([P.cp_off for P in self.patches])[pr]
            Pc_dof = np.array([P.dof_off for P in self.patches])[pc]; Pc_ncp = np.array([P.ncp for P in self.patches])[pc]
            Pc_cp = np.array([P.cp_off for P in self.patches])[pc]
            off_in_row = 3 * (gstart - rstart) + np.where(lower, 0, 3 * S_all[cr])
            self._cpl = dict(cr=cr, cc=cc, off_in_row=off_in_row, gsize=gsize, rank=rank,
                             Pr_dof=Pr_dof, Pr_ncp=Pr_ncp, Pr_cp=Pr_cp, Pc_dof=Pc_dof, Pc_ncp=Pc_ncp, Pc_cp=Pc_cp)
            for i in range(3):
                base = indptr[Pr_dof + i * Pr_ncp + (cr - Pr_cp)] + off_in_row
                for j in range(3):
                    indices[base + j * gsize + rank] = Pc_dof + j * Pc_ncp + (cc - Pc_cp)
        else:
            self._cpl = None
        self.K_indptr, self.K_indices = indptr, indices

    def k_positions(self, r, c):
        """Positions in K.vals of the 3x3 block (row scalar CP r, col scalar CP c):
        array [n,3,3] (i,j).  Works for same-patch (own stencil) and coupled pairs."""
        r = np.asarray(r, dtype=np.int64); c = np.asarray(c, dtype=np.int64)
        pr = self.scalar_patch[r]; pc = self.scalar_patch[c]
        dof = np.array([P.dof_off for P in self.patches]); ncp = np.array([P.ncp for P in self.patches])
        cpo = np.array([P.cp_off for P in self.patches]); nu_ = np.array([P.n_u for P in self.patches])
        pos = np.empty((len(r), 3, 3), dtype=np.int64)
        same = pr == pc
        if same.any():
            rs, cs, ps = r[same], c[same], pr[same]
            lr = rs - cpo[ps]; lc = cs - cpo[ps]
            I = lr % nu_[ps]; J = lr // nu_[ps]; Ic = lc % nu_[ps]; Jc = lc // nu_[ps]
            lo_u = np.empty(len(rs), dtype=np.int64); hi_u = np.empty_like(lo_u); lo_v = np.empty_like(lo_u)
            for P in self.patches:
                m = ps == P.index
                lo_u[m] = P.cp_lo_u[I[m]]; hi_u[m] = P.cp_hi_u[I[m]]; lo_v[m] = P.cp_lo_v[J[m]]
            WI = hi_u - lo_u + 1
            inside = (Ic >= lo_u) & (Ic <= hi_u) & (Jc >= lo_v)
            if not inside.all():
                raise ValueError("same-patch coupling outside the element stencil")
            loc = (Jc - lo_v) * WI + (Ic - lo_u)
            S = self.S_all[rs]
            for i in range(3):
                base = self.K_indptr[dof[ps] + i * ncp[ps] + lr] + self.row_nlow[rs]
                for j in range(3):
                    pos[same, i, j] = base + j * S + loc
        if (~same).any():
            C = self._cpl
            key = r[~same] * self.n_scalar + c[~same]
            k = np.searchsorted(self.cpl_keys, key)
            assert np.all(self.cpl_keys[k] == key)
            for i in range(3):
                base = self.K_indptr[C["Pr_dof"][k] + i * C["Pr_ncp"][k] + (C["cr"][k] - C["Pr_cp"][k])] + C["off_in_row"][k]
                for j in range(3):
                    pos[~same, i, j] = base + j * C["gsize"][k] + C["rank"][k]
        return pos

    # --------------------------------------------------------------- P patterns
    def _P_patterns(self):
        self.P_indptr, self.P_indices = [], []
        cache = {}
        for fi, field in enumerate(self.opt_field):
            key = tuple(self.shopt_surf_inds[fi])
            if key in cache:              # same patch list => same pattern and column offsets: share the arrays
                self.P_indptr.append(cache[key][0]); self.P_indices.append(cache[key][1])
                continue
            rowlen = np.zeros(self.N, dtype=np.int64)
            for s in self.shopt_surf_inds[fi]:
                P = self.patches[s]
                S = self.S_all[P.cp_off:P.cp_off + P.ncp]
                rowlen[P.dof_off:P.dof_off + 3 * P.ncp] = np.tile(S, 3)
            indptr = np.zeros(self.N + 1, dtype=np.int64); np.cumsum(rowlen, out=indptr[1:])
            indices = np.empty(int(indptr[-1]), dtype=np.int32)
            for s in self.shopt_surf_inds[fi]:
                P = self.patches[s]
                cand, mask, S = self._own_stencil(P)
                loc = np.cumsum(mask, axis=1) - 1
                a_idx, m_idx = np.nonzero(mask)
                for i in range(3):
                    base = indptr[P.dof_off + i * P.ncp + a_idx]
                    indices[base + loc[a_idx, m_idx]] = P.pcol_off[field] + cand[a_idx, m_idx]
            self.P_indptr.append(indptr); self.P_indices.append(indices)
            cache[key] = (indptr, indices)

    # ---------------------------------------------------------------- T pattern
    def _T_pattern(self):
        rowlens, cols_all = [], []
        for P in self.patches:
            I = np.tile(np.arange(P.n_u), P.n_v); J = np.repeat(np.arange(P.n_v), P.n_u)
            if P.th_kind == 0:
                cnt = np.ones(P.ncp, dtype=np.int64)
                cols = np.full(P.ncp, P.th_off, dtype=np.int64)
            elif P.th_kind == 1:
                wu = P.el_hi_u[I] - P.el_lo_u[I] + 2; wv = P.el_hi_v[J] - P.el_lo_v[J] + 2
                cnt = wu * wv
                k = _ragged_arange(cnt)
                a = np.repeat(np.arange(P.ncp), cnt)
                vI = P.el_lo_u[I][a] + k % wu[a]; vJ = P.el_lo_v[J][a] + k // wu[a]
                cols = P.th_off + vI + vJ * (P.neu + 1)
            else:
                cand, mask, S = self._own_stencil(P)
                cnt = S
                cols = P.th_off + cand[mask]
            rowlens.append(np.tile(cnt, 3)); cols_all.append(np.tile(cols, 3))
        rowlen = np.concatenate(rowlens)
        self.T_indptr = np.zeros(self.N + 1, dtype=np.int64); np.cumsum(rowlen, out=self.T_indptr[1:])
        self.T_indices = np.concatenate(cols_all).astype(np.int32)

    # ------------------------------------------------------------------ penalty
    def _size_and_thickness_at(self, P, xi):
        """Element size (half the parametric cell diameter times |dX/dxi|_F) and
        initial thickness at mortar points: inputs of the frozen alpha_d, alpha_r
        (hypothesis H-pen-3 of the oracle; nonmatching_opt.py:1115 uses them as stored)."""
        conn, D = bsp.surface_point_tables(P.ku, P.kv, 3, 3, P.cp[:, 3], xi)
        X = P.cp[:, :3][conn]
        g1 = np.einsum("qa,qac->qc", D[:, 1], X); g2 = np.einsum("qa,qac->qc", D[:, 2], X)
        su = bsp.span_index(P.ku, 3, xi[:, 0]); sv = bsp.span_index(P.kv, 3, xi[:, 1])
        hu = P.ku[su + 1] - P.ku[su]; hv = P.kv[sv + 1] - P.kv[sv]
        h = 0.5 * np.sqrt(hu * hu + hv * hv) * np.sqrt((g1 * g1).sum(1) + (g2 * g2).sum(1))
        if P.th_kind == 0:
            t = np.full(len(xi), P.theta0[0])
        elif P.th_kind == 2:
            c1, D1 = bsp.surface_point_tables(P.ku, P.kv, 3, 3, np.ones(P.ncp), xi)
            t = (D1[:, 0] * P.theta0[c1]).sum(1)
        else:
            eu = np.searchsorted(P.spans_u, su); ev = np.searchsorted(P.spans_v, sv)
            x = (xi[:, 0] - P.ku[su]) / hu; y = (xi[:, 1] - P.kv[sv]) / hv
            v00 = eu + ev * (P.neu + 1); th = P.theta0
            t00, t10, t01, t11 = th[v00], th[v00 + 1], th[v00 + P.neu + 1], th[v00 + P.neu + 2]
            t = np.where(x >= y, (1 - x) * t00 + (x - y) * t10 + y * t11, (1 - y) * t00 + (y - x) * t01 + x * t11)
        return h, t

    def _penalty(self):
        ev_lists = {k: [] for k in ("connA", "connB", "connC0", "connC1", "basA", "basB", "basC0", "basC1",
                                    "tpar", "alpha", "dofA", "dofB")}
        self.itf_alpha = []
        self.itf_eval_range = []; n_acc = 0
        override = self.problem.get("alpha_override")
        for ii, T in enumerate(self.itf):
            PA, PB = self.patches[T["sA"]], self.patches[T["sB"]]
            c, v = T["c"], T["v"]
            hA, tA = self._size_and_thickness_at(PA, T["xiA"]); hB, tB = self._size_and_thickness_at(PB, T["xiB"])
            h = 0.5 * (hA + hB)
            ad = np.minimum(self.alpha * PA.E * tA / (h * (1 - PA.nu ** 2)), self.alpha * PB.E * tB / (h * (1 - PB.nu ** 2)))
            ar = np.minimum(self.alpha * PA.E * tA ** 3 / (12 * h * (1 - PA.nu ** 2)),
                            self.alpha * PB.E * tB ** 3 / (12 * h * (1 - PB.nu ** 2)))
            if override is not None:       # coarse level of the preconditioner: keep the fine penalty stiffness
                ad, ar = override[ii]
            self.itf_alpha.append((ad, ar))
            n = len(v)
            self.itf_eval_range.append((n_acc, n_acc + n)); n_acc += n
            ev_lists["connA"].append(PA.cp_off + T["connA"][v]); ev_lists["connB"].append(PB.cp_off + T["connB"][v])
            ev_lists["connC0"].append(PA.cp_off + T["connA"][c]); ev_lists["connC1"].append(PA.cp_off + T["connA"][c + 1])
            ev_lists["basA"].append(T["DA"][v][:, 0:3, :]); ev_lists["basB"].append(T["DB"][v][:, 0:3, :])
            ev_lists["basC0"].append(T["DA"][c][:, 0, :]); ev_lists["basC1"].append(T["DA"][c + 1][:, 0, :])
            ev_lists["tpar"].append(T["tpar"]); ev_lists["alpha"].append(np.stack([ad[v], ar[v]], axis=1))
            ev_lists["dofA"].append(np.tile([PA.dof_off, PA.ncp, PA.cp_off], (n, 1)))
            ev_lists["dofB"].append(np.tile([PB.dof_off, PB.ncp, PB.cp_off], (n, 1)))
        pen = {}
        if self.itf:
            for k, lst in ev_lists.items():
                a = np.concatenate(lst)
                pen[k] = np.ascontiguousarray(a.astype(np.int32) if k.startswith(("conn", "dof")) else a.astype(np.float64))
            n_eval = len(pen["tpar"])
        else:
            n_eval = 0
        pen["n_eval"] = n_eval
        self.pen = pen
        if n_eval == 0:
            return
        dof = np.array([P.dof_off for P in self.patches]); ncp = np.array([P.ncp for P in self.patches])
        cpo = np.array([P.cp_off for P in self.patches])
        # ---- R gather: destination = scalar CP (3 rows) ----
        nodes = np.concatenate([pen["connA"], pen["connB"]], axis=1).astype(np.int64)   # [n_eval, 32]
        item = (np.arange(n_eval, dtype=np.int64)[:, None] * 32 + np.arange(32)[None, :]).ravel()
        dest = nodes.ravel()
        if self.own_patches is not None:                 # sharded: destinations in own patches only
            keep = self.own_patches[self.scalar_patch[dest]]
            item, dest = item[keep], dest[keep]
        order = np.argsort(dest, kind="stable")
        ud, start = np.unique(dest[order], return_index=True)
        pen["R_ptr"] = np.append(start, len(dest)).astype(np.int64)
        pen["R_item"] = item[order].astype(np.int32)
        ps = self.scalar_patch[ud]
        pen["R_row"] = np.stack([dof[ps] + i * ncp[ps] + (ud - cpo[ps]) for i in range(3)], axis=1).astype(np.int32)
        pen["nR"] = len(ud)
        pen["R_dest_patch"] = ps.astype(np.int32)
        # ---- K gather: destination = (row CP, col CP) ----
        ev_ids = np.arange(n_eval, dtype=np.int64)
        nd = nodes
        if self.own_patches is not None:                 # sharded: evaluations that touch an own patch ...
            touch = self.own_patches[self.scalar_patch[nodes[:, 0]]] | self.own_patches[self.scalar_patch[nodes[:, 16]]]
            ev_ids, nd = ev_ids[touch], nodes[touch]
        r = np.repeat(nd, 32, axis=1).ravel(); c = np.tile(nd, (1, 32)).ravel()
        la = np.repeat(np.arange(32), 32); lb = np.tile(np.arange(32), 32)
        item = (ev_ids[:, None] * 1024 + (la * 32 + lb)[None, :]).ravel()
        if self.own_patches is not None:                 # ... and of those the destinations whose row is owned
            keep = self.own_patches[self.scalar_patch[r]]
            r, c, item = r[keep], c[keep], item[keep]
        key = r * self.n_scalar + c
        order = np.argsort(key, kind="stable")
        uk, start = np.unique(key[order], return_index=True)
        pen["K_ptr"] = np.append(start, len(key)).astype(np.int64)
        pen["K_item"] = item[order].astype(np.int32)
        ur, uc = uk // self.n_scalar, uk % self.n_scalar
        pos = self.k_positions(ur, uc)
        # BC masking (zeroRowsColumns): entries in BC rows/cols receive nothing
        pr, pc = self.scalar_patch[ur], self.scalar_patch[uc]
        for i in range(3):
            rbc = self.bc_mask[dof[pr] + i * ncp[pr] + (ur - cpo[pr])] != 0
            for j in range(3):
                cbc = self.bc_mask[dof[pc] + j * ncp[pc] + (uc - cpo[pc])] != 0
                pos[:, i, j] = np.where(rbc | cbc, -1, pos[:, i, j])
        pen["K_pos"] = np.ascontiguousarray(pos.reshape(-1, 9))
        pen["nK"] = len(uk)
        pen["K_dest_patch"] = pr.astype(np.int32)
        # ---- dR/dCP_f penalty part ----
        # the gather structure depends on the patch list only: build it once per distinct list
        self.penP, cache = [], {}
        for fi, field in enumerate(self.opt_field):
            key = tuple(self.shopt_surf_inds[fi])
            if key not in cache:
                cache[key] = self._penalty_P(field, nodes)
            pp = dict(cache[key]); pp["field"] = field
            self.penP.append(pp)

    def _penalty_P(self, field, nodes):
        """Gather lists of the penalty part of dR/dCP_f.  Built interface by interface (the item lists hold
        32 x 6 x 16 entries per evaluation: one interface at a time keeps the temporaries small on coupling-heavy
        models), then grouped into ROUNDS of interfaces whose destination sets are disjoint: inside a round every
        destination has one owner thread, rounds run one after the other and accumulate (+=) -- deterministic."""
        pen = self.pen
        n_s = self.n_scalar
        dof = np.array([P.dof_off for P in self.patches]); ncp = np.array([P.ncp for P in self.patches])
        cpo = np.array([P.cp_off for P in self.patches])
        pcol = np.array([P.pcol_off[field] for P in self.patches])
        names = ("connC0", "connC1", "connA", "connA", "connB", "connB")          # xblock 0..5
        la16 = np.repeat(np.arange(32, dtype=np.int32), 16); lb32 = np.tile(np.arange(16, dtype=np.int32), 32)
        per_itf = []
        for ii, (e0, e1) in enumerate(self.itf_eval_range):
            ne = e1 - e0
            rs, cs, evs, codes = [], [], [], []
            nd = nodes[e0:e1]
            if self.own_patches is not None and not (self.own_patches[self.itf[ii]["sA"]] or self.own_patches[self.itf[ii]["sB"]]):
                per_itf.append(None)                     # sharded: neither side is an own patch
                continue
            for xb, nm in enumerate(names):
                conn = pen[nm][e0:e1].astype(np.int64)
                r = np.repeat(nd, 16, axis=1)                     # [ne, 512]
                c = np.tile(conn, (1, 32))
                keep = pcol[self.scalar_patch[c]] >= 0
                if self.own_patches is not None:
                    keep &= self.own_patches[self.scalar_patch[r]]
                if not keep.any():
                    continue
                code = np.broadcast_to((la16 | (xb << 5) | (lb32 << 8))[None, :], (ne, 512))
                ev = np.broadcast_to(np.arange(e0, e1, dtype=np.int32)[:, None], (ne, 512))
                rs.append(r[keep]); cs.append(c[keep]); evs.append(ev[keep]); codes.append(code[keep])
            if not rs:
                per_itf.append(None)
                continue
            r = np.concatenate(rs); c = np.concatenate(cs); ev = np.concatenate(evs); code = np.concatenate(codes)
            key = r * n_s + c
            order = np.argsort(key, kind="stable")
            uk, start = np.unique(key[order], return_index=True)
            per_itf.append(dict(uk=uk, ptr=np.append(start, len(key)).astype(np.int64),
                                item_eval=ev[order].astype(np.int32), item_code=code[order].astype(np.int32)))
        out = dict(field=field)
        live = [i for i, d in enumerate(per_itf) if d is not None]
        if not live:
            out.update(n_dest=0, rounds=[])
            return out
        # rounds: greedy colouring of the "destination sets intersect" graph (only interfaces sharing a patch can)
        color = {}
        for i in live:
            pi = set((self.itf[i]["sA"], self.itf[i]["sB"]))
            used = set()
            for j in live:
                if j >= i:
                    break
                if pi & set((self.itf[j]["sA"], self.itf[j]["sB"])) and color[j] not in used:
                    if np.intersect1d(per_itf[i]["uk"], per_itf[j]["uk"], assume_unique=True).size:
                        used.add(color[j])
            k = 0
            while k in used:
                k += 1
            color[i] = k
        # CSR of the penalty part over ALL destinations: rows = dofs, cols = pcol + local cp
        uk_all = np.unique(np.concatenate([per_itf[i]["uk"] for i in live]))
        ur, uc = uk_all // n_s, uk_all % n_s
        pr, pc = self.scalar_patch[ur], self.scalar_patch[uc]
        col = pcol[pc] + (uc - cpo[pc])
        rows = np.stack([dof[pr] + i * ncp[pr] + (ur - cpo[pr]) for i in range(3)], axis=1)     # [n,3]
        rr = rows.ravel(); cc = np.repeat(col, 3)
        o2 = np.lexsort((cc, rr))
        indptr = np.zeros(self.N + 1, dtype=np.int64)
        np.cumsum(np.bincount(rr, minlength=self.N), out=indptr[1:])
        out["indptr"] = indptr; out["indices"] = cc[o2].astype(np.int32)
        pos_all = np.empty(len(rr), dtype=np.int64); pos_all[o2] = np.arange(len(rr))
        pos_all = pos_all.reshape(-1, 3)
        pos_all = np.where(self.bc_mask[rows] != 0, -1, pos_all)     # apply_row_bcs, diag = 0
        out["nnz"] = len(rr)
        out["n_dest"] = len(uk_all)
        rounds = []
        for k in range(max(color.values()) + 1):
            members = [i for i in live if color[i] == k]
            ptrs, evs, codes, uks, off = [], [], [], [], 0
            for i in members:
                d = per_itf[i]
                ptrs.append(d["ptr"][:-1] + off); off += int(d["ptr"][-1])
                evs.append(d["item_eval"]); codes.append(d["item_code"]); uks.append(d["uk"])
            uk = np.concatenate(uks)
            idx = np.searchsorted(uk_all, uk)
            rounds.append(dict(ptr=np.append(np.concatenate(ptrs), off).astype(np.int64),
                               item_eval=np.concatenate(evs), item_code=np.concatenate(codes),
                               pos=np.ascontiguousarray(pos_all[idx]), n_dest=len(uk),
                               dest_patch=self.scalar_patch[uk // n_s].astype(np.int32)))
        out["rounds"] = rounds
        return out

    # -------------------------------------------------------------- const loads
    def _const_force(self):
        """Loads independent of u and of the design: PointSource
        (nonmatching_opt.py:735-738) and dead edge tractions on spline.ds
        (demos_csdl_alpha/thickness_opt/plate_const_th_opt_wint.py:139-150)."""
        f = np.zeros(self.N)
        for pl in self.problem.get("point_loads", []):
            P = self.patches[pl["patch"]]
            conn, D = bsp.surface_point_tables(P.ku, P.kv, 3, 3, np.ones(P.ncp), np.asarray(pl["xi"], float).reshape(1, 2))
            np.add.at(f, P.dof_off + pl["field"] * P.ncp + conn[0], pl["value"] * D[0, 0])
        for el in self.problem.get("edge_loads", []):
            P = self.patches[el["patch"]]
            d, side = el["direction"], el["side"]
            kt = P.kv if d == 0 else P.ku
            kf = P.ku if d == 0 else P.kv
            spans = bsp.nonempty_spans(kt, 3)
            g, w = quad.gauss01((P.quad_deg + 2) // 2)
            h = kt[spans + 1] - kt[spans]
            tau = (kt[spans][:, None] + h[:, None] * g[None, :]).ravel()
            wt = (h[:, None] * w[None, :]).ravel()
            fixed = np.full_like(tau, kf[-1] if side == 1 else kf[0])
            xi = np.stack([fixed, tau], 1) if d == 0 else np.stack([tau, fixed], 1)
            conn, D = bsp.surface_point_tables(P.ku, P.kv, 3, 3, P.cp[:, 3], xi)
            gt = np.einsum("qa,qac->qc", D[:, 2 if d == 0 else 1], P.cp[:, :3][conn])
            jac = np.sqrt((gt * gt).sum(1))
            for c in range(3):
                np.add.at(f, P.dof_off + c * P.ncp + conn, -(wt * jac)[:, None] * D[:, 0] * el["traction"][c])
        self.f_const = f
