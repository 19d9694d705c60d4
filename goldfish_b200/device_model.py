"""Device-resident model: uploads the symbolic phase to HBM and drives the
CUDA kernels through the C ABI (include/goldfish_b200.h).

PyTorch is used for device memory and streams only (torch.empty / .data_ptr());
every number is produced by the hand-written kernels of goldfish_b200/csrc.

HBM layout (all FP64 unless noted):
  cp     [n_scalar][4]   homogeneous control points, one 32-byte record per CP
  u      [N]             displacement, per patch field-blocked [ux | uy | uz]
  theta  [n_th]          thickness dofs
  K      CSR (int64 indptr, int32 indices, f64 vals), pattern fixed by Symbolic
  P[f]   CSR shell part of dR/dCP_f  + penP[f] small CSR (penalty part)
  T      CSR dR/dthickness
  tab_u/tab_v [spans][nq][3][4]  1-D basis tables at the quadrature points
"""
import ctypes as C
import numpy as np
import torch

from . import _capi as capi
from .symbolic import Symbolic


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else C.c_void_p(0)


_NCCL_COMM = {}


def _nccl_comm(lib, dist, device):
    """The library-side NCCL communicator of this process (one per device), see csrc/gf_dist.cu."""
    key = str(device)
    if key not in _NCCL_COMM:
        idt = torch.zeros(128, dtype=torch.uint8)
        if dist.get_rank() == 0:
            buf = (C.c_ubyte * 128)()
            capi.check(lib.gf_dist_unique_id(buf), "gf_dist_unique_id")
            idt = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
        idt = idt.to(device)
        dist.broadcast(idt, src=0)
        raw = bytes(idt.cpu().numpy().tobytes())
        d = capi.GfDist()
        capi.check(lib.gf_dist_init(C.byref(d), C.c_char_p(raw), dist.get_rank(), dist.get_world_size()), "gf_dist_init")
        _NCCL_COMM[key] = d.comm
    return _NCCL_COMM[key]


class DeviceCsr:
    """CSR matrix whose arrays live in HBM."""

    def __init__(self, nrows, ncols, indptr, indices, device, share=None):
        self.nrows, self.ncols = int(nrows), int(ncols)
        self.indptr_h = np.ascontiguousarray(indptr, dtype=np.int64)
        self.indices_h = np.ascontiguousarray(indices, dtype=np.int32)
        self.nnz = int(self.indptr_h[-1])
        if share is not None:             # same pattern as another matrix: one copy of the index arrays in HBM
            self.indptr, self.indices = share.indptr, share.indices
        else:
            self.indptr = torch.from_numpy(self.indptr_h).to(device)
            self.indices = torch.from_numpy(self.indices_h).to(device)
        self.vals = torch.zeros(max(self.nnz, 1), dtype=torch.float64, device=device)
        self._t = None
        self._twin = share
        self.device = device

    def c_struct(self):
        s = capi.GfCsr()
        s.nrows, s.ncols, s.nnz = self.nrows, self.ncols, self.nnz
        s.indptr, s.indices, s.vals = _ptr(self.indptr), _ptr(self.indices), _ptr(self.vals)
        return s

    def transpose_map(self):
        """Host-built transpose structure for the deterministic A^T x gather."""
        if self._t is None and self._twin is not None:          # same pattern: same transpose structure
            self._t = self._twin.transpose_map()
        if self._t is None:
            if self.indices_h is None:
                self.indices_h = self.indices.cpu().numpy()
            rows = np.repeat(np.arange(self.nrows, dtype=np.int64), np.diff(self.indptr_h))
            order = np.argsort(self.indices_h, kind="stable")
            tptr = np.zeros(self.ncols + 1, dtype=np.int64)
            np.cumsum(np.bincount(self.indices_h, minlength=self.ncols), out=tptr[1:])
            self._t_arrays = (torch.from_numpy(tptr).to(self.device),
                              torch.from_numpy(rows[order].astype(np.int32)).to(self.device),
                              torch.from_numpy(order.astype(np.int64)).to(self.device))
            t = capi.GfCsrT()
            t.nrows, t.nnz = self.ncols, self.nnz
            t.indptr, t.indices, t.perm = [_ptr(a) for a in self._t_arrays]
            self._t = t
        return self._t

    def drop_host_copy(self):
        """Large runs: keep the column indices in HBM only (downloaded again if a host routine asks)."""
        self.indices_h = None

    def to_scipy(self):
        import scipy.sparse as sp
        if self.indices_h is None:
            self.indices_h = self.indices.cpu().numpy()
        return sp.csr_matrix((self.vals[:self.nnz].cpu().numpy(), self.indices_h, self.indptr_h),
                             shape=(self.nrows, self.ncols))


class DeviceModel:
    def __init__(self, problem, opt_field=(), shopt_surf_inds=(), device=None, symbolic=None,
                 precond="schwarz", schwarz_layers=2, coarse_nc="auto", schwarz_sub="auto", distributed=None, lean=False):
        if not torch.cuda.is_available():
            raise capi.GoldfishError("goldfish_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = capi.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        # ---- patch-sharded multi-GPU mode (one process per GPU, torch.distributed / NCCL) ----
        import torch.distributed as tdist
        if distributed is None:
            distributed = tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1
        self.dist = tdist if distributed else None
        self.rank = tdist.get_rank() if distributed else 0
        self.world = tdist.get_world_size() if distributed else 1
        from .partition import lpt_partition, shard_symbolic
        nel = [(len(np.unique(P["knots"][0])) - 1) * (len(np.unique(P["knots"][1])) - 1) for P in problem["patches"]]
        if self.world > len(nel):
            raise ValueError("patch-sharded run with more ranks (%d) than spline patches (%d): every rank must own a patch"
                             % (self.world, len(nel)))
        self.owner = lpt_partition(nel, self.world)
        if symbolic is not None:
            self.sym = symbolic
        else:
            self.sym = Symbolic(problem, opt_field, shopt_surf_inds,
                                own_patches=(self.owner == self.rank) if distributed else None)
        S = self.sym
        dv = self.device
        self._keep = []
        self.own_patches = [P.index for P in S.patches if self.owner[P.index] == self.rank]
        self._shard = shard_symbolic(S, self.owner, self.rank)
        self.own_ranges = self._shard["own_ranges"]
        self._color_elem, self._color_ptr = self._shard["color_elem"], self._shard["color_ptr"]

        def up(a, dtype=None):
            a = np.ascontiguousarray(a if dtype is None else np.asarray(a).astype(dtype))
            t = torch.from_numpy(a).to(dv)
            self._keep.append(t)
            return t

        # patch descriptors
        descs = (capi.GfPatchDesc * len(S.patches))()
        for d, P in zip(descs, S.patches):
            d.n_u, d.n_v, d.neu, d.nev = P.n_u, P.n_v, P.neu, P.nev
            d.cp_off, d.dof_off = P.cp_off, P.dof_off
            d.th_off, d.th_kind, d.nth = P.th_off, P.th_kind, P.nth
            d.span_u_off, d.span_v_off, d.cpd_u_off, d.cpd_v_off = P.span_u_off, P.span_v_off, P.cpd_u_off, P.cpd_v_off
            d.rational = P.rational
            d.el_off = P.el_off
            for f in range(3):
                d.pcol_off[f] = P.pcol_off[f]
                d.f[f] = float(P.f[f])
            d.E, d.nu = P.E, P.nu
        raw = np.frombuffer(bytes(descs), dtype=np.uint8).copy()
        self.t_patches = up(raw)
        self.t = {}
        for k, a, dt in (("elem_patch", S.elem_patch, np.int32), ("elem_eu", S.elem_eu, np.int32),
                         ("elem_ev", S.elem_ev, np.int32), ("color_elem", self._color_elem if len(self._color_elem) else np.zeros(1, np.int32), np.int32),
                         ("tab_u", S.tab_u, np.float64), ("tab_v", S.tab_v, np.float64),
                         ("first_cp_u", S.first_cp_u, np.int32), ("first_cp_v", S.first_cp_v, np.int32),
                         ("span_h_u", S.span_h_u, np.float64), ("span_h_v", S.span_h_v, np.float64),
                         ("qw", S.qw, np.float64), ("tw_lin", S.tw_lin, np.float64),
                         ("bc", S.bc_mask, np.uint8), ("bc_list", S.bc_list, np.int32),
                         ("row_nlow", S.row_nlow, np.int32),
                         ("f_const", S.f_const if self.rank == 0 else np.zeros_like(S.f_const), np.float64)):
            self.t[k] = up(a, dt)
        for k, a in S.dirs.items():
            self.t[k] = up(a, np.int32)
        self.color_ptr_h = np.ascontiguousarray(self._color_ptr, dtype=np.int32)
        own_rows = np.zeros(S.N, dtype=bool)
        for b0, b1 in self.own_ranges:
            own_rows[b0:b1] = True
        self.t["bc_list_own"] = up(S.bc_list[own_rows[S.bc_list]] if len(S.bc_list) else np.zeros(0, np.int32), np.int32)
        self.n_bc_own = int(own_rows[S.bc_list].sum()) if len(S.bc_list) else 0
        # state
        self.cp = up(S.cp0, np.float64)
        self.u = torch.zeros(S.N, dtype=torch.float64, device=dv)
        self.theta = up(S.theta0, np.float64)
        # operators
        self.K = DeviceCsr(S.N, S.N, S.K_indptr, S.K_indices, dv)
        self.P = []
        for i in range(len(S.opt_field)):
            twin = next((self.P[j] for j in range(i) if S.P_indices[j] is S.P_indices[i]), None)
            self.P.append(DeviceCsr(S.N, S.P_ncols[i], S.P_indptr[i], S.P_indices[i], dv, share=twin))
        self.T = DeviceCsr(S.N, S.n_th, S.T_indptr, S.T_indices, dv)
        self.penP = []
        # outputs
        self.R = torch.zeros(S.N, dtype=torch.float64, device=dv)
        self.WV = torch.zeros(max(S.num_elements, 1) * 2, dtype=torch.float64, device=dv)
        self.wv_sum = torch.zeros(2, dtype=torch.float64, device=dv)
        self.dWdu = torch.zeros(S.N, dtype=torch.float64, device=dv)
        self.dWdP = [torch.zeros(max(n, 1), dtype=torch.float64, device=dv) for n in S.P_ncols]
        self.dVdP = [torch.zeros(max(n, 1), dtype=torch.float64, device=dv) for n in S.P_ncols]
        self.dWdt = torch.zeros(max(S.n_th, 1), dtype=torch.float64, device=dv)
        self.dVdt = torch.zeros(max(S.n_th, 1), dtype=torch.float64, device=dv)
        self.dt_el = torch.zeros(max(S.num_elements, 1) * 2, dtype=torch.float64, device=dv)
        self._build_penalty(up)
        self._build_struct()
        # PCG workspace
        n = S.N
        self.w_r, self.w_z, self.w_p, self.w_Ap, self.w_dinv = [torch.zeros(n, dtype=torch.float64, device=dv) for _ in range(5)]
        self.w_scal = torch.zeros(16, dtype=torch.float64, device=dv)
        self.w_partial = torch.zeros(4096 * 4, dtype=torch.float64, device=dv)
        self.w_scal_h = torch.zeros(16, dtype=torch.float64).pin_memory()
        w = capi.GfPcgWork()
        w.r, w.z, w.p, w.Ap, w.dinv = [_ptr(t) for t in (self.w_r, self.w_z, self.w_p, self.w_Ap, self.w_dinv)]
        w.scal, w.partial, w.scal_h = _ptr(self.w_scal), _ptr(self.w_partial), C.c_void_p(self.w_scal_h.data_ptr())
        # node-wise tangent product: control points of the owned patches (all of them on one GPU)
        own_p = [P for P in S.patches if self.owner[P.index] == self.rank]
        row0 = np.concatenate([P.dof_off + np.arange(P.ncp, dtype=np.int64) for P in own_p]) if own_p else np.zeros(0, np.int64)
        stride = np.concatenate([np.full(P.ncp, P.ncp, dtype=np.int32) for P in own_p]) if own_p else np.zeros(0, np.int32)
        self._node_rows = (up(row0), up(stride))
        import os as _osn
        if _osn.environ.get("GF_SPMV_ROWWISE", "0") != "1":
            w.nodes.row0, w.nodes.stride, w.nodes.n = _ptr(self._node_rows[0]), _ptr(self._node_rows[1]), len(row0)
        self.pcg_work = w
        if precond not in ("schwarz", "jacobi"):
            raise ValueError("Undefined preconditioner: {}".format(precond))
        self.precond = precond
        self.schwarz_layers = schwarz_layers
        self.schwarz_sub = schwarz_sub
        if schwarz_sub == "auto":
            # one CTA sweeps one sub-domain, and the sweep is a latency chain whose length grows with the block: keep
            # at least ~one wave (148 SMs) of blocks per GPU by shrinking the sub-domains as the ranks multiply
            from .schwarz import SchwarzSetup as _SS
            self.schwarz_sub = _SS.choose_subdomains(S.patches, self.world, schwarz_layers)
        import os as _os0
        if _os0.environ.get("GF_SW_SUB"):                  # tuning experiments: "48" or "24,96"
            v = [int(x) for x in _os0.environ["GF_SW_SUB"].split(",")]
            self.schwarz_sub = v[0] if len(v) == 1 else (v[0], v[1])
        if _os0.environ.get("GF_SW_LAYERS"):
            self.schwarz_layers = int(_os0.environ["GF_SW_LAYERS"])
        max_ne = max(max(P.neu, P.nev) for P in S.patches)
        self.coarse_ratio = None
        if coarse_nc == "auto":
            # Coarse spline level: every patch coarsened by one ratio r >= 7 in both directions, r grown until the level
            # has <= 24 k dofs on one GPU (its sweeps are one latency chain on a thread-block cluster that must hide
            # behind the fine sweeps: 23 k dofs = 720 block steps = 1.4 ms at C3) or <= 48 k in sharded runs (there the
            # coarse solve is a dense product with a row slab of Kc^-1, so the level does not shrink with the GPU count).
            if max_ne < 16:
                coarse_nc = 0
            else:
                from . import coarse as coarse_mod
                self.coarse_ratio = coarse_mod.coarsening_ratio(problem, 24000 if self.world < 2 else 48000)
                coarse_nc = max(8, int(np.ceil(max_ne / self.coarse_ratio)))
        self.coarse_nc = int(coarse_nc) if precond == "schwarz" else 0
        self.problem = problem
        self._pc = None
        self._sw = None
        self._sw_factored = False
        self._coarse_factored = False
        # small systems: the factorisation costs less than the extra CG iterations a lagged
        # preconditioner needs, so refactor whenever K has been re-assembled
        self.eager_refactor = S.N < 100000
        self._K_version = 0
        self._fact_version = -1
        import os as _os
        # relative recurrence residual; the TRUE residual of these systems (kappa ~ 1e10..1e12) stagnates near
        # 1e-9..1e-10, and every parity test also passes at 1e-10, so 1e-11 keeps a margin without idle iterations
        self.krylov_rtol = float(_os.environ.get("GF_KRYLOV_RTOL", "1e-11"))
        self.krylov_max_it = 200000 if precond == "jacobi" else 20000      # per pass; Schwarz-PCG needs tens to hundreds
        # target for the TRUE relative residual |b - K x| / |b| of every solve (None: trust the recurrence)
        self.true_rtol = float(_os.environ.get("GF_TRUE_RTOL", "1e-8"))
        self.pass_rtol = float(_os.environ.get("GF_PASS_RTOL", "1e-6"))       # recurrence tolerance of the first pass
        self.max_refine = 3
        # small systems (latency-bound, an extra pass costs well under a millisecond) always get two correction passes
        # on the double-double residual: the reference's own fixtures are the worst conditioned ones (C1 plate:
        # kappa ~ 1.5e12, where a 1e-9 true residual still leaves 3e-8 in the adjoint vector)
        self.polish = S.N < 200000
        self.gmres_fallback = True
        self.fallback_used = False
        self.last_true_relres = None
        self.krylov_check_every = 50 if precond == "jacobi" else 5
        self.last_krylov_its = 0
        self.last_relres = 0.0
        self.stats = {"launches": 0}
        self.state_epoch = 0
        self._epochs = {}
        # coarse-level refresh policy (opt-in automation: the iteration-growth trigger has not been measured on an
        # optimisation run yet; refresh_coarse() itself is what an optimiser driver calls every few design updates)
        self.design_epoch = 0
        self._coarse_design_epoch = 0
        self.auto_refresh_coarse = False
        self._its_ref = None
        self._want_coarse_refresh = False
        if lean:
            # >= 10 M dof runs: the index arrays of K / dR/dCP / dR/dt and the coupling lists live in HBM from here on;
            # drop the host copies (GBs per rank) once the transpose structures that are built from them exist
            for M in self.P + [self.T] + [pp[0] for pp in self.penP if pp is not None]:
                M.transpose_map()
            for M in [self.K] + self.P + [self.T] + [pp[0] for pp in self.penP if pp is not None]:
                M.drop_host_copy()
            S.K_indices = None
            S.P_indices = [None for _ in S.P_indices]
            S.T_indices = None
            for k in ("K_item", "K_pos", "K_ptr", "R_item", "R_ptr"):
                S.pen.pop(k, None); self._shard["pen"].pop(k, None)
            for pp in list(getattr(S, "penP", [])) + list(self._shard.get("penP", [])):
                pp.pop("rounds", None); pp.pop("indices", None)
            self._penP_cache_keys = None
            import gc
            gc.collect()

    def ensure(self, **what):
        """assemble() only what is stale w.r.t. the current u / design state."""
        need = {k: True for k, v in what.items() if v and self._epochs.get(k) != self.state_epoch}
        if need:
            self.assemble(**need)
            for k in need:
                self._epochs[k] = self.state_epoch

    def touch(self):
        """Mark u / design variables as changed (invalidates cached linearisations)."""
        self.state_epoch += 1

    # ---------------------------------------------------------------- structs
    def _build_penalty(self, up):
        S = self.sym
        pen = S.pen
        self.pen_t = {}
        q = capi.GfPenalty()
        q.n_eval = pen["n_eval"]
        self._penP_cache = {}
        pen = self._shard["pen"]          # destinations restricted to owned rows (all of them on 1 GPU)
        if pen["n_eval"] > 0:
            for k in ("connA", "connB", "connC0", "connC1", "basA", "basB", "basC0", "basC1", "tpar", "alpha",
                      "dofA", "dofB", "R_ptr", "R_item", "R_row", "K_ptr", "K_item", "K_pos"):
                self.pen_t[k] = up(pen[k] if len(pen[k]) else np.zeros(1, pen[k].dtype))
                setattr(q, k, _ptr(self.pen_t[k]))
            ne = pen["n_eval"]
            self.pen_g = torch.zeros(ne * 18, dtype=torch.float64, device=self.device)
            self.pen_Huu = torch.zeros(ne * 324, dtype=torch.float64, device=self.device)
            self.pen_HuX = torch.zeros(ne * 324, dtype=torch.float64, device=self.device)
            q.g, q.Huu, q.HuX = _ptr(self.pen_g), _ptr(self.pen_Huu), _ptr(self.pen_HuX)
            q.nR, q.nK = pen["nR"], pen["nK"]
            for pp, pp_all in zip(self._shard["penP"], S.penP):
                # the part exists on EVERY rank as soon as the model has intersections (possibly with an empty pattern
                # on a rank that owns none of its destinations): DeviceMat products issue one all-reduce per part
                ip = pp.get("indptr", np.zeros(S.N + 1, dtype=np.int64)); ix = pp.get("indices", np.zeros(0, dtype=np.int32))
                twin = next((q[0] for q, qa in zip(self.penP, S.penP) if q is not None and qa.get("indices") is not None
                             and qa.get("indices") is pp_all.get("indices")), None)
                M = DeviceCsr(S.N, S.P_ncols[S.opt_field.index(pp["field"])], ip, ix, self.device, share=twin)
                # one gather struct per ROUND of interfaces with disjoint destinations (rounds accumulate); a rank that
                # owns none of the destinations keeps the all-zero part, so that every rank holds the same number of
                # parts (one all-reduce per part in DeviceMat products)
                structs = []
                for rd in pp["rounds"]:
                    if rd["n_dest"] == 0:
                        continue
                    s = capi.GfPenaltyP()
                    s.n_dest = rd["n_dest"]
                    key = id(rd["item_eval"])
                    if key not in self._penP_cache:       # the same gather lists serve the three fields
                        self._penP_cache[key] = [up(rd[k] if len(rd[k]) else np.zeros(1, rd[k].dtype)) for k in ("ptr", "item_eval", "item_code", "pos")]
                    arrs = self._penP_cache[key]
                    s.ptr, s.item_eval, s.item_code, s.pos = [_ptr(a) for a in arrs]
                    s.vals = _ptr(M.vals)
                    s.field = pp["field"]
                    structs.append(s)
                self.penP.append((M, structs))
        else:
            self.penP = [None for _ in S.opt_field]
        self.pen_struct = q

    def _build_struct(self):
        S = self.sym
        m = capi.GfModel()
        m.num_patches, m.num_elements, m.nq, m.num_colors = len(S.patches), S.num_elements, S.nq, S.num_colors
        m.N, m.n_scalar, m.n_th = S.N, S.n_scalar, S.n_th
        m.patches = _ptr(self.t_patches)
        for k in ("elem_patch", "elem_eu", "elem_ev", "color_elem", "tab_u", "tab_v", "first_cp_u", "first_cp_v",
                  "span_h_u", "span_h_v", "qw", "tw_lin", "bc", "bc_list", "row_nlow",
                  "cp_lo_u", "cp_hi_u", "el_lo_u", "el_hi_u", "cp_lo_v", "cp_hi_v", "el_lo_v", "el_hi_v"):
            setattr(m, k, _ptr(self.t[k]))
        m.color_ptr_h = self.color_ptr_h.ctypes.data_as(C.c_void_p)
        m.cp, m.u, m.theta = _ptr(self.cp), _ptr(self.u), _ptr(self.theta)
        m.n_bc = len(S.bc_list)
        m.K = self.K.c_struct()
        for f in range(3):
            if f in S.opt_field:
                m.P[f] = self.P[S.opt_field.index(f)].c_struct()
        m.T = self.T.c_struct()
        self.model = m
        m2 = capi.GfModel.from_buffer_copy(m)      # same model, zero-dof list restricted to owned rows
        m2.bc_list, m2.n_bc = _ptr(self.t["bc_list_own"]), self.n_bc_own
        self.model_own_bc = m2
        o = capi.GfShellOut()
        o.R, o.WV, o.dWdu = _ptr(self.R), _ptr(self.WV), _ptr(self.dWdu)
        for f in range(3):
            if f in S.opt_field:
                i = S.opt_field.index(f)
                o.dWdP[f] = self.dWdP[i].data_ptr()
                o.dVdP[f] = self.dVdP[i].data_ptr()
        o.dWdt, o.dVdt = _ptr(self.dWdt), _ptr(self.dVdt)
        o.dt_el = _ptr(self.dt_el)
        self.out = o

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ state
    def set_u(self, u):
        u = torch.as_tensor(u, dtype=torch.float64).reshape(-1)
        assert u.numel() == self.sym.N
        u = u.to(self.device, non_blocking=True)
        if torch.equal(u, self.u):
            return                       # same state (e.g. update_uIGA right after the solve): keep cached operators
        self.u.copy_(u)
        self.touch()

    def set_theta(self, th):
        th = torch.as_tensor(th, dtype=torch.float64).reshape(-1)
        assert th.numel() == self.sym.n_th
        th = th.to(self.device, non_blocking=True)
        if torch.equal(th, self.theta):
            return
        self.theta.copy_(th)
        self.design_epoch += 1
        self.touch()

    def set_cp(self, field, arr, surf_inds=None):
        S = self.sym
        if surf_inds is None:
            surf_inds = range(len(S.patches))
        arr = torch.as_tensor(arr, dtype=torch.float64).reshape(-1)
        cpv = self.cp.view(-1, 4)
        o = 0
        for s in surf_inds:
            P = S.patches[s]
            cpv[P.cp_off:P.cp_off + P.ncp, field].copy_(arr[o:o + P.ncp], non_blocking=True)
            o += P.ncp
        assert o == arr.numel()
        self.design_epoch += 1
        self.touch()

    def get_cp(self, field, surf_inds=None):
        S = self.sym
        if surf_inds is None:
            surf_inds = range(len(S.patches))
        cpv = self.cp.view(-1, 4)
        return torch.cat([cpv[S.patches[s].cp_off:S.patches[s].cp_off + S.patches[s].ncp, field] for s in surf_inds])

    # --------------------------------------------------------------- assembly
    def assemble(self, residual=False, tangent=False, functionals=False, shape=False, thickness=False):
        """One pass over shells + coupling.  Results stay in HBM:
        R (BC rows zeroed), K (BCs applied), P[f]/penP[f] (BC rows zeroed), T."""
        lib, st, S = self.lib, self._stream(), self.sym
        what = 0
        if residual:
            what |= capi.GF_OUT_R; self.R.copy_(self.t["f_const"])
        if tangent:
            what |= capi.GF_OUT_K; self.K.vals.zero_()
        if functionals:
            what |= capi.GF_OUT_W
        if shape and S.opt_field:
            what |= capi.GF_OUT_P
            for M in self.P:
                M.vals.zero_()
            for a in self.dWdP + self.dVdP:
                a.zero_()
        if thickness:
            what |= capi.GF_OUT_T
            self.T.vals.zero_(); self.dWdt.zero_(); self.dVdt.zero_(); self.dWdu.zero_()
        if what:
            capi.check(lib.gf_shell_assemble(C.byref(self.model), what, C.byref(self.out), st), "gf_shell_assemble")
        if S.pen["n_eval"] > 0 and (residual or tangent or (shape and S.opt_field)):
            with_X = 1 if (shape and S.opt_field) else 0
            capi.check(lib.gf_penalty_points(C.byref(self.model), C.byref(self.pen_struct), with_X, st), "gf_penalty_points")
            if residual:
                capi.check(lib.gf_penalty_gather_R(C.byref(self.model), C.byref(self.pen_struct), _ptr(self.R), st), "gather_R")
            if tangent:
                capi.check(lib.gf_penalty_gather_K(C.byref(self.model), C.byref(self.pen_struct), st), "gather_K")
            if with_X:
                for pp in self.penP:
                    if pp is not None:
                        pp[0].vals.zero_()
                        for rd in pp[1]:
                            capi.check(lib.gf_penalty_gather_P(C.byref(self.pen_struct), C.byref(rd), st), "gather_P")
        if residual:
            self.allreduce(self.R)
            capi.check(lib.gf_mask_vec(C.byref(self.model), _ptr(self.R), st), "gf_mask_vec")
        if tangent:
            capi.check(lib.gf_bc_set_diag(C.byref(self.model_own_bc), 1.0, st), "gf_bc_set_diag")
            self._K_full = self.dist is None
            self._K_version += 1
        if functionals:
            capi.check(lib.gf_reduce_wv(S.num_elements, _ptr(self.WV), _ptr(self.wv_sum), st), "gf_reduce_wv")
            self.allreduce(self.wv_sum)
        if self.dist is not None:
            if shape and S.opt_field:
                for a in self.dWdP + self.dVdP:
                    self.allreduce(a)
            if thickness:
                for a in (self.dWdt, self.dVdt, self.dWdu):
                    self.allreduce(a)

    def allreduce(self, t):
        """Sum a device tensor over the ranks (no-op in single-GPU runs)."""
        if self.dist is not None:
            self.dist.all_reduce(t)
        return t

    def replicate_K(self):
        """Distributed runs keep only the owned rows of K after assembly; the
        preconditioner set-up needs the rows of the overlap nodes too, so the
        values are summed over the ranks once per factorisation (NVLink all-reduce)."""
        if self.dist is not None and not getattr(self, "_K_full", False):
            self.dist.all_reduce(self.K.vals)
            self._K_full = True

    # ------------------------------------------------------------ linear algebra
    def spmv(self, A, x, y, alpha=1.0, beta=0.0, transpose=False):
        st = self._stream()
        cs = A.c_struct()
        if transpose:
            capi.check(self.lib.gf_spmv_t(C.byref(cs), C.byref(A.transpose_map()), _ptr(x), _ptr(y), alpha, beta, st), "gf_spmv_t")
        else:
            capi.check(self.lib.gf_spmv(C.byref(cs), _ptr(x), _ptr(y), alpha, beta, st), "gf_spmv")
        return y

    def spmv_node(self, x, y, alpha=1.0, beta=0.0):
        """y = beta y + alpha K x with the node-wise kernel the Krylov solvers use (gf_spmv_node: the three field
        rows of a control point share one index / x read; bitwise the same y as gf_spmv,
        scripts/gpu_spmv_node_check.py).  Rows of the owned patches only in sharded runs."""
        r0, st = self._node_rows
        capi.check(self.lib.gf_spmv_node(C.byref(self.K.c_struct()), _ptr(r0), _ptr(st), r0.numel(), _ptr(x), _ptr(y),
                                         alpha, beta, self._stream()), "gf_spmv_node")
        return y

    def spmv_global(self, A, x, y, alpha=1.0, beta=0.0, transpose=False):
        """y = beta y + alpha A x (or A^T x) with x, y replicated on every rank."""
        if self.dist is None:
            return self.spmv(A, x, y, alpha, beta, transpose)
        tmp = torch.zeros_like(y)
        if A is self.K:
            # K may or may not be replicated at this point: use the owned rows only (K symmetric)
            cs = A.c_struct()
            for b0, b1 in self.own_ranges:
                sub = capi.GfCsr.from_buffer_copy(cs)
                sub.indptr = C.c_void_p(A.indptr.data_ptr() + 8 * int(b0)); sub.nrows = int(b1 - b0)
                capi.check(self.lib.gf_spmv(C.byref(sub), _ptr(x), C.c_void_p(tmp.data_ptr() + 8 * int(b0)), 1.0, 0.0,
                                            self._stream()), "gf_spmv")
        else:
            self.spmv(A, x, tmp, 1.0, 0.0, transpose)      # rows of other ranks hold zeros
        self.allreduce(tmp)
        return self.axpby(alpha, tmp, beta, y)

    def axpby(self, a, x, b, y):
        capi.check(self.lib.gf_axpby(x.numel(), a, _ptr(x), b, _ptr(y), self._stream()), "gf_axpby")
        return y

    def dot(self, x, y):
        capi.check(self.lib.gf_dot(x.numel(), _ptr(x), _ptr(y), _ptr(self.w_partial), _ptr(self.w_scal[8:]),
                                   self._stream()), "gf_dot")
        return float(self.w_scal[8].item())

    def _precond_struct(self):
        """GfPrecond: fine overlapping blocks (+ coarse spline level)."""
        if self._pc is None:
            pc = capi.GfPrecond()
            pc.fine = C.pointer(self._schwarz())
            pc.dist = C.pointer(self._dist_struct())
            self._coarse = None
            if self.coarse_nc > 0:
                from . import coarse as coarse_mod
                cpr, P = coarse_mod.build(self.problem, nc=self.coarse_nc, ratio=self.coarse_ratio)
                cpr["alpha_override"] = self.sym.itf_alpha
                self._coarse_problem = cpr
                cm = DeviceModel(cpr, device=self.device, precond="schwarz", coarse_nc=0, distributed=False)
                cm._single_block = True
                Pd = DeviceCsr(P.shape[0], P.shape[1], P.indptr, P.indices, self.device)
                Pd.vals.copy_(torch.from_numpy(P.data))
                Rt = P.T.tocsr(); Rt.sort_indices()
                Rd = DeviceCsr(Rt.shape[0], Rt.shape[1], Rt.indptr, Rt.indices, self.device)
                Rd.vals.copy_(torch.from_numpy(Rt.data))
                rc = torch.zeros(P.shape[1], dtype=torch.float64, device=self.device)
                zc = torch.zeros(P.shape[1], dtype=torch.float64, device=self.device)
                if self.dist is None:
                    pc.coarse = C.pointer(cm._schwarz())
                pc.P, pc.Rt = Pd.c_struct(), Rd.c_struct()
                pc.rc, pc.zc = _ptr(rc), _ptr(zc)
                pc.bc_c, pc.n_bc_c = _ptr(cm.t["bc_list"]), len(cm.sym.bc_list)
                self._coarse = (cm, Pd, Rd, rc, zc)
            self._pc = pc
        return self._pc

    def _schwarz(self):
        """Build (once) the overlapping-Schwarz block structure and its HBM storage."""
        if self._sw is None:
            from .schwarz import SchwarzSetup, NB
            SW = SchwarzSetup(self.sym, layers=self.schwarz_layers, sub=self.schwarz_sub,
                              single_block=getattr(self, "_single_block", False),
                              own_patches=(self.owner == self.rank) if self.dist is not None else None)
            A = SW.arrays()
            dv = self.device
            t = {k: torch.from_numpy(np.ascontiguousarray(A[k])).to(dv)
                 for k in ("n_pad", "nbr", "off_j", "mbj", "rlen", "off_col", "off_y", "off_inv", "glob", "gs", "ls", "off_g", "zptr", "zsrc")}
            t["band"] = torch.zeros(A["band_len"], dtype=torch.float64, device=dv)
            t["band32"] = torch.zeros(A["band_len"], dtype=torch.float32, device=dv)
            t["invd"] = torch.zeros(A["inv_len"], dtype=torch.float64, device=dv)
            t["y"] = torch.zeros(A["n_y"], dtype=torch.float64, device=dv)
            t["s"] = torch.zeros(A["n_y"], dtype=torch.float64, device=dv)
            t["barrier"] = torch.zeros(A["nblocks"], dtype=torch.int32, device=dv)
            t["flag"] = torch.zeros(1, dtype=torch.int32, device=dv)
            step_mb = np.ascontiguousarray(A["step_mb"], dtype=np.int32)
            s = capi.GfSchwarz()
            s.nblocks, s.nb = A["nblocks"], NB
            s.max_nbr, s.max_mb, s.max_n_pad = A["max_nbr"], A["max_mb"], A["max_n_pad"]
            s.debug_flags = 0
            s.n_y, s.band_len = A["n_y"], A["band_len"]
            for k in ("n_pad", "nbr", "off_j", "mbj", "rlen", "off_col", "off_y", "off_inv", "glob", "gs", "ls", "off_g", "zptr", "zsrc",
                      "band", "band32", "invd", "y", "s", "barrier", "flag"):
                setattr(s, k, _ptr(t[k]))
            s.step_mb_h = step_mb.ctypes.data_as(C.c_void_p)
            self._sw = (s, t, step_mb, A)
        return self._sw[0]

    def set_sweep_mode(self, mode):
        """Triangular-sweep kernel of the fine Schwarz blocks: "auto" (by block count), "single" (one CTA
        per block, vector in shared memory; coarse block on a thread-block cluster), "single_nocluster" (coarse
        block as a barrier group inside the fine launch) or "group" (CTA group per block, global-memory barrier)."""
        self._schwarz().debug_flags = {"auto": 0, "single": 4, "group": 8, "single_nocluster": 4 | 16}[mode]

    def _dist_struct(self):
        """GfDist: owned row ranges + the library's own NCCL communicator (created once per process from a unique
        id broadcast over torch.distributed; shared by every model of the process)."""
        if getattr(self, "_dist_c", None) is None:
            d = capi.GfDist()
            if self.dist is not None:
                self._ranges_h = np.ascontiguousarray(self.own_ranges if len(self.own_ranges) else np.zeros((1, 2), np.int64))
                d.n_ranges = max(1, len(self.own_ranges))
                d.ranges_h = self._ranges_h.ctypes.data_as(C.c_void_p)
                d.comm, d.rank, d.world = _nccl_comm(self.lib, self.dist, self.device), self.rank, self.world
            self._dist_c = d
        return self._dist_c

    def _gmres(self, b, x, rtol, max_it, restart=None):
        """Right-preconditioned GMRES(restart) with the current preconditioner (fallback of _krylov)."""
        n = self.sym.N
        if restart is None:            # indefinite systems stagnate under short restarts: as long as ~4 GB of basis allow
            restart = int(max(20, min(100, n, 4e9 / (16.0 * n))))
        if getattr(self, "_gm", None) is None or self._gm[2] != restart:
            dv = self.device
            t = dict(V=torch.empty((restart + 1) * n, dtype=torch.float64, device=dv),
                     Z=torch.empty(restart * n, dtype=torch.float64, device=dv), t=torch.empty(n, dtype=torch.float64, device=dv),
                     hdev=torch.zeros(2 * (restart + 2) + 1, dtype=torch.float64, device=dv),
                     partial=torch.zeros(1024 * (restart + 1), dtype=torch.float64, device=dv),
                     h_host=torch.zeros(2 * (restart + 2) + 1, dtype=torch.float64).pin_memory())
            w = capi.GfGmresWork()
            w.V, w.Z, w.t, w.hdev, w.partial = [_ptr(t[k]) for k in ("V", "Z", "t", "hdev", "partial")]
            w.h_host = C.c_void_p(t["h_host"].data_ptr())
            w.nodes = self.pcg_work.nodes
            self._gm = (w, t, restart)
        w, _, restart = self._gm
        pre = C.byref(self._precond_struct()) if self.precond == "schwarz" else None
        its = C.c_int(0); rel = C.c_double(0.0)
        rc = self.lib.gf_gmres(C.byref(self.K.c_struct()), _ptr(b), _ptr(x), C.byref(w), pre, C.byref(self._dist_struct()),
                               rtol, restart, max_it, C.byref(its), C.byref(rel), self._stream())
        capi.check(rc, "gf_gmres")
        return its.value, rel.value

    def _dense_coarse_inverse(self, cm, pc):
        """Sharded runs: this rank's row slab of Kc^-1 (FP64).  One-off set-up per coarse refresh, done with the
        dense Cholesky of the library stack (cuSOLVER through torch) -- plumbing, not the hot path: every Krylov
        iteration then applies the slab with the hand-written k_dense_rows product (GfPrecond.cinv)."""
        nc = cm.sym.N
        Kc = cm.K
        rows = torch.repeat_interleave(torch.arange(nc, device=self.device), torch.from_numpy(np.diff(Kc.indptr_h)).to(self.device))
        D = torch.zeros((nc, nc), dtype=torch.float64, device=self.device)
        D.index_put_((rows, Kc.indices.long()), Kc.vals[:Kc.nnz], accumulate=True)
        del rows
        D = 0.5 * (D + D.T)
        L = torch.linalg.cholesky(D)
        del D
        inv = torch.cholesky_inverse(L)
        del L
        r0 = (nc * self.rank) // self.world; r1 = (nc * (self.rank + 1)) // self.world
        self._cinv = inv[r0:r1].contiguous().clone()
        del inv
        torch.cuda.empty_cache()
        pc.cinv, pc.cinv_row0, pc.cinv_rows = _ptr(self._cinv), r0, r1 - r0

    def refresh_coarse(self):
        """Bring the coarse level to the CURRENT design and re-factor it at the next preconditioner set-up: the coarse
        control net is the least-squares restriction of the current fine control net (the map that built it from the
        initial design), the coarse thickness the patch means of the current thickness dofs."""
        self._coarse_factored = False
        self._its_ref = None
        self._want_coarse_refresh = False
        if getattr(self, "_coarse", None) is None:
            return
        from . import coarse as coarse_mod
        cm = self._coarse[0]
        S = self.sym
        cpc, thc = coarse_mod.restrict_design(self._coarse_problem, self.cp.detach().cpu().numpy().reshape(-1, 4),
                                              self.theta.detach().cpu().numpy(),
                                              [(P.n_u, P.n_v, P.cp_off, P.th_off, P.nth) for P in S.patches])
        cm.cp.copy_(torch.from_numpy(np.ascontiguousarray(cpc)).to(self.device).reshape(cm.cp.shape))
        cm.theta.copy_(torch.from_numpy(thc).to(self.device))
        cm.touch()
        self._coarse_design_epoch = self.design_epoch

    def factor_preconditioner(self, _reference=False):
        """(Re)build the preconditioner from the current K values."""
        st = self._stream()
        if self._want_coarse_refresh:
            self.refresh_coarse()
        self.replicate_K()
        cs = self.K.c_struct()
        if self.precond == "schwarz":
            pc = self._precond_struct()
            rc = self.lib.gf_schwarz_factor(C.byref(self._schwarz()), C.byref(cs), st)
            if rc == capi.GF_ERR_BREAKDOWN and not _reference:
                # The tangent at this state is not positive definite (e.g. past a buckling point), so its
                # blocks have no Cholesky factor.  Precondition with the tangent of the SAME design at u = 0
                # instead (SPD) and let the Krylov fallback (GMRES) deal with the indefinite operator.
                u_keep = self.u.clone()
                self.u.zero_()
                self.assemble(tangent=True)
                self.factor_preconditioner(_reference=True)
                self.u.copy_(u_keep)
                self.assemble(tangent=True)
                self.precond_is_reference = True
                self._fact_version = self._K_version
                return
            capi.check(rc, "gf_schwarz_factor")
            self.precond_is_reference = bool(_reference)
            if self._coarse is not None and not self._coarse_factored:
                # The coarse operator is the reference configuration's tangent (u = 0, initial design) on the
                # coarse spline space: it does not depend on the state, so it is assembled and factored once
                # (refresh_coarse() forces a rebuild).
                cm = self._coarse[0]
                cm.assemble(tangent=True)
                if self.dist is None:
                    capi.check(self.lib.gf_schwarz_factor(C.byref(cm._schwarz()), C.byref(cm.K.c_struct()), st),
                               "gf_schwarz_factor(coarse)")
                else:
                    self._dense_coarse_inverse(cm, pc)
                self._coarse_factored = True
        capi.check(self.lib.gf_jacobi_setup(C.byref(cs), _ptr(self.w_dinv), st), "gf_jacobi_setup")
        self._sw_factored = True
        self._fact_version = self._K_version

    def precond_apply(self, r, z=None):
        """z = M^-1 r with the factored preconditioner (what every CG iteration calls)."""
        if z is None:
            z = torch.empty_like(r)
        if not self._sw_factored:
            self.factor_preconditioner()
        # sharded runs sum the ranks' contributions in the PCG work vector z: run there and copy out
        zz = self.w_z if self.dist is not None else z
        capi.check(self.lib.gf_precond_apply(C.byref(self._precond_struct()), _ptr(r), _ptr(zz), self.sym.N, self._stream()),
                   "gf_precond_apply")
        if zz is not z:
            z.copy_(zz)
        return z

    def solve(self, b, x=None, rtol=None, max_it=None, refactor=None):
        """x = K^{-1} b by preconditioned CG (K symmetric => also K^{-T} b).
        The preconditioner is factored on first use and whenever refactor=True;
        otherwise the last factorisation is reused (lagged preconditioner)."""
        if x is None:
            x = torch.empty_like(b)
        if refactor or not self._sw_factored or (self.eager_refactor and self._fact_version != self._K_version):
            self.factor_preconditioner()
        # CG's recurrence residual drifts from b - K x on these systems (kappa ~ 1e10..1e12: after ~80 iterations the
        # TRUE residual sits orders of magnitude above a 1e-11 recurrence residual), so the solve is a short
        # sequence of passes with residual replacement: each pass runs PCG from x = 0 on the true residual of the
        # previous iterate, to a tolerance relative to THAT residual, and the true residual is measured with one
        # extra product after every pass.  Same total iteration count as one long pass; a true residual below
        # `true_rtol` (the reference's LU solve, utils/opt_utils.py:176, is exact to ~kappa*eps).
        self.last_true_relres = None
        if rtol is not None or self.true_rtol is None:
            its, rel = self._krylov(b, x, self.krylov_rtol if rtol is None else rtol, max_it)
            self.last_krylov_its, self.last_relres = its, rel
            return x
        if getattr(self, "_w_res", None) is None:
            self._w_res, self._w_cor = torch.empty_like(b), torch.empty_like(b)
        bn = self.dot(b, b) ** 0.5
        its, rel = self._krylov(b, x, self.pass_rtol, max_it)
        self.last_krylov_its, self.last_relres = its, rel
        if self.auto_refresh_coarse and getattr(self, "_coarse", None) is not None:
            from .coarse import refresh_due
            if self._its_ref is None:
                self._its_ref = its                   # first solve after the coarse level was (re)built
            elif refresh_due(its, self._its_ref, self.design_epoch != self._coarse_design_epoch):
                self._want_coarse_refresh = True      # acted upon at the next factorisation
        if not bn > 0.0:                  # zero right-hand side: x = 0 exactly
            self.last_true_relres = 0.0
            return x
        for k in range(self.max_refine + 1):
            capi.check(self.lib.gf_residual_dd(C.byref(self.K.c_struct()), C.byref(self._dist_struct()), _ptr(x), _ptr(b),
                                               _ptr(self._w_res), self._stream()), "gf_residual_dd")
            tr = (self.dot(self._w_res, self._w_res) ** 0.5 / bn) if bn > 0 else 0.0
            self.last_true_relres = tr
            if k == self.max_refine or (tr <= self.true_rtol and not (self.polish and k < 2)):
                break
            cor_rtol = min(1e-2 if tr <= self.true_rtol else 1e-1, max(0.3 * self.true_rtol / max(tr, 1e-300), 1e-9))
            try:
                its2, rel2 = self._krylov(self._w_res, self._w_cor, cor_rtol, max_it)
            except capi.GoldfishNotConverged:
                break                      # keep the iterate of the previous pass; its true residual is reported
            self.last_krylov_its += its2
            self.last_relres = rel2 * tr
            self.axpby(1.0, self._w_cor, 1.0, x)
        return x

    def _krylov(self, b, x, rtol, max_it=None):
        """One Krylov solve of K x = b from x = 0: PCG; on breakdown (p.Ap <= 0: the tangent is indefinite, e.g. past
        a buckling point, where the reference's LU still returns a step) the same system is handed to
        right-preconditioned restarted GMRES with the same preconditioner."""
        st = self._stream()
        cs = self.K.c_struct()
        pre = C.byref(self._precond_struct()) if self.precond == "schwarz" else None
        its = C.c_int(0); rel = C.c_double(0.0)
        max_it = self.krylov_max_it if max_it is None else max_it
        rc = self.lib.gf_pcg(C.byref(cs), _ptr(b), _ptr(x), C.byref(self.pcg_work), pre, C.byref(self._dist_struct()),
                             rtol, 0.0, max_it, self.krylov_check_every, C.byref(its), C.byref(rel), st)
        if rc == capi.GF_ERR_BREAKDOWN and self.gmres_fallback:
            n_it = its.value
            its2, rel2 = self._gmres(b, x, rtol, max_it)
            self.fallback_used = True
            return n_it + its2, rel2
        capi.check(rc, "gf_pcg")
        return its.value, rel.value

    def dRdCP_matrix(self, i):
        """dR/dCP of opt-field slot i as one handle (shell + penalty parts)."""
        from .vecmat import DeviceMat
        return DeviceMat(self, [self.P[i]] + ([self.penP[i][0]] if self.penP[i] is not None else []))

    def newton(self, max_it=30, rtol=1e-3, verbose=False, accept_stagnation=False):
        """PENGoLINS solve_nonlinear_nonmatching_problem(iga_dofs=True): Newton
        from u = 0, stop when |R|/|R0| < rtol (disp_imop.py:38-44)."""
        self.u.zero_()
        self.touch()
        self.newton_stagnated = False
        ref = None
        hist, kits, trel = [], [], []
        du = torch.empty_like(self.u)
        rhs = torch.empty_like(self.u)
        for it in range(max_it + 1):
            self.assemble(residual=True, tangent=True, functionals=True)      # W, V ride along for free
            for k in ("residual", "tangent", "functionals"):
                self._epochs[k] = self.state_epoch                                # valid until u changes
            nrm = self.dot(self.R, self.R) ** 0.5
            if it == 0:
                ref = nrm
            rel = nrm / ref if ref > 0 else 0.0
            hist.append(rel)
            if verbose:
                print("newton", it, nrm, rel, self.last_krylov_its)
            if (it > 0 and rel < rtol) or ref == 0.0:
                break
            # FP64 floor of the residual evaluation: |R| is a difference of internal forces ~ |K||u|, so |R|/|R0| cannot
            # fall below ~eps |K||u| / |R0| (7e-7 on the flat-skinned wing box, whose load is tiny against its membrane
            # stiffness).  A tighter rtol than that floor is met by stagnation: no halving over one step below 1e-5.
            if accept_stagnation and it >= 2 and rel < 1e-5 and rel > 0.5 * hist[-2]:
                self.newton_stagnated = True
                break
            if it == max_it:
                self.newton_history, self.newton_krylov_its, self.newton_true_relres = hist, kits, trel
                raise capi.GoldfishNotConverged("Nonlinear solver failed to converge in %d iterations" % max_it)
            self.axpby(-1.0, self.R, 0.0, rhs)
            self.solve(rhs, du, refactor=(it == 0))
            kits.append(self.last_krylov_its); trel.append(self.last_true_relres)
            self.axpby(1.0, du, 1.0, self.u)
            self.touch()
        self.newton_history, self.newton_krylov_its, self.newton_true_relres = hist, kits, trel
        return self.u
