"""Host set-up of the overlapping additive-Schwarz preconditioner (once per topology).

One block per patch: the patch's control points plus ``layers`` graph layers of
the neighbouring patches' control points across each intersection (the scalar
CP graph is the K pattern collapsed over the 3 fields).  Inside a block the
nodes are ordered so that the block matrix is BANDED: own nodes in natural
order (slow index = the longer parametric direction), overlap nodes inserted
at the row of their nearest own node plus their offset along the slow
direction.  3 dofs per node are interleaved.  csrc/gf_schwarz.cu factors and
solves the bands; this file only produces index arrays.

Mirrors PENGoLINS' ``solve_nonmatching_mat(..., solver='ksp')`` (CG +
additive PCFIELDSPLIT with per-patch LU), with overlap.
"""
import numpy as np
import scipy.sparse as sp
from scipy.spatial import cKDTree

NB = 64


class SchwarzSetup:
    def __init__(self, sym, layers=2, single_block=False, sub=(24, 96), own_patches=None):
        """own_patches (sharded runs): boolean array over the patches; only the blocks of those patches are built,
        and the control-point graph only for them and the patches they intersect."""
        S = sym
        self.sym, self.layers = sym, layers
        n_s = S.n_scalar
        need = None
        if own_patches is not None and not single_block:
            own_patches = np.asarray(own_patches, dtype=bool)
            need = own_patches.copy()
            for T in S.itf:
                if own_patches[T["sA"]] or own_patches[T["sB"]]:
                    need[T["sA"]] = need[T["sB"]] = True
        # scalar CP graph: own stencils + coupled pairs
        rows, cols = [], []
        for P in S.patches:
            if need is not None and not need[P.index]:
                continue
            cand, mask, _ = S._own_stencil(P)
            a, m = np.nonzero(mask)
            rows.append(P.cp_off + a); cols.append(P.cp_off + cand[a, m])
        if len(S.cpl_keys):
            rows.append(S.cpl_keys // n_s); cols.append(S.cpl_keys % n_s)
        r = np.concatenate(rows); c = np.concatenate(cols)
        G = sp.csr_matrix((np.ones(len(r), dtype=np.int8), (r, c)), shape=(n_s, n_s))
        G.sum_duplicates()
        self.G = G
        dof = np.array([P.dof_off for P in S.patches]); ncp = np.array([P.ncp for P in S.patches])
        cpo = np.array([P.cp_off for P in S.patches])
        Xall = S.cp0[:, :3] / S.cp0[:, 3:4]
        self.blocks = []
        if single_block:
            # coarse level: ONE block with every node, reverse Cuthill-McKee ordered
            from scipy.sparse.csgraph import reverse_cuthill_mckee
            nodes = np.asarray(reverse_cuthill_mckee(G.astype(np.int32), symmetric_mode=True), dtype=np.int64)
            self.blocks.append(self._finish_block(nodes, G, n_s, dof, ncp, cpo, n_own=n_s))
            return
        self.sub = sub
        for P, (i0, i1, j0, j1) in self._subdomains(S.patches, sub):
            if own_patches is not None and not own_patches[P.index]:
                continue
            # own set = a rectangle of the patch's CP grid (at most sub[0] x sub[1] nodes; 24 x 96 measured best on B200 at
            # 1 M DOF: the narrow side sets the band width, the long side keeps the overlap volume down):
            # short band => short triangular-solve chains and cheap factorisation;
            # the coarse spline level carries the global coupling.
            II, JJ = np.meshgrid(np.arange(i0, i1), np.arange(j0, j1), indexing="xy")
            own = (P.cp_off + II + JJ * P.n_u).ravel()
            cur = own
            for _ in range(layers):
                cur = np.union1d(cur, G[cur].indices)
            extra = np.setdiff1d(cur, own)
            # natural keys of own nodes; slow direction = the one with more CPs
            I = (own - P.cp_off) % P.n_u - i0; J = (own - P.cp_off) // P.n_u - j0
            nru, nrv = i1 - i0, j1 - j0
            swap = nru > nrv
            slow_own, fast_own = (I, J) if swap else (J, I)
            ks, kf = slow_own.astype(np.float64), fast_own.astype(np.float64)
            if len(extra):
                Xo = Xall[own].reshape(nrv, nru, 3)
                tree = cKDTree(Xall[own])
                _, nn = tree.query(Xall[extra])
                In, Jn = I[nn], J[nn]

                def direction(di, dj):
                    a_i = np.clip(In + di, 0, nru - 1); a_j = np.clip(Jn + dj, 0, nrv - 1)
                    b_i = np.clip(In - di, 0, nru - 1); b_j = np.clip(Jn - dj, 0, nrv - 1)
                    d = Xo[a_j, a_i] - Xo[b_j, b_i]
                    steps = (a_i - b_i) + (a_j - b_j)
                    h = np.linalg.norm(d, axis=1) / np.maximum(steps, 1)
                    e = d / np.maximum(np.linalg.norm(d, axis=1), 1e-300)[:, None]
                    return e, np.maximum(h, 1e-300)
                eu, hu = direction(1, 0); ev, hv = direction(0, 1)
                off = Xall[extra] - Xall[own][nn]
                pu = (off * eu).sum(1); pv = (off * ev).sum(1)
                ou = pu / hu; ov = pv / hv
                # overlap nodes of a patch that stands ON this one (T-/X-junction: rib on skin) lie out of the own
                # surface: fold their normal distance into the slow key, so that the sheet is laid down along the
                # slow direction instead of piling all its layers into one row (which would widen the band)
                on = np.sqrt(np.maximum((off * off).sum(1) - pu * pu - pv * pv, 0.0)) / (hu if swap else hv)
                on = np.where(on > 0.25, on, 0.0) * float(__import__("os").environ.get("GF_SW_FOLD", "1"))
                s_e, f_e = (In + ou + on, Jn + ov) if swap else (Jn + ov + on, In + ou)
                ks = np.concatenate([ks, s_e]); kf = np.concatenate([kf, f_e])
            nodes = np.concatenate([own, extra])
            order = np.lexsort((kf, ks))
            best = self._finish_block(nodes[order], G, n_s, dof, ncp, cpo, n_own=len(own))
            if len(extra) and best["mb"] > 10:
                # wide band: overlap nodes that do not continue the own grid (sheets standing on it, several
                # interfaces meeting).  Try a reverse Cuthill-McKee ordering of the block's own graph and keep the
                # ordering with the smaller band storage.
                from scipy.sparse.csgraph import reverse_cuthill_mckee
                sub = G[nodes][:, nodes]
                rcm = np.asarray(reverse_cuthill_mckee(sub.astype(np.int32), symmetric_mode=True), dtype=np.int64)
                alt = self._finish_block(nodes[rcm], G, n_s, dof, ncp, cpo, n_own=len(own))
                if int(alt["mbj"].sum()) < int(best["mbj"].sum()):
                    best = alt
            self.blocks.append(best)
            self.blocks[-1]["patch"] = P.index

    def keep_blocks(self, mask):
        """Distributed runs: each rank factors and solves only the blocks of its own patches."""
        self.blocks = [b for b, m in zip(self.blocks, mask) if m]

    @staticmethod
    def choose_subdomains(patches, world, layers=2, candidates=((24, 96), (24, 48), (24, 24), (12, 24))):
        """Sub-domain shape for `world` GPUs from a two-term cost model of one preconditioner application per GPU:
        streaming the solve-form panels (HBM bound: bytes / 4.8 TB/s, what k_sw_solve1 reaches) against the
        triangular-sweep chain of one block (latency bound: 2 x block rows x 2.7 us per step, times the number of
        waves when the GPU's blocks do not all fit on its 148 SMs).  Small sub-domains shorten the chain and cost
        overlap bytes: one GPU at 1 M dofs is bandwidth bound and keeps 24 x 96 (measured best,
        profiles/r1_sweep_tuning_*.jsonl); with more GPUs the bytes per GPU shrink and the chain takes over
        (profiles/r2_sweep_tuning_subdomains.jsonl: the iteration count does not grow with smaller sub-domains)."""
        ov = 3 * layers                                   # overlap in control-point layers per side
        best, best_t = candidates[0], None
        for a, b in candidates:
            nblk, bytes_tot, rows_max, npad_max = 0, 0.0, 0, 0
            for P in patches:
                su = max(1, int(np.ceil(P.n_u / a))); sv = max(1, int(np.ceil(P.n_v / b)))
                wa = min(P.n_u, int(np.ceil(P.n_u / su)) + 2 * ov); wb = min(P.n_v, int(np.ceil(P.n_v / sv)) + 2 * ov)
                n_pad = 3 * wa * wb
                rows = int(np.ceil(n_pad / NB))
                mb = int(np.ceil(9.0 * min(wa, wb) / NB)) + 1
                nblk += su * sv
                bytes_tot += 0.7 * su * sv * rows * ((mb + 1) * NB * NB * 4 * 2 + NB * NB * 8)
                rows_max = max(rows_max, rows); npad_max = max(npad_max, n_pad)
            slots = 148 * max(1, int(220 * 1024 // (npad_max * 8)))
            waves = int(np.ceil(nblk / world / slots))
            t = max(bytes_tot / world / 4.8e12, 2 * rows_max * 2.7e-6 * waves)
            if best_t is None or t < 0.97 * best_t:
                best, best_t = (a, b), t
        return best

    @staticmethod
    def count_subdomains(patches, sub):
        sub_u, sub_v = (sub, sub) if np.isscalar(sub) else sub
        return int(sum(max(1, int(np.ceil(P.n_u / sub_u))) * max(1, int(np.ceil(P.n_v / sub_v))) for P in patches))

    @staticmethod
    def _subdomains(patches, sub):
        """Rectangles (i0, i1, j0, j1) of at most sub x sub (or sub[0] x sub[1]) control points tiling each patch."""
        sub_u, sub_v = (sub, sub) if np.isscalar(sub) else sub
        for P in patches:
            su = max(1, int(np.ceil(P.n_u / sub_u))); sv = max(1, int(np.ceil(P.n_v / sub_v)))
            eu = np.round(np.linspace(0, P.n_u, su + 1)).astype(int); ev = np.round(np.linspace(0, P.n_v, sv + 1)).astype(int)
            for b in range(sv):
                for a in range(su):
                    yield P, (eu[a], eu[a + 1], ev[b], ev[b + 1])

    def _finish_block(self, nodes, G, n_s, dof, ncp, cpo, n_own):
        """Envelope (variable panel heights) and dof maps of one ordered block."""
        S = self.sym
        pos = np.full(n_s, -1, dtype=np.int64); pos[nodes] = np.arange(len(nodes))
        sub = G[nodes]
        pr = np.repeat(np.arange(len(nodes)), np.diff(sub.indptr)); pc = pos[sub.indices]
        ok = pc >= 0
        beta = int(np.abs(pr[ok] - pc[ok]).max())
        n = 3 * len(nodes)
        bw = 3 * beta + 2
        n_pad = ((n + NB - 1) // NB) * NB
        nbr_ = n_pad // NB
        # block-row envelope: first block column reached by each block row
        # (dof = 3*node + field; fields couple fully, so node extremes suffice)
        first_node = np.full(len(nodes), len(nodes), dtype=np.int64)
        np.minimum.at(first_node, pr[ok], pc[ok])
        row_blk = (3 * np.arange(len(nodes)) + 2) // NB          # block row of the node's last dof
        row_blk0 = (3 * np.arange(len(nodes))) // NB             # ... and of its first dof
        fc = np.arange(nbr_, dtype=np.int64)
        np.minimum.at(fc, row_blk, (3 * first_node) // NB)
        np.minimum.at(fc, row_blk0, (3 * first_node) // NB)
        for b_ in range(nbr_ - 2, -1, -1):                       # monotone envelope
            fc[b_] = min(fc[b_], fc[b_ + 1])
        rlen = np.arange(nbr_) - fc
        last = np.searchsorted(fc, np.arange(nbr_), side="right") - 1   # last row whose envelope reaches col j
        mbj = np.maximum(last - np.arange(nbr_), 0)
        ps = S.scalar_patch[nodes]
        glob = np.full(n_pad, -1, dtype=np.int32)
        for f in range(3):
            glob[f:n:3] = dof[ps] + f * ncp[ps] + (nodes - cpo[ps])
        return dict(nodes=nodes, n=n, n_pad=n_pad, bw=bw, nbr=nbr_, mbj=mbj.astype(np.int32),
                    rlen=rlen.astype(np.int32), mb=int(mbj.max()), glob=glob, n_own=n_own)

    def arrays(self):
        """Flat index arrays for GfSchwarz."""
        S = self.sym
        nb = len(self.blocks)
        n_pad = np.array([b["n_pad"] for b in self.blocks], dtype=np.int32)
        nbr = np.array([b["nbr"] for b in self.blocks], dtype=np.int32)
        mbj = np.concatenate([b["mbj"] for b in self.blocks]).astype(np.int32)
        rlen = np.concatenate([b["rlen"] for b in self.blocks]).astype(np.int32)
        off_j = np.concatenate([[0], np.cumsum(nbr.astype(np.int64))[:-1]]).astype(np.int64)
        col_sz = (mbj.astype(np.int64) + 1) * NB * NB
        off_col = np.concatenate([[0], np.cumsum(col_sz)[:-1]]).astype(np.int64)
        band_len = int(col_sz.sum())
        step_mb = np.zeros(int(nbr.max()), dtype=np.int32)
        for b in self.blocks:
            step_mb[:b["nbr"]] = np.maximum(step_mb[:b["nbr"]], b["mbj"])
        off_y = np.concatenate([[0], np.cumsum(n_pad.astype(np.int64))[:-1]]).astype(np.int64)
        inv_sz = nbr.astype(np.int64) * NB * NB
        off_inv = np.concatenate([[0], np.cumsum(inv_sz)[:-1]]).astype(np.int64)
        glob = np.concatenate([b["glob"] for b in self.blocks])
        # global -> local lookup of each block: sorted global dofs + their local index
        gs, ls, off_g = [], [], [0]
        for b in self.blocks:
            g = b["glob"]; m = np.nonzero(g >= 0)[0]
            o = np.argsort(g[m], kind="stable")
            gs.append(g[m][o]); ls.append(m[o].astype(np.int32)); off_g.append(off_g[-1] + len(m))
        gs = np.concatenate(gs).astype(np.int32); ls = np.concatenate(ls); off_g = np.asarray(off_g, dtype=np.int64)
        # prolongation gather: dof d <- every block-local copy, in block order
        src = np.nonzero(glob >= 0)[0]
        d = glob[src]
        o = np.argsort(d, kind="stable")
        zsrc = src[o].astype(np.int64)
        zptr = np.zeros(S.N + 1, dtype=np.int64)
        np.cumsum(np.bincount(d, minlength=S.N), out=zptr[1:])
        return dict(nblocks=nb, n_pad=n_pad, nbr=nbr, off_j=off_j, mbj=mbj, rlen=rlen, off_col=off_col,
                    step_mb=step_mb, off_y=off_y, off_inv=off_inv,
                    glob=glob.astype(np.int32), gs=gs, ls=ls, off_g=off_g, zptr=zptr, zsrc=zsrc, n_y=int(n_pad.sum()),
                    band_len=band_len, inv_len=int(inv_sz.sum()),
                    max_nbr=int(nbr.max()), max_mb=int(mbj.max()), max_n_pad=int(n_pad.max()))
