"""Coarse spline level of the two-level Schwarz preconditioner (host set-up).

The fine spline space of every patch contains the space on a coarser knot
vector (a subset of its knots), so the prolongation is exact knot insertion:
P = kron(P_v, P_u) per patch and field.  The coarse operator is obtained by
re-discretising the same shells + penalty coupling on the coarse patches with
the same kernels (a second, small DeviceModel), which keeps everything on the
GPU and avoids a sparse triple product.
"""
import numpy as np
import scipy.sparse as sp
from . import bsplines as bsp


def _coarse_knots(kn, p, nc):
    uniq = np.unique(kn)
    ne = len(uniq) - 1
    if ne <= nc:
        return kn.copy(), np.zeros(0)
    keep = np.unique(np.round(np.linspace(0, ne, nc + 1)).astype(int))
    interior_keep = uniq[keep[1:-1]]
    coarse = np.concatenate([np.full(p + 1, uniq[0]), interior_keep, np.full(p + 1, uniq[-1])])
    missing = np.setdiff1d(uniq[1:-1], interior_keep)
    return coarse, missing


def coarsening_ratio(problem, target_dofs, r_min=7.0):
    """Uniform coarsening ratio r (fine elements per coarse element and direction, >= r_min) for which the coarse
    model has about target_dofs dofs: every patch is coarsened isotropically, so long thin patches (spars, ribs)
    keep their aspect ratio instead of getting a fixed number of coarse elements per side."""
    ne = [(len(np.unique(pd["knots"][0])) - 1, len(np.unique(pd["knots"][1])) - 1) for pd in problem["patches"]]
    p = problem["patches"][0]["p"][0]

    def dofs(r):
        return sum(3 * (max(min(a, 8), int(np.ceil(a / r))) + p) * (max(min(b, 8), int(np.ceil(b / r))) + p) for a, b in ne)
    r = float(r_min)
    while dofs(r) > target_dofs and r < 1e4:
        r *= 1.05
    return r


def build(problem, nc=8, ratio=None):
    """Returns (coarse_problem, P) with P the (N x Nc) scipy CSR prolongation.  nc: coarse elements per side of
    every patch; ratio (overrides nc): fine elements per coarse element, per direction and patch."""
    patches_c, blocks, restrict = [], [], []
    for pd in problem["patches"]:
        p = pd["p"][0]
        ku, kv = [np.asarray(k, dtype=np.float64) for k in pd["knots"]]
        n_u, n_v = bsp.num_basis(ku, p), bsp.num_basis(kv, p)
        if ratio is not None:
            neu, nev = len(np.unique(ku)) - 1, len(np.unique(kv)) - 1
            ncu = max(min(neu, 8), int(np.ceil(neu / ratio))); ncv = max(min(nev, 8), int(np.ceil(nev / ratio)))    # at least 8 per side
        else:
            ncu = ncv = nc
        cu, mu = _coarse_knots(ku, p, ncu); cv, mv = _coarse_knots(kv, p, ncv)
        Pu, ku2 = bsp.knot_insertion_operator(cu, p, mu); Pv, kv2 = bsp.knot_insertion_operator(cv, p, mv)
        assert np.allclose(ku2, ku) and np.allclose(kv2, kv)
        X = np.asarray(pd["cp"], dtype=np.float64).reshape(n_v, n_u, 4)
        # coarse control net: least-squares inverse of the refinement
        Xc = np.linalg.lstsq(Pv, X.reshape(n_v, -1), rcond=None)[0].reshape(Pv.shape[1], n_u, 4)
        Xc = np.linalg.lstsq(Pu, Xc.transpose(1, 0, 2).reshape(n_u, -1), rcond=None)[0]
        Xc = Xc.reshape(Pu.shape[1], Pv.shape[1], 4).transpose(1, 0, 2)
        nc_u, nc_v = Pu.shape[1], Pv.shape[1]
        restrict.append((np.linalg.pinv(Pv), np.linalg.pinv(Pu)))             # least-squares restriction of a control net
        P2 = sp.kron(sp.csr_matrix(Pv), sp.csr_matrix(Pu), format="csr")      # (n_v n_u) x (nc_v nc_u)
        ncp, ncpc = n_u * n_v, nc_u * nc_v
        # coarse zero-dofs: a coarse dof is constrained iff its dominant fine dof is
        bc = np.zeros(3 * ncp, dtype=bool); bc[np.asarray(pd.get("bc_dofs", []), dtype=np.int64)] = True
        dom = np.asarray(abs(P2).argmax(axis=0)).ravel()
        bc_c = np.concatenate([f * ncpc + np.nonzero(bc[f * ncp + dom])[0] for f in range(3)])
        th = pd["thickness"]
        tval = float(np.mean(np.atleast_1d(th["values"])))
        patches_c.append(dict(p=pd["p"], knots=(cu, cv), cp=Xc.reshape(-1, 4), bc_dofs=bc_c.astype(np.int64),
                              quad_deg=pd["quad_deg"], thickness=dict(kind="const", values=tval),
                              body_force=(0.0, 0.0, 0.0), E=pd.get("E", problem["E"]), nu=pd.get("nu", problem["nu"])))
        blocks.append(sp.block_diag([P2, P2, P2], format="csr"))
    coarse = dict(name=problem.get("name", "") + "_coarse", patches=patches_c, E=problem["E"], nu=problem["nu"],
                  interfaces=problem.get("interfaces", []), penalty_coefficient=problem.get("penalty_coefficient", 1e3),
                  point_loads=[], edge_loads=[], restrict=restrict)
    P = sp.block_diag(blocks, format="csr")
    P.sort_indices()
    return coarse, P


def restrict_design(coarse, cp_fine, theta_fine, fine_patches):
    """Coarse control net and thickness of the CURRENT design: per patch the least-squares inverse of the knot
    insertion applied to the fine homogeneous control net (the same map `build` applies to the initial design), and
    the patch mean of the thickness dofs.  cp_fine: [n_scalar, 4]; fine_patches: (n_u, n_v, cp_off, th_off, nth) per
    patch.  Returns (cp_coarse [n_scalar_c, 4], theta_coarse [n_patches])."""
    cps, ths = [], []
    for (Rv, Ru), (n_u, n_v, cp_off, th_off, nth) in zip(coarse["restrict"], fine_patches):
        X = np.asarray(cp_fine[cp_off:cp_off + n_u * n_v], dtype=np.float64).reshape(n_v, n_u, 4)
        Xc = np.einsum("av,vuk->auk", Rv, X)
        Xc = np.einsum("bu,auk->abk", Ru, Xc)                                  # [nc_v, nc_u, 4]
        cps.append(Xc.reshape(-1, 4))
        ths.append(float(np.mean(theta_fine[th_off:th_off + nth])))
    return np.concatenate(cps), np.asarray(ths)


def refresh_due(its, its_ref, design_changed, growth=1.6, slack=10):
    """Iteration-growth trigger of the automatic coarse refresh: the coarse level belongs to an older design and the
    first Krylov pass needs clearly more iterations than it did right after the level was built."""
    return bool(design_changed and its_ref is not None and its > growth * its_ref + slack)
