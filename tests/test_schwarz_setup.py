"""CPU: host set-up of the two-level Schwarz preconditioner (goldfish_b200/schwarz.py,
coarse.py): block envelopes contain the block matrices, every dof is covered, the coarse
prolongation reproduces the fine geometry, and -- emulated with scipy block solves -- the
preconditioner makes CG converge in tens of iterations where per-patch blocks stall."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import cases
from goldfish_b200 import problems, coarse
from goldfish_b200.symbolic import Symbolic
from goldfish_b200.schwarz import SchwarzSetup, NB
from oracle.model import OracleModel


def _pcg(K, b, M, rtol=1e-10, maxit=400):
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy(); rz = r @ z; bn = np.linalg.norm(b)
    for it in range(maxit):
        Ap = K @ p; a = rz / (p @ Ap); x += a * p; r -= a * Ap
        if np.linalg.norm(r) < rtol * bn:
            return x, it + 1
        z = M(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return x, maxit


def test_block_envelopes_and_coverage():
    pr = problems.cylinder(n_el=12)
    S = Symbolic(pr)
    K = sp.csr_matrix((np.ones(S.K_indptr[-1]), S.K_indices, S.K_indptr), shape=(S.N, S.N))
    for sub in (48, 8):
        SW = SchwarzSetup(S, layers=2, sub=sub)
        A = SW.arrays()
        covered = np.zeros(S.N, dtype=int)
        for bl in SW.blocks:
            g = bl["glob"]; m = g >= 0
            covered[g[m]] += 1
            loc = np.full(S.N, -1); loc[g[m]] = np.nonzero(m)[0]
            coo = K[g[m]].tocoo()
            lr = np.nonzero(m)[0][coo.row]; lc = loc[coo.col]
            ok = (lc >= 0) & (lc <= lr)
            br, bj = lr[ok] // NB, lc[ok] // NB
            assert np.all(br - bj <= bl["mbj"][bj])               # every entry lies inside the stored panels
            assert np.all(br - bj <= bl["rlen"][br])
            assert bl["n_pad"] % NB == 0 and len(g) == bl["n_pad"]
        assert covered.min() >= 1                                  # additive Schwarz covers every dof
        assert np.array_equal(np.diff(A["zptr"]), covered)         # prolongation gather lists every copy once
        assert A["band_len"] == int(((A["mbj"].astype(np.int64) + 1) * NB * NB).sum())
    one = SchwarzSetup(S, single_block=True)
    assert len(one.blocks) == 1 and one.blocks[0]["n"] == S.N


def test_coarse_prolongation_is_exact_refinement():
    pr = problems.scordelis_lo(num_el=8)               # rational patches
    cpr, P = coarse.build(pr, nc=4)
    S, Sc = Symbolic(pr), Symbolic(cpr)
    assert P.shape == (S.N, Sc.N)
    for Pf, Pc in zip(S.patches, Sc.patches):
        blk = P[Pf.dof_off:Pf.dof_off + Pf.ncp, Pc.dof_off:Pc.dof_off + Pc.ncp]
        assert np.abs(blk @ Pc.cp - Pf.cp).max() < 1e-11            # homogeneous control net refined exactly
        assert np.abs(np.asarray(blk.sum(1)).ravel() - 1).max() < 1e-12   # partition of unity


def _cpu_tangent(pr, alpha=None):
    """K at u = 0 from the compiled CPU port (shells + penalty)."""
    from oracle.cpu_port import CpuModel
    from goldfish_b200 import _capi as capi
    if alpha is not None:
        pr = dict(pr); pr["alpha_override"] = alpha      # coarse level: keep the fine penalty stiffness
    cm = CpuModel(pr)
    cm.set_u(np.zeros(cm.S.N))
    cm.shell(capi.GF_OUT_R | capi.GF_OUT_K)
    return cm, cm.K_matrix().tocsr(), -cm.residual()


def test_two_level_preconditioner_emulated():
    pr = problems.cylinder(n_el=16)
    cm, K, b = _cpu_tangent(pr)
    S = cm.S
    SW = SchwarzSetup(S, layers=2, sub=48)
    lus = []
    for bl in SW.blocks:
        g = bl["glob"][bl["glob"] >= 0]
        lus.append((g, spla.splu(K[g][:, g].tocsc())))
    cpr, P = coarse.build(pr, nc=8)
    cpr["alpha_override"] = S.itf_alpha                    # keep the fine penalty stiffness on the coarse level
    cc, Kc, _ = _cpu_tangent(cpr, alpha=S.itf_alpha)
    luc = spla.splu(Kc.tocsc())
    bc_c = cc.S.bc_list

    def one_level(r):
        z = np.zeros_like(r)
        for g, lu in lus:
            z[g] += lu.solve(r[g])
        return z

    def two_level(r):
        rc = P.T @ r; rc[bc_c] = 0.0
        return one_level(r) + P @ luc.solve(rc)

    xe = spla.splu(K.tocsc()).solve(b)
    x1, it1 = _pcg(K, b, one_level)
    x2, it2 = _pcg(K, b, two_level)
    assert it2 < it1 and it2 < 80
    assert np.linalg.norm(x2 - xe) < 1e-6 * np.linalg.norm(xe)
