"""Regression goldens of the ORACLE for BASELINE configs C1 (plate) and C2 (T-beam).

The reference holds no golden vectors for this path (its tests assert nothing,
SURVEY.md section 4) and cannot run here, so these are outputs of the CPU
restatement -- "parity unpinned" -- used (a) to keep the oracle from drifting
and (b) to check the CUDA path at the full C1/C2 sizes without re-running the
slow oracle on the GPU box.     python tests/golden/make_oracle_goldens.py
The LU solves (Newton steps and adjoint) are iteratively refined with an extended-precision
residual (OracleModel.solve(refine=3)): the goldens hold the solution of the linear systems
themselves, so the CUDA path is held to north_star's 1e-8 on C1 as well (kappa ~ 1.5e12).
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.model import OracleModel
import cases

HERE = os.path.dirname(os.path.abspath(__file__))


def make(name, pr, kw):
    m = OracleModel(pr)
    u = cases.random_state(m.N, m.bc_global)
    m.set_u(u)
    out = dict(u=u, R=m.residual(), W=m.energy(), V=m.volume(), dWdu=m.dWdu(apply_bcs=False),
               dWdt=m.dWdt(), dVdt=m.dVdt())
    K = m.stiffness(); out.update(K_indptr=K.indptr, K_indices=K.indices, K_data=K.data)
    T = m.dRdt(); out.update(T_indptr=T.indptr, T_indices=T.indices, T_data=T.data)
    for i, f in enumerate(kw.get("opt_field", [])):
        A = m.dRdCP(f, kw["shopt_surf_inds"][i])
        out["P%d_indptr" % f] = A.indptr; out["P%d_indices" % f] = A.indices; out["P%d_data" % f] = A.data
        out["dWdP%d" % f] = m.dWdCP(f, kw["shopt_surf_inds"][i]); out["dVdP%d" % f] = m.dVdCP(f, kw["shopt_surf_inds"][i])
    un = m.solve_nonlinear(max_it=30, rtol=1e-3, refine=3)
    out.update(u_newton=un, newton_hist=np.array(m.newton_history))
    # adjoint total derivative of W_int w.r.t. thickness dofs at the Newton state
    Kn = m.stiffness(); lam = m.solve(Kn, m.dWdu(apply_bcs=True), transpose=True, refine=3)
    out.update(W_newton=m.energy(), lam=lam, dWdt_total=m.dWdt() - m.dRdt().T @ lam)
    np.savez_compressed(os.path.join(HERE, name + "_golden.npz"), **out)
    print(name, m.N, "K nnz", K.nnz, "newton its", len(m.newton_history) - 1, "W", out["W_newton"])


if __name__ == "__main__":
    make("tbeam_c2", *cases.tbeam_c2())
    make("plate_c1", *cases.plate_c1())
