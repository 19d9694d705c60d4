"""GPU: the Krylov path past a buckling point.  A strongly compressed T-beam has an indefinite tangent: block
Cholesky of that tangent fails and CG breaks down, while the reference's LU (utils/opt_utils.py:176) still returns a
Newton step.  The library then preconditions with the u = 0 tangent of the same design and solves with
right-preconditioned GMRES; the result is compared with a sparse LU of the very same matrix."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla
import torch

from goldfish_b200 import problems

pytestmark = pytest.mark.gpu


def _compressed_state(dm, strain):
    """u_y = -strain * y on every control point (uniform axial shortening of the beam)."""
    S = dm.sym
    u = np.zeros(S.N)
    for P in S.patches:
        X = S.cp0[P.cp_off:P.cp_off + P.ncp]
        u[P.dof_off + P.ncp:P.dof_off + 2 * P.ncp] = -strain * X[:, 1] / X[:, 3]
    u[S.bc_list] = 0.0
    return u


def test_indefinite_tangent_falls_back_to_gmres(built_lib):
    from goldfish_b200.device_model import DeviceModel
    pr = problems.tbeam(num_el=8, body_force=(0.0, 0.0, 1.0))
    dm = DeviceModel(pr)
    dm.set_u(_compressed_state(dm, 0.002))        # just past buckling: three negative eigenvalues
    dm.assemble(residual=True, tangent=True)
    K = dm.K.to_scipy()

    dense_min = np.linalg.eigvalsh(K.toarray()).min()
    assert dense_min < 0.0                                   # K is indefinite
    b = -dm.R.clone()
    x = dm.solve(b, refactor=True).cpu().numpy()
    assert dm.fallback_used and dm.precond_is_reference
    xe = spla.splu(K.tocsc()).solve(b.cpu().numpy())
    # an indefinite system: the residual is what GMRES controls (and what the LU answer itself is good to)
    assert dm.last_true_relres < 1e-8
    assert np.linalg.norm(K @ x - b.cpu().numpy()) < 1e-8 * np.linalg.norm(b.cpu().numpy())
    assert np.linalg.norm(x - xe) < 1e-5 * np.linalg.norm(xe)


def test_gmres_matches_pcg_on_spd_system(built_lib):
    """GMRES alone (forced) on an SPD tangent gives the CG / LU solution."""
    from goldfish_b200.device_model import DeviceModel
    pr = problems.cylinder(n_el=16)
    dm = DeviceModel(pr)
    dm.set_u(np.zeros(dm.sym.N))
    dm.assemble(residual=True, tangent=True)
    b = -dm.R.clone()
    x_cg = dm.solve(b).clone()
    x_gm = torch.empty_like(b)
    its, rel = dm._gmres(b, x_gm, 1e-11, 2000)
    assert rel < 1e-11 and its > 0
    # GMRES minimises the residual: with kappa(K) ~ 1e10 a 1e-12 residual bounds the error far less tightly than
    # CG's energy-norm minimisation does, so compare residuals, and solutions only loosely
    r = b.clone()
    dm.spmv(dm.K, x_gm, r, alpha=-1.0, beta=1.0)
    assert float(torch.linalg.vector_norm(r) / torch.linalg.vector_norm(b)) < 1e-9
    assert float(torch.linalg.vector_norm(x_gm - x_cg) / torch.linalg.vector_norm(x_cg)) < 1e-2
