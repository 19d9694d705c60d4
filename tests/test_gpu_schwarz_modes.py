"""GPU: the triangular-sweep kernels of the Schwarz preconditioner (fine blocks: one CTA per block with
the vector in shared memory / a CTA group per block with a global-memory barrier; coarse block: one
thread-block cluster / a barrier group) apply the same operator: same CG iteration counts, same
solution, and that solution is the LU one."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla
import torch

from goldfish_b200 import problems

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("sub", [(8, 16), 48])
def test_sweep_kernels_agree(built_lib, sub):
    from goldfish_b200.device_model import DeviceModel
    pr = problems.cylinder(n_el=24)
    dm = DeviceModel(pr, schwarz_sub=sub)
    dm.set_u(np.zeros(dm.sym.N))
    dm.assemble(residual=True, tangent=True)
    dm.factor_preconditioner()
    b = -dm.R.clone()
    rng = np.random.default_rng(3)
    r = torch.from_numpy(rng.standard_normal(dm.sym.N)).to(dm.device)
    r[torch.from_numpy(np.asarray(dm.sym.bc_list)).to(dm.device).long()] = 0.0
    out = {}
    for mode in ("group", "single_nocluster", "single"):
        dm.set_sweep_mode(mode)
        z = dm.precond_apply(r).cpu().numpy().copy()
        z2 = dm.precond_apply(r).cpu().numpy().copy()
        assert np.array_equal(z, z2)                          # bit-reproducible from call to call
        x = dm.solve(b).cpu().numpy().copy()
        out[mode] = (z, x, dm.last_krylov_its)
    torch.cuda.synchronize()
    zg, xg, ig = out["group"]
    for mode in ("single_nocluster", "single"):
        zs, xs, i_s = out[mode]
        assert np.isfinite(zs).all() and rel(zs, zg) < 1e-3      # FP32 panel products, different summation order
        assert abs(ig - i_s) <= 15                               # two passes, each checked every 5 iterations
        assert rel(xs, xg) < 1e-8
    xe = spla.splu(dm.K.to_scipy().tocsc()).solve(b.cpu().numpy())
    assert rel(xs, xe) < 1e-6
