"""GPU: the whole analysis + adjoint iteration at a MID size (8-patch non-matching cylinder, n_el = 32, 35.7 k dofs,
the bench workload's topology and design perturbation) against the compiled CPU port with its LU path
(oracle/cpu_port.py: multifrontal LU per Newton step, re-factorised transpose for the adjoint, iteratively refined):
displacements, adjoint vector and every total gradient to north_star's 1e-8; objective to 1e-8."""
import numpy as np
import pytest
import torch

import bench
from oracle.cpu_port import CpuModel

pytestmark = pytest.mark.gpu


def test_iteration_matches_cpu_lu_path_at_mid_size(built_lib):
    from goldfish_b200.device_model import DeviceModel
    pr, kw = bench.workload(32)
    cm = CpuModel(pr, **kw)
    cp, th = bench.design_state(cm.S)
    cm.cp[:] = cp; cm.theta[:] = th
    _, g_cpu = cm.iteration(newton_rtol=1e-3, refine=2)
    dm = DeviceModel(pr, **kw)
    dm.cp.copy_(torch.from_numpy(cp)); dm.set_theta(th)
    st = bench.Step(dm)
    st()
    torch.cuda.synchronize()
    rel = lambda a, b: float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))
    assert st.info["newton_its"] == cm.newton_its
    assert rel(dm.u.cpu().numpy(), cm.u) < 1e-8
    assert abs(float(dm.wv_sum[0]) - cm.W) < 1e-8 * abs(cm.W)
    assert rel(st.lam.cpu().numpy(), cm.lam) < 1e-8
    for i in range(3):
        assert rel(st.gP[i].cpu().numpy(), g_cpu[i]) < 1e-8
    assert rel(st.gT.cpu().numpy(), g_cpu[3]) < 1e-8
    assert max(t for t in st.info["true_relres"] if t is not None) < 1e-8
