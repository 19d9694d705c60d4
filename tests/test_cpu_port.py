"""CPU: the compiled C++/OpenMP port (oracle/c/kl_cpu.cpp) agrees with the numpy
oracle -- two independent CPU restatements (jets vs. dual numbers over the
hand-derived first variation) of the same shell quadrature."""
import numpy as np
import scipy.sparse as sp
import pytest
import cases
from oracle.model import OracleModel
from oracle.cpu_port import CpuModel
from goldfish_b200 import _capi as capi


@pytest.mark.parametrize("case", ["tbeam_small", "slr_small"])
def test_port_matches_numpy_oracle(case):
    pr, kw = getattr(cases, case)()
    cm = CpuModel(pr, **kw)
    om = OracleModel(pr)
    u = cases.random_state(om.N, om.bc_global)
    cm.set_u(u); om.set_u(u)
    cm.shell(capi.GF_OUT_R | capi.GF_OUT_K | capi.GF_OUT_W | capi.GF_OUT_P | capi.GF_OUT_T)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(cm.residual(), om.residual()) < 1e-11
    K = cm.K_matrix(); Ko = om.stiffness()
    assert abs(K - Ko).max() < 1e-11 * abs(Ko).max()
    T = sp.csr_matrix((cm.Tv, cm._idx["T"][1], cm._idx["T"][0]), shape=(om.N, om.n_th))
    assert abs(T - om.dRdt()).max() < 1e-11 * abs(om.dRdt()).max()
    assert abs(cm.WV[0::2].sum() - om.energy()) < 1e-11 * om.energy()
    assert rel(cm.dWdt, om.dWdt()) < 1e-11 and rel(cm.dVdt, om.dVdt()) < 1e-11
    for i, f in enumerate(kw["opt_field"]):
        Psh = sp.csr_matrix((cm.Pv[i], cm._idx["P"][i][1], cm._idx["P"][i][0]), shape=(om.N, cm.S.P_ncols[i]))
        Ao = om.dRdCP_fields([f], kw["shopt_surf_inds"][i], penalty=False)[0]
        assert abs(Psh - Ao).max() < 1e-11 * abs(Ao).max()
        assert rel(cm.dWdP[i], om.dWdCP(f, kw["shopt_surf_inds"][i])) < 1e-11
