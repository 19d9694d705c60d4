"""CPU: sparsity patterns and DoF maps of the symbolic phase are bit-exact
with the oracle (north_star: 'sparsity patterns and DOF maps bit-exact')."""
import numpy as np
import pytest
import scipy.sparse as sp
from oracle.model import OracleModel
from goldfish_b200.symbolic import Symbolic
import cases


@pytest.mark.parametrize("case", ["tbeam_small", "slr_small", "plate_c1"])
def test_patterns_match_oracle(case):
    pr, kw = getattr(cases, case)()
    S = Symbolic(pr, **kw)
    m = OracleModel(pr)
    assert S.N == m.N and S.n_th == m.n_th
    K = m.stiffness()
    assert np.array_equal(K.indptr, S.K_indptr) and np.array_equal(K.indices, S.K_indices)
    T = m.dRdt()
    assert np.array_equal(T.indptr, S.T_indptr) and np.array_equal(T.indices, S.T_indices)
    assert np.array_equal(np.sort(m.bc_global), S.bc_list)
    assert np.abs(S.f_const - m.f_const).max() <= 1e-13 * max(1.0, np.abs(m.f_const).max())
    for fi, f in enumerate(S.opt_field):
        A = m.dRdCP(f, S.shopt_surf_inds[fi])
        sh = sp.csr_matrix((np.ones(S.P_indptr[fi][-1]), S.P_indices[fi], S.P_indptr[fi]), shape=A.shape)
        pp = S.penP[fi]
        pen = sp.csr_matrix((np.ones(pp["nnz"]), pp["indices"], pp["indptr"]), shape=A.shape)
        U = (sh + pen).tocsr(); U.sort_indices()
        assert np.array_equal(U.indptr, A.indptr) and np.array_equal(U.indices, A.indices)
    al = np.concatenate([np.stack([I.alpha_d[I.ev_v], I.alpha_r[I.ev_v]], 1) for I in m.interfaces])
    assert np.abs(al - S.pen["alpha"]).max() < 1e-13 * np.abs(al).max()


def test_colouring_is_conflict_free():
    pr, kw = cases.plate_c1()
    S = Symbolic(pr)
    for c in range(S.num_colors):
        els = S.color_elem[S.color_ptr[c]:S.color_ptr[c + 1]]
        seen = set()
        for e in els:
            P = S.patches[S.elem_patch[e]]
            I0 = P.first_u[S.elem_eu[e]]; J0 = P.first_v[S.elem_ev[e]]
            cps = {(P.index, I0 + a, J0 + b) for a in range(4) for b in range(4)}
            assert not (cps & seen)
            seen |= cps


def test_empty_and_ragged_inputs():
    pr, kw = cases.tbeam_small()
    pr["interfaces"] = []                                  # no intersections at all
    S = Symbolic(pr)
    assert S.pen["n_eval"] == 0 and S.row_nlow.max() == 0
    pr2, _ = cases.tbeam_small()
    pr2["patches"][0]["p"] = (2, 3)
    with pytest.raises(ValueError):
        Symbolic(pr2)


@pytest.mark.parametrize("case", ["slr_small", "plate_c1"])
def test_field_rows_of_a_control_point_share_one_column_list(case):
    """Row layout the kernels rely on: the three field rows of a control point hold the SAME
    column list, made of groups [field 0 | field 1 | field 2] over one CP list (own stencil, then one group per
    coupled patch).  This is what lets the scatter compute positions arithmetically, and what a node-wise SpMV
    (one index read per three rows) will use."""
    pr, kw = getattr(cases, case)()
    S = Symbolic(pr, **kw)
    ip, ix = S.K_indptr, S.K_indices
    for P in S.patches:
        r0 = P.dof_off + np.arange(P.ncp)
        for i in (1, 2):
            ri = r0 + i * P.ncp
            assert np.array_equal(ip[ri + 1] - ip[ri], ip[r0 + 1] - ip[r0])
        for a in range(0, P.ncp, max(1, P.ncp // 64)):                    # sample of control points
            c0 = ix[ip[r0[a]]:ip[r0[a] + 1]]
            for i in (1, 2):
                r = r0[a] + i * P.ncp
                assert np.array_equal(ix[ip[r]:ip[r + 1]], c0)
            # own-stencil group: three consecutive blocks over the same CP list, shifted by ncp
            nl, Sa = int(S.row_nlow[P.cp_off + a]), int(S.S_all[P.cp_off + a])
            own = c0[nl:nl + 3 * Sa].reshape(3, Sa)
            assert np.array_equal(own[1], own[0] + P.ncp) and np.array_equal(own[2], own[0] + 2 * P.ncp)
            assert own[0].min() >= P.dof_off and own[0].max() < P.dof_off + P.ncp


def test_node_wise_product_addressing_emulated():
    """numpy emulation of gf_spmv_node's addressing (node_row0 / node_stride as DeviceModel.spmv_node builds them):
    index list of the field-0 row applied to the value streams of the three field rows == K x."""
    pr, kw = cases.slr_small()
    S = Symbolic(pr, **kw)
    rng = np.random.default_rng(0)
    vals = rng.standard_normal(S.K_indptr[-1]); x = rng.standard_normal(S.N)
    K = sp.csr_matrix((vals, S.K_indices, S.K_indptr), shape=(S.N, S.N))
    row0 = np.concatenate([P.dof_off + np.arange(P.ncp, dtype=np.int64) for P in S.patches])
    stride = np.concatenate([np.full(P.ncp, P.ncp, dtype=np.int32) for P in S.patches])
    y = np.zeros(S.N)
    for r0, st in zip(row0, stride):
        s0 = S.K_indptr[r0]; ln = S.K_indptr[r0 + 1] - s0
        xv = x[S.K_indices[s0:s0 + ln]]
        for i in range(3):
            si = S.K_indptr[r0 + i * st]
            y[r0 + i * st] = vals[si:si + ln] @ xv
    assert len(row0) == S.n_scalar and np.abs(y - K @ x).max() < 1e-12
