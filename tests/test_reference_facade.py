"""CPU: the reference's OWN operation file, unmodified, on top of this repo's shim namespace.

/root/reference/GOLDFISH/operations/disp_imop.py does `from GOLDFISH.nonmatching_opt_ffd import *` and then only
calls get_petsc_vec_array / update_nest_vec / A_x_b / AT_x_b / solve_Ax_b / solve_ATx_b and the NonMatchingOpt
methods.  Here that file is loaded byte-for-byte with `GOLDFISH.nonmatching_opt_ffd` bound to
goldfish_b200.opt_utils, and run -- side by side with goldfish_b200/operations/disp_imop.py -- over one backend
object: a numpy/oracle stand-in for NonMatchingOpt with petsc4py-shaped handles (the CUDA model needs a GPU and the
GPU box has no /root/reference, so the backend is the oracle; what is tested is the FACADE contract: same calls, same
`+=` / `[:]=` semantics, same numbers from both files).  Skipped where /root/reference is absent."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import scipy.sparse.linalg as spla

import cases
from oracle.model import OracleModel

REF = "/root/reference/GOLDFISH/operations/disp_imop.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference sources not present")


class NpVec:
    type = "nest"

    def __init__(self, a, owner=None):
        self.data_np = np.array(a, dtype=np.float64); self.owner = owner

    array = property(lambda self: self.data_np.copy())
    data = property(lambda self: self.data_np)

    def copy(self): return NpVec(self.data_np, self.owner)
    def setArray(self, a): self.data_np[:] = a
    def assemble(self): return None


class NpMat:
    def __init__(self, A, owner, is_K=False):
        self.A, self.owner, self.is_K = A.tocsr(), owner, is_K

    def copy(self): return NpMat(self.A.copy(), self.owner, self.is_K)
    def mult(self, x, y): y.data_np[:] = self.A @ x.data_np
    def multTranspose(self, x, y): y.data_np[:] = self.A.T @ x.data_np


class OracleBackend:
    """The L2 surface the operation files touch, over the numpy oracle."""
    comm = None
    use_aero_pressure = False
    var_thickness = False

    def __init__(self, pr, kw):
        self.om = OracleModel(pr)
        self.opt_field = kw["opt_field"]; self.surf = kw["shopt_surf_inds"]
        self.opt_shape, self.opt_thickness = True, True
        n_sc = sum(P.ncp for P in self.om.patches)
        self.vec_iga_nest = NpVec(np.zeros(self.om.N), self)
        self.vec_scalar_iga_nest = NpVec(np.zeros(n_sc), self)
        self.cpdes_iga_nest = [NpVec(np.zeros(n_sc), self) for _ in self.opt_field]
        self.h_th_nest = NpVec(np.zeros(self.om.n_th), self)
        self.dm = self
        self._cache = {}

    def _cached(self, key, fn):
        """The oracle is slow and this test is about the facade, not the operators: evaluate each operator once per state."""
        k = (key, self.om.u.tobytes())
        if k not in self._cache:
            self._cache[k] = fn()
        return self._cache[k]

    def solve(self, b, x):                       # what opt_utils._solve calls on the matrix owner
        lu = self._cached("lu", lambda: spla.splu(self.om.stiffness().tocsc()))
        x[:] = lu.solve(b)

    def RIGA(self): return NpVec(self._cached("R", self.om.residual), self)
    def dRIGAduIGA(self): return NpMat(self._cached("K", self.om.stiffness), self, is_K=True)

    def dRIGAdCPIGA(self, field):
        return NpMat(self._cached(("P", field), lambda: self.om.dRdCP(field, self.surf[self.opt_field.index(field)])), self)

    def dRIGAdh_th(self): return NpMat(self._cached("T", self.om.dRdt), self)

    def solve_nonlinear_nonmatching_problem(self, max_it=30, zero_mortar_funcs=True, rtol=1e-3, iga_dofs=True):
        k = ("newton", max_it, rtol)
        if k not in self._cache:
            self._cache[k] = self.om.solve_nonlinear(max_it=max_it, rtol=rtol)
        self.om.set_u(self._cache[k])
        return None, NpVec(self._cache[k], self)


def _load_reference_class():
    from goldfish_b200 import opt_utils
    shim = types.ModuleType("GOLDFISH.nonmatching_opt_ffd")
    for k in ("get_petsc_vec_array", "update_nest_vec", "A_x_b", "AT_x_b", "solve_Ax_b", "solve_ATx_b", "A_x", "AT_x"):
        setattr(shim, k, getattr(opt_utils, k))
    shim.np = np
    pkg = types.ModuleType("GOLDFISH"); pkg.__path__ = []
    saved = {k: sys.modules.get(k) for k in ("GOLDFISH", "GOLDFISH.nonmatching_opt_ffd")}
    sys.modules["GOLDFISH"], sys.modules["GOLDFISH.nonmatching_opt_ffd"] = pkg, shim
    try:
        spec = importlib.util.spec_from_file_location("_ref_disp_imop", REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)              # the reference file itself, byte for byte
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod.DispImOpeartion


def test_reference_disp_imop_runs_unmodified_and_matches_our_facade():
    from goldfish_b200.operations.disp_imop import DispImOpeartion as Ours
    Ref = _load_reference_class()
    pr, kw = cases.tbeam_small()
    kw = dict(opt_field=[0, 2], shopt_surf_inds=[[0, 1], [0, 1]])
    nm = OracleBackend(pr, kw)
    a, b = Ref(nm), Ours(nm)
    ua, ub = a.solve_nonlinear(max_it=30, rtol=1e-3), b.solve_nonlinear(max_it=30, rtol=1e-3)
    assert np.array_equal(ua, ub)
    nm.om.set_u(ua)
    assert np.array_equal(a.apply_nonlinear(), b.apply_nonlinear())
    a.linearize(); b.linearize()
    rng = np.random.default_rng(0)
    N, n_sc, n_th = nm.om.N, nm.vec_scalar_iga_nest.data_np.size, nm.om.n_th
    d_in = [rng.standard_normal(n_sc), rng.standard_normal(n_sc), rng.standard_normal(n_th)]
    d_out = rng.standard_normal(N)
    ra, rb = np.ones(N), np.ones(N)                  # accumulate (+=) into the caller's array
    a.apply_linear_fwd([x.copy() for x in d_in], d_out.copy(), ra)
    b.apply_linear_fwd([x.copy() for x in d_in], d_out.copy(), rb)
    assert np.allclose(ra, rb, rtol=1e-14, atol=0) and not np.allclose(ra, 1.0)
    res = rng.standard_normal(N)
    ia, ib = [np.ones(n_sc), np.ones(n_sc), np.ones(n_th)], [np.ones(n_sc), np.ones(n_sc), np.ones(n_th)]
    oa, ob = np.ones(N), np.ones(N)
    a.apply_linear_rev(ia, oa, res.copy()); b.apply_linear_rev(ib, ob, res.copy())
    assert all(np.allclose(x, y, rtol=1e-14, atol=0) for x, y in zip(ia + [oa], ib + [ob]))
    xa, xb = np.zeros(N), np.zeros(N)                # overwrite ([:] =)
    rhs = nm.om.dWdu(apply_bcs=True)
    a.solve_linear_rev(rhs.copy(), xa); b.solve_linear_rev(rhs.copy(), xb)
    assert np.allclose(xa, xb, rtol=1e-12, atol=0) and np.linalg.norm(xa) > 0
    fa, fb = np.zeros(N), np.zeros(N)
    a.solve_linear_fwd(fa, res.copy()); b.solve_linear_fwd(fb, res.copy())
    assert np.allclose(fa, fb, rtol=1e-12, atol=0)
