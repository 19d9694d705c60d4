"""ORACLE / CPU BASELINE (test + measurement infrastructure, never the product path).

Host driver of oracle/c/mf_lu.cpp: the multi-core sparse direct solver standing in for the reference's
PETSc PCLU + MUMPS solves (/root/reference/GOLDFISH/utils/opt_utils.py:156-209).  Analysis (nested dissection
of the control-point graph + symbolic factorisation, oracle/nested_dissection.py) is done once per pattern;
`factor` / `solve` run the numeric phases with all host threads (OpenMP over independent subtrees, threaded
BLAS on the large fronts near the root).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

from .nested_dissection import nested_dissection, symbolic

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "c", "libgfo_mf.so")


def _capsule_ptr(mod, name):
    cap = mod.__pyx_capi__[name]
    C.pythonapi.PyCapsule_GetName.restype = C.c_char_p
    C.pythonapi.PyCapsule_GetName.argtypes = [C.py_object]
    C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
    C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    return C.c_void_p(C.pythonapi.PyCapsule_GetPointer(cap, C.pythonapi.PyCapsule_GetName(cap)))


def _openblas_set_threads():
    """Pointer to the set-num-threads entry of the OpenBLAS that scipy's BLAS capsules resolve to (or None)."""
    try:
        from threadpoolctl import ThreadpoolController
        import scipy.linalg  # noqa: F401  (loads the library)
        for lib in ThreadpoolController().lib_controllers:
            if lib.internal_api == "openblas" and "scipy.libs" in (lib.filepath or ""):
                dll = lib.dynlib
                for nm in ("openblas_set_num_threads", "openblas_set_num_threads64_", "scipy_openblas_set_num_threads64_",
                           "scipy_openblas_set_num_threads"):
                    fn = getattr(dll, nm, None)
                    if fn is not None:
                        return C.cast(fn, C.c_void_p), dll
    except Exception:
        pass
    return C.c_void_p(0), None


class MultifrontalLU:
    def __init__(self, G, X, node_dofs, N, leaf=64, big=1200):
        """G: scalar control-point graph (scipy CSR, symmetric pattern); X: (n_s, 3) coordinates;
        node_dofs: (n_s, 3) global dof of each control point's three fields; N: dofs."""
        if not os.path.exists(LIB):
            subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "c")])
        lib = self.lib = C.CDLL(LIB)
        lib.mf_create.restype = C.c_void_p
        lib.mf_create.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_int64]
        lib.mf_set_blas.argtypes = [C.c_void_p] * 8 + [C.c_int]
        lib.mf_factor.argtypes = [C.c_void_p] * 9
        lib.mf_solve.argtypes = [C.c_void_p] * 4
        lib.mf_flops.restype = C.c_double; lib.mf_flops.argtypes = [C.c_void_p]
        lib.mf_lu_doubles.restype = C.c_int64; lib.mf_lu_doubles.argtypes = [C.c_void_p]
        lib.mf_destroy.argtypes = [C.c_void_p]; lib.mf_free_numeric.argtypes = [C.c_void_p]
        fronts, parent = nested_dissection(G, X, leaf)
        pos_n, first_n, border_n, _ = symbolic(fronts, parent, G)
        nf = len(fronts)
        perm_nodes = np.concatenate(fronts)
        node_dofs = np.asarray(node_dofs, dtype=np.int64)
        self.perm = np.ascontiguousarray(node_dofs[perm_nodes].reshape(-1))          # permuted position -> dof
        assert len(self.perm) == N and len(np.unique(self.perm)) == N
        self.pos = np.empty(N, dtype=np.int64); self.pos[self.perm] = np.arange(N)
        k = np.array([3 * len(f) for f in fronts], dtype=np.int32)
        u = np.array([3 * len(b) for b in border_n], dtype=np.int32)
        first = (3 * np.asarray(first_n[:-1])).astype(np.int64)
        bptr = np.zeros(nf + 1, dtype=np.int64); np.cumsum(u, out=bptr[1:])
        bidx = np.concatenate([(3 * b[:, None] + np.arange(3)[None, :]).reshape(-1) for b in border_n]).astype(np.int64) \
            if bptr[-1] else np.zeros(0, dtype=np.int64)
        self._keep = (k, u, first, bptr, bidx, np.ascontiguousarray(parent, dtype=np.int32))
        p = lambda a: C.c_void_p(a.ctypes.data)
        self.h = C.c_void_p(lib.mf_create(nf, p(k), p(u), p(first), p(bptr), p(bidx), p(self._keep[5]), N))
        import scipy.linalg.cython_blas as cb
        import scipy.linalg.cython_lapack as cl
        st, self._dll = _openblas_set_threads()
        lib.mf_set_blas(self.h, _capsule_ptr(cb, "dgemm"), _capsule_ptr(cb, "dtrsm"), _capsule_ptr(cl, "dgetrf"),
                        _capsule_ptr(cl, "dlaswp"), _capsule_ptr(cb, "dgemv"), _capsule_ptr(cb, "dtrsv"), st, big)
        nthr = int(os.environ.get("GFO_THREADS", "0"))
        if nthr > 0:
            lib.mf_set_threads.argtypes = [C.c_void_p, C.c_int]
            lib.mf_set_threads(self.h, nthr)
        self.N, self.nfronts = N, nf
        self.flops = lib.mf_flops(self.h)
        self.lu_bytes = 8 * lib.mf_lu_doubles(self.h)
        self.max_front = int((k + u).max())
        self.blas_threads_controlled = bool(st.value)

    def factor(self, A, AT=None):
        """Numeric LU of the CSR matrix A (AT = its transpose as CSR; None: A is symmetric)."""
        A = A.tocsr() if not sp.isspmatrix_csr(A) else A
        AT = A if AT is None else AT
        self._A = (np.ascontiguousarray(A.indptr, dtype=np.int64), np.ascontiguousarray(A.indices, dtype=np.int32),
                   np.ascontiguousarray(A.data, dtype=np.float64),
                   np.ascontiguousarray(AT.indptr, dtype=np.int64), np.ascontiguousarray(AT.indices, dtype=np.int32),
                   np.ascontiguousarray(AT.data, dtype=np.float64))
        p = lambda a: C.c_void_p(a.ctypes.data)
        info = self.lib.mf_factor(self.h, *[p(a) for a in self._A], p(self.perm), p(self.pos))
        if info != 0:
            raise RuntimeError("multifrontal LU: zero pivot (dgetrf info = %d)" % info)
        return self

    def solve(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        self.lib.mf_solve(self.h, C.c_void_p(b.ctypes.data), C.c_void_p(x.ctypes.data), C.c_void_p(self.perm.ctypes.data))
        return x

    def free_numeric(self):
        self.lib.mf_free_numeric(self.h)

    def __del__(self):
        try:
            self.lib.mf_destroy(self.h)
        except Exception:
            pass
