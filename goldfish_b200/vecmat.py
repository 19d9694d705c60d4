"""petsc4py-shaped handles over HBM buffers.

The reference's L3 facade only ever touches PETSc objects through a handful of
members: Vec ``.array .copy() .getNestSubVecs()
.getSizes() .setArray() .assemble() .norm() .zeroEntries()`` and Mat
``.copy() .transpose() .getValuesCSR() .size .mult() .multTranspose()``.
These duck-typed classes provide exactly that over device memory so that code
written against /root/reference/GOLDFISH/operations/*.py keeps its shape.
"""
import numpy as np
import torch


class DeviceVec:
    """Nested vector: one contiguous FP64 device buffer + the sub-vector sizes."""
    type = "nest"

    def __init__(self, data, sizes=None, owner=None):
        self.data = data                      # torch tensor on the GPU
        self.sizes = list(sizes) if sizes is not None else [data.numel()]
        self.owner = owner

    @classmethod
    def zeros(cls, sizes, device, owner=None):
        return cls(torch.zeros(int(sum(sizes)), dtype=torch.float64, device=device), sizes, owner)

    # -- petsc4py surface ---------------------------------------------------
    @property
    def array(self):
        return self.data.cpu().numpy()

    @property
    def size(self):
        return self.data.numel()

    def getSizes(self):
        return (self.data.numel(), self.data.numel())

    def getSize(self):
        return self.data.numel()

    def getOwnershipRange(self):
        return (0, self.data.numel())

    def copy(self):
        return DeviceVec(self.data.clone(), self.sizes, self.owner)

    def setArray(self, arr):
        self.data.copy_(torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float64)).reshape(-1), non_blocking=True)

    def zeroEntries(self):
        self.data.zero_()

    def assemble(self):
        return None

    def ghostUpdate(self):
        return None

    def norm(self):
        if self.owner is not None:
            return self.owner.dot(self.data, self.data) ** 0.5
        return float(np.linalg.norm(self.array))

    def getNestSubVecs(self):
        out, o = [], 0
        for n in self.sizes:
            out.append(DeviceVec(self.data[o:o + n], [n], self.owner))
            o += n
        return out


class DeviceMat:
    """Sum of CSR parts living in HBM (one part for K / dR/dt, shell + penalty
    parts for dR/dCP).  ``transposed`` only flips which product ``mult`` does."""

    def __init__(self, owner, parts, transposed=False, is_K=False):
        self.owner, self.parts, self.transposed, self.is_K = owner, parts, transposed, is_K

    @property
    def shape(self):
        r, c = self.parts[0].nrows, self.parts[0].ncols
        return (c, r) if self.transposed else (r, c)

    size = shape
    sizes = shape

    def getSizes(self):
        r, c = self.shape
        return ((r, r), (c, c))

    def copy(self):
        # values are only replaced by the next linearize(); a handle copy is
        # all solve_linear_* needs (disp_imop.py:131,138 deep-copies K first)
        return DeviceMat(self.owner, self.parts, self.transposed, self.is_K)

    def transpose(self):
        self.transposed = not self.transposed
        return self

    def _apply(self, x, y, transpose):
        for i, A in enumerate(self.parts):
            self.owner.spmv_global(A, x, y, 1.0, 0.0 if i == 0 else 1.0, transpose=transpose)
        return y

    def mult(self, x, y):
        self._apply(x.data, y.data, self.transposed)

    def multTranspose(self, x, y):
        self._apply(x.data, y.data, not self.transposed)

    def to_scipy(self):
        """Merged CSR (sorted columns); structural zeros are kept, as PETSc
        keeps them (scipy's `+` would silently drop entries that sum to 0.0)."""
        import scipy.sparse as sp
        mats = [P.to_scipy().tocoo() for P in self.parts]
        A = sp.coo_matrix((np.concatenate([m.data for m in mats]),
                           (np.concatenate([m.row for m in mats]), np.concatenate([m.col for m in mats]))),
                          shape=mats[0].shape).tocsr()
        A.sum_duplicates()
        A.sort_indices()
        return A.T.tocsr() if self.transposed else A

    def getValuesCSR(self):
        A = self.to_scipy()
        return A.indptr, A.indices, A.data
