"""Marshalling and solve helpers with the names of
/root/reference/GOLDFISH/utils/opt_utils.py (and the PENGoLINS helpers it
star-imports) that the operations layer calls."""
import numpy as np
import torch
from .vecmat import DeviceVec, DeviceMat


def get_petsc_vec_array(petsc_vec, comm=None):
    """opt_utils.py:28-54 -- global values as ndarray (device -> host copy)."""
    return petsc_vec.array


def update_nest_vec(vec_array, nest_vec, comm=None):
    """opt_utils.py:70-103 -- assign a numpy array to a nested vector."""
    if getattr(nest_vec, "type", None) != "nest":
        raise TypeError("Type of PETSc vector is not nest.")
    nest_vec.setArray(vec_array)
    nest_vec.assemble()


def A_x_b(A, x, b):
    """b = A x (PENGoLINS A_x_b; disp_imop.py:68,76,86)."""
    A.mult(x, b)


def AT_x_b(A, x, b):
    """b = A^T x (disp_imop.py:109,115,121)."""
    A.multTranspose(x, b)


def A_x(A, x):
    b = DeviceVec.zeros([A.shape[0]], x.data.device, x.owner)
    A.mult(x, b)
    return b


def AT_x(A, x):
    b = DeviceVec.zeros([A.shape[1]], x.data.device, x.owner)
    A.multTranspose(x, b)
    return b


def _solve(A, b, array):
    if not getattr(A, "is_K", False):
        raise TypeError("solve_Ax_b is only defined for the tangent dR/du on this path")
    dm = A.owner
    x = b.copy()
    dm.solve(b.data, x.data)
    return x.array if array else x


def solve_Ax_b(A, b, array=False, comm=None):
    """opt_utils.py:156-181: x = K^{-1} b (sparse LU there, preconditioned CG here)."""
    return _solve(A, b, array)


def solve_ATx_b(A, b, array=False, comm=None):
    """opt_utils.py:183-209: x = K^{-T} b.  K is symmetric by construction
    (nonmatching_opt.py:804-809), so the transpose solve is the same solve."""
    return _solve(A, b, array)
