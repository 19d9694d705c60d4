"""ctypes binding of include/goldfish_b200.h (the C ABI of the CUDA library).

The library is built in-tree by ``__graft_entry__.build()`` into
``goldfish_b200/libgoldfish_b200.so``.  There is NO CPU fallback: if the
library or a CUDA device is missing, loading/compute raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GF_LIB", os.path.join(_HERE, "libgoldfish_b200.so"))   # GF_LIB: tuning experiments only

GF_OUT_R, GF_OUT_K, GF_OUT_W, GF_OUT_P, GF_OUT_T = 1, 2, 4, 8, 16
GF_ERR_BADARG, GF_ERR_CUDA, GF_ERR_NOCONV, GF_ERR_BREAKDOWN, GF_ERR_NAN = 1, 2, 3, 4, 5
GF_ERRORS = {1: "bad argument", 2: "CUDA error", 3: "not converged", 4: "breakdown", 5: "NaN"}

c_i32, c_i64, c_f64, c_vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p


class GfPatchDesc(C.Structure):
    _fields_ = [("n_u", c_i32), ("n_v", c_i32), ("neu", c_i32), ("nev", c_i32),
                ("cp_off", c_i32), ("dof_off", c_i32),
                ("th_off", c_i32), ("th_kind", c_i32), ("nth", c_i32),
                ("span_u_off", c_i32), ("span_v_off", c_i32),
                ("cpd_u_off", c_i32), ("cpd_v_off", c_i32),
                ("rational", c_i32), ("pcol_off", c_i32 * 3), ("el_off", c_i32),
                ("E", c_f64), ("nu", c_f64), ("f", c_f64 * 3)]


class GfCsr(C.Structure):
    _fields_ = [("nrows", c_i64), ("ncols", c_i64), ("nnz", c_i64),
                ("indptr", c_vp), ("indices", c_vp), ("vals", c_vp)]


class GfModel(C.Structure):
    _fields_ = [("num_patches", c_i32), ("num_elements", c_i32), ("nq", c_i32), ("num_colors", c_i32),
                ("N", c_i64), ("n_scalar", c_i64), ("n_th", c_i64),
                ("patches", c_vp), ("elem_patch", c_vp), ("elem_eu", c_vp), ("elem_ev", c_vp),
                ("color_elem", c_vp), ("color_ptr_h", c_vp),
                ("tab_u", c_vp), ("tab_v", c_vp), ("first_cp_u", c_vp), ("first_cp_v", c_vp),
                ("span_h_u", c_vp), ("span_h_v", c_vp), ("qw", c_vp), ("tw_lin", c_vp),
                ("cp_lo_u", c_vp), ("cp_hi_u", c_vp), ("el_lo_u", c_vp), ("el_hi_u", c_vp),
                ("cp_lo_v", c_vp), ("cp_hi_v", c_vp), ("el_lo_v", c_vp), ("el_hi_v", c_vp),
                ("cp", c_vp), ("u", c_vp), ("theta", c_vp), ("bc", c_vp), ("bc_list", c_vp),
                ("n_bc", c_i64), ("row_nlow", c_vp),
                ("K", GfCsr), ("P", GfCsr * 3), ("T", GfCsr)]


class GfShellOut(C.Structure):
    _fields_ = [("R", c_vp), ("WV", c_vp), ("dWdu", c_vp), ("dWdP", c_vp * 3), ("dVdP", c_vp * 3),
                ("dWdt", c_vp), ("dVdt", c_vp), ("dt_el", c_vp)]


class GfPenalty(C.Structure):
    _fields_ = [("n_eval", c_i64),
                ("connA", c_vp), ("connB", c_vp), ("connC0", c_vp), ("connC1", c_vp),
                ("basA", c_vp), ("basB", c_vp), ("basC0", c_vp), ("basC1", c_vp),
                ("tpar", c_vp), ("alpha", c_vp), ("dofA", c_vp), ("dofB", c_vp),
                ("g", c_vp), ("Huu", c_vp), ("HuX", c_vp),
                ("nR", c_i64), ("R_ptr", c_vp), ("R_item", c_vp), ("R_row", c_vp),
                ("nK", c_i64), ("K_ptr", c_vp), ("K_item", c_vp), ("K_pos", c_vp)]


class GfPenaltyP(C.Structure):
    _fields_ = [("n_dest", c_i64), ("ptr", c_vp), ("item_eval", c_vp), ("item_code", c_vp),
                ("pos", c_vp), ("vals", c_vp), ("field", c_i32)]


class GfCsrT(C.Structure):
    _fields_ = [("nrows", c_i64), ("nnz", c_i64), ("indptr", c_vp), ("indices", c_vp), ("perm", c_vp)]


class GfSchwarz(C.Structure):
    _fields_ = [("nblocks", c_i32), ("nb", c_i32), ("max_nbr", c_i32), ("max_mb", c_i32), ("max_n_pad", c_i32),
                ("debug_flags", c_i32), ("n_y", c_i64), ("band_len", c_i64),
                ("n_pad", c_vp), ("nbr", c_vp), ("off_j", c_vp), ("mbj", c_vp), ("rlen", c_vp), ("off_col", c_vp),
                ("step_mb_h", c_vp), ("off_y", c_vp), ("off_inv", c_vp),
                ("glob", c_vp), ("gs", c_vp), ("ls", c_vp), ("off_g", c_vp), ("zptr", c_vp), ("zsrc", c_vp),
                ("band", c_vp), ("band32", c_vp), ("invd", c_vp), ("y", c_vp), ("s", c_vp), ("barrier", c_vp), ("flag", c_vp)]


class GfDist(C.Structure):
    _fields_ = [("n_ranges", c_i32), ("rank", c_i32), ("world", c_i32), ("pad_", c_i32), ("ranges_h", c_vp), ("comm", c_vp)]


class GfPrecond(C.Structure):
    _fields_ = [("fine", C.POINTER(GfSchwarz)), ("coarse", C.POINTER(GfSchwarz)), ("P", GfCsr), ("Rt", GfCsr),
                ("rc", c_vp), ("zc", c_vp), ("bc_c", c_vp), ("n_bc_c", c_i64), ("dist", C.POINTER(GfDist)),
                ("cinv", c_vp), ("cinv_row0", c_i64), ("cinv_rows", c_i64)]


class GfNodeRows(C.Structure):
    _fields_ = [("row0", c_vp), ("stride", c_vp), ("n", c_i64)]


class GfPcgWork(C.Structure):
    _fields_ = [("r", c_vp), ("z", c_vp), ("p", c_vp), ("Ap", c_vp), ("dinv", c_vp),
                ("scal", c_vp), ("partial", c_vp), ("scal_h", c_vp), ("nodes", GfNodeRows)]


class GfGmresWork(C.Structure):
    _fields_ = [("V", c_vp), ("Z", c_vp), ("t", c_vp), ("hdev", c_vp), ("partial", c_vp), ("h_host", c_vp), ("nodes", GfNodeRows)]


# (struct, last field) in the order of gf_abi_layout's ids
ABI_STRUCTS = [(GfPatchDesc, "f"), (GfCsr, "vals"), (GfModel, "T"), (GfShellOut, "dt_el"), (GfPenalty, "K_pos"),
               (GfPenaltyP, "field"), (GfCsrT, "perm"), (GfSchwarz, "flag"), (GfDist, "comm"), (GfPrecond, "cinv_rows"),
               (GfPcgWork, "nodes"), (GfGmresWork, "nodes")]


def check_abi(lib):
    """Compare every ctypes mirror with the compiled header (size and offset of the last field)."""
    for i, (T, last) in enumerate(ABI_STRUCTS):
        size, off = c_i64(0), c_i64(0)
        if lib.gf_abi_layout(i, C.byref(size), C.byref(off)) != 0:
            raise GoldfishError("gf_abi_layout(%d) failed" % i)
        if size.value != C.sizeof(T) or off.value != getattr(T, last).offset:
            raise GoldfishError("ABI mismatch for %s: library %d/%d, binding %d/%d"
                                % (T.__name__, size.value, off.value, C.sizeof(T), getattr(T, last).offset))


# every symbol include/goldfish_b200.h declares, with its argument types
SIGNATURES = {
    "gf_shell_assemble": [C.POINTER(GfModel), C.c_int, C.POINTER(GfShellOut), c_vp],
    "gf_bc_set_diag": [C.POINTER(GfModel), c_f64, c_vp],
    "gf_penalty_points": [C.POINTER(GfModel), C.POINTER(GfPenalty), C.c_int, c_vp],
    "gf_penalty_gather_R": [C.POINTER(GfModel), C.POINTER(GfPenalty), c_vp, c_vp],
    "gf_penalty_gather_K": [C.POINTER(GfModel), C.POINTER(GfPenalty), c_vp],
    "gf_penalty_gather_P": [C.POINTER(GfPenalty), C.POINTER(GfPenaltyP), c_vp],
    "gf_mask_vec": [C.POINTER(GfModel), c_vp, c_vp],
    "gf_spmv": [C.POINTER(GfCsr), c_vp, c_vp, c_f64, c_f64, c_vp],
    "gf_spmv_node": [C.POINTER(GfCsr), c_vp, c_vp, c_i64, c_vp, c_vp, c_f64, c_f64, c_vp],
    "gf_spmv_t": [C.POINTER(GfCsr), C.POINTER(GfCsrT), c_vp, c_vp, c_f64, c_f64, c_vp],
    "gf_pcg": [C.POINTER(GfCsr), c_vp, c_vp, C.POINTER(GfPcgWork), C.POINTER(GfPrecond), C.POINTER(GfDist), c_f64, c_f64, C.c_int, C.c_int,
               C.POINTER(C.c_int), C.POINTER(c_f64), c_vp],
    "gf_residual_dd": [C.POINTER(GfCsr), C.POINTER(GfDist), c_vp, c_vp, c_vp, c_vp],
    "gf_gmres": [C.POINTER(GfCsr), c_vp, c_vp, C.POINTER(GfGmresWork), C.POINTER(GfPrecond), C.POINTER(GfDist), c_f64, C.c_int, C.c_int,
                 C.POINTER(C.c_int), C.POINTER(c_f64), c_vp],
    "gf_dist_unique_id": [c_vp],
    "gf_dist_init": [C.POINTER(GfDist), c_vp, C.c_int, C.c_int],
    "gf_dist_allreduce": [C.POINTER(GfDist), c_vp, c_i64, c_vp],
    "gf_dist_destroy": [C.POINTER(GfDist)],
    "gf_schwarz_factor": [C.POINTER(GfSchwarz), C.POINTER(GfCsr), c_vp],
    "gf_schwarz_apply": [C.POINTER(GfSchwarz), c_vp, c_vp, c_i64, c_vp],
    "gf_schwarz_apply2": [C.POINTER(GfSchwarz), c_vp, c_vp, c_i64, C.POINTER(GfSchwarz), c_vp, c_vp, c_i64, c_vp],
    "gf_dot_slot0": [c_i64, c_vp, c_vp, c_vp, C.c_int, c_vp],
    "gf_schwarz_sweeps": [C.POINTER(GfSchwarz), c_vp, c_vp],
    "gf_precond_apply": [C.POINTER(GfPrecond), c_vp, c_vp, c_i64, c_vp],
    "gf_jacobi_setup": [C.POINTER(GfCsr), c_vp, c_vp],
    "gf_axpby": [c_i64, c_f64, c_vp, c_f64, c_vp, c_vp],
    "gf_dot": [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp],
    "gf_reduce_wv": [c_i64, c_vp, c_vp, c_vp],
    "gf_peak_fp64": [C.c_int, C.c_int, C.c_int, c_vp, C.POINTER(c_f64), c_vp],
    "gf_last_error": [],
    "gf_version": [],
    "gf_abi_layout": [C.c_int, C.POINTER(c_i64), C.POINTER(c_i64)],
    "gf_launch_count": [],
}

_lib = None


class GoldfishError(RuntimeError):
    pass


def load():
    """Load the CUDA library; raises (never falls back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GoldfishError(
            "goldfish_b200 CUDA library not built: %s is missing "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_char_p if name == "gf_last_error" else (C.c_longlong if name == "gf_launch_count" else C.c_int)
    check_abi(lib)                      # the ctypes mirrors above must match the compiled header
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().gf_last_error()
        msg = msg.decode() if msg else ""
        if rc == 3:
            raise GoldfishNotConverged("%s: %s" % (what, msg))
        raise GoldfishError("%s failed (%s): %s" % (what, GF_ERRORS.get(rc, rc), msg))


class GoldfishNotConverged(GoldfishError):
    pass
