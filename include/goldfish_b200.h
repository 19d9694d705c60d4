/* goldfish_b200 -- C ABI of the B200-native analysis + adjoint hot path.
 *
 * Drop-in boundary for the arithmetic that the reference reaches through
 * dolfin/FFC/PETSc/MUMPS from GOLDFISH's NonMatchingOpt.
 * Plain pointers and sizes only; every pointer inside GfModel is a DEVICE
 * pointer unless its name ends in _h.  The host language on top is Python
 * (the reference is Python): goldfish_b200/_capi.py binds these with ctypes.
 *
 * Each entry point cites the reference interface it replaces
 * (paths relative to /root/reference/GOLDFISH/).
 *
 * Return value: 0 = ok, GF_ERR_* otherwise (gf_last_error() gives the text).
 */
#ifndef GOLDFISH_B200_H
#define GOLDFISH_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GF_OK 0
#define GF_ERR_BADARG 1
#define GF_ERR_CUDA 2
#define GF_ERR_NOCONV 3   /* Krylov / Newton did not reach the tolerance */
#define GF_ERR_BREAKDOWN 4
#define GF_ERR_NAN 5

/* thickness discretisation of a patch */
#define GF_TH_CONST 0  /* one dof per patch (Function(V_control)/Constant + HthMapComp) */
#define GF_TH_LINEAR 1 /* CG1 on the two-triangle mesh of knot-span vertices (V_linear) */
#define GF_TH_IGA 2    /* spline dofs, t = sum_a N_a theta_a (var_thickness=True) */

typedef struct GfPatchDesc {
  int32_t n_u, n_v;             /* control points per direction                   */
  int32_t neu, nev;             /* non-empty knot spans (Bezier elements) per dir */
  int32_t cp_off;               /* first scalar CP in the global scalar numbering */
  int32_t dof_off;              /* first displacement dof (field-blocked per patch) */
  int32_t th_off, th_kind, nth; /* thickness dofs                                 */
  int32_t span_u_off, span_v_off; /* offsets into per-span tables                 */
  int32_t cpd_u_off, cpd_v_off; /* offsets into per-direction CP tables           */
  int32_t rational;             /* any weight != 1                                */
  int32_t pcol_off[3];          /* column offset of the patch in dR/dCP_f, or -1  */
  int32_t el_off;               /* first element of the patch (global element ids)   */
  double E, nu;
  double f[3];                  /* dead load per unit reference area              */
} GfPatchDesc;

typedef struct GfCsr {          /* CSR with 64-bit row pointers, 32-bit columns   */
  int64_t nrows, ncols, nnz;
  const int64_t* indptr;
  const int32_t* indices;
  double* vals;
} GfCsr;

typedef struct GfModel {
  /* ---- sizes ---- */
  int32_t num_patches, num_elements, nq, num_colors;
  int64_t N;                    /* displacement dofs                               */
  int64_t n_scalar;             /* scalar control points                           */
  int64_t n_th;                 /* thickness dofs                                  */
  /* ---- patches / elements ---- */
  const GfPatchDesc* patches;   /* [num_patches]                                   */
  const int32_t* elem_patch;    /* [num_elements]                                  */
  const int32_t* elem_eu;       /* [num_elements] span index in u (0..neu-1)       */
  const int32_t* elem_ev;
  const int32_t* color_elem;    /* element ids grouped by colour                   */
  const int32_t* color_ptr_h;   /* HOST [num_colors+1]                             */
  /* ---- 1-D basis tables at the element quadrature points ---- */
  const double* tab_u;          /* [spans_u][nq][3][4] value, d1, d2 of the 4 fns  */
  const double* tab_v;
  const int32_t* first_cp_u;    /* [spans_u] first CP index of the span            */
  const int32_t* first_cp_v;
  const double* span_h_u;       /* [spans_u] span length                           */
  const double* span_h_v;
  const double* qw;             /* [nq] reference weights (sum 1)                  */
  const double* tw_lin;         /* [nq][4] barycentric weights of (v00,v10,v01,v11) */
  /* per-direction CP tables: stencil and element ranges of a CP index */
  const int32_t* cp_lo_u; const int32_t* cp_hi_u; const int32_t* el_lo_u; const int32_t* el_hi_u;
  const int32_t* cp_lo_v; const int32_t* cp_hi_v; const int32_t* el_lo_v; const int32_t* el_hi_v;
  /* ---- state ---- */
  const double* cp;             /* [n_scalar][4] homogeneous control points        */
  const double* u;              /* [N] displacement (IGA dofs)                     */
  const double* theta;          /* [n_th] thickness dofs                           */
  const uint8_t* bc;            /* [N] 1 = zero-dof                                */
  const int32_t* bc_list;       /* [n_bc] the zero-dofs                            */
  int64_t n_bc;
  const int32_t* row_nlow;      /* [n_scalar] coupling columns left of the own block (already x3) */
  /* ---- operators ---- */
  GfCsr K;                      /* dR/du, merged shell + penalty pattern           */
  GfCsr P[3];                   /* shell part of dR/dCP_f (structured)             */
  GfCsr T;                      /* dR/dthickness                                   */
} GfModel;

/* What one assembly pass produces (bit mask). */
#define GF_OUT_R 1      /* residual (shell part incl. dead loads)   -> out_R    */
#define GF_OUT_K 2      /* tangent                                   -> K.vals   */
#define GF_OUT_W 4      /* W_int and V, per element                  -> out_WV   */
#define GF_OUT_P 8      /* dR/dCP_f, dW/dCP_f, dV/dCP_f              -> P[f], out_dWdP, out_dVdP */
#define GF_OUT_T 16     /* dR/dt, dW/dt, dV/dt, dW/du                -> T, out_dWdt, ... */

typedef struct GfShellOut {
  double* R;        /* [N]   accumulated (+=); caller zeroes                        */
  double* WV;       /* [num_elements][2] per-element W_int, V (overwritten)         */
  double* dWdu;     /* [N]   (+=)                                                   */
  double* dWdP[3];  /* [ncols of P[f]] (+=)                                         */
  double* dVdP[3];
  double* dWdt;     /* [n_th] (+=)                                                  */
  double* dVdt;
  double* dt_el;    /* [num_elements][2] per-element dW/dt, dV/dt of GF_TH_CONST patches */
} GfShellOut;

/* Shell quadrature + scatter.  Replaces assemble(residuals[s]) /
 * assemble(residuals_deriv[s]) / assemble(dR_dcp_symexp) / assemble(dR_dh_th_symexp)
 * and the M^T . M extraction products:
 * nonmatching_opt.py:733-739, :779-781, :852, :936, :688 (AT_R_B), and the
 * functionals of operations/int_energy_exop.py:55-107, operations/volume_exop.py:46-84. */
int gf_shell_assemble(const GfModel* m, int what, const GfShellOut* out, void* stream);

/* Zero-dof handling of K after all contributions are in: rows/cols were
 * masked during scatter; this writes diag = 1 (nonmatching_opt.py:693-700). */
int gf_bc_set_diag(const GfModel* m, double diag, void* stream);

/* ---- penalty coupling (nonmatching_opt.py:745-752, :789-801, :864-867;
 *      utils/opt_utils.py:212-260) ---- */
typedef struct GfPenalty {
  int64_t n_eval;                 /* (cell, end vertex) evaluations                 */
  const int32_t* connA;           /* [n_eval][16] global scalar CPs, side A @ v     */
  const int32_t* connB;           /* [n_eval][16] side B @ v                        */
  const int32_t* connC0;          /* [n_eval][16] side A @ cell vertex c            */
  const int32_t* connC1;          /* [n_eval][16] side A @ cell vertex c+1          */
  const double* basA;             /* [n_eval][3][16] phi, phi_u, phi_v  @ v         */
  const double* basB;
  const double* basC0;            /* [n_eval][16] phi @ c                           */
  const double* basC1;
  const double* tpar;             /* [n_eval][2]                                    */
  const double* alpha;            /* [n_eval][2] alpha_d, alpha_r                   */
  const int32_t* dofA;            /* [n_eval][3] dof_off, ncp, cp_off of patch A    */
  const int32_t* dofB;
  /* point results */
  double* g;                      /* [n_eval][18]                                   */
  double* Huu;                    /* [n_eval][18][18]                               */
  double* HuX;                    /* [n_eval][18][18]  d grad_u / d Xv              */
  /* deterministic gather lists (host-built) */
  int64_t nR; const int64_t* R_ptr; const int32_t* R_item;   /* item = eval*32 + local node */
  const int32_t* R_row;           /* [nR][3] destination dofs                       */
  int64_t nK; const int64_t* K_ptr; const int32_t* K_item;   /* item = eval*1024 + la*32 + lb */
  const int64_t* K_pos;           /* [nK][9] positions in K.vals, -1 = masked (bc)  */
} GfPenalty;

int gf_penalty_points(const GfModel* m, const GfPenalty* p, int with_X, void* stream);
int gf_penalty_gather_R(const GfModel* m, const GfPenalty* p, double* R, void* stream);
int gf_penalty_gather_K(const GfModel* m, const GfPenalty* p, void* stream);

/* Penalty part of dR/dCP_f, kept as its own small CSR (its pattern reaches
 * one element beyond the shell stencil through the chord-length term).
 * One thread per destination (row node, column CP): 3 values (row fields), ACCUMULATED into vals: the host
 * groups the intersections into rounds with disjoint destination sets and calls this once per round
 * (fixed order => deterministic); the caller zeroes vals before the first round.
 * item code: la (5 bits: side*16 + a) | xblock << 5 (3 bits) | lb << 8 (4 bits);
 * xblock: 0 X_A(c), 1 X_A(c+1), 2 X_A,1, 3 X_A,2, 4 X_B,1, 5 X_B,2. */
typedef struct GfPenaltyP {
  int64_t n_dest; const int64_t* ptr;
  const int32_t* item_eval; const int32_t* item_code;
  const int64_t* pos;             /* [n_dest][3] positions in vals, -1 = masked     */
  double* vals;
  int32_t field;
} GfPenaltyP;
int gf_penalty_gather_P(const GfPenalty* p, const GfPenaltyP* pp, void* stream);
/* zero the listed entries of a vector (apply_bcs_vec, nonmatching_opt.py:655) */
int gf_mask_vec(const GfModel* m, double* v, void* stream);

/* ---- linear algebra (PETSc MatMult/MatMultTranspose, KSP; utils/opt_utils.py:106-209,
 *      operations/disp_imop.py:58-142) ---- */
int gf_spmv(const GfCsr* A, const double* x, double* y, double alpha, double beta, void* stream);
/* y = beta*y + alpha*A^T x through a host-built transpose map (deterministic gather) */
/* EXPERIMENTAL (round-2 candidate for the default tangent product): node-wise product for
 * matrices whose rows r0, r0+stride, r0+2*stride (the three displacement fields of one control point) share one
 * column list; node_row0[n] = first row of node n, node_stride[n] = its patch's control-point count. */
int gf_spmv_node(const GfCsr* A, const int64_t* node_row0, const int32_t* node_stride, int64_t n_nodes,
                 const double* x, double* y, double alpha, double beta, void* stream);
typedef struct GfCsrT { int64_t nrows, nnz; const int64_t* indptr; const int32_t* indices; const int64_t* perm; } GfCsrT;
int gf_spmv_t(const GfCsr* A, const GfCsrT* At, const double* x, double* y, double alpha, double beta, void* stream);

/* Overlapping additive Schwarz with banded block-Cholesky sub-solves (the CG +
 * per-patch-LU fieldsplit option of PENGoLINS, SURVEY.md Appendix A.5, made
 * overlapping so the penalty springs sit inside a block).  goldfish_b200/schwarz.py
 * builds the block orderings; all arrays are device pointers except *_h. */
typedef struct GfSchwarz {
  int32_t nblocks, nb;            /* nb = 64                                          */
  int32_t max_nbr, max_mb, max_n_pad;
  int32_t debug_flags;            /* 0 in production; timing experiments: 1 skip block GEMVs, 2 skip barriers;
                                     sweep kernel choice: 4 = one CTA per block, 8 = CTA group per block,
                                     16 = no thread-block cluster for the coarse block (default: single CTA + cluster when they fit);
                                     32 = timing experiments: skip the fine sweeps */
  int64_t n_y, band_len;          /* total padded local dofs; band storage length     */
  const int32_t* n_pad;           /* [nblocks] padded local size (multiple of nb)     */
  const int32_t* nbr;             /* [nblocks] block rows                             */
  const int64_t* off_j;           /* [nblocks] offsets into the per-block-column arrays */
  const int32_t* mbj;             /* [sum nbr] panel height (blocks below the diagonal) of each block column */
  const int32_t* rlen;            /* [sum nbr] blocks left of the diagonal in each block row */
  const int64_t* off_col;         /* [sum nbr] offset of each block column's panel in band */
  const int32_t* step_mb_h;       /* HOST [max_nbr] max panel height per step over the blocks */
  const int64_t* off_y;           /* [nblocks] offsets into y / glob                  */
  const int64_t* off_inv;         /* [nblocks] offsets into invd                      */
  const int32_t* glob;            /* [n_y] local -> global dof, -1 = padding          */
  const int32_t* gs;              /* per block: sorted global dofs ...                */
  const int32_t* ls;              /* ... and their local index (global -> local lookup) */
  const int64_t* off_g;           /* [nblocks+1] offsets into gs / ls                 */
  const int64_t* zptr;            /* [N+1] prolongation gather                        */
  const int64_t* zsrc;            /* [..] indices into y                              */
  double* band;                   /* factor storage: nb x nb blocks                   */
  float* band32;                  /* FP32 copy of the solve-form panels (streamed by the sweeps) */
  double* invd;                   /* inverses of the diagonal factor blocks           */
  double* y;                      /* [n_y] block-local vectors                        */
  double* s;                      /* [n_y] backward-sweep accumulators                */
  uint32_t* barrier;              /* [nblocks] group barrier counters                 */
  int32_t* flag;                  /* device flag: non-SPD block met                   */
} GfSchwarz;
int gf_schwarz_factor(const GfSchwarz* s, const GfCsr* K, void* stream);
int gf_schwarz_apply(const GfSchwarz* s, const double* r, double* z, int64_t n, void* stream);
/* fine blocks + one coarse block, triangular sweeps of both in one cooperative launch.  With at least
 * half as many fine blocks as SMs every fine block is swept by ONE CTA out of shared memory (r_f is
 * gathered by the kernel); otherwise a group of CTAs shares each block through a global-memory barrier. */
int gf_schwarz_apply2(const GfSchwarz* fine, const double* r_f, double* z_f, int64_t n_f,
                      const GfSchwarz* coarse, const double* r_c, double* z_c, int64_t n_c, void* stream);
/* the fine triangular sweeps alone (k_sw_solve1, one CTA per block); result stays in s->y (bench.py roofline) */
int gf_schwarz_sweeps(const GfSchwarz* s, const double* r, void* stream);
int gf_dot_slot0(int64_t n, const double* x, const double* y, double* partial2, int grid, void* stream);

/* Two-level preconditioner  z = sum_i R_i^T A_i^-1 R_i r  +  P Kc^-1 P^T r :
 * `fine` = overlapping patch blocks, `coarse` = ONE block holding the factor of
 * the coarse-spline operator (same shells + coupling re-discretised on a coarser
 * knot vector by the same kernels), P = knot-insertion prolongation, Rt = P^T. */
/* Patch-sharded multi-GPU runs (one process per GPU): vectors are replicated, matrix
 * rows / Schwarz blocks are owned.  The library computes the owned part of A p and of the
 * preconditioned residual and sums them over the ranks itself: NCCL all-reduce over NVLink, issued on the
 * compute stream from inside the Krylov loop (no host callback).  The host only carries the 128-byte unique
 * id from rank 0 to the others over its own process group (torch.distributed broadcast). */
typedef struct GfDist {
  int32_t n_ranges;               /* 0 = single process                               */
  int32_t rank, world, pad_;
  const int64_t* ranges_h;        /* HOST [n_ranges][2] owned row ranges [begin, end)  */
  void* comm;                     /* ncclComm_t, set by gf_dist_init                   */
} GfDist;
int gf_dist_unique_id(void* id128);                                     /* rank 0: ncclGetUniqueId            */
int gf_dist_init(GfDist* d, const void* id128, int rank, int world);    /* collective: ncclCommInitRank       */
int gf_dist_allreduce(const GfDist* d, double* buf, int64_t n, void* stream);   /* in-place FP64 sum          */
int gf_dist_destroy(GfDist* d);

typedef struct GfPrecond {
  const GfSchwarz* fine;
  const GfSchwarz* coarse;        /* may be NULL: one-level                           */
  GfCsr P, Rt;
  double* rc; double* zc;         /* [Nc] work vectors                                */
  const int32_t* bc_c;            /* coarse zero-dofs                                 */
  int64_t n_bc_c;
  const GfDist* dist;             /* NULL or n_ranges == 0: single process            */
  /* Sharded runs: the coarse solve as a dense product.  Every rank holds a ROW SLAB of the (symmetric) inverse of
   * the coarse operator, FP64 row-major [cinv_rows][Nc]; it computes z_c[slab] = Kc^-1[slab, :] r_c, prolongates
   * only that part, and the all-reduce that already sums the fine-block contributions completes z.  A 2 * n_blockrows
   * latency chain (k_sw_coarse_cluster, replicated on every GPU) becomes an HBM-bound product split over the GPUs
   * with no extra exchange step.  NULL: band factor `coarse` (single-GPU path). */
  const double* cinv;
  int64_t cinv_row0, cinv_rows;
} GfPrecond;
int gf_precond_apply(const GfPrecond* pc, const double* r, double* z, int64_t n, void* stream);

/* Node-wise row structure of the tangent: rows row0[n], row0[n]+stride[n], row0[n]+2 stride[n] (the three
 * displacement fields of control point n) share one column list, so the Krylov product reads each column index
 * and each x entry once per three non-zeros (9.33 B per non-zero instead of 12; bitwise the same y as gf_spmv).
 * Sharded runs list the control points of the owned patches only.  n = 0: plain row-wise product. */
typedef struct GfNodeRows { const int64_t* row0; const int32_t* stride; int64_t n; } GfNodeRows;

typedef struct GfPcgWork {
  double* r; double* z; double* p; double* Ap;  /* [n] each */
  double* dinv;       /* [n] inverse diagonal (Jacobi) or 3x3 blocks, see precond */
  double* scal;       /* [16] device scalars                                        */
  double* partial;    /* [4096*4] per-CTA partial sums                              */
  double* scal_h;     /* pinned host [16]                                           */
  GfNodeRows nodes;
} GfPcgWork;
/* Preconditioned CG on K x = b (K symmetric: nonmatching_opt.py:804-809).
 * Replaces solve_nonmatching_mat(..., 'direct') (utils/opt_utils.py:176,204). */
int gf_pcg(const GfCsr* A, const double* b, double* x, const GfPcgWork* w, const GfPrecond* precond,
           const GfDist* dist, double rtol, double atol, int max_it, int check_every, int* iters,
           double* relres, void* stream);
int gf_jacobi_setup(const GfCsr* A, double* dinv, void* stream);
/* r = b - A x with exact products (FMA) and double-double accumulation: the residual of iterative refinement
 * (the counterpart of the extended-precision residual the LU goldens are refined with).  Sharded: owned rows, summed. */
int gf_residual_dd(const GfCsr* A, const GfDist* dist, const double* x, const double* b, double* r, void* stream);

/* Flexible (right-preconditioned) restarted GMRES(restart) on K x = b with the same preconditioner: the fallback when
 * the tangent is indefinite and CG reports GF_ERR_BREAKDOWN (the reference's LU still returns a Newton step
 * there, utils/opt_utils.py:176).  `iters` counts matrix-vector products. */
typedef struct GfGmresWork {
  double* V;          /* [(restart+1)][n] Krylov basis                              */
  double* Z;          /* [restart][n] preconditioned basis z_j = M^-1 v_j (flexible GMRES) */
  double* t;          /* [n]                                                        */
  double* hdev;       /* [2*(restart+2)+1] device scalars                           */
  double* partial;    /* [1024*(restart+1)] per-CTA partial sums                    */
  double* h_host;     /* pinned host [2*(restart+2)+1]                              */
  GfNodeRows nodes;
} GfGmresWork;
int gf_gmres(const GfCsr* A, const double* b, double* x, const GfGmresWork* w, const GfPrecond* precond,
             const GfDist* dist, double rtol, int restart, int max_it, int* iters, double* relres, void* stream);

/* small vector helpers on device */
int gf_axpby(int64_t n, double a, const double* x, double b, double* y, void* stream);
int gf_dot(int64_t n, const double* x, const double* y, double* partial, double* out_dev, void* stream);
int gf_reduce_wv(int64_t num_elements, const double* WV, double* out2_dev, void* stream);

/* FP64 peak micro-benchmarks (mode 0: DFMA chains, 1: DMMA m8n8k4): one launch of `grid` x 256 threads; the
 * caller times it with CUDA events; *flops receives the operation count.  Measurement infrastructure. */
int gf_peak_fp64(int mode, int grid, int iters, double* out, double* flops, void* stream);

const char* gf_last_error(void);
int gf_version(void);
/* sizeof / offsetof(last field) of the public structs, in declaration order (0 GfPatchDesc, 1 GfCsr, 2 GfModel,
 * 3 GfShellOut, 4 GfPenalty, 5 GfPenaltyP, 6 GfCsrT, 7 GfSchwarz, 8 GfDist, 9 GfPrecond, 10 GfPcgWork, 11 GfGmresWork): lets a
 * binding verify its mirror of this header (tests/test_capi_symbols.py does, without a GPU). */
int gf_abi_layout(int which, int64_t* size, int64_t* last_offset);
/* number of kernels this library has launched so far (bench.py's gpu_launches) */
long long gf_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
