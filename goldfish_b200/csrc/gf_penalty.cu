// Penalty coupling along intersection curves (sm_100a, FP64).
//
// Replaces PENGoLINS transfer_penalty_residual / transfer_penalty_residual_deriv
// and GOLDFISH transfer_dRmdcpm_sub (/root/reference/GOLDFISH/nonmatching_opt.py:
// 745-752, 789-801, 864-867; /root/reference/GOLDFISH/utils/opt_utils.py:212-260).
// With vertex quadrature every mortar-space matrix is block diagonal per
// (cell, end vertex) evaluation, so the coupling term is  sum_q B_q^T H_q B_q
// with an 18 x 18 point Hessian.
//
//  1. k_penalty_points: one warp per evaluation; lane d carries the dual
//     direction d -> column d of H_uu (pass 0) or H_uX (pass 1).
//  2. gather kernels: one thread per DESTINATION (node, or node pair) sums the
//     contributions of the evaluations that touch it in a fixed, host-built
//     order -> deterministic, no atomics, K.vals updated with a plain +=.
#include "gf_common.cuh"

namespace gf {

__global__ void __launch_bounds__(128)
k_penalty_points(GfModel M, GfPenalty Q, int with_X) {
  __shared__ double sh[4][36];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ev = (int64_t)blockIdx.x * 4 + warp;
  if (ev >= Q.n_eval) return;
  double* g = sh[warp];
  // ---- mortar values: 18 u-variables, 18 X-variables ----
  if (lane < 18) {
    const int side = lane / 9, kc = lane % 9, k = kc / 3, c = kc % 3;
    const int32_t* conn = (side ? Q.connB : Q.connA) + ev * 16;
    const double* bas = (side ? Q.basB : Q.basA) + ev * 48 + k * 16;
    const int32_t* dd = (side ? Q.dofB : Q.dofA) + ev * 3;
    const double* u = M.u + dd[0] + (size_t)c * dd[1] - dd[2];
    double s = 0.0;
    for (int a = 0; a < 16; ++a) s = fma(bas[a], u[conn[a]], s);
    g[lane] = s;
    // X variable `lane`
    const int blk = lane / 3;  // 0: X_A(c) 1: X_A(c+1) 2: X_A,1 3: X_A,2 4: X_B,1 5: X_B,2
    const int32_t* cx; const double* bx;
    if (blk == 0) { cx = Q.connC0 + ev * 16; bx = Q.basC0 + ev * 16; }
    else if (blk == 1) { cx = Q.connC1 + ev * 16; bx = Q.basC1 + ev * 16; }
    else if (blk < 4) { cx = Q.connA + ev * 16; bx = Q.basA + ev * 48 + (blk - 1) * 16; }
    else { cx = Q.connB + ev * 16; bx = Q.basB + ev * 48 + (blk - 3) * 16; }
    double x = 0.0;
    for (int a = 0; a < 16; ++a) x = fma(bx[a], M.cp[(size_t)cx[a] * 4 + (lane % 3)], x);
    g[18 + lane] = x;
  }
  __syncwarp();
  const double tp[2] = {Q.tpar[ev * 2], Q.tpar[ev * 2 + 1]};
  const double ad = Q.alpha[ev * 2], ar = Q.alpha[ev * 2 + 1];
  for (int pass = 0; pass < (with_X ? 2 : 1); ++pass) {
    Dual uv[18], Xv[18], grad[18], e;
#pragma unroll
    for (int k = 0; k < 18; ++k) {
      uv[k] = Dual(g[k], (pass == 0 && lane == k) ? 1.0 : 0.0);
      Xv[k] = Dual(g[18 + k], (pass == 1 && lane == k) ? 1.0 : 0.0);
    }
    penalty_point<Dual>(uv, Xv, tp, ad, ar, e, grad);
    if (lane < 18) {
      double* H = (pass == 0 ? Q.Huu : Q.HuX) + ev * 324;
#pragma unroll
      for (int m = 0; m < 18; ++m) H[m * 18 + lane] = grad[m].d;
    }
    if (pass == 0 && lane == 0) {
#pragma unroll
      for (int m = 0; m < 18; ++m) Q.g[ev * 18 + m] = grad[m].v;
    }
    __syncwarp();
  }
}

// R[node, i] += sum_items sum_k bas[k][a] g[side*9 + 3k + i]
__global__ void k_penalty_gather_R(GfPenalty Q, double* R) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Q.nR) return;
  double s[3] = {0.0, 0.0, 0.0};
  for (int64_t it = Q.R_ptr[n]; it < Q.R_ptr[n + 1]; ++it) {
    const int32_t item = Q.R_item[it];
    const int64_t ev = item >> 5;
    const int ln = item & 31, side = ln >> 4, a = ln & 15;
    const double* bas = (side ? Q.basB : Q.basA) + ev * 48;
    const double* g = Q.g + ev * 18 + side * 9;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double b = bas[k * 16 + a];
#pragma unroll
      for (int i = 0; i < 3; ++i) s[i] = fma(b, g[3 * k + i], s[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) R[Q.R_row[n * 3 + i]] += s[i];
}

// K[(node r, i), (node c, j)] += sum_items sum_kl basR[k][a] Huu[sa*9+3k+i][sb*9+3l+j] basC[l][b]
__global__ void k_penalty_gather_K(GfPenalty Q, double* Kvals) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Q.nK) return;
  double s[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) s[i] = 0.0;
  for (int64_t it = Q.K_ptr[n]; it < Q.K_ptr[n + 1]; ++it) {
    const int32_t item = Q.K_item[it];
    const int64_t ev = item >> 10;
    const int la = (item >> 5) & 31, lb = item & 31;
    const int sa = la >> 4, a = la & 15, sb = lb >> 4, b = lb & 15;
    const double* bR = (sa ? Q.basB : Q.basA) + ev * 48;
    const double* bC = (sb ? Q.basB : Q.basA) + ev * 48;
    const double* H = Q.Huu + ev * 324 + (sa * 9) * 18 + sb * 9;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double br = bR[k * 16 + a];
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        const double w = br * bC[l * 16 + b];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) s[i * 3 + j] = fma(w, H[(3 * k + i) * 18 + 3 * l + j], s[i * 3 + j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int64_t pos = Q.K_pos[n * 9 + i];
    if (pos >= 0) Kvals[pos] += s[i];
  }
}

__global__ void k_penalty_gather_P(GfPenalty Q, GfPenaltyP PP) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= PP.n_dest) return;
  double s[3] = {0.0, 0.0, 0.0};
  for (int64_t it = PP.ptr[n]; it < PP.ptr[n + 1]; ++it) {
    const int64_t ev = PP.item_eval[it];
    const int code = PP.item_code[it];
    const int la = code & 31, xb = (code >> 5) & 7, lb = (code >> 8) & 15;
    const int sa = la >> 4, a = la & 15;
    const double* bR = (sa ? Q.basB : Q.basA) + ev * 48;
    double cC;
    if (xb == 0) cC = Q.basC0[ev * 16 + lb];
    else if (xb == 1) cC = Q.basC1[ev * 16 + lb];
    else if (xb < 4) cC = Q.basA[ev * 48 + (xb - 1) * 16 + lb];
    else cC = Q.basB[ev * 48 + (xb - 3) * 16 + lb];
    const double* H = Q.HuX + ev * 324 + (sa * 9) * 18 + xb * 3 + PP.field;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double w = bR[k * 16 + a] * cC;
#pragma unroll
      for (int i = 0; i < 3; ++i) s[i] = fma(w, H[(3 * k + i) * 18], s[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int64_t pos = PP.pos[n * 3 + i];
    if (pos >= 0) PP.vals[pos] += s[i];     // rounds of interfaces accumulate; the caller zeroes vals first
  }
}

__global__ void k_mask_vec(GfModel M, double* v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M.n_bc; i += (int64_t)gridDim.x * blockDim.x)
    v[M.bc_list[i]] = 0.0;
}

}  // namespace gf

using namespace gf;

extern "C" int gf_penalty_points(const GfModel* m, const GfPenalty* p, int with_X, void* stream) {
  if (!m || !p) return set_error(GF_ERR_BADARG, "gf_penalty_points: null argument");
  if (p->n_eval <= 0) return GF_OK;
  k_penalty_points<<<(unsigned)((p->n_eval + 3) / 4), 128, 0, (cudaStream_t)stream>>>(*m, *p, with_X);
  return check_launch("k_penalty_points");
}
extern "C" int gf_penalty_gather_R(const GfModel* m, const GfPenalty* p, double* R, void* stream) {
  (void)m;
  if (p->nR <= 0) return GF_OK;
  k_penalty_gather_R<<<(unsigned)((p->nR + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*p, R);
  return check_launch("k_penalty_gather_R");
}
extern "C" int gf_penalty_gather_K(const GfModel* m, const GfPenalty* p, void* stream) {
  if (p->nK <= 0) return GF_OK;
  k_penalty_gather_K<<<(unsigned)((p->nK + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*p, m->K.vals);
  return check_launch("k_penalty_gather_K");
}
extern "C" int gf_penalty_gather_P(const GfPenalty* p, const GfPenaltyP* pp, void* stream) {
  if (pp->n_dest <= 0) return GF_OK;
  k_penalty_gather_P<<<(unsigned)((pp->n_dest + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*p, *pp);
  return check_launch("k_penalty_gather_P");
}
extern "C" int gf_mask_vec(const GfModel* m, double* v, void* stream) {
  if (m->n_bc <= 0) return GF_OK;
  k_mask_vec<<<(unsigned)((m->n_bc + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*m, v);
  return check_launch("k_mask_vec");
}
